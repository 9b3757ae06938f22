// Micro-benchmark: dependent-chain latencies that bound one LDL^T step on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double seed) {
    __shared__ double sm[64];
    const int lane = threadIdx.x & 31;
    double a = seed + lane * 1e-3, b = 1.0000001, c = 1e-9;
    long long t0, t1;
    // (0) dependent DFMA
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; ++i) a = fma(a, b, c);
    t1 = clock64(); if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // (1) dependent DMUL
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; ++i) a = a * b;
    t1 = clock64(); if (threadIdx.x == 0) cyc[1] = t1 - t0;
    // (2) dependent __drcp_rn
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) a = __drcp_rn(a) + 0.5;
    t1 = clock64(); if (threadIdx.x == 0) cyc[2] = t1 - t0;
    // (3) dependent 1.0 / a
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) a = 1.0 / a + 0.5;
    t1 = clock64(); if (threadIdx.x == 0) cyc[3] = t1 - t0;
    // (4) STS -> syncwarp -> LDS round trip (dependent)
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) { sm[lane] = a; __syncwarp(); a = sm[(lane + 1) & 31]; __syncwarp(); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[4] = t1 - t0;
    // (5) shuffle of a double (dependent)
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) a = __shfl_sync(0xffffffffu, a, (lane + 1) & 31);
    t1 = clock64(); if (threadIdx.x == 0) cyc[5] = t1 - t0;
    // (6) independent DFMA throughput: 8 chains
    double e[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) e[q] = a + q;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i)
#pragma unroll
        for (int q = 0; q < 8; ++q) e[q] = fma(e[q], b, c);
    t1 = clock64(); if (threadIdx.x == 0) cyc[6] = t1 - t0;
#pragma unroll
    for (int q = 0; q < 8; ++q) a += e[q];
    // (7) fast reciprocal: MUFU.RCP64H + 2 Newton steps, no slow path
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) {
        double y;
        asm volatile("{ .reg .b32 lo, hi; mov.b64 {lo, hi}, %1; rcp.approx.ftz.f64 %0, %1; }" : "=d"(y) : "d"(a));
        double r = fma(-a, y, 1.0); y = fma(y, r, y); r = fma(-a, y, 1.0); y = fma(y, r, y);
        a = y + 0.5;
    }
    t1 = clock64(); if (threadIdx.x == 0) cyc[7] = t1 - t0;
    // (8) log
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 16; ++i) a = log(a) + 3.0;
    t1 = clock64(); if (threadIdx.x == 0) cyc[8] = t1 - t0;
    out[threadIdx.x] = a;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 16 * 8);
    for (int nw : {1, 4, 12}) {
        k<<<1, 32 * nw>>>(out, cyc, 1.5); k<<<1, 32 * nw>>>(out, cyc, 1.5);
        cudaDeviceSynchronize();
        long long h[16]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%2d warps: DFMA dep %.1f  DMUL dep %.1f  drcp_rn(+add) %.1f  1/x(+add) %.1f  STS-sync-LDS-sync %.1f  shfl64 %.1f  DFMA x8 indep %.2f/instr  rcp.approx+2NR(+add) %.1f  log(+add) %.1f\n", nw,
               h[0] / 256.0, h[1] / 256.0, h[2] / 64.0, h[3] / 64.0, h[4] / 64.0, h[5] / 64.0, h[6] / 512.0, h[7] / 64.0, h[8] / 16.0);
    }
    return 0;
}
