// Micro-benchmark: one warp adds frame rows from shared memory in float32, lane = dimension
// (the KL2 sum chain), in three codings.  cycles per row.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int D = 39, ROWS = 64;
__global__ void k(const float* x, float* out, long long* cyc, int nst) {
    extern __shared__ float buf[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int e = threadIdx.x; e < ROWS * D; e += blockDim.x) buf[e] = x[e];
    __syncthreads();
    const bool second = lane + 32 < D;
    float s0 = 0.f, s1 = 0.f;
    const float* b = buf + lane;
    long long t0 = clock64();
    // (A) simple: 8 rows, loads then adds
    for (int st = 0; st < nst; ++st)
        for (int r = 0; r + 8 <= ROWS; r += 8) {
            float u[8], v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) { u[q] = b[(r + q) * D]; v[q] = second ? b[(r + q) * D + 32] : 0.f; }
#pragma unroll
            for (int q = 0; q < 8; ++q) { s0 = __fadd_rn(s0, u[q]); s1 = __fadd_rn(s1, v[q]); }
        }
    long long t1 = clock64();
    if (lane == 0) cyc[warp * 4 + 0] = t1 - t0;
    // (B) 16 rows loads then adds
    t0 = clock64();
    for (int st = 0; st < nst; ++st)
        for (int r = 0; r + 16 <= ROWS; r += 16) {
            float u[16], v[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) { u[q] = b[(r + q) * D]; v[q] = second ? b[(r + q) * D + 32] : 0.f; }
#pragma unroll
            for (int q = 0; q < 16; ++q) { s0 = __fadd_rn(s0, u[q]); s1 = __fadd_rn(s1, v[q]); }
        }
    t1 = clock64();
    if (lane == 0) cyc[warp * 4 + 1] = t1 - t0;
    // (C) only s0 (32 dims), 16 rows
    t0 = clock64();
    for (int st = 0; st < nst; ++st)
        for (int r = 0; r + 16 <= ROWS; r += 16) {
            float u[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) u[q] = b[(r + q) * D];
#pragma unroll
            for (int q = 0; q < 16; ++q) s0 = __fadd_rn(s0, u[q]);
        }
    t1 = clock64();
    if (lane == 0) cyc[warp * 4 + 2] = t1 - t0;
    // (D) dependent FADD latency
    t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < 1024; ++i) s0 = __fadd_rn(s0, s1);
    t1 = clock64();
    if (lane == 0) cyc[warp * 4 + 3] = t1 - t0;
    out[threadIdx.x] = s0 + s1;
}
int main() {
    float* x; float* out; long long* cyc;
    cudaMalloc(&x, ROWS * D * 4); cudaMemset(x, 0, ROWS * D * 4); cudaMalloc(&out, 4096); cudaMalloc(&cyc, 64 * 8);
    const int nst = 64;
    for (int nw : {1, 4}) {
        k<<<1, 32 * nw, ROWS * D * 4>>>(x, out, cyc, nst); k<<<1, 32 * nw, ROWS * D * 4>>>(x, out, cyc, nst);
        cudaDeviceSynchronize();
        long long h[4]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%d warps: cycles/row  (A) 8-row batches %.1f  (B) 16-row %.1f  (C) 32 dims only %.1f   FADD dependent %.2f\n", nw,
               h[0] / (double)(nst * ROWS), h[1] / (double)(nst * ROWS), h[2] / (double)(nst * ROWS), h[3] / 1024.0);
    }
    return 0;
}
