// Micro-benchmark (not a test): shared-memory pipe cost of the access patterns of the warp LDL^T strip
// (broadcast reads of 4 / 8 distinct addresses, 64- and 128-bit; stores by the 4 owner lanes), 12 warps per SM.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o bench_smem bench_smem.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(384, 1) k(double* out, long long* cyc, int iters) {
    __shared__ __align__(16) double sm[12][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, i = lane >> 3, j = lane & 7;
    double* w = sm[warp];
    for (int q = lane; q < 256; q += 32) w[q] = q * 0.5 + warp;
    __syncwarp();
    unsigned base = (unsigned)__cvta_generic_to_shared(w);
    unsigned addr;
    if (MODE == 0) addr = base + lane * 8;                 // 64-bit, 32 distinct consecutive
    if (MODE == 1) addr = base + i * 8;                    // 64-bit, 4 distinct (rows: vn[i + 4 ri])
    if (MODE == 2) addr = base + j * 8;                    // 64-bit, 8 distinct (columns)
    if (MODE == 3) addr = base + i * 80;                   // 128-bit, 4 distinct chunks
    if (MODE == 4) addr = base + j * 48;                   // 128-bit, 8 distinct chunks
    if (MODE == 5) addr = base + lane * 16;                // 128-bit, 32 distinct consecutive
    if (MODE == 6) addr = base + i * 8;                    // store 64-bit by 4 lanes (j == 3)
    if (MODE == 7) addr = base + i * 80;                   // store 128-bit by 4 lanes
    if (MODE == 8) addr = base + lane * 8;                 // store 64-bit by all lanes
    if (MODE == 9) addr = base + i * 8;                    // store 64-bit by 8 lanes (i < 1?)  -> 8 active lanes (lane < 8)
    if (MODE == 10) addr = base + lane * 4;                // 32-bit shuffle stand-in
    double acc0 = lane * 0.25, acc1 = lane * 0.5;
    unsigned x[4] = {0, 0, 0, 0};
    __syncthreads();
    const long long c0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (MODE <= 2) {
                unsigned lo, hi;
                asm volatile("ld.shared.v2.u32 {%0, %1}, [%2+%3];" : "=r"(lo), "=r"(hi) : "r"(addr), "n"(0), "r"(it) : "memory");
                x[u & 3] += lo;
            } else if (MODE <= 5) {
                unsigned a0, a1, a2, a3;
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(addr), "r"(it) : "memory");
                x[u & 3] += a0 + a3;
            } else if (MODE == 6) {
                asm volatile("{\n.reg .pred q;\nsetp.eq.s32 q, %0, 3;\n@q st.shared.f64 [%1], %2;\n}" ::"r"(j), "r"(addr), "d"(acc0) : "memory");
            } else if (MODE == 7) {
                asm volatile("{\n.reg .pred q;\nsetp.eq.s32 q, %0, 3;\n@q st.shared.v2.f64 [%1], {%2, %3};\n}" ::"r"(j), "r"(addr), "d"(acc0), "d"(acc1) : "memory");
            } else if (MODE == 8) {
                asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(acc0), "r"(it) : "memory");
            } else if (MODE == 9) {
                asm volatile("{\n.reg .pred q;\nsetp.lt.s32 q, %0, 8;\n@q st.shared.f64 [%1], %2;\n}" ::"r"(lane), "r"(addr), "d"(acc0) : "memory");
            } else {
                x[u & 3] += __shfl_sync(0xffffffffu, x[(u + 1) & 3] + it, (lane + 5) & 31);
            }
        }
    }
    const long long c1 = clock64();
    if (lane == 0) cyc[warp] = c1 - c0;
    out[threadIdx.x] = acc0 + acc1 + (double)(x[0] ^ x[1] ^ x[2] ^ x[3]);
}

template <int MODE>
void run(const char* name, double* out, long long* cyc) {
    const int iters = 2000;
    for (int nw : {1, 12}) {
        k<MODE><<<1, nw * 32>>>(out, cyc, iters);
        k<MODE><<<1, nw * 32>>>(out, cyc, iters);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
        long long h[12];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%-44s %2d warps: %.2f cycles per instruction per warp, %.2f SM cycles per instruction\n", name, nw,
               h[0] / (double)(iters * 16), h[0] / (double)(iters * 16) / nw);
    }
}

int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 384 * 8); cudaMalloc(&cyc, 12 * 8);
    run<0>("LDS.64  32 distinct", out, cyc);
    run<1>("LDS.64  4 distinct (broadcast rows)", out, cyc);
    run<2>("LDS.64  8 distinct (broadcast columns)", out, cyc);
    run<3>("LDS.128 4 distinct chunks, stride 80 B", out, cyc);
    run<4>("LDS.128 8 distinct chunks, stride 48 B", out, cyc);
    run<5>("LDS.128 32 distinct", out, cyc);
    run<6>("STS.64  4 lanes active", out, cyc);
    run<7>("STS.128 4 lanes active", out, cyc);
    run<8>("STS.64  32 lanes", out, cyc);
    run<9>("STS.64  8 lanes active", out, cyc);
    run<10>("SHFL.32", out, cyc);
    return 0;
}
