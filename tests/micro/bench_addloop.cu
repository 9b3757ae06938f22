// Micro-benchmark (not a test): the serial float32 row-add loop of the KL2 sum chains in isolation.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int D39 = 39;
__global__ void k(float* out, long long* cyc, int iters, int variant, const float* gsrc) {
    __shared__ __align__(16) float ring[32 * 40 + 64];
    const int lane = threadIdx.x & 31;
    for (int q = threadIdx.x; q < 32 * D39; q += blockDim.x) ring[q] = q * 1e-3f;
    __syncthreads();
    const int off2 = lane + 32 < D39 ? 32 : 0;
    float s0 = 0.f, s1 = 0.f, t0 = 0.f, t1 = 0.f;
    const long long c0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const float* buf = ring + lane;
        int row = 0; const int sseg = 32;
        if (variant == 0) {
            float ua[8], va[8], ub[8], vb[8];
            auto load8 = [&](float (&u)[8], float (&v)[8], int rr) {
                const float* p = buf + rr * D39;
#pragma unroll
                for (int q = 0; q < 8; ++q) { u[q] = p[q * D39]; v[q] = p[q * D39 + off2]; }
            };
            auto add8 = [&](const float (&u)[8], const float (&v)[8]) {
#pragma unroll
                for (int q = 0; q < 8; ++q) { s0 = __fadd_rn(s0, u[q]); s1 = __fadd_rn(s1, v[q]); }
            };
            load8(ua, va, row);
            for (;;) {
                bool more = row + 16 <= sseg;
                if (more) load8(ub, vb, row + 8);
                add8(ua, va);
                row += 8;
                if (!more) break;
                more = row + 16 <= sseg;
                if (more) load8(ua, va, row + 8);
                add8(ub, vb);
                row += 8;
                if (!more) break;
            }
        } else if (variant == 1) {
#pragma unroll 8
            for (; row < sseg; ++row) { s0 = __fadd_rn(s0, buf[row * D39]); s1 = __fadd_rn(s1, buf[row * D39 + off2]); }
        } else if (variant == 2) {      // one LDS per row
#pragma unroll 16
            for (; row < sseg; ++row) { s0 = __fadd_rn(s0, buf[row * D39]); }
        } else if (variant == 3) {      // one LDS per row, four chains
#pragma unroll 16
            for (; row < sseg; ++row) { const float u = buf[row * D39]; s0 = __fadd_rn(s0, u); s1 = __fadd_rn(s1, u); t0 = __fadd_rn(t0, u); t1 = __fadd_rn(t1, u); }
        } else if (variant == 4) {      // global loads, two per row
#pragma unroll 16
            for (; row < sseg; ++row) { s0 = __fadd_rn(s0, __ldg(gsrc + lane + row * D39)); s1 = __fadd_rn(s1, __ldg(gsrc + lane + row * D39 + off2)); }
        } else {                        // LDS.64: two consecutive floats per lane (20 lanes x 2 dims), one per row
#pragma unroll 16
            for (; row < sseg; ++row) { const float2 u = *reinterpret_cast<const float2*>(ring + 2 * lane + row * 40); s0 = __fadd_rn(s0, u.x); s1 = __fadd_rn(s1, u.y); }
        }
        asm volatile("" ::: "memory");
    }
    const long long c1 = clock64();
    out[threadIdx.x] = s0 + s1 + t0 + t1;
    if (threadIdx.x == 0) cyc[0] = c1 - c0;
}
int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 1024 * 4); cudaMalloc(&cyc, 8);
    float* gsrc; cudaMalloc(&gsrc, 64 * 1024); cudaMemset(gsrc, 0, 64 * 1024);
    for (int v : {0, 1, 2, 3, 4, 5}) for (int nw : {1, 4}) {
        k<<<1, 32 * nw>>>(out, cyc, 1000, v, gsrc); k<<<1, 32 * nw>>>(out, cyc, 1000, v, gsrc);
        cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("variant %d warps %d: %.2f cycles per row\n", v, nw, h / (1000.0 * 32));
    }
    return 0;
}
