// stats.cuh - K1: fp64 prefix sums of x and x x^T over the frames of a recording.
//
// Replaces every np.cov / np.mean over a slice (spk-change-detection.py:87-94,
// 107-108, 126-127): after this pass the sufficient statistics of ANY window
// [a, b) are P[b] - P[a], an O(d^2) difference, instead of an O(n d^2) pass
// over the raw frames.
//
// HBM layout: P is (n + 1) records of REC = 820 doubles (layout.cuh), record t
// = statistics of frames [0, t).  Frames are centred on a per-file shift vector
// first (covariances are shift invariant; centring keeps the prefix magnitudes,
// hence the cancellation in P[b] - P[a], small).
//
// Three launches, all bit-reproducible run to run (a look-back scan would add
// tile aggregates in a timing-dependent association, and change points must be
// bit-identical between runs):
//   A  tile_sums   one CTA per TILE frames: the 820 sums of its tile
//   B  tile_scan   exclusive scan of the tile sums, one thread per component
//   C  tile_write  one CTA per tile: running sums from its base, written per frame
// Algorithmic traffic per frame: 2 x 156 B read + 6,560 B written (HBM-bound).
#pragma once

#include "common.cuh"

namespace spk {

constexpr int K1_TILE = 128;          // frames per tile
constexpr int K1_THREADS = 832;       // 26 warps; threads 0..819 own one component each
constexpr int K1_XS = 40;             // smem row stride: D values + a constant 1

// component q of a record is  sum_t xs[t][c_row[q]] * xs[t][c_col[q]]  where
// column D of xs is the constant 1 (first moments, frame count)
__constant__ uint8_t c_row[REC];
__constant__ uint8_t c_col[REC];

inline void fill_lut(uint8_t* row, uint8_t* col) {
    for (int r = 0; r < D39; ++r)
        for (int c = 0; c <= r; ++c) { row[L39::pos(r, c)] = (uint8_t)r; col[L39::pos(r, c)] = (uint8_t)c; }
    for (int j = 0; j < D39; ++j) { row[L39::VEC + j] = (uint8_t)j; col[L39::VEC + j] = (uint8_t)D39; }
    row[L39::CNT] = (uint8_t)D39; col[L39::CNT] = (uint8_t)D39;
}

// per-file shift: mean of up to ~2048 evenly spaced frames (any constant works;
// it only has to be near the mean and identical for every record of the file)
__global__ void __launch_bounds__(1024) k1_shift(const float* __restrict__ x, int64_t n, double* __restrict__ shift) {
    __shared__ double part[25][K1_XS];
    const int g = threadIdx.x / K1_XS, j = threadIdx.x % K1_XS;
    const int64_t step = n > 2048 ? n / 2048 : 1;
    const int64_t ns = (n + step - 1) / step;
    double acc = 0.0;
    if (g < 25 && j < D39)
        for (int64_t s = g; s < ns; s += 25) acc += (double)x[(s * step) * D39 + j];
    if (g < 25) part[g][j] = acc;
    __syncthreads();
    if (threadIdx.x < K1_XS) {
        double t = 0.0;
        for (int k = 0; k < 25; ++k) t += part[k][threadIdx.x];
        shift[threadIdx.x] = (threadIdx.x < D39 && ns > 0) ? t / (double)ns : 0.0;
    }
}

__device__ __forceinline__ void k1_load_tile(const float* __restrict__ x, int64_t n, int64_t f0,
                                             const double* __restrict__ shift, double (*xs)[K1_XS]) {
    const int64_t left = n - f0;
    const int valid = left < K1_TILE ? (int)left : K1_TILE;
    const float* src = x + f0 * D39;
    const int total = valid * D39;
    // the tile start is 16-byte aligned whenever the matrix base is (128*39*4 B per tile)
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        for (int i = threadIdx.x; i < total / 4; i += K1_THREADS) {
            const float4 v = __ldg(s4 + i);
            const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = 4 * i + u, t = idx / D39, j = idx - t * D39;
                xs[t][j] = (double)e[u] - shift[j];
            }
        }
        for (int idx = (total / 4) * 4 + threadIdx.x; idx < total; idx += K1_THREADS) {
            const int t = idx / D39, j = idx - t * D39;
            xs[t][j] = (double)__ldg(src + idx) - shift[j];
        }
    } else {
        for (int idx = threadIdx.x; idx < total; idx += K1_THREADS) {
            const int t = idx / D39, j = idx - t * D39;
            xs[t][j] = (double)__ldg(src + idx) - shift[j];
        }
    }
    for (int t = threadIdx.x; t < K1_TILE; t += K1_THREADS) xs[t][D39] = t < valid ? 1.0 : 0.0;
    // frames past the end of the recording contribute nothing
    for (int idx = total + threadIdx.x; idx < K1_TILE * D39; idx += K1_THREADS) {
        const int t = idx / D39, j = idx - t * D39;
        xs[t][j] = 0.0;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(K1_THREADS) k1_tile_sums(const float* __restrict__ x, int64_t n,
                                                           const double* __restrict__ shift,
                                                           double* __restrict__ tile) {
    __shared__ __align__(16) double xs[K1_TILE][K1_XS];
    const int64_t f0 = (int64_t)blockIdx.x * K1_TILE;
    k1_load_tile(x, n, f0, shift, xs);
    const int q = threadIdx.x;
    if (q >= REC) return;
    const int r = c_row[q], c = c_col[q];
    double a0 = 0.0, a1 = 0.0;
#pragma unroll 8
    for (int t = 0; t < K1_TILE; t += 2) {
        a0 = fma(xs[t][r], xs[t][c], a0);
        a1 = fma(xs[t + 1][r], xs[t + 1][c], a1);
    }
    tile[(int64_t)blockIdx.x * REC + q] = a0 + a1;
}

// exclusive scan over tiles, one thread per component, fixed left-to-right order
__global__ void __launch_bounds__(128) k1_tile_scan(double* __restrict__ tile, int64_t ntiles) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= REC) return;
    double run = 0.0;
    int64_t t = 0;
    for (; t + 8 <= ntiles; t += 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = tile[(t + u) * REC + q];
#pragma unroll
        for (int u = 0; u < 8; ++u) { tile[(t + u) * REC + q] = run; run += v[u]; }
    }
    for (; t < ntiles; ++t) { const double v = tile[t * REC + q]; tile[t * REC + q] = run; run += v; }
}

__global__ void __launch_bounds__(K1_THREADS) k1_tile_write(const float* __restrict__ x, int64_t n,
                                                            const double* __restrict__ shift,
                                                            const double* __restrict__ tile,
                                                            double* __restrict__ P) {
    __shared__ __align__(16) double xs[K1_TILE][K1_XS];
    const int64_t f0 = (int64_t)blockIdx.x * K1_TILE;
    k1_load_tile(x, n, f0, shift, xs);
    const int q = threadIdx.x;
    if (q >= REC) return;
    const int64_t left = n - f0;
    const int valid = left < K1_TILE ? (int)left : K1_TILE;
    const int r = c_row[q], c = c_col[q];
    double acc = tile[(int64_t)blockIdx.x * REC + q];
    double* out = P + (f0 + 1) * REC + q;
    if (blockIdx.x == 0) P[q] = 0.0;
#pragma unroll 4
    for (int t = 0; t < valid; ++t) {
        acc = fma(xs[t][r], xs[t][c], acc);
        __stcs(out + (int64_t)t * REC, acc);      // streaming store: written once, read later by other kernels
    }
}

}  // namespace spk
