// stats.cuh - K1: fp64 prefix sums of x and x x^T over the frames of a recording.
//
// Replaces every np.cov / np.mean over a slice (spk-change-detection.py:87-94,
// 107-108, 126-127): after this pass the sufficient statistics of ANY window
// [a, b) are an O(d^2) difference of records instead of an O(n d^2) pass over
// the raw frames.
//
// Accuracy is the design constraint.  The reference computes each window's
// covariance directly, so its entries are good to ~eps; a plain running prefix
// would give P[b] - P[a] an error of eps * (prefix magnitude), i.e. (t / n) eps
// relative - 7,000 eps for a half-second window at the end of an hour - and
// speakers with ill-conditioned covariances (cond 1e9 occurs in the synthetic
// data) amplify that into the 1e-5 range of ln|S|.  Hence two levels:
//
//   P[t]  (n + 1 records of REC doubles, lane-paired layout, layout.cuh):
//         statistics of frames [block_start(t), t) - the prefix RESTARTS at
//         every block of K1_TILE = 128 frames, so its magnitude is bounded by one
//         block;
//   C[j]  (nblocks + 1 records of REC double2): statistics of frames
//         [0, j * 128) as a DOUBLE-DOUBLE (hi, lo) pair, accumulated with
//         error-free additions.
//
//   sum over [a, b) = (P[b] - P[a]) + ((C[jb].hi - C[ja].hi) + (C[jb].lo - C[ja].lo))
//
// with ja, jb the blocks of a and b; the second bracket vanishes (and is not
// loaded) when both fall into one block.  Every window sum is then good to a
// few eps of ITS OWN magnitude wherever it lies in the recording.
//
// Frames are centred on a per-file shift vector first (covariances are shift
// invariant; centring keeps even the in-block cancellation small).
//
// Three launches, bit-reproducible run to run:
//   tile_write  one CTA per block: running sums written per frame + block total
//   chunk_sums / chunk_scan  double-double exclusive scan of the block totals
// Algorithmic traffic per frame: 156 B read + 6,560 B written (HBM-bound); the
// block level adds 13,120 B per 128 frames (+1.6 %).
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
__device__ __forceinline__ void k1_shift_body(const float* __restrict__ x, int64_t n, double* __restrict__ shift) {
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
static __global__ void __launch_bounds__(1024) k1_shift(const float* __restrict__ x, int64_t n, double* __restrict__ shift) {
    k1_shift_body(x, n, shift);
}

// ---- packed batches of recordings (spkdiar_features_upload_batch) -------------------------
// Recording r occupies the packed frame rows [base, base + n); base is a multiple of K1_TILE
// and the recording owns n / K1_TILE + 1 blocks (its last block is partial or empty), so that
// the record at its end - P[base + n], C[block of base + n] - never coincides with the zero
// record that starts the next recording.  Shift, block-local prefix and block-level scan all
// restart per recording: the statistics of a packed recording are bit-identical to those of
// the same recording uploaded alone.
struct RecTab { int64_t base; int64_t n; };

static __global__ void __launch_bounds__(1024) k1_shift_batch(const float* __restrict__ x, const RecTab* __restrict__ tab,
                                                       double* __restrict__ shift) {
    const RecTab t = tab[blockIdx.x];
    k1_shift_body(x + t.base * D39, t.n, shift + (int64_t)blockIdx.x * K1_XS);
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

// ---- exclusive double-double scan over the block totals ------------------------------
// C[j] = sum of the totals of blocks < j, per component, as (hi, lo).  The blocks are cut
// into K1_CHUNKS contiguous chunks; thread (chunk, component) first sums its chunk, then
// adds the totals of the chunks before it IN ORDER and rescans its chunk.  The association
// is fixed by the block count alone, so the result is bit-reproducible, and consecutive
// threads own consecutive components, so every access is coalesced.
constexpr int K1_CHUNKS = 64;

__device__ __forceinline__ void dd_add(double& hi, double& lo, double v) {       // (hi, lo) += v
    const double s = __dadd_rn(hi, v);                         // TwoSum(hi, v)
    const double bb = __dsub_rn(s, hi);
    const double err = __dadd_rn(__dsub_rn(hi, __dsub_rn(s, bb)), __dsub_rn(v, bb));
    lo = __dadd_rn(lo, err);
    const double h2 = __dadd_rn(s, lo);                        // renormalise (FastTwoSum)
    lo = __dsub_rn(lo, __dsub_rn(h2, s));
    hi = h2;
}

__device__ __forceinline__ void k1_chunk_sums_body(const double* __restrict__ tile, int64_t ntiles,
                                                   double2* __restrict__ chunk_tot) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= K1_CHUNKS * REC) return;
    const int ch = idx / REC, q = idx - ch * REC;
    const int64_t per = (ntiles + K1_CHUNKS - 1) / K1_CHUNKS;
    const int64_t t0 = ch * per, t1 = (t0 + per < ntiles) ? t0 + per : ntiles;
    double hi = 0.0, lo = 0.0;
    int64_t t = t0;
    for (; t + 8 <= t1; t += 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = tile[(t + u) * REC + q];
#pragma unroll
        for (int u = 0; u < 8; ++u) dd_add(hi, lo, v[u]);
    }
    for (; t < t1; ++t) dd_add(hi, lo, tile[t * REC + q]);
    chunk_tot[idx] = make_double2(hi, lo);
}
static __global__ void __launch_bounds__(128) k1_chunk_sums(const double* __restrict__ tile, int64_t ntiles,
                                                     double2* __restrict__ chunk_tot) {
    k1_chunk_sums_body(tile, ntiles, chunk_tot);
}

__device__ __forceinline__ void k1_chunk_scan_body(const double* __restrict__ tile, int64_t ntiles,
                                                   const double2* __restrict__ chunk_tot,
                                                   double2* __restrict__ C, bool write_total) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= K1_CHUNKS * REC) return;
    const int ch = idx / REC, q = idx - ch * REC;
    const int64_t per = (ntiles + K1_CHUNKS - 1) / K1_CHUNKS;
    const int64_t t0 = ch * per, t1 = (t0 + per < ntiles) ? t0 + per : ntiles;
    double hi = 0.0, lo = 0.0;
    for (int c = 0; c < ch; ++c) {                             // totals of the chunks before mine, in order
        const double2 v = chunk_tot[c * REC + q];
        dd_add(hi, lo, v.x);
        dd_add(hi, lo, v.y);
    }
    int64_t t = t0;
    for (; t + 8 <= t1; t += 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = tile[(t + u) * REC + q];
#pragma unroll
        for (int u = 0; u < 8; ++u) { C[(t + u) * REC + q] = make_double2(hi, lo); dd_add(hi, lo, v[u]); }
    }
    for (; t < t1; ++t) { C[t * REC + q] = make_double2(hi, lo); dd_add(hi, lo, tile[t * REC + q]); }
    // the thread that owns the last block also writes the grand total
    if (write_total && t1 == ntiles && t0 < ntiles) C[ntiles * REC + q] = make_double2(hi, lo);
}
static __global__ void __launch_bounds__(128) k1_chunk_scan(const double* __restrict__ tile, int64_t ntiles,
                                                     const double2* __restrict__ chunk_tot,
                                                     double2* __restrict__ C) {
    k1_chunk_scan_body(tile, ntiles, chunk_tot, C, true);
}
// packed batch: blockIdx.y = recording.  The scan runs over ceil(n / K1_TILE) blocks exactly as for
// a recording uploaded alone; the total lands in the recording's own extra block when n is a
// multiple of K1_TILE and is not written otherwise (the slot would be the next recording's zero).
static __global__ void __launch_bounds__(128) k1_chunk_sums_batch(const double* __restrict__ tile, const RecTab* __restrict__ tab,
                                                           double2* __restrict__ chunk_tot) {
    const RecTab t = tab[blockIdx.y];
    const int64_t nt = (t.n + K1_TILE - 1) / K1_TILE;
    if (nt == 0) return;
    k1_chunk_sums_body(tile + (t.base / K1_TILE) * REC, nt, chunk_tot + (int64_t)blockIdx.y * K1_CHUNKS * REC);
}
static __global__ void __launch_bounds__(128) k1_chunk_scan_batch(const double* __restrict__ tile, const RecTab* __restrict__ tab,
                                                           const double2* __restrict__ chunk_tot,
                                                           double2* __restrict__ C) {
    const RecTab t = tab[blockIdx.y];
    const int64_t nt = (t.n + K1_TILE - 1) / K1_TILE;
    const int64_t t0 = t.base / K1_TILE;
    if (nt == 0) {                                   // an empty recording: its one block prefix is zero
        const int idx = blockIdx.x * blockDim.x + threadIdx.x;
        if (idx < REC) C[t0 * REC + idx] = make_double2(0.0, 0.0);
        return;
    }
    k1_chunk_scan_body(tile + t0 * REC, nt, chunk_tot + (int64_t)blockIdx.y * K1_CHUNKS * REC, C + t0 * REC,
                       t.n % K1_TILE == 0);
}

// one CTA per block of K1_TILE frames: P[f0] = 0, P[f0 + t + 1] = running sums
// (t + 1 < K1_TILE), block total -> tile[]
__device__ __forceinline__ void k1_tile_write_body(const float* __restrict__ x, int64_t n,
                                                   const double* __restrict__ shift,
                                                   double* __restrict__ tile,
                                                   double* __restrict__ P) {
    __shared__ __align__(16) double xs[K1_TILE][K1_XS];
    const int64_t f0 = (int64_t)blockIdx.x * K1_TILE;
    k1_load_tile(x, n, f0, shift, xs);
    const int q = threadIdx.x;
    if (q >= REC) return;
    const int64_t left = n - f0;
    const int valid = left < K1_TILE ? (int)left : K1_TILE;
    const int r = c_row[q], c = c_col[q];
    double acc = 0.0;
    double* out = P + f0 * REC + q;
    __stcs(out, 0.0);                              // the prefix restarts at the block boundary
#pragma unroll 4
    for (int t = 0; t < valid; ++t) {
        acc = fma(xs[t][r], xs[t][c], acc);
        // streaming store: written once, read later by other kernels.  The record at the
        // next block boundary belongs to the next block (it is that block's zero).
        if (t + 1 < K1_TILE) __stcs(out + (int64_t)(t + 1) * REC, acc);
    }
    tile[(int64_t)blockIdx.x * REC + q] = acc;
    // a recording that ends exactly on a block boundary still needs its last (zero) record
    if (valid == K1_TILE && f0 + K1_TILE == n) __stcs(out + (int64_t)K1_TILE * REC, 0.0);
}
static __global__ void __launch_bounds__(K1_THREADS) k1_tile_write(const float* __restrict__ x, int64_t n,
                                                            const double* __restrict__ shift,
                                                            double* __restrict__ tile,
                                                            double* __restrict__ P) {
    k1_tile_write_body(x, n, shift, tile, P);
}
// packed batch: blockIdx.y = recording, blockIdx.x = block of the recording (n / K1_TILE + 1 of them)
static __global__ void __launch_bounds__(K1_THREADS) k1_tile_write_batch(const float* __restrict__ x,
                                                                  const RecTab* __restrict__ tab,
                                                                  const double* __restrict__ shift,
                                                                  double* __restrict__ tile,
                                                                  double* __restrict__ P) {
    const RecTab t = tab[blockIdx.y];
    if ((int64_t)blockIdx.x > t.n / K1_TILE) return;
    k1_tile_write_body(x + t.base * D39, t.n, shift + (int64_t)blockIdx.y * K1_XS,
                       tile + (t.base / K1_TILE) * REC, P + t.base * REC);
}

// ---- K5: statistics records of frame ranges straight from the frames ------------------------------
// What get_spk_features + np.cov of the clustering scripts need (spk-clustering.py:46-52, 91-96): (n, sum x,
// sum x x^T) of every initial cluster.  When no window search runs on the recording (spk-clustering.py started
// on its own, BASELINE configs 3 and 5) the 6,560 B-per-frame prefix of K1 is never needed: a task = one range of
// at most K5_SPAN frames, one CTA, thread q accumulates component q over the frames in order (same products, same
// shift as K1), 156 B read per frame and one record written per range.  Longer ranges are cut into tasks whose
// partial records k5_reduce adds in order.
constexpr int64_t K5_SPAN = 4096;
static __global__ void __launch_bounds__(K1_THREADS) k5_direct(const float* __restrict__ x, const double* __restrict__ shift,
                                                        const int64_t* __restrict__ task, int64_t ntask,
                                                        double* __restrict__ rec, double* __restrict__ part) {
    __shared__ __align__(16) double xs[K1_TILE][K1_XS];
    const int64_t a = task[blockIdx.x], b = task[ntask + blockIdx.x], dst = task[2 * ntask + blockIdx.x];
    const int q = threadIdx.x;
    const int r = q < REC ? c_row[q] : 0, c = q < REC ? c_col[q] : 0;
    double acc = 0.0;
    for (int64_t f0 = a; f0 < b; f0 += K1_TILE) {
        __syncthreads();                               // the previous tile has been consumed
        k1_load_tile(x, b, f0, shift, xs);             // frames at and past b count as absent
        const int64_t left = b - f0;
        const int valid = left < K1_TILE ? (int)left : K1_TILE;
        if (q < REC) {
#pragma unroll 4
            for (int t = 0; t < valid; ++t) acc = fma(xs[t][r], xs[t][c], acc);
        }
    }
    if (q < REC) (dst >= 0 ? rec + dst * REC : part + (-dst - 1) * REC)[q] = acc;
}
// red[0][k] = cluster, red[1][k] = its first partial record, red[2][k] = how many
static __global__ void __launch_bounds__(K1_THREADS) k5_reduce(const double* __restrict__ part, const int64_t* __restrict__ red,
                                                        int64_t nred, double* __restrict__ rec) {
    const int64_t s = red[blockIdx.x], p0 = red[nred + blockIdx.x], np = red[2 * nred + blockIdx.x];
    const int q = threadIdx.x;
    if (q >= REC) return;
    double acc = 0.0;
    for (int64_t p = 0; p < np; ++p) acc += part[(p0 + p) * REC + q];
    rec[s * REC + q] = acc;
}

}  // namespace spk
