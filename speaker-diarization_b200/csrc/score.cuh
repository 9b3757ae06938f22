// score.cuh - K2: batched window / pair scoring, one warp per log-determinant.
//
// Replaces the bodies of bic / glr / kl2 (spk-change-detection.py:72-133,
// spk-clustering.py:81-133).  The arithmetic that combines log-determinants
// into a distance mirrors the reference expression by expression (explicit
// _rn intrinsics: no FMA contraction), so that the only deviation from the
// reference is the last-bits difference between an LDL^T log-determinant and
// np.log(scipy.linalg.det(.)).
#pragma once

#include "common.cuh"
#include "ldl.cuh"

namespace spk {

static_assert(STAT_BLOCK == 128, "WinSrc block size must equal K1_TILE");
constexpr int SC_WARPS = 4;                 // warps per CTA in the scoring kernels
constexpr int SC_THREADS = SC_WARPS * 32;
constexpr size_t SC_SMEM = SC_WARPS * sizeof(WarpScratch);     // dynamic shared memory of the scoring kernels

// ---- distance formulas (fp64, reference operation order) ----------------------------
// BIC, spk-change-detection.py:95-99 / spk-clustering.py:95-99:
//   d = 0.5*N*ln|S| - 0.5*N1*ln|S1| - 0.5*N2*ln|S2|;  d -= lambda*0.5*(p+0.5*p*(p+1))*ln N
// the penalty depends on N = N1 + N2 only: one log per window, not per candidate
__device__ __forceinline__ double bic_pen(double lambda, double N) {
    const double p = (double)c_dim;                 // the dimension of the feature files (ldl.cuh)
    const double k = __dmul_rn(__dmul_rn(lambda, 0.5), __dadd_rn(p, __dmul_rn(__dmul_rn(0.5, p), p + 1.0)));
    return __dmul_rn(k, log(N));
}
__device__ __forceinline__ double bic_combine_pen(double N1, double N2, double ld1, double ld2,
                                                  double ld, double pen) {
    const double N = N1 + N2;
    const double t0 = __dmul_rn(__dmul_rn(0.5, N), ld);
    const double c1 = __dmul_rn(__dmul_rn(0.5, N1), ld1);
    const double t2 = __dmul_rn(__dmul_rn(0.5, N2), ld2);
    return __dsub_rn(__dsub_rn(__dsub_rn(t0, c1), t2), pen);
}
__device__ __forceinline__ double bic_combine(double N1, double N2, double ld1, double ld2,
                                              double ld, double lambda) {
    return bic_combine_pen(N1, N2, ld1, ld2, ld, bic_pen(lambda, N1 + N2));
}
// GLR, spk-change-detection.py:114-115:
//   d = -(N/2) * ((N1/N)*ln|S1| + (N2/N)*ln|S2| - ln|(N1/N) S1 + (N2/N) S2|)
__device__ __forceinline__ double glr_combine(double N1, double N2, double ld1, double ld2, double ldm) {
    const double N = N1 + N2;
    const double a = __dmul_rn(__ddiv_rn(N1, N), ld1);
    const double b = __dmul_rn(__ddiv_rn(N2, N), ld2);
    const double s = __dsub_rn(__dadd_rn(a, b), ldm);
    return __dmul_rn(-__ddiv_rn(N, 2.0), s);
}
__device__ __forceinline__ void glr_weights(double N1, double N2, double& wx, double& wy) {
    const double N = N1 + N2;
    wx = (N1 / N) / (N1 - 1.0);           // (N1/N) * S1 with S1 = M1/(N1-1)
    wy = (N2 / N) / (N2 - 1.0);
}
__device__ __forceinline__ double range_map(double v) {     // np.log(det()) range, Q12
    if (v < -744.4400719213812) return -d_inf();
    if (v > 709.782712893384) return d_inf();
    return v;
}

// ---- one log-determinant term of a candidate ------------------------------------------
// term 0: left [a,m)   term 1: right [m,b)   term 2: pooled [a,b) (BIC) or the GLR mix
template <class SrcX, class SrcY>
__device__ __forceinline__ double logdet_term(int term, int metric, const SrcX& X, const SrcY& Y,
                                              WarpScratch& w, int lane) {
    const int kind = term == 0 ? FORM_X : (term == 1 ? FORM_Y : (metric == SPKDIAR_BIC ? FORM_POOL : FORM_MIX));
    // memory phase: the operand records -> shared memory
    const double* rx = w.rec[0];
    const double* ry = w.rec[1];
    if (kind != FORM_Y) rx = stage_record(X, w.rec[0], lane);
    if (kind != FORM_X) ry = stage_record(Y, w.rec[1], lane);
    __syncwarp();
    // compute phase
    double a[Grid<D39>::NSLOT];
    const SmemSrc sx{rx}, sy{ry};
    double wx = 1.0, wy = 1.0;
    if (kind == FORM_MIX) glr_weights(sx(L39::CNT), sy(L39::CNT), wx, wy);
    const double n = form_matrix<D39>(a, kind, sx, sy, wx, wy, w, lane);
    const double lm = ldl_logdet<D39>(a, w, lane);
    if (kind == FORM_MIX) return range_map(lm);
    return finish_logdet(lm, n, D39);
}

// ---- windows of one recording: candidate k = (a, m, b) -----------------------------------
static __global__ void __launch_bounds__(SC_THREADS, 3)
win_terms_kernel(const Stats st, const int64_t* __restrict__ a,
                 const int64_t* __restrict__ m, const int64_t* __restrict__ b,
                 int64_t ncand, int metric, double* __restrict__ terms) {
    extern __shared__ __align__(16) unsigned char sc_smem[];
    WarpScratch* ws = reinterpret_cast<WarpScratch*>(sc_smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t ntask = ncand * 3;
    for (int64_t id = (int64_t)blockIdx.x * SC_WARPS + warp; id < ntask; id += (int64_t)gridDim.x * SC_WARPS) {
        const int64_t k = id / 3;
        const int term = (int)(id - 3 * k);
        const WinSrc X(st, a[k], m[k], REC);
        const WinSrc Y(st, m[k], b[k], REC);
        const double v = logdet_term(term, metric, X, Y, ws[warp], lane);
        if (lane == 0) terms[id] = v;
    }
}

static __global__ void win_combine_kernel(const int64_t* __restrict__ a, const int64_t* __restrict__ m,
                                   const int64_t* __restrict__ b, int64_t ncand, int metric,
                                   double lambda, const double* __restrict__ terms,
                                   double* __restrict__ out) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= ncand) return;
    const double N1 = (double)(m[k] - a[k]), N2 = (double)(b[k] - m[k]);
    const double* t = terms + 3 * k;
    out[k] = metric == SPKDIAR_BIC ? bic_combine(N1, N2, t[0], t[1], t[2], lambda)
                                   : glr_combine(N1, N2, t[0], t[1], t[2]);
}

// ---- KL2 (reference semantics: diagonal-only formula, float32 sequential means) ----------
// spk-change-detection.py:124-133, SURVEY.md Q3 / Q4.
constexpr int KL2_ROWS = 13;     // frame rows per staging buffer
constexpr int KL2_RING = 3;      // buffers in flight
struct Kl2Scratch {
    LdlScratch w;
    // three uses that never overlap in time: the staged operand record (until the
    // matrix is formed), the stored factor L (factorisation + inverse), the frame-row
    // ring of the float32 means (after both sides are done)
    union {
        double rec[REC];
        double Lsm[(D39 * (D39 - 1)) / 2 + 3];
        float  ring[KL2_RING][KL2_ROWS * D39];
    };
    double pinv[VS];
    double dS[2][VS];      // diag(S) of side 0 / 1
    double dP[2][VS];      // diag(S^-1)
    float  mean[2][VS];    // float32 means
};
static_assert(sizeof(float) * KL2_RING * KL2_ROWS * D39 <= sizeof(double) * REC, "ring must fit the record buffer");

// float32 sequential mean over the concatenation of `nr` frame ranges, exactly
// as np.mean(arr, 0) accumulates a float32 matrix row by row:
// lane j owns dimension j (and j + 32), adds row after row with __fadd_rn, and
// divides by the count in float32.  The adds are a serial chain by definition;
// what can be hidden is the memory latency: rows are staged through a 3-deep
// ring of 16-row shared-memory buffers filled by cp.async, so 48 rows are always
// in flight.
__device__ __forceinline__ void kl2_stage(const float* __restrict__ src, int nel, float* buf, int lane) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(buf);
    for (int e = lane; e < nel; e += 32)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 4u * e), "l"(src + e) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void seq_mean_f32(const float* __restrict__ x, const int64_t* ra,
                                             const int64_t* rb, int64_t nr, int lane,
                                             float (*ring)[KL2_ROWS * D39], float* out) {
    float s0 = 0.f, s1 = 0.f;
    int64_t cnt = 0;
    const bool second = lane + 32 < D39;
    for (int64_t r = 0; r < nr; ++r) {
        const int64_t a = ra[r], b = rb[r];
        if (b <= a) continue;
        cnt += b - a;
        const int64_t ntile = (b - a + KL2_ROWS - 1) / KL2_ROWS;
        auto rows_of = [&](int64_t t) { const int64_t left = b - (a + t * KL2_ROWS); return (int)(left < KL2_ROWS ? left : KL2_ROWS); };
        for (int64_t t = 0; t < KL2_RING - 1; ++t) {            // prologue: fill the ring
            if (t < ntile) kl2_stage(x + (a + t * KL2_ROWS) * D39, rows_of(t) * D39, ring[t % KL2_RING], lane);
            else asm volatile("cp.async.commit_group;" ::: "memory");
        }
        for (int64_t t = 0; t < ntile; ++t) {
            const int64_t nx = t + KL2_RING - 1;
            if (nx < ntile) kl2_stage(x + (a + nx * KL2_ROWS) * D39, rows_of(nx) * D39, ring[nx % KL2_RING], lane);
            else asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group %0;" ::"n"(KL2_RING - 1) : "memory");
            __syncwarp();
            const float* buf = ring[t % KL2_RING];
            const int rows = rows_of(t);
#pragma unroll 4
            for (int q = 0; q < rows; ++q) {
                s0 = __fadd_rn(s0, buf[q * D39 + lane]);
                if (second) s1 = __fadd_rn(s1, buf[q * D39 + lane + 32]);
            }
            __syncwarp();                                        // buffer free for the stage after next
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    const float fn = (float)cnt;
    out[lane] = __fdiv_rn(s0, fn);
    if (second) out[lane + 32] = __fdiv_rn(s1, fn);
}

// one side of a KL2 evaluation: diag(S) and diag(S^-1) of the window / cluster `gsrc`
// into dS / dP (shared memory); `k` is the calling warp's own scratch.
// `k` is any scratch type with members w (LdlScratch), rec / Lsm (a union) and pinv.
template <class Src, class Scr>
__device__ __forceinline__ void kl2_side_one(const Src& gsrc, Scr& k, double* dS, double* dP, int lane) {
    const SmemSrc src{stage_record(gsrc, k.rec, lane)};
    __syncwarp();
    double a[Grid<D39>::NSLOT];
    const double n = form_matrix<D39>(a, FORM_X, src, src, 1.0, 1.0, k.w, lane);
    const double rn1 = 1.0 / (n - 1.0);
    for (int j = lane; j < D39; j += 32) {
        const double s = k.w.s0[j];
        dS[j] = (src(L39::pos_diag(j)) - s * s / n) * rn1;
    }
    __syncwarp();
    // factorisation with the inverse of the factor accumulated in the finished register slots
    ldl_logdet_inv<D39>(a, k.w, lane, k.pinv, n - 1.0, dP);
    __syncwarp();
}

// both sides by one warp.  The loop is deliberately not unrolled: one copy of the
// factorisation per kernel.
template <class Src>
__device__ __forceinline__ void kl2_sides(const Src& X, const Src& Y, Kl2Scratch& k, int lane) {
#pragma unroll 1
    for (int side = 0; side < 2; ++side) {
        const Src gsrc = side ? Y : X;
        kl2_side_one(gsrc, k, k.dS[side], k.dP[side], lane);
    }
}

__device__ __forceinline__ double kl2_finish(Kl2Scratch& k, int lane, double* t1_out, double* t2_out) {
    double t1 = 0.0, t2 = 0.0;
    for (int j = lane; j < D39; j += 32) {
        const double delta = (double)__fsub_rn(k.mean[0][j], k.mean[1][j]);
        t1 += (k.dS[0][j] - k.dS[1][j]) * (k.dP[1][j] - k.dP[0][j]);
        t2 += __dmul_rn(__dmul_rn(k.dP[0][j] + k.dP[1][j], delta), delta);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        t1 += __shfl_xor_sync(0xffffffffu, t1, o);
        t2 += __shfl_xor_sync(0xffffffffu, t2, o);
    }
    *t1_out = t1; *t2_out = t2;
    return __dadd_rn(__dmul_rn(0.5, t1), __dmul_rn(0.5, t2));
}

constexpr int KS = 2 * VS;              // doubles per cached KL2 side: diag(S)[40], diag(S^-1)[40]

// distance of one KL2 pair from the two cached sides (diag S [VS], diag S^-1 [VS]) and running float32 sums
// (one warp; the growing-window search caches them per offset, the clustering engine per cluster)
// (LSHARED: the left side and sum sit in shared memory - the merged cluster of the clustering engine)
template <bool LSHARED = false>
__device__ __forceinline__ double kl2_distance_cached(const double* sideL, const double* sideR,
                                                  const float* sumL, const float* sumR,
                                                  double nL, double nR, int lane) {
    double t1 = 0.0, t2 = 0.0;
    const float fl = (float)nL, fr = (float)nR;
    for (int j = lane; j < D39; j += 32) {
        const float m0 = __fdiv_rn(LSHARED ? sumL[j] : __ldcg(sumL + j), fl), m1 = __fdiv_rn(__ldcg(sumR + j), fr);
        const double delta = (double)__fsub_rn(m0, m1);
        const double s0 = LSHARED ? sideL[j] : __ldcg(sideL + j), p0 = LSHARED ? sideL[VS + j] : __ldcg(sideL + VS + j);
        const double s1 = __ldcg(sideR + j), p1 = __ldcg(sideR + VS + j);
        t1 += (s0 - s1) * (p1 - p0);
        t2 += __dmul_rn(__dmul_rn(p0 + p1, delta), delta);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        t1 += __shfl_xor_sync(0xffffffffu, t1, o);
        t2 += __shfl_xor_sync(0xffffffffu, t2, o);
    }
    return __dadd_rn(__dmul_rn(0.5, t1), __dmul_rn(0.5, t2));
}

static __global__ void __launch_bounds__(SC_THREADS, 2)
win_kl2_kernel(const Stats st, const float* __restrict__ x,
               const int64_t* __restrict__ a, const int64_t* __restrict__ m,
               const int64_t* __restrict__ b, int64_t ncand, double* __restrict__ out,
               double* __restrict__ terms) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Kl2Scratch* ks = reinterpret_cast<Kl2Scratch*>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Kl2Scratch& k = ks[warp];
    for (int64_t id = (int64_t)blockIdx.x * SC_WARPS + warp; id < ncand; id += (int64_t)gridDim.x * SC_WARPS) {
        const int64_t aa = a[id], mm = m[id], bb = b[id];
        kl2_sides(WinSrc(st, aa, mm, REC), WinSrc(st, mm, bb, REC), k, lane);
        seq_mean_f32(x, &aa, &mm, 1, lane, k.ring, k.mean[0]);
        seq_mean_f32(x, &mm, &bb, 1, lane, k.ring, k.mean[1]);
        __syncwarp();
        double t1, t2;
        const double d = kl2_finish(k, lane, &t1, &t2);
        if (lane == 0) {
            out[id] = d;
            if (terms) { terms[3 * id] = t1; terms[3 * id + 1] = t2; terms[3 * id + 2] = 0.0; }
        }
        __syncwarp();
    }
}

// ---- sets of frame ranges (clusters given by the host) -----------------------------------
// record of a set = sum over its ranges of (P[b] - P[a]), plain cluster record
static __global__ void __launch_bounds__(256)
set_records_kernel(const Stats st, const int64_t* __restrict__ off,
                   const int64_t* __restrict__ ra, const int64_t* __restrict__ rb,
                   int64_t nsets, double* __restrict__ rec) {
    const int64_t s = blockIdx.x;
    if (s >= nsets) return;
    for (int q = threadIdx.x; q < REC; q += blockDim.x) {
        double acc = 0.0;
        for (int64_t r = off[s]; r < off[s + 1]; ++r) acc += WinSrc(st, ra[r], rb[r], REC)(q);
        rec[s * REC + q] = acc;
    }
}

// pair p scores record recX[p] against recY[p]
static __global__ void __launch_bounds__(SC_THREADS, 3)
pair_terms_kernel(const double* recX, const double* recY, int64_t npairs, int metric,
                  double* __restrict__ terms) {
    extern __shared__ __align__(16) unsigned char sc_smem[];
    WarpScratch* ws = reinterpret_cast<WarpScratch*>(sc_smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t ntask = npairs * 3;
    for (int64_t id = (int64_t)blockIdx.x * SC_WARPS + warp; id < ntask; id += (int64_t)gridDim.x * SC_WARPS) {
        const int64_t k = id / 3;
        const int term = (int)(id - 3 * k);
        const RecSrc X{recX + k * REC};
        const RecSrc Y{recY + k * REC};
        const double v = logdet_term(term, metric, X, Y, ws[warp], lane);
        if (lane == 0) terms[id] = v;
    }
}

static __global__ void pair_combine_kernel(const double* __restrict__ recX, const double* __restrict__ recY,
                                    int64_t npairs, int metric, double lambda,
                                    const double* __restrict__ terms, double* __restrict__ out) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= npairs) return;
    const double N1 = recX[k * REC + L39::CNT], N2 = recY[k * REC + L39::CNT];
    const double* t = terms + 3 * k;
    out[k] = metric == SPKDIAR_BIC ? bic_combine(N1, N2, t[0], t[1], t[2], lambda)
                                   : glr_combine(N1, N2, t[0], t[1], t[2]);
}

static __global__ void __launch_bounds__(SC_THREADS, 2)
pair_kl2_kernel(const double* recX, const double* recY, const float* __restrict__ x,
                const int64_t* __restrict__ off1, const int64_t* __restrict__ a1, const int64_t* __restrict__ b1,
                const int64_t* __restrict__ off2, const int64_t* __restrict__ a2, const int64_t* __restrict__ b2,
                int64_t npairs, double* __restrict__ out, double* __restrict__ terms) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Kl2Scratch* ks = reinterpret_cast<Kl2Scratch*>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Kl2Scratch& k = ks[warp];
    for (int64_t id = (int64_t)blockIdx.x * SC_WARPS + warp; id < npairs; id += (int64_t)gridDim.x * SC_WARPS) {
        kl2_sides(RecSrc{recX + id * REC}, RecSrc{recY + id * REC}, k, lane);
        seq_mean_f32(x, a1 + off1[id], b1 + off1[id], off1[id + 1] - off1[id], lane, k.ring, k.mean[0]);
        seq_mean_f32(x, a2 + off2[id], b2 + off2[id], off2[id + 1] - off2[id], lane, k.ring, k.mean[1]);
        __syncwarp();
        double t1, t2;
        const double d = kl2_finish(k, lane, &t1, &t2);
        if (lane == 0) {
            out[id] = d;
            if (terms) { terms[3 * id] = t1; terms[3 * id + 1] = t2; terms[3 * id + 2] = 0.0; }
        }
        __syncwarp();
    }
}

}  // namespace spk
