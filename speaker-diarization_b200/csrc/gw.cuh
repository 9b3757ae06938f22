// gw.cuh - K3: the growing-window speaker-turn search as one persistent kernel.
//
// Replaces dist_gw (spk-change-detection.py:180-288).  The reference's loop is
// sequential: a window [start, end) is scanned for its best split, then either
// a change is written and `start` jumps there, or `end` grows by a schedule.
// Here the whole loop lives on the device, no host round trip per window:
//
//   * one CHAIN per recipe line; chains are independent and are pulled from a
//     queue by GROUPS of CTAs (group size = grid / #groups, chosen by the host:
//     one chain -> one group spanning the whole GPU, many chains -> one CTA each);
//   * per WAVE a group (1) plans a batch of windows: because the growth schedule
//     of `end` is deterministic while no change is found, the next few windows
//     of the same `start` are scored speculatively in the same wave; (2) every
//     warp of the group factorises one covariance per task (left / right /
//     pooled or GLR-mix term of one candidate); (3) after a group barrier each
//     CTA redundantly reduces the terms to distances and applies the reference's
//     decision rules in window order - identical inputs, identical code, so all
//     CTAs of the group reach the same decision without a broadcast;
//   * a positive window triggers the fine-tune wave (step-1 candidates around
//     the coarse maximum), then `start` advances and the left-term cache resets.
//
// All position arithmetic (`start`, `end`, `ws`, `dws`, candidate offsets) is
// IEEE fp64 in the reference's operation order; frame indices are truncations
// of those doubles.  Candidate offsets T[k] = fl(T[k-1] + istep)
// come from a host-built table so that non-dyadic frame rates round as in Python.
#pragma once

#include <climits>

#include "common.cuh"
#include "score.cuh"

namespace spk {

constexpr int GW_WARPS = 12;
constexpr int GW_THREADS = GW_WARPS * 32;
constexpr int GW_BMAX = GW_WARPS;       // windows per speculative batch (one deciding warp each)
constexpr int GW_JMAX = 128;            // fine-tune candidates (2*istep + 1 <= JMAX)
#define GW_NEG_INIT (-9223372036854775808.0)   /* -sys.maxint - 1 as a double, CD:203 */

struct GwDev {                 // kernel parameters
    Stats st;                  // two-level frame statistics
    const float* x;            // frames (KL2 means)
    const double* T;           // candidate offset table, kmax entries
    int64_t kmax;
    const int64_t* seg_a;      // chains
    const int64_t* seg_b;
    int32_t nchain;
    int32_t group_ctas;        // CTAs per group
    int32_t ngroups;
    int32_t bmax;              // speculation depth actually used (<= GW_BMAX)
    double rate, winsize, winstep, deltaws, threshold, lambda, minfeas, istep;
    int32_t metric;
    // workspaces, per group
    double* left;              // [ngroups][kmax]
    double* right;             // [ngroups][2 parity][bmax][kmax][rterms]
    double* pooled;            // [ngroups][2 parity][GW_BMAX]
    double* fine;              // [ngroups][2 parity][3][GW_JMAX]
    unsigned long long* bar;   // [ngroups] barrier counters
    int32_t* next_chain;       // queue cursor
    int32_t* group_chain;      // [ngroups][2] published next chain per group
    // output
    spkdiar_gw_window* win;
    int64_t win_cap;
    unsigned long long* nwin;  // records produced (may exceed win_cap -> E_CAPACITY)
    unsigned long long* dbg;   // optional phase cycle counters of CTA 0: plan, eval, barrier, decide, waves, tasks
};

struct GwPlan {                // shared memory, written by thread 0
    int mode;                  // 0 coarse, 1 fine, 2 chain finished
    int nW;                    // windows in this batch
    int nL;                    // new left terms
    int k0;                    // first new left k
    int ntask;
    int parity;
    int nJ;                    // fine candidates
    int rterms;
    double e[GW_BMAX];         // window ends
    double ws_after[GW_BMAX];  // growth state after a negative window w
    double dws_after[GW_BMAX];
    double e_after[GW_BMAX];
    int last[GW_BMAX];         // a negative window w ends the chain
    int K[GW_BMAX];            // coarse candidates of window w
    int sec[GW_BMAX + 1];      // task offsets of the right sections
    double fi[GW_JMAX];        // fine offsets i_j
    double pend_pl;            // pooled term of the window waiting for its fine tune
    long long row_lo[GW_WARPS];   // KL2: per-warp frame-row ranges of the cooperative mean pass
    long long row_hi[GW_WARPS];
    // decision scratch
    double bd[GW_BMAX];
    int bk[GW_BMAX];
    int ninf[GW_BMAX];
};

__device__ __forceinline__ void gw_group_barrier(unsigned long long* ctr, unsigned long long& target,
                                                 int group_ctas) {
    __syncthreads();
    if (group_ctas > 1) {
        if (threadIdx.x == 0) {
            target += (unsigned long long)group_ctas;
            __threadfence();
            atomicAdd(ctr, 1ULL);
            while (*((volatile unsigned long long*)ctr) < target) { }
            __threadfence();
        }
        __syncthreads();
    }
}

// count of k with T[k] < lim (T strictly increasing), from an arithmetic guess
__device__ __forceinline__ int gw_count_below(const double* __restrict__ T, int64_t kmax, double lim,
                                              double minfeas, double istep) {
    double g = (lim - minfeas) / istep;
    int64_t k = g > 0.0 ? (int64_t)g : 0;
    if (k > kmax) k = kmax;
    while (k < kmax && __ldg(T + k) < lim) ++k;
    while (k > 0 && !(__ldg(T + k - 1) < lim)) --k;
    return (int)k;
}

// ---- KL2: float32 sequential means, streamed cooperatively by the CTA ---------------
// np.mean(arr, 0) of the reference adds the float32 rows one after the other
// (SURVEY.md Q4), a serial chain per candidate and per side that re-reads the whole
// window.  The candidates one CTA evaluates in a wave are neighbours (same window,
// offsets 0.1 s apart), so their row ranges overlap almost completely: the CTA streams
// the rows ONCE through a 4-stage shared-memory ring (cp.async, 128 rows per stage) and
// every warp adds the rows of its own range [ra, rb) from shared memory, lane = dimension.
constexpr int GW_CROWS = 128;
constexpr int GW_CSTAGES = 4;
constexpr size_t GW_CRING_BYTES = sizeof(float) * GW_CSTAGES * GW_CROWS * D39;

__device__ __forceinline__ void gw_cta_means(const float* __restrict__ x, long long ra, long long rb,
                                             GwPlan& plan, float* cring, int warp, int lane, float* out) {
    if (lane == 0) { plan.row_lo[warp] = ra < rb ? ra : LLONG_MAX; plan.row_hi[warp] = ra < rb ? rb : LLONG_MIN; }
    __syncthreads();
    long long A = LLONG_MAX, B = LLONG_MIN;
#pragma unroll
    for (int w = 0; w < GW_WARPS; ++w) {
        A = plan.row_lo[w] < A ? plan.row_lo[w] : A;
        B = plan.row_hi[w] > B ? plan.row_hi[w] : B;
    }
    float s0 = 0.f, s1 = 0.f;
    const bool second = lane + 32 < D39;
    if (A < B) {
        const long long nst = (B - A + GW_CROWS - 1) / GW_CROWS;
        auto issue = [&](long long i) {
            if (i < nst) {
                const long long r0 = A + i * GW_CROWS;
                const long long left = B - r0;
                const int nel = (int)(left < GW_CROWS ? left : GW_CROWS) * D39;
                const float* src = x + r0 * D39;
                const unsigned dst = (unsigned)__cvta_generic_to_shared(cring + (i % GW_CSTAGES) * (GW_CROWS * D39));
                for (int e = threadIdx.x; e < nel; e += GW_THREADS)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 4u * e), "l"(src + e) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        for (int p = 0; p < GW_CSTAGES - 1; ++p) issue(p);
        for (long long i = 0; i < nst; ++i) {
            issue(i + GW_CSTAGES - 1);
            asm volatile("cp.async.wait_group %0;" ::"n"(GW_CSTAGES - 1) : "memory");
            __syncthreads();                                    // stage i has landed for everybody
            const long long r0 = A + i * GW_CROWS;
            const long long lo = ra > r0 ? ra : r0;
            long long hi = r0 + GW_CROWS;
            hi = rb < hi ? rb : hi;
            const float* buf = cring + (i % GW_CSTAGES) * (GW_CROWS * D39) + lane;
            long long r = lo;
            for (; r + 8 <= hi; r += 8) {                       // loads first, then the serial add chain
                float u[8], v[8];
                const float* row = buf + (int)(r - r0) * D39;
#pragma unroll
                for (int q = 0; q < 8; ++q) { u[q] = row[q * D39]; v[q] = second ? row[q * D39 + 32] : 0.f; }
#pragma unroll
                for (int q = 0; q < 8; ++q) { s0 = __fadd_rn(s0, u[q]); s1 = __fadd_rn(s1, v[q]); }
            }
            for (; r < hi; ++r) {
                const float* row = buf + (int)(r - r0) * D39;
                s0 = __fadd_rn(s0, row[0]);
                if (second) s1 = __fadd_rn(s1, row[32]);
            }
            __syncthreads();                                    // stage buffer free for a later stage
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const float fn = (float)(rb - ra);
    out[lane] = __fdiv_rn(s0, fn);
    if (second) out[lane + 32] = __fdiv_rn(s1, fn);
}

template <bool KL2>
__global__ void __launch_bounds__(GW_THREADS, 1) gw_kernel(const GwDev g) {
    extern __shared__ __align__(16) unsigned char gw_smem[];
    GwPlan& plan = *reinterpret_cast<GwPlan*>(gw_smem);
    unsigned char* scratch_base = gw_smem + ((sizeof(GwPlan) + 15) & ~(size_t)15);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = blockIdx.x / g.group_ctas;
    const int rank = blockIdx.x - group * g.group_ctas;      // CTA rank in its group
    const int gwarps = g.group_ctas * GW_WARPS;
    const int gpairs = gwarps / 2;            // KL2: two warps (one per side) work on one candidate
    const int rterms = KL2 ? 1 : (g.metric == SPKDIAR_GLR ? 2 : 1);
    double* left = g.left + (int64_t)group * g.kmax;
    double* right = g.right + (int64_t)group * 2 * g.bmax * g.kmax * rterms;
    double* pooled = g.pooled + (int64_t)group * 2 * GW_BMAX;
    double* fine = g.fine + (int64_t)group * 2 * 3 * GW_JMAX;
    unsigned long long* bar = g.bar + group;
    unsigned long long bar_target = 0;
    int wave = 0;               // parity source for the double-buffered term arrays
    int chain_pub = 0;          // parity of the published next-chain slot

    long long t_plan = 0, t_eval = 0, t_bar = 0, t_dec = 0, n_wave = 0, n_task = 0, t_e1 = 0, t_e2 = 0;
    int chain = group;
    while (chain < g.nchain) {
        // ---- chain state (identical in thread 0 of every CTA of the group) ----
        const int64_t base = g.seg_a[chain];
        const int64_t nfr = g.seg_b[chain] - base;
        const double n = (double)nfr;
        double start = 0.0;
        double end = start + g.winsize * 2;
        double ws = g.minfeas, dws = g.deltaws;
        int left_valid = 0;
        int seq = 0;
        bool done = !(end <= n);
        // pending positive window (between the coarse and the fine wave)
        double pend_e = 0.0, pend_maxi = 0.0, pend_maxd = 0.0;
        int pend_ncand = 0, pend_ninf = 0;
        bool want_fine = false;

        while (!done) {
            // ================= PLAN =================
            const long long c0 = clock64();
            const int parity = wave & 1;
            if (threadIdx.x == 0) {
                plan.parity = parity;
                plan.rterms = rterms;
                if (!want_fine) {
                    plan.mode = 0;
                    double e = end, w_ = ws, dw = dws;
                    int nW = 0;
                    for (int w = 0; w < g.bmax; ++w) {
                        plan.e[w] = e;
                        nW = w + 1;
                        // negative-branch growth, CD:273-284
                        int last = 0;
                        if (e + w_ <= n) {
                            e += w_;
                            if (w_ < g.winstep) { w_ += dw; dw *= 2; }
                            if (w_ > g.winstep) w_ = g.winstep;
                        } else if (e != n) {
                            e = n;
                        } else {
                            last = 1;
                        }
                        plan.e_after[w] = e; plan.ws_after[w] = w_; plan.dws_after[w] = dw; plan.last[w] = last;
                        if (last) break;
                    }
                    plan.nW = nW;
                } else {
                    plan.mode = 1;
                    // fine-tune offsets, CD:235-251: i = maxi - istep; while i < maxi + istep: ...; i += 1
                    double i = pend_maxi - g.istep;
                    const double endtune = pend_maxi + g.istep;
                    int nJ = 0;
                    while (i < endtune && nJ < GW_JMAX) { plan.fi[nJ++] = i; i += 1; }
                    plan.nJ = nJ;
                    plan.nW = 1;
                    plan.e[0] = pend_e;
                }
            }
            __syncthreads();
            if (plan.mode == 0) {
                if (threadIdx.x < plan.nW) {
                    const double lim = plan.e[threadIdx.x] - start - g.minfeas;     // CD:204
                    plan.K[threadIdx.x] = gw_count_below(g.T, g.kmax, lim, g.minfeas, g.istep);
                }
                __syncthreads();
                if (threadIdx.x == 0) {
                    // cut the batch to the group's one-round capacity (always keep window 0)
                    int nW = 0, tasks = 0, kmaxw = left_valid;
                    for (int w = 0; w < plan.nW; ++w) {
                        const int newl = plan.K[w] > kmaxw ? plan.K[w] - kmaxw : 0;
                        const int t = plan.K[w] * rterms + newl + ((!KL2 && g.metric == SPKDIAR_BIC) ? 1 : 0);
                        if (w > 0 && tasks + t > (KL2 ? gpairs : gwarps)) break;
                        tasks += t;
                        if (plan.K[w] > kmaxw) kmaxw = plan.K[w];
                        nW = w + 1;
                    }
                    plan.nW = nW;
                    plan.k0 = left_valid;
                    plan.nL = KL2 ? 0 : (kmaxw - left_valid);
                    int off = plan.nL + ((!KL2 && g.metric == SPKDIAR_BIC) ? nW : 0);
                    for (int w = 0; w < nW; ++w) { plan.sec[w] = off; off += plan.K[w] * rterms; }
                    plan.sec[nW] = off;
                    plan.ntask = off;
                }
            } else if (threadIdx.x == 0) {
                plan.ntask = plan.nJ * (KL2 ? 1 : (g.metric == SPKDIAR_GLR ? 3 : 2));
            }
            __syncthreads();

            // ================= EVALUATE =================
            const long long c1 = clock64();
            const int64_t s0 = base + (int64_t)start;
            // A wave with fewer tasks than warps spreads over all SMs instead of filling the first
            // CTAs.  BIC / GLR: task t -> CTA t % group_ctas.  KL2: every CTA takes a CONTIGUOUS
            // chunk of the round (neighbouring candidates) for the cooperative mean pass.
            for (int r0 = 0; r0 < plan.ntask; r0 += (KL2 ? gpairs : gwarps)) {
                int id;
                bool has;
                if (KL2) {
                    // two warps per candidate (left side / right side), gpairs candidates per round
                    const int nround = plan.ntask - r0 < gpairs ? plan.ntask - r0 : gpairs;
                    const int chunk = (nround + g.group_ctas - 1) / g.group_ctas;
                    const int pi = warp >> 1;
                    id = r0 + rank * chunk + pi;
                    has = pi < chunk && rank * chunk + pi < nround;
                } else {
                    id = r0 + warp * g.group_ctas + rank;
                    has = id < plan.ntask;
                }
                int64_t mm = s0, ee = s0;
                int term = 0;
                double* dst = nullptr;
                if (has) {
                    if (plan.mode == 0) {
                        const int npool = (!KL2 && g.metric == SPKDIAR_BIC) ? plan.nW : 0;
                        if (id < plan.nL) {                         // left term of a new coarse offset
                            const int k = plan.k0 + id;
                            mm = base + (int64_t)(start + __ldg(g.T + k));
                            ee = mm; term = 0; dst = left + k;
                        } else if (id < plan.nL + npool) {          // pooled term of window w (BIC)
                            const int w = id - plan.nL;
                            mm = s0; ee = base + (int64_t)plan.e[w]; term = 2;
                            dst = pooled + parity * GW_BMAX + w;
                        } else {                                    // right (and GLR mix) terms
                            int w = 0;
                            while (id >= plan.sec[w + 1]) ++w;
                            const int r = id - plan.sec[w];
                            const int k = r / rterms, sub = r - k * rterms;
                            mm = base + (int64_t)(start + __ldg(g.T + k));
                            ee = base + (int64_t)plan.e[w];
                            term = KL2 ? 3 : (sub == 0 ? 1 : 2);
                            dst = right + (((int64_t)parity * g.bmax + w) * g.kmax + k) * rterms + sub;
                        }
                    } else {
                        const int per = KL2 ? 1 : (g.metric == SPKDIAR_GLR ? 3 : 2);
                        const int j = id / per, sub = id - j * per;
                        mm = base + (int64_t)(start + plan.fi[j]);
                        ee = base + (int64_t)plan.e[0];
                        term = KL2 ? 3 : sub;
                        dst = fine + ((int64_t)parity * 3 + sub) * GW_JMAX + j;
                    }
                }
                if (KL2) {
                    Kl2Scratch* kall = reinterpret_cast<Kl2Scratch*>(scratch_base);
                    Kl2Scratch& own = kall[warp];               // factorisation scratch of this warp
                    Kl2Scratch& pair = kall[warp & ~1];         // results of the candidate (both sides)
                    float* cring = reinterpret_cast<float*>(scratch_base + GW_WARPS * sizeof(Kl2Scratch));
                    const int side = warp & 1;
                    const long long ra = has ? (side ? mm : s0) : 0;
                    const long long rb = has ? (side ? ee : mm) : 0;
                    const long long k0 = clock64();
                    if (has) kl2_side_one(WinSrc(g.st, ra, rb, REC), own, pair.dS[side], pair.dP[side], lane);
                    const long long k1 = clock64();
                    gw_cta_means(g.x, ra, rb, plan, cring, warp, lane, pair.mean[side]);   // ends with __syncthreads
                    t_e1 += k1 - k0; t_e2 += clock64() - k1;
                    if (has && side == 0) {
                        double t1, t2;
                        const double v = kl2_finish(pair, lane, &t1, &t2);
                        if (lane == 0) *dst = v;
                    }
                    __syncthreads();                            // pair results consumed before the next round
                } else if (has) {
                    WarpScratch& wsr = reinterpret_cast<WarpScratch*>(scratch_base)[warp];
                    const WinSrc X(g.st, s0, mm, REC);
                    const WinSrc Y(g.st, mm, ee, REC);
                    const double v = logdet_term(term, g.metric, X, Y, wsr, lane);
                    if (lane == 0) *dst = v;
                }
            }
            const long long c2 = clock64();
            gw_group_barrier(bar, bar_target, g.group_ctas);
            const long long c3 = clock64();
            ++wave;

            // ================= DECIDE =================
            if (plan.mode == 0) {
                if (warp < plan.nW) {
                    const int w = warp;
                    const int64_t e0 = (int64_t)plan.e[w];
                    const int64_t s0r = (int64_t)start;
                    const double pl = (!KL2 && g.metric == SPKDIAR_BIC) ? __ldcg(pooled + parity * GW_BMAX + w) : 0.0;
                    const double pen = (!KL2 && g.metric == SPKDIAR_BIC) ? bic_pen(g.lambda, (double)(e0 - s0r)) : 0.0;
                    double bd = GW_NEG_INIT; int bk = -1; int ninf = 0;
                    for (int k = lane; k < plan.K[w]; k += 32) {
                        const int64_t m = (int64_t)(start + __ldg(g.T + k));
                        const double N1 = (double)(m - s0r), N2 = (double)(e0 - m);
                        const double* rp = right + (((int64_t)parity * g.bmax + w) * g.kmax + k) * rterms;
                        double d;
                        if (KL2) d = __ldcg(rp);
                        else if (g.metric == SPKDIAR_BIC) d = bic_combine_pen(N1, N2, __ldcg(left + k), __ldcg(rp), pl, pen);
                        else d = glr_combine(N1, N2, __ldcg(left + k), __ldcg(rp), __ldcg(rp + 1));
                        if (d == d_inf() || d == -d_inf()) ++ninf;          // CD:219-220
                        else if (d > bd) { bd = d; bk = k; }                // CD:215-217 (strict, first wins)
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const double od = __shfl_xor_sync(0xffffffffu, bd, o);
                        const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
                        ninf += __shfl_xor_sync(0xffffffffu, ninf, o);
                        if (ok >= 0 && (bk < 0 || od > bd || (od == bd && ok < bk))) { bd = od; bk = ok; }
                    }
                    if (lane == 0) { plan.bd[w] = bd; plan.bk[w] = bk; plan.ninf[w] = ninf; }
                }
                __syncthreads();
                if (threadIdx.x == 0) {
                    int kmaxw = left_valid;
                    // negative windows ahead of the first positive one: their records take
                    // consecutive slots claimed with ONE atomic
                    int nneg = 0;
                    for (int w = 0; w < plan.nW; ++w) {
                        if (plan.bd[w] > g.threshold && plan.bk[w] >= 0) break;
                        ++nneg;
                        if (plan.last[w]) break;
                    }
                    unsigned long long slot = 0;
                    if (rank == 0 && nneg > 0) slot = atomicAdd(g.nwin, (unsigned long long)nneg);
                    for (int w = 0; w < plan.nW; ++w) {
                        if (plan.K[w] > kmaxw) kmaxw = plan.K[w];
                        const double maxd = plan.bd[w];
                        const double maxi = plan.bk[w] >= 0 ? __ldg(g.T + plan.bk[w]) : 0.0;
                        const bool positive = maxd > g.threshold && plan.bk[w] >= 0;       // CD:230
                        if (positive) {
                            want_fine = true;
                            pend_e = plan.e[w]; pend_maxi = maxi; pend_maxd = maxd;
                            pend_ncand = plan.K[w]; pend_ninf = plan.ninf[w];
                            if (!KL2 && g.metric == SPKDIAR_BIC) plan.pend_pl = __ldcg(pooled + parity * GW_BMAX + w);
                            end = plan.e[w];
                            break;
                        }
                        if (rank == 0) {                                                    // negative window record
                            if ((int64_t)slot < g.win_cap) {
                                spkdiar_gw_window r;
                                r.start = start; r.end = plan.e[w]; r.maxi = maxi; r.maxd = maxd;
                                r.maxi_fine = 0.0; r.maxd_fine = 0.0; r.positive = 0; r.chain = chain;
                                r.ncand = plan.bk[w] >= 0 ? plan.K[w] : -plan.K[w] - 1;
                                r.ninf = plan.ninf[w]; r.seq = seq; r.pad = 0;
                                g.win[slot] = r;
                            }
                            ++slot;
                        }
                        ++seq;
                        end = plan.e_after[w]; ws = plan.ws_after[w]; dws = plan.dws_after[w];
                        if (plan.last[w]) { done = true; break; }
                    }
                    if (!KL2) left_valid = kmaxw;
                    plan.mode = done ? 2 : (want_fine ? 1 : 0);
                }
            } else {
                // fine-tune decision, CD:237-251: strict improvement over the coarse maximum, first wins
                if (warp == 0) {
                    const int64_t e0 = (int64_t)plan.e[0];
                    const int64_t s0r = (int64_t)start;
                    const double* f0 = fine + (int64_t)parity * 3 * GW_JMAX;
                    const double pl = plan.pend_pl;
                    const double pen = (!KL2 && g.metric == SPKDIAR_BIC) ? bic_pen(g.lambda, (double)(e0 - s0r)) : 0.0;
                    double bd = GW_NEG_INIT; int bj = -1; int ninf = 0;
                    for (int j = lane; j < plan.nJ; j += 32) {
                        const int64_t m = (int64_t)(start + plan.fi[j]);
                        const double N1 = (double)(m - s0r), N2 = (double)(e0 - m);
                        double d;
                        if (KL2) d = __ldcg(f0 + j);
                        else if (g.metric == SPKDIAR_BIC) d = bic_combine_pen(N1, N2, __ldcg(f0 + j), __ldcg(f0 + GW_JMAX + j), pl, pen);
                        else d = glr_combine(N1, N2, __ldcg(f0 + j), __ldcg(f0 + GW_JMAX + j), __ldcg(f0 + 2 * GW_JMAX + j));
                        if (d == d_inf() || d == -d_inf()) ++ninf;
                        else if (d > bd) { bd = d; bj = j; }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const double od = __shfl_xor_sync(0xffffffffu, bd, o);
                        const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
                        ninf += __shfl_xor_sync(0xffffffffu, ninf, o);
                        if (oj >= 0 && (bj < 0 || od > bd || (od == bd && oj < bj))) { bd = od; bj = oj; }
                    }
                    if (lane == 0) { plan.bd[0] = bd; plan.bk[0] = bj; plan.ninf[0] = ninf; }
                }
                __syncthreads();
                if (threadIdx.x == 0) {
                    double maxd = pend_maxd, maxi = pend_maxi;
                    if (plan.bk[0] >= 0 && plan.bd[0] > pend_maxd) { maxd = plan.bd[0]; maxi = plan.fi[plan.bk[0]]; }
                    if (rank == 0) {
                        const unsigned long long slot = atomicAdd(g.nwin, 1ULL);
                        if ((int64_t)slot < g.win_cap) {
                            spkdiar_gw_window r;
                            r.start = start; r.end = pend_e; r.maxi = pend_maxi; r.maxd = pend_maxd;
                            r.maxi_fine = maxi; r.maxd_fine = maxd; r.positive = 1; r.chain = chain;
                            r.ncand = pend_ncand; r.ninf = pend_ninf + plan.ninf[0]; r.seq = seq; r.pad = 0;
                            g.win[slot] = r;
                        }
                    }
                    ++seq;
                    want_fine = false;
                    left_valid = 0;                                     // CD:256
                    start += maxi;                                      // CD:263
                    if (start + g.winsize * 2 <= n) {                   // CD:264-268
                        end = start + g.winsize * 2;
                        ws = g.minfeas; dws = g.deltaws;
                    } else {
                        done = true;                                    // CD:269-270
                    }
                    plan.mode = done ? 2 : 0;
                }
            }
            __syncthreads();
            // every thread follows thread 0's view of the chain state
            done = plan.mode == 2;
            want_fine = plan.mode == 1;
            // `start`, `end`, ... live in thread 0; the others need start / left_valid for the next wave
            if (threadIdx.x == 0) { plan.e_after[0] = start; plan.K[0] = left_valid; }
            __syncthreads();
            start = plan.e_after[0];
            left_valid = plan.K[0];
            __syncthreads();
            t_plan += c1 - c0; t_eval += c2 - c1; t_bar += c3 - c2; t_dec += clock64() - c3; ++n_wave; n_task += plan.ntask;
        }

        // ---- next chain for this group ----
        if (g.ngroups >= g.nchain) break;           // every chain had its own group
        if (rank == 0 && threadIdx.x == 0) {
            const int c = g.ngroups + atomicAdd(g.next_chain, 1);
            g.group_chain[group * 2 + chain_pub] = c;
        }
        gw_group_barrier(bar, bar_target, g.group_ctas);
        chain = *((volatile int32_t*)(g.group_chain + group * 2 + chain_pub));
        chain_pub ^= 1;
    }
    if (g.dbg && blockIdx.x == 0 && threadIdx.x == 0) {
        g.dbg[0] = t_plan; g.dbg[1] = t_eval; g.dbg[2] = t_bar; g.dbg[3] = t_dec; g.dbg[4] = n_wave; g.dbg[5] = n_task; g.dbg[6] = t_e1; g.dbg[7] = t_e2;
    }
}

inline cudaError_t gw_configure() {
    const size_t plan = (sizeof(GwPlan) + 15) & ~(size_t)15;
    cudaError_t e = cudaFuncSetAttribute(gw_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(plan + GW_WARPS * sizeof(WarpScratch)));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(gw_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(plan + GW_WARPS * sizeof(Kl2Scratch) + GW_CRING_BYTES));
}

}  // namespace spk
