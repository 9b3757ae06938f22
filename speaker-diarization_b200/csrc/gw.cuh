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
//
// KL2 (spk-change-detection.py:124-133) needs, besides diag(S) and diag(S^-1) of
// both sides, the reference's float32 means: np.mean adds the float32 rows of the
// slice one after the other, so a mean is a serial chain of fp32
// additions that depends on where the slice STARTS.  While `start` does not move
//   - all left means [start, m_k) are snapshots of ONE running sum from `start`;
//   - the right sum [m_k, end) of candidate k is a running sum from m_k that only
//     has to be EXTENDED when `end` grows;
// so the kernel keeps one running sum per candidate offset ("sum chain") alive in
// HBM across windows and waves, and a wave only adds the new rows.  The rows are
// streamed through a shared-memory ring filled by bulk asynchronous copies (TMA),
// four chains (one per chain warp) sharing one stream per CTA while the other warps
// factorise.  Left
// sides are cached per offset exactly like the left BIC term.  A KL2 wave has two
// steps separated by a group barrier: sides + sums, then the distances.
#pragma once

#include <climits>

#include "common.cuh"
#include "score.cuh"
#include "tma.cuh"

namespace spk {

constexpr int GW_WARPS = 12;
constexpr int GW_THREADS = GW_WARPS * 32;
constexpr int GW_BMAX = GW_WARPS;       // windows per speculative batch (one deciding warp each)
constexpr int GW_JMAX = 160;            // fine-tune candidates (2*istep + 1 <= JMAX; fused first wave: every frame
                                        // a fine tune of window 0 could look at)
#define GW_NEG_INIT (-9223372036854775808.0)   /* -sys.maxint - 1 as a double, CD:203 */

constexpr int GW_RING_ROWS = 64;        // frame rows per ring stage (a multiple of 4: 16-byte spans)
constexpr int GW_RING_STAGES = 6;
constexpr int GW_STAGE_FLOATS = GW_RING_ROWS * D39;
constexpr int GW_CHAIN_WARPS = 4;       // KL2: the last four warps (one per sub-partition) run the sum chains
constexpr int GW_CHAIN_WARP0 = GW_WARPS - GW_CHAIN_WARPS;

struct GwDev {                 // kernel parameters
    Stats st;                  // two-level frame statistics
    const float* x;            // frames (KL2 means)
    int64_t nrows;             // frames in the recording
    const double* T;           // candidate offset table, kmax entries
    int64_t kmax;
    int32_t t_exact;           // T[k] == minfeas + k * istep bit for bit: no table loads
    const int64_t* seg_a;      // chains
    const int64_t* seg_b;
    int32_t nchain;
    int32_t group_ctas;        // CTAs per group
    int32_t ngroups;
    int32_t bmax;              // speculation depth actually used (<= GW_BMAX)
    double rate, winsize, winstep, deltaws, threshold, lambda, minfeas, istep;
    int32_t metric;
    int32_t kl2_depth[3];      // KL2 speculation depth of the 1st, 2nd and later coarse waves after a change
    int32_t fuse_ok;           // KL2: the first wave after a change may carry the fine tune of window 0 (see GwPlan::fuse)
    // workspaces, per group
    double* left;              // [ngroups][kmax]
    double* right;             // [ngroups][2 parity][bmax][kmax][rterms]
    double* pooled;            // [ngroups][2 parity][GW_BMAX]
    double* fine;              // [ngroups][2 parity][3][GW_JMAX]
    // KL2 workspaces, per group
    int64_t kcap;              // candidate slots per parity
    double* kside_left;        // [ngroups][kmax][KS]
    double* kside_right;       // [ngroups][2][kcap][KS]
    double* kside_fine;        // [ngroups][2][2 sides][GW_JMAX][KS]
    float* ksum_left;          // [ngroups][kmax][VS]      running sum of [start, m_k)
    float* ksum_right;         // [ngroups][2][kcap][VS]   running sum of [m_k, end_w)
    float* ksum_fine;          // [ngroups][2][2 sides][GW_JMAX][VS]
    unsigned long long* bar;   // [ngroups] barrier counters
    int32_t* next_chain;       // queue cursor
    int32_t* group_chain;      // [ngroups][2] published next chain per group
    // output
    spkdiar_gw_window* win;
    int64_t win_cap;
    unsigned long long* nwin;  // output slots claimed (blocks of GW_SLOT_BLOCK; unused slots keep chain = -1)
    unsigned long long* nrec;  // records produced
    // split mode (abi_gw.inc: one long chain cut into sub-chains that run side by side, one CTA each):
    // a sub-chain publishes the absolute position of every change it detects and stops as soon as one
    // of them is also a change of a LATER sub-chain - from an identical `start` on, two searches are
    // identical (CD:263-268: after a change the whole state is a function of `start`) -, or when it
    // passes its give-up limits.  chg == nullptr: plain mode.
    const double* start0;      // [nchain] initial chain-relative start (0 or a half frame)
    const double* stop;        // [nchain][2] give up after a change at start >= stop[0] / a window end >= stop[1]
    const int32_t* sync;       // [nchain][3] own slot, first and last + 1 slot to compare with
    const double* slot_pos0;   // [nslot] absolute start position of the slot's sub-chain (increasing within a range)
    double* chg;               // [nslot][chg_cap] published absolute change positions
    int32_t* nchg;             // [nslot]
    int32_t chg_cap;
    int32_t* reason;           // [nchain] 0: the search ended as the reference's does, 1: synced, 2: gave up
    long long* trace;          // optional per-wave trace of CTA 0 (KL2)
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
    int kmaxw;                 // largest candidate count of the batch
    int pend_bk;               // coarse maximum of the window waiting for its fine tune
    int left_valid;            // offsets k < left_valid have a cached left term / left side
    int side_next;             // KL2: next unclaimed side task of this CTA's chunk
    // KL2, FUSED first wave after a change: besides the coarse candidates of the batch the wave scores every
    // frame a fine tune of WINDOW 0 could look at (fi[q] = frame mlo + q, nJ = their number), so that a change
    // found in window 0 - 93 % of the changes of the reference's KL2 on the 1-hour bench recording - is fine
    // tuned by the deciding warp from values already there and needs no fine wave of its own
    int fuse;
    int nJ0;                   // fine candidates per coarse maximum (CD:235-251)
    long long mlo;             // chain-relative frame of fi[0]
    double fbd;                // best fine candidate of window 0 (strictly above nothing yet: GW_NEG_INIT)
    int fbj, fninf;
    double start;              // window start of the chain (fp64, chain-relative)
    double pl[GW_BMAX];        // BIC: pooled term of window w
    // KL2 sum chains that survive from the previous coarse wave of the same `start`
    int chain_valid;           // offsets k < chain_valid have a running right sum ...
    int chain_base;            // ... in slot chain_base + k of parity chain_parity ...
    int chain_parity;
    long long chain_row;       // ... that ends at this frame row
    double e[GW_BMAX];         // window ends
    double ws_after[GW_BMAX];  // growth state after a negative window w
    double dws_after[GW_BMAX];
    double e_after[GW_BMAX];
    int last[GW_BMAX];         // a negative window w ends the chain
    int K[GW_BMAX];            // coarse candidates of window w
    int sec[GW_BMAX + 1];      // BIC/GLR: task offsets of the right sections; KL2: candidate slot offsets
    double fi[GW_JMAX];        // fine offsets i_j
    double pend_pl;            // pooled term of the window waiting for its fine tune
    // decision scratch
    double bd[GW_BMAX];
    int bk[GW_BMAX];
    int ninf[GW_BMAX];
};

// per-warp factorisation scratch of the KL2 kernel
struct GwKl2Warp {
    LdlScratch w;
    union {
        double rec[REC];
        double Lsm[(D39 * (D39 - 1)) / 2 + 3];
    };
    double pinv[VS];
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

__device__ __forceinline__ double gw_Tp(const GwDev* g, int64_t k) {      // see gw_T below
    return g->t_exact ? __dadd_rn(g->minfeas, __dmul_rn((double)k, g->istep)) : __ldg(g->T + k);
}

// ---- KL2: sum chains ---------------------------------------------------------------------
// One chain = one float32 running sum over consecutive frame rows, lane j owning
// dimensions j and j + 32, with SNAPSHOTS of the sum stored at given rows.
//   kind 0  coarse right chain of offset k = a: rows from m_k (or from where the previous
//           wave stopped) up to the end of every window w >= w0 of the batch
//   kind 1  coarse left chain: continues the running sum of [start, .) through the new offsets
//   kind 2  fine right chain of fine candidate j = a: rows [m_j, end)
//   kind 3  fine left chain: from the nearest cached coarse offset through the fine offsets
// Chain ct of a wave belongs to CTA ct % group_ctas; a CTA runs its chains four at a time,
// one per chain warp, over ONE stream of rows: one thread feeds a shared-memory ring with bulk copies
// (TMA) of GW_RING_ROWS rows, every warp adds the rows of its own range from the ring.
struct GwChain { int kind, a, w0, nsnap; long long pos; const float* init; };

struct GwChainCtx {
    const GwDev* g; const GwPlan* plan;
    int64_t base; double start;
    float* sum_left; float* sum_right; float* sum_fine;     // this group's arrays
};

__device__ __forceinline__ long long gw_snap_row(const GwChainCtx& c, const GwChain& ch, int i) {
    switch (ch.kind) {
        case 0:  return c.base + (int64_t)c.plan->e[ch.w0 + i];
        case 1:  return c.base + (int64_t)(c.start + gw_Tp(c.g, c.plan->k0 + i));
        case 2:  return c.base + (int64_t)c.plan->e[0];
        default: return c.base + (int64_t)(c.start + c.plan->fi[i]);
    }
}
__device__ __forceinline__ float* gw_snap_dst(const GwChainCtx& c, const GwChain& ch, int i) {
    const int parity = c.plan->parity;
    switch (ch.kind) {
        case 0:  return c.sum_right + ((int64_t)parity * c.g->kcap + c.plan->sec[ch.w0 + i] + ch.a) * VS;
        case 1:  return c.sum_left + (int64_t)(c.plan->k0 + i) * VS;
        case 2:  return c.sum_fine + ((int64_t)(parity * 2 + 1) * GW_JMAX + ch.a) * VS;
        default: return c.sum_fine + ((int64_t)(parity * 2 + 0) * GW_JMAX + i) * VS;
    }
}

// One round of the sum pass: each of the GW_CHAIN_WARPS chain warps of the CTA runs (at
// most) one chain; they synchronise among themselves on named barrier 1, the other warps
// of the CTA factorise meanwhile.  `range` is a shared-memory scratch of 2 * GW_CHAIN_WARPS
// long longs, `phase` the mbarrier phase bit of every ring stage (identical in all chain
// threads, lives as long as the kernel).  cw = index of the calling chain warp.
__device__ __forceinline__ void gw_chain_sync() {
    asm volatile("bar.sync 1, %0;" ::"n"(GW_CHAIN_WARPS * 32) : "memory");
}
__device__ __forceinline__ void gw_sum_round(const GwChainCtx& c, const GwChain& ch, bool active,
                                             long long* range, float* ring, uint64_t* bars, uint32_t& phase,
                                             int cw, int lane, long long& dbg_rows, long long& dbg_wait) {
    const float* __restrict__ x = c.g->x;
    const bool second = lane + 32 < D39;
    float s0 = 0.f, s1 = 0.f;
    int si = 0;
    // Snapshot rows, relative to ch.pos: lane l computes the row of snapshot 32 b + l once per batch b,
    // the chain takes them by shuffle, always ONE snapshot ahead, so that neither the offset table nor
    // the double -> integer conversion sits between two segments of the serial add chain.
    int relbatch = -1, myrel = 0;
    auto fetch = [&](int i) -> int {
        if (i >= ch.nsnap) return INT_MAX;
        if ((i >> 5) != relbatch) {
            relbatch = i >> 5;
            const int ii = 32 * relbatch + lane;
            long long row = 0;
            if (ii < ch.nsnap)
                row = ch.kind == 1 ? c.base + (int64_t)(c.start + gw_Tp(c.g, (int64_t)c.plan->k0 + ii))
                                   : gw_snap_row(c, ch, ii);
            myrel = ii < ch.nsnap ? (int)(row - ch.pos) : INT_MAX;
        }
        return __shfl_sync(0xffffffffu, myrel, i & 31);
    };
    int snap_cur = INT_MAX, snap_nxt = INT_MAX;      // rows relative to ch.pos
    long long rb_abs = 0;
    auto store_snap = [&]() {
        float* dst = gw_snap_dst(c, ch, si);
        dst[lane] = s0;
        if (second) dst[lane + 32] = s1;
        ++si;
        snap_cur = snap_nxt;
        snap_nxt = fetch(si + 1);
    };
    if (active) {
        if (ch.init) {                               // written by another CTA one wave ago: L2
            s0 = __ldcg(ch.init + lane);
            if (second) s1 = __ldcg(ch.init + 32 + lane);
        }
        rb_abs = ch.pos + fetch(ch.nsnap - 1);       // last row + 1
        snap_cur = fetch(0);
        snap_nxt = fetch(1);
        while (si < ch.nsnap && snap_cur <= 0) store_snap();
        if (si >= ch.nsnap) active = false;
    }
    if (lane == 0) {
        range[2 * cw] = active ? ch.pos : LLONG_MAX;
        range[2 * cw + 1] = active ? rb_abs : LLONG_MIN;
    }
    gw_chain_sync();
    long long A = LLONG_MAX, B = LLONG_MIN;
#pragma unroll
    for (int w = 0; w < GW_CHAIN_WARPS; ++w) {
        A = range[2 * w] < A ? range[2 * w] : A;
        B = range[2 * w + 1] > B ? range[2 * w + 1] : B;
    }
    gw_chain_sync();                                 // `range` may be rewritten by the next round
    if (!(A < B)) return;
    A &= ~3LL;                                       // 16-byte aligned start of the stream
    const int nst = (int)((B - A + GW_RING_ROWS - 1) / GW_RING_ROWS);
    const long long file_bytes = c.g->nrows * (long long)(D39 * sizeof(float));
    auto stage_bytes = [&](int i) {
        long long nb = file_bytes - (A + (long long)i * GW_RING_ROWS) * (long long)(D39 * sizeof(float));
        if (nb > (long long)(GW_STAGE_FLOATS * sizeof(float))) nb = GW_STAGE_FLOATS * sizeof(float);
        return (unsigned)(nb & ~15LL);
    };
    const bool issuer = cw == 0 && lane == 0;
    auto issue = [&](int i) {                        // one thread only
        const int s = i % GW_RING_STAGES;
        const unsigned nb = stage_bytes(i);
        mbar_expect_tx(bars + s, nb);
        bulk_g2s(ring + s * GW_STAGE_FLOATS, x + (A + (long long)i * GW_RING_ROWS) * D39, nb, bars + s);
    };
    if (issuer)
        for (int p = 0; p < GW_RING_STAGES && p < nst; ++p) issue(p);
    // positions relative to A from here on (a batch spans far less than 2^31 rows)
    const int org = active ? (int)(ch.pos - A) : 0;          // snapshot rows are relative to ch.pos
    int pos = org;
    const int rb = active ? (int)(rb_abs - A) : 0;
    if (active) dbg_rows += rb - pos;
    for (int i = 0; i < nst; ++i) {
        const int s = i % GW_RING_STAGES;
        const long long w0c = clock64();
        mbar_wait(bars + s, (phase >> s) & 1u);
        phase ^= 1u << s;
        dbg_wait += clock64() - w0c;
        const int r0 = i * GW_RING_ROWS;
        int hi = r0 + GW_RING_ROWS;
        hi = rb < hi ? rb : hi;
        if (active && pos < hi) {
            const int safe = r0 + (int)(stage_bytes(i) / (D39 * sizeof(float)));     // rows fully in the ring
            const float* buf = ring + s * GW_STAGE_FLOATS + lane - r0 * D39;         // buf[r * D39] = row r
            int r = pos > r0 ? pos : r0;
            for (;;) {
                const int snap = snap_cur == INT_MAX ? INT_MAX : snap_cur + org;
                if (r == snap) { store_snap(); continue; }       // the sum of the rows below `snap`
                if (r >= hi) break;
                const int seg = snap < hi ? snap : hi;
                const int sseg = seg < safe ? seg : safe;
                if (r + 8 <= sseg) {                             // 8 rows at a time; the loads of the next 8 rows
                    float ua[8], va[8], ub[8], vb[8];            // are in flight while the serial add chain runs
                    auto load8 = [&](float (&u)[8], float (&v)[8], int rr) {
                        const float* row = buf + rr * D39;
#pragma unroll
                        for (int q = 0; q < 8; ++q) { u[q] = row[q * D39]; v[q] = second ? row[q * D39 + 32] : 0.f; }
                    };
                    auto add8 = [&](const float (&u)[8], const float (&v)[8]) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) { s0 = __fadd_rn(s0, u[q]); s1 = __fadd_rn(s1, v[q]); }
                    };
                    load8(ua, va, r);
                    for (;;) {
                        bool more = r + 16 <= sseg;
                        if (more) load8(ub, vb, r + 8);
                        add8(ua, va);
                        r += 8;
                        if (!more) break;
                        more = r + 16 <= sseg;
                        if (more) load8(ua, va, r + 8);
                        add8(ub, vb);
                        r += 8;
                        if (!more) break;
                    }
                }
                for (; r < sseg; ++r) {
                    s0 = __fadd_rn(s0, buf[r * D39]);
                    if (second) s1 = __fadd_rn(s1, buf[r * D39 + 32]);
                }
                for (; r < seg; ++r) {                           // the last rows of the file: not 16-byte copyable
                    const float* row = x + (A + r) * D39 + lane;
                    s0 = __fadd_rn(s0, __ldg(row));
                    if (second) s1 = __fadd_rn(s1, __ldg(row + 32));
                }
            }
            pos = hi;
        }
        gw_chain_sync();                                         // stage i has been consumed by every chain warp
        if (issuer && i + GW_RING_STAGES < nst) issue(i + GW_RING_STAGES);
    }
}

// chain ct of the current wave (see the kinds above)
__device__ __forceinline__ GwChain gw_make_chain(const GwChainCtx& c, int ct, int nright, int left_valid,
                                                 int64_t s0, bool fine) {
    const GwDev& g = *c.g;
    const GwPlan& plan = *c.plan;
    GwChain ch;
    if (!fine && ct < nright) {                            // coarse right chain of offset k
        const int k = ct;
        ch.kind = 0; ch.a = k;
        int w0 = 0;
        while (w0 < plan.nW && plan.K[w0] <= k) ++w0;
        ch.w0 = w0; ch.nsnap = plan.nW - w0;
        if (k < plan.chain_valid) {
            ch.pos = plan.chain_row;
            ch.init = c.sum_right + ((int64_t)plan.chain_parity * g.kcap + plan.chain_base + k) * VS;
        } else {
            ch.pos = c.base + (int64_t)(c.start + gw_Tp(&g, k));
            ch.init = nullptr;
        }
    } else if (!fine) {                                     // coarse left chain
        ch.kind = 1; ch.a = 0; ch.w0 = 0; ch.nsnap = plan.nL;
        if (plan.k0 > 0) {
            ch.pos = c.base + (int64_t)(c.start + gw_Tp(&g, plan.k0 - 1));
            ch.init = c.sum_left + (int64_t)(plan.k0 - 1) * VS;
        } else {
            ch.pos = s0; ch.init = nullptr;
        }
    } else if (ct < nright) {                               // fine right chain
        ch.kind = 2; ch.a = ct; ch.w0 = 0; ch.nsnap = 1;
        ch.pos = c.base + (int64_t)(c.start + plan.fi[ct]);
        ch.init = nullptr;
    } else {                                                // fine left chain
        ch.kind = 3; ch.a = 0; ch.w0 = 0; ch.nsnap = plan.nJ;
        const long long m0 = c.base + (int64_t)(c.start + plan.fi[0]);
        int kk = plan.pend_bk - 1;                          // nearest cached offset at or below m0
        if (kk >= left_valid) kk = left_valid - 1;
        while (kk + 1 < left_valid && c.base + (int64_t)(c.start + gw_Tp(&g, kk + 1)) <= m0) ++kk;
        while (kk >= 0 && c.base + (int64_t)(c.start + gw_Tp(&g, kk)) > m0) --kk;
        if (kk >= 0) {
            ch.pos = c.base + (int64_t)(c.start + gw_Tp(&g, kk));
            ch.init = c.sum_left + (int64_t)kk * VS;
        } else {
            ch.pos = s0; ch.init = nullptr;
        }
    }
    return ch;
}

// candidate offset k: the host table, or its closed form when the host found the
// table to BE minfeas + k * istep exactly (dyadic frame rates such as 100 or 125 fps)
__device__ __forceinline__ double gw_T(const GwDev& g, int64_t k) {
    return g.t_exact ? __dadd_rn(g.minfeas, __dmul_rn((double)k, g.istep)) : __ldg(g.T + k);
}
// count of k with T[k] < lim (T strictly increasing), from an arithmetic guess
__device__ __forceinline__ int gw_count_below(const GwDev& g, double lim) {
    double q = (lim - g.minfeas) / g.istep;
    int64_t k = q > 0.0 ? (int64_t)q : 0;
    if (k > g.kmax) k = g.kmax;
    while (k < g.kmax && gw_T(g, k) < lim) ++k;
    while (k > 0 && !(gw_T(g, k - 1) < lim)) --k;
    return (int)k;
}

constexpr int GW_SLOT_BLOCK = 64;       // window records are claimed 64 at a time

// thread-0 state of one chain (the reference's loop variables, CD:189-200)
struct GwState {
    double start, end, ws, dws;
    int left_valid, seq, coarse_waves;
    bool done, want_fine;
    double pend_e, pend_maxi, pend_maxd;
    int pend_ncand, pend_ninf;
};

template <bool KL2, bool SPLIT>
static __global__ void __launch_bounds__(GW_THREADS, 1) gw_kernel(const GwDev g) {
    extern __shared__ __align__(16) unsigned char gw_smem[];
    GwPlan& plan = *reinterpret_cast<GwPlan*>(gw_smem);
    unsigned char* scratch_base = gw_smem + ((sizeof(GwPlan) + 15) & ~(size_t)15);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = blockIdx.x / g.group_ctas;
    const int rank = blockIdx.x - group * g.group_ctas;      // CTA rank in its group
    const int gwarps = g.group_ctas * GW_WARPS;
    const bool is_bic = !KL2 && g.metric == SPKDIAR_BIC;
    const int rterms = KL2 ? 1 : (g.metric == SPKDIAR_GLR ? 2 : 1);
    double* left = g.left + (int64_t)group * g.kmax;
    double* right = g.right + (int64_t)group * 2 * g.bmax * g.kmax * rterms;
    double* pooled = g.pooled + (int64_t)group * 2 * GW_BMAX;
    double* fine = g.fine + (int64_t)group * 2 * 3 * GW_JMAX;
    // KL2 arrays of this group
    double* kside_left = KL2 ? g.kside_left + (int64_t)group * g.kmax * KS : nullptr;
    double* kside_right = KL2 ? g.kside_right + (int64_t)group * 2 * g.kcap * KS : nullptr;
    double* kside_fine = KL2 ? g.kside_fine + (int64_t)group * 4 * GW_JMAX * KS : nullptr;
    float* ksum_left = KL2 ? g.ksum_left + (int64_t)group * g.kmax * VS : nullptr;
    float* ksum_right = KL2 ? g.ksum_right + (int64_t)group * 2 * g.kcap * VS : nullptr;
    float* ksum_fine = KL2 ? g.ksum_fine + (int64_t)group * 4 * GW_JMAX * VS : nullptr;
    GwKl2Warp* kwarps = reinterpret_cast<GwKl2Warp*>(scratch_base);
    float* ring = reinterpret_cast<float*>(scratch_base + GW_WARPS * sizeof(GwKl2Warp));
    uint64_t* ring_bar = reinterpret_cast<uint64_t*>(ring + GW_RING_STAGES * GW_STAGE_FLOATS);
    long long* range = reinterpret_cast<long long*>(ring_bar + GW_RING_STAGES);
    uint32_t ring_phase = 0;
    if (KL2) {
        if (threadIdx.x == 0) {
            for (int s = 0; s < GW_RING_STAGES; ++s) mbar_init(ring_bar + s, 1);
            mbar_fence_init();
        }
        __syncthreads();
    }
    unsigned long long* bar = g.bar + group;
    unsigned long long bar_target = 0;
    int wave = 0;               // parity source for the double-buffered term arrays
    int chain_pub = 0;          // parity of the published next-chain slot
    unsigned long long slot_next = 0, slot_end = 0;     // claimed output slots (rank 0, thread 0)

    long long t_plan = 0, t_eval = 0, t_bar = 0, t_dec = 0, n_wave = 0, n_task = 0, t_e1 = 0, t_e2 = 0;
    long long n_rows = 0, t_wait = 0, t_last_chain = 0;
    long long t_chain = 0;      // KL2: cycles of warp 0 in the sum chains / of warp 1 in the sides

    // ---- warp 0: describe the next wave in `plan` from the chain state.  The state is kept
    // identical in all 32 lanes; lane w works on window w of the batch, the batch cut and the
    // task offsets are warp scans - no serial walk over shared memory. ----
    auto plan_wave = [&](GwState& st, double n) {
        if (lane == 0) {
            plan.parity = wave & 1;
            plan.side_next = 0;
            plan.start = st.start;
            plan.left_valid = st.left_valid;
            plan.mode = st.done ? 2 : (st.want_fine ? 1 : 0);
        }
        if (st.done) return;
        if (st.want_fine) {
            // fine-tune offsets, CD:235-251: i = maxi - istep; while i < maxi + istep: ...; i += 1
            double i = st.pend_maxi - g.istep;
            const double endtune = st.pend_maxi + g.istep;
            int nJ = 0;
            while (i < endtune && nJ < GW_JMAX) { if ((nJ & 31) == lane) plan.fi[nJ] = i; ++nJ; i += 1; }
            if (lane == 0) {
                plan.fuse = 0;
                plan.nJ = nJ;
                plan.nW = 1;
                plan.e[0] = st.pend_e;
                plan.ntask = nJ * (KL2 ? 2 : (g.metric == SPKDIAR_GLR ? 3 : 2));
            }
            return;
        }
        // the growth schedule of `end` while no change is found, CD:273-284 (every lane, in registers)
        double e = st.end, w_ = st.ws, dw = st.dws;
        double my_e = 0.0, my_eaft = 0.0, my_wsa = 0.0, my_dwsa = 0.0;
        int my_last = 0, nW = 0;
        // KL2 speculates less right after a change: the sum chains of a batch cost rows x offsets,
        // i.e. grow with the square of the depth, and most changes show within a few windows
        int depth = g.bmax;
        if (KL2 && g.bmax > 1) depth = st.coarse_waves == 0 ? g.kl2_depth[0] : (st.coarse_waves == 1 ? g.kl2_depth[1] : g.kl2_depth[2]);
        for (int w = 0; w < depth; ++w) {
            if (lane == w) my_e = e;
            nW = w + 1;
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
            if (lane == w) { my_eaft = e; my_wsa = w_; my_dwsa = dw; my_last = last; }
            if (last) break;
        }
        const bool in = lane < nW;
        const int Kw = in ? gw_count_below(g, my_e - st.start - g.minfeas) : 0;                 // CD:204
        int prevK = __shfl_up_sync(0xffffffffu, Kw, 1);
        if (lane == 0 || prevK < st.left_valid) prevK = st.left_valid;
        const int newl = Kw > prevK ? Kw - prevK : 0;
        const int unit = in ? Kw * rterms : 0;
        int tsc = in ? unit + newl + (is_bic ? 1 : 0) : 0, ksc = unit;     // inclusive scans: tasks, right terms
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int tv = __shfl_up_sync(0xffffffffu, tsc, o), kv = __shfl_up_sync(0xffffffffu, ksc, o);
            if (lane >= o) { tsc += tv; ksc += kv; }
        }
        // cut the batch to the group's one-round capacity (always keep window 0)
        const unsigned over = __ballot_sync(0xffffffffu, in && lane > 0 && tsc > gwarps);
        const int nWc = over ? __ffs(over) - 1 : nW;
        const int klast = __shfl_sync(0xffffffffu, Kw, nWc - 1);
        const int kmaxw = klast > st.left_valid ? klast : st.left_valid;
        const int nL = kmaxw - st.left_valid;
        const int off0 = KL2 ? 0 : nL + (is_bic ? nWc : 0);
        if (lane < nWc) {
            plan.e[lane] = my_e; plan.e_after[lane] = my_eaft; plan.ws_after[lane] = my_wsa;
            plan.dws_after[lane] = my_dwsa; plan.last[lane] = my_last; plan.K[lane] = Kw;
            plan.sec[lane] = off0 + ksc - unit;
            if (lane == nWc - 1) { plan.sec[nWc] = off0 + ksc; plan.ntask = (KL2 ? nL : 0) + off0 + ksc; }
        }
        if (lane == 0) { plan.nW = nWc; plan.k0 = st.left_valid; plan.nL = nL; plan.kmaxw = kmaxw; plan.fuse = 0; }
        if (KL2 && g.fuse_ok && st.coarse_waves == 0 && st.left_valid == 0) {
            const int K0 = __shfl_sync(0xffffffffu, Kw, 0);
            int nJ0 = 0;                                                    // CD:235-251 for the first offset
            for (double i = gw_T(g, 0) - g.istep; i < gw_T(g, 0) + g.istep && nJ0 < GW_JMAX; i += 1) ++nJ0;
            if (K0 > 0 && nJ0 > 0) {
                const double lo = gw_T(g, 0) - g.istep;
                const double hi = (gw_T(g, K0 - 1) - g.istep) + (double)(nJ0 - 1);
                const long long mlo = (long long)(st.start + lo), mhi = (long long)(st.start + hi);
                const long long nQ = mhi - mlo + 1;
                const long long e0 = (long long)__shfl_sync(0xffffffffu, my_e, 0);
                if (nQ <= GW_JMAX && mlo > (long long)st.start && mhi < e0) {
                    for (int q = lane; q < (int)nQ; q += 32) plan.fi[q] = (double)(mlo + q) - st.start;
                    __syncwarp();
                    if (lane == 0) {
                        plan.fuse = 1; plan.nJ = (int)nQ; plan.nJ0 = nJ0; plan.mlo = mlo; plan.pend_bk = 0;
                        plan.ntask += 2 * (int)nQ;
                    }
                }
            }
        }
    };
    // claim `cnt` consecutive output slots (warp 0; every lane keeps the same cursor)
    auto claim_slots = [&](int cnt) -> unsigned long long {
        if (slot_end - slot_next < (unsigned long long)cnt) {
            unsigned long long b0 = 0;
            if (lane == 0) b0 = atomicAdd(g.nwin, (unsigned long long)GW_SLOT_BLOCK);
            slot_next = __shfl_sync(0xffffffffu, b0, 0);
            slot_end = slot_next + GW_SLOT_BLOCK;
        }
        const unsigned long long s0 = slot_next;
        slot_next += cnt;
        return s0;
    };

    int chain = group;
    while (chain < g.nchain) {
        const int64_t base = g.seg_a[chain];
        const int64_t nfr = g.seg_b[chain] - base;
        const double n = (double)nfr;
        GwState st;
        st.start = SPLIT ? __ldg(g.start0 + chain) : 0.0;
        int reason = 0, nchg_mine = 0;              // split mode (warp 0, identical in all lanes)
        st.end = st.start + g.winsize * 2;
        st.ws = g.minfeas; st.dws = g.deltaws;
        st.left_valid = 0; st.seq = 0; st.coarse_waves = 0;
        st.done = !(st.end <= n);
        st.want_fine = false;
        st.pend_e = st.pend_maxi = st.pend_maxd = 0.0; st.pend_ncand = st.pend_ninf = 0;
        if (warp == 0) {
            if (lane == 0) plan.chain_valid = 0;
            plan_wave(st, n);
        }
        __syncthreads();

        while (plan.mode != 2) {
            // ================= EVALUATE =================
            const long long c1 = clock64();
            const int parity = plan.parity;
            const double start = plan.start;
            const int left_valid = plan.left_valid;
            const int64_t s0 = base + (int64_t)start;
            if (KL2) {
                const long long k0c = clock64();
                if (warp >= GW_CHAIN_WARP0) {
                    // ---- sum chains: chain ct belongs to CTA ct % group_ctas, four per round ----
                    GwChainCtx cc{&g, &plan, base, start, ksum_left, ksum_right, ksum_fine};
                    const int cw = warp - GW_CHAIN_WARP0;
                    const int nright = plan.mode == 0 ? plan.kmaxw : plan.nJ;
                    const int nct = nright + ((plan.mode == 0 ? plan.nL : plan.nJ) > 0 ? 1 : 0);
                    // a fused wave also runs the chains of a fine wave over fi[]: nJ right chains and the left one
                    const int nctf = (plan.mode == 0 && plan.fuse) ? plan.nJ + 1 : 0;
                    // four NEIGHBOURING chains per CTA (one per chain warp): their row ranges nearly coincide, so
                    // the one stream of rows the CTA's ring carries serves all four
                    for (int c0 = rank * GW_CHAIN_WARPS; c0 < nct + nctf; c0 += g.group_ctas * GW_CHAIN_WARPS) {
                        const int ct = c0 + cw;
                        GwChain ch;
                        ch.kind = 0; ch.a = 0; ch.w0 = 0; ch.nsnap = 0; ch.pos = 0; ch.init = nullptr;
                        if (ct < nct) ch = gw_make_chain(cc, ct, nright, left_valid, s0, plan.mode != 0);
                        else if (ct < nct + nctf - 1) ch = gw_make_chain(cc, ct - nct, plan.nJ, left_valid, s0, true);
                        if (nctf > 0 && ct == nct + nctf - 1) {
                            // The left sums of a fused wave: ONE running sum from the window start with a snapshot at
                            // EVERY frame from fi[0] on.  A snapshot per row is the worst case of the ring machinery
                            // (measured 900 cycles per snapshot), so this chain reads its rows straight from global
                            // memory, sixteen loads in flight, and stores the running sum after every row.
                            const float* row = g.x + s0 * D39 + lane;
                            const bool second = lane + 32 < D39;
                            const int nrow = (int)(plan.mlo - (long long)start) + plan.nJ - 1;     // rows s0 .. last fine frame
                            const int first = (int)(plan.mlo - (long long)start);                  // snapshot q after `first + q` rows
                            float* dst = ksum_fine + ((int64_t)(parity * 2 + 0) * GW_JMAX) * VS;
                            float a0 = 0.f, a1 = 0.f;
                            for (int r0 = 0; r0 < nrow; r0 += 16) {
                                float u[16], v[16];
#pragma unroll
                                for (int q = 0; q < 16; ++q) {
                                    const bool ok = r0 + q < nrow;
                                    u[q] = ok ? __ldg(row + (int64_t)(r0 + q) * D39) : 0.f;
                                    v[q] = (ok && second) ? __ldg(row + (int64_t)(r0 + q) * D39 + 32) : 0.f;
                                }
#pragma unroll
                                for (int q = 0; q < 16; ++q) {
                                    if (r0 + q < nrow) {
                                        a0 = __fadd_rn(a0, u[q]); a1 = __fadd_rn(a1, v[q]);
                                        const int sq = r0 + q + 1 - first;                          // rows summed so far = first + sq
                                        if (sq >= 0 && sq < plan.nJ) {
                                            dst[(int64_t)sq * VS + lane] = a0;
                                            if (second) dst[(int64_t)sq * VS + lane + 32] = a1;
                                        }
                                    }
                                }
                            }
                            n_rows += nrow;
                        }
                        gw_sum_round(cc, ch, ct < nct + nctf - (nctf > 0 ? 1 : 0) && ch.nsnap > 0, range, ring, ring_bar, ring_phase,
                                     cw, lane, n_rows, t_wait);
                    }
                    if (cw == 0) { t_chain += clock64() - k0c; t_last_chain = clock64() - k0c; }
                    if (g.trace && n_wave == 1 && lane == 0) g.trace[8 * 4096 + 16 * blockIdx.x + 12 + cw] = clock64() - k0c;
                }
                {
                    // ---- sides: diag(S), diag(S^-1) of one window per task.  The CTA owns a contiguous
                    // chunk of tasks (see below); its warps claim them one by one, the chain warps join
                    // when their sums are done ----
                    const long long k2c = clock64();
                    GwKl2Warp& own = kwarps[warp];
                    const int chunk = (plan.ntask + g.group_ctas - 1) / g.group_ctas;
                    const int id_end = (rank + 1) * chunk < plan.ntask ? (rank + 1) * chunk : plan.ntask;
                    for (;;) {
                        int id = 0;
                        if (lane == 0) id = rank * chunk + atomicAdd(&plan.side_next, 1);
                        id = __shfl_sync(0xffffffffu, id, 0);
                        if (id >= id_end) break;
                        int64_t ra, rb;
                        double* dst;
                        const int ntask0 = plan.mode == 0 ? plan.nL + plan.sec[plan.nW] : 0;    // the rest: fine sides
                        if (id < ntask0) {
                            if (id < plan.nL) {                                 // left side of a new offset
                                const int k = plan.k0 + id;
                                ra = s0; rb = base + (int64_t)(start + gw_T(g, k));
                                dst = kside_left + (int64_t)k * KS;
                            } else {                                            // right side of candidate slot t
                                const int t = id - plan.nL;
                                int w = 0;
                                while (t >= plan.sec[w + 1]) ++w;
                                const int k = t - plan.sec[w];
                                ra = base + (int64_t)(start + gw_T(g, k));
                                rb = base + (int64_t)plan.e[w];
                                dst = kside_right + ((int64_t)parity * g.kcap + t) * KS;
                            }
                        } else {
                            const int idf = id - ntask0;
                            const int side = idf >= plan.nJ ? 1 : 0, j = idf - side * plan.nJ;
                            const int64_t mm = base + (int64_t)(start + plan.fi[j]);
                            ra = side ? mm : s0;
                            rb = side ? base + (int64_t)plan.e[0] : mm;
                            dst = kside_fine + ((int64_t)(parity * 2 + side) * GW_JMAX + j) * KS;
                        }
                        kl2_side_one(WinSrc(g.st, ra, rb, REC), own, dst, dst + VS, lane);
                    }
                    if (warp == 0) t_chain += clock64() - k2c;
                }
                const long long k1c = clock64();
                if (g.trace && n_wave == 1 && lane == 0) g.trace[8 * 4096 + 16 * blockIdx.x + warp] = k1c - k0c;
                gw_group_barrier(bar, bar_target, g.group_ctas);
                t_e1 += k1c - k0c; t_e2 += clock64() - k1c;
                if (g.trace && blockIdx.x == 0 && threadIdx.x == 0 && n_wave < 4096) {
                    long long* tr = g.trace + 8 * n_wave;
                    tr[0] = plan.mode; tr[1] = plan.nW; tr[2] = plan.ntask; tr[3] = plan.mode == 0 ? plan.kmaxw : plan.nJ;
                    tr[4] = k1c - k0c; tr[5] = clock64() - k1c; tr[6] = (long long)(plan.e[plan.nW - 1] - start); tr[7] = t_last_chain;
                }
                // ---- distances: one warp per candidate ----
                const int ncand0 = plan.mode == 0 ? plan.sec[plan.nW] : 0;               // the rest: fine candidates
                const int ncand = ncand0 + ((plan.mode != 0 || plan.fuse) ? plan.nJ : 0);
                for (int tt = warp * g.group_ctas + rank; tt < ncand; tt += gwarps) {
                    const double *sl, *sr; const float *ml, *mr; int64_t mm, ee; double* dst;
                    const int t = tt < ncand0 ? tt : tt - ncand0;
                    if (tt < ncand0) {
                        int w = 0;
                        while (t >= plan.sec[w + 1]) ++w;
                        const int k = t - plan.sec[w];
                        mm = base + (int64_t)(start + gw_T(g, k));
                        ee = base + (int64_t)plan.e[w];
                        sl = kside_left + (int64_t)k * KS;              ml = ksum_left + (int64_t)k * VS;
                        sr = kside_right + ((int64_t)parity * g.kcap + t) * KS;
                        mr = ksum_right + ((int64_t)parity * g.kcap + t) * VS;
                        dst = right + ((int64_t)parity * g.bmax + w) * g.kmax + k;
                    } else {
                        mm = base + (int64_t)(start + plan.fi[t]);
                        ee = base + (int64_t)plan.e[0];
                        sl = kside_fine + ((int64_t)(parity * 2 + 0) * GW_JMAX + t) * KS;
                        sr = kside_fine + ((int64_t)(parity * 2 + 1) * GW_JMAX + t) * KS;
                        ml = ksum_fine + ((int64_t)(parity * 2 + 0) * GW_JMAX + t) * VS;
                        mr = ksum_fine + ((int64_t)(parity * 2 + 1) * GW_JMAX + t) * VS;
                        dst = fine + (int64_t)parity * 3 * GW_JMAX + t;
                    }
                    const double v = kl2_distance_cached(sl, sr, ml, mr, (double)(mm - s0), (double)(ee - mm), lane);
                    if (lane == 0) *dst = v;
                }
            } else {
                // A wave with fewer tasks than warps spreads over all SMs instead of filling the first
                // CTAs, in CONTIGUOUS chunks: neighbouring tasks are neighbouring offsets of one window,
                // they share the records at the window end and (13 offsets at a time) the block prefix at
                // the split, so most operand loads of a CTA's twelve warps meet in L1 instead of L2.
                const int chunk = (plan.ntask + g.group_ctas - 1) / g.group_ctas;
                const int id_end = (rank + 1) * chunk < plan.ntask ? (rank + 1) * chunk : plan.ntask;
                for (int id = rank * chunk + warp; id < id_end; id += GW_WARPS) {
                    int64_t mm = s0, ee = s0;
                    int term = 0;
                    double* dst = nullptr;
                    if (plan.mode == 0) {
                        const int npool = is_bic ? plan.nW : 0;
                        if (id < plan.nL) {                         // left term of a new coarse offset
                            const int k = plan.k0 + id;
                            mm = base + (int64_t)(start + gw_T(g, k));
                            ee = mm; term = 0; dst = left + k;
                        } else if (id < plan.nL + npool) {          // pooled term of window w (BIC)
                            const int w = id - plan.nL;
                            mm = s0; ee = base + (int64_t)plan.e[w]; term = 2;
                            dst = pooled + parity * GW_BMAX + w;
                        } else {                                    // right (and GLR mix) terms
                            int w = 0;
                            while (id >= plan.sec[w + 1]) ++w;
                            const int r = id - plan.sec[w];
                            // GLR: first the right terms of all offsets, then the (more expensive) mix terms, so
                            // that the warps of a round work on the same kind of term
                            const int sub = r >= plan.K[w] ? 1 : 0, k = r - sub * plan.K[w];
                            mm = base + (int64_t)(start + gw_T(g, k));
                            ee = base + (int64_t)plan.e[w];
                            term = sub == 0 ? 1 : 2;
                            dst = right + (((int64_t)parity * g.bmax + w) * g.kmax + k) * rterms + sub;
                        }
                    } else {
                        const int per = g.metric == SPKDIAR_GLR ? 3 : 2;
                        const int j = id / per, sub = id - j * per;
                        mm = base + (int64_t)(start + plan.fi[j]);
                        ee = base + (int64_t)plan.e[0];
                        term = sub;
                        dst = fine + ((int64_t)parity * 3 + sub) * GW_JMAX + j;
                    }
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

            // ================= DECIDE: best candidate per window, one warp per window =================
            const int nwin = plan.mode == 0 ? plan.nW : 1;
            if (warp < nwin) {
                const int w = warp;
                const bool coarse = plan.mode == 0;
                const int ncand = coarse ? plan.K[w] : plan.nJ;
                const int64_t e0 = (int64_t)plan.e[w];
                const int64_t s0r = (int64_t)start;
                const double* f0 = fine + (int64_t)parity * 3 * GW_JMAX;
                const double pl = !is_bic ? 0.0 : (coarse ? __ldcg(pooled + parity * GW_BMAX + w) : plan.pend_pl);
                const double pen = is_bic ? bic_pen(g.lambda, (double)(e0 - s0r)) : 0.0;
                double bd = GW_NEG_INIT; int bk = -1; int ninf = 0;
                for (int c0 = 0; c0 < ncand; c0 += 128) {           // four candidates per lane: loads first
                    double off[4], t0[4], t1[4], t2[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int k = c0 + 32 * u + lane;
                        off[u] = t0[u] = t1[u] = t2[u] = 0.0;
                        if (k < ncand) {
                            if (coarse) {
                                const double* rp = right + (((int64_t)parity * g.bmax + w) * g.kmax + k) * rterms;
                                off[u] = gw_T(g, k);
                                t1[u] = __ldcg(rp);
                                if (!KL2) t0[u] = __ldcg(left + k);
                                if (!KL2 && !is_bic) t2[u] = __ldcg(rp + 1);
                            } else {
                                off[u] = plan.fi[k];
                                t0[u] = __ldcg(f0 + k);
                                if (!KL2) t1[u] = __ldcg(f0 + GW_JMAX + k);
                                if (!KL2 && !is_bic) t2[u] = __ldcg(f0 + 2 * GW_JMAX + k);
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int k = c0 + 32 * u + lane;
                        if (k < ncand) {
                            const int64_t m = (int64_t)(start + off[u]);
                            const double N1 = (double)(m - s0r), N2 = (double)(e0 - m);
                            double d;
                            if (KL2) d = coarse ? t1[u] : t0[u];
                            else if (is_bic) d = bic_combine_pen(N1, N2, t0[u], t1[u], pl, pen);
                            else d = glr_combine(N1, N2, t0[u], t1[u], t2[u]);
                            if (d == d_inf() || d == -d_inf()) ++ninf;          // CD:219-220
                            else if (d > bd) { bd = d; bk = k; }                // CD:215-217 (strict, first wins)
                        }
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double od = __shfl_xor_sync(0xffffffffu, bd, o);
                    const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
                    ninf += __shfl_xor_sync(0xffffffffu, ninf, o);
                    if (ok >= 0 && (bk < 0 || od > bd || (od == bd && ok < bk))) { bd = od; bk = ok; }
                }
                if (lane == 0) { plan.bd[w] = bd; plan.bk[w] = bk; plan.ninf[w] = ninf; plan.pl[w] = pl; }
                if (KL2 && coarse && w == 0 && plan.fuse && bk >= 0 && bd > g.threshold) {
                    // fused wave, window 0 is positive: its fine tune (CD:235-251) over the dense values
                    const double i0 = gw_T(g, bk) - g.istep;
                    double fd = GW_NEG_INIT; int fj = -1; int fninf = 0;
                    for (int j = lane; j < plan.nJ0; j += 32) {
                        const long long m = (long long)(start + (i0 + (double)j));
                        const double d = __ldcg(f0 + (m - plan.mlo));
                        if (d == d_inf() || d == -d_inf()) ++fninf;
                        else if (d > fd) { fd = d; fj = j; }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const double od = __shfl_xor_sync(0xffffffffu, fd, o);
                        const int oj = __shfl_xor_sync(0xffffffffu, fj, o);
                        fninf += __shfl_xor_sync(0xffffffffu, fninf, o);
                        if (oj >= 0 && (fj < 0 || od > fd || (od == fd && oj < fj))) { fd = od; fj = oj; }
                    }
                    if (lane == 0) { plan.fbd = fd; plan.fbj = fj; plan.fninf = fninf; }
                }
            }
            __syncthreads();
            const long long c4 = clock64();

            // ================= warp 0: apply the reference's decision rules, plan the next wave =================
            if (warp == 0) {
                bool fused_pos = false;            // the change of window 0 was fine tuned inside this (fused) wave
                if (plan.mode == 0) {
                    // lane w looks at window w; windows count in order up to the first positive one
                    // (it goes to the fine tune) or the first negative one that ends the chain
                    const int nW = plan.nW;
                    const bool in = lane < nW;
                    const double bd = in ? plan.bd[lane] : 0.0;
                    const int bk = in ? plan.bk[lane] : -1;
                    const int ninfw = in ? plan.ninf[lane] : 0;
                    const int Kw = in ? plan.K[lane] : 0;
                    const double ew = in ? plan.e[lane] : 0.0, eaft = in ? plan.e_after[lane] : 0.0;
                    const double wsa = in ? plan.ws_after[lane] : 0.0, dwsa = in ? plan.dws_after[lane] : 0.0;
                    const double plw = in ? plan.pl[lane] : 0.0;
                    const double maxi = bk >= 0 ? gw_T(g, bk) : 0.0;
                    const bool pos = in && bd > g.threshold && bk >= 0;                        // CD:230
                    const unsigned mpos = __ballot_sync(0xffffffffu, pos);
                    const unsigned mlast = __ballot_sync(0xffffffffu, in && plan.last[in ? lane : 0] != 0);
                    const int fpos = mpos ? __ffs(mpos) - 1 : 32, flast = mlast ? __ffs(mlast) - 1 : 32;
                    const bool has_pos = fpos < nW && fpos <= flast;
                    const int nneg = has_pos ? fpos : (flast + 1 < nW ? flast + 1 : nW);
                    const bool fin = !has_pos && flast < nW;
                    if (rank == 0 && nneg > 0) {                                               // negative window records
                        const unsigned long long sb = claim_slots(nneg);
                        if (lane < nneg && (int64_t)(sb + lane) < g.win_cap) {
                            spkdiar_gw_window r;
                            r.start = st.start; r.end = ew; r.maxi = maxi; r.maxd = bd;
                            r.maxi_fine = 0.0; r.maxd_fine = 0.0; r.positive = 0; r.chain = chain;
                            r.ncand = bk >= 0 ? Kw : -Kw - 1;
                            r.ninf = ninfw; r.seq = st.seq + lane; r.pad = 0;
                            g.win[sb + lane] = r;
                        }
                    }
                    st.seq += nneg;
                    if (nneg > 0) {
                        st.end = __shfl_sync(0xffffffffu, eaft, nneg - 1);
                        st.ws = __shfl_sync(0xffffffffu, wsa, nneg - 1);
                        st.dws = __shfl_sync(0xffffffffu, dwsa, nneg - 1);
                    }
                    if (has_pos) {
                        st.want_fine = true;
                        st.pend_e = __shfl_sync(0xffffffffu, ew, fpos);
                        st.pend_maxi = __shfl_sync(0xffffffffu, maxi, fpos);
                        st.pend_maxd = __shfl_sync(0xffffffffu, bd, fpos);
                        st.pend_ncand = __shfl_sync(0xffffffffu, Kw, fpos);
                        st.pend_ninf = __shfl_sync(0xffffffffu, ninfw, fpos);
                        st.end = st.pend_e;
                        if (lane == fpos) { plan.pend_bk = bk; plan.pend_pl = plw; }
                        fused_pos = KL2 && plan.fuse && fpos == 0;
                    }
                    if (fin) st.done = true;
                    if (SPLIT && !st.done && !has_pos && st.end >= __ldg(g.stop + 2 * chain + 1)) { st.done = true; reason = 2; }
                    // every offset of the batch now has its left term / left side, and (KL2) a
                    // running right sum that ends at the last window of the batch
                    st.left_valid = plan.kmaxw;
                    ++st.coarse_waves;
                    if (KL2 && lane == 0) {
                        plan.chain_valid = plan.kmaxw;
                        plan.chain_base = plan.sec[nW - 1];
                        plan.chain_parity = parity;
                        plan.chain_row = base + (int64_t)plan.e[nW - 1];
                    }
                }
                if (plan.mode != 0 || fused_pos) {
                    // fine-tune decision, CD:237-251: strict improvement over the coarse maximum, first wins
                    __syncwarp();
                    double maxd = st.pend_maxd, maxi = st.pend_maxi;
                    const int fbest = fused_pos ? plan.fbj : plan.bk[0];
                    const double fbd = fused_pos ? plan.fbd : plan.bd[0];
                    const int fninf = fused_pos ? plan.fninf : plan.ninf[0];
                    if (fbest >= 0 && fbd > st.pend_maxd) {
                        maxd = fbd;
                        maxi = fused_pos ? (st.pend_maxi - g.istep) + (double)fbest : plan.fi[fbest];
                    }
                    if (rank == 0) {
                        const unsigned long long sb = claim_slots(1);
                        if (lane == 0 && (int64_t)sb < g.win_cap) {
                            spkdiar_gw_window r;
                            r.start = st.start; r.end = st.pend_e; r.maxi = st.pend_maxi; r.maxd = st.pend_maxd;
                            r.maxi_fine = maxi; r.maxd_fine = maxd; r.positive = 1; r.chain = chain;
                            r.ncand = st.pend_ncand; r.ninf = st.pend_ninf + fninf; r.seq = st.seq; r.pad = 0;
                            g.win[sb] = r;
                        }
                    }
                    ++st.seq;
                    st.want_fine = false;
                    st.coarse_waves = 0;
                    st.left_valid = 0;                                  // CD:256
                    if (lane == 0) plan.chain_valid = 0;
                    st.start += maxi;                                   // CD:263
                    if (st.start + g.winsize * 2 <= n) {                // CD:264-268
                        st.end = st.start + g.winsize * 2;
                        st.ws = g.minfeas; st.dws = g.deltaws;
                    } else {
                        st.done = true;                                 // CD:269-270
                    }
                    if (SPLIT && !st.done) {
                        const double apos = (double)base + st.start;
                        const int slot = __ldg(g.sync + 3 * chain), lo = __ldg(g.sync + 3 * chain + 1), hi = __ldg(g.sync + 3 * chain + 2);
                        double* mine = g.chg + (size_t)slot * g.chg_cap;
                        if (nchg_mine < g.chg_cap && lane == 0) {
                            mine[nchg_mine] = apos;
                            __threadfence();
                            *((volatile int32_t*)(g.nchg + slot)) = nchg_mine + 1;
                        }
                        ++nchg_mine;
                        __syncwarp();
                        const int nmine = nchg_mine < g.chg_cap ? nchg_mine : g.chg_cap;
                        bool hit = false;
                        for (int sl = lo; sl < hi && !hit; ++sl) {
                            const double p0 = __ldg(g.slot_pos0 + sl);
                            if (!(p0 <= apos)) break;                    // later sub-chains start beyond this change
                            int nn = *((volatile int32_t*)(g.nchg + sl));
                            nn = nn < g.chg_cap ? nn : g.chg_cap;
                            __threadfence();
                            int m0 = 0;                                  // my changes at or beyond that sub-chain's start
                            while (m0 < nmine && __ldcg(mine + m0) < p0) ++m0;
                            for (int i = lane; i < nn; i += 32) {
                                const double v = __ldcg(g.chg + (size_t)sl * g.chg_cap + i);
                                for (int m = m0; m < nmine; ++m) hit |= __ldcg(mine + m) == v;
                            }
                            hit = __any_sync(0xffffffffu, hit);
                        }
                        if (hit) { st.done = true; reason = 1; }
                        else if (st.start >= __ldg(g.stop + 2 * chain)) { st.done = true; reason = 2; }
                    }
                }
                __syncwarp();                                           // reads of the old plan before it is rewritten
                plan_wave(st, n);
            }
            __syncthreads();
            t_eval += c2 - c1; t_bar += c3 - c2; t_dec += c4 - c3; t_plan += clock64() - c4; ++n_wave; n_task += plan.ntask;
        }
        if (rank == 0 && threadIdx.x == 0) {
            atomicAdd(g.nrec, (unsigned long long)st.seq);
            if (SPLIT) g.reason[chain] = reason;
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
    if (KL2 && g.dbg && lane == 0 && (warp == 0 || warp == GW_CHAIN_WARP0)) g.dbg[8 + 4 * blockIdx.x + (warp == 0 ? 1 : 0)] = t_chain;
    if (KL2 && g.dbg && lane == 0 && warp == GW_CHAIN_WARP0) { g.dbg[8 + 4 * blockIdx.x + 2] = n_rows; g.dbg[8 + 4 * blockIdx.x + 3] = t_wait; }
    if (g.dbg && blockIdx.x == 0 && threadIdx.x == 0) {
        g.dbg[0] = t_plan; g.dbg[1] = t_eval; g.dbg[2] = t_bar; g.dbg[3] = t_dec; g.dbg[4] = n_wave; g.dbg[5] = n_task; g.dbg[6] = t_e1; g.dbg[7] = t_e2;
    }
}

inline size_t gw_smem_bytes(bool kl2) {
    const size_t plan = (sizeof(GwPlan) + 15) & ~(size_t)15;
    if (!kl2) return plan + GW_WARPS * sizeof(WarpScratch);
    return plan + GW_WARPS * sizeof(GwKl2Warp) + sizeof(float) * GW_RING_STAGES * GW_STAGE_FLOATS
           + sizeof(uint64_t) * GW_RING_STAGES + sizeof(long long) * 2 * GW_CHAIN_WARPS;
}

cudaError_t gw_set_dim(int d) { return set_dim_symbol(d); }

cudaError_t gw_configure() {
    cudaError_t e = cudaFuncSetAttribute(gw_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)gw_smem_bytes(false));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(gw_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gw_smem_bytes(false));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(gw_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gw_smem_bytes(true));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(gw_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)gw_smem_bytes(true));
}

}  // namespace spk
