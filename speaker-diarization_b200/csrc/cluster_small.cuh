// cluster_small.cuh - K7 for up to ~2,000 clusters: the merge loop with an OWNER WARP per cluster.
//
// Same results as cl_merge_loop (cluster.cuh), bit for bit - same records, same device functions for the
// log-determinants and the distances, the same ndarray.argmin order (spk-clustering.py:203-237,
// spk-clustering2.py:187-215) - but organised around what a merge costs when there are only a few hundred
// clusters: latency.  The general engine deals the pairs of a rescoring to whatever warp comes next, so per
// merge every warp stages a cluster record from L2, the cached row minima live in global memory and the
// loop needs two grid barriers; its 23 us per merge at 361 clusters are mostly memory round trips.  Here
//
//   * every cluster k belongs to ONE warp for the whole run (k % owners picks the warp, k / owners its
//     slot); the warp keeps the cluster's 820-double record, its ln|S_k| and the minimum of row k of the pair
//     matrix in shared memory.  A rescoring then stages nothing: the merged record X (rec[a] + rec[b], formed
//     once per CTA) and the own record are both on chip, the warp forms the pooled matrix and factorises it;
//   * the last warp of every CTA owns nothing: it factorises the merged cluster itself (ln|S_ab|, needed by
//     every pair) while the owners factorise their pooled matrices;
//   * there is NO grid barrier.  After a merge every owner PUBLISHES one 24-byte entry per cluster - the new
//     minimum of its row, its distance to the merged cluster, and a tag (merge number | column) stored with
//     release semantics.  At the top of the next merge every thread of every CTA spins (acquire) on the tags of
//     the rows it looks at until they carry the current merge number, folds (row minimum) and (entry of the
//     merged row = the published distance) into its candidate, and the CTA reduces to the argmin: same data,
//     same code, same decision in every CTA - no broadcast.  The wait is for exactly the data that is needed;
//     entries live in three buffers (merge number mod 3), which is enough because nobody can publish merge
//     m + 2 before everybody has published merge m + 1, i.e. has finished reading merge m;
//   * cluster records are double-buffered by a version bit per cluster, so the owner of `a` can store the
//     merged record while slower CTAs still read the old one (the owner publishes a tag for the merged row as
//     well: whoever has seen it sees the record);
//   * a CTA none of whose clusters is alive any more leaves the kernel (CTA 0 owns cluster 0, which never
//     dies: the survivor of a merge is the smaller index).
//
// 256 threads per CTA (one CTA per SM, co-resident: launched cooperatively): the warp LDL^T wants more than
// the 168 registers a 384-thread CTA leaves it.
#pragma once

#include "cluster.cuh"

namespace spk {

constexpr int CS_WARPS = 8;                    // 7 owner warps + 1 for ln|S_ab|
constexpr int CS_OWNERS = CS_WARPS - 1;
constexpr int CS_THREADS = CS_WARPS * 32;
constexpr int CS_OWN = 2;                      // clusters per owner warp
constexpr long long CS_SPIN_LIMIT = 1LL << 24; // polls before a missing publication is reported as an error

// what an owner publishes per cluster and merge
struct CsPub {
    double v;                                  // minimum of row k over the alive columns ...
    double d;                                  // distance of cluster k to the merged cluster of this merge
    unsigned long long tag;                    // (merge number + 1) << 32 | column of the minimum (0xffffffff: none)
};

struct CsDev {
    double* rec[2];                            // [n][REC] x 2 versions; version 0 holds the initial records
    const double* ld;                          // [n] initial ln|S_k|
    double* M;                                 // [n][n]
    uint8_t* alive_out;
    int64_t n;
    int metric; double lambda; double threshold; int max_spk; int variant;
    CsPub* pub;                                // [3][n], zero-initialised (tag 0 = nothing published)
    unsigned long long* bar;                   // grid barrier (test hook only)
    unsigned long long* stat;                  // as ClDev::stat
    spkdiar_merge* out; int64_t cap;
    long long* nmerge; double* final_min;
    unsigned long long* dbg;
    int* err;                                  // set when a publication did not arrive
    double* rowlog; long long rowlog_cap;
};

struct CsSmem {
    LdlScratch* ws;        // [CS_WARPS]
    double* Y;             // [CS_OWNERS][CS_OWN][REC]
    double* X;             // [REC]
    unsigned long long* wkey;   // [2][CS_WARPS] per-warp candidates of the block argmin (by merge parity)
    long long* widx;            // [2][CS_WARPS]
    double* shd;           // [2]
    double* own_ld;        // [CS_OWNERS][CS_OWN] ln|S_k| of the owned clusters
    double* own_v;         // [CS_OWNERS][CS_OWN] minimum of row k ...
    int32_t* own_c;        // [CS_OWNERS][CS_OWN] ... and its column (-1: none, -2: unknown, rescan)
    uint32_t* abits;       // [ceil(n/32)]
    uint8_t* ver;          // [n]
};
__device__ __forceinline__ CsSmem cs_carve(unsigned char* base, int64_t n) {
    CsSmem m;
    m.ws = reinterpret_cast<LdlScratch*>(base);
    m.Y = reinterpret_cast<double*>(base + CS_WARPS * sizeof(LdlScratch));
    m.X = m.Y + (size_t)CS_OWNERS * CS_OWN * REC;
    m.wkey = reinterpret_cast<unsigned long long*>(m.X + REC);
    m.widx = reinterpret_cast<long long*>(m.wkey + 2 * CS_WARPS);
    m.shd = reinterpret_cast<double*>(m.widx + 2 * CS_WARPS);
    m.own_ld = m.shd + 2;
    m.own_v = m.own_ld + CS_OWNERS * CS_OWN;
    m.own_c = reinterpret_cast<int32_t*>(m.own_v + CS_OWNERS * CS_OWN);
    m.abits = reinterpret_cast<uint32_t*>(m.own_c + CS_OWNERS * CS_OWN);
    m.ver = reinterpret_cast<uint8_t*>(m.abits + (n + 31) / 32);
    return m;
}
inline size_t cs_smem_bytes(int64_t n) {
    return CS_WARPS * sizeof(LdlScratch) + ((size_t)CS_OWNERS * CS_OWN + 1) * REC * sizeof(double)
           + 4 * CS_WARPS * 8 + 2 * sizeof(double) + CS_OWNERS * CS_OWN * 20
           + (size_t)((n + 31) / 32) * 4 + (size_t)n + 32;
}

// one log-determinant term with both operand records already in shared memory (the arithmetic of
// logdet_term, score.cuh)
__device__ __forceinline__ double cs_logdet(int kind, const double* rx, const double* ry, LdlScratch& w, int lane) {
    double a[Grid<D39>::NSLOT];
    const SmemSrc sx{rx}, sy{ry};
    double wx = 1.0, wy = 1.0;
    if (kind == FORM_MIX) glr_weights(sx(L39::CNT), sy(L39::CNT), wx, wy);
    const double n = form_matrix<D39>(a, kind, sx, sy, wx, wy, w, lane);
    const double lm = ldl_logdet<D39>(a, w, lane);
    if (kind == FORM_MIX) return range_map(lm);
    return finish_logdet(lm, n, D39);
}

// minimum of row r over the alive columns, ndarray.argmin order; column `ca` (>= 0) is taken as `va` instead
// of being loaded (the warp has just written it).  All lanes return the result.
__device__ __forceinline__ ClBest cs_row_scan(const double* __restrict__ row, int64_t n, const uint32_t* abits,
                                              int64_t ca, double va, int lane) {
    ClBest rb{d_inf(), INT64_MAX};
    for (int64_t c0 = 0; c0 < n; c0 += 32 * 8) {
        double v[8];
        bool ok[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int64_t c = c0 + 32 * u + lane;
            ok[u] = c < n && ((abits[c >> 5] >> (c & 31)) & 1u);
            v[u] = (ok[u] && c != ca) ? __ldcg(row + c) : va;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (ok[u]) cl_take(rb, v[u], c0 + 32 * u + lane);
    }
    return cl_warp_best(rb);
}

// ndarray.argmin order as ONE unsigned comparison: key 0 for NaN (first), else the order-preserving image of
// the double; ties by the flat index.  (key, idx) pairs are then reduced with plain integer compares.
__device__ __forceinline__ unsigned long long cs_key(double v) { return v != v ? 0ULL : cl_ord(v == 0.0 ? 0.0 : v); }   // -0 == +0
__device__ __forceinline__ double cs_unkey(unsigned long long k) {
    if (k == 0ULL) return d_nan();
    const unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffULL) : ~k;
    return __longlong_as_double((long long)u);
}
__device__ __forceinline__ void cs_take(unsigned long long& bk, long long& bi, unsigned long long k, long long i) {
    if (k < bk || (k == bk && i < bi)) { bk = k; bi = i; }
}
__device__ __forceinline__ void cs_publish(CsPub* p, double v, double d, unsigned long long tag) {
    p->v = v; p->d = d;
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(&p->tag), "l"(tag) : "memory");
}
__device__ __forceinline__ unsigned long long cs_ld_acquire(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

static __global__ void __launch_bounds__(CS_THREADS, 1) cl_small_loop(const CsDev g) {
    extern __shared__ __align__(16) unsigned char cs_smem[];
    const int64_t n = g.n;
    const CsSmem sm = cs_carve(cs_smem, n);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool owner = warp < CS_OWNERS;
    const int64_t owners = (int64_t)gridDim.x * CS_OWNERS;
    const int64_t gw = (int64_t)warp * gridDim.x + blockIdx.x;          // owner number: CTAs interleaved
    const int nwords = (int)((n + 31) / 32);
    const int kind = g.metric == SPKDIAR_BIC ? FORM_POOL : FORM_MIX;
    const double diag = g.variant == 1 ? CL_MAXINT_D : d_inf();
    const bool lockstep = g.rowlog != nullptr;             // test hook: grid barriers, nobody leaves

    for (int wd = threadIdx.x; wd < nwords; wd += CS_THREADS) {
        const int64_t lo = (int64_t)wd * 32;
        sm.abits[wd] = (n - lo >= 32) ? 0xffffffffu : ((1u << (int)(n - lo)) - 1u);
    }
    for (int64_t i = threadIdx.x; i < n; i += CS_THREADS) sm.ver[i] = 0;
    __syncthreads();

    // ---- the owners load their clusters: record -> shared memory, ln|S_k|, minimum of row k ----
    // (per-slot state lives in shared memory: the loops over the slots stay rolled, ONE copy of the
    // factorisation's straight-line code, no register pressure from the bookkeeping)
    double st_max = -d_inf(), st_min = d_inf();          // finite distances this warp computed (variant 1)
    bool st_any = false;
    double* const own_ld = sm.own_ld + warp * CS_OWN;
    double* const own_v = sm.own_v + warp * CS_OWN;
    int32_t* const own_c = sm.own_c + warp * CS_OWN;
#pragma unroll 1
    for (int j = 0; j < CS_OWN && owner; ++j) {
        const int64_t k = gw + (int64_t)j * owners;
        if (k < n) {
            double* y = sm.Y + ((size_t)warp * CS_OWN + j) * REC;
            for (int q = lane; q < REC; q += 32) y[q] = __ldcg(g.rec[0] + k * REC + q);
            const ClBest rb = cs_row_scan(g.M + k * n, n, sm.abits, -1, 0.0, lane);
            if (lane == 0) {
                own_ld[j] = __ldcg(g.ld + k);
                own_v[j] = rb.v; own_c[j] = rb.idx == INT64_MAX ? -1 : (int32_t)rb.idx;
                cs_publish(g.pub + k, rb.v, 0.0, (1ULL << 32) | (unsigned int)own_c[j]);
            }
        }
    }
    __syncwarp();
    unsigned long long bar_target = 0;

    int64_t nalive = n, a_prev = -1;
    long long nm = 0;
    bool failed = false;
    double det_max = 0.0, det_min = CL_MAXINT_D;           // spk-clustering.py:418-419
    long long t_scan = 0, t_pick = 0, t_score = 0, t_post = 0, t_ldl = 0;
    for (;;) {
        const long long c0 = clock64();
        // ---------- the decision: wait for what the owners published for this merge, take the argmin ----------
        // rows other than the merged one: their published minimum; the merged row a_prev: its entries are the
        // published distances (plus its diagonal)
        unsigned long long bk = ~0ULL;
        long long bidx = INT64_MAX;
        {
            const CsPub* pb = g.pub + (size_t)(nm % 3) * n;
            const unsigned long long want = (unsigned long long)(nm + 1);
            for (int64_t r = threadIdx.x; r < n; r += CS_THREADS) {
                if (!((sm.abits[r >> 5] >> (r & 31)) & 1u)) continue;
                const CsPub* p = pb + r;
                unsigned long long tag;
                long long spins = 0;
                while (((tag = cs_ld_acquire(&p->tag)) >> 32) != want) {
                    if (++spins > CS_SPIN_LIMIT || ((spins & 1023) == 0 && *((volatile int*)g.err))) { failed = true; break; }
                }
                if (failed) break;
                if (r == a_prev) continue;                  // the merged row: waited for (its record), entries below
                const int32_t c = (int32_t)(unsigned int)(tag & 0xffffffffULL);
                if (c >= 0) cs_take(bk, bidx, cs_key(__ldcg(&p->v)), r * n + c);
                if (a_prev >= 0) cs_take(bk, bidx, cs_key(__ldcg(&p->d)), a_prev * n + r);
            }
            if (a_prev >= 0 && threadIdx.x == 0) cs_take(bk, bidx, cs_key(diag), a_prev * n + a_prev);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long ok = __shfl_xor_sync(0xffffffffu, bk, o);
            const long long oi = __shfl_xor_sync(0xffffffffu, bidx, o);
            cs_take(bk, bidx, ok, oi);
        }
        const int par = (int)(nm & 1);
        if (lane == 0) { sm.wkey[par * CS_WARPS + warp] = bk; sm.widx[par * CS_WARPS + warp] = bidx; }
        if (__syncthreads_or(failed ? 1 : 0)) {
            if (threadIdx.x == 0) *g.err = 1;
            break;
        }
        bk = sm.wkey[par * CS_WARPS + (lane & (CS_WARPS - 1))];
        bidx = sm.widx[par * CS_WARPS + (lane & (CS_WARPS - 1))];
#pragma unroll
        for (int o = CS_WARPS / 2; o > 0; o >>= 1) {
            const unsigned long long ok = __shfl_xor_sync(0xffffffffu, bk, o);
            const long long oi = __shfl_xor_sync(0xffffffffu, bidx, o);
            cs_take(bk, bidx, ok, oi);
        }
        if (lockstep && a_prev >= 0) {                      // test hook: the row merge nm - 1 rewrote, for the host's replay
            if (nm - 1 < g.rowlog_cap)
                for (int64_t cidx = (int64_t)blockIdx.x * CS_THREADS + threadIdx.x; cidx < n; cidx += (int64_t)gridDim.x * CS_THREADS)
                    g.rowlog[(nm - 1) * n + cidx] = __ldcg(g.M + a_prev * n + cidx);
            cl_grid_barrier(g.bar, bar_target);             // before this merge rewrites the same row
        }
        const double mind = cs_unkey(bk);
        const int64_t bi = bidx / n, bj = bidx - (bidx / n) * n;
        const int64_t a = bi < bj ? bi : bj, b = bi < bj ? bj : bi;
        // ---------- stop test, spk-clustering.py:207-208 ----------
        const bool go = (mind <= g.threshold) || (g.max_spk > 0 && nalive > (int64_t)g.max_spk);
        if (!go || a == b || bidx == INT64_MAX) {
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                if (g.dbg) { g.dbg[0] = t_scan; g.dbg[1] = t_ldl; g.dbg[2] = t_pick; g.dbg[3] = t_score; g.dbg[4] = t_post; g.dbg[5] = nm; }
                *g.nmerge = nm;
                *g.final_min = mind;
                g.stat[2] = (unsigned long long)__double_as_longlong(det_max);
                g.stat[3] = (unsigned long long)__double_as_longlong(det_min);
            }
            break;
        }
        if (mind > det_max) det_max = mind;                 // spk-clustering.py:210-213
        if (mind < det_min) det_min = mind;
        if (blockIdx.x == 0 && threadIdx.x == 0 && nm < g.cap) {
            spkdiar_merge mr; mr.a = (int32_t)a; mr.b = (int32_t)b; mr.d = mind;
            g.out[nm] = mr;
        }
        const long long c1 = clock64();
        // ---------- the merged record, once per CTA ----------
        const int va = sm.ver[a];
        {
            const double* ra = g.rec[va] + a * REC;
            const double* rb = g.rec[sm.ver[b]] + b * REC;
            for (int q = threadIdx.x; q < REC; q += CS_THREADS) sm.X[q] = __ldcg(ra + q) + __ldcg(rb + q);
        }
        // does this CTA still own a live cluster after this merge?
        bool mine_alive = false;
#pragma unroll
        for (int j = 0; j < CS_OWN; ++j) {
            const int64_t k = gw + (int64_t)j * owners;
            mine_alive |= owner && k < n && k != b && ((sm.abits[k >> 5] >> (k & 31)) & 1u);
        }
        const int stay = __syncthreads_or(mine_alive ? 1 : 0);      // X complete; everybody has read ver / abits
        if (threadIdx.x == 0) { sm.abits[b >> 5] &= ~(1u << (b & 31)); sm.ver[a] = (uint8_t)(va ^ 1); }
        if (!stay && !lockstep) break;                      // nothing left to score here, nobody waits for this CTA
        const long long c2 = clock64();
        // ---------- rescoring: the last warp factorises the merged cluster, the owners their pooled matrices ----------
        if (!owner) {
            const double v = cs_logdet(FORM_X, sm.X, sm.X, sm.ws[warp], lane);
            if (lane == 0) sm.shd[0] = v;
        }
        double t0 = 0.0, t1 = 0.0;                          // pooled terms of slot 0 / 1
#pragma unroll 1
        for (int j = 0; j < CS_OWN && owner; ++j) {
            const int64_t k = gw + (int64_t)j * owners;
            if (k < n && k != a && k != b && ((sm.abits[k >> 5] >> (k & 31)) & 1u)) {
                const double t = cs_logdet(kind, sm.X, sm.Y + ((size_t)warp * CS_OWN + j) * REC, sm.ws[warp], lane);
                if (j == 0) t0 = t; else t1 = t;
            }
        }
        const long long c2a = clock64();
        __syncthreads();                                    // ln|S_ab|; bit b cleared, version of a switched
        const long long c3 = clock64();
        const double ld_ab = sm.shd[0];
        const double N1 = sm.X[L39::CNT];
        const unsigned long long tag_hi = (unsigned long long)(nm + 2) << 32;
        CsPub* const pn = g.pub + (size_t)((nm + 1) % 3) * n;
#pragma unroll 1
        for (int j = 0; j < CS_OWN && owner; ++j) {
            const int64_t k = gw + (int64_t)j * owners;
            if (k >= n) break;
            if (k == a) {
                // the merged cluster replaces a: record (shared memory + the other version in HBM), ln|S|
                double* y = sm.Y + ((size_t)warp * CS_OWN + j) * REC;
                double* dst = g.rec[va ^ 1] + a * REC;
                for (int q = lane; q < REC; q += 32) { const double x = sm.X[q]; y[q] = x; dst[q] = x; }
                __threadfence();
                __syncwarp();
                if (lane == 0) {
                    own_ld[j] = ld_ab;
                    own_c[j] = -2;                          // row a: its entries are the distances published below
                    cs_publish(pn + k, 0.0, 0.0, tag_hi | 0xffffffffULL);
                }
                continue;
            }
            if (!((sm.abits[k >> 5] >> (k & 31)) & 1u)) continue;
            const double N2 = sm.Y[((size_t)warp * CS_OWN + j) * REC + L39::CNT];
            const double t = j == 0 ? t0 : t1;
            const double d = g.metric == SPKDIAR_BIC ? bic_combine(N1, N2, ld_ab, own_ld[j], t, g.lambda)
                                                     : glr_combine(N1, N2, ld_ab, own_ld[j], t);
            bool rescan = false;
            double rv = own_v[j];
            int32_t rc = own_c[j];
            if (g.variant == 1 && d == d && d != d_inf() && d != -d_inf()) {
                st_any = true;
                if (d > st_max) st_max = d;
                if (d < st_min) st_min = d;
            }
            if (rc == -2 || rc == (int32_t)b) rescan = true;        // (-2: this was the merged row of the last merge)
            else if (g.variant == 1) {
                if (rc == (int32_t)a) {
                    // the row's minimum sat in the rewritten column: it stays there unless it got worse
                    if (cl_before(rv, a, d, a)) rescan = true; else rv = d;
                } else if (rc < 0 || cl_before(d, a, rv, rc)) { rv = d; rc = (int32_t)a; }
            }                                                       // variant 2: column a keeps its stale entries (Q5)
            if (lane == 0) {
                g.M[a * n + k] = d;                                 // row a
                if (g.variant == 1) g.M[k * n + a] = d;             // and column a
            }
            if (rescan) {
                const ClBest rb = cs_row_scan(g.M + k * n, n, sm.abits, g.variant == 1 ? a : -1, d, lane);
                rv = rb.v; rc = rb.idx == INT64_MAX ? -1 : (int32_t)rb.idx;
            }
            __syncwarp();
            if (lane == 0) {
                own_v[j] = rv; own_c[j] = rc;
                cs_publish(pn + k, rv, d, tag_hi | (unsigned int)rc);
            }
            __syncwarp();
        }
        const long long c4 = clock64();
        t_scan += c1 - c0; t_pick += c2 - c1; t_score += c3 - c2; t_post += c4 - c3; t_ldl += c2a - c2;
        a_prev = a;
        --nalive;
        ++nm;
    }
    // max / min over every finite distance the rescorings produced (spk-clustering.py:233-237)
    if (g.variant == 1 && owner && lane == 0 && st_any) {
        atomicMax(g.stat + 0, cl_ord(st_max));
        atomicMin(g.stat + 1, cl_ord(st_min));
    }
    if (blockIdx.x == 0) {
        __syncthreads();
        for (int64_t i = threadIdx.x; i < n; i += CS_THREADS) g.alive_out[i] = (sm.abits[i >> 5] >> (i & 31)) & 1u;
    }
}

cudaError_t cluster_small_configure() {
    return cudaFuncSetAttribute(cl_small_loop, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
}

}  // namespace spk
