// cluster.cuh - K5-K7: the agglomerative clustering engine.
//
// Replaces spk_cluster_hi of spk-clustering.py:178-260 (variant 1) and
// spk-clustering2.py:173-229 (variant 2).  The reference re-concatenates the raw
// frames of both clusters and runs np.cov for every pair it scores; here every
// cluster is one 820-double sufficient-statistics record resident in HBM
// (pooled statistics of two clusters = the sum of their records), the pair
// matrix stays on the device, and the merge loop is ONE persistent cooperative
// kernel: per merge a grid-wide exact argmin (NaN-first, first flat index wins,
// as ndarray.min()/argmin() behave - SURVEY.md Q8), the stop test, and the
// rescoring of the merged row (and column in variant 1), separated by two grid
// barriers.  Variant 2 keeps the reference's stale entries (Q5): the matrix is
// the full N x N array, only row `a` is rewritten after a merge.
//
// Indices inside the kernel are ORIGINAL cluster indices with an alive mask;
// deleting row/column b in the reference preserves order, so the flat-index
// tie-break is the same.  The host wrapper converts to compacted indices.
#pragma once

#include "common.cuh"
#include "score.cuh"

struct spkdiar_clus {
    spkdiar_ctx* ctx = nullptr;
    spkdiar_feat* feat = nullptr;
    int64_t n = 0;               // initial clusters
    int metric = SPKDIAR_BIC;
    double lambda = 1.3;
    int64_t* seg = nullptr;      // device: a[n], b[n]
    int64_t* hseg = nullptr;     // host copy (new[]): the direct records (K5) are planned on the host
    double* rec = nullptr;       // [n][REC]
    double* ld = nullptr;        // [n] ln|S_i|
    double* M = nullptr;         // [n][n]
    uint8_t* alive = nullptr;    // [n] (final state, for the test hook)
    // KL2 (spk-clustering.py:124-133): per cluster diag(S), diag(S^-1) and the running float32 sum of its frames in
    // the order the reference concatenates them; the turns of a cluster as a linked list of initial segments
    double* kside = nullptr;     // [n][KS]
    float* ksum = nullptr;       // [n][VS]
    int32_t* klist = nullptr;    // [3][n]: next segment, head and tail of cluster
    bool ran = false;
    double* rowlog_host = nullptr;   // test hook (spkdiar_cluster_rowlog): row a after every merge
    int64_t rowlog_cap = 0;
    uint64_t counters[8] = {0};      // phase cycle counters of the last persistent run (spkdiar_cluster_counters)
};

namespace spk {

constexpr int CL_WARPS = 12;
constexpr int CL_THREADS = CL_WARPS * 32;
#define CL_MAXINT_D 9223372036854775807.0      /* float(sys.maxint) == 2^63 */

struct ClBest { double v; int64_t idx; };

// "a comes before b" under ndarray.argmin(): NaN first, then smaller value, then smaller flat index
__device__ __forceinline__ bool cl_before(double av, int64_t ai, double bv, int64_t bi) {
    const bool an = av != av, bn = bv != bv;
    if (an || bn) return an && (!bn || ai < bi);
    return av < bv || (av == bv && ai < bi);
}
__device__ __forceinline__ void cl_take(ClBest& x, double v, int64_t i) {
    if (cl_before(v, i, x.v, x.idx)) { x.v = v; x.idx = i; }
}
__device__ __forceinline__ ClBest cl_warp_best(ClBest x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, x.v, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, (long long)x.idx, o);
        cl_take(x, ov, oi);
    }
    return x;
}

// order-preserving map double -> uint64 for atomicMax / atomicMin
__device__ __forceinline__ unsigned long long cl_ord(double d) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(d);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ULL);
}
inline double cl_unord(unsigned long long k) {
    const unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffULL) : ~k;
    double d;
    memcpy(&d, &u, sizeof(d));
    return d;
}
__device__ __forceinline__ void cl_track(double d, unsigned long long* stat) {   // stat[0]=max, stat[1]=min over finite d
    if (d == d_inf() || d == -d_inf() || d != d) return;    // `d > max` / `d < min` are false for NaN
    atomicMax(stat + 0, cl_ord(d));
    atomicMin(stat + 1, cl_ord(d));
}
// The same statistics kept in registers while a warp works through its pairs and folded into the two global
// words ONCE at the end: every pair sending two atomics to the same two addresses (2 x 12,000 per merge at
// 24,000 clusters, 2 x 1.96 M in the fill of config 3) queues them up in one L2 slice.  Maximum and minimum
// of ordered keys do not depend on the order, so the result is the same to the bit.
struct ClTrack {
    unsigned long long hi = 0ULL, lo = ~0ULL;               // ordered keys; (0, ~0) = nothing seen
    __device__ __forceinline__ void see(double d) {
        if (d == d_inf() || d == -d_inf() || d != d) return;
        const unsigned long long k = cl_ord(d);
        hi = k > hi ? k : hi;
        lo = k < lo ? k : lo;
    }
    __device__ __forceinline__ void flush(unsigned long long* stat) const {
        if (hi >= lo) { atomicMax(stat + 0, hi); atomicMin(stat + 1, lo); }
    }
};

static __global__ void __launch_bounds__(256)
cl_init_records(const Stats st, const int64_t* __restrict__ seg, int64_t n,
                double* __restrict__ rec) {
    const int64_t s = blockIdx.x;
    const WinSrc w(st, seg[s], seg[n + s], REC);
    for (int q = threadIdx.x; q < REC; q += blockDim.x) rec[s * REC + q] = w(q);
}

static __global__ void __launch_bounds__(SC_THREADS, 3)
cl_self_logdet(const double* rec, int64_t n, double* __restrict__ ld) {
    extern __shared__ __align__(16) unsigned char sc_smem[];
    WarpScratch* ws = reinterpret_cast<WarpScratch*>(sc_smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t i = (int64_t)blockIdx.x * SC_WARPS + warp; i < n; i += (int64_t)gridDim.x * SC_WARPS) {
        const RecSrc X{rec + i * REC};
        const double v = logdet_term(0, SPKDIAR_BIC, X, X, ws[warp], lane);
        if (lane == 0) ld[i] = v;
    }
}

static __global__ void cl_fill_const(double* __restrict__ M, int64_t n, double offdiag, double diag) {
    const int64_t total = n * n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / n, c = i - r * n;
        M[i] = r == c ? diag : offdiag;
    }
}

// pair p of the strict upper triangle, row-major -> (i, j)
__device__ __forceinline__ void cl_pair(int64_t p, int64_t n, int64_t& i, int64_t& j) {
    const double b = 2.0 * (double)n - 1.0;
    int64_t r = (int64_t)((b - sqrt(b * b - 8.0 * (double)p)) * 0.5);
    if (r < 0) r = 0;
    if (r > n - 2) r = n - 2;
    // first pair of row r is r*n - r*(r+1)/2
    while (r > 0 && r * n - (r * (r + 1)) / 2 > p) --r;
    while ((r + 1) * n - ((r + 1) * (r + 2)) / 2 <= p) ++r;
    i = r;
    j = p - (r * n - (r * (r + 1)) / 2) + r + 1;
}

// distance of clusters X (arr1) and Y (arr2) from their records and cached ln|S|
template <class SrcX, class SrcY>
__device__ __forceinline__ double cl_pair_distance(int metric, double lambda, const SrcX& X, const SrcY& Y,
                                                   double ldx, double ldy, WarpScratch& w, int lane) {
    const double t = logdet_term(2, metric, X, Y, w, lane);
    const double N1 = X(L39::CNT), N2 = Y(L39::CNT);
    return metric == SPKDIAR_BIC ? bic_combine(N1, N2, ldx, ldy, t, lambda)
                                 : glr_combine(N1, N2, ldx, ldy, t);
}

// initial fill, spk-clustering.py:188-200 / spk-clustering2.py:180-184
static __global__ void __launch_bounds__(SC_THREADS, 3)
cl_fill_pairs(const double* rec, const double* __restrict__ ld, int64_t n, int metric, double lambda,
              int variant, double* __restrict__ M, unsigned long long* stat) {
    extern __shared__ __align__(16) unsigned char sc_smem[];
    WarpScratch* ws = reinterpret_cast<WarpScratch*>(sc_smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t npair = (n * (n - 1)) / 2;
    // every warp takes a CONTIGUOUS run of pairs (row-major over the upper triangle): the record of cluster i
    // stays staged in shared memory while the run stays in row i - one record load per pair instead of two
    const int64_t gw = (int64_t)blockIdx.x * SC_WARPS + warp, nw = (int64_t)gridDim.x * SC_WARPS;
    const int64_t per = (npair + nw - 1) / nw;
    int64_t p = gw * per;
    const int64_t pe = p + per < npair ? p + per : npair;
    if (p >= pe) return;
    WarpScratch& w = ws[warp];
    int64_t i, j, staged = -1;
    cl_pair(p, n, i, j);
    ClTrack track;
    for (; p < pe; ++p) {
        if (i != staged) {
            __syncwarp();
            stage_record(RecSrc{rec + i * REC}, w.rec[0], lane);
            __syncwarp();
            staged = i;
        }
        const SmemSrc X{w.rec[0]};
        const RecSrc Y{rec + j * REC};
        // (ln|S_i|, ln|S_j| are read AFTER the factorisation: nothing that can wait is held across it)
        const double t = logdet_term(2, metric, X, Y, w, lane);
        if (lane == 0) {
            const double N1 = X(L39::CNT), N2 = Y(L39::CNT);
            const double d = metric == SPKDIAR_BIC ? bic_combine(N1, N2, ld[i], ld[j], t, lambda)
                                                   : glr_combine(N1, N2, ld[i], ld[j], t);
            M[i * n + j] = d;
            if (variant == 1) { M[j * n + i] = d; track.see(d); }
        }
        if (++j == n) { ++i; j = i + 1; }
    }
    if (lane == 0) track.flush(stat);
}

// One mailbox slot per (parity, sending rank): the sender writes the payload, then the sequence
// number with release semantics at system scope; the receiver polls the sequence number with
// acquire semantics.  Slots live in the RECEIVER's memory, senders write them through
// peer-mapped pointers (NVLink).
struct ClMail { double v; long long idx; unsigned long long seq; unsigned long long pad; };
constexpr int CL_MAX_RANKS = 16;
__device__ __forceinline__ void cl_st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long cl_ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

struct ClDev {
    double* rec; double* ld; double* M; uint8_t* alive_out;
    int64_t n;
    int metric; double lambda; double threshold; int max_spk; int variant;
    int32_t rank, nranks;        // pair (r, c) is scored and kept by rank (r + c) % nranks
    double* rowmin_v;            // [n] minimum of row r over the alive (owned) columns ...
    int32_t* rowmin_c;           // [n] ... and its (first) column, -1: none
    int32_t* repoch;             // [n] iteration whose flagged list holds row r
    int32_t* flist;              // [2][n] flagged rows of the even / odd iterations (-(a+1): the merged row)
    int32_t* fcount;             // [2]
    ClBest* slotsB;              // [grid] per-CTA minimum of the row rewritten by the last rescoring
    ClBest* slots;               // [grid] per-CTA candidates, double-buffered by merge parity: [2][grid]
    unsigned long long* bar;
    unsigned long long* stat;    // [0]=max [1]=min (ordered keys) over finite distances; [2]=max_det [3]=min_det
    spkdiar_merge* out; int64_t cap;
    long long* nmerge;           // merges performed
    double* final_min;           // the minimum that stopped the loop
    unsigned long long* dbg;     // optional phase cycle counters of CTA 0
    double* rowlog;              // test hook: [rowlog_cap][n] row a of M as rewritten by merge nm (else null)
    long long rowlog_cap;
    // host-driven (sharded) run only: state that the persistent kernel keeps on chip
    uint32_t* abits_g;           // [ceil(n/32)] alive mask
    double* pend;                // [REC + 2] merged record, ln|S_ab|, index a (+1; 0 = nothing pending)
    ClBest* local_best;          // this rank's candidate of the current iteration
    unsigned int* ticket;        // last-CTA election of the argmin kernel
    // device-decided run (NCCL exchange on the stream, no host round trip per merge)
    int* stopped;                // set when the stop test fails; later launches return at once
    const ClBest* gathered;      // [nranks] candidates of all ranks (all-gather target)
    double* det;                 // [0] max_det [1] min_det
    // peer-memory exchange inside the persistent kernel
    ClMail* mbox[CL_MAX_RANKS];  // mbox[r]: rank r's mailbox [3][nranks] as mapped on THIS device (mbox[rank]: local);
                                 // sets 0 / 1: candidates of even / odd merges, set 2: the final statistics round
    unsigned long long seq_base; // sequence numbers of this run start above it
    int* err;                    // set when a peer did not answer in time
    // KL2 (single GPU): cached sides and running sums per cluster, turn lists, the frames
    double* kside; float* ksum; int32_t* knext; int32_t* khead; int32_t* ktail;
    const float* x; const int64_t* seg;
};

__device__ __forceinline__ void cl_grid_barrier(unsigned long long* ctr, unsigned long long& target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += (unsigned long long)gridDim.x;
        __threadfence();
        atomicAdd(ctr, 1ULL);
        while (*((volatile unsigned long long*)ctr) < target) { }
        __threadfence();
    }
    __syncthreads();
}

// shared memory of the merge kernels: per-warp scratch, merged record [REC], small arrays, alive bitmask
struct ClSmem {
    WarpScratch* ws; double* merged; ClBest* wbest; double* shd; int* wbusy; double* kside; float* ksum; uint32_t* abits;
};
__device__ __forceinline__ ClSmem cl_carve(unsigned char* base) {
    ClSmem m;
    m.ws = reinterpret_cast<WarpScratch*>(base);
    m.merged = reinterpret_cast<double*>(base + CL_WARPS * sizeof(WarpScratch));
    m.wbest = reinterpret_cast<ClBest*>(m.merged + REC);
    m.shd = reinterpret_cast<double*>(m.wbest + CL_WARPS);          // [0] ld_ab
    m.wbusy = reinterpret_cast<int*>(m.shd + 2);                    // [CL_WARPS] warp has a pair in round 0
    m.kside = reinterpret_cast<double*>(m.wbusy + CL_WARPS);        // [KS] KL2: side of the merged cluster
    m.ksum = reinterpret_cast<float*>(m.kside + KS);                // [VS] ... and its running sum
    m.abits = reinterpret_cast<uint32_t*>(m.ksum + VS);
    return m;
}
__device__ __forceinline__ bool cl_own(const ClDev& g, int64_t r, int64_t c) {
    return g.nranks == 1 || (int32_t)((r + c) % g.nranks) == g.rank;
}

// ---------- KL2 in the clustering engine ----------
// the scratch kl2_side_one wants, laid over a warp's WarpScratch (operand buffer 0 = record / factor, buffer 1 = 1 / pivots)
struct ClKl2Scr { LdlScratch& w; double* rec; double* pinv; };
__device__ __forceinline__ ClKl2Scr cl_kl2_scr(WarpScratch& ws) { return ClKl2Scr{ws, ws.rec[0], ws.rec[1]}; }

// np.mean of a float32 matrix sums row after row: continue the running sum (s0, s1) of lane's
// dimensions over the frames [a, b), sixteen loads in flight
__device__ __forceinline__ void cl_kl2_add_rows(const float* __restrict__ x, int64_t a, int64_t b, int lane, float& s0, float& s1) {
    const int off2 = lane + 32 < D39 ? 32 : 0;
    const float* row = x + a * D39 + lane;
    const int64_t nrow = b - a;
    for (int64_t r0 = 0; r0 < nrow; r0 += 16) {
        float u[16], v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const bool ok = r0 + q < nrow;
            u[q] = ok ? __ldg(row + (r0 + q) * D39) : 0.f;
            v[q] = ok ? __ldg(row + (r0 + q) * D39 + off2) : 0.f;
        }
#pragma unroll
        for (int q = 0; q < 16; ++q)
            if (r0 + q < nrow) { s0 = __fadd_rn(s0, u[q]); s1 = __fadd_rn(s1, v[q]); }
    }
}

// initial clusters: side and running sum of segment s, one warp each
static __global__ void __launch_bounds__(SC_THREADS, 3)
cl_kl2_init(const double* __restrict__ rec, const float* __restrict__ x, const int64_t* __restrict__ seg, int64_t n,
            double* __restrict__ kside, float* __restrict__ ksum, int32_t* __restrict__ klist) {
    extern __shared__ __align__(16) unsigned char sc_smem[];
    WarpScratch* ws = reinterpret_cast<WarpScratch*>(sc_smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t s = (int64_t)blockIdx.x * SC_WARPS + warp; s < n; s += (int64_t)gridDim.x * SC_WARPS) {
        ClKl2Scr k = cl_kl2_scr(ws[warp]);
        __shared__ double side[SC_WARPS][KS];
        kl2_side_one(RecSrc{rec + s * REC}, k, side[warp], side[warp] + VS, lane);
        for (int q = lane; q < KS; q += 32) kside[s * KS + q] = (q % VS) < D39 ? side[warp][q] : 0.0;
        float s0 = 0.f, s1 = 0.f;
        cl_kl2_add_rows(x, seg[s], seg[n + s], lane, s0, s1);
        ksum[s * VS + lane] = s0;
        if (lane + 32 < D39) ksum[s * VS + lane + 32] = s1;
        if (lane == 0) { klist[s] = -1; klist[n + s] = (int32_t)s; klist[2 * n + s] = (int32_t)s; }
        __syncwarp();
    }
}

// initial fill with the KL2 distances, one warp per pair (cached sides: no factorisation here)
static __global__ void __launch_bounds__(256)
cl_fill_pairs_kl2(const double* __restrict__ rec, const double* __restrict__ kside, const float* __restrict__ ksum, int64_t n,
                  int variant, double* __restrict__ M, unsigned long long* stat) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t npair = (n * (n - 1)) / 2;
    ClTrack track;
    for (int64_t p = (int64_t)blockIdx.x * 8 + warp; p < npair; p += (int64_t)gridDim.x * 8) {
        int64_t i, j;
        cl_pair(p, n, i, j);
        const double d = kl2_distance_cached(kside + i * KS, kside + j * KS, ksum + i * VS, ksum + j * VS,
                                             rec[i * REC + L39::CNT], rec[j * REC + L39::CNT], lane);
        if (lane == 0) {
            M[i * n + j] = d;
            if (variant == 1) { M[j * n + i] = d; track.see(d); }
        }
    }
    if (lane == 0) track.flush(stat);
}

// ---------- ARGMIN over the alive (owned) part of the matrix, from the row-minimum cache ----------
// rowmin[r] = first minimum of row r over the alive columns (ndarray.argmin order: NaN first,
// then value, then column).  A merge rewrites row a, kills column b and - variant 1 - rewrites
// column a.  Row a and every row whose cached minimum is no longer trustworthy are put on the
// FLAGGED LIST of the next iteration (epoch-stamped, so nobody has to clear flags); all other
// rows just compare their cached minimum with the one new entry.  An iteration then costs
// one pass over the n cached minima (a row per LANE) plus a few CTA-wide row rescans instead
// of n^2 entries; row a itself is rebuilt from the per-CTA minima of the rescoring.
// Returns this CTA's candidate in thread 0.
__device__ __forceinline__ ClBest cl_phase_argmin(const ClDev& g, long long nm, const ClSmem& sm,
                                                  int warp, int lane, int64_t gwarp, int64_t nwarps) {
    const int64_t n = g.n;
    const uint32_t* abits = sm.abits;
    ClBest* wbest = sm.wbest;
    const int par = (int)(nm & 1);
    if (blockIdx.x == 0 && threadIdx.x == 0) g.fcount[par ^ 1] = 0;         // list of the next iteration
    ClBest mine{d_inf(), INT64_MAX};
    const int nfl = __ldcg(g.fcount + par);
    const int32_t* fl = g.flist + (int64_t)par * n;
    for (int li = blockIdx.x; li < nfl; li += gridDim.x) {                  // one flagged row per CTA at a time
        const int32_t code = __ldcg(fl + li);
        const int64_t r = code < 0 ? -(int64_t)code - 1 : code;
        ClBest rb{d_inf(), INT64_MAX};
        if (code < 0) {
            // the merged row: its entries are the distances of the last rescoring (minimum per
            // CTA in slotsB) and the diagonal
            for (int t = threadIdx.x; t <= (int)gridDim.x; t += CL_THREADS) {
                if (t < (int)gridDim.x) {
                    const long long c = __ldcg((const long long*)&g.slotsB[t].idx);
                    if (c != INT64_MAX) cl_take(rb, __ldcg(&g.slotsB[t].v), c);
                } else if (cl_own(g, r, r)) {
                    cl_take(rb, __ldcg(g.M + r * n + r), r);
                }
            }
        } else {
            const double* row = g.M + r * n;
            for (int64_t c0 = 0; c0 < n; c0 += CL_THREADS * 4) {            // four independent loads per thread
                double v[4];
                bool ok[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int64_t c = c0 + CL_THREADS * u + threadIdx.x;
                    ok[u] = c < n && ((abits[c >> 5] >> (c & 31)) & 1u) && cl_own(g, r, c);
                    v[u] = ok[u] ? __ldcg(row + c) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (ok[u]) cl_take(rb, v[u], c0 + CL_THREADS * u + threadIdx.x);
            }
        }
        rb = cl_warp_best(rb);
        if (lane == 0) wbest[warp] = rb;
        __syncthreads();
        if (threadIdx.x == 0) {
            ClBest bb = wbest[0];
            for (int w = 1; w < CL_WARPS; ++w) cl_take(bb, wbest[w].v, wbest[w].idx);
            g.rowmin_v[r] = bb.v; g.rowmin_c[r] = bb.idx == INT64_MAX ? -1 : (int32_t)bb.idx;
            if (bb.idx != INT64_MAX) cl_take(mine, bb.v, r * n + bb.idx);
        }
        __syncthreads();
    }
    // the cached minima of all other alive rows, one row per lane
    const int32_t epoch = (int32_t)nm;
    for (int64_t r = gwarp * 32 + lane; r < n; r += nwarps * 32) {
        if (!((abits[r >> 5] >> (r & 31)) & 1u)) continue;
        if (__ldcg(g.repoch + r) == epoch) continue;                        // on the flagged list: its CTA has it
        const int32_t rc = __ldcg(g.rowmin_c + r);
        if (rc >= 0) cl_take(mine, __ldcg(g.rowmin_v + r), r * n + rc);
    }
    mine = cl_warp_best(mine);
    if (lane == 0) wbest[warp] = mine;
    __syncthreads();
    ClBest out{d_inf(), INT64_MAX};
    if (threadIdx.x == 0) {
        out = wbest[0];
        for (int w = 1; w < CL_WARPS; ++w) cl_take(out, wbest[w].v, wbest[w].idx);
    }
    __syncthreads();
    return out;
}

// ---------- MERGE + RESCORE: clusters a and b (a < b) become a ----------
// sm.merged must hold rec[a] + rec[b] and bit b of sm.abits must be cleared (and a
// __syncthreads passed).  Scores the merged cluster against every alive cluster k whose
// pair (a, k) this rank owns, writes row a (and column a in variant 1), maintains the
// row-minimum cache and the flagged list of iteration nm + 1.  Returns ln|S_ab|.
// (KL2 is a compile-time switch: with the KL2 branches in the same instantiation the BIC / GLR rescoring of large
// problems ran a third slower - 217 k -> 290 k cycles per merge at 23,881 clusters)
template <bool KL2 = false>
__device__ __forceinline__ double cl_phase_apply(const ClDev& g, long long nm, int64_t a, int64_t b, const ClSmem& sm,
                                                 int warp, int lane, int64_t gwarp, int64_t nwarps) {
    const int64_t n = g.n;
    const uint32_t* abits = sm.abits;
    const int nwords = (int)((n + 31) / 32);
    const int par = (int)(nm & 1);
    const SmemSrc X{sm.merged};
    // ln|S_ab| is needed by every pair of the rescoring.  Every CTA computes it itself (no
    // broadcast): by a warp that has no pair in the first round if there is one (the usual
    // case: fewer alive clusters than warps), else by warp 0 ahead of its own pairs - while
    // the other warps already factorise their first pooled matrix.  One __syncthreads later
    // everybody knows it.
    // Pairs are dealt to the warps by ORDINAL among the alive clusters (a dense numbering: no warp
    // gets two pairs while another has none because of where the dead indices happen to lie).
    // select(o) = index of the o-th alive cluster: lanes count the bits of their share of the
    // mask words, a warp scan finds the lane that holds the target, that lane walks its words.
    // Sharded run: this rank scores the pairs (a, k) with (a + k) % nranks == rank, i.e. the clusters k of ONE
    // residue class - the ordinals count only those (for a power-of-two number of ranks the class is the same bit
    // pattern in every mask word), so that a rank's share of the row is dealt densely over its warps too.  Other
    // rank counts keep the dense numbering of all alive clusters and skip the pairs of the other ranks.
    const bool by_class = g.nranks > 1 && g.nranks <= 32 && (g.nranks & (g.nranks - 1)) == 0;
    uint32_t cls = 0xffffffffu;
    if (by_class) {
        const int c0 = (int)(((int64_t)g.rank - a % g.nranks + g.nranks) % g.nranks);
        cls = 0;
        for (int bit = c0; bit < 32; bit += g.nranks) cls |= 1u << bit;
    }
    const int wpl = (nwords + 31) / 32;                       // mask words per lane
    int mycnt = 0;
    for (int w = lane * wpl; w < (lane + 1) * wpl && w < nwords; ++w) mycnt += __popc(abits[w] & cls);
    int incl = mycnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const int total_alive = __shfl_sync(0xffffffffu, incl, 31);
    auto select = [&](int64_t o) -> int64_t {                  // warp-uniform o; n when o >= #alive
        if (o >= total_alive) return n;
        const int excl = incl - mycnt;
        const bool mine = o >= excl && o < incl;
        int64_t found = 0;
        if (mine) {
            int rem = (int)o - excl;
            for (int w = lane * wpl;; ++w) {
                const uint32_t word = abits[w] & cls;
                const int c = __popc(word);
                if (rem < c) { found = (int64_t)w * 32 + (__fns(word, 0, rem + 1)); break; }
                rem -= c;
            }
        }
        const unsigned who = __ballot_sync(0xffffffffu, mine);
        return __shfl_sync(0xffffffffu, found, __ffs(who) - 1);
    };
    int64_t ord = gwarp;
    int64_t k = select(ord);
    auto mine_pair = [&](int64_t kk) { return kk < n && kk != a && cl_own(g, a, kk); };
    constexpr bool kl2 = KL2;
    if constexpr (KL2) {
        // KL2 (spk-clustering.py:124-133): a pair costs no factorisation - diag(S), diag(S^-1) and the float32
        // sum of every cluster are cached.  What the merge needs first: the side of the merged cluster (warp 0,
        // one factorisation + inverse) and its running sum, a's sum continued over b's frames turn by turn in the
        // order the reference concatenates them (warp 1).  Every CTA does both itself: no broadcast.
        if (warp == 0) {
            ClKl2Scr scr = cl_kl2_scr(sm.ws[0]);
            kl2_side_one(X, scr, sm.kside, sm.kside + VS, lane);
        } else if (warp == 1) {
            float s0 = __ldcg(g.ksum + a * VS + lane), s1 = lane + 32 < D39 ? __ldcg(g.ksum + a * VS + lane + 32) : 0.f;
            for (int32_t s = __ldcg(g.khead + b); s >= 0; s = __ldcg(g.knext + s))
                cl_kl2_add_rows(g.x, g.seg[s], g.seg[n + s], lane, s0, s1);
            sm.ksum[lane] = s0;
            if (lane + 32 < D39) sm.ksum[lane + 32] = s1;
        }
    }
    if (lane == 0) sm.wbusy[warp] = (kl2 || mine_pair(k)) ? 1 : 0;
    __syncthreads();
    int ldw = 0;
#pragma unroll
    for (int w = CL_WARPS - 1; w >= 0; --w) if (!sm.wbusy[w]) ldw = w;
    if (kl2) ldw = -1;                            // nobody computes ln|S_ab|
    const double N1 = sm.merged[L39::CNT];
    double ld_ab = 0.0;
    ClBest rowa{d_inf(), INT64_MAX};             // lane 0: best (distance, k) this warp produced for row a
    ClTrack track;                               // lane 0: extremes of the distances this warp computed
    for (int round = 0;; ++round) {
        bool has; int term; int64_t kk;
        if (round == 0 && warp == ldw && !sm.wbusy[ldw]) { has = true; term = 0; kk = a; }       // a spare warp
        else if (round == 0 && warp == ldw) { has = true; term = 0; kk = a; ord -= nwarps; }     // none spare: warp 0 first
        else { has = mine_pair(k); term = 2; kk = has ? k : a; }
        if (round > 0 && k >= n) break;
        double t = 0.0;
        double N2 = 0.0, ldk = 0.0, vold = 0.0;
        int32_t cmin = -1;
        if (kl2 && has && term == 2 && lane == 0) N2 = __ldcg(g.rec + kk * REC + L39::CNT);
        double dk = 0.0;
        if constexpr (KL2) {
            if (has) {
                N2 = __shfl_sync(0xffffffffu, N2, 0);
                dk = kl2_distance_cached<true>(sm.kside, g.kside + kk * KS, sm.ksum, g.ksum + kk * VS, N1, N2, lane);
            }
        } else {
            if (has) {
                const RecSrc Y{g.rec + kk * REC};
                t = logdet_term(term, g.metric, X, Y, sm.ws[warp], lane);
            }
        }
        if (round == 0 && !kl2) {
            if (warp == ldw && lane == 0) sm.shd[0] = t;
            __syncthreads();
            ld_ab = sm.shd[0];
        }
        if (has && term == 2 && lane == 0) {
            // (requested AFTER the factorisation: held across it, these seven registers cost more in spills inside
            // it than the round trips they would hide - measured 217 k against 187 k cycles per merge at 23,881
            // clusters; keeping the per-warp results in shared memory instead of registers was slower again, 200 k)
            if (!kl2) { N2 = __ldcg(g.rec + kk * REC + L39::CNT); ldk = __ldcg(g.ld + kk); }
            cmin = __ldcg(g.rowmin_c + kk);
            if (g.variant == 1) vold = __ldcg(g.rowmin_v + kk);
            const double d = kl2 ? dk : (g.metric == SPKDIAR_BIC ? bic_combine(N1, N2, ld_ab, ldk, t, g.lambda)
                                                                 : glr_combine(N1, N2, ld_ab, ldk, t));
            g.M[a * n + kk] = d;                                            // row a
            cl_take(rowa, d, kk);                                            // minimum of the new row a
            bool flag = false;
            if (g.variant == 1) {
                g.M[kk * n + a] = d; track.see(d);                          // and column a
                if (cmin == (int32_t)b) flag = true;
                else if (cmin == (int32_t)a) {
                    // the row's minimum sat in the rewritten column: it stays there unless it got worse
                    if (cl_before(vold, a, d, a)) flag = true; else g.rowmin_v[kk] = d;
                } else if (cmin < 0 || cl_before(d, a, vold, cmin)) { g.rowmin_v[kk] = d; g.rowmin_c[kk] = (int32_t)a; }
            } else if (cmin == (int32_t)b) {
                flag = true;                                                // column a keeps its stale entries (Q5)
            }
            if (flag) {
                g.repoch[kk] = (int32_t)(nm + 1);
                const int at = atomicAdd(g.fcount + (par ^ 1), 1);
                g.flist[(int64_t)(par ^ 1) * n + at] = (int32_t)kk;
            }
        }
        ord += nwarps;
        k = select(ord);
    }
    if (lane == 0) track.flush(g.stat);
    // A row whose pair with a belongs to ANOTHER rank still loses column b here: if its cached
    // minimum sat there (or in the column a another rank rewrites - not this rank's entry, so the
    // local row only loses it) it has to be rescanned.  One lane per row.
    if (g.nranks > 1) {
        for (int64_t r = gwarp * 32 + lane; r < n; r += nwarps * 32) {
            if (!((abits[r >> 5] >> (r & 31)) & 1u) || r == a || cl_own(g, a, r)) continue;
            const int32_t cmin = __ldcg(g.rowmin_c + r);
            if (cmin == (int32_t)b) {
                g.repoch[r] = (int32_t)(nm + 1);
                const int at = atomicAdd(g.fcount + (par ^ 1), 1);
                g.flist[(int64_t)(par ^ 1) * n + at] = (int32_t)r;
            }
        }
    }
    // per-CTA minimum of the new row a -> slotsB; row a goes on the flagged list as "merged row"
    if (lane == 0) sm.wbest[warp] = rowa;
    __syncthreads();
    if (threadIdx.x == 0) {
        ClBest bb = sm.wbest[0];
        for (int w = 1; w < CL_WARPS; ++w) cl_take(bb, sm.wbest[w].v, sm.wbest[w].idx);
        g.slotsB[blockIdx.x] = bb;
        if (blockIdx.x == 0) {
            g.repoch[a] = (int32_t)(nm + 1);
            const int at = atomicAdd(g.fcount + (par ^ 1), 1);
            g.flist[(int64_t)(par ^ 1) * n + at] = -(int32_t)a - 1;
        }
    }
    return ld_ab;
}

// the merge loop as ONE persistent cooperative kernel (single GPU)
template <bool KL2>
static __global__ void __launch_bounds__(CL_THREADS, 1) cl_merge_loop(const ClDev g) {
    extern __shared__ __align__(16) unsigned char cl_smem[];
    const ClSmem sm = cl_carve(cl_smem);
    __shared__ ClBest gbest;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n = g.n;
    // warp numbering interleaves the CTAs: with fewer rows than warps every SM gets some, and every CTA
    // keeps idle warps (one of them computes ln|S_ab| during the rescoring)
    const int64_t gwarp = (int64_t)warp * gridDim.x + blockIdx.x;
    const int64_t nwarps = (int64_t)gridDim.x * CL_WARPS;
    const int nwords = (int)((n + 31) / 32);
    for (int wd = threadIdx.x; wd < nwords; wd += CL_THREADS) {
        const int64_t lo = (int64_t)wd * 32;
        sm.abits[wd] = (n - lo >= 32) ? 0xffffffffu : ((1u << (int)(n - lo)) - 1u);
    }
    __syncthreads();
    unsigned long long bar_target = 0;
    int64_t nalive = n;
    long long nm = 0;
    double det_max = 0.0, det_min = CL_MAXINT_D;           // spk-clustering.py:418-419

    long long t_scan = 0, t_b1 = 0, t_pick = 0, t_score = 0, t_b2 = 0;
    for (;;) {
        const long long c0 = clock64();
        const ClBest cta = cl_phase_argmin(g, nm, sm, warp, lane, gwarp, nwarps);
        if (threadIdx.x == 0) g.slots[(nm & 1) * gridDim.x + blockIdx.x] = cta;
        const long long c1 = clock64();
        cl_grid_barrier(g.bar, bar_target);
        const long long c2 = clock64();
        if (warp == 0) {
            ClBest b{d_inf(), INT64_MAX};
            const ClBest* sl = g.slots + (nm & 1) * gridDim.x;
            for (int s = lane; s < (int)gridDim.x; s += 32) {
                const double v = __ldcg(&sl[s].v);
                const long long i = __ldcg((const long long*)&sl[s].idx);
                cl_take(b, v, i);
            }
            b = cl_warp_best(b);
            if (g.nranks > 1) {
                // ---------- exchange with the other ranks through peer memory ----------
                // CTA 0 posts this rank's candidate into every rank's mailbox (lane r -> rank r);
                // EVERY CTA then reads its own device's mailbox until all candidates of this
                // iteration have arrived, and takes the global minimum in ndarray.argmin order.
                b.v = __shfl_sync(0xffffffffu, b.v, 0);
                b.idx = __shfl_sync(0xffffffffu, (long long)b.idx, 0);
                const unsigned long long want = g.seq_base + (unsigned long long)nm + 1ULL;
                const int slot = (int)(nm & 1) * g.nranks;
                if (blockIdx.x == 0 && lane < g.nranks) {
                    ClMail* dst = g.mbox[lane] + slot + g.rank;
                    dst->v = b.v; dst->idx = b.idx;
                    cl_st_release_sys(&dst->seq, want);
                }
                ClBest got{d_inf(), INT64_MAX};
                bool late = false;
                if (lane < g.nranks) {
                    const ClMail* src = g.mbox[g.rank] + slot + lane;
                    long long spins = 0;
                    while (cl_ld_acquire_sys(&src->seq) != want) {
                        if (++spins > 20000000LL || *((volatile int*)g.err)) { late = true; break; }
                    }
                    got.v = *((volatile const double*)&src->v);
                    got.idx = *((volatile const long long*)&src->idx);
                }
                if (__any_sync(0xffffffffu, late)) {
                    if (lane == 0) *g.err = 1;
                    got.v = d_inf(); got.idx = INT64_MAX;          // stops the loop below (a == b)
                }
                b = cl_warp_best(got);
            }
            if (lane == 0) gbest = b;
        }
        __syncthreads();
        const double mind = gbest.v;
        const int64_t bi = gbest.idx / n, bj = gbest.idx - (gbest.idx / n) * n;
        const int64_t a = bi < bj ? bi : bj, b = bi < bj ? bj : bi;
        // ---------- stop test, spk-clustering.py:207-208 ----------
        const bool go = (mind <= g.threshold) || (g.max_spk > 0 && nalive > (int64_t)g.max_spk);
        if (!go || a == b || gbest.idx == INT64_MAX) {
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                if (g.dbg) { g.dbg[0] = t_scan; g.dbg[1] = t_b1; g.dbg[2] = t_pick; g.dbg[3] = t_score; g.dbg[4] = t_b2; g.dbg[5] = nm; }
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
        const long long c3 = clock64();
        for (int q = threadIdx.x; q < REC; q += CL_THREADS)
            sm.merged[q] = __ldcg(g.rec + a * REC + q) + __ldcg(g.rec + b * REC + q);
        if (threadIdx.x == 0) sm.abits[b >> 5] &= ~(1u << (b & 31));
        __syncthreads();
        const double ld_ab = cl_phase_apply<KL2>(g, nm, a, b, sm, warp, lane, gwarp, nwarps);
        const long long c4 = clock64();
        cl_grid_barrier(g.bar, bar_target);
        t_scan += c1 - c0; t_b1 += c2 - c1; t_pick += c3 - c2; t_score += c4 - c3; t_b2 += clock64() - c4;
        // ---------- commit (CTA 0): the merged record replaces a's ----------
        if (blockIdx.x == 0) {
            for (int q = threadIdx.x; q < REC; q += CL_THREADS) g.rec[a * REC + q] = sm.merged[q];
            if (threadIdx.x == 0) g.ld[a] = ld_ab;
            if constexpr (KL2) {
                for (int q = threadIdx.x; q < KS; q += CL_THREADS) g.kside[a * KS + q] = (q % VS) < D39 ? sm.kside[q] : 0.0;
                if (threadIdx.x < D39) g.ksum[a * VS + threadIdx.x] = sm.ksum[threadIdx.x];
                if (threadIdx.x == 0) {                      // a's turns, then b's (spk-clustering.py:218: extend)
                    g.knext[g.ktail[a]] = g.khead[b];
                    g.ktail[a] = g.ktail[b];
                }
            }
        }
        if (g.rowlog && nm < g.rowlog_cap)                  // test hook: the rewritten row, for the host's argmin replay
            for (int64_t cidx = (int64_t)blockIdx.x * CL_THREADS + threadIdx.x; cidx < n; cidx += (int64_t)gridDim.x * CL_THREADS)
                g.rowlog[nm * n + cidx] = __ldcg(g.M + a * n + cidx);
        --nalive;
        ++nm;
    }
    if (g.nranks > 1 && blockIdx.x == 0 && warp == 0 && !*((volatile int*)g.err)) {
        // one more round: (max, min) over every finite distance any rank computed.  All ranks left the
        // loop at the same iteration, so the sequence number nm + 2 agrees.  The round has its OWN slots (the
        // third set): a rank that has collected it and returns may start the next run and post that run's first
        // candidate (parity-0 slots) while a slower peer is still reading this round - the two never share a slot,
        // and nobody can reach the next run's statistics round before every rank has left this run.
        const unsigned long long want = g.seq_base + (unsigned long long)nm + 2ULL;
        const int slot = 2 * g.nranks;
        const unsigned long long kmax = __ldcg(g.stat + 0), kmin = __ldcg(g.stat + 1);
        if (lane < g.nranks) {
            ClMail* dst = g.mbox[lane] + slot + g.rank;
            dst->v = __longlong_as_double((long long)kmax); dst->idx = (long long)kmin;
            cl_st_release_sys(&dst->seq, want);
        }
        unsigned long long gmax = kmax, gmin = kmin;
        bool late = false;
        if (lane < g.nranks) {
            const ClMail* src = g.mbox[g.rank] + slot + lane;
            long long spins = 0;
            while (cl_ld_acquire_sys(&src->seq) != want) { if (++spins > 20000000LL) { late = true; break; } }
            gmax = (unsigned long long)__double_as_longlong(*((volatile const double*)&src->v));
            gmin = (unsigned long long)*((volatile const long long*)&src->idx);
        }
        if (__any_sync(0xffffffffu, late)) {
            if (lane == 0) *g.err = 1;
        } else {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {                       // ordered keys: plain integer max / min
                const unsigned long long om = __shfl_xor_sync(0xffffffffu, gmax, o), on = __shfl_xor_sync(0xffffffffu, gmin, o);
                gmax = om > gmax ? om : gmax; gmin = on < gmin ? on : gmin;
            }
            if (lane == 0) { g.stat[0] = gmax; g.stat[1] = gmin; }
        }
    }
    if (blockIdx.x == 0) {
        __syncthreads();
        for (int64_t i = threadIdx.x; i < n; i += CL_THREADS) g.alive_out[i] = (sm.abits[i >> 5] >> (i & 31)) & 1u;
    }
}

// ---- the same two phases as separate launches (row-sharded run: the host exchanges the ranks'
// candidates between them) ----
static __global__ void __launch_bounds__(CL_THREADS, 1) cl_shard_argmin(const ClDev g, long long nm) {
    if (g.stopped && *((volatile int*)g.stopped)) return;
    extern __shared__ __align__(16) unsigned char cl_smem[];
    const ClSmem sm = cl_carve(cl_smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t gwarp = (int64_t)warp * gridDim.x + blockIdx.x;
    const int64_t nwarps = (int64_t)gridDim.x * CL_WARPS;
    const int nwords = (int)((g.n + 31) / 32);
    for (int wd = threadIdx.x; wd < nwords; wd += CL_THREADS) sm.abits[wd] = g.abits_g[wd];
    // commit the merge of the previous iteration (nobody reads records in this kernel)
    if (blockIdx.x == 0) {
        const int64_t pa = (int64_t)g.pend[REC + 1] - 1;
        if (pa >= 0) {
            for (int q = threadIdx.x; q < REC; q += CL_THREADS) g.rec[pa * REC + q] = g.pend[q];
            if (threadIdx.x == 0) g.ld[pa] = g.pend[REC];
        }
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) g.pend[REC + 1] = 0.0;
    const ClBest cta = cl_phase_argmin(g, nm, sm, warp, lane, gwarp, nwarps);
    __shared__ bool last;
    if (threadIdx.x == 0) {
        g.slots[blockIdx.x] = cta;
        __threadfence();
        last = atomicAdd(g.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && warp == 0) {
        __threadfence();
        ClBest b{d_inf(), INT64_MAX};
        for (int s = lane; s < (int)gridDim.x; s += 32) {
            const double v = __ldcg(&g.slots[s].v);
            const long long i = __ldcg((const long long*)&g.slots[s].idx);
            cl_take(b, v, i);
        }
        b = cl_warp_best(b);
        if (lane == 0) { *g.local_best = b; *g.ticket = 0u; }
    }
}

static __global__ void __launch_bounds__(CL_THREADS, 1) cl_shard_apply(const ClDev g, long long nm, int64_t a, int64_t b) {
    extern __shared__ __align__(16) unsigned char cl_smem[];
    const ClSmem sm = cl_carve(cl_smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t gwarp = (int64_t)warp * gridDim.x + blockIdx.x;
    const int64_t nwarps = (int64_t)gridDim.x * CL_WARPS;
    const int nwords = (int)((g.n + 31) / 32);
    for (int wd = threadIdx.x; wd < nwords; wd += CL_THREADS) {
        uint32_t w = g.abits_g[wd];
        if (wd == (int)(b >> 5)) w &= ~(1u << (b & 31));            // whoever reads it before CTA 0 stores it
        sm.abits[wd] = w;
    }
    for (int q = threadIdx.x; q < REC; q += CL_THREADS)
        sm.merged[q] = __ldcg(g.rec + a * REC + q) + __ldcg(g.rec + b * REC + q);
    __syncthreads();
    const double ld_ab = cl_phase_apply(g, nm, a, b, sm, warp, lane, gwarp, nwarps);
    if (blockIdx.x == 0) {
        // the merged record is committed by the next argmin launch (other CTAs still read rec[a] here)
        for (int q = threadIdx.x; q < REC; q += CL_THREADS) g.pend[q] = sm.merged[q];
        if (threadIdx.x == 0) {
            g.pend[REC] = ld_ab;
            g.pend[REC + 1] = (double)(a + 1);
            atomicAnd(g.abits_g + (b >> 5), ~(1u << (b & 31)));
        }
    }
}

// cl_shard_apply with the decision taken on the device: every thread picks the global minimum of
// the gathered candidates (ndarray.argmin order) and evaluates the stop test of CL1:207-208
static __global__ void __launch_bounds__(CL_THREADS, 1) cl_shard_apply_dev(const ClDev g, long long nm) {
    if (*((volatile int*)g.stopped)) return;
    ClBest best{__ldcg(&g.gathered[0].v), __ldcg((const long long*)&g.gathered[0].idx)};
    for (int r = 1; r < g.nranks; ++r)
        cl_take(best, __ldcg(&g.gathered[r].v), __ldcg((const long long*)&g.gathered[r].idx));
    const int64_t n = g.n;
    const double mind = best.v;
    const int64_t bi = best.idx / n, bj = best.idx - (best.idx / n) * n;
    const int64_t a = bi < bj ? bi : bj, b = bi < bj ? bj : bi;
    const int64_t nalive = n - nm;
    const bool go = (mind <= g.threshold) || (g.max_spk > 0 && nalive > (int64_t)g.max_spk);
    if (!go || a == b || best.idx == INT64_MAX) {
        if (blockIdx.x == 0 && threadIdx.x == 0) { *g.final_min = mind; __threadfence(); *g.stopped = 1; }
        return;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (nm < g.cap) { spkdiar_merge mr; mr.a = (int32_t)a; mr.b = (int32_t)b; mr.d = mind; g.out[nm] = mr; }
        *g.nmerge = nm + 1;
        if (mind > g.det[0]) g.det[0] = mind;               // spk-clustering.py:210-213
        if (mind < g.det[1]) g.det[1] = mind;
    }
    extern __shared__ __align__(16) unsigned char cl_smem[];
    const ClSmem sm = cl_carve(cl_smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t gwarp = (int64_t)warp * gridDim.x + blockIdx.x;
    const int64_t nwarps = (int64_t)gridDim.x * CL_WARPS;
    const int nwords = (int)((g.n + 31) / 32);
    for (int wd = threadIdx.x; wd < nwords; wd += CL_THREADS) {
        uint32_t w = g.abits_g[wd];
        if (wd == (int)(b >> 5)) w &= ~(1u << (b & 31));
        sm.abits[wd] = w;
    }
    for (int q = threadIdx.x; q < REC; q += CL_THREADS)
        sm.merged[q] = __ldcg(g.rec + a * REC + q) + __ldcg(g.rec + b * REC + q);
    __syncthreads();
    const double ld_ab = cl_phase_apply(g, nm, a, b, sm, warp, lane, gwarp, nwarps);
    if (blockIdx.x == 0) {
        for (int q = threadIdx.x; q < REC; q += CL_THREADS) g.pend[q] = sm.merged[q];
        if (threadIdx.x == 0) {
            g.pend[REC] = ld_ab;
            g.pend[REC + 1] = (double)(a + 1);
            atomicAnd(g.abits_g + (b >> 5), ~(1u << (b & 31)));
        }
    }
}

// initial fill of the pairs this rank owns (nranks > 1)
static __global__ void __launch_bounds__(SC_THREADS, 3)
cl_fill_pairs_shard(const double* rec, const double* __restrict__ ld, int64_t n, int metric, double lambda,
                    int rank, int nranks, double* __restrict__ M, unsigned long long* stat) {
    extern __shared__ __align__(16) unsigned char sc_smem[];
    WarpScratch* ws = reinterpret_cast<WarpScratch*>(sc_smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // owned pairs of row i: j = i + 1 + ((rank - 2 i - 1) mod nranks) + t * nranks; rows are dealt to the warps
    // cyclically, a row's owned pairs in turn
    const int64_t gw = (int64_t)blockIdx.x * SC_WARPS + warp, nw = (int64_t)gridDim.x * SC_WARPS;
    // linear index over (i, t): row i has cnt(i) = number of j > i with (i + j) % nranks == rank
    ClTrack track;
    for (int64_t i = 0; i < n - 1; ++i) {
        int64_t j0 = i + 1 + ((((int64_t)rank - 2 * i - 1) % nranks) + nranks) % nranks;
        // warps stride over the owned pairs of the row, offset by the row so that short rows do not
        // always start at warp 0
        for (int64_t t = (gw + nw - (i % nw)) % nw; j0 + t * nranks < n; t += nw) {
            const int64_t j = j0 + t * nranks;
            const RecSrc X{rec + i * REC}, Y{rec + j * REC};
            const double d = cl_pair_distance(metric, lambda, X, Y, ld[i], ld[j], ws[warp], lane);
            if (lane == 0) { M[i * n + j] = d; M[j * n + i] = d; track.see(d); }
        }
    }
    if (lane == 0) track.flush(stat);
}

// spk-clustering2.py:220: distances.max() over the compacted matrix (NaN propagates)
static __global__ void cl_alive_max(const double* __restrict__ M, const uint8_t* __restrict__ alive, int64_t n,
                             unsigned long long* out /* [0]=ordered max, [1]=nan flag */) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / n, c = i - r * n;
        if (!alive[r] || !alive[c]) continue;
        const double v = M[i];
        if (v != v) atomicExch(out + 1, 1ULL);
        else atomicMax(out + 0, cl_ord(v));
    }
}

inline size_t cl_smem_bytes(int64_t n) {
    return CL_WARPS * sizeof(WarpScratch) + REC * sizeof(double) + CL_WARPS * sizeof(ClBest) + 2 * sizeof(double)
           + CL_WARPS * sizeof(int) + KS * sizeof(double) + VS * sizeof(float)
           + (size_t)((n + 31) / 32) * sizeof(uint32_t) + 16;
}

cudaError_t cluster_small_configure();         // cluster_small.cuh
cudaError_t cluster_inorder_configure();       // cluster_inorder.cuh
cudaError_t cluster_set_dim(int d) { return set_dim_symbol(d); }

cudaError_t cluster_configure() {
    cudaError_t e = cudaFuncSetAttribute(cl_merge_loop<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(cl_merge_loop<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return e;
    e = cluster_inorder_configure();
    if (e != cudaSuccess) return e;
    e = cluster_small_configure();
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(cl_shard_argmin, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(cl_shard_apply, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(cl_shard_apply_dev, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(cl_fill_pairs_shard, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SC_SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(cl_kl2_init, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SC_SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(cl_self_logdet, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SC_SMEM);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(cl_fill_pairs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SC_SMEM);
}

}  // namespace spk
