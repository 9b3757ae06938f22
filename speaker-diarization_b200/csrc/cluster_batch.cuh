// cluster_batch.cuh - many SMALL agglomerative clustering problems in one launch.
//
// BASELINE config 4 (spk-diarization2.py over a corpus): every recording ends in
// spk-clustering.py -m hi over its own few dozen turns.  One such problem cannot fill
// a GPU - the resident engine of cluster.cuh spreads its merge loop over all SMs and
// pays two grid barriers per merge -, but a corpus has hundreds of them and they are
// independent.  Here ONE CTA runs ONE problem from the segment list to the last merge
// (records, ln|S_i|, pair matrix, merge loop; __syncthreads instead of grid barriers),
// CTAs pull problems from a queue, and a launch covers a whole batch of recordings.
//
// Arithmetic, operand order and the argmin rule (NaN first, value, flat index over the
// ORIGINAL indices with an alive mask) are those of cluster.cuh, through the same
// device functions: merge sequences, distances and statistics are bit-identical to
// spkdiar_cluster_run on the same segments (tests/test_gpu_batch.py).
#pragma once

#include "cluster.cuh"

namespace spk {

struct ClBatchDev {
    Stats st;
    const int64_t* first;        // [nprob + 1] problem p owns segments first[p] .. first[p + 1]
    const int64_t* seg_a;        // packed frame ranges of the initial clusters
    const int64_t* seg_b;
    int32_t nprob;
    int32_t nmax;                // largest problem
    int metric; double lambda; double threshold; int max_spk; int variant;
    // workspaces, one set per CTA
    double* rec;                 // [grid][nmax][REC]
    double* ld;                  // [grid][nmax]
    double* M;                   // [grid][nmax * nmax]  (row stride = n of the current problem)
    double* t;                   // [grid][nmax]         pooled terms of the current rescoring
    int32_t* next;               // problem queue cursor
    // results
    spkdiar_merge* out;          // [first[nprob]]: the merges of problem p start at first[p] (original indices)
    int64_t* nmerge;             // [nprob]
    double* stats;               // [nprob][4]
};

__device__ __forceinline__ double cl_unord_dev(unsigned long long k) {      // inverse of cl_ord
    const unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffULL) : ~k;
    return __longlong_as_double((long long)u);
}

inline size_t cl_batch_smem_bytes(int64_t nmax) {
    return CL_WARPS * sizeof(WarpScratch) + REC * sizeof(double) + CL_WARPS * sizeof(ClBest)
           + 4 * sizeof(unsigned long long) + (size_t)((nmax + 31) / 32) * sizeof(uint32_t) + 16;
}

static __global__ void __launch_bounds__(CL_THREADS, 1) cl_batch_kernel(const ClBatchDev g) {
    extern __shared__ __align__(16) unsigned char cl_smem[];
    WarpScratch* ws = reinterpret_cast<WarpScratch*>(cl_smem);
    double* merged = reinterpret_cast<double*>(cl_smem + CL_WARPS * sizeof(WarpScratch));
    ClBest* wbest = reinterpret_cast<ClBest*>(merged + REC);
    unsigned long long* sstat = reinterpret_cast<unsigned long long*>(wbest + CL_WARPS);   // [0] max [1] min (ordered keys), [2] NaN seen
    uint32_t* abits = reinterpret_cast<uint32_t*>(sstat + 4);
    __shared__ int s_prob;
    __shared__ ClBest gbest;
    __shared__ double s_ldab;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* rec = g.rec + (size_t)blockIdx.x * g.nmax * REC;
    double* ld = g.ld + (size_t)blockIdx.x * g.nmax;
    double* M = g.M + (size_t)blockIdx.x * g.nmax * g.nmax;
    double* tt = g.t + (size_t)blockIdx.x * g.nmax;

    int prob = blockIdx.x;
    while (prob < g.nprob) {
        const int64_t f0 = g.first[prob];
        const int n = (int)(g.first[prob + 1] - f0);
        const int nwords = (n + 31) / 32;
        // ---- records of the initial clusters (cl_init_records) ----
        for (int s = warp; s < n; s += CL_WARPS) {
            const WinSrc w(g.st, g.seg_a[f0 + s], g.seg_b[f0 + s], REC);
            for (int q = lane; q < REC; q += 32) rec[(size_t)s * REC + q] = w(q);
        }
        for (int wd = threadIdx.x; wd < nwords; wd += CL_THREADS) {
            const int lo = wd * 32;
            abits[wd] = (n - lo >= 32) ? 0xffffffffu : ((1u << (n - lo)) - 1u);
        }
        if (threadIdx.x == 0) {
            // max_dist = 0, min_dist = sys.maxint as ordered keys (spk-clustering.py:416-417)
            sstat[0] = cl_ord(0.0); sstat[1] = cl_ord(CL_MAXINT_D); sstat[2] = 0ULL; sstat[3] = 0ULL;
        }
        // ---- pair matrix: constant fill (spk-clustering.py:185-187 / spk-clustering2.py:178) ----
        for (int i = threadIdx.x; i < n * n; i += CL_THREADS) {
            const int r = i / n, c = i - r * n;
            M[i] = g.variant == 1 ? (r == c ? CL_MAXINT_D : 0.0) : d_inf();
        }
        __syncthreads();
        // ---- ln|S_i| (cl_self_logdet) ----
        for (int s = warp; s < n; s += CL_WARPS) {
            const RecSrc X{rec + (size_t)s * REC};
            const double v = logdet_term(0, SPKDIAR_BIC, X, X, ws[warp], lane);
            if (lane == 0) ld[s] = v;
        }
        __syncthreads();
        // ---- initial fill (cl_fill_pairs) ----
        const int64_t npair = ((int64_t)n * (n - 1)) / 2;
        for (int64_t p = warp; p < npair; p += CL_WARPS) {
            int64_t i, j;
            cl_pair(p, n, i, j);
            const RecSrc X{rec + i * REC}, Y{rec + j * REC};
            const double d = cl_pair_distance(g.metric, g.lambda, X, Y, __ldcg(ld + i), __ldcg(ld + j), ws[warp], lane);
            if (lane == 0) {
                M[i * n + j] = d;
                if (g.variant == 1) { M[j * n + i] = d; cl_track(d, sstat); }
            }
        }
        __syncthreads();
        // ---- merge loop ----
        int nalive = n;
        long long nm = 0;
        double det_max = 0.0, det_min = CL_MAXINT_D;           // spk-clustering.py:418-419
        double final_min = 0.0;
        for (;;) {
            // exact argmin over the alive part of the matrix
            ClBest mine{d_inf(), INT64_MAX};
            for (int i = threadIdx.x; i < n * n; i += CL_THREADS) {
                const int r = i / n, c = i - r * n;
                if (!((abits[r >> 5] >> (r & 31)) & 1u) || !((abits[c >> 5] >> (c & 31)) & 1u)) continue;
                cl_take(mine, __ldcg(M + i), (int64_t)i);
            }
            mine = cl_warp_best(mine);
            if (lane == 0) wbest[warp] = mine;
            __syncthreads();
            if (threadIdx.x == 0) {
                ClBest bb = wbest[0];
                for (int w = 1; w < CL_WARPS; ++w) cl_take(bb, wbest[w].v, wbest[w].idx);
                gbest = bb;
            }
            __syncthreads();
            const double mind = gbest.v;
            const int64_t bi = gbest.idx / n, bj = gbest.idx - (gbest.idx / n) * n;
            const int64_t a = bi < bj ? bi : bj, b = bi < bj ? bj : bi;
            // stop test, spk-clustering.py:207-208
            const bool go = (mind <= g.threshold) || (g.max_spk > 0 && nalive > g.max_spk);
            if (!go || a == b || gbest.idx == INT64_MAX) { final_min = mind; break; }
            if (mind > det_max) det_max = mind;                 // spk-clustering.py:210-213
            if (mind < det_min) det_min = mind;
            if (threadIdx.x == 0) {
                spkdiar_merge mr; mr.a = (int32_t)a; mr.b = (int32_t)b; mr.d = mind;
                g.out[f0 + nm] = mr;
                abits[b >> 5] &= ~(1u << (b & 31));
            }
            for (int q = threadIdx.x; q < REC; q += CL_THREADS)
                merged[q] = __ldcg(rec + a * REC + q) + __ldcg(rec + b * REC + q);
            __syncthreads();
            // rescoring: task 0 = ln|S_ab|, task 1 + k = pooled term of (ab, k) for every alive k != a
            const SmemSrc X{merged};
            for (int task = warp; task <= n; task += CL_WARPS) {
                if (task == 0) {
                    const RecSrc Y{rec + a * REC};
                    const double v = logdet_term(0, g.metric, X, Y, ws[warp], lane);
                    if (lane == 0) s_ldab = v;
                } else {
                    const int k = task - 1;
                    if (k == (int)a || !((abits[k >> 5] >> (k & 31)) & 1u)) continue;
                    const RecSrc Y{rec + (size_t)k * REC};
                    const double v = logdet_term(2, g.metric, X, Y, ws[warp], lane);
                    if (lane == 0) tt[k] = v;
                }
            }
            __syncthreads();
            const double ld_ab = s_ldab;
            const double N1 = merged[L39::CNT];
            for (int k = threadIdx.x; k < n; k += CL_THREADS) {
                if (k == (int)a || !((abits[k >> 5] >> (k & 31)) & 1u)) continue;
                const double N2 = __ldcg(rec + (size_t)k * REC + L39::CNT);
                const double ldk = __ldcg(ld + k);
                const double tk = __ldcg(tt + k);
                const double d = g.metric == SPKDIAR_BIC ? bic_combine(N1, N2, ld_ab, ldk, tk, g.lambda)
                                                         : glr_combine(N1, N2, ld_ab, ldk, tk);
                M[a * n + k] = d;                                                // row a
                if (g.variant == 1) { M[(int64_t)k * n + a] = d; cl_track(d, sstat); }   // and column a
            }
            for (int q = threadIdx.x; q < REC; q += CL_THREADS) rec[a * REC + q] = merged[q];
            if (threadIdx.x == 0) ld[a] = ld_ab;
            __syncthreads();
            --nalive;
            ++nm;
        }
        // ---- results ----
        if (g.variant == 2) {
            // spk-clustering2.py:220: distances.max() over the compacted matrix (NaN propagates)
            for (int i = threadIdx.x; i < n * n; i += CL_THREADS) {
                const int r = i / n, c = i - r * n;
                if (!((abits[r >> 5] >> (r & 31)) & 1u) || !((abits[c >> 5] >> (c & 31)) & 1u)) continue;
                const double v = __ldcg(M + i);
                if (v != v) atomicExch(sstat + 2, 1ULL);
                else atomicMax(sstat + 3, cl_ord(v));
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            g.nmerge[prob] = nm;
            double* so = g.stats + 4 * (size_t)prob;
            if (g.variant == 1) {
                so[0] = cl_unord_dev(sstat[0]);
                so[1] = cl_unord_dev(sstat[1]);
                so[2] = det_max; so[3] = det_min;
            } else {
                so[0] = sstat[2] ? d_nan() : cl_unord_dev(sstat[3]);
                so[1] = final_min; so[2] = 0.0; so[3] = CL_MAXINT_D;
            }
            s_prob = (int)gridDim.x + atomicAdd(g.next, 1);
        }
        __syncthreads();
        prob = s_prob;
        __syncthreads();
    }
}

cudaError_t cluster_batch_configure() {
    return cudaFuncSetAttribute(cl_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
}

}  // namespace spk
