// cluster_inorder.cuh - spk_cluster_in (spk-clustering.py:136-175, spk-clustering2.py:135-170) with the line loop on
// the device.
//
// In-order clustering takes the recipe lines one after the other: the new segment is scored against every speaker
// found so far, joins the nearest one if that distance is at most the threshold, else becomes a new speaker.  Every
// line depends on the decisions before it, so the host-driven form (clustering.py::_cluster_in with one
// spkdiar_score_sets call and one synchronisation per line) pays a launch + a round trip per line for a handful of
// factorisations.  Here ONE CTA walks all lines of a recording: speaker records (sum of the records of their
// segments, added in the order the segments joined - the order spkdiar_score_sets sums a set's ranges) and their
// ln|S| stay resident, a line costs one round of factorisations (one warp per speaker, plus one for the segment) and
// two __syncthreads.  The kernel writes every distance it computed (the host replays the script's bookkeeping -
// -tt lines, max / min statistics, the `d` quirk of spk-clustering.py:164-167 - from them) and the decision per
// line; same device functions (logdet_term, bic_combine / glr_combine) as the scoring kernels, so the distances
// are those of the host-driven loop bit for bit.  BIC and GLR; KL2 stays host-driven.
#pragma once

#include "cluster.cuh"

namespace spk {

struct InDev {
    Stats st;
    int64_t nlines;
    const int64_t* la; const int64_t* lb;      // [nlines] frame range of every line
    int metric; double lambda; double threshold;
    double* srec;                              // [nspk0 + nlines][REC] speaker records (the first nspk0 filled by the caller)
    double* sld;                               // [nspk0 + nlines] ln|S| of the speakers
    double* pooled;                            // [nspk0 + nlines] scratch: pooled / mix term of the current line
    int32_t nspk0;
    double* dist; int64_t dist_cap;            // distances of line l at dist[first[l] .. first[l + 1])
    int64_t* first;                            // [nlines + 1]
    int32_t* best;                             // [nlines] speaker joined, -1: new speaker
    int* err;                                  // 1: dist_cap too small
};

inline size_t cl_inorder_smem_bytes() { return CL_WARPS * sizeof(WarpScratch) + REC * sizeof(double) + 64; }

static __global__ void __launch_bounds__(CL_THREADS, 1) cl_inorder_kernel(const InDev g) {
    extern __shared__ __align__(16) unsigned char in_smem[];
    WarpScratch* ws = reinterpret_cast<WarpScratch*>(in_smem);
    double* segrec = reinterpret_cast<double*>(in_smem + CL_WARPS * sizeof(WarpScratch));
    __shared__ double s_ldseg;
    __shared__ int s_best, s_stop;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int nspk = g.nspk0;
    for (int s = warp; s < nspk; s += CL_WARPS) {
        const RecSrc X{g.srec + (size_t)s * REC};
        const double v = logdet_term(0, g.metric, X, X, ws[warp], lane);
        if (lane == 0) g.sld[s] = v;
    }
    if (threadIdx.x == 0) { g.first[0] = 0; s_stop = 0; }
    __syncthreads();
    int64_t off = 0;
    for (int64_t l = 0; l < g.nlines; ++l) {
        {
            const WinSrc w(g.st, g.la[l], g.lb[l], REC);
            for (int q = threadIdx.x; q < REC; q += CL_THREADS) segrec[q] = w(q);
        }
        __syncthreads();
        const SmemSrc Y{segrec};
        if (nspk > 0 && off + nspk > g.dist_cap) {
            if (threadIdx.x == 0) { *g.err = 1; s_stop = 1; }
        }
        __syncthreads();
        if (s_stop) return;
        // one round of factorisations: the segment itself, and the pooled (BIC) / mixed (GLR) matrix per speaker
        for (int t = warp; t < nspk + 1; t += CL_WARPS) {
            if (t == 0) {
                const double v = logdet_term(1, g.metric, Y, Y, ws[warp], lane);
                if (lane == 0) s_ldseg = v;
            } else {
                const RecSrc X{g.srec + (size_t)(t - 1) * REC};
                const double v = logdet_term(2, g.metric, X, Y, ws[warp], lane);
                if (lane == 0) g.pooled[t - 1] = v;
            }
        }
        __syncthreads();
        const double N2 = segrec[L39::CNT], ldseg = s_ldseg;
        for (int s = threadIdx.x; s < nspk; s += CL_THREADS) {
            const double N1 = g.srec[(size_t)s * REC + L39::CNT];
            const double ld1 = g.sld[s], ldp = g.pooled[s];
            g.dist[off + s] = g.metric == SPKDIAR_BIC ? bic_combine(N1, N2, ld1, ldseg, ldp, g.lambda)
                                                      : glr_combine(N1, N2, ld1, ldseg, ldp);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            // spk-clustering.py:147-160: the nearest speaker among the distances that are not +-inf (strict <, the
            // first wins; a NaN never wins), starting from sys.maxint
            double mind = CL_MAXINT_D;
            int best = -1;
            for (int s = 0; s < nspk; ++s) {
                const double d = g.dist[off + s];
                if (d == d_inf() || d == -d_inf()) continue;
                if (d < mind) { mind = d; best = s; }
            }
            if (!(nspk > 0 && mind <= g.threshold)) best = -1;
            s_best = best;
            g.best[l] = best;
            g.first[l + 1] = off + nspk;
        }
        __syncthreads();
        const int best = s_best;
        off += nspk;
        if (best >= 0) {
            // the segment joins speaker `best`: its record grows by the segment's, its ln|S| is recomputed
            double* r = g.srec + (size_t)best * REC;
            for (int q = threadIdx.x; q < REC; q += CL_THREADS) r[q] = r[q] + segrec[q];
            __syncthreads();
            if (warp == 0) {
                const RecSrc X{r};
                const double v = logdet_term(0, g.metric, X, X, ws[0], lane);
                if (lane == 0) g.sld[best] = v;
            }
        } else {
            double* r = g.srec + (size_t)nspk * REC;
            for (int q = threadIdx.x; q < REC; q += CL_THREADS) r[q] = 0.0 + segrec[q];     // as a set of one range is summed
            if (threadIdx.x == 0) g.sld[nspk] = ldseg;      // the same matrix, the same factorisation
            nspk += 1;
        }
        __syncthreads();
    }
}

cudaError_t cluster_inorder_configure() {
    return cudaFuncSetAttribute(cl_inorder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cl_inorder_smem_bytes());
}

}  // namespace spk
