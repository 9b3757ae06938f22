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

cudaError_t cd_merge_chain_configure();
cudaError_t cluster_inorder_configure() {
    cudaError_t e = cudaFuncSetAttribute(cl_inorder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cl_inorder_smem_bytes());
    return e != cudaSuccess ? e : cd_merge_chain_configure();
}

}  // namespace spk

// ---- merge mode of the change detector (spk-change-detection.py:136-177, 375-394) on the device ----
// `-m m` walks the recipe lines of a wav: the previous segment (possibly already merged with its predecessors) is
// scored against the next line; below the threshold the two are merged (the previous one now ends where the next
// ends), else the previous one is written and the next takes its place.  One dependent decision per line; here one
// CTA does the whole chain (three warps: left, right and pooled / mixed term of a step; thread 0 decides).
// BIC goes through the reference's memo of the first left term (the mutable default of `bic`):
// c1_memo in/out, NaN = not set yet, `use_memo` = --bic-cache reference.
namespace spk {

struct MgDev {
    Stats st;
    int64_t nlines;
    const int64_t* la; const int64_t* lb;      // [nlines] clamped frame bounds of every line (start, end)
    int metric; double lambda; double threshold; int use_memo;
    double* terms;                             // [nlines - 1][3] left, right, pooled / mixed ln|S| of step k
    double* dist;                              // [nlines - 1]
    int32_t* merged;                           // [nlines - 1] 1: line k + 1 was merged into the previous segment
    double* memo;                              // [1] c1 of the reference's memo (NaN: empty)
};

static __global__ void __launch_bounds__(CL_THREADS, 1) cd_merge_chain_kernel(const MgDev g) {
    extern __shared__ __align__(16) unsigned char in_smem[];
    WarpScratch* ws = reinterpret_cast<WarpScratch*>(in_smem);
    __shared__ double s_t[3];
    __shared__ long long s_pa, s_pb;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { s_pa = g.la[0]; s_pb = g.lb[0]; }
    __syncthreads();
    double memo = *g.memo;
    for (int64_t k = 0; k + 1 < g.nlines; ++k) {
        const int64_t a1 = s_pa, b1 = s_pb > s_pa ? s_pb : s_pa;
        const int64_t a2 = g.la[k + 1], b2 = g.lb[k + 1] > a2 ? g.lb[k + 1] : a2;
        if (warp < 3) {
            const WinSrc X(g.st, a1, b1, REC), Y(g.st, a2, b2, REC);
            const double v = logdet_term(warp, g.metric, X, Y, ws[warp], lane);
            if (lane == 0) s_t[warp] = v;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const double N1 = (double)(b1 - a1), N2 = (double)(b2 - a2);
            double d;
            if (g.metric == SPKDIAR_BIC) {
                double c1 = __dmul_rn(__dmul_rn(0.5, N1), s_t[0]);
                if (g.use_memo) { if (memo != memo) memo = c1; else c1 = memo; }
                const double N = N1 + N2;
                const double t0 = __dmul_rn(__dmul_rn(0.5, N), s_t[2]);
                const double t2 = __dmul_rn(__dmul_rn(0.5, N2), s_t[1]);
                d = __dsub_rn(__dsub_rn(__dsub_rn(t0, c1), t2), bic_pen(g.lambda, N));
            } else {
                d = glr_combine(N1, N2, s_t[0], s_t[1], s_t[2]);
            }
            const bool merge = d < g.threshold && d != d_inf() && d != -d_inf();       // CD:163-166
            g.terms[3 * k] = s_t[0]; g.terms[3 * k + 1] = s_t[1]; g.terms[3 * k + 2] = s_t[2];
            g.dist[k] = d;
            g.merged[k] = merge ? 1 : 0;
            if (merge) s_pb = g.lb[k + 1];                                              // prev now ends where next ends
            else { s_pa = g.la[k + 1]; s_pb = g.lb[k + 1]; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *g.memo = memo;
}

cudaError_t cd_merge_chain_configure() {
    return cudaFuncSetAttribute(cd_merge_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cl_inorder_smem_bytes());
}

}  // namespace spk
