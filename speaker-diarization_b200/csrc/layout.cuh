// layout.cuh - the "lane-paired" packed layout of a symmetric DxD matrix.
//
// Every sufficient-statistics record in HBM (frame prefix records written by
// the statistics kernels, per-cluster records of the clustering engine) is
// RECORD = TRI + D + 1 doubles:
//
//   [0, TRI)        second moments  sum (x-c)_r (x-c)_k   (lower triangle, k <= r)
//   [TRI, TRI + D)  first moments   sum (x-c)_k
//   [TRI + D]       frame count
//
// The triangle is NOT stored row-major.  It is stored in the order the
// warp-level LDL^T factorisation wants to read it: lane l (l < NL) owns the
// two rows  rh = D-1-l ("hi", D-l entries)  and  l ("lo", l+1 entries, only
// when l < rh), so that every lane owns about the same number of entries, and
// register slot hi[k] / lo[k] of consecutive lanes sit at consecutive
// addresses:
//
//   lo[k] of lane l (k <= l < NLO)      at  off_lo(k) + (l - k)
//   hi[k] of lane l (l <= min(NL-1, D-1-k))  at  off_hi(k) + l
//
// A warp therefore loads a whole record with coalesced 8-byte loads into
// statically indexed registers - no shared-memory staging, no shuffles.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace spk {

template <int D>
struct Layout {
    static constexpr int NL  = (D + 1) / 2;        // lanes that own rows
    static constexpr int NLO = D / 2;              // lanes that also own a "lo" row
    static constexpr int TRI = D * (D + 1) / 2;
    static constexpr int REC = TRI + D + 1;        // doubles per record
    static constexpr int VEC = TRI;                // offset of first moments
    static constexpr int CNT = TRI + D;            // offset of frame count

    __host__ __device__ static constexpr int off_lo(int k) {
        return k * NLO - (k * (k - 1)) / 2;
    }
    static constexpr int TLO = NLO * (NLO + 1) / 2;
    __host__ __device__ static constexpr int cnt_hi(int k) {
        return (D - k) < NL ? (D - k) : NL;
    }
    __host__ __device__ static constexpr int off_hi(int k) {
        int o = TLO;
        for (int j = 0; j < k; ++j) o += cnt_hi(j);
        return o;
    }
    // closed form of off_hi for run-time k
    __host__ __device__ static constexpr int off_hi_rt(int k) {
        constexpr int K0 = D - NL + 1;             // columns 0..K0-1 are held by all NL lanes
        return k <= K0 ? TLO + k * NL
                       : TLO + K0 * NL + (k - K0) * D - ((K0 + k - 1) * (k - K0)) / 2;
    }
    __host__ __device__ static constexpr bool off_hi_rt_ok() {
        for (int k = 0; k <= D; ++k) if (off_hi_rt(k) != off_hi(k)) return false;
        return true;
    }
    // position of the diagonal element (j, j), run-time j
    __host__ __device__ static constexpr int pos_diag(int j) {
        return (j < D - 1 - j) ? off_lo(j) : off_hi_rt(j) + (D - 1 - j);
    }
    // position of element (r, c), c <= r
    __host__ __device__ static constexpr int pos(int r, int c) {
        return (r < D - 1 - r) ? off_lo(c) + (r - c)          // a "lo" row of lane r
                               : off_hi(c) + (D - 1 - r);     // a "hi" row of lane D-1-r
    }
};

static_assert(Layout<39>::off_hi(39) == Layout<39>::TRI, "layout must tile the triangle");
static_assert(Layout<39>::REC == 820, "record size");
static_assert(Layout<39>::off_hi_rt_ok(), "closed form of off_hi");

}  // namespace spk
