// layout.cuh - (1) the layout of a sufficient-statistics record in HBM and
// (2) which lane of a warp owns which entry of the DxD matrix it factorises.
//
// (1) Every record (frame prefix records written by the statistics kernels,
// per-cluster records of the clustering engine) is RECORD = TRI + D + 1 doubles:
//
//   [0, TRI)        second moments  sum (x-c)_r (x-c)_k,  k <= r, packed row-major:
//                   entry (r, k) at r (r + 1) / 2 + k
//   [TRI, TRI + D)  first moments   sum (x-c)_k
//   [TRI + D]       frame count
//
// (2) The factorisation is a right-looking LDL^T whose trailing submatrix shrinks
// with every step, so a lane that owns whole rows or columns idles more and more.
// The lower triangle is therefore dealt out BLOCK-CYCLICALLY over the warp seen as
// a 4 x 8 grid: lane (i, j), i = lane / 8, j = lane % 8, owns the entries (r, k)
// with r % 4 == i and k % 8 == j.  At every step the live entries are spread evenly
// over all 32 lanes (a lane updates at most 30 of its own entries in the first step,
// about n_live / 32 later), where a row-per-lane layout needs 58 updates per lane on
// 20 lanes.  A lane keeps its entries in statically indexed registers: local row
// ri = r / 4 (0..9), local column kj = k / 8 (0..4); only the slots with ri >= 2 kj
// can lie on or below the diagonal, 30 slots in all.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace spk {

template <int D>
struct Layout {
    static constexpr int TRI = D * (D + 1) / 2;
    static constexpr int REC = TRI + D + 1;        // doubles per record
    static constexpr int VEC = TRI;                // offset of first moments
    static constexpr int CNT = TRI + D;            // offset of frame count
    __host__ __device__ static constexpr int pos(int r, int c) { return (r * (r + 1)) / 2 + c; }   // c <= r
    __host__ __device__ static constexpr int pos_diag(int j) { return (j * (j + 3)) / 2; }
};

template <int D>
struct Grid {
    static constexpr int PR = 4, PC = 8;                       // lane grid
    static constexpr int NRI = (D + PR - 1) / PR;              // local rows    (10 for D = 39)
    static constexpr int NKJ = (D + PC - 1) / PC;              // local columns ( 5 for D = 39)
    static constexpr int RPK = PC / PR;                        // local rows per local column step (2)
    // slots of local column kj: local rows RPK * kj .. NRI - 1
    __host__ __device__ static constexpr int off(int kj) { return kj * NRI - RPK * (kj * (kj - 1)) / 2; }
    static constexpr int NSLOT = off(NKJ);                     // 30 for D = 39
    __host__ __device__ static constexpr int slot(int kj, int ri) { return off(kj) + ri - RPK * kj; }
    __host__ __device__ static constexpr int ri_first(int kj) { return RPK * kj; }
};

static_assert(Layout<39>::REC == 820, "record size");
static_assert(Grid<39>::NSLOT == 30, "slots per lane");
static_assert(Grid<39>::NRI * Grid<39>::PR >= 39 && Grid<39>::NKJ * Grid<39>::PC >= 39, "grid covers the matrix");

}  // namespace spk
