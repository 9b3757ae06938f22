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
// a PR x PC = 8 x 4 grid: lane (i, j), i = lane / 4, j = lane % 4, owns the entries (r, k)
// with r % 8 == i and k % 4 == j.  At every step the live entries are spread evenly
// over all 32 lanes (a lane updates at most 30 of its own entries in the first step,
// about n_live / 32 later), where a row-per-lane layout needs 58 updates per lane on
// 20 lanes.  A lane keeps its entries in statically indexed registers: local row
// ri = r / 8 (0..4), local column kj = k / 4 (0..9); only the slots with ri >= kj / 2
// can lie on or below the diagonal, 30 slots in all.
// Why 8 x 4 and not 4 x 8 (the shape of round 1): with twelve warps on an SM the
// factorisation is bound by the shared-memory pipe, not by the fp64 pipe - the column of
// every step travels through a shared-memory strip, a store instruction occupies that pipe
// for two cycles however few lanes are active (measured: tests/micro/bench_smem.cu), and a
// column is published by the lanes that own it, one store instruction per local row.  Eight
// owner lanes with at most 5 local rows need 114 store instructions per factorisation where
// four lanes with 10 local rows needed 209; the loads (310) and the slots (30) are the same,
// the multiply-adds drop from 509 to 469 warp instructions and the scalings from 209 to 114.
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
    static constexpr int PR = 8, PC = 4;                       // lane grid: lane = i * PC + j
    static_assert(PR * PC == 32, "one warp");
    static constexpr int NRI = (D + PR - 1) / PR;              // local rows    ( 5 for D = 39)
    static constexpr int NKJ = (D + PC - 1) / PC;              // local columns (10 for D = 39)
    // slots of local column kj: local rows ri_first(kj) .. NRI - 1 (the others lie above the diagonal)
    __host__ __device__ static constexpr int ri_first(int kj) { return (kj * PC) / PR; }
    __host__ __device__ static constexpr int off(int kj) {
        int o = 0;
        for (int k = 0; k < kj; ++k) o += NRI - ri_first(k);
        return o;
    }
    static constexpr int NSLOT = off(NKJ);                     // 30 for D = 39
    __host__ __device__ static constexpr int slot(int kj, int ri) { return off(kj) + ri - ri_first(kj); }
    __device__ static __forceinline__ int lane_i(int lane) { return lane / PC; }
    __device__ static __forceinline__ int lane_j(int lane) { return lane % PC; }
};

static_assert(Layout<39>::REC == 820, "record size");
static_assert(Grid<39>::NSLOT == 30, "slots per lane");
static_assert(Grid<39>::NRI * Grid<39>::PR >= 39 && Grid<39>::NKJ * Grid<39>::PC >= 39, "grid covers the matrix");

}  // namespace spk
