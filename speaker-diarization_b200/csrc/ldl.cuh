// ldl.cuh - one warp factorises one DxD symmetric positive-definite matrix.
//
// This is the arithmetic core that replaces  np.log(scipy.linalg.det(np.cov(.)))
// of the reference (spk-change-detection.py:87-96,107-115; spk-clustering.py:
// 91-96,107-115): the covariance of a window is formed from a DIFFERENCE of
// sufficient-statistics records, factorised as L diag(p) L^T without square
// roots, and  ln|S| = sum ln p_c - D ln(n-1).
//
// Register layout: see layout.cuh (block-cyclic 8 x 4 lane grid, 30 slots per
// lane).  The factorisation is right-looking: at step c the owners of column c
// publish it to a 2x40-double shared-memory strip, one __syncwarp later all lanes
// read the pivot and the entries they need back and apply the rank-1 update to
// their own slots from registers.
#pragma once

#include "layout.cuh"

namespace spk {

constexpr int VS = 40;     // strip stride in doubles (>= D + 1, keeps 16-byte alignment)

constexpr int REC39 = Layout<39>::REC;

// Feature files of fewer than 39 dimensions (the reference's loader takes the dimension from the file header,
// spk-change-detection.py:37-41): frames are stored zero-padded to 39 columns, so every record, strip and
// register tile keeps its one compiled shape, and the matrix that is factorised gets an IDENTITY block for the
// padding: its diagonal is set to the value that makes S_kk = 1 there (form_matrix), so ln|S| and diag(S^-1) of
// the real block come out unchanged.  c_dim is the real dimension (BIC penalty, "fewer frames than dimensions");
// one value per device at a time, set by the upload (spkdiar.cu) in every translation unit.
static __constant__ int c_dim = 39;
static inline cudaError_t set_dim_symbol(int d) { return cudaMemcpyToSymbol(c_dim, &d, sizeof(int)); }

// per-warp shared-memory scratch of the factorisation
struct LdlScratch {
    double v[2][VS];       // column strips, double-buffered by step parity
    double s0[VS];         // first moments of operand X
    double s1[VS];         // first moments of operand Y (GLR mix only)
};
// ... plus the two operand records, staged from HBM/L2 before the matrix is formed
struct WarpScratch : LdlScratch {
    double rec[2][REC39];
};

// ---- operand sources ----------------------------------------------------------
// A source yields element q of a statistics record.

// frames [a, b) of a recording from the two-level statistics of stats.cuh:
// block-local prefix records plus the double-double block prefix.  Both arrays
// are immutable while any scoring kernel runs -> read-only path.
struct Stats {
    const double* P;       // (n + 1) block-local prefix records
    const double2* C;      // (nblocks + 1) double-double block prefix records
};
constexpr int STAT_BLOCK = 128;    // == K1_TILE
struct WinSrc {
    const double* __restrict__ pb;    // record at b
    const double* __restrict__ pa;    // record at a
    const double2* __restrict__ cb;   // block prefix of b's block
    const double2* __restrict__ ca;   // block prefix of a's block
    double n;                         // b - a
    bool cross;                       // a and b lie in different blocks
    __device__ __forceinline__ WinSrc(const Stats& st, long long a, long long b, int rec)
        : pb(st.P + b * rec), pa(st.P + a * rec),
          cb(st.C + (b / STAT_BLOCK) * rec), ca(st.C + (a / STAT_BLOCK) * rec),
          n((double)(b - a)), cross((b / STAT_BLOCK) != (a / STAT_BLOCK)) {}
    __device__ __forceinline__ double operator()(int q) const {
        double v = __ldg(pb + q) - __ldg(pa + q);
        if (cross) {
            const double2 hb = __ldg(cb + q), ha = __ldg(ca + q);
            v += (hb.x - ha.x) + (hb.y - ha.y);
        }
        return v;
    }
};
// a cluster record that a persistent kernel may have rewritten: bypass L1.
struct RecSrc {
    const double* p;
    __device__ __forceinline__ double operator()(int q) const { return __ldcg(p + q); }
};
// a record already in shared memory
struct SmemSrc {
    const double* p;
    __device__ __forceinline__ double operator()(int q) const { return p[q]; }
};
template <class A, class B>
struct SumSrc {
    A a; B b;
    __device__ __forceinline__ double operator()(int q) const { return a(q) + b(q); }
};

// ---- staging -------------------------------------------------------------------
// Stage one statistics record into shared memory.  This is the MEMORY phase of a
// task: lane-strided, fully coalesced, B independent element loads in flight per
// lane and almost no other live registers - whereas loading straight into the
// register tile of the factorisation left one load in flight at a time (ncu: 62 % of
// the stall samples on the consuming DADDs, profiles/r01_win_terms_before.txt).
template <int B, class Src>
__device__ __forceinline__ void stage_batches(const Src& src, double* buf, int lane) {
#pragma unroll 1
    for (int q0 = 0; q0 < REC39; q0 += 32 * B) {
        double t[B];
#pragma unroll
        for (int u = 0; u < B; ++u) {
            const int q = q0 + 32 * u + lane;
            t[u] = q < REC39 ? src(q) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < B; ++u) {
            const int q = q0 + 32 * u + lane;
            if (q < REC39) buf[q] = t[u];
        }
    }
}
template <class Src>
__device__ __forceinline__ const double* stage_record(const Src& src, double* buf, int lane) {
    stage_batches<13>(src, buf, lane);                // 26 elements per lane: two round trips
    return buf;
}
// a window of the two-level statistics: two loads per element inside one block, four
// (two of them 16 bytes wide) across blocks
__device__ __forceinline__ const double* stage_record(const WinSrc& src, double* buf, int lane) {
    if (src.cross) stage_batches<9>(src, buf, lane);
    else stage_batches<13>(src, buf, lane);
    return buf;
}
// a record that already sits in shared memory is used in place
__device__ __forceinline__ const double* stage_record(const SmemSrc& src, double*, int) { return src.p; }

// ---- forming the matrix to factorise, in registers --------------------------------
// One straight-line code path for every kind of matrix (a control-flow merge of
// four separately formed register arrays makes the compiler park them in local
// memory); the kinds differ only in warp-uniform weights and predicates:
//
//   FORM_X     M = Qx - sx sx^T / nx                         (left / cluster X)
//   FORM_Y     M = Qy - sy sy^T / ny                         (right / cluster Y)
//   FORM_POOL  M = (Qx+Qy) - (sx+sy)(sx+sy)^T / (nx+ny)      (BIC pooled term)
//   FORM_MIX   M = wx (Qx - sx sx^T/nx) + wy (Qy - sy sy^T/ny)   (GLR, CD:114-115)
//
// Lane (i, j) = (lane / PC, lane % PC) fills its slots (kj, ri) <-> entry (i + PR ri, j + PC kj),
// see layout.cuh.  Returns the frame count the matrix stands for (nx, ny, nx+ny; 0 for MIX).
enum { FORM_X = 0, FORM_Y = 1, FORM_POOL = 2, FORM_MIX = 3 };

template <int D, class SrcX, class SrcY>
__device__ __forceinline__ double form_matrix(double (&a)[Grid<D>::NSLOT],
                                              int kind, const SrcX& x, const SrcY& y, double wx, double wy,
                                              LdlScratch& w, int lane) {
    using L = Layout<D>;
    using G = Grid<D>;
    const bool ux = kind != FORM_Y, uy = kind != FORM_X;
    const bool two = kind == FORM_MIX;                 // second rank-1 correction
    const double nx = ux ? x(L::CNT) : 0.0, ny = uy ? y(L::CNT) : 0.0;
    // s0 / s1: the first-moment vectors of the rank-1 corrections
    for (int t = lane; t < VS; t += 32) {
        const double sx = (ux && t < D) ? x(L::VEC + t) : 0.0, sy = (uy && t < D) ? y(L::VEC + t) : 0.0;
        w.s0[t] = two ? sx : sx + sy;
        w.s1[t] = two ? sy : 0.0;
    }
    __syncwarp();
    double ax = ux ? 1.0 : 0.0, ay = uy ? 1.0 : 0.0, c1, c2 = 0.0;
    if (two) { ax = wx; ay = wy; c1 = wx / nx; c2 = wy / ny; }
    else c1 = 1.0 / (nx + ny);
    const int i = G::lane_i(lane), j = G::lane_j(lane);
    const int dim = c_dim;
    double u1[G::NRI], u2[G::NRI];
#pragma unroll
    for (int ri = 0; ri < G::NRI; ++ri) {
        u1[ri] = w.s0[i + G::PR * ri] * c1;            // index <= 39 < VS
        u2[ri] = w.s1[i + G::PR * ri] * c2;
    }
#pragma unroll
    for (int kj = 0; kj < G::NKJ; ++kj) {
        const int k = j + G::PC * kj;
        const double sk0 = w.s0[k], sk1 = w.s1[k];
#pragma unroll
        for (int ri = G::ri_first(kj); ri < G::NRI; ++ri) {
            const int r = i + G::PR * ri;
            const int q = (r < D && k <= r) ? L::pos(r, k) : 0;       // slots above the diagonal / outside: any value
            double m = 0.0;
            if (ux) m = ax * x(q);
            if (uy) m = fma(ay, y(q), m);
            m = fma(-u1[ri], sk0, m);
            if (two) m = fma(-u2[ri], sk1, m);
            a[G::slot(kj, ri)] = m;
        }
    }
    // identity block of the zero-padded dimensions of a file with fewer than D of them (a warp-uniform branch that
    // files of D dimensions skip; only slots that can lie on the diagonal test anything: r - k = PR ri - PC kj +
    // (i - j) with -PC < i - j < PR)
    if (dim < D) {
        const double pad = two ? 1.0 : nx + ny - 1.0;        // M = (n - 1) S, or S itself for the GLR mix
#pragma unroll
        for (int kj = 0; kj < G::NKJ; ++kj) {
#pragma unroll
            for (int ri = G::ri_first(kj); ri < G::NRI; ++ri) {
                if (G::PC * kj - G::PR * ri > -G::PC && G::PC * kj - G::PR * ri < G::PR) {
                    const int r = i + G::PR * ri, k = j + G::PC * kj;
                    if (r == k && k >= dim) a[G::slot(kj, ri)] = pad;
                }
            }
        }
    }
    __syncwarp();
    return two ? 0.0 : nx + ny;
}

// ---- factorisation ---------------------------------------------------------------
// Returns ln|M| (sum of the logs of the D pivots), NaN when a pivot is not > 0.
// One elimination step, column C known at compile time (template recursion: nvcc does
// not fully unroll a 39-trip loop with this much body, and a rolled loop would index
// the register array dynamically, i.e. push it to local memory):
//   the owners of column C publish it to the shared-memory strip; every lane reads the
//   pivot, the column entries of its own rows (scaled by 1/pivot) and of its own columns,
//   and applies the rank-1 update to its live slots.  Slots that are finished or lie
//   outside the triangle receive garbage that is never published.
// predicated shared-memory store without a branch: the elimination steps stay one
// straight-line block, so the instruction scheduler can overlap the tail of one step
// (updates nobody waits for) with the latency chain of the next (strip round trip,
// reciprocal, first multiply)
__device__ __forceinline__ void st_shared_if(bool p, double* addr, double v) {
    asm volatile("{\n.reg .pred q;\nsetp.ne.b32 q, %0, 0;\n@q st.shared.f64 [%1], %2;\n}\n"
                 ::"r"((int)p), "r"((unsigned)__cvta_generic_to_shared(addr)), "d"(v) : "memory");
}
// 1 / a to within an ulp without the slow-path branch of __drcp_rn: hardware seed y
// (20 bits), e = 1 - a y, 1 / a = y (1 + e + e^2 + O(e^3)).  a <= 0, inf, NaN or
// subnormal give inf / NaN, which the caller reports as a failed factorisation.
__device__ __forceinline__ double fast_rcp(double a) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    const double e = fma(-a, y, 1.0);
    const double f = fma(e, e, e);
    return fma(y, f, y);
}

// does lane row-group i of local row ri hold a row r = i + PR ri with C < r < D?  After unrolling, ri and C
// are constants: only the boundary group (rows around C) and the last group (rows around D) test anything
template <int D>
__device__ __forceinline__ bool live_row(bool p, int i, int ri, int C) {
    using G = Grid<D>;
    const int lo = G::PR * ri;
    if (lo <= C) p = p && (i > C - lo);
    if (lo + G::PR - 1 >= D) p = p && (i < D - lo);
    return p;
}

// One elimination step.  On entry the lane already holds what it needs of column C
// (read from the strip by the previous step): the pivot, the entries vr[] of its own
// rows and vk[] of its own columns, and q = 1 / pivot is on its way.  The step
// (1) updates column C + 1 and publishes it, (2) reads its part of column C + 1 back
// and starts the next reciprocal, (3) applies the rest of the rank-1 update.
// Measured on B200: one warp alone is a latency chain (strip round trip -> reciprocal ->
// multiply -> FMA per step); twelve warps on an SM are bound by the shared-memory pipe -
// 2 cycles per store instruction + 1 per load accounted for 11.9 k of the 12.3 k cycles of
// the 4 x 8 layout - hence the layout with the fewest stores (layout.cuh) and no traffic
// beyond the column itself (a look-ahead of the next pivot through an extra strip slot
// saved 9 % alone and cost 4 pipe cycles per step under load: removed).
template <int D, int C>
struct LdlStep {
    using G = Grid<D>;
    static __device__ __forceinline__ void run(double (&a)[G::NSLOT], LdlScratch& w, int lane, int i, int j,
                                               const double (&vr)[G::NRI], const double (&vk)[G::NKJ],
                                               double piv, double q, double& p0, double& p1) {
        constexpr int ri_a = (C + 1) / G::PR;          // first local row that can hold a row > C
        constexpr int kj_a = (C + 1) / G::PC;          // first local column that can hold a column > C
        constexpr int kn = (C + 1) / G::PC, jn = (C + 1) % G::PC;     // local column / owner lanes of column C + 1
        double* vn = w.v[(C + 1) & 1];
        if ((C & 31) == lane) { if (C < 32) p0 = piv; else p1 = piv; }
        double lr[G::NRI];
#pragma unroll
        for (int ri = ri_a; ri < G::NRI; ++ri) lr[ri] = vr[ri] * q;
        double nvr[G::NRI], nvk[G::NKJ], npiv = 1.0, nq = 1.0;
        if (C + 1 < D) {
            // (1) column C + 1 first: update it and hand it over at once
            const bool pj = j == jn;
#pragma unroll
            for (int ri = (ri_a > G::ri_first(kn) ? ri_a : G::ri_first(kn)); ri < G::NRI; ++ri) {
                const double m = fma(-lr[ri], vk[kn], a[G::slot(kn, ri)]);
                a[G::slot(kn, ri)] = m;
                st_shared_if(live_row<D>(pj, i, ri, C), vn + i + G::PR * ri, m);
            }
            __syncwarp();
            // (2) what the next step needs of column C + 1
            constexpr int ri_n = (C + 2) / G::PR, kj_n = (C + 2) / G::PC;
            npiv = vn[C + 1];
#pragma unroll
            for (int ri = ri_n; ri < G::NRI; ++ri) nvr[ri] = vn[i + G::PR * ri];
#pragma unroll
            for (int kj = kj_n; kj < G::NKJ; ++kj) nvk[kj] = vn[j + G::PC * kj];
            nq = fast_rcp(npiv);
            // (3) the rest of the update of step C
#pragma unroll
            for (int kj = kj_a; kj < G::NKJ; ++kj) {
                if (kj == kn) continue;
#pragma unroll
                for (int ri = (ri_a > G::ri_first(kj) ? ri_a : G::ri_first(kj)); ri < G::NRI; ++ri)
                    a[G::slot(kj, ri)] = fma(-lr[ri], vk[kj], a[G::slot(kj, ri)]);
            }
        }
        LdlStep<D, C + 1>::run(a, w, lane, i, j, nvr, nvk, npiv, nq, p0, p1);
    }
};
template <int D>
struct LdlStep<D, D> {
    using G = Grid<D>;
    static __device__ __forceinline__ void run(double (&)[G::NSLOT], LdlScratch&, int, int, int,
                                               const double (&)[G::NRI], const double (&)[G::NKJ], double, double,
                                               double&, double&) {}
};

template <int D>
__device__ __forceinline__ double ldl_logdet(double (&a)[Grid<D>::NSLOT], LdlScratch& w, int lane) {
    using G = Grid<D>;
    static_assert(D <= 64 && G::PR * G::NRI <= VS && G::PC * G::NKJ <= VS, "pivots in two registers per lane, strip covers the grid");
    double p0 = 1.0, p1 = 1.0;             // pivots c == lane and c == lane + 32
    const int i = G::lane_i(lane), j = G::lane_j(lane);
    __syncwarp();                          // the strips may still be read by a slower lane of the previous task
#pragma unroll
    for (int ri = 0; ri < G::NRI; ++ri)    // prologue: publish column 0 and read it back
        st_shared_if(live_row<D>(j == 0, i, ri, -1), w.v[0] + i + G::PR * ri, a[G::slot(0, ri)]);
    __syncwarp();
    double vr[G::NRI], vk[G::NKJ];
    const double piv = w.v[0][0];
#pragma unroll
    for (int ri = 0; ri < G::NRI; ++ri) vr[ri] = w.v[0][i + G::PR * ri];
#pragma unroll
    for (int kj = 0; kj < G::NKJ; ++kj) vk[kj] = w.v[0][j + G::PC * kj];
    LdlStep<D, 0>::run(a, w, lane, i, j, vr, vk, piv, fast_rcp(piv), p0, p1);
    double s = log(p0);
    if (D > 32) s += log(p1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    // A pivot that is not > 0: the matrix is singular to working precision (a cluster of fewer frames than
    // dimensions, a constant feature) or the input was not finite.  The reference takes np.log(det(S)) of such a
    // matrix: LAPACK's determinant underflows to 0 (-inf, and the pair distance +inf, which the agglomeration
    // ignores) or comes out as rounding noise of either sign - nothing to reproduce bit for bit.  Ruling: the
    // FIRST pivot that fails decides - NaN if it is NaN (non-finite input), -inf otherwise (det -> 0); pivots
    // after it are garbage either way.  (A NaN here would stop the whole agglomeration: ndarray.argmin returns
    // the first NaN and `mind <= threshold` is False, spk-clustering.py:203-208.)
    const unsigned m0 = __ballot_sync(0xffffffffu, !(p0 > 0.0)), m1 = __ballot_sync(0xffffffffu, !(p1 > 0.0));
    __syncwarp();
    if (m0 | m1) {
        const double first = m0 ? __shfl_sync(0xffffffffu, p0, __ffs(m0) - 1) : __shfl_sync(0xffffffffu, p1, __ffs(m1) - 1);
        return first != first ? __longlong_as_double(0x7ff8000000000000LL) : -__longlong_as_double(0x7ff0000000000000LL);
    }
    return s;
}

// ---- factorisation fused with the inverse of the factor (KL2) -------------
// KL2 needs diag(M^-1) = diag(L^-T D^-1 L^-1): diag(M^-1)_k = q_k + sum_{r > k} X_rk^2 q_r with
// X = L^-1 (unit lower triangular) and q = 1 / pivot.  X is accumulated DURING the elimination:
// step C applies X_r,: -= l_rC X_C,: to the rows r > C (X_rC starts as -l_rC), i.e. to the
// entries (r, k) with r > C > k - exactly the register slots the elimination has FINISHED with
// (column k < C of M is done).  So the same 30 slots per lane hold M's live part (k > C) and X's
// grown part (k < C), every step updates all of them with one FMA each, and the inverse costs no
// extra dependent chain: row C of X travels through the second pair of strips next to column C.
// On entry of step C the lane holds the pivot, vr[ri] = M_rC of its rows and cf[kj] = M_kC of its
// columns k > C or X_Ck of its columns k < C.
template <int D, int C>
struct LdlInvStep {
    using G = Grid<D>;
    static __device__ __forceinline__ void run(double (&a)[G::NSLOT], LdlScratch& w, int lane, int i, int j,
                                               const double (&vr)[G::NRI], const double (&cf)[G::NKJ],
                                               double piv, double q,
                                               double& p0, double& p1, bool& bad, double* pinv) {
        constexpr int kjb = C / G::PC, jC = C % G::PC;                // local column / owner lanes of column C
        constexpr int rib = C / G::PR, iC = C % G::PR;                // local row that holds row C
        constexpr int kn = (C + 1) / G::PC, jn = (C + 1) % G::PC;     // ... of column C + 1
        constexpr int rn = (C + 1) / G::PR, in = (C + 1) % G::PR;     // ... of row C + 1
        double* vn = w.v[(C + 1) & 1];
        double* xn = (C + 1) & 1 ? w.s1 : w.s0;                       // X row strips (free after form_matrix)
        bad |= !(piv > 0.0);
        if ((C & 31) == lane) { if (C < 32) p0 = piv; else p1 = piv; }
        st_shared_if(lane == 0, pinv + C, q);
        // l_rC of the lane's rows r > C (0 for r <= C: those slots are finished and must not move)
        double lr[G::NRI];
#pragma unroll
        for (int ri = rib; ri < G::NRI; ++ri) {
            const double l = vr[ri] * q;
            lr[ri] = (ri > rib || i > iC) ? l : 0.0;
        }
        // one update per slot (kj, ri), ri >= max(ri_first(kj), rib):
        //   k > C : M_rk -= l_rC M_kC       k < C : X_rk -= l_rC X_Ck       k == C : X_rC = -l_rC
        auto update = [&](int kj, int ri) {
            const int sl = G::slot(kj, ri);
            const double t = fma(-lr[ri], cf[kj], a[sl]);
            if (kj == kjb) a[sl] = (j == jC && (ri > rib || i > iC)) ? -lr[ri] : t;
            else a[sl] = t;
        };
        double nvr[G::NRI], ncf[G::NKJ], npiv = 1.0, nq = 1.0;
        if (C + 1 < D) {
            // first what the next step needs: local column kn (column C + 1 of M) and local row rn
            // (row C + 1 of X)
#pragma unroll
            for (int ri = (rib > G::ri_first(kn) ? rib : G::ri_first(kn)); ri < G::NRI; ++ri) update(kn, ri);
#pragma unroll
            for (int kj = 0; kj < G::NKJ; ++kj)
                if (kj != kn && rn >= G::ri_first(kj) && rn >= rib) update(kj, rn);
#pragma unroll
            for (int ri = (rib > G::ri_first(kn) ? rib : G::ri_first(kn)); ri < G::NRI; ++ri)
                st_shared_if(live_row<D>(j == jn, i, ri, C), vn + i + G::PR * ri, a[G::slot(kn, ri)]);
#pragma unroll
            for (int kj = 0; kj <= kn; ++kj) {
                if (rn < G::ri_first(kj)) continue;
                const int k = j + G::PC * kj;
                st_shared_if(i == in && k <= C, xn + k, a[G::slot(kj, rn)]);
            }
            __syncwarp();
            constexpr int ri_n = (C + 1) / G::PR;
            npiv = vn[C + 1];
#pragma unroll
            for (int ri = ri_n; ri < G::NRI; ++ri) nvr[ri] = vn[i + G::PR * ri];
#pragma unroll
            for (int kj = 0; kj < G::NKJ; ++kj) {
                const int k = j + G::PC * kj;
                // columns right of C + 1: M_k,C+1; left of it: X_C+1,k; the lane's own column C + 1: unused
                const double* src = (kj > kn || (kj == kn && j > jn)) ? vn : xn;
                ncf[kj] = src[k];
            }
            nq = fast_rcp(npiv);
            // then the rest
#pragma unroll
            for (int kj = 0; kj < G::NKJ; ++kj) {
                if (kj == kn) continue;
#pragma unroll
                for (int ri = (rib > G::ri_first(kj) ? rib : G::ri_first(kj)); ri < G::NRI; ++ri)
                    if (ri != rn) update(kj, ri);
            }
        }
        LdlInvStep<D, C + 1>::run(a, w, lane, i, j, nvr, ncf, npiv, nq, p0, p1, bad, pinv);
    }
};
template <int D>
struct LdlInvStep<D, D> {
    using G = Grid<D>;
    static __device__ __forceinline__ void run(double (&)[G::NSLOT], LdlScratch&, int, int, int,
                                               const double (&)[G::NRI], const double (&)[G::NKJ], double, double,
                                               double&, double&, bool&, double*) {}
};

// ln|M| and diag(M^-1) (into dinv[0..D-1], scaled by `scale`; NaN when a pivot is not > 0).
// pinv: VS doubles of shared memory.  Uses w.s0 / w.s1 as strips: form_matrix's vectors are gone after.
template <int D>
__device__ __forceinline__ double ldl_logdet_inv(double (&a)[Grid<D>::NSLOT], LdlScratch& w, int lane,
                                                 double* pinv, double scale, double* dinv) {
    using G = Grid<D>;
    double p0 = 1.0, p1 = 1.0;
    bool bad = false;
    const int i = G::lane_i(lane), j = G::lane_j(lane);
    __syncwarp();
#pragma unroll
    for (int ri = 0; ri < G::NRI; ++ri)    // prologue: publish column 0 and read it back (row 0 of X is empty)
        st_shared_if(live_row<D>(j == 0, i, ri, -1), w.v[0] + i + G::PR * ri, a[G::slot(0, ri)]);
    __syncwarp();
    double vr[G::NRI], cf[G::NKJ];
    const double piv = w.v[0][0];
#pragma unroll
    for (int ri = 0; ri < G::NRI; ++ri) vr[ri] = w.v[0][i + G::PR * ri];
#pragma unroll
    for (int kj = 0; kj < G::NKJ; ++kj) cf[kj] = w.v[0][j + G::PC * kj];
    LdlInvStep<D, 0>::run(a, w, lane, i, j, vr, cf, piv, fast_rcp(piv), p0, p1, bad, pinv);
    double s = log(p0);
    if (D > 32) s += log(p1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __syncwarp();                                              // pinv complete
    // diag(M^-1)_k = q_k + sum_{r > k} X_rk^2 q_r: per lane over its rows, then over the four lanes of a column
    double qr[G::NRI];
#pragma unroll
    for (int ri = 0; ri < G::NRI; ++ri) { const int r = i + G::PR * ri; qr[ri] = r < D ? pinv[r] : 0.0; }
#pragma unroll
    for (int kj = 0; kj < G::NKJ; ++kj) {
        const int k = j + G::PC * kj;
        double acc = 0.0;
#pragma unroll
        for (int ri = G::ri_first(kj); ri < G::NRI; ++ri) {
            const int r = i + G::PR * ri;
            const double x = (r > k && r < D) ? a[G::slot(kj, ri)] : 0.0;
            acc = fma(x * x, qr[ri], acc);
        }
#pragma unroll
        for (int o = G::PC; o < 32; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);      // over the PR lanes of a column
        if (i == 0 && k < D) dinv[k] = bad ? __longlong_as_double(0x7ff8000000000000LL) : (acc + pinv[k]) * scale;
    }
    __syncwarp();
    return bad ? __longlong_as_double(0x7ff8000000000000LL) : s;
}

// ln|S| of the reference from ln|M|:  S = M / (n - 1).  Also applies the range
// mapping of np.log(det(S)): a determinant that under/overflows
// fp64 makes the reference see -inf / +inf.
__device__ __forceinline__ double finish_logdet(double ln_m, double n, int d) {
    // no more frames than dimensions: the covariance is singular by construction (see ldl_logdet)
    if (n > 1.0 && n <= (double)c_dim) return -__longlong_as_double(0x7ff0000000000000LL);
    double v = ln_m - (double)d * log(n - 1.0);
    if (v < -744.4400719213812) v = -__longlong_as_double(0x7ff0000000000000LL);
    else if (v > 709.782712893384) v = __longlong_as_double(0x7ff0000000000000LL);
    return v;
}

}  // namespace spk
