// ldl.cuh - one warp factorises one DxD symmetric positive-definite matrix.
//
// This is the arithmetic core that replaces  np.log(scipy.linalg.det(np.cov(.)))
// of the reference (spk-change-detection.py:87-96,107-115; spk-clustering.py:
// 91-96,107-115): the covariance of a window is formed from a DIFFERENCE of
// sufficient-statistics records, factorised as L diag(p) L^T without square
// roots, and  ln|S| = sum ln p_c - D ln(n-1).
//
// Register layout: see layout.cuh.  Lane l holds row rh = D-1-l in hi[0..D-1]
// (entries k <= rh are meaningful) and row l in lo[0..NLO-1] (k <= l, only when
// l < NLO).  The factorisation is right-looking: at step c every lane
// publishes its entry of column c to a 2x(D+1)-double shared-memory strip, one
// __syncwarp later all lanes read the pivot and the column back as broadcast
// loads and apply the rank-1 update to their own rows from registers.  Slots a
// lane does not own carry garbage that is never published.
#pragma once

#include "layout.cuh"

namespace spk {

constexpr int VS = 40;     // strip stride in doubles (>= D + 1, keeps 16-byte alignment)

constexpr int REC39 = Layout<39>::REC;

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
// task: lane-strided, fully coalesced, 8 independent element loads in flight per
// lane and almost no live registers - whereas loading straight into the 58-double
// register tile of the factorisation left one load in flight at a time (ncu: 62 % of
// the stall samples on the consuming DADDs, profiles/r01_win_terms_before.txt).
template <class Src>
__device__ __forceinline__ const double* stage_record(const Src& src, double* buf, int lane) {
#pragma unroll 1
    for (int q0 = 0; q0 < REC39; q0 += 32 * 8) {
        double t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int q = q0 + 32 * u + lane;
            t[u] = q < REC39 ? src(q) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int q = q0 + 32 * u + lane;
            if (q < REC39) buf[q] = t[u];
        }
    }
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
// Returns the frame count the matrix stands for (nx, ny, nx+ny; 0 for MIX).
enum { FORM_X = 0, FORM_Y = 1, FORM_POOL = 2, FORM_MIX = 3 };

template <int D, class SrcX, class SrcY>
__device__ __forceinline__ double form_matrix(double (&hi)[D], double (&lo)[Layout<D>::NLO > 0 ? Layout<D>::NLO : 1],
                                              int kind, const SrcX& x, const SrcY& y, double wx, double wy,
                                              LdlScratch& w, int lane) {
    using L = Layout<D>;
    const bool ux = kind != FORM_Y, uy = kind != FORM_X;
    const bool two = kind == FORM_MIX;                 // second rank-1 correction
    const double nx = ux ? x(L::CNT) : 0.0, ny = uy ? y(L::CNT) : 0.0;
    // v1 / v2: the first-moment vectors of the rank-1 corrections
    for (int j = lane; j < D; j += 32) {
        const double sx = ux ? x(L::VEC + j) : 0.0, sy = uy ? y(L::VEC + j) : 0.0;
        w.s0[j] = two ? sx : sx + sy;
        w.s1[j] = two ? sy : 0.0;
    }
    __syncwarp();
    double ax = ux ? 1.0 : 0.0, ay = uy ? 1.0 : 0.0, c1, c2 = 0.0;
    if (two) { ax = wx; ay = wy; c1 = wx / nx; c2 = wy / ny; }
    else c1 = 1.0 / (nx + ny);
    const int rh = (D - 1 - lane) >= 0 ? (D - 1 - lane) : 0;
    const int rl = lane < D ? lane : 0;
    const double u1h = w.s0[rh] * c1, u2h = w.s1[rh] * c2;
    const double u1l = w.s0[rl] * c1, u2l = w.s1[rl] * c2;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        const int q = L::off_hi(k) + lane;
        double m = 0.0;
        if (ux) m = ax * x(q);
        if (uy) m = fma(ay, y(q), m);
        m = fma(-u1h, w.s0[k], m);
        if (two) m = fma(-u2h, w.s1[k], m);
        hi[k] = m;
    }
#pragma unroll
    for (int k = 0; k < L::NLO; ++k) {
        const int q = L::off_lo(k) - k + lane;
        double m = 0.0;
        if (ux) m = ax * x(q);
        if (uy) m = fma(ay, y(q), m);
        m = fma(-u1l, w.s0[k], m);
        if (two) m = fma(-u2l, w.s1[k], m);
        lo[k] = m;
    }
    __syncwarp();
    return two ? 0.0 : nx + ny;
}

// ---- factorisation ---------------------------------------------------------------
// Returns ln|M| (sum of the logs of the D pivots), NaN when a pivot is not > 0.
// STORE: also leave the strict lower triangle of L (row-major, row r at
// r(r-1)/2) in Lsm and the reciprocal pivots in pinv (both shared memory).
// One elimination step, column C known at compile time (template recursion: nvcc does
// not fully unroll a 39-trip loop with this much body, and a rolled loop would index
// the register arrays dynamically, i.e. push them to local memory).
template <int D, bool STORE, int C>
struct LdlStep {
    using L = Layout<D>;
    static __device__ __forceinline__ void run(double (&hi)[D], double (&lo)[L::NLO > 0 ? L::NLO : 1],
                                               LdlScratch& w, int lane, double& p0, double& p1, bool& bad,
                                               double* Lsm, double* pinv) {
        const int rh = D - 1 - lane;           // my hi row (valid when lane < NL)
        double* v = w.v[C & 1];
        if (lane < L::NL && rh >= C) v[rh] = hi[C];
        if (C < L::NLO) { if (lane < L::NLO && lane >= C) v[lane] = lo[C < L::NLO ? C : 0]; }
        __syncwarp();
        const double piv = v[C];
        bad |= !(piv > 0.0);
        if ((C & 31) == lane) { if (C < 32) p0 = piv; else p1 = piv; }
        const double r = __drcp_rn(piv);
        const double lh = hi[C] * r;
        if (STORE) {
            if (lane < L::NL && rh > C) Lsm[(rh * (rh - 1)) / 2 + C] = lh;
            if (lane == 0) pinv[C] = r;
        }
#pragma unroll
        for (int c2 = C + 1; c2 < D; ++c2) hi[c2] = fma(-lh, v[c2], hi[c2]);
        if (C < L::NLO) {
            const double ll = lo[C < L::NLO ? C : 0] * r;
            if (STORE) { if (lane < L::NLO && lane > C) Lsm[(lane * (lane - 1)) / 2 + C] = ll; }
#pragma unroll
            for (int c2 = C + 1; c2 < L::NLO; ++c2) lo[c2] = fma(-ll, v[c2], lo[c2]);
        }
        LdlStep<D, STORE, C + 1>::run(hi, lo, w, lane, p0, p1, bad, Lsm, pinv);
    }
};
template <int D, bool STORE>
struct LdlStep<D, STORE, D> {
    using L = Layout<D>;
    static __device__ __forceinline__ void run(double (&)[D], double (&)[L::NLO > 0 ? L::NLO : 1], LdlScratch&, int,
                                               double&, double&, bool&, double*, double*) {}
};

template <int D, bool STORE>
__device__ __forceinline__ double ldl_logdet(double (&hi)[D], double (&lo)[Layout<D>::NLO > 0 ? Layout<D>::NLO : 1],
                                             LdlScratch& w, int lane,
                                             double* Lsm = nullptr, double* pinv = nullptr) {
    double p0 = 1.0, p1 = 1.0;             // pivots c == lane and c == lane + 32
    bool bad = false;
    LdlStep<D, STORE, 0>::run(hi, lo, w, lane, p0, p1, bad, Lsm, pinv);
    double s = log(p0);
    if (D > 32) s += log(p1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __syncwarp();
    return bad ? __longlong_as_double(0x7ff8000000000000LL) : s;
}

// ln|S| of the reference from ln|M|:  S = M / (n - 1).  Also applies the range
// mapping of np.log(det(S)): a determinant that under/overflows
// fp64 makes the reference see -inf / +inf.
__device__ __forceinline__ double finish_logdet(double ln_m, double n, int d) {
    double v = ln_m - (double)d * log(n - 1.0);
    if (v < -744.4400719213812) v = -__longlong_as_double(0x7ff0000000000000LL);
    else if (v > 709.782712893384) v = __longlong_as_double(0x7ff0000000000000LL);
    return v;
}

// ---- diag(M^-1) from the stored factor (KL2 only) --------------------
// X = L^-1 is unit lower triangular; column j of X is owned by lane l as the
// pair (j = l, rows l..D-1) and (j = D-1-l, rows D-1-l..D-1).
//   diag(M^-1)_j = sum_{r >= j} X_rj^2 / p_r
// Returns the value for column `lane` in ga and for column D-1-lane in gb.
template <int D, int R>
struct InvRow {
    using L = Layout<D>;
    static constexpr int HB = D - L::NL;
    static __device__ __forceinline__ void run(const double* Lsm, int lane, double (&xa)[D], double (&xb)[L::NL]) {
        const double* row = Lsm + (R * (R - 1)) / 2;
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            if (k & 1) a1 = fma(row[k], xa[k], a1); else a0 = fma(row[k], xa[k], a0);
        }
        // rows above the column's own diagonal stay 0; the diagonal stays 1
        if (R > lane) xa[R] = -(a0 + a1);
        if (R > HB) {
            double b0 = 0.0, b1 = 0.0;
#pragma unroll
            for (int k = HB; k < R; ++k) {
                if (k & 1) b1 = fma(row[k], xb[k - HB], b1); else b0 = fma(row[k], xb[k - HB], b0);
            }
            if (R > D - 1 - lane) xb[R > HB ? R - HB : 0] = -(b0 + b1);
        }
        InvRow<D, R + 1>::run(Lsm, lane, xa, xb);
    }
};
template <int D>
struct InvRow<D, D> {
    using L = Layout<D>;
    static __device__ __forceinline__ void run(const double*, int, double (&)[D], double (&)[L::NL]) {}
};

template <int D>
__device__ __forceinline__ void inv_diag(const double* Lsm, const double* pinv, int lane,
                                         double& ga, double& gb) {
    using L = Layout<D>;
    constexpr int HB = D - L::NL;          // first row of the short columns (19 for D = 39)
    double xa[D];                          // column `lane`, indexed by row
    double xb[L::NL];                      // column D-1-lane, rows HB..D-1 -> index r - HB
#pragma unroll
    for (int r = 0; r < D; ++r) xa[r] = (r == lane) ? 1.0 : 0.0;
#pragma unroll
    for (int r = HB; r < D; ++r) xb[r - HB] = (r == D - 1 - lane) ? 1.0 : 0.0;
    InvRow<D, 1>::run(Lsm, lane, xa, xb);
    ga = 0.0; gb = 0.0;
#pragma unroll
    for (int r = 0; r < D; ++r) ga = fma(xa[r] * xa[r], pinv[r], ga);
#pragma unroll
    for (int r = HB; r < D; ++r) gb = fma(xb[r - HB] * xb[r - HB], pinv[r], gb);
}

}  // namespace spk
