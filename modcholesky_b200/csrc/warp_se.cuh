// warp_se.cuh -- SE90 / SE99 (se90.rs:38-181, se99.rs:35-201), one warp per matrix, the
// matrix resident in registers for the whole factorisation (n <= NMAX <= 32).
// See warp_common.cuh for the data layout.  HBM traffic per matrix is exactly the
// algorithmic 16 n^2 + 16 n bytes (+4 for phase2_start): A is read once with coalesced
// row-of-the-batch loads, L / e / p are written once.
#pragma once
#include "warp_common.cuh"

namespace modchol {

// Per-warp shared memory (doubles): two 32-entry column lines, plus (STRICT only) a padded
// copy of the trailing block used once per matrix by the ndarray-ordered Gershgorin sums.
template <int NMAX, bool kStrict>
__host__ __device__ constexpr int se_warp_smem_doubles()
{
    return 64 + (kStrict ? NMAX * (NMAX + 1) : 0);
}

template <int NMAX, bool kStrict>
struct SeWarp {
    double r[NMAX];
    double dg;      // current diagonal element of this lane's row
    double g;       // lower Gershgorin bound (phase 2)
    double e_mine;  // E entry of this lane's ORIGINAL index (se99.rs:194-198 un-permute is free)
    int pos;        // permuted position of this lane's row
    bool valid;     // lane < n
    // warp-uniform
    double gamma, taugamma, delta_prev;
    int n;
    double* buf0;   // column line (unscaled values, STRICT norms)
    double* buf1;   // column line (scaled values, trailing update)
};

// Pivot scaling: s = sqrt(piv) and the scaled column element cs = col / s
// (se90.rs:84-87, se99.rs:96-99).  FAST mode uses y = rsqrt(piv): s = piv*y, cs = col*y, for
// pivots in the safe range; every special value takes the reference's formulas.
template <bool kStrict>
__device__ __forceinline__ void se_scale(double piv, double col, double& s, double& cs,
                                         double& inv_piv)
{
    if (kStrict || !fast_range_ok(piv)) {
        s = dsqrt(piv);
        cs = ddiv(col, s);
        inv_piv = 0.0;  // unused on this path
    } else {
        const double y = rsqrt(piv);
        s = dmul(piv, y);
        cs = dmul(col, y);
        inv_piv = dmul(y, y);
    }
}

// Elimination step (se90.rs:84-93 == :142-151, se99.rs:96-105 == :162-171): pivot lane q,
// s = sqrt of the pivot (already including E), cs = this lane's scaled column element,
// dg_new = this lane's updated diagonal dg - cs*cs.
template <int J, int NMAX, bool kStrict>
__device__ __forceinline__ void se_eliminate(SeWarp<NMAX, kStrict>& w, int lane, int q, double s,
                                             double cs, double dg_new, bool below)
{
    if (lane == q)
        w.r[J] = s;
    else
        w.r[J] = cs;
    w.buf1[w.pos] = cs;
    __syncwarp();
    if (below)
        w.dg = dg_new;
    // r[k] -= cs * (scaled column element of the row at position k), k = J+1 .. n-1.
    // Columns are processed in groups so that padding groups (k >= n) are skipped.
    constexpr int K0 = J + 1;
    if constexpr ((K0 & 1) && K0 < NMAX) {
        w.r[K0] = nmul_add<kStrict>(w.r[K0], cs, w.buf1[K0]);
    }
    constexpr int K1 = (K0 + 1) & ~1;
    static_for<0, (NMAX - K1 + 7) / 8>([&](auto gi) {
        constexpr int base = K1 + 8 * decltype(gi)::value;
        if (base < w.n) {
#pragma unroll
            for (int k = base; k < base + 8 && k < NMAX; k += 2) {
                const double2 v = *reinterpret_cast<const double2*>(&w.buf1[k]);
                w.r[k] = nmul_add<kStrict>(w.r[k], cs, v.x);
                if (k + 1 < NMAX)
                    w.r[k + 1] = nmul_add<kStrict>(w.r[k + 1], cs, v.y);
            }
        }
    });
    __syncwarp();
}

// Pivot bookkeeping common to both phases: make position m the new position J.
template <int J, int NMAX, bool kStrict>
__device__ __forceinline__ int se_apply_pivot(SeWarp<NMAX, kStrict>& w, int m)
{
    if (m != J) {
        swap_columns_uniform<J, NMAX>(w.r, m);
        if (w.pos == J)
            w.pos = m;
        else if (w.pos == m)
            w.pos = J;
    }
    return __ffs(__ballot_sync(kFull, w.pos == J)) - 1;
}

// One phase-1 step.  Returns false when the factorisation must switch to phase 2 at J.
template <int kAlgo, int J, int NMAX, bool kStrict>
__device__ __forceinline__ bool se_phase1_step(SeWarp<NMAX, kStrict>& w, int lane)
{
    const bool active = w.valid && w.pos >= J;
    double dmax;
    const int m = warp_argmax_first(w.dg, active, w.pos, J, dmax);
    if (kAlgo == kSE99) {  // se99.rs:62-73 (before the pivot is applied)
        bool win;
        const double dmin = warp_extreme<false>(w.dg, active, INFINITY, win);
        if (dmax < w.taugamma || dmin < dmul(-MODCHOL_MU, dmax))
            return false;
    }
    const int q = se_apply_pivot<J, NMAX, kStrict>(w, m);  // se90.rs:63-68, se99.rs:76-81
    const double col = w.r[J];
    const double piv = __shfl_sync(kFull, w.dg, q);
    const bool below = w.valid && w.pos > J;
    double sq, cs, ip;
    se_scale<kStrict>(piv, col, sq, cs, ip);
    const double dg_new = nmul_add<kStrict>(w.dg, cs, cs);
    // look-ahead min_i (l_ii - l_ij^2 / l_jj)  (se90.rs:70-77, se99.rs:82-89).  In FAST mode
    // this is the updated diagonal itself (l_ij^2 / l_jj == cs^2 up to rounding).
    double nv;
    if (kStrict || !fast_range_ok(piv))
        nv = dsub(w.dg, ddiv(dmul(col, col), piv));
    else
        nv = dg_new;
    bool win;
    const double la = warp_extreme<false>(nv, below, INFINITY, win);
    const double thresh = (kAlgo == kSE99) ? dmul(-MODCHOL_MU, w.gamma) : w.taugamma;
    if (la < thresh)
        return false;  // the pivot at J stays applied (se99.rs:90-93)
    se_eliminate<J, NMAX, kStrict>(w, lane, q, sq, cs, dg_new, below);
    return true;
}

// Lower Gershgorin bounds of the trailing block starting at k (se90.rs:105-110,
// se99.rs:125-130): g_i = l_ii - S(|l_i,k..i-1|) - S(|l_i+1..n-1,i|), S = ndarray sum.
template <int NMAX, bool kStrict>
__device__ __forceinline__ void se_gershgorin_init(SeWarp<NMAX, kStrict>& w, int lane, int k,
                                                   double* rows)
{
    const int n = w.n;
    if constexpr (kStrict) {
        // every lane parks its row in shared memory (stride NMAX+1: conflict-free), then walks
        // its two index ranges with run-time offsets and compile-time partial-sum slots.
        double* row = rows + (lane < NMAX ? lane : 0) * (NMAX + 1);
        if (lane < NMAX) {
#pragma unroll
            for (int c = 0; c < NMAX; ++c)
                row[c] = w.r[c];
        }
        __syncwarp();
        double s[2];
#pragma unroll
        for (int part = 0; part < 2; ++part) {
            const int start = part == 0 ? k : w.pos + 1;
            int len = part == 0 ? w.pos - k : n - 1 - w.pos;
            if (!(w.valid && w.pos >= k))
                len = 0;
            const int chunked = len & ~7;
            double p[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                p[u] = 0.0;
#pragma unroll
            for (int t = 0; t < ((NMAX - 1) & ~7); ++t)
                if (t < chunked)
                    p[t & 7] = dadd(p[t & 7], fabs(row[start + t]));
            double acc = 0.0;
            acc = dadd(acc, dadd(p[0], p[4]));
            acc = dadd(acc, dadd(p[1], p[5]));
            acc = dadd(acc, dadd(p[2], p[6]));
            acc = dadd(acc, dadd(p[3], p[7]));
#pragma unroll
            for (int i = 0; i < 7; ++i)
                if (chunked + i < len)
                    acc = dadd(acc, fabs(row[start + chunked + i]));
            s[part] = acc;
        }
        w.g = dsub(dsub(w.dg, s[0]), s[1]);
        __syncwarp();
    } else {
        (void)rows;
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int c = 0; c < NMAX; c += 2) {
            if (c >= k && c != w.pos && c < n)
                a0 = dadd(a0, fabs(w.r[c]));
            if (c + 1 < NMAX && c + 1 >= k && c + 1 != w.pos && c + 1 < n)
                a1 = dadd(a1, fabs(w.r[c + 1]));
        }
        w.g = dsub(w.dg, dadd(a0, a1));
    }
}

// One phase-2 step (se90.rs:113-152, se99.rs:133-172).
template <int J, int NMAX, bool kStrict>
__device__ __forceinline__ void se_phase2_step(SeWarp<NMAX, kStrict>& w, int lane)
{
    const bool active = w.valid && w.pos >= J;
    double gmax;
    const int m = warp_argmax_first(w.g, active, w.pos, J, gmax);
    const int q = se_apply_pivot<J, NMAX, kStrict>(w, m);
    const double col = w.r[J];
    double piv = __shfl_sync(kFull, w.dg, q);
    const bool below = w.valid && w.pos > J;
    const double acol = below ? fabs(col) : 0.0;
    double normj;  // se90.rs:124, se99.rs:144
    if constexpr (kStrict) {
        w.buf0[w.pos] = acol;
        __syncwarp();
        const int len = w.n - J - 1;
        const double x = (lane < len) ? w.buf0[J + 1 + lane] : 0.0;
        normj = nd_sum_lanes(x, len);
    } else {
        normj = warp_sum_tree(acol);
    }
    const double ej = phase2_e(piv, normj, w.taugamma, w.delta_prev);  // se90.rs:125-131
    if (ej > 0.0) {
        piv = dadd(piv, ej);
        w.delta_prev = ej;
    }
    if (lane == q)
        w.e_mine = ej;
    // bound update (se90.rs:134-139, se99.rs:154-159)
    double sq, cs, ip;
    se_scale<kStrict>(piv, col, sq, cs, ip);
    if (fabs(dsub(piv, normj)) > MODCHOL_EPS) {
        double t;
        if (kStrict || !fast_range_ok(piv))
            t = dsub(1.0, ddiv(normj, piv));
        else
            t = __fma_rn(-normj, ip, 1.0);
        if (below)
            w.g = mul_add<kStrict>(w.g, acol, t);
    }
    se_eliminate<J, NMAX, kStrict>(w, lane, q, sq, cs, nmul_add<kStrict>(w.dg, cs, cs), below);
}

template <int kAlgo, int NMAX, bool kStrict, int kWarps>
__global__ void __launch_bounds__(kWarps * 32)
se_warp_kernel(int n, int vec_ok, long long batch, const double* __restrict__ A, double* __restrict__ L,
               double* __restrict__ E, unsigned long long* __restrict__ P, int* __restrict__ K)
{
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double* my = sm + (size_t)wib * se_warp_smem_doubles<NMAX, kStrict>();
    SeWarp<NMAX, kStrict> w;
    w.buf0 = my;
    w.buf1 = my + 32;
    double* rows = my + 64;
    w.n = n;
    w.valid = lane < n;
    const size_t nn = (size_t)n * n;
    const long long nwarps = (long long)gridDim.x * kWarps;

    for (long long b = (long long)blockIdx.x * kWarps + wib; b < batch; b += nwarps) {
        // ---- load: lane i takes column i of A (== row i, A symmetric), coalesced ----------
        const double* Ab = A + (size_t)b * nn;
#pragma unroll
        for (int c = 0; c < NMAX; ++c)
            w.r[c] = (c < n && w.valid) ? __ldg(Ab + (size_t)c * n + lane) : 0.0;
        w.pos = lane;
        w.e_mine = 0.0;
        w.g = 0.0;
        w.delta_prev = 0.0;
        w.dg = get_col<NMAX>(w.r, lane);  // per-lane diagonal (select chain, once per matrix)
        {
            bool win;  // gamma = max |a_ii| folded from 0.0 (se90.rs:55-57, se99.rs:54-56)
            w.gamma = warp_extreme<true>(fabs(w.dg), w.valid, 0.0, win);
            w.taugamma = dmul(MODCHOL_TAU, w.gamma);
        }

        // ---- phase 1 (se90.rs:61-96, se99.rs:61-109) --------------------------------------
        int kstart = -1;
        static_for<0, NMAX>([&](auto jc) {
            constexpr int J = decltype(jc)::value;
            if (kstart < 0 && J < n) {
                if (!se_phase1_step<kAlgo, J, NMAX, kStrict>(w, lane))
                    kstart = J;
            }
        });

        // ---- phase 2 ----------------------------------------------------------------------
        if (kstart >= 0) {
            if (kAlgo == kSE99 && kstart == n - 1) {  // se99.rs:115-119
                const int q = __ffs(__ballot_sync(kFull, w.valid && w.pos == n - 1)) - 1;
                const double ljj = __shfl_sync(kFull, w.dg, q);
                const double ej = tail1x1_e(ljj, w.gamma);
                if (lane == q)
                    w.e_mine = ej;
                set_col<NMAX>(w.r, n - 1, lane == q, dsqrt(dadd(ljj, ej)));
            } else {
                se_gershgorin_init<NMAX, kStrict>(w, lane, kstart, rows);
                static_for<0, (NMAX > 2 ? NMAX - 2 : 0)>([&](auto jc) {
                    constexpr int J = decltype(jc)::value;
                    if (J >= kstart && J + 2 < n)
                        se_phase2_step<J, NMAX, kStrict>(w, lane);
                });
                // final 2x2 block, never pivoted (se90.rs:155-166, se99.rs:175-186)
                const int qa = __ffs(__ballot_sync(kFull, w.valid && w.pos == n - 2)) - 1;
                const int qb = __ffs(__ballot_sync(kFull, w.valid && w.pos == n - 1)) - 1;
                double a00 = __shfl_sync(kFull, w.dg, qa);
                double a11 = __shfl_sync(kFull, w.dg, qb);
                const double a01 = __shfl_sync(kFull, get_col<NMAX>(w.r, n - 1), qa);
                double a10 = __shfl_sync(kFull, get_col<NMAX>(w.r, n - 2), qb);
                double lhi, llo;
                eigenvalues_2x2(a00, a01, a10, a11, lhi, llo);
                const double en = tail2x2_e(kAlgo, lhi, llo, w.gamma, w.delta_prev);
                if (lane == qa || lane == qb)
                    w.e_mine = en;
                if (en > 0.0) {
                    a00 = dadd(a00, en);
                    a11 = dadd(a11, en);
                }
                a00 = dsqrt(a00);
                a10 = ddiv(a10, a00);
                a11 = dsqrt(dsub(a11, dmul(a10, a10)));
                set_col<NMAX>(w.r, n - 2, lane == qa, a00);
                set_col<NMAX>(w.r, n - 2, lane == qb, a10);
                set_col<NMAX>(w.r, n - 1, lane == qb, a11);
            }
        }

        // ---- epilogue: row `pos` of L (strict upper explicitly zero), e, p ------------------
        double* Lrow = L + (size_t)b * nn + (size_t)w.pos * n;
        if (w.valid) {
            if (vec_ok) {  // n even and L 16-byte aligned
#pragma unroll
                for (int c = 0; c < NMAX; c += 2) {
                    if (c < n) {
                        double2 v;
                        v.x = (c <= w.pos) ? w.r[c] : 0.0;
                        v.y = (c + 1 <= w.pos) ? w.r[c + 1 < NMAX ? c + 1 : c] : 0.0;
                        *reinterpret_cast<double2*>(Lrow + c) = v;
                    }
                }
            } else {
#pragma unroll
                for (int c = 0; c < NMAX; ++c)
                    if (c < n)
                        Lrow[c] = (c <= w.pos) ? w.r[c] : 0.0;
            }
            E[(size_t)b * n + lane] = w.e_mine;
            P[(size_t)b * n + w.pos] = (unsigned long long)lane;
        }
        if (K && lane == 0)
            K[b] = kstart;
    }
}

}  // namespace modchol
