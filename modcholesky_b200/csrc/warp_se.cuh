// warp_se.cuh -- SE90 / SE99 (se90.rs:38-181, se99.rs:35-201), one warp per matrix, the
// matrix resident in registers for the whole factorisation (n <= NMAX <= 32).  The ALTERNATE
// small-n family (modchol_set_small_kernel_family(1)); the default is wsm_se.cuh.
// See warp_common.cuh for the data layout.  HBM traffic per matrix is exactly the
// algorithmic 16 n^2 + 16 n bytes (+4 for phase2_start): A is read once with coalesced
// row-of-the-batch loads, L / e / p are written once.
//
// Code shape.  A first version unrolled all steps of both phases (static register indices
// everywhere): 650 KB of SASS, instruction-cache hit rate 51 %, "no instruction" the top
// stall (profiles/r01_notes.md).  This version runs ONE run-time loop over the step index j
// with one copy of all scalar logic; only the three operations that must name registers
// statically -- exchange columns j <-> m, read column j, write column j + trailing update --
// are dispatched through `switch (j)` into per-j specialisations.
#pragma once
#include "warp_common.cuh"

namespace modchol {

// Lanes per matrix: the smallest power of two >= NMAX.  32 / G matrices share a warp; every
// collective runs on the group's member mask, so the scalar logic is issued once for all of
// them.
template <int NMAX>
__host__ __device__ constexpr int se_group_lanes()
{
    return NMAX <= 8 ? 8 : (NMAX <= 16 ? 16 : 32);
}

// Per-warp shared memory (doubles): two 32-entry column lines (G entries per group), plus
// (STRICT only) a padded copy of each lane's row used once per matrix by the ndarray-ordered
// Gershgorin sums.
template <int NMAX, bool kStrict>
__host__ __device__ constexpr int se_warp_smem_doubles()
{
    return 64 + (kStrict ? 32 * (NMAX + 1) : 0);
}

#define MODCHOL_ALLJ(X)                                                                        \
    X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15)       \
    X(16) X(17) X(18) X(19) X(20) X(21) X(22) X(23) X(24) X(25) X(26) X(27) X(28) X(29) X(30)  \
    X(31)

// r[k] -= cs * line[k] for k = J+1 .. n-1 (line = scaled pivot column in position order);
// columns go in groups of eight so that padding groups (k >= n) are skipped.
template <int J, int NMAX, bool kStrict>
__device__ __forceinline__ void trailing_update(double (&r)[NMAX], double cs, const double* line,
                                                int n)
{
    constexpr int K0 = J + 1;
    if constexpr ((K0 & 1) && K0 < NMAX)
        r[K0] = nmul_add<kStrict>(r[K0], cs, line[K0]);
    constexpr int K1 = (K0 + 1) & ~1;
    static_for<0, (NMAX - K1 + 7) / 8>([&](auto gi) {
        constexpr int base = K1 + 8 * decltype(gi)::value;
        if (base < n) {
#pragma unroll
            for (int k = base; k < base + 8 && k < NMAX; k += 2) {
                const double2 v = *reinterpret_cast<const double2*>(&line[k]);
                r[k] = nmul_add<kStrict>(r[k], cs, v.x);
                if (k + 1 < NMAX)
                    r[k + 1] = nmul_add<kStrict>(r[k + 1], cs, v.y);
            }
        }
    });
}

// dispatch 2: r[j] = newj (sqrt of the pivot in the pivot lane, scaled column elsewhere),
// then the trailing update of se90.rs:88-92 / se99.rs:100-104.
// Returns the updated column j+1 (the next step's `cj`).
template <int NMAX, bool kStrict>
__device__ __forceinline__ double dispatch_write_update(double (&r)[NMAX], int j, double newj,
                                                        double cs, const double* line, int n)
{
    double next = 0.0;
#define MODCHOL_X(J)                                             \
    case J:                                                      \
        if constexpr (J < NMAX) {                                \
            r[J] = newj;                                         \
            trailing_update<J, NMAX, kStrict>(r, cs, line, n);   \
            if constexpr (J + 1 < NMAX)                          \
                next = r[J + 1];                                 \
        }                                                        \
        break;
    switch (j) {
        MODCHOL_ALLJ(MODCHOL_X)
    default:
        break;
    }
#undef MODCHOL_X
    return next;
}

// Lower Gershgorin bounds of the trailing block starting at k (se90.rs:105-110,
// se99.rs:125-130): g_i = l_ii - S(|l_i,k..i-1|) - S(|l_i+1..n-1,i|), S = ndarray sum.
template <int NMAX, bool kStrict>
__device__ __forceinline__ double gershgorin_init(unsigned mask, const double (&r)[NMAX], double dg,
                                                  int pos, bool valid, int lane, int k, int n,
                                                  double* rows)
{
    if constexpr (kStrict) {
        // every lane parks its row in shared memory (stride NMAX+1: conflict-free), then walks
        // its two index ranges with run-time offsets and compile-time partial-sum slots.
        double* row = rows + lane * (NMAX + 1);
#pragma unroll
        for (int c = 0; c < NMAX; ++c)
            row[c] = r[c];
        __syncwarp(mask);
        double s[2];
#pragma unroll
        for (int part = 0; part < 2; ++part) {
            const int start = part == 0 ? k : pos + 1;
            int len = part == 0 ? pos - k : n - 1 - pos;
            if (!(valid && pos >= k))
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
        __syncwarp(mask);
        return dsub(dsub(dg, s[0]), s[1]);
    } else {
        (void)rows;
        (void)lane;
        (void)valid;
        (void)n;
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int c = 0; c < NMAX; c += 2) {
            // padding columns (c >= n) hold zeros
            if (c >= k && c != pos)
                a0 = dadd(a0, fabs(r[c]));
            if (c + 1 < NMAX && c + 1 >= k && c + 1 != pos)
                a1 = dadd(a1, fabs(r[c + 1]));
        }
        return dsub(dg, dadd(a0, a1));
    }
}

template <int kAlgo, int NMAX, bool kStrict, int kWarps, int kMinBlocks>
__global__ void __launch_bounds__(kWarps * 32, kMinBlocks)
se_warp_kernel(int n, int vec_ok, long long batch, const double* __restrict__ A,
               double* __restrict__ L, double* __restrict__ E, unsigned long long* __restrict__ P,
               int* __restrict__ K)
{
    extern __shared__ double sm[];
    constexpr int G = se_group_lanes<NMAX>();   // lanes per matrix
    constexpr int M = 32 / G;                   // matrices per warp
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int grp = lane / G, gl = lane % G;    // group within the warp, lane within the group
    const unsigned gmask = (G == 32) ? kFull : (((1u << G) - 1u) << (grp * G));
    double* my = sm + (size_t)wib * se_warp_smem_doubles<NMAX, kStrict>();
    double* buf0 = my + grp * G;        // column line, unscaled |values| (STRICT norms)
    double* buf1 = my + 32 + grp * G;   // column line, scaled values (trailing update)
    double* rows = my + 64;
    const size_t nn = (size_t)n * n;
    const long long nwarps = (long long)gridDim.x * kWarps;
    const long long nslots = (batch + M - 1) / M;   // warp-sized work items

    for (long long slot = (long long)blockIdx.x * kWarps + wib; slot < nslots; slot += nwarps) {
        const long long b = slot * M + grp;
        const bool valid = gl < n && b < batch;
        // ---- load: lane i takes column i of A (== row i, A symmetric), coalesced ----------
        double r[NMAX];
        const double* Ab = A + (size_t)(b < batch ? b : 0) * nn + gl;
#pragma unroll
        for (int c = 0; c < NMAX; ++c)
            r[c] = (c < n && valid) ? __ldg(Ab + c * n) : 0.0;
        const bool live = b < batch;
        int pos = gl;            // permuted position of this lane's row
        double e_mine = 0.0;     // E entry of this lane's ORIGINAL index (the un-permute of
                                 // se99.rs:194-198 is free in this layout)
        double g = 0.0;          // lower Gershgorin bound (phase 2)
        double delta_prev = 0.0;
        // current diagonal element of this lane's row (r[pos] would need a dynamic index)
        double dg = valid ? __ldg(Ab + gl * n) : 0.0;
        bool win;
        // gamma = max |a_ii| folded from 0.0 (se90.rs:55-57, se99.rs:54-56)
        const double gamma = warp_extreme<true>(gmask, fabs(dg), valid, 0.0, win);
        const double taugamma = dmul(MODCHOL_TAU, gamma);

        double cj = r[0];        // register copy of the current column j (see take_column)
        int kstart = -1;

        // B: make position m the new position j (se90.rs:64-68,116-121); returns this lane's
        // element of the pivot column
        auto bring_pivot = [&](int j, int m) -> double {
            if (m != j) {
                cj = take_column<NMAX>(r, m, cj);   // cj = r[m]; r[m] = old column j
                if (pos == j)
                    pos = m;
                else if (pos == m)
                    pos = j;
            }
            return cj;
        };
        // D: s = sqrt(pivot) and the scaled column element (se90.rs:84-87, se99.rs:96-99).
        // FAST: y = rsqrt(piv), s = piv*y, cs = col*y, 1/piv = y*y for pivots in the safe
        // range; everything else takes the reference's formulas.
        auto scale = [&](double piv, double col, double& s, double& cs, double& inv_piv) -> bool {
            const bool fast = !kStrict && fast_range_ok(piv);
            if (fast) {
                const double y = rsqrt_normal(piv);
                s = dmul(piv, y);
                cs = dmul(col, y);
                inv_piv = dmul(y, y);
            } else {
                s = dsqrt(piv);
                cs = ddiv(col, s);
                inv_piv = 0.0;
            }
            return fast;
        };
        // ---- one loop over the step j; the scalar part is phase specific and straight-line, the
        //      elimination step (the big switch dispatch) exists once for both phases ----------
        bool phase1 = true;
        int j = 0;
        const int nsteps = live ? n : 0;   // an idle group (batch tail) does nothing
        while (true) {
            int q;
            bool below;
            double s, cs, dg_new;
            if (phase1) {
                // ---- phase 1 (se90.rs:61-96, se99.rs:61-109) ------------------------------
                if (j >= nsteps)
                    break;
                const bool active = valid && pos >= j;
                double dmax;  // se90.rs:63, se99.rs:76
                const int m = warp_argmax_first(gmask, dg, active, pos, j, dmax);
                bool to_phase2 = false;
                if (kAlgo == kSE99) {  // se99.rs:62-73, tested before the pivot is applied
                    const double dmin = warp_extreme<false>(gmask, dg, active, INFINITY, win);
                    to_phase2 = dmax < taugamma || dmin < dmul(-MODCHOL_MU, dmax);
                }
                if (!to_phase2) {
                    const double col = bring_pivot(j, m);
                    q = __ffs(__ballot_sync(gmask, pos == j) >> (grp * G)) - 1;
                    const double piv = __shfl_sync(gmask, dg, q, G);
                    below = valid && pos > j;
                    double inv_piv;
                    const bool fast = scale(piv, col, s, cs, inv_piv);
                    dg_new = nmul_add<kStrict>(dg, cs, cs);
                    // look-ahead min_i (l_ii - l_ij^2 / l_jj) (se90.rs:70-77, se99.rs:82-89);
                    // in FAST mode that is the updated diagonal itself.
                    const double nv = fast ? dg_new : dsub(dg, ddiv(dmul(col, col), piv));
                    const double la = warp_extreme<false>(gmask, nv, below, INFINITY, win);
                    const double thresh = (kAlgo == kSE99) ? dmul(-MODCHOL_MU, gamma) : taugamma;
                    if (la < thresh) {  // the pivot at j stays applied (se99.rs:90-93)
                        to_phase2 = true;
                        put_column<NMAX>(r, j, cj);   // materialise column j (see take_column)
                    }
                }
                if (to_phase2) {
                    phase1 = false;
                    kstart = j;
                    if (!(kAlgo == kSE99 && kstart == n - 1))
                        g = gershgorin_init<NMAX, kStrict>(gmask, r, dg, pos, valid, lane, kstart, n, rows);
                    continue;
                }
            } else {
                // ---- phase 2 step (se90.rs:113-152, se99.rs:133-172) -----------------------
                if (j + 2 >= n)
                    break;
                const bool active = valid && pos >= j;
                double gmax;  // se90.rs:115, se99.rs:135
                const int m = warp_argmax_first(gmask, g, active, pos, j, gmax);
                const double col = bring_pivot(j, m);
                q = __ffs(__ballot_sync(gmask, pos == j) >> (grp * G)) - 1;
                double piv = __shfl_sync(gmask, dg, q, G);
                below = valid && pos > j;
                // E_jj (se90.rs:124-131, se99.rs:144-151)
                const double acol = below ? abs_bits(col) : 0.0;
                double normj;
                if constexpr (kStrict) {
                    buf0[pos] = acol;
                    __syncwarp(gmask);
                    const int len = n - j - 1;
                    const double x = (gl < len) ? buf0[j + 1 + gl] : 0.0;
                    normj = nd_sum_lanes<G>(gmask, x, len);
                } else {
                    normj = warp_sum_tree<G>(gmask, acol);
                }
                // e_j = max(max(0, delta_prev), -l_jj + max(normj, tau*gamma)).  delta_prev is
                // always >= +0 and never NaN, so max(0, delta_prev) == delta_prev, and the two
                // remaining maxima have a never-NaN second operand.
                const double ej =
                    max_b_not_nan(dadd(-piv, max_b_not_nan(normj, taugamma)), delta_prev);
                if (kStrict) {
                    if (ej > 0.0) {
                        piv = dadd(piv, ej);
                        delta_prev = ej;
                    }
                } else {  // ej >= delta_prev >= 0: adding an exact zero changes nothing
                    piv = dadd(piv, ej);
                    delta_prev = ej;
                }
                if (gl == q)
                    e_mine = ej;
                double inv_piv;
                const bool fast = scale(piv, col, s, cs, inv_piv);
                dg_new = nmul_add<kStrict>(dg, cs, cs);
                // bound update (se90.rs:134-139, se99.rs:154-159)
                if (fabs(dsub(piv, normj)) > MODCHOL_EPS) {
                    const double t =
                        fast ? __fma_rn(-normj, inv_piv, 1.0) : dsub(1.0, ddiv(normj, piv));
                    if (below)
                        g = mul_add<kStrict>(g, acol, t);
                }
            }
            // ---- elimination step (se90.rs:84-93 == :142-151, se99.rs:96-105 == :162-171) ----
            buf1[pos] = cs;
            __syncwarp(gmask);
            if (below)
                dg = dg_new;
            cj = dispatch_write_update<NMAX, kStrict>(r, j, gl == q ? s : cs, cs, buf1, n);
            __syncwarp(gmask);
            ++j;
        }

        // ---- tails ------------------------------------------------------------------------
        if (kstart >= 0) {
            if (kAlgo == kSE99 && kstart == n - 1) {  // se99.rs:115-119
                const int q = __ffs(__ballot_sync(gmask, valid && pos == n - 1) >> (grp * G)) - 1;
                const double ljj = __shfl_sync(gmask, dg, q, G);
                const double ej = tail1x1_e(ljj, gamma);
                if (gl == q)
                    e_mine = ej;
                put_column_if<NMAX>(r, n - 1, gl == q, dsqrt(dadd(ljj, ej)));
            } else {
                // final 2x2 block, never pivoted (se90.rs:155-166, se99.rs:175-186)
                const int qa = __ffs(__ballot_sync(gmask, valid && pos == n - 2) >> (grp * G)) - 1;
                const int qb = __ffs(__ballot_sync(gmask, valid && pos == n - 1) >> (grp * G)) - 1;
                double a00 = __shfl_sync(gmask, dg, qa, G);
                double a11 = __shfl_sync(gmask, dg, qb, G);
                const double a01 = __shfl_sync(gmask, read_column<NMAX>(r, n - 1), qa, G);
                double a10 = __shfl_sync(gmask, read_column<NMAX>(r, n - 2), qb, G);
                double lhi, llo;
                eigenvalues_2x2(a00, a01, a10, a11, lhi, llo);
                const double en = tail2x2_e(kAlgo, lhi, llo, gamma, delta_prev);
                if (gl == qa || gl == qb)
                    e_mine = en;
                if (en > 0.0) {
                    a00 = dadd(a00, en);
                    a11 = dadd(a11, en);
                }
                a00 = dsqrt(a00);
                a10 = ddiv(a10, a00);
                a11 = dsqrt(dsub(a11, dmul(a10, a10)));
                put_column_if<NMAX>(r, n - 2, gl == qa || gl == qb, gl == qa ? a00 : a10);
                put_column_if<NMAX>(r, n - 1, gl == qb, a11);
            }
        }

        // ---- epilogue: row `pos` of L (strict upper explicitly zero), e, p ------------------
        double* Lrow = L + (size_t)(b < batch ? b : 0) * nn + (size_t)pos * n;
        if (valid) {
            if (vec_ok) {  // n even and L 16-byte aligned
#pragma unroll
                for (int c = 0; c < NMAX; c += 2) {
                    if (c < n) {
                        double2 v;
                        v.x = (c <= pos) ? r[c] : 0.0;
                        v.y = (c + 1 <= pos) ? r[c + 1 < NMAX ? c + 1 : c] : 0.0;
                        *reinterpret_cast<double2*>(Lrow + c) = v;
                    }
                }
            } else {
#pragma unroll
                for (int c = 0; c < NMAX; ++c)
                    if (c < n)
                        Lrow[c] = (c <= pos) ? r[c] : 0.0;
            }
            E[(size_t)b * n + gl] = e_mine;
            P[(size_t)b * n + pos] = (unsigned long long)gl;
        }
        if (K && gl == 0 && live)
            K[b] = kstart;
    }
}

}  // namespace modchol
