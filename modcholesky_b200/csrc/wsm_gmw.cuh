// wsm_gmw.cuh -- GMW81 (gmw81.rs:39-134) with the matrix resident in shared memory: one warp per
// matrix, packed lower triangle, lane = permuted position, R rows per lane (n <= 32 R).
// Same storage and swap as wsm_se.cuh.  STRICT: LEFT-looking like the reference so that every
// rounding is reproduced (FAST uses the symmetric form described at its loop below):
//   W(i,s), s <  j : c_is for the rows i >= j still to be eliminated (gmw81.rs:84); for a
//                    finished row i < j the FINAL L entry l_is * sqrt(d_s) (gmw81.rs:112-125)
//   W(i,s), s >= j : the permuted input (never modified by the reference either)
//   cd             : running diagonal of c (gmw81.rs:104-109), per row
//   dpos/sqd/rdpos : d_s, sqrt(d_s), 1/d_s for the position s == this lane's row (rows ARE
//                    positions here, so the same lane that owns row s owns column s's d)
#pragma once
#include "wsm_se.cuh"

namespace modchol {

// sum_{s<j} line[s] * W(i,s) for one row (byte address rowb), ndarray order or FMA chains
template <bool kStrict, typename AddrT>
__device__ __forceinline__ double wsm_gmw_dot(unsigned lineb, AddrT rowb, int j)
{
    if constexpr (kStrict) {
        double p[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
            p[u] = 0.0;
        int s = 0;
        for (; s + 8 <= j; s += 8) {
            double l[8], w[8];
#pragma unroll
            for (int u = 0; u < 8; u += 2) {
                const double2 v = lds_v2f64(lineb + 8u * (s + u));
                l[u] = v.x;
                l[u + 1] = v.y;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
                w[u] = lds_f64(rowb + 8u * (s + u));
#pragma unroll
            for (int u = 0; u < 8; ++u)
                p[u] = dadd(p[u], dmul(l[u], w[u]));
        }
        double acc = 0.0;
        acc = dadd(acc, dadd(p[0], p[4]));
        acc = dadd(acc, dadd(p[1], p[5]));
        acc = dadd(acc, dadd(p[2], p[6]));
        acc = dadd(acc, dadd(p[3], p[7]));
        for (; s < j; ++s)
            acc = dadd(acc, dmul(lds_f64(lineb + 8u * s), lds_f64(rowb + 8u * s)));
        return acc;
    } else {
        double a[4] = {0.0, 0.0, 0.0, 0.0};
        int s = 0;
        for (; s + 4 <= j; s += 4) {
            const double2 l01 = lds_v2f64(lineb + 8u * s), l23 = lds_v2f64(lineb + 8u * s + 16);
            double w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                w[u] = lds_f64(rowb + 8u * (s + u));
            a[0] = __fma_rn(l01.x, w[0], a[0]);
            a[1] = __fma_rn(l01.y, w[1], a[1]);
            a[2] = __fma_rn(l23.x, w[2], a[2]);
            a[3] = __fma_rn(l23.y, w[3], a[3]);
        }
        for (; s < j; ++s)
            a[0] = __fma_rn(lds_f64(lineb + 8u * s), lds_f64(rowb + 8u * s), a[0]);
        return dadd(dadd(a[0], a[1]), dadd(a[2], a[3]));
    }
}

template <int G, int R, bool kStrict, int kWarps, int kMinBlocks>
__global__ void __launch_bounds__(kWarps * 32, kMinBlocks)
wsm_gmw_kernel(int n, int wstride, long long batch, const double* __restrict__ A,
               double* __restrict__ L, double* __restrict__ E, unsigned long long* __restrict__ P,
               int* __restrict__ K, const double* Bm, double* X, int nrhs)   // X may alias Bm
{
    extern __shared__ double sm[];
    constexpr int kPack = 32 / G;   // matrices per warp, G lanes each (see wsm_se.cuh)
    int lane = threadIdx.x & 31;
    asm volatile("mov.u32 %0, %0;" : "+r"(lane));   // pinned in a register (see wsm_se.cuh)
    const int wib = threadIdx.x >> 5;
    const int sub = lane & (G - 1), grp = lane / G;
    const unsigned mask = wsm_group_mask<G>(grp);
    double* W = sm + (size_t)(wib * kPack + grp) * wstride;
    unsigned wb = (unsigned)__cvta_generic_to_shared(W);
    asm volatile("mov.u32 %0, %0;" : "+r"(wb));
    const unsigned scrb = wb + 8u * (unsigned)((tri(n) + 1) & ~1);
    const unsigned lineb = scrb + 128u;
    const size_t nn = (size_t)n * n;
    const long long nslots = (long long)gridDim.x * kWarps * kPack;

    int row[R];
    unsigned tib[R];
    bool rv[R];
#pragma unroll
    for (int s = 0; s < R; ++s) {
        row[s] = sub + G * s;
        rv[s] = row[s] < n;
        tib[s] = wb + 8u * (unsigned)tri(rv[s] ? row[s] : 0);
        asm volatile("mov.u32 %0, %0;" : "+r"(tib[s]));
    }

    for (long long b = ((long long)blockIdx.x * kWarps + wib) * kPack + grp; b < batch; b += nslots) {
        const double* Ab = A + (size_t)b * nn;
        // ---- load the lower triangle; off-diagonal maximum on the fly (gmw81.rs:52-58) --------
        __syncwarp(mask);
        const double om = wsm_load_lower<G, R, true>(Ab, n, wb, sub);
        __syncwarp(mask);

        double cd[R], ev[R], dpos[R], sqd[R], rdpos[R], av[R], omv[R], dummy[R];
        int perm[R];
        bool win[R];
#pragma unroll
        for (int s = 0; s < R; ++s) {
            cd[s] = rv[s] ? lds_f64(tib[s] + 8u * row[s]) : 0.0;   // gmw81.rs:66-67
            ev[s] = 0.0;
            dpos[s] = sqd[s] = rdpos[s] = 0.0;
            dummy[s] = 0.0;
            perm[s] = row[s];
            av[s] = fabs(cd[s]);
            omv[s] = om;
        }
        bool all[R];
#pragma unroll
        for (int s = 0; s < R; ++s)
            all[s] = s == 0;
        // gmw81.rs:49-64 (A symmetric: the lower off-diagonals carry the off-diagonal maximum)
        const double diag_max = wsm_extreme<true, R>(mask, av, rv, 0.0, win);
        const double off_max = wsm_extreme<true, R>(mask, omv, all, 0.0, win);
        const double nf = (double)n;
        const double delta = dmul(MODCHOL_EPS, fmax(1.0, dadd(diag_max, off_max)));
        const double beta =
            dsqrt(fmax(fmax(diag_max, ddiv(off_max, dsqrt(dsub(dmul(nf, nf), 1.0)))), MODCHOL_EPS));
        const double rbeta = kStrict ? 0.0 : ddiv(1.0, beta);

        unsigned tjb = wb;
        if constexpr (!kStrict) {
            // ---- FAST: the same factorisation in symmetric form --------------------------------
            // W(i,s), s < j holds l~_is = c_is / sqrt(d_s) -- the FINAL L entry -- instead of c_is:
            //   c_ij = a_ij - sum_s l~_is l~_js          (gmw81.rs:79-85 with l_js c_is = l~_js l~_is)
            //   c_ii  runs in cd:  cd_i -= l~_ij^2       (gmw81.rs:104-109)
            // so row j is read in place (no l_js line, no final scaling of the row), finished rows
            // skip the dot product and, for one matrix per warp and one row per lane, the blocked
            // DMMA update of wsm_se.cuh applies.  Same pivots, d_j, e_j up to rounding.
            constexpr bool kDmma = G == 32 && R == 1 && MODCHOL_WSM_DMMA;
            int sbase = 0;
            for (int j = 0; j < n; tjb += 8u * (unsigned)(j + 1), ++j) {
                if constexpr (kDmma) {
                    if (j - sbase == 4) {
                        wsm_rank4_update_dmma<4 * R>(wb, n, j, sbase, lane);
                        sbase = j;
                        __syncwarp(mask);
                    }
                }
                bool active[R], below[R];
#pragma unroll
                for (int s = 0; s < R; ++s) {
                    active[s] = rv[s] && row[s] >= j;
                    below[s] = rv[s] && row[s] > j;
                    av[s] = fabs(cd[s]);
                }
                double best;   // pivot on max |c_ii| (gmw81.rs:71-78)
                const int m = wsm_argmax_first<R>(mask, av, active, row, j, best);
                if (m != j)
                    wsm_swap_positions<G, R>(mask, wb, scrb, tib, row, rv, sub, j, m, tjb, cd, dummy, perm);
                // column j, its diagonal element included: c_jj comes from the same dot product as
                // in the reference (gmw81.rs:83-85 with i = j), not from the running diagonal,
                // whose different cancellation noise would show in E where E is tiny
                double col[R], acol[R];
                wsm_left_column<G, R, false, true>(j, tjb, sbase, tib, row, rv, col);
                const double cjj = wsm_bcast_row<G, R>(mask, col, j);
#pragma unroll
                for (int s = 0; s < R; ++s)
                    acol[s] = fabs(col[s]);
                // theta = max_{i>j} |c_ij| (gmw81.rs:87-100), d_j (gmw81.rs:102)
                const double theta = wsm_extreme<true, R>(mask, acol, below, 0.0, win);
                const double tb = dmul(theta, rbeta);
                const double dj = fmax(fmax(fabs(cjj), dmul(tb, tb)), delta);
                double sq, y;
                if (fast_range_ok(dj)) {
                    y = rsqrt_normal(dj);
                    sq = dmul(dj, y);
                } else {
                    sq = dsqrt(dj);
                    y = ddiv(1.0, sq);
                }
#pragma unroll
                for (int s = 0; s < R; ++s) {
                    const double ls = dmul(col[s], y);
                    sts_f64_if(below[s] || row[s] == j, below[s] ? tib[s] + 8u * j : tjb + 8u * j,
                               below[s] ? ls : sq);
                    if (below[s])
                        cd[s] = __fma_rn(-ls, ls, cd[s]);
                    if (row[s] == j)
                        ev[s] = dsub(dj, cjj);                           // gmw81.rs:110
                }
                __syncwarp(mask);
            }
        }
        for (int j = kStrict ? 0 : n; j < n; tjb += 8u * (unsigned)(j + 1), ++j) {
            // ---- pivot on max |c_ii| (gmw81.rs:71-78) ---------------------------------------
            bool active[R];
#pragma unroll
            for (int s = 0; s < R; ++s) {
                active[s] = rv[s] && row[s] >= j;
                av[s] = fabs(cd[s]);
            }
            double best;
            const int m = wsm_argmax_first<R>(mask, av, active, row, j, best);
            if (m != j)
                wsm_swap_positions<G, R>(mask, wb, scrb, tib, row, rv, sub, j, m, tjb, cd, dummy, perm);
            // ---- l_js = c_js / d_s, s < j (gmw81.rs:79-81): lane s owns d_s -----------------
            double ljs[R];
#pragma unroll
            for (int s = 0; s < R; ++s) {
                ljs[s] = 0.0;
                if (rv[s] && row[s] < j) {
                    const double cjs = lds_f64(tjb + 8u * row[s]);
                    ljs[s] = kStrict ? ddiv(cjs, dpos[s]) : dmul(cjs, rdpos[s]);
                    sts_f64(lineb + 8u * row[s], ljs[s]);
                }
            }
            __syncwarp(mask);
            // ---- c_ij = l_ij - sum_s l_js c_is for i >= j (gmw81.rs:83-85) ------------------
            double cij[R], acij[R];
            bool below[R];
#pragma unroll
            for (int s = 0; s < R; ++s) {
                below[s] = rv[s] && row[s] > j;
                cij[s] = 0.0;
                if (rv[s] && row[s] >= j) {
                    const double acc = wsm_gmw_dot<kStrict>(lineb, tib[s], j);
                    cij[s] = dsub(lds_f64(tib[s] + 8u * j), acc);
                }
                acij[s] = fabs(cij[s]);
            }
            // theta = max_{i>j} |c_ij|, 0 for the last column (gmw81.rs:87-100)
            const double theta = wsm_extreme<true, R>(mask, acij, below, 0.0, win);
            const double cjj = wsm_bcast_row<G, R>(mask, cij, j);
            const double tb = kStrict ? ddiv(theta, beta) : dmul(theta, rbeta);
            const double dj = fmax(fmax(fabs(cjj), dmul(tb, tb)), delta);   // gmw81.rs:102
            double sq, rdj = 0.0;
            const bool fast = !kStrict && fast_range_ok(dj);
            if (fast) {
                const double y = rsqrt_normal(dj);
                sq = dmul(dj, y);
                rdj = dmul(y, y);
            } else {
                sq = dsqrt(dj);
                if (!kStrict)
                    rdj = ddiv(1.0, dj);
            }
            __syncwarp(mask);   // row j's lane has read its c_js (dot product) before they are replaced
#pragma unroll
            for (int s = 0; s < R; ++s) {
                // row j is final from here on: L_js = l_js * sqrt(d_s); the reference's dense dot
                // adds exact zeros, which only turns a -0 into +0 (gmw81.rs:112-125)
                if (rv[s] && row[s] < j)
                    sts_f64(tjb + 8u * row[s], dadd(0.0, dmul(ljs[s], sqd[s])));
                if (below[s]) {
                    sts_f64(tib[s] + 8u * j, cij[s]);
                    const double c2 = dmul(cij[s], cij[s]);   // gmw81.rs:104-109
                    cd[s] = kStrict ? dsub(cd[s], ddiv(c2, dj)) : __fma_rn(-c2, rdj, cd[s]);
                }
                if (row[s] == j) {
                    ev[s] = dsub(dj, cjj);                                   // gmw81.rs:110
                    sts_f64(tjb + 8u * j, dadd(0.0, dmul(1.0, sq)));         // L_jj = sqrt(d_j)
                    dpos[s] = dj;
                    sqd[s] = sq;
                    rdpos[s] = rdj;
                }
            }
            __syncwarp(mask);
        }

        // ---- epilogue ---------------------------------------------------------------------
        __syncwarp(mask);
        if (L)
            wsm_store_lower<G, R>(L + (size_t)b * nn, n, wb, sub);
#pragma unroll
        for (int s = 0; s < R; ++s)
            if (rv[s]) {
                if (E)
                    E[(size_t)b * n + perm[s]] = ev[s];                      // gmw81.rs:127-131
                if (P)
                    P[(size_t)b * n + row[s]] = (unsigned long long)perm[s];
            }
        if (K && sub == 0)
            K[b] = -1;
        if (X)   // fused solve, see wsm_solve_in_place
            wsm_solve_in_place<G, R>(mask, n, wb, tib, row, rv, perm, Bm + (size_t)b * nrhs * n,
                                     X + (size_t)b * nrhs * n, nrhs);
    }
}

}  // namespace modchol
