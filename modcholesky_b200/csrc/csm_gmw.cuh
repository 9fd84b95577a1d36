// csm_gmw.cuh -- GMW81 (gmw81.rs:39-134), one thread block per matrix, packed lower triangle in
// shared memory, thread = permuted position.  Block-level sibling of wsm_gmw.cuh (same storage
// conventions: W(i,s) = c_is for s < j and i >= j, final L entries for finished rows, untouched
// input for s >= j; LEFT-looking dot products in ndarray order; FAST: the symmetric form with
// W(i,s) = c_is / sqrt(d_s)) with the two-level collectives of csm_se.cuh.  64 < n <= 256
// (kSpill above 236).
#pragma once
#include "csm_se.cuh"
#include "wsm_gmw.cuh"

namespace modchol {

template <bool kStrict, int kThreads, int kMinBlocks, bool kSpill = false, bool kDmma = false>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
csm_gmw_kernel(int n, long long batch, const double* __restrict__ A, double* __restrict__ L,
               double* __restrict__ E, unsigned long long* __restrict__ P, int* __restrict__ K,
               double* __restrict__ spill = nullptr, int S = 0)
{
    extern __shared__ double sm[];
    double* W = sm;
    using addr_t = typename CsmRows<kSpill>::addr_t;
    CsmRows<kSpill> rows;   // see csm_se.cuh: rows [0, S) may live in a global spill buffer
    rows.S = kSpill ? S : 0;
    rows.triS = tri(rows.S);
    const int tpad = (tri(n) - rows.triS + 1) & ~1;
    double* linep = W + tpad;
    const int lpad = (n + 8 + 1) & ~1;
    CsmScratch sc = csm_scratch(linep + lpad);
    const unsigned wsh = (unsigned)__cvta_generic_to_shared(W);
    if constexpr (kSpill) {
        rows.sbase = (unsigned long long)(size_t)W;
        rows.gb = (unsigned long long)(size_t)(spill + (size_t)blockIdx.x * rows.triS);
    } else {
        rows.sbase = wsh;
        rows.gb = 0;
    }
    const unsigned lineb = wsh + 8u * (unsigned)tpad;
    const unsigned panelb = wsh + 8u * (unsigned)(tpad + lpad + csm_scratch_doubles());   // kDmma
    const int tid = threadIdx.x;
    const int nwarps = kThreads >> 5;
    const size_t nn = (size_t)n * n;
    const int row = tid;
    const bool rv = row < n;
    const addr_t tib = rows.row(rv ? row : 0);

    for (long long b = blockIdx.x; b < batch; b += gridDim.x) {
        const double* Ab = A + (size_t)b * nn;
        // ---- load the lower triangle; off-diagonal maximum on the fly (gmw81.rs:52-58) --------
        __syncthreads();
        double om = 0.0;
        for (int i0 = 0; i0 < n; i0 += 4) {
            double v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u;
                v[u] = (i < n && tid <= i) ? __ldg(Ab + i * n + tid) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u;
                sts_f64_if(i < n && tid <= i, rows.row(i < n ? i : 0) + 8u * (unsigned)tid, v[u]);
                if (i < n && tid < i)
                    om = fmax(om, fabs(v[u]));
            }
        }
        __syncthreads();

        double cd = rv ? lds_f64(tib + 8u * row) : 0.0;   // gmw81.rs:66-67
        double ev = 0.0, dpos = 0.0, sqd = 0.0, rdpos = 0.0;
        int perm = row;
        bool win;
        // gmw81.rs:49-64
        const double diag_max = csm_extreme<true>(sc, nwarps, fabs(cd), rv, 0.0, win);
        const double off_max = csm_extreme<true>(sc, nwarps, om, true, 0.0, win);
        const double nf = (double)n;
        const double delta = dmul(MODCHOL_EPS, fmax(1.0, dadd(diag_max, off_max)));
        const double beta =
            dsqrt(fmax(fmax(diag_max, ddiv(off_max, dsqrt(dsub(dmul(nf, nf), 1.0)))), MODCHOL_EPS));
        const double rbeta = kStrict ? 0.0 : ddiv(1.0, beta);

        // symmetric exchange of positions j < m (see wsm_swap_positions) + row scalars
        auto swap_positions = [&](int j, int m, addr_t tjb) {
            const addr_t tmb = rows.row(m);
            const int i = row;
            const bool act = rv && i != m;
            addr_t a1 = (i <= j) ? tjb + 8u * i : tib + 8u * j;
            addr_t a2 = (i == j) ? tmb + 8u * m : ((i < m) ? tmb + 8u * i : tib + 8u * m);
            if (!act) {
                a1 = tib;
                a2 = tib;
            }
            const double t1 = lds_f64(a1), t2 = lds_f64(a2);
            if (i == j || i == m) {
                double* q = sc.scal + (i == j ? 0 : 4);
                q[0] = cd;
                q[1] = (double)perm;
            }
            __syncthreads();
            sts_f64_if(act, a1, t2);
            sts_f64_if(act, a2, t1);
            if (i == j || i == m) {
                const double* q = sc.scal + (i == j ? 4 : 0);
                cd = q[0];
                perm = (int)q[1];
            }
            __syncthreads();
        };
        addr_t tjb = rows.row(0);
        if constexpr (!kStrict) {
            // ---- FAST: symmetric form (see wsm_gmw.cuh): W(i,s) = c_is / sqrt(d_s), the final L
            // entry; row j read in place, finished rows skip the dot product, blocked DMMA update
            int sbase = 0;
            for (int j = 0; j < n; tjb = rows.next(tjb, j), ++j) {
                if constexpr (kDmma) {
                    if (j - sbase == 4) {
                        csm_blocked_update<kSpill>(rows, n, j, sbase, panelb, tib, row, rv, tid, nwarps);
                        sbase = j;
                    }
                }
                const bool active = rv && row >= j, below = rv && row > j;
                double best;   // pivot on max |c_ii| (gmw81.rs:71-78)
                const int m = csm_argmax_first(sc, nwarps, fabs(cd), active, row, j, best);
                if (m != j)
                    swap_positions(j, m, tjb);
                // column j with its diagonal element (c_jj from the same dot product, gmw81.rs:83-85)
                const double col = csm_left_column<false, true, addr_t>(j, tjb, sbase, tib, row, rv);
                double* D = sc.dbl + sc.buf * 8;
                if (row == j)
                    D[0] = col;
                // theta = max_{i>j} |c_ij| (gmw81.rs:87-100); c_jj rides on the same barrier
                const double theta = csm_extreme<true>(sc, nwarps, fabs(col), below, 0.0, win);
                const double cjj = D[0];
                const double tb = dmul(theta, rbeta);
                const double dj = fmax(fmax(fabs(cjj), dmul(tb, tb)), delta);   // gmw81.rs:102
                double sq, y;
                if (fast_range_ok(dj)) {
                    y = rsqrt_normal(dj);
                    sq = dmul(dj, y);
                } else {
                    sq = dsqrt(dj);
                    y = ddiv(1.0, sq);
                }
                const double ls = dmul(col, y);
                sts_f64_if(below || row == j, below ? tib + 8u * j : tjb + 8u * j, below ? ls : sq);
                if (below)
                    cd = __fma_rn(-ls, ls, cd);
                if (row == j)
                    ev = dsub(dj, cjj);                                  // gmw81.rs:110
                __syncthreads();
            }
        }
        for (int j = kStrict ? 0 : n; j < n; tjb = rows.next(tjb, j), ++j) {
            // ---- pivot on max |c_ii| (gmw81.rs:71-78) ---------------------------------------
            const bool active = rv && row >= j;
            double best;
            const int m = csm_argmax_first(sc, nwarps, fabs(cd), active, row, j, best);
            if (m != j)
                swap_positions(j, m, tjb);
            // ---- l_js = c_js / d_s, s < j (gmw81.rs:79-81): thread s owns d_s ---------------
            double ljs = 0.0;
            if (rv && row < j) {
                const double cjs = lds_f64(tjb + 8u * row);
                ljs = kStrict ? ddiv(cjs, dpos) : dmul(cjs, rdpos);
                sts_f64(lineb + 8u * row, ljs);
            }
            __syncthreads();
            // ---- c_ij = l_ij - sum_s l_js c_is for i >= j (gmw81.rs:83-85) ------------------
            const bool below = rv && row > j;
            double cij = 0.0;
            if (rv && row >= j)
                cij = dsub(lds_f64(tib + 8u * j), wsm_gmw_dot<kStrict>(lineb, tib, j));
            // theta = max_{i>j} |c_ij| (gmw81.rs:87-100); the sync inside also orders the dot
            // product's reads of row j before the writes below
            const double theta = csm_extreme<true>(sc, nwarps, fabs(cij), below, 0.0, win);
            double* D = sc.dbl + sc.buf * 8;
            if (row == j)
                D[0] = cij;
            __syncthreads();
            sc.buf ^= 1;
            const double cjj = D[0];
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
            // row j is final: L_js = l_js * sqrt(d_s) (+0: the dense dot of gmw81.rs:125)
            if (rv && row < j)
                sts_f64(tjb + 8u * row, dadd(0.0, dmul(ljs, sqd)));
            if (below) {
                sts_f64(tib + 8u * j, cij);
                const double c2 = dmul(cij, cij);   // gmw81.rs:104-109
                cd = kStrict ? dsub(cd, ddiv(c2, dj)) : __fma_rn(-c2, rdj, cd);
            }
            if (row == j) {
                ev = dsub(dj, cjj);                                      // gmw81.rs:110
                sts_f64(tjb + 8u * j, dadd(0.0, dmul(1.0, sq)));
                dpos = dj;
                sqd = sq;
                rdpos = rdj;
            }
            __syncthreads();
        }

        // ---- epilogue ---------------------------------------------------------------------
        double* Lb = L + (size_t)b * nn;
        for (int i0 = 0; i0 < n; i0 += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u;
                if (i < n && tid < n)
                    Lb[i * n + tid] = (tid <= i) ? lds_f64(rows.row(i) + 8u * (unsigned)tid) : 0.0;
            }
        }
        if (rv) {
            E[(size_t)b * n + perm] = ev;                                // gmw81.rs:127-131
            P[(size_t)b * n + row] = (unsigned long long)perm;
        }
        if (K && tid == 0)
            K[b] = -1;
    }
}

}  // namespace modchol
