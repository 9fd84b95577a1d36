// warp_gmw.cuh -- GMW81 (gmw81.rs:39-134), G lanes per matrix, matrix resident in registers.
//
// Same layout as the SE kernels (warp_common.cuh): lane i owns original row i; r[c] is the
// element (row of this lane, permuted column position c).  GMW81 is LEFT-looking in the
// reference, and so is this kernel, so that STRICT mode reproduces every rounding:
//   for c <  j : r[c] = c_{row,c}            (gmw81.rs:84, computed when column c was the pivot)
//   for c >= j : r[c] = the permuted input   (the reference never modifies `l` there)
//   cd          = running diagonal of c      (gmw81.rs:104-109), per lane
//   dpos / sqd  = d_s and sqrt(d_s) for the position s == lane-within-group
// Row j of the output L (= l_{j,s} * sqrt(d_s), gmw81.rs:112-125) is final the moment row j
// becomes the pivot row, and is written then, one coalesced row per step -- the dense
// `l.dot(&dout)` of the reference (2 n^3 flops) is never formed.
#pragma once
#include "warp_se.cuh"

namespace modchol {

template <int NMAX>
__host__ __device__ constexpr int gmw_warp_smem_doubles()
{
    return 64;  // two 32-entry lines per warp (G entries per group)
}

// sum_{s<J} line[s] * r[s] in ndarray's unrolled_fold order (STRICT) or four FMA chains.
template <int J, int NMAX, bool kStrict>
__device__ __forceinline__ double gmw_dot(const double (&r)[NMAX], const double* line)
{
    if constexpr (J == 0) {
        return 0.0;
    } else if constexpr (kStrict) {
        constexpr int C = J & ~7;
        double p[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
            p[u] = 0.0;
#pragma unroll
        for (int s = 0; s < C; s += 2) {
            const double2 v = *reinterpret_cast<const double2*>(&line[s]);
            p[s & 7] = dadd(p[s & 7], dmul(v.x, r[s]));
            p[(s + 1) & 7] = dadd(p[(s + 1) & 7], dmul(v.y, r[s + 1]));
        }
        double acc = 0.0;
        acc = dadd(acc, dadd(p[0], p[4]));
        acc = dadd(acc, dadd(p[1], p[5]));
        acc = dadd(acc, dadd(p[2], p[6]));
        acc = dadd(acc, dadd(p[3], p[7]));
#pragma unroll
        for (int s = C; s < J; ++s)
            acc = dadd(acc, dmul(line[s], r[s]));
        return acc;
    } else {
        double a[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int s = 0; s + 1 < J; s += 2) {
            const double2 v = *reinterpret_cast<const double2*>(&line[s]);
            a[s & 3] = __fma_rn(v.x, r[s], a[s & 3]);
            a[(s + 1) & 3] = __fma_rn(v.y, r[s + 1], a[(s + 1) & 3]);
        }
        if constexpr (J & 1)
            a[(J - 1) & 3] = __fma_rn(line[J - 1], r[J - 1], a[(J - 1) & 3]);
        return dadd(dadd(a[0], a[1]), dadd(a[2], a[3]));
    }
}

// dispatch: r[j] = c_ij = l_ij - dot_j  (gmw81.rs:83-85), l_ij = cj (the register copy of
// column j); returns c_ij and, through `next`, column j+1 (the next step's cj).
template <int NMAX, bool kStrict>
__device__ __forceinline__ double dispatch_gmw_column(double (&r)[NMAX], int j, double cj,
                                                      const double* line, double& next)
{
    double cij = 0.0;
#define MODCHOL_X(J)                                                   \
    case J:                                                            \
        if constexpr (J < NMAX) {                                      \
            cij = dsub(cj, gmw_dot<J, NMAX, kStrict>(r, line));        \
            r[J] = cij;                                                \
            if constexpr (J + 1 < NMAX)                                \
                next = r[J + 1];                                       \
        }                                                              \
        break;
    switch (j) {
        MODCHOL_ALLJ(MODCHOL_X)
    default:
        break;
    }
#undef MODCHOL_X
    return cij;
}

template <int NMAX, bool kStrict, int kWarps, int kMinBlocks>
__global__ void __launch_bounds__(kWarps * 32, kMinBlocks)
gmw_warp_kernel(int n, long long batch, const double* __restrict__ A, double* __restrict__ L,
                double* __restrict__ E, unsigned long long* __restrict__ P, int* __restrict__ K)
{
    extern __shared__ double sm[];
    constexpr int G = se_group_lanes<NMAX>();
    constexpr int M = 32 / G;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int grp = lane / G, gl = lane % G;
    const unsigned gmask = (G == 32) ? kFull : (((1u << G) - 1u) << (grp * G));
    double* my = sm + (size_t)wib * gmw_warp_smem_doubles<NMAX>();
    double* rowline = my + grp * G;       // pivot row c_{j,s}, by position s
    double* lline = my + 32 + grp * G;    // l_{j,s} = c_{j,s} / d_s, by position s
    const size_t nn = (size_t)n * n;
    const long long nwarps = (long long)gridDim.x * kWarps;
    const long long nslots = (batch + M - 1) / M;

    for (long long slot = (long long)blockIdx.x * kWarps + wib; slot < nslots; slot += nwarps) {
        const long long b = slot * M + grp;
        const bool live = b < batch;
        const bool valid = gl < n && live;
        double r[NMAX];
        const double* Ab = A + (size_t)(live ? b : 0) * nn + gl;
#pragma unroll
        for (int c = 0; c < NMAX; ++c)
            r[c] = (c < n && valid) ? __ldg(Ab + c * n) : 0.0;
        int pos = gl;
        double e_mine = 0.0;
        double cd = get_col<NMAX>(r, gl);   // gmw81.rs:66-67: diag(c) = diag(A)
        double dpos = 0.0, sqd = 0.0, rdpos = 0.0;
        double cj = r[0];   // register copy of the current column j (see take_column)

        // gmw81.rs:49-64
        bool win;
        const double diag_max = warp_extreme<true>(gmask, fabs(cd), valid, 0.0, win);
        double om = 0.0;
#pragma unroll
        for (int c = 0; c < NMAX; ++c)
            if (c != gl && c < n)
                om = fmax(om, fabs(r[c]));
        const double off_max = warp_extreme<true>(gmask, om, valid, 0.0, win);
        const double nf = (double)n;
        const double delta = dmul(MODCHOL_EPS, fmax(1.0, dadd(diag_max, off_max)));
        const double beta =
            dsqrt(fmax(fmax(diag_max, ddiv(off_max, dsqrt(dsub(dmul(nf, nf), 1.0)))), MODCHOL_EPS));
        const double rbeta = kStrict ? 0.0 : ddiv(1.0, beta);

        double* Lb = L + (size_t)(live ? b : 0) * nn;
        const int nsteps = live ? n : 0;
        for (int j = 0; j < nsteps; ++j) {
            // ---- pivot on max |c_ii| (gmw81.rs:71-78) ---------------------------------------
            const bool active = valid && pos >= j;
            double best;
            const int m = warp_argmax_first(gmask, fabs(cd), active, pos, j, best);
            if (m != j) {
                cj = take_column<NMAX>(r, m, cj);   // cj = r[m]; r[m] = old column j
                if (pos == j)
                    pos = m;
                else if (pos == m)
                    pos = j;
            }
            const int q = __ffs(__ballot_sync(gmask, pos == j) >> (grp * G)) - 1;
            const bool below = valid && pos > j;
            // ---- l_js = c_js / d_s, s < j (gmw81.rs:79-81): the pivot lane parks its row, the
            //      position lanes divide, everybody reads the quotients back ------------------
            if (gl == q) {
                static_for<0, NMAX / 8>([&](auto gi) {
                    constexpr int base = 8 * decltype(gi)::value;
                    if (base < j) {
#pragma unroll
                        for (int c = base; c < base + 8; c += 2)
                            *reinterpret_cast<double2*>(&rowline[c]) = make_double2(r[c], r[c + 1]);
                    }
                });
            }
            __syncwarp(gmask);
            double ljs = 0.0;
            if (gl < j) {
                const double cjs = rowline[gl];
                ljs = kStrict ? ddiv(cjs, dpos) : dmul(cjs, rdpos);
                lline[gl] = ljs;
            }
            __syncwarp(gmask);
            // ---- c_ij = l_ij - sum_s l_js c_is (gmw81.rs:83-85) ------------------------------
            double cnext = 0.0;
            const double cij = dispatch_gmw_column<NMAX, kStrict>(r, j, cj, lline, cnext);
            cj = cnext;
            // theta = max_{i>j} |c_ij|, 0 for the last column (gmw81.rs:87-100)
            const double theta = warp_extreme<true>(gmask, fabs(cij), below, 0.0, win);
            const double cjj = __shfl_sync(gmask, cij, q, G);
            const double tb = kStrict ? ddiv(theta, beta) : dmul(theta, rbeta);
            const double dj = fmax(fmax(fabs(cjj), dmul(tb, tb)), delta);  // gmw81.rs:102
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
            // running diagonal (gmw81.rs:104-109)
            if (below) {
                const double c2 = dmul(cij, cij);
                cd = kStrict ? dsub(cd, ddiv(c2, dj)) : __fma_rn(-c2, rdj, cd);
            }
            if (gl == q)
                e_mine = dsub(dj, cjj);  // gmw81.rs:110
            // ---- row j of L = [l_js sqrt(d_s)] , sqrt(d_j), 0...  (gmw81.rs:112-125) ---------
            if (gl < n) {
                double v = 0.0;
                if (gl < j)
                    v = dmul(ljs, sqd);
                else if (gl == j)
                    v = dmul(1.0, sq);
                Lb[(size_t)j * n + gl] = dadd(0.0, v);  // the dense dot's zero terms: -0 -> +0
            }
            if (gl == j) {
                dpos = dj;
                sqd = sq;
                rdpos = rdj;
            }
            __syncwarp(gmask);
        }

        if (valid) {
            E[(size_t)b * n + gl] = e_mine;                         // gmw81.rs:127-131
            P[(size_t)b * n + pos] = (unsigned long long)gl;
        }
        if (K && gl == 0 && live)
            K[b] = -1;
    }
}

}  // namespace modchol
