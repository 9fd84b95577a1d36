// wipp_kernels.cuh -- the implicit-pivoting FAST kernels of wip_kernels.cuh for 4 < n <= 16 with
// SEVERAL matrices per warp: G = 16 lanes per matrix (two matrices, 8 < n <= 16) or G = 8 (four
// matrices, 4 < n <= 8); lane g G + i owns row i of the warp's g-th matrix (gmw81.rs:39-134,
// se90.rs:38-181, se99.rs:35-201).  With one matrix per warp half (or three quarters) of the
// lanes idle through every step; here every lane works.
//
// What is warp-wide and what is per group of 16 lanes:
//   * the step counter j and the count of pending panel columns are warp-uniform: both matrices
//     retire column j in the same iteration, whatever phase each is in, so the trailing update
//     (mma.sync.m8n8k4.f64 needs the whole warp) and the tensor-memory store of column j of L
//     (tcgen05.st is warp-wide) sit at convergence points of the loop;
//   * every collective is issued with the FULL member mask and is segmented by hand (the first
//     version, on 16-lane member masks with per-group branches, ran every step once per group:
//     1 950 warp-instructions per matrix against 1 899 unpacked): two REDUX (one per group),
//     ballots split in halves, shuffles with a per-lane source.  All control flow is warp-uniform:
//     the code of a phase runs when ANY group is in that phase, with the other group's updates
//     predicated off;
//   * a phase switch does NOT flush the pending columns (that would need the tensor pipe in
//     divergent code): the Gershgorin bounds (se90.rs:105-110, se99.rs:125-130) are formed from
//     the stored slots minus the pending columns, element by element -- a cold path;
//   * norms use a shuffle tree over the 16 lanes (the DMMA sum of wip_kernels.cuh is warp-wide).
// Storage per matrix: packed lower triangle T(G) doubles, a row-major panel of the pending
// columns (G rows of kWppRowLd doubles, 128-bit reads conflict free), L in tensor memory (2 G
// columns per lane).
#pragma once
#include "wip_kernels.cuh"

namespace modchol {

constexpr int kWppKB = 4;                 // steps between two trailing updates
constexpr int kWppRowLd = kWppKB + 2;     // 48-byte panel rows
constexpr int kWppTmemCols = 32;
// doubles per matrix: triangle + panel, padded so that consecutive matrices start 64 bytes apart
// modulo 128 (two groups share a shared-memory wavefront of 8-byte accesses)
__host__ __device__ constexpr int wpp_mat_doubles(int G) { return G == 16 ? 232 : ((G * (G + 1) / 2 + G * kWppRowLd + 15) & ~15) + 8; }

// ---- collectives over the 32/G groups of a warp, all at once, full member mask ----------------
template <int G>
__device__ __forceinline__ int seg_max_s32(int grp, int v)
{
    if constexpr (G == 16) {
        const int r0 = __reduce_max_sync(kFull, grp == 0 ? v : (int)0x80000000);
        const int r1 = __reduce_max_sync(kFull, grp == 1 ? v : (int)0x80000000);
        return grp ? r1 : r0;
    }
    int mine = __reduce_max_sync(kFull, grp == 0 ? v : (int)0x80000000);
#pragma unroll
    for (int q = 1; q < 32 / G; ++q) {
        const int r = __reduce_max_sync(kFull, grp == q ? v : (int)0x80000000);
        mine = grp == q ? r : mine;
    }
    return mine;
}
template <int G>
__device__ __forceinline__ unsigned seg_max_u32(int grp, unsigned v)
{
    if constexpr (G == 16) {
        const unsigned r0 = __reduce_max_sync(kFull, grp == 0 ? v : 0u);
        const unsigned r1 = __reduce_max_sync(kFull, grp == 1 ? v : 0u);
        return grp ? r1 : r0;
    }
    unsigned mine = __reduce_max_sync(kFull, grp == 0 ? v : 0u);
#pragma unroll
    for (int q = 1; q < 32 / G; ++q) {
        const unsigned r = __reduce_max_sync(kFull, grp == q ? v : 0u);
        mine = grp == q ? r : mine;
    }
    return mine;
}
template <int G>
__device__ __forceinline__ unsigned seg_min_u32(int grp, unsigned v)
{
    if constexpr (G == 16) {
        const unsigned r0 = __reduce_min_sync(kFull, grp == 0 ? v : 0xffffffffu);
        const unsigned r1 = __reduce_min_sync(kFull, grp == 1 ? v : 0xffffffffu);
        return grp ? r1 : r0;
    }
    unsigned mine = __reduce_min_sync(kFull, grp == 0 ? v : 0xffffffffu);
#pragma unroll
    for (int q = 1; q < 32 / G; ++q) {
        const unsigned r = __reduce_min_sync(kFull, grp == q ? v : 0xffffffffu);
        mine = grp == q ? r : mine;
    }
    return mine;
}
// this group's part of a ballot, as bits of ABSOLUTE lane numbers
// (gbits: the lanes of this group, computed once per kernel)
__device__ __forceinline__ unsigned seg_ballot(unsigned gbits, bool pred)
{
    return __ballot_sync(kFull, pred) & gbits;
}

// lane (absolute) of the first maximum in permuted-position order (utils.rs:52-91) over the
// candidate lanes of each group; no candidate: the row at position j stays
template <int G>
__device__ __forceinline__ int wipp_pivot_lane(int grp, unsigned gbits, int khi, unsigned klo, bool cand, bool fin, int pos, int j)
{
    const int mh = seg_max_s32<G>(grp, cand ? khi : (int)0x80000000);
    const bool c1 = cand && khi == mh;
    unsigned b1 = seg_ballot(gbits, c1);
    if (__any_sync(kFull, __popc(b1) != 1)) {
        const unsigned ml = seg_max_u32<G>(grp, c1 ? klo : 0u);
        const bool win = c1 && klo == ml;
        unsigned mpos = seg_min_u32<G>(grp, win ? (unsigned)pos : 1024u);
        if (mpos > 255u)
            mpos = (unsigned)j;
        b1 = seg_ballot(gbits, !fin && pos == (int)mpos);
    }
    return b1 ? 31 - __clz(b1) : G * grp;   // (a group without any row alive: any of its lanes)
}

// max (kMax) / min over the candidate lanes of each group, NaN never taken, `none` without candidates
template <bool kMax, int G>
__device__ __forceinline__ double seg_extreme(int grp, double v, bool cand, double none)
{
    int khi;
    unsigned klo;
    make_key(v, khi, klo);
    cand = cand && (v == v);
    if (!kMax) {
        khi = ~khi;
        klo = ~klo;
    }
    const int mh = seg_max_s32<G>(grp, cand ? khi : (int)0x80000000);
    const unsigned ml = seg_max_u32<G>(grp, (cand && khi == mh) ? klo : 0u);
    if (mh == (int)0x80000000)
        return none;
    return kMax ? key_value(mh, ml) : key_value(~mh, ~ml);
}

template <int G>
__device__ __forceinline__ double wipp_sum(double x)
{
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1)
        x = dadd(x, __shfl_xor_sync(kFull, x, o));
    return x;
}

// kSolve: fused consumer step, as in wip_kernel (every group solves its own systems; the
// substitution's broadcasts are shuffles of width G under the full member mask).
template <int kAlgo, int G, int kN, int kMinBlocks, bool kSolve = false>
__global__ void __launch_bounds__(kWipWarps * 32, kMinBlocks)
wipp_kernel(int n_rt, long long batch, const double* __restrict__ A, double* __restrict__ L,
            double* __restrict__ E, unsigned long long* __restrict__ P, int* __restrict__ K,
            const double* Bm = nullptr, double* X = nullptr, int nrhs = 0)
{
    extern __shared__ double sm[];
    __shared__ unsigned tmem_base;
    const int n = kN ? kN : n_rt;
    int lane = threadIdx.x & 31;
    // through a shuffle: ptxas cannot re-derive the lane number (and the group index and masks that
    // follow from it) from S2R inside the step loop, as it does under register pressure otherwise
    lane = __shfl_sync(kFull, lane, lane);
    const int wib = __shfl_sync(kFull, (int)(threadIdx.x >> 5), 0);
    constexpr int kGroups = 32 / G;           // matrices per warp
    constexpr int NB = G / 8;                 // 8x8 block rows per matrix
    constexpr int kWppTri = G * (G + 1) / 2;
    constexpr int kWppMat = wpp_mat_doubles(G);
    const int sub = lane & (G - 1), grp = lane / G;
    const unsigned gbits = ((1u << G) - 1u) << (G * grp);   // the lanes of this group

    if (wib == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"((unsigned)__cvta_generic_to_shared(&tmem_base)), "n"(kWppTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tw = tmem_base + ((unsigned)(wib * 32) << 16);

    unsigned wbw = (unsigned)__cvta_generic_to_shared(sm + (size_t)wib * kGroups * kWppMat);   // this warp's matrices
    asm volatile("mov.u32 %0, %0;" : "+r"(wbw));
    const unsigned wb = wbw + 8u * (unsigned)(grp * kWppMat);     // this group's matrix
    const unsigned pnb = wb + 8u * kWppTri;                       // its panel
    const unsigned tib = wb + 8u * (unsigned)tri(sub);            // slot (sub, 0)
    const unsigned prow = pnb + 8u * (unsigned)(kWppRowLd * sub); // this lane's panel row
    const int r8 = lane >> 2, c4 = lane & 3;
    const bool pd0 = 2 * c4 <= r8, pd1 = 2 * c4 + 1 <= r8;
    const size_t nn = (size_t)n * n;
    const long long nslots = (long long)gridDim.x * kWipWarps * kGroups;
    const bool wide = kN != 0 && kN % 8 == 0 && (((size_t)A | (size_t)L) & 31) == 0;

    for (long long b0 = ((long long)blockIdx.x * kWipWarps + wib) * kGroups; b0 < batch; b0 += nslots) {
        const long long b = b0 + grp;
        const bool gact = b < batch;              // this group has a matrix
        const bool has = gact && sub < n;
        const double* Ab = A + (size_t)(gact ? b : b0) * nn;
        // ---- load: row by row into the packed triangle (lane = column) -------------------------
        double dg = has ? __ldg(Ab + (size_t)sub * (n + 1)) : 0.0;
        double asum = 0.0;
        unsigned long long om = 0ull;
        __syncwarp();   // the previous matrices' readers of the triangles are done
        {
            const double* src = Ab + sub;
            unsigned dst = wb + 8u * (unsigned)sub;
            auto take = [&](int i, double v) {
                sts_f64_if(sub <= i, dst, v);
                dst += 8u * (unsigned)(i + 1);
                if constexpr (kAlgo == kGMW81) {
                    const unsigned long long a = (unsigned long long)__double_as_longlong(v) & 0x7fffffffffffffffull;
                    om = (sub != i && a > om) ? a : om;
                } else {
                    asum = dadd(asum, fabs(v));
                }
            };
            int i = 0;
            for (; i + 8 <= n; i += 8) {
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    v[u] = has ? __ldg(src + (size_t)u * n) : 0.0;
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    take(i + u, v[u]);
                src += (size_t)8 * n;
            }
            for (; i < n; ++i) {
                take(i, has ? __ldg(src) : 0.0);
                src += n;
            }
            asum = dsub(asum, fabs(dg));
        }
#pragma unroll
        for (int c = 0; c < kWppKB; c += 2)
            asm volatile("st.shared.v2.f64 [%0], {%1, %1};" :: "r"(prow + 8u * c), "d"(0.0) : "memory");
        __syncwarp();

        // warp-wide: C(i,k) -= sum_c P(i,c) P(k,c) on the three 8x8 tiles of BOTH matrices; then every
        // lane clears its own panel row
        auto rank_update = [&]() {
#pragma unroll
            for (int m = 0; m < kGroups; ++m) {
                const unsigned mb = wbw + 8u * (unsigned)(m * kWppMat);
                const unsigned pf = mb + 8u * (unsigned)(kWppTri + kWppRowLd * r8 + c4);
                const unsigned fr0 = mb + 8u * (unsigned)(tri(r8) + 2 * c4);
                if constexpr (NB == 2) {
                    double f[2], c0[3], c1[3];
#pragma unroll
                    for (int q = 0; q < 2; ++q)
                        f[q] = lds_f64(pf + 8u * (unsigned)(8 * q * kWppRowLd));
                    const unsigned ra1 = fr0 + 8u * (unsigned)(tri(8) + 8 * r8);
                    c0[0] = lds_f64(fr0);        c1[0] = lds_f64(fr0 + 8);         // tile (0,0)
                    c0[1] = lds_f64(ra1);        c1[1] = lds_f64(ra1 + 8);         // tile (1,0)
                    c0[2] = lds_f64(ra1 + 64u);  c1[2] = lds_f64(ra1 + 72u);       // tile (1,1)
                    const double nf0 = wip_neg(f[0]), nf1 = wip_neg(f[1]);
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                                 : "+d"(c0[0]), "+d"(c1[0]) : "d"(nf0), "d"(f[0]));
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                                 : "+d"(c0[1]), "+d"(c1[1]) : "d"(nf1), "d"(f[0]));
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                                 : "+d"(c0[2]), "+d"(c1[2]) : "d"(nf1), "d"(f[1]));
                    sts_f64_if(pd0, fr0, c0[0]);
                    sts_f64_if(pd1, fr0 + 8, c1[0]);
                    sts_f64(ra1, c0[1]);
                    sts_f64(ra1 + 8, c1[1]);
                    sts_f64_if(pd0, ra1 + 64u, c0[2]);
                    sts_f64_if(pd1, ra1 + 72u, c1[2]);
                    continue;
                }
                double f[NB], c0[NB * (NB + 1) / 2], c1[NB * (NB + 1) / 2];
#pragma unroll
                for (int q = 0; q < NB; ++q)
                    f[q] = lds_f64(pf + 8u * (unsigned)(8 * q * kWppRowLd));
#pragma unroll
                for (int bi = 0; bi < NB; ++bi)
#pragma unroll
                    for (int bj = 0; bj <= bi; ++bj) {
                        const unsigned ra = fr0 + 8u * (unsigned)(tri(8 * bi) + 8 * bi * r8) + 64u * bj;
                        c0[bi * (bi + 1) / 2 + bj] = lds_f64(ra);
                        c1[bi * (bi + 1) / 2 + bj] = lds_f64(ra + 8);
                    }
#pragma unroll
                for (int bi = 0; bi < NB; ++bi)
#pragma unroll
                    for (int bj = 0; bj <= bi; ++bj) {
                        const int t = bi * (bi + 1) / 2 + bj;
                        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                                     : "+d"(c0[t]), "+d"(c1[t]) : "d"(wip_neg(f[bi])), "d"(f[bj]));
                    }
#pragma unroll
                for (int bi = 0; bi < NB; ++bi)
#pragma unroll
                    for (int bj = 0; bj <= bi; ++bj) {
                        const int t = bi * (bi + 1) / 2 + bj;
                        const unsigned ra = fr0 + 8u * (unsigned)(tri(8 * bi) + 8 * bi * r8) + 64u * bj;
                        if (bi == bj) {
                            sts_f64_if(pd0, ra, c0[t]);
                            sts_f64_if(pd1, ra + 8, c1[t]);
                        } else {
                            sts_f64(ra, c0[t]);
                            sts_f64(ra + 8, c1[t]);
                        }
                    }
            }
            __syncwarp();
#pragma unroll
            for (int c = 0; c < kWppKB; c += 2)
                asm volatile("st.shared.v2.f64 [%0], {%1, %1};" :: "r"(prow + 8u * c), "d"(0.0) : "memory");
            __syncwarp();
        };

        double ev = 0.0, g = 0.0;
        int pos = sub;
        bool fin = !has;
        unsigned fmask = (n >= G ? 0u : (0xffffffffu << n)) & ((1u << G) - 1u);   // finished rows of this matrix (bit = row)
        int cnt = 0, kstart = -1;
        bool win;

        // current value of C(sub, r) of this group's matrix: stored slot minus the pending columns
        // (r: ABSOLUTE lane of the pivot row; rr = its row index)
        auto column = [&](int rr) -> double {
            const unsigned tr = wb + 4u * (unsigned)(rr * (rr + 1));
            double col = lds_f64(sub > rr ? tib + 8u * (unsigned)rr : tr + 8u * (unsigned)sub);
            if (cnt > 0) {
                const unsigned pq = pnb + 8u * (unsigned)(kWppRowLd * rr);
                const double2 o0 = lds_v2f64(prow), o1 = lds_v2f64(prow + 16u);
                const double2 q0 = lds_v2f64(pq), q1 = lds_v2f64(pq + 16u);
                double s0 = dmul(o0.x, q0.x), s1 = dmul(o0.y, q0.y);
                s0 = __fma_rn(o1.x, q1.x, s0);
                s1 = __fma_rn(o1.y, q1.y, s1);
                col = dsub(col, dadd(s0, s1));
            }
            return col;
        };
        // the pivot row r of every group with `mine` takes position j
        auto take_position = [&](bool mine, int r, int j) {
            const int pr = __shfl_sync(kFull, pos, r);
            const int np = lane == r ? j : (pos == j ? pr : pos);
            pos = mine ? np : pos;
        };
        // group-local part of retiring column j: panel entry, row r finished
        auto retire = [&](bool mine, int r, double cs) {
            sts_f64_if(mine, prow + 8u * (unsigned)cnt, cs);
            fin = fin || (mine && lane == r);
            if (mine)
                fmask |= 1u << (r & (G - 1));
        };

        // per-group algorithm state
        int phase = !gact ? 3 : (kAlgo == kGMW81 ? 0 : 1);   // 0 GMW81, 1 SE phase 1, 2 SE phase 2, 3 nothing left
        double gamma, taugamma, delta = 0.0, rbeta = 0.0, delta_prev = 0.0, dmin_carry = 0.0;
        bool have_dmin = false;
        double lnext = 0.0;   // the last column's entry of L computed by a 2x2 tail
        // gamma / xi_diag = max |a_ii| folded from 0 (gmw81.rs:49-51, se90.rs:55-57, se99.rs:54-56)
        gamma = seg_extreme<true, G>(grp, fabs(dg), has, 0.0);
        taugamma = dmul(MODCHOL_TAU, gamma);
        if constexpr (kAlgo == kGMW81) {
            const unsigned ohi = seg_max_u32<G>(grp, (unsigned)(om >> 32));
            const unsigned olo = seg_max_u32<G>(grp, (unsigned)(om >> 32) == ohi ? (unsigned)om : 0u);
            const double off_max = __hiloint2double((int)ohi, (int)olo);
            const double nf = (double)n;
            delta = dmul(MODCHOL_EPS, fmax(1.0, dadd(gamma, off_max)));
            const double beta =
                dsqrt(fmax(fmax(gamma, ddiv(off_max, dsqrt(dsub(dmul(nf, nf), 1.0)))), MODCHOL_EPS));
            rbeta = ddiv(1.0, beta);
        } else {
            g = dsub(dg, asum);   // Gershgorin bound if phase 2 starts at column 0
        }

        for (int j = 0; j < n; ++j) {
            double lval = 0.0;   // this lane's entry of column j of L
            if constexpr (kAlgo == kGMW81) {
                // ---- GMW81, symmetric form (see wsm_gmw.cuh) ----------------------------------------
                const bool mine = phase == 0;
                const int khi = __double2hiint(dg) & 0x7fffffff;
                const unsigned klo = (unsigned)__double2loint(dg);
                const int r = wipp_pivot_lane<G>(grp, gbits, khi, klo, mine && !fin && dg == dg, fin, pos, j);   // gmw81.rs:71-78
                take_position(mine, r, j);
                const double col = column(r & (G - 1));
                const double cjj = __shfl_sync(kFull, col, r);
                const bool live = !fin && lane != r;
                const double theta = seg_extreme<true, G>(grp, fabs(col), live, 0.0);   // gmw81.rs:87-100
                const double tb = dmul(theta, rbeta);
                const double dj = fmax(fmax(fabs(cjj), dmul(tb, tb)), delta);        // gmw81.rs:102
                double y = rsqrt_normal(dj);
                double sq = dmul(dj, y);
                if (!fast_range_ok(dj))
                    wip_slow_gmw(dj, sq, y);
                const double cs = fin ? 0.0 : dmul(col, y);
                dg = mine ? __fma_rn(-cs, cs, dg) : dg;                              // gmw81.rs:104-109
                ev = (mine && lane == r) ? dsub(dj, cjj) : ev;                       // gmw81.rs:110
                lval = mine ? (lane == r ? sq : cs) : 0.0;
                retire(mine, r, cs);
            } else {
                if (__any_sync(kFull, phase == 1)) {
                    // ---- SE phase 1 (se90.rs:61-96, se99.rs:61-109) ---------------------------------
                    const bool mine = phase == 1;
                    int khi;
                    unsigned klo;
                    wip_key(dg, khi, klo);
                    const int r = wipp_pivot_lane<G>(grp, gbits, khi, klo, mine && !fin && dg == dg, fin, pos, j);
                    const double piv = __shfl_sync(kFull, dg, r);
                    bool sw = false;
                    if (kAlgo == kSE99) {   // se99.rs:62-73, before the pivot is applied
                        double dmin = dmin_carry;
                        if (__any_sync(kFull, mine && !have_dmin)) {
                            const double fresh = seg_extreme<false, G>(grp, dg, mine && !fin, INFINITY);
                            dmin = have_dmin ? dmin_carry : fresh;
                        }
                        const double dmax = (piv == piv) ? piv : -INFINITY;
                        sw = mine && (dmax < taugamma || dmin < dmul(-MODCHOL_MU, dmax));
                    }
                    const bool go = mine && !sw;   // this group applies its pivot
                    take_position(go, r, j);
                    const double col = column(r & (G - 1));
                    const bool live = go && !fin && lane != r;
                    const bool fast = fast_range_ok(piv);
                    const double y = rsqrt_normal(piv);
                    double sq = dmul(piv, y);
                    double cs = dmul(col, y);
                    double dgn = __fma_rn(-cs, cs, dg);
                    double nv = dgn;   // look-ahead value (se90.rs:70-77, se99.rs:82-89)
                    if (!fast)
                        wip_slow_p1(piv, col, dg, sq, cs, dgn, nv);
                    const double la = seg_extreme<false, G>(grp, nv, live, INFINITY);
                    const double thresh = (kAlgo == kSE99) ? dmul(-MODCHOL_MU, gamma) : taugamma;
                    const bool sw2 = go && la < thresh;   // the pivot at j stays applied (se99.rs:90-93)
                    const bool done = go && !sw2;         // column j is retired in phase 1
                    have_dmin = go ? fast : have_dmin;
                    dmin_carry = go ? la : dmin_carry;
                    cs = fin ? 0.0 : cs;
                    dg = done ? dgn : dg;
                    lval = done ? (lane == r ? sq : cs) : lval;
                    retire(done, r, cs);
                    const bool swn = sw || sw2;
                    if (__any_sync(kFull, swn)) {
                        kstart = swn ? j : kstart;
                        phase = swn ? 2 : phase;
                        if (__any_sync(kFull, swn && j > 0 && !(kAlgo == kSE99 && j == n - 1))) {
                            // Gershgorin bounds of the trailing block as the reference holds it
                            // (se90.rs:105-110, se99.rs:125-130): stored slots minus the pending columns,
                            // element by element (no flush: the trailing update is a warp-wide operation)
                            const double2 o0 = lds_v2f64(prow), o1 = lds_v2f64(prow + 16u);
                            double sacc = 0.0;
                            unsigned tk = wb;
                            for (int k = 0; k < n; ++k) {
                                const bool alive = ((fmask >> k) & 1u) == 0u;
                                double v = lds_f64(sub > k ? tib + 8u * (unsigned)k : tk + 8u * (unsigned)sub);
                                const unsigned pq = pnb + 8u * (unsigned)(kWppRowLd * k);
                                const double2 q0 = lds_v2f64(pq), q1 = lds_v2f64(pq + 16u);
                                double s0 = dmul(o0.x, q0.x), s1 = dmul(o0.y, q0.y);
                                s0 = __fma_rn(o1.x, q1.x, s0);
                                s1 = __fma_rn(o1.y, q1.y, s1);
                                v = dsub(v, dadd(s0, s1));
                                if (alive && k != sub)
                                    sacc = dadd(sacc, fabs(v));
                                tk += 8u * (unsigned)(k + 1);
                            }
                            g = (swn && j > 0) ? dsub(dg, sacc) : g;
                        }
                    }
                }
                if (__any_sync(kFull, phase == 2)) {
                    // ---- SE phase 2 (se90.rs:101-167, se99.rs:121-187) ------------------------------
                    const bool mine = phase == 2;
                    if (j + 2 < n) {
                        int khi;   // se90.rs:115, se99.rs:135
                        unsigned klo;
                        wip_key(g, khi, klo);
                        const int r = wipp_pivot_lane<G>(grp, gbits, khi, klo, mine && !fin && g == g, fin, pos, j);
                        take_position(mine, r, j);
                        double piv = __shfl_sync(kFull, dg, r);
                        const double col = column(r & (G - 1));
                        const bool live = mine && !fin && lane != r;
                        const double acol = live ? abs_bits(col) : 0.0;
                        // E_jj (se90.rs:124-131, se99.rs:144-151)
                        const double normj = wipp_sum<G>(acol);
                        const double ej = max_b_not_nan(dadd(-piv, max_b_not_nan(normj, taugamma)), delta_prev);
                        piv = dadd(piv, ej);
                        delta_prev = mine ? ej : delta_prev;
                        // bound update (se90.rs:134-139, se99.rs:154-159) and column scaling
                        const bool upd = fabs(dsub(piv, normj)) > MODCHOL_EPS;
                        const double y = rsqrt_normal(piv);
                        double sq = dmul(piv, y);
                        double t = __fma_rn(-normj, dmul(y, y), 1.0);
                        double cs = dmul(col, y);
                        if (!fast_range_ok(piv))
                            wip_slow_p2(piv, normj, col, sq, t, cs);
                        cs = fin ? 0.0 : cs;
                        dg = mine ? __fma_rn(-cs, cs, dg) : dg;
                        g = (mine && upd) ? __fma_rn(acol, t, g) : g;
                        ev = (mine && lane == r) ? ej : ev;
                        lval = mine ? (lane == r ? sq : cs) : lval;
                        retire(mine, r, cs);
                    } else if (j == n - 2) {
                        // final 2x2 block, never pivoted (se90.rs:155-166, se99.rs:175-186): the two rows
                        // alive sit at positions n-2 and n-1
                        const unsigned bq0 = seg_ballot(gbits, !fin && pos == n - 2);
                        const unsigned bq1 = seg_ballot(gbits, !fin && pos == n - 1);
                        const int u0 = bq0 ? 31 - __clz(bq0) : G * grp, u1 = bq1 ? 31 - __clz(bq1) : G * grp;
                        double a00 = __shfl_sync(kFull, dg, u0);
                        double a11 = __shfl_sync(kFull, dg, u1);
                        const double c2 = column(u0 & (G - 1));
                        double a10 = __shfl_sync(kFull, c2, u1);
                        double lhi, llo;
                        eigenvalues_2x2(a00, a10, a10, a11, lhi, llo);   // lower storage: m[0,1] == m[1,0]
                        const double en = tail2x2_e(kAlgo, lhi, llo, gamma, delta_prev);
                        if (en > 0.0) {
                            a00 = dadd(a00, en);
                            a11 = dadd(a11, en);
                        }
                        a00 = dsqrt(a00);
                        a10 = ddiv(a10, a00);
                        a11 = dsqrt(dsub(a11, dmul(a10, a10)));
                        ev = (mine && (lane == u0 || lane == u1)) ? en : ev;
                        lval = mine ? (lane == u0 ? a00 : (lane == u1 ? a10 : 0.0)) : lval;
                        lnext = mine ? (lane == u1 ? a11 : 0.0) : lnext;
                    } else {   // j == n - 1: the last entry of a 2x2 tail, or SE99's 1x1 tail (se99.rs:115-119)
                        const bool one = mine && kAlgo == kSE99 && kstart == n - 1;
                        const unsigned ba = seg_ballot(gbits, !fin);
                        const int u = ba ? 31 - __clz(ba) : G * grp;
                        const double ljj = __shfl_sync(kFull, dg, u);
                        const double ej = tail1x1_e(ljj, gamma);
                        ev = (one && lane == u) ? ej : ev;
                        const double v1 = lane == u ? dsqrt(dadd(ljj, ej)) : 0.0;
                        lval = mine ? (one ? v1 : lnext) : lval;
                        phase = mine ? 3 : phase;
                    }
                }
            }
            // ---- warp-wide: column j of L into tensor memory; trailing update every kWppKB steps ----
            __syncwarp();
            wip_tmem_store(tw + 2u * (unsigned)j, lval);
            if (++cnt == kWppKB) {
                if (j + 1 < n)
                    rank_update();
                cnt = 0;
            }
        }

        // ---- epilogue: this lane's row of L leaves tensor memory for row `pos` of its matrix ----
        __syncwarp();
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        {
            double* dstp = L ? L + (size_t)(gact ? b : b0) * nn + (size_t)pos * n : nullptr;
            const unsigned rowb = wb + 8u * (unsigned)tri(pos);   // row pos of the position-ordered triangle
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                if (8 * c >= n)
                    break;
                double v[8];
                wip_tmem_load8(tw + 16u * c, v);
                if constexpr (kSolve) {
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        sts_f64_if(has && 8 * c + u <= pos, rowb + 8u * (unsigned)(8 * c + u), v[u]);
                }
                if (dstp && has) {
                    if (wide) {
                        stg_v4f64(dstp + 8 * c, v[0], v[1], v[2], v[3]);
                        stg_v4f64(dstp + 8 * c + 4, v[4], v[5], v[6], v[7]);
                    } else {
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            if (8 * c + u < n)
                                dstp[8 * c + u] = v[u];
                    }
                }
            }
        }
        if (has) {
            if (E)
                E[(size_t)b * n + sub] = ev;
            if (P)
                P[(size_t)b * n + pos] = (unsigned long long)sub;
        }
        if (K && gact && sub == 0)
            K[b] = kstart;
        if constexpr (kSolve) {
            // lane -> position through the (now free) panel area, then the substitutions on the
            // position-ordered triangle
            sts_f64(pnb + 8u * (unsigned)pos, (double)sub);
            __syncwarp();
            const unsigned tibs[1] = {tib};
            const int rows[1] = {sub};
            const bool rvs[1] = {has};
            const int perms[1] = {(int)lds_f64(pnb + 8u * (unsigned)sub)};
            const size_t off = (size_t)(gact ? b : b0) * nrhs * n;
            wsm_solve_in_place<G, 1>(kFull, n, wb, tibs, rows, rvs, perms, Bm + off, X + off, nrhs);
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (wib == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                     :: "r"(tmem_base), "n"(kWppTmemCols) : "memory");
}

}  // namespace modchol
