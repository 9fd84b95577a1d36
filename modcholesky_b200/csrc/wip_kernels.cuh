// wip_kernels.cuh -- FAST-mode GMW81 / SE90 / SE99 for 8 < n <= 32 with IMPLICIT pivoting:
// one warp per matrix, lane i owns ORIGINAL row i for the whole factorisation and nothing is
// ever swapped (gmw81.rs:39-134, se90.rs:38-181, se99.rs:35-201; swap of utils.rs:30-49 is
// replaced by the per-lane `pos`, the permuted position the reference would hold the row at).
//
//   * storage: packed lower triangle in shared memory in ORIGINAL coordinates, element
//     (i,k), i >= k, at T(i)+k.  The slot of (i, r) is dead as soon as row/column r has been
//     the pivot, so the factor is stored in place: after step t with pivot row r the slot
//     (max(i,r), min(i,r)) holds L(i,t) for every row i still alive and (r,r) holds L(r,t).
//   * a step reads the pivot column straight from those slots: lane i > r reads T(i)+r
//     (the triangular numbers are a permutation modulo 16 on aligned runs of 16 rows), lane
//     i < r reads T(r)+i (contiguous) -- one LDS per lane, no swap, no row scalars to exchange.
//   * the trailing block is brought up to date every FOUR steps on the FP64 tensor pipe
//     (mma.sync.m8n8k4.f64): the four new columns go through a 32x4 panel in shared memory
//     whose layout makes both the per-step column write and the fragment load conflict free;
//     between updates a column needs at most three fused multiply-adds per lane.  Tiles are
//     fixed 8x8 blocks of the original index space; a store is predicated on "row and column
//     both still alive" so that finished slots keep their L entries.
//   * pivot search: order-preserving integer keys + ONE REDUX when the high word of the
//     maximum is unique (practically always); ties fall back to the full first-in-position
//     search of utils.rs:52-91 on (hi, lo, pos).
//   * the Gershgorin bounds of a phase 2 that starts at column 0 (symmetric-uniform inputs)
//     come from the absolute column sums accumulated while the matrix is loaded.
// FAST arithmetic only (fused multiply-add, rsqrt scaling, tree-ordered norms): results agree
// with the oracle within the north-star tolerances, not bit for bit -- STRICT mode keeps the
// position-ordered kernels of wsm_se.cuh / wsm_gmw.cuh.
#pragma once
#include "wsm_se.cuh"

namespace modchol {

constexpr int kWipTri = 528;       // T(32) doubles of packed triangle
constexpr int kWipPanelLd = 36;    // panel column stride: == 4 (mod 16) doubles, see above
constexpr int kWipDoubles = kWipTri + 4 * kWipPanelLd + 16;   // + inverse permutation (32 ints)

using IC0 = std::integral_constant<int, 0>;
using IC1 = std::integral_constant<int, 1>;
using IC2 = std::integral_constant<int, 2>;
using IC3 = std::integral_constant<int, 3>;

__device__ __forceinline__ void sts_u32(unsigned a, unsigned v)
{
    asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned lds_u32(unsigned a)
{
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}

// order-preserving key without the -0 fold of make_key (FAST: +-0 ties are rounding ties)
__device__ __forceinline__ void wip_key(double v, int& khi, unsigned& klo)
{
    const int hi = __double2hiint(v);
    const int m = hi >> 31;
    khi = hi ^ (m & 0x7fffffff);
    klo = (unsigned)__double2loint(v) ^ (unsigned)m;
}

// Lane that holds the first maximum (in permuted-position order, utils.rs:52-91) of the keyed
// value over the candidate lanes.  NaN lanes are not candidates; without any candidate the
// row at position j stays (the reference's index 0).
__device__ __forceinline__ int wip_pivot_lane(int khi, unsigned klo, bool cand, bool fin, int pos, int j)
{
    const int mh = __reduce_max_sync(kFull, cand ? khi : (int)0x80000000);
    const bool c1 = cand && khi == mh;
    const unsigned b1 = __ballot_sync(kFull, c1);
    if (__popc(b1) == 1)
        return __ffs(b1) - 1;
    const unsigned ml = __reduce_max_sync(kFull, c1 ? klo : 0u);
    const bool win = c1 && klo == ml;
    unsigned mpos = __reduce_min_sync(kFull, win ? (unsigned)pos : 1024u);
    if (mpos > 255u)
        mpos = (unsigned)j;
    return __ffs(__ballot_sync(kFull, !fin && pos == (int)mpos)) - 1;
}

// Trailing update with the cnt (<= 4) columns of the panel:
//     C(i,k) -= sum_c P(i,c) P(k,c)      for every pair of rows i >= k that are both alive,
// one DMMA per 8x8 tile of the original index space.  Fragments: lane l holds
// A[l/4][l%4] = P(8 bi + l/4, l%4), the same pattern serves as B for block column bj, and
// C[l/4][2 (l%4) + {0,1}].  fmask = finished rows (bit i), including the rows >= n.
template <int NB>
__device__ __noinline__ void wip_rank_update(unsigned wb, unsigned pnb, unsigned fmask, int cnt, int lane)
{
    const int r8 = lane >> 2, c4 = lane & 3;
    double frag[NB];
    const unsigned pf = pnb + 8u * (unsigned)(kWipPanelLd * c4 + r8);
#pragma unroll
    for (int b = 0; b < NB; ++b)
        frag[b] = lds_f64_or_zero(c4 < cnt, pf + 64u * b);
    const unsigned rowdead = fmask >> r8;          // bit 8 bi     : row 8 bi + r8 is finished
    const unsigned coldead = fmask >> (2 * c4);    // bit 8 bj + e : column 8 bj + 2 c4 + e
    const unsigned r0 = wb + 8u * (unsigned)(tri(r8) + 2 * c4);
#pragma unroll
    for (int bi = 0; bi < NB; ++bi) {
        if (((fmask >> (8 * bi)) & 0xffu) == 0xffu)
            continue;
        // T(8 bi + r8) = T(8 bi) + T(r8) + 8 bi r8
        const unsigned ra = r0 + 8u * (unsigned)(tri(8 * bi) + 8 * bi * r8);
        const bool rl = ((rowdead >> (8 * bi)) & 1u) == 0u;
        const double na = -frag[bi];
#pragma unroll
        for (int bj = 0; bj <= bi; ++bj) {
            if (((fmask >> (8 * bj)) & 0xffu) == 0xffu)
                continue;
            bool p0 = rl && ((coldead >> (8 * bj)) & 1u) == 0u;
            bool p1 = rl && ((coldead >> (8 * bj + 1)) & 1u) == 0u;
            if (bi == bj) {
                p0 = p0 && 2 * c4 <= r8;
                p1 = p1 && 2 * c4 + 1 <= r8;
            }
            const unsigned a = ra + 64u * bj;
            const double c0 = lds_f64(a), c1 = lds_f64(a + 8);
            double d0, d1;
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%4, %5};"
                         : "=d"(d0), "=d"(d1) : "d"(na), "d"(frag[bj]), "d"(c0), "d"(c1));
            sts_f64_if(p0, a, d0);
            sts_f64_if(p1, a + 8, d1);
        }
    }
}

template <int kAlgo, int NB, int kWarps, int kMinBlocks>
__global__ void __launch_bounds__(kWarps * 32, kMinBlocks)
wip_kernel(int n, long long batch, const double* __restrict__ A, double* __restrict__ L,
           double* __restrict__ E, unsigned long long* __restrict__ P, int* __restrict__ K)
{
    extern __shared__ double sm[];
    int lane = threadIdx.x & 31;
    asm volatile("mov.u32 %0, %0;" : "+r"(lane));   // pinned (see wsm_se.cuh)
    const int wib = threadIdx.x >> 5;
    unsigned wb = (unsigned)__cvta_generic_to_shared(sm + (size_t)wib * kWipDoubles);
    asm volatile("mov.u32 %0, %0;" : "+r"(wb));
    const unsigned pnb = wb + 8u * kWipTri;                  // 4 panel columns of kWipPanelLd
    const unsigned invb = pnb + 8u * 4 * kWipPanelLd;        // inverse permutation, 32 ints
    unsigned tib = wb + 8u * (unsigned)tri(lane);            // byte address of slot (lane, 0)
    asm volatile("mov.u32 %0, %0;" : "+r"(tib));
    const unsigned pcol = pnb + 8u * (unsigned)lane;         // this lane's entry of panel column 0
    const size_t nn = (size_t)n * n;
    const long long nslots = (long long)gridDim.x * kWarps;
    const bool has = lane < n;

    for (long long b = (long long)blockIdx.x * kWarps + wib; b < batch; b += nslots) {
        // ---- load: row i of A by the whole warp (lane = column), lower part into the packed
        // triangle; this lane's diagonal element and absolute off-diagonal column sum (== row
        // sum: A symmetric) on the fly -------------------------------------------------------
        const double* src = A + (size_t)b * nn + lane;
        unsigned dst = wb + 8u * (unsigned)lane;
        double dg = 0.0, asum = 0.0;
        unsigned long long omax = 0ull;   // GMW81: max |a_ik|, i != k, as an integer (gmw81.rs:52-58)
        __syncwarp();
        auto take = [&](int i, double v) {
            sts_f64_if(lane <= i, dst, v);
            dst += 8u * (unsigned)(i + 1);
            if (lane == i) {
                dg = v;
            } else if (kAlgo == kGMW81) {
                const unsigned long long a = (unsigned long long)__double_as_longlong(v) & 0x7fffffffffffffffull;
                omax = a > omax ? a : omax;   // |.| of non-NaN doubles order like their bits
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
        __syncwarp();

        double ev = 0.0;
        int pos = lane;                       // permuted position of this lane's row
        bool fin = !has;                      // row already eliminated (or no row)
        unsigned fmask = n >= 32 ? 0u : (0xffffffffu << n);
        int j = 0, cnt = 0, kstart = -1;
        double csr[3] = {0.0, 0.0, 0.0};      // this row's entries of the pending panel columns
        bool win;

        // current value of C(lane, r): stored slot minus the C pending panel columns
        auto column = [&](auto ctag, int r, unsigned& addr) -> double {
            constexpr int C = decltype(ctag)::value;
            const unsigned tr = wb + 8u * (unsigned)tri(r);
            addr = lane > r ? tib + 8u * (unsigned)r : tr + 8u * (unsigned)lane;
            double col = lds_f64(addr);
#pragma unroll
            for (int c = 0; c < C; ++c)
                col = __fma_rn(-csr[c], lds_f64(pnb + 8u * (unsigned)(kWipPanelLd * c + r)), col);
            return col;
        };
        auto column_rt = [&](int c, int r, unsigned& addr) -> double {
            switch (c) {
            case 0: return column(IC0{}, r, addr);
            case 1: return column(IC1{}, r, addr);
            case 2: return column(IC2{}, r, addr);
            default: return column(IC3{}, r, addr);
            }
        };
        // s = sqrt(pivot), y with cs = col * y, 1 / pivot (FAST formulas inside the safe range)
        auto scale = [&](double piv, double& sq, double& y, double& inv_piv) -> bool {
            const bool fast = fast_range_ok(piv);
            if (fast) {
                y = rsqrt_normal(piv);
                sq = dmul(piv, y);
                inv_piv = dmul(y, y);
            } else {
                sq = dsqrt(piv);
                y = 0.0;
                inv_piv = 0.0;
            }
            return fast;
        };
        // the pivot row r takes position j; the row that sat there takes r's old position
        auto take_position = [&](int r) {
            const int pr = __shfl_sync(kFull, pos, r);
            if (pos == j)
                pos = pr;
            if (lane == r)
                pos = j;
        };
        // column j of L: cs into the dead slots of column r and into panel column C, sq on the
        // diagonal; row r is finished
        auto retire = [&](auto ctag, int r, unsigned addr, double cs, double sq) {
            constexpr int C = decltype(ctag)::value;
            sts_f64_if(!fin, addr, lane == r ? sq : cs);
            sts_f64(pcol + 8u * (unsigned)(kWipPanelLd * C), cs);
            if constexpr (C < 3)
                csr[C] = cs;
            if (lane == r)
                fin = true;
            fmask |= 1u << r;
            __syncwarp();
            ++j;
        };

        if constexpr (kAlgo == kGMW81) {
            // ---- GMW81, symmetric form (see wsm_gmw.cuh): slots hold c_is / sqrt(d_s) ---------
            const double diag_max = warp_extreme<true>(kFull, fabs(dg), has, 0.0, win);
            const unsigned ohi = __reduce_max_sync(kFull, (unsigned)(omax >> 32));
            const unsigned olo = __reduce_max_sync(kFull, (unsigned)(omax >> 32) == ohi ? (unsigned)omax : 0u);
            const double off_max = __hiloint2double((int)ohi, (int)olo);
            const double nf = (double)n;
            const double delta = dmul(MODCHOL_EPS, fmax(1.0, dadd(diag_max, off_max)));
            const double beta =
                dsqrt(fmax(fmax(diag_max, ddiv(off_max, dsqrt(dsub(dmul(nf, nf), 1.0)))), MODCHOL_EPS));
            const double rbeta = ddiv(1.0, beta);
            auto gmw_step = [&](auto ctag) {
                // pivot on max |c_ii| (gmw81.rs:71-78)
                const int khi = __double2hiint(dg) & 0x7fffffff;
                const unsigned klo = (unsigned)__double2loint(dg);
                const int r = wip_pivot_lane(khi, klo, !fin && dg == dg, fin, pos, j);
                take_position(r);
                unsigned addr;
                const double col = column(ctag, r, addr);   // lane r: c_jj from the same updates
                const double cjj = __shfl_sync(kFull, col, r);
                const bool live = !fin && lane != r;
                // theta = max_{i>j} |c_ij| (gmw81.rs:87-100), d_j (gmw81.rs:102)
                const double theta = warp_extreme<true>(kFull, fabs(col), live, 0.0, win);
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
                const double cs = dmul(col, y);
                if (live)
                    dg = __fma_rn(-cs, cs, dg);                 // gmw81.rs:104-109
                if (lane == r)
                    ev = dsub(dj, cjj);                         // gmw81.rs:110
                retire(ctag, r, addr, cs, sq);
            };
            while (j < n) {
                gmw_step(IC0{});
                if (j == n) break;
                gmw_step(IC1{});
                if (j == n) break;
                gmw_step(IC2{});
                if (j == n) break;
                gmw_step(IC3{});
                if (j == n) break;
                wip_rank_update<NB>(wb, pnb, fmask, 4, lane);
                __syncwarp();
            }
        } else {
            double g = dsub(dg, asum);            // Gershgorin bound if phase 2 starts at column 0
            // gamma = max |a_ii| folded from 0.0 (se90.rs:55-57, se99.rs:54-56)
            const double gamma = warp_extreme<true>(kFull, fabs(dg), has, 0.0, win);
            const double taugamma = dmul(MODCHOL_TAU, gamma);
            double delta_prev = 0.0;
            bool have_dmin = false;
            double dmin_carry = 0.0;

            // ---- phase 1 (se90.rs:61-96, se99.rs:61-109): 0 go on, 1 switch to phase 2, 2 done --
            auto p1_step = [&](auto ctag) -> int {
                int khi;
                unsigned klo;
                wip_key(dg, khi, klo);
                const int r = wip_pivot_lane(khi, klo, !fin && dg == dg, fin, pos, j);
                const double piv = __shfl_sync(kFull, dg, r);
                if (kAlgo == kSE99) {   // se99.rs:62-73, before the pivot is applied
                    // the minimum of the diagonal over the rows alive IS the previous step's
                    // look-ahead minimum (same rows, same values)
                    const double dmin = have_dmin ? dmin_carry
                                                  : warp_extreme<false>(kFull, dg, !fin, INFINITY, win);
                    const double dmax = (piv == piv) ? piv : -INFINITY;
                    if (dmax < taugamma || dmin < dmul(-MODCHOL_MU, dmax)) {
                        kstart = j;
                        return 1;
                    }
                }
                take_position(r);
                unsigned addr;
                const double col = column(ctag, r, addr);
                const bool live = !fin && lane != r;
                double sq, y, inv_piv;
                const bool fast = scale(piv, sq, y, inv_piv);
                const double cs = fast ? dmul(col, y) : ddiv(col, sq);
                const double dgn = __fma_rn(-cs, cs, dg);
                // look-ahead value (se90.rs:70-77, se99.rs:82-89); FAST: the updated diagonal
                const double nv = fast ? dgn : dsub(dg, ddiv(dmul(col, col), piv));
                const double la = warp_extreme<false>(kFull, nv, live, INFINITY, win);
                have_dmin = fast;
                dmin_carry = la;
                const double thresh = (kAlgo == kSE99) ? dmul(-MODCHOL_MU, gamma) : taugamma;
                if (la < thresh) {   // the pivot at j stays applied (se99.rs:90-93)
                    kstart = j;
                    return 1;
                }
                if (live)
                    dg = dgn;
                retire(ctag, r, addr, cs, sq);
                return j == n ? 2 : 0;
            };
            int st;
            for (;;) {
                cnt = 0;
                if ((st = p1_step(IC0{})) != 0) break;
                cnt = 1;
                if ((st = p1_step(IC1{})) != 0) break;
                cnt = 2;
                if ((st = p1_step(IC2{})) != 0) break;
                cnt = 3;
                if ((st = p1_step(IC3{})) != 0) break;
                wip_rank_update<NB>(wb, pnb, fmask, 4, lane);
                __syncwarp();
            }

            // ---- phase 2 -------------------------------------------------------------------
            if (st == 1 && kAlgo == kSE99 && kstart == n - 1) {   // se99.rs:115-119
                const int u = __ffs(__ballot_sync(kFull, !fin)) - 1;
                const double ljj = __shfl_sync(kFull, dg, u);
                const double ej = tail1x1_e(ljj, gamma);
                if (lane == u) {
                    ev = ej;
                    sts_f64(tib + 8u * (unsigned)lane, dsqrt(dadd(ljj, ej)));
                }
            } else if (st == 1) {
                if (j > 0) {
                    // the Gershgorin bounds need the trailing block as the reference holds it: apply
                    // the pending columns, then g_i = a_ii - sum_{k alive, k != i} |a_ik|
                    // (se90.rs:105-110, se99.rs:125-130)
                    if (cnt > 0) {
                        wip_rank_update<NB>(wb, pnb, fmask, cnt, lane);
                        __syncwarp();
                        cnt = 0;
                    }
                    double s = 0.0;
                    unsigned tk = wb;
                    for (int k = 0; k < n; ++k) {
                        const bool alive = ((fmask >> k) & 1u) == 0u;
                        const double v = lds_f64(lane > k ? tib + 8u * (unsigned)k : tk + 8u * (unsigned)lane);
                        if (alive && k != lane)
                            s = dadd(s, fabs(v));
                        tk += 8u * (unsigned)(k + 1);
                    }
                    g = dsub(dg, s);
                }
                auto p2_step = [&](auto ctag) {
                    int khi;   // se90.rs:115, se99.rs:135
                    unsigned klo;
                    wip_key(g, khi, klo);
                    const int r = wip_pivot_lane(khi, klo, !fin && g == g, fin, pos, j);
                    take_position(r);
                    double piv = __shfl_sync(kFull, dg, r);
                    unsigned addr;
                    const double col = column(ctag, r, addr);
                    const bool live = !fin && lane != r;
                    const double acol = live ? abs_bits(col) : 0.0;
                    // E_jj (se90.rs:124-131, se99.rs:144-151)
                    const double normj = warp_sum_tree<32>(kFull, acol);
                    const double ej = max_b_not_nan(dadd(-piv, max_b_not_nan(normj, taugamma)), delta_prev);
                    piv = dadd(piv, ej);
                    delta_prev = ej;
                    double sq, y, inv_piv;
                    const bool fast = scale(piv, sq, y, inv_piv);
                    // bound update (se90.rs:134-139, se99.rs:154-159)
                    const bool upd = fabs(dsub(piv, normj)) > MODCHOL_EPS;
                    double t = 0.0;
                    if (upd)
                        t = fast ? __fma_rn(-normj, inv_piv, 1.0) : dsub(1.0, ddiv(normj, piv));
                    const double cs = fast ? dmul(col, y) : ddiv(col, sq);
                    if (live) {
                        dg = __fma_rn(-cs, cs, dg);
                        if (upd)
                            g = __fma_rn(acol, t, g);
                    }
                    if (lane == r)
                        ev = ej;
                    retire(ctag, r, addr, cs, sq);
                };
                while (j + 2 < n) {
                    p2_step(IC0{});
                    cnt = 1;
                    if (j + 2 >= n) break;
                    p2_step(IC1{});
                    cnt = 2;
                    if (j + 2 >= n) break;
                    p2_step(IC2{});
                    cnt = 3;
                    if (j + 2 >= n) break;
                    p2_step(IC3{});
                    cnt = 0;
                    wip_rank_update<NB>(wb, pnb, fmask, 4, lane);
                    __syncwarp();
                }
                // final 2x2 block, never pivoted (se90.rs:155-166, se99.rs:175-186): the two rows
                // alive sit at positions n-2 and n-1
                const int u0 = __ffs(__ballot_sync(kFull, !fin && pos == n - 2)) - 1;
                const int u1 = __ffs(__ballot_sync(kFull, !fin && pos == n - 1)) - 1;
                double a00 = __shfl_sync(kFull, dg, u0);
                double a11 = __shfl_sync(kFull, dg, u1);
                unsigned addr;
                const double c2 = column_rt(cnt, u0, addr);
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
                if (lane == u0) {
                    ev = en;
                    sts_f64(tib + 8u * (unsigned)lane, a00);
                }
                if (lane == u1) {
                    ev = en;
                    sts_f64(addr, a10);
                    sts_f64(tib + 8u * (unsigned)lane, a11);
                }
            }
        }

        // ---- epilogue: L row q = the slots of row inv[q] in pivot order (strict upper
        // explicitly zero, se99.rs:189-192), e in original order (se99.rs:194-198), p ---------
        if (has)
            sts_u32(invb + 4u * (unsigned)pos, (unsigned)lane);
        __syncwarp();
        const int ct = has ? (int)lds_u32(invb + 4u * (unsigned)lane) : 0;   // row at position `lane`
        const unsigned tct = wb + 8u * (unsigned)tri(ct);
        if (L) {
            double* dstp = L + (size_t)b * nn + lane;
#pragma unroll 4
            for (int q = 0; q < n; ++q) {
                const int iq = __shfl_sync(kFull, ct, q);
                const unsigned tq = __shfl_sync(kFull, tct, q);
                const unsigned a = iq > ct ? tq + 8u * (unsigned)ct : tct + 8u * (unsigned)iq;
                const double v = lds_f64_or_zero(lane <= q, a);
                if (has)
                    *dstp = v;
                dstp += n;
            }
        }
        if (has) {
            if (E)
                E[(size_t)b * n + lane] = ev;
            if (P)
                P[(size_t)b * n + lane] = (unsigned long long)ct;
        }
        if (K && lane == 0)
            K[b] = kstart;
    }
}

}  // namespace modchol
