// wip2_kernels.cuh -- the implicit-pivoting FAST kernels of wip_kernels.cuh for 32 < n <= 64:
// one warp per matrix, lane l owns the ORIGINAL rows l and l + 32 for the whole factorisation
// (gmw81.rs:39-134, se90.rs:38-181, se99.rs:35-201; no swap, see wip_kernels.cuh).
//
// Same data flow as the one-row kernels:
//   * C (symmetric trailing matrix, original coordinates) as a packed lower triangle in shared
//     memory, T(64) = 2 080 doubles per warp, brought up to date every kWipKB steps by
//     mma.sync.m8n8k4.f64 on its 36 tiles (8x8), fragments from the panel of pending columns;
//   * a step reads one slot per owned row (lane i > r: T(i) + r, else T(r) + i) and subtracts the
//     pending panel columns (registers x one broadcast LDS each);
//   * L goes to TENSOR MEMORY, lane l's two rows at columns [0, 128) and [128, 256) of its TMEM
//     lane (tcgen05.st per step), read back in the epilogue (tcgen05.ld) and written to rows
//     pos of the output with 256-bit stores; a finished row stores exact zeros;
//   * pivot search over two candidates per lane: lane-local choice, then the REDUX search of
//     wip_pivot_lane on (hi, lo, pos).
// Shared memory 21 KB and 256 tensor-memory columns per matrix: two 4-warp blocks per SM.  The
// wsm_se.cuh kernels this replaces in FAST mode spend ~400 instructions per step on a
// left-looking column of up to 63 terms; here a step is ~190 (profiles/r02_notes.md).
#pragma once
#include "wip_kernels.cuh"

namespace modchol {

constexpr int kWip2Tri = 2080;       // T(64)
constexpr int kWip2Ld = 68;          // panel column stride: >= 64, == 4 (mod 16)
constexpr int kWip2Doubles = kWip2Tri + kWipKB * kWip2Ld;
constexpr int kWip2TmemCols = 256;   // two rows of 64 doubles per lane

// kSolve: fused consumer step, as in wip_kernel (two positions per lane in the substitution).
template <int kAlgo, int NB, int kN, int kMinBlocks, bool kSolve = false>
__global__ void __launch_bounds__(kWipWarps * 32, kMinBlocks)
wip2_kernel(int n_rt, long long batch, const double* __restrict__ A, double* __restrict__ L,
            double* __restrict__ E, unsigned long long* __restrict__ P, int* __restrict__ K,
            const double* Bm = nullptr, double* X = nullptr, int nrhs = 0)
{
    extern __shared__ double sm[];
    __shared__ unsigned tmem_base;
    constexpr int KC = kWipKB / 4;
    const int n = kN ? kN : n_rt;
    int lane = threadIdx.x & 31;
    asm volatile("mov.u32 %0, %0;" : "+r"(lane));
    const int wib = __shfl_sync(kFull, (int)(threadIdx.x >> 5), 0);

    if (wib == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"((unsigned)__cvta_generic_to_shared(&tmem_base)), "n"(kWip2TmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tw = tmem_base + ((unsigned)(wib * 32) << 16);

    unsigned wb = (unsigned)__cvta_generic_to_shared(sm + (size_t)wib * kWip2Doubles);
    asm volatile("mov.u32 %0, %0;" : "+r"(wb));
    const unsigned pnb = wb + 8u * kWip2Tri;
    const unsigned tib0 = wb + 8u * (unsigned)tri(lane);          // slot (lane, 0)
    const unsigned tib1 = wb + 8u * (unsigned)tri(lane + 32);     // slot (lane + 32, 0)
    const unsigned pcol = pnb + 8u * (unsigned)lane;              // this lane's first entry of panel column 0
    const int r8 = lane >> 2, c4 = lane & 3;
    const unsigned pf = pnb + 8u * (unsigned)(kWip2Ld * c4 + r8);  // fragment source
    const unsigned fr0 = wb + 8u * (unsigned)(tri(r8) + 2 * c4);   // slot (r8, 2 c4)
    const bool pd0 = 2 * c4 <= r8, pd1 = 2 * c4 + 1 <= r8;
    const size_t nn = (size_t)n * n;
    const long long nslots = (long long)gridDim.x * kWipWarps;
    const bool has0 = lane < n, has1 = lane + 32 < n;
    const bool wide = kN != 0 && kN % 8 == 0 && (((size_t)A | (size_t)L) & 31) == 0;

    for (long long b = (long long)blockIdx.x * kWipWarps + wib; b < batch; b += nslots) {
        const double* Ab = A + (size_t)b * nn;
        // ---- load: row by row into the packed triangle (lane = column, two columns per lane);
        // SE: absolute off-diagonal column sums (== row sums, A is symmetric); GMW81: off-diagonal
        // maximum (gmw81.rs:52-58) as integers -----------------------------------------------------
        double dg0 = has0 ? __ldg(Ab + (size_t)lane * (n + 1)) : 0.0;
        double dg1 = has1 ? __ldg(Ab + (size_t)(lane + 32) * (n + 1)) : 0.0;
        double as0 = 0.0, as1 = 0.0;
        unsigned long long om = 0ull;
        __syncwarp();   // the previous matrix's readers of the triangle are done
        {
            auto take = [&](int i, double v0, double v1) {
                const unsigned ti = wb + 4u * (unsigned)(i * (i + 1));   // 8 T(i)
                sts_f64_if(lane <= i, ti + 8u * (unsigned)lane, v0);
                sts_f64_if(lane + 32 <= i, ti + 8u * (unsigned)(lane + 32), v1);
                if constexpr (kAlgo == kGMW81) {
                    const unsigned long long a0 = (unsigned long long)__double_as_longlong(v0) & 0x7fffffffffffffffull;
                    const unsigned long long a1 = (unsigned long long)__double_as_longlong(v1) & 0x7fffffffffffffffull;
                    om = (lane != i && a0 > om) ? a0 : om;
                    om = (lane + 32 != i && a1 > om) ? a1 : om;
                } else {
                    as0 = dadd(as0, fabs(v0));
                    as1 = dadd(as1, fabs(v1));
                }
            };
            const double* src = Ab + lane;
            int i = 0;
            for (; i + 4 <= n; i += 4) {
                double v0[4], v1[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    v0[u] = has0 ? __ldg(src + (size_t)u * n) : 0.0;
                    v1[u] = has1 ? __ldg(src + (size_t)u * n + 32) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    take(i + u, v0[u], v1[u]);
                src += (size_t)4 * n;
            }
            for (; i < n; ++i) {
                take(i, has0 ? __ldg(src) : 0.0, has1 ? __ldg(src + 32) : 0.0);
                src += n;
            }
            as0 = dsub(as0, fabs(dg0));
            as1 = dsub(as1, fabs(dg1));
        }
        __syncwarp();

        // C(i,k) -= sum_{c < cnt} P(i,c) P(k,c), tile by tile (see wip_kernels.cuh)
        auto rank_update = [&](int cnt) {
            // per block row: every load first, then the tensor-pipe operations, then the stores -- the
            // accessors are volatile asm (kept in program order), so a tile-by-tile loop would wait
            // for a shared-memory round trip in front of every DMMA
            double f[KC][NB];
#pragma unroll
            for (int h = 0; h < KC; ++h)
#pragma unroll
                for (int q = 0; q < NB; ++q) {
                    const unsigned a = pf + 8u * (unsigned)(4 * kWip2Ld * h) + 64u * q;
                    f[h][q] = cnt == kWipKB ? lds_f64(a) : lds_f64_or_zero(4 * h + c4 < cnt, a);
                }
#pragma unroll
            for (int bi = 0; bi < NB; ++bi) {
                double nf[KC], c0[NB], c1[NB];
                const unsigned ra = fr0 + 8u * (unsigned)(tri(8 * bi) + 8 * bi * r8);
#pragma unroll
                for (int h = 0; h < KC; ++h)
                    nf[h] = wip_neg(f[h][bi]);
#pragma unroll
                for (int bj = 0; bj <= bi; ++bj) {
                    c0[bj] = lds_f64(ra + 64u * bj);
                    c1[bj] = lds_f64(ra + 64u * bj + 8);
                }
#pragma unroll
                for (int bj = 0; bj <= bi; ++bj)
#pragma unroll
                    for (int h = 0; h < KC; ++h)
                        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                                     : "+d"(c0[bj]), "+d"(c1[bj]) : "d"(nf[h]), "d"(f[h][bj]));
#pragma unroll
                for (int bj = 0; bj <= bi; ++bj) {
                    if (bi == bj) {
                        sts_f64_if(pd0, ra + 64u * bj, c0[bj]);
                        sts_f64_if(pd1, ra + 64u * bj + 8, c1[bj]);
                    } else {
                        sts_f64(ra + 64u * bj, c0[bj]);
                        sts_f64(ra + 64u * bj + 8, c1[bj]);
                    }
                }
            }
            __syncwarp();
        };

        double ev0 = 0.0, ev1 = 0.0;
        int pos0 = lane, pos1 = lane + 32;
        bool fin0 = !has0, fin1 = !has1;
        unsigned long long fmask = n >= 64 ? 0ull : (~0ull << n);
        int j = 0, cnt = 0, kstart = -1;
        double cs0r[kWipKB - 1], cs1r[kWipKB - 1];   // these rows' entries of the pending panel columns
#pragma unroll
        for (int c = 0; c < kWipKB - 1; ++c)
            cs0r[c] = cs1r[c] = 0.0;
        bool win;

        // (lane, slot) of the row that holds the first maximum, in permuted-position order
        // (utils.rs:52-91), of the keyed values over the candidate rows; NaN rows are not
        // candidates; without a candidate the row at position j stays
        auto pivot_row = [&](int kh0, unsigned kl0, bool c0, int kh1, unsigned kl1, bool c1, int& rl, int& rs) {
            const bool pick1 = c1 && (!c0 || kh1 > kh0 || (kh1 == kh0 && (kl1 > kl0 || (kl1 == kl0 && pos1 < pos0))));
            const int kh = pick1 ? kh1 : kh0;
            const unsigned kl = pick1 ? kl1 : kl0;
            const int ps = pick1 ? pos1 : pos0;
            const bool cand = c0 || c1;
            const int mh = __reduce_max_sync(kFull, cand ? kh : (int)0x80000000);
            const bool a1 = cand && kh == mh;
            unsigned b1 = __ballot_sync(kFull, a1);
            int slot = pick1 ? 1 : 0;
            if (__popc(b1) != 1) {
                const unsigned ml = __reduce_max_sync(kFull, a1 ? kl : 0u);
                const bool w = a1 && kl == ml;
                unsigned mpos = __reduce_min_sync(kFull, w ? (unsigned)ps : 1024u);
                if (mpos > 255u)
                    mpos = (unsigned)j;
                const bool h0 = !fin0 && pos0 == (int)mpos, h1 = !fin1 && pos1 == (int)mpos;
                b1 = __ballot_sync(kFull, h0 || h1);
                slot = h1 ? 1 : 0;
            }
            rl = 31 - __clz(b1);
            rs = __shfl_sync(kFull, slot, rl);
        };
        auto take_position = [&](int rl, int rs) {
            const int pr = __shfl_sync(kFull, rs ? pos1 : pos0, rl);
            pos0 = pos0 == j ? pr : pos0;
            pos1 = pos1 == j ? pr : pos1;
            if (lane == rl) {
                if (rs) pos1 = j; else pos0 = j;
            }
        };
        // current values of C(lane, r), C(lane + 32, r): stored slots minus the C pending columns
        auto column = [&](int C, int r, double& col0, double& col1) {
            const unsigned tr = wb + 4u * (unsigned)(r * (r + 1));
            col0 = lds_f64(lane > r ? tib0 + 8u * (unsigned)r : tr + 8u * (unsigned)lane);
            col1 = lds_f64(lane + 32 > r ? tib1 + 8u * (unsigned)r : tr + 8u * (unsigned)(lane + 32));
            const unsigned pr = pnb + 8u * (unsigned)r;
            double e0 = 0.0, e1 = 0.0;   // odd pending columns apart: two chains of half the length
            auto corr = [&](auto ctag) {
                constexpr int c = decltype(ctag)::value;
                if constexpr (c < kWipKB - 1) {
                    const double p = lds_f64(pr + 8u * (unsigned)(kWip2Ld * c));
                    if constexpr (c & 1) {
                        e0 = __fma_rn(-cs0r[c], p, e0);
                        e1 = __fma_rn(-cs1r[c], p, e1);
                    } else {
                        col0 = __fma_rn(-cs0r[c], p, col0);
                        col1 = __fma_rn(-cs1r[c], p, col1);
                    }
                }
            };
            if (C >= 4) {
                corr(IC<0>{}); corr(IC<1>{}); corr(IC<2>{}); corr(IC<3>{});
                if (C >= 6) {
                    corr(IC<4>{}); corr(IC<5>{});
                    if (C >= 7)
                        corr(IC<6>{});
                } else if (C == 5) {
                    corr(IC<4>{});
                }
            } else if (C >= 2) {
                corr(IC<0>{}); corr(IC<1>{});
                if (C == 3)
                    corr(IC<2>{});
            } else if (C == 1) {
                corr(IC<0>{});
            }
            col0 = dadd(col0, e0);
            col1 = dadd(col1, e1);
        };
        // column j of L: cs (sq for the pivot row) into tensor memory and panel column C
        auto retire = [&](int C, int rl, int rs, double cs0, double cs1, double sq) {
            const bool me = lane == rl;
            wip_tmem_store(tw + 2u * (unsigned)j, (me && rs == 0) ? sq : cs0);
            wip_tmem_store(tw + 128u + 2u * (unsigned)j, (me && rs == 1) ? sq : cs1);
            const unsigned pc = pcol + 8u * (unsigned)(kWip2Ld * C);
            sts_f64(pc, cs0);
            sts_f64(pc + 256u, cs1);
            auto keep = [&](auto ctag) {
                constexpr int c = decltype(ctag)::value;
                if constexpr (c < kWipKB - 1) {
                    cs0r[c] = cs0;
                    cs1r[c] = cs1;
                }
            };
            if (C >= 4) {
                if (C >= 6) {
                    if (C == 6) keep(IC<6>{});
                } else if (C == 5) {
                    keep(IC<5>{});
                } else {
                    keep(IC<4>{});
                }
            } else if (C >= 2) {
                if (C == 3) keep(IC<3>{}); else keep(IC<2>{});
            } else if (C == 1) {
                keep(IC<1>{});
            } else {
                keep(IC<0>{});
            }
            fin0 = fin0 || (me && rs == 0);
            fin1 = fin1 || (me && rs == 1);
            fmask |= 1ull << (rl + 32 * rs);
            __syncwarp();
            ++j;
        };
        auto next_count = [&]() {
            if (++cnt == kWipKB) {
                rank_update(kWipKB);
                cnt = 0;
            }
        };

        if constexpr (kAlgo == kGMW81) {
            // ---- GMW81, symmetric form (see wsm_gmw.cuh) --------------------------------------------
            const double dmx = fmax(has0 ? fabs(dg0) : 0.0, has1 ? fabs(dg1) : 0.0);
            const double diag_max = warp_extreme<true>(kFull, dmx, true, 0.0, win);
            const unsigned ohi = __reduce_max_sync(kFull, (unsigned)(om >> 32));
            const unsigned olo = __reduce_max_sync(kFull, (unsigned)(om >> 32) == ohi ? (unsigned)om : 0u);
            const double off_max = __hiloint2double((int)ohi, (int)olo);
            const double nf = (double)n;
            const double delta = dmul(MODCHOL_EPS, fmax(1.0, dadd(diag_max, off_max)));
            const double beta =
                dsqrt(fmax(fmax(diag_max, ddiv(off_max, dsqrt(dsub(dmul(nf, nf), 1.0)))), MODCHOL_EPS));
            const double rbeta = ddiv(1.0, beta);
            while (j < n) {
                // pivot on max |c_ii| (gmw81.rs:71-78)
                int rl, rs;
                pivot_row(__double2hiint(dg0) & 0x7fffffff, (unsigned)__double2loint(dg0), !fin0 && dg0 == dg0,
                          __double2hiint(dg1) & 0x7fffffff, (unsigned)__double2loint(dg1), !fin1 && dg1 == dg1, rl, rs);
                take_position(rl, rs);
                const int r = rl + 32 * rs;
                double col0, col1;
                column(cnt, r, col0, col1);
                const double cjj = __shfl_sync(kFull, rs ? col1 : col0, rl);   // c_jj from the same updates
                const bool me = lane == rl;
                const bool live0 = !fin0 && !(me && rs == 0), live1 = !fin1 && !(me && rs == 1);
                // theta = max_{i>j} |c_ij| (gmw81.rs:87-100), d_j (gmw81.rs:102)
                const double th = fmax(live0 ? fabs(col0) : 0.0, live1 ? fabs(col1) : 0.0);
                const double theta = warp_extreme<true>(kFull, th, true, 0.0, win);
                const double tb = dmul(theta, rbeta);
                const double dj = fmax(fmax(fabs(cjj), dmul(tb, tb)), delta);
                double y = rsqrt_normal(dj);
                double sq = dmul(dj, y);
                if (!fast_range_ok(dj))
                    wip_slow_gmw(dj, sq, y);
                const double cs0 = fin0 ? 0.0 : dmul(col0, y), cs1 = fin1 ? 0.0 : dmul(col1, y);
                dg0 = __fma_rn(-cs0, cs0, dg0);                      // gmw81.rs:104-109
                dg1 = __fma_rn(-cs1, cs1, dg1);
                const double ej = dsub(dj, cjj);                     // gmw81.rs:110
                ev0 = (me && rs == 0) ? ej : ev0;
                ev1 = (me && rs == 1) ? ej : ev1;
                retire(cnt, rl, rs, cs0, cs1, sq);
                if (j < n)
                    next_count();
            }
        } else {
            double g0 = dsub(dg0, as0), g1 = dsub(dg1, as1);   // Gershgorin bounds if phase 2 starts at column 0
            // gamma = max |a_ii| folded from 0.0 (se90.rs:55-57, se99.rs:54-56)
            const double dmx = fmax(has0 ? fabs(dg0) : 0.0, has1 ? fabs(dg1) : 0.0);
            const double gamma = warp_extreme<true>(kFull, dmx, true, 0.0, win);
            const double taugamma = dmul(MODCHOL_TAU, gamma);
            double delta_prev = 0.0;
            bool have_dmin = false;
            double dmin_carry = 0.0;
            int st = 0;

            // ---- phase 1 (se90.rs:61-96, se99.rs:61-109): 0 go on, 1 switch to phase 2, 2 done ----
            for (;;) {
                int kh0, kh1, rl, rs;
                unsigned kl0, kl1;
                wip_key(dg0, kh0, kl0);
                wip_key(dg1, kh1, kl1);
                pivot_row(kh0, kl0, !fin0 && dg0 == dg0, kh1, kl1, !fin1 && dg1 == dg1, rl, rs);
                const double piv = __shfl_sync(kFull, rs ? dg1 : dg0, rl);
                if (kAlgo == kSE99) {   // se99.rs:62-73, before the pivot is applied
                    double dmin = dmin_carry;
                    if (!have_dmin) {
                        const double lm = fmin(fin0 ? INFINITY : dg0, fin1 ? INFINITY : dg1);
                        dmin = warp_extreme<false>(kFull, lm, true, INFINITY, win);
                    }
                    const double dmax = (piv == piv) ? piv : -INFINITY;
                    if (dmax < taugamma || dmin < dmul(-MODCHOL_MU, dmax)) {
                        kstart = j;
                        st = 1;
                        break;
                    }
                }
                take_position(rl, rs);
                const int r = rl + 32 * rs;
                double col0, col1;
                column(cnt, r, col0, col1);
                const bool me = lane == rl;
                const bool live0 = !fin0 && !(me && rs == 0), live1 = !fin1 && !(me && rs == 1);
                const bool fast = fast_range_ok(piv);
                const double y = rsqrt_normal(piv);
                double sq = dmul(piv, y);
                double cs0 = dmul(col0, y), cs1 = dmul(col1, y);
                double dn0 = __fma_rn(-cs0, cs0, dg0), dn1 = __fma_rn(-cs1, cs1, dg1);
                double nv0 = dn0, nv1 = dn1;   // look-ahead values (se90.rs:70-77, se99.rs:82-89)
                if (!fast) {
                    double sq1;
                    wip_slow_p1(piv, col0, dg0, sq, cs0, dn0, nv0);
                    wip_slow_p1(piv, col1, dg1, sq1, cs1, dn1, nv1);
                }
                const double lm = fmin(live0 ? nv0 : INFINITY, live1 ? nv1 : INFINITY);
                const double la = warp_extreme<false>(kFull, lm, true, INFINITY, win);
                have_dmin = fast;
                dmin_carry = la;
                const double thresh = (kAlgo == kSE99) ? dmul(-MODCHOL_MU, gamma) : taugamma;
                if (la < thresh) {   // the pivot at j stays applied (se99.rs:90-93)
                    kstart = j;
                    st = 1;
                    break;
                }
                cs0 = fin0 ? 0.0 : cs0;
                cs1 = fin1 ? 0.0 : cs1;
                dg0 = dn0;
                dg1 = dn1;
                retire(cnt, rl, rs, cs0, cs1, sq);
                if (j == n) {
                    st = 2;
                    break;
                }
                next_count();
            }

            // ---- phase 2 -------------------------------------------------------------------------
            if (st == 1 && kAlgo == kSE99 && kstart == n - 1) {   // se99.rs:115-119
                const unsigned bb0 = __ballot_sync(kFull, !fin0), bb1 = __ballot_sync(kFull, !fin1);
                const int us = bb1 ? 1 : 0;
                const int ul = 31 - __clz(bb1 ? bb1 : bb0);
                const double ljj = __shfl_sync(kFull, us ? dg1 : dg0, ul);
                const double ej = tail1x1_e(ljj, gamma);
                const bool me = lane == ul;
                ev0 = (me && us == 0) ? ej : ev0;
                ev1 = (me && us == 1) ? ej : ev1;
                const double v = dsqrt(dadd(ljj, ej));
                wip_tmem_store(tw + 2u * (unsigned)(n - 1), (me && us == 0) ? v : 0.0);
                wip_tmem_store(tw + 128u + 2u * (unsigned)(n - 1), (me && us == 1) ? v : 0.0);
            } else if (st == 1) {
                if (j > 0) {
                    // the Gershgorin bounds need the trailing block as the reference holds it: apply
                    // the pending columns, then g_i = a_ii - sum_{k alive, k != i} |a_ik|
                    // (se90.rs:105-110, se99.rs:125-130)
                    if (cnt > 0) {
                        rank_update(cnt);
                        cnt = 0;
                    }
                    double s0 = 0.0, s1 = 0.0;
                    unsigned tk = wb;
                    for (int k = 0; k < n; ++k) {
                        const bool alive = ((fmask >> k) & 1ull) == 0ull;
                        const double v0 = lds_f64(lane > k ? tib0 + 8u * (unsigned)k : tk + 8u * (unsigned)lane);
                        const double v1 = lds_f64(lane + 32 > k ? tib1 + 8u * (unsigned)k : tk + 8u * (unsigned)(lane + 32));
                        if (alive && k != lane)
                            s0 = dadd(s0, fabs(v0));
                        if (alive && k != lane + 32)
                            s1 = dadd(s1, fabs(v1));
                        tk += 8u * (unsigned)(k + 1);
                    }
                    g0 = dsub(dg0, s0);
                    g1 = dsub(dg1, s1);
                }
                while (j + 2 < n) {
                    int kh0, kh1, rl, rs;   // se90.rs:115, se99.rs:135
                    unsigned kl0, kl1;
                    wip_key(g0, kh0, kl0);
                    wip_key(g1, kh1, kl1);
                    pivot_row(kh0, kl0, !fin0 && g0 == g0, kh1, kl1, !fin1 && g1 == g1, rl, rs);
                    take_position(rl, rs);
                    const int r = rl + 32 * rs;
                    double piv = __shfl_sync(kFull, rs ? dg1 : dg0, rl);
                    double col0, col1;
                    column(cnt, r, col0, col1);
                    const bool me = lane == rl;
                    const bool live0 = !fin0 && !(me && rs == 0), live1 = !fin1 && !(me && rs == 1);
                    const double ac0 = live0 ? abs_bits(col0) : 0.0, ac1 = live1 ? abs_bits(col1) : 0.0;
                    // E_jj (se90.rs:124-131, se99.rs:144-151)
                    const double normj = wip_sum32(dadd(ac0, ac1));
                    const double ej = max_b_not_nan(dadd(-piv, max_b_not_nan(normj, taugamma)), delta_prev);
                    piv = dadd(piv, ej);
                    delta_prev = ej;
                    // bound update (se90.rs:134-139, se99.rs:154-159) and column scaling
                    const bool upd = fabs(dsub(piv, normj)) > MODCHOL_EPS;
                    const double y = rsqrt_normal(piv);
                    double sq = dmul(piv, y);
                    double t = __fma_rn(-normj, dmul(y, y), 1.0);
                    double cs0 = dmul(col0, y), cs1 = dmul(col1, y);
                    if (!fast_range_ok(piv)) {
                        double sq1, t1;
                        wip_slow_p2(piv, normj, col0, sq, t, cs0);
                        wip_slow_p2(piv, normj, col1, sq1, t1, cs1);
                    }
                    cs0 = fin0 ? 0.0 : cs0;   // finished rows receive exact zeros
                    cs1 = fin1 ? 0.0 : cs1;
                    dg0 = __fma_rn(-cs0, cs0, dg0);
                    dg1 = __fma_rn(-cs1, cs1, dg1);
                    g0 = upd ? __fma_rn(ac0, t, g0) : g0;
                    g1 = upd ? __fma_rn(ac1, t, g1) : g1;
                    ev0 = (me && rs == 0) ? ej : ev0;
                    ev1 = (me && rs == 1) ? ej : ev1;
                    retire(cnt, rl, rs, cs0, cs1, sq);
                    next_count();
                }
                // final 2x2 block, never pivoted (se90.rs:155-166, se99.rs:175-186): the two rows
                // alive sit at positions n-2 and n-1
                const bool q00 = !fin0 && pos0 == n - 2, q01 = !fin1 && pos1 == n - 2;
                const bool q10 = !fin0 && pos0 == n - 1, q11 = !fin1 && pos1 == n - 1;
                const unsigned bq0 = __ballot_sync(kFull, q00 || q01), bq1 = __ballot_sync(kFull, q10 || q11);
                const int u0l = 31 - __clz(bq0), u1l = 31 - __clz(bq1);
                const int u0s = __shfl_sync(kFull, q01 ? 1 : 0, u0l), u1s = __shfl_sync(kFull, q11 ? 1 : 0, u1l);
                double a00 = __shfl_sync(kFull, u0s ? dg1 : dg0, u0l);
                double a11 = __shfl_sync(kFull, u1s ? dg1 : dg0, u1l);
                double c20, c21;
                column(cnt, u0l + 32 * u0s, c20, c21);
                double a10 = __shfl_sync(kFull, u1s ? c21 : c20, u1l);
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
                const bool m00 = lane == u0l && u0s == 0, m01 = lane == u0l && u0s == 1;
                const bool m10 = lane == u1l && u1s == 0, m11 = lane == u1l && u1s == 1;
                ev0 = (m00 || m10) ? en : ev0;
                ev1 = (m01 || m11) ? en : ev1;
                wip_tmem_store(tw + 2u * (unsigned)(n - 2), m00 ? a00 : (m10 ? a10 : 0.0));
                wip_tmem_store(tw + 128u + 2u * (unsigned)(n - 2), m01 ? a00 : (m11 ? a10 : 0.0));
                wip_tmem_store(tw + 2u * (unsigned)(n - 1), m10 ? a11 : 0.0);
                wip_tmem_store(tw + 128u + 2u * (unsigned)(n - 1), m11 ? a11 : 0.0);
            }
        }

        // ---- epilogue: the two rows of L leave tensor memory for rows pos0 / pos1 of the output;
        // e in original order (se99.rs:194-198); p -------------------------------------------------
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        if constexpr (kSolve)
            __syncwarp();   // every lane is done reading C from the triangle
        if (L || kSolve) {
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const bool has = s ? has1 : has0;
                const int ps = s ? pos1 : pos0;
                double* dstp = L + (size_t)b * nn + (size_t)ps * n;
                const unsigned rowb = wb + 8u * (unsigned)tri(ps);   // row ps of the position-ordered triangle
#pragma unroll
                for (int c = 0; c < NB; ++c) {
                    if (8 * c >= n)
                        break;
                    double v[8];
                    wip_tmem_load8(tw + 128u * s + 16u * c, v);
                    if (L && wide && has) {
                        stg_v4f64(dstp + 8 * c, v[0], v[1], v[2], v[3]);
                        stg_v4f64(dstp + 8 * c + 4, v[4], v[5], v[6], v[7]);
                    } else if (L && !wide && has) {
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            if (8 * c + u < n)
                                dstp[8 * c + u] = v[u];
                    }
                    if constexpr (kSolve) {
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            sts_f64_if(has && 8 * c + u <= ps, rowb + 8u * (unsigned)(8 * c + u), v[u]);
                    }
                }
            }
        }
        if (has0) {
            if (E) E[(size_t)b * n + lane] = ev0;
            if (P) P[(size_t)b * n + pos0] = (unsigned long long)lane;
        }
        if (has1) {
            if (E) E[(size_t)b * n + lane + 32] = ev1;
            if (P) P[(size_t)b * n + pos1] = (unsigned long long)(lane + 32);
        }
        if (K && lane == 0)
            K[b] = kstart;
        if constexpr (kSolve) {
            // lane -> positions lane, lane + 32 through the (now free) panel area, then the
            // substitutions on the position-ordered triangle
            sts_f64(pnb + 8u * (unsigned)pos0, (double)lane);
            sts_f64(pnb + 8u * (unsigned)pos1, (double)(lane + 32));
            __syncwarp();
            const unsigned tibs[2] = {tib0, tib1};
            const int rows[2] = {lane, lane + 32};
            const bool rvs[2] = {has0, has1};
            const int perms[2] = {(int)lds_f64(pnb + 8u * (unsigned)lane), (int)lds_f64(pnb + 8u * (unsigned)(lane + 32))};
            wsm_solve_in_place<32, 2>(kFull, n, wb, tibs, rows, rvs, perms, Bm + (size_t)b * nrhs * n,
                                      X + (size_t)b * nrhs * n, nrhs);
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (wib == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                     :: "r"(tmem_base), "n"(kWip2TmemCols) : "memory");
}

}  // namespace modchol
