// lwm_kernels.cuh -- FAST-mode GMW81 / SE90 / SE99 for the large size class up to n = 256:
// ONE WARP per matrix (lane l owns the permuted positions l, l + 32, ...), physical pivoting,
// column panels in shared memory, the trailing matrix streamed through shared memory by the
// bulk-copy engine (TMA: cp.async.bulk + mbarrier), updated there on the FP64 tensor pipe.
//
// Why one warp.  A pivot step is a chain of dependent reductions (pivot search -> column ->
// theta / norm / look-ahead -> scaling).  With a thread block per matrix every link costs a block
// barrier and each of the 8 warps repeats the scalar work: 2 170 warp-instructions and ~6 700
// cycles per step at n = 256 (profiles/r02_notes.md).  One warp does the same step in ~300
// instructions with warp-level reductions only; throughput then comes from the NUMBER of matrices
// in flight (three per SM, 444 per chip -- the most whose working triangles still fit the L2).
//
//   storage   the working matrix lives in place in the caller's output buffer: the strictly lower
//             trailing part T(i,k), i > k, TRANSPOSED in the strict upper triangle l[k n + i] (the
//             symmetric input delivers it as is), finished L in the lower triangle (final place).
//             Per-position scalars (running diagonal, Gershgorin bounds, e, p) in shared memory.
//   step j    pivot search over the scalars; symmetric exchange j <-> m fused with the read of
//             column j (one round trip to L2); minus the pending panel columns (left-looking inside
//             the panel); E rule; scaled column appended to the panel (gmw81.rs:70-111,
//             se90.rs:61-96 / 113-152, se99.rs:61-109 / 133-172).
//   panel     every KB columns (or at the phase switch): the panel's entries of L are written out
//             and the trailing block is updated, T -= P P^T: block rows of 8 storage rows are brought
//             into a two-stage shared-memory ring by cp.async.bulk (one copy per row, completion on
//             an mbarrier), updated by mma.sync.m8n8k4.f64 (KB/4 per 8x8 tile, fragments from the
//             panel), and written back by cp.async.bulk shared -> global while the next block row
//             is already in flight.  The same ring copies A into the working buffer at the start
//             (GMW81: with the off-diagonal maximum of gmw81.rs:52-58 taken on the way), feeds the
//             Gershgorin row sums at the phase switch (se90.rs:105-110, se99.rs:125-130) and a
//             zero line is bulk-stored over the strict upper triangle at the end.
//   odd n / unaligned buffers take the same ring with ordinary loads and stores.
#pragma once
#include "bsm_kernels.cuh"

namespace modchol {

constexpr int kLwmChunk = 128;      // columns of a staged block row (8 rows x 128 doubles = 8 KB)
constexpr int kLwmStageLd = 136;    // row stride of a stage: 1 088 B == 64 (mod 128): LDS.128 conflict free

__host__ __device__ inline size_t lwm_smem_bytes(int n, int kb)
{
    const size_t npad = (size_t)((n + 7) & ~7);
    return 8 * ((size_t)kb * bsm_ldp(n) + 5 * npad + 2 * 8 * kLwmStageLd + 2) + 4 * npad;
}

__device__ __forceinline__ void sts_v2f64(unsigned a, double x, double y)
{
    smem_check(a, 16u);
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" :: "r"(a), "d"(x), "d"(y) : "memory");
}

struct LwmRing {
    unsigned stb;      // two stages of 8 x kLwmStageLd doubles
    unsigned mbar;     // two mbarriers
    unsigned ph[2];    // their phase bits
    bool tma;          // bulk copies (16-byte aligned rows) or ordinary loads / stores
};

// Streams the block rows bj >= b0 of `src` (storage rows 8 bj .. 8 bj + 7, columns 8 bj .. n) through
// the ring in chunks of kLwmChunk columns: fn(bj, ib, ie, stage) sees the chunk [ib, ie) of the
// eight rows at stage + 8 (r kLwmStageLd + i - ib); with kStore the chunk then goes to `dst`.
template <bool kStore, typename Fn>
__device__ __forceinline__ void lwm_stream(LwmRing& rg, const double* __restrict__ src, double* __restrict__ dst,
                                           int n, int b0, int lane, Fn&& fn)
{
    const int nb = (n + 7) >> 3;
    if (b0 >= nb)
        return;
    constexpr unsigned kStageBytes = 8u * 8u * kLwmStageLd;
    auto load = [&](int s, int bj, int ib) {
        const int ie = min(n, ib + kLwmChunk), rows = min(8, n - 8 * bj);
        const unsigned st = rg.stb + (unsigned)s * kStageBytes;
        if (rg.tma) {
            const unsigned bytes = 8u * (unsigned)(ie - ib);
            if (lane == 0)
                mbar_expect_tx(rg.mbar + 8u * s, bytes * (unsigned)rows);
            __syncwarp();
            if (lane < rows)
                bulk_g2s(st + 8u * (unsigned)(lane * kLwmStageLd), src + (size_t)(8 * bj + lane) * n + ib, bytes,
                         rg.mbar + 8u * s);
        } else {
            for (int r = 0; r < rows; ++r)
                for (int i = ib + lane; i < ie; i += 32)
                    sts_f64(st + 8u * (unsigned)(r * kLwmStageLd + i - ib), __ldcg(src + (size_t)(8 * bj + r) * n + i));
            __syncwarp();
        }
    };
    if (rg.tma) {
        fence_proxy_async();   // this warp's ordinary global stores come before the bulk reads
        __syncwarp();
    } else {
        __syncwarp();
    }
    int bj = b0, ib = 8 * b0, s = 0;
    load(0, bj, ib);
    for (;;) {
        int nbj = bj, nib = ib + kLwmChunk;
        if (nib >= n) {
            nbj = bj + 1;
            nib = 8 * nbj;
        }
        const bool more = nbj < nb;
        if (more) {
            if (kStore && rg.tma) {   // the write-back that last used the other stage has read it
                if (lane < 8)
                    bulk_wait_read0();
                __syncwarp();
            }
            load(s ^ 1, nbj, nib);
        }
        if (rg.tma) {
            mbar_wait(rg.mbar + 8u * s, rg.ph[s]);
            rg.ph[s] ^= 1u;
        }
        const int ie = min(n, ib + kLwmChunk), rows = min(8, n - 8 * bj);
        const unsigned st = rg.stb + (unsigned)s * kStageBytes;
        fn(bj, ib, ie, st);
        if (kStore) {
            if (rg.tma) {
                fence_proxy_async();   // the updated stage (ordinary shared stores) before the bulk read
                __syncwarp();
                if (lane < rows) {
                    bulk_s2g(dst + (size_t)(8 * bj + lane) * n + ib, st + 8u * (unsigned)(lane * kLwmStageLd),
                             8u * (unsigned)(ie - ib));
                    bulk_commit();
                }
            } else {
                __syncwarp();
                for (int r = 0; r < rows; ++r)
                    for (int i = ib + lane; i < ie; i += 32)
                        dst[(size_t)(8 * bj + r) * n + i] = lds_f64(st + 8u * (unsigned)(r * kLwmStageLd + i - ib));
                __syncwarp();
            }
        } else {
            __syncwarp();   // every lane is done with the stage before it is loaded again
        }
        if (!more)
            break;
        bj = nbj;
        ib = nib;
        s ^= 1;
    }
    if (kStore && rg.tma) {
        if (lane < 8)
            bulk_wait0();      // written and visible
        __syncwarp();
    }
}

// warp-wide first maximum by position (utils.rs:52-91) of vals[i] (|vals[i]| with kAbs), i in
// [j, n): NaN never replaces the running maximum, a NaN at position j sticks, nothing to take -> j
template <int R, bool kAbs>
__device__ __forceinline__ int lwm_argmax(unsigned valb, int j, int n, int lane)
{
    double bv = 0.0;
    int bi = -1;
#pragma unroll
    for (int q = 0; q < R; ++q) {
        const int i = lane + 32 * q;
        if (i >= j && i < n) {
            double v = lds_f64(valb + 8u * (unsigned)i);
            if (kAbs)
                v = fabs(v);
            if (v == v && (bi < 0 || v > bv)) {
                bv = v;
                bi = i;
            }
        }
    }
    int khi;
    unsigned klo;
    wip_key(bv, khi, klo);
    const bool cand = bi >= 0;
    const int mh = __reduce_max_sync(kFull, cand ? khi : (int)0x80000000);
    const bool c1 = cand && khi == mh;
    const unsigned ml = __reduce_max_sync(kFull, c1 ? klo : 0u);
    const unsigned mp = __reduce_min_sync(kFull, (c1 && klo == ml) ? (unsigned)bi : 0xffffu);
    const double head = lds_f64(valb + 8u * (unsigned)j);
    return (mh == (int)0x80000000 || mp >= 0xffffu || head != head) ? j : (int)mp;
}

template <int kAlgo, int R, int KB>
__global__ void __launch_bounds__(32)
lwm_kernel(int n, long long batch, const double* __restrict__ A, double* __restrict__ L,
           double* __restrict__ E, unsigned long long* __restrict__ P, int* __restrict__ K)
{
    extern __shared__ __align__(16) double sm[];
    constexpr int KC = KB / 4;
    const int lane = threadIdx.x;
    const int r8 = lane >> 2, c4 = lane & 3;
    const int ldp = bsm_ldp(n), npad = (n + 7) & ~7;
    const unsigned pb = (unsigned)__cvta_generic_to_shared(sm);      // panel: column t at pb + 8 t ldp
    const unsigned tdb = pb + 8u * (unsigned)(KB * ldp);             // diagonal of T as of the last update
    const unsigned cdb = tdb + 8u * (unsigned)npad;                  // running diagonal, by position
    const unsigned gb = cdb + 8u * (unsigned)npad;                   // Gershgorin bounds (SE phase 2)
    const unsigned evb = gb + 8u * (unsigned)npad;                   // e, by position
    const unsigned rsb = evb + 8u * (unsigned)npad;                  // row sums / zero line
    LwmRing rg;
    rg.stb = rsb + 8u * (unsigned)npad;
    rg.mbar = rg.stb + 2u * 8u * 8u * kLwmStageLd;
    const unsigned permb = rg.mbar + 16u;                            // ints
    rg.ph[0] = rg.ph[1] = 0u;
    rg.tma = (n & 1) == 0 && ((((size_t)A) | ((size_t)L)) & 15) == 0;
    if (lane == 0) {
        mbar_init(rg.mbar, 1);
        mbar_init(rg.mbar + 8u, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    const size_t nn = (size_t)n * n;
    const unsigned pf = pb + 8u * (unsigned)(c4 * ldp + r8);         // P(r8, c4): fragment source

    for (long long b = blockIdx.x; b < batch; b += gridDim.x) {
        const double* Ab = A + (size_t)b * nn;
        double* Lb = L + (size_t)b * nn;

        // ---- in: A -> working buffer through the ring; diagonal and identity permutation ----------
        double om = 0.0;
        lwm_stream<true>(rg, Ab, Lb, n, 0, lane, [&](int bj, int ib, int ie, unsigned st) {
            if constexpr (kAlgo == kGMW81) {   // off-diagonal maximum (gmw81.rs:52-58)
                const int rows = min(8, n - 8 * bj);
                for (int r = 0; r < rows; ++r) {
                    const int k = 8 * bj + r;
#pragma unroll
                    for (int t = 0; t < kLwmChunk / 32; ++t) {
                        const int i = ib + lane + 32 * t;
                        if (i < ie && i > k)
                            om = fmax(om, fabs(lds_f64(st + 8u * (unsigned)(r * kLwmStageLd + i - ib))));
                    }
                }
            }
        });
        double dmaxabs = 0.0;
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const int i = lane + 32 * q;
            if (i < npad) {
                const double d = i < n ? __ldg(Ab + (size_t)i * (n + 1)) : 0.0;
                sts_f64(tdb + 8u * (unsigned)i, d);
                sts_f64(cdb + 8u * (unsigned)i, d);
                sts_f64(evb + 8u * (unsigned)i, 0.0);
                asm volatile("st.shared.u32 [%0], %1;" :: "r"(permb + 4u * (unsigned)i), "r"(i) : "memory");
                if (i < n)
                    dmaxabs = fmax(dmaxabs, fabs(d));
            }
        }
        bool win;
        // gamma / xi_diag = max |a_ii| folded from 0 (gmw81.rs:49-51, se90.rs:55-57, se99.rs:54-56)
        const double gamma = warp_extreme<true>(kFull, dmaxabs, true, 0.0, win);
        double delta = 0.0, rbeta = 0.0;
        if constexpr (kAlgo == kGMW81) {
            const double off_max = warp_extreme<true>(kFull, om, true, 0.0, win);
            const double nf = (double)n;
            delta = dmul(MODCHOL_EPS, fmax(1.0, dadd(gamma, off_max)));
            const double beta =
                dsqrt(fmax(fmax(gamma, ddiv(off_max, dsqrt(dsub(dmul(nf, nf), 1.0)))), MODCHOL_EPS));
            rbeta = ddiv(1.0, beta);
        }
        const double taugamma = dmul(MODCHOL_TAU, gamma);
        __syncwarp();

        // ---- the pieces of a step ---------------------------------------------------------------
        int j = 0, c = 0, j0 = 0, kstart = -1;
        double col[R];
        double sjj;   // sum_t P(j,t)^2 of the pending panel columns

        // symmetric exchange j <-> m (utils.rs:30-49) fused with the read of column j, then minus the
        // pending panel columns: col[q] = current T(i, j), i = lane + 32 q > j
        auto exchange_and_column = [&](int m) {
            const bool swp = m != j;
            double x[R], y[R];
            double *a1[R], *a2[R];
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const int i = lane + 32 * q;
                a1[q] = Lb + (size_t)j * n + i;
                a2[q] = i < m ? Lb + (size_t)i * n + m : Lb + (size_t)m * n + i;
                x[q] = y[q] = 0.0;
                if (i > j && i < n) {
                    x[q] = __ldcg(a1[q]);   // L2 only: the bulk copies write behind the L1
                    if (swp && i != m)
                        y[q] = __ldcg(a2[q]);
                }
            }
            if (swp) {
                // finished rows of L, panel rows, scalars
                for (int t = lane; t < j0; t += 32) {
                    const double u = __ldcg(Lb + (size_t)j * n + t), w = __ldcg(Lb + (size_t)m * n + t);
                    Lb[(size_t)j * n + t] = w;
                    Lb[(size_t)m * n + t] = u;
                }
                if (lane < c) {
                    const unsigned a = pb + 8u * (unsigned)(lane * ldp);
                    const double u = lds_f64(a + 8u * j), w = lds_f64(a + 8u * m);
                    sts_f64(a + 8u * j, w);
                    sts_f64(a + 8u * m, u);
                }
                if (lane >= 28) {
                    if (lane == 31) {
                        unsigned u, w;
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(u) : "r"(permb + 4u * j) : "memory");
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(permb + 4u * m) : "memory");
                        asm volatile("st.shared.u32 [%0], %1;" :: "r"(permb + 4u * j), "r"(w) : "memory");
                        asm volatile("st.shared.u32 [%0], %1;" :: "r"(permb + 4u * m), "r"(u) : "memory");
                    } else {
                        const unsigned a = lane == 28 ? tdb : (lane == 29 ? cdb : gb);
                        const double u = lds_f64(a + 8u * j), w = lds_f64(a + 8u * m);
                        sts_f64(a + 8u * j, w);
                        sts_f64(a + 8u * m, u);
                    }
                }
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    const int i = lane + 32 * q;
                    if (i > j && i < n && i != m) {
                        *a1[q] = y[q];
                        *a2[q] = x[q];
                        x[q] = y[q];
                    }
                }
            }
            __syncwarp();
            // left-looking inside the panel: minus P(i, 0..c) . P(j, 0..c)
            double s[R];
#pragma unroll
            for (int q = 0; q < R; ++q)
                s[q] = 0.0;
            sjj = 0.0;
            for (int t = 0; t < c; ++t) {
                const unsigned a = pb + 8u * (unsigned)(t * ldp);
                const double pj = lds_f64(a + 8u * j);
                sjj = __fma_rn(pj, pj, sjj);
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    const int i = lane + 32 * q;
                    if (32 * q < npad)   // (rows beyond npad do not exist in the panel)
                        s[q] = __fma_rn(lds_f64(a + 8u * (unsigned)min(i, npad - 1)), pj, s[q]);
                }
            }
#pragma unroll
            for (int q = 0; q < R; ++q)
                col[q] = dsub(x[q], s[q]);
        };
        // column c of the panel: cs[q] for the rows below j, sq on the diagonal; e_j
        auto append = [&](const double (&cs)[R], double sq, double ej) {
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const int i = lane + 32 * q;
                if (i > j && i < n)
                    sts_f64(pb + 8u * (unsigned)(c * ldp + i), cs[q]);
            }
            if (lane == 0) {
                sts_f64(pb + 8u * (unsigned)(c * ldp + j), sq);
                sts_f64(evb + 8u * (unsigned)j, ej);
            }
            __syncwarp();
            ++j;
            ++c;
        };
        // end of a panel: its entries of L go to their final place, the trailing block is updated
        auto flush = [&]() {
            if (c == 0)
                return;
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const int i = lane + 32 * q;
                if (i >= j0 && i < n) {
                    const int cnt = min(c, i - j0 + 1);
                    double* dst = Lb + (size_t)i * n + j0;
                    for (int t = 0; t < cnt; ++t)
                        dst[t] = lds_f64(pb + 8u * (unsigned)(t * ldp + i));
                }
            }
            if (j < n) {
                lwm_stream<true>(rg, Lb, Lb, n, j >> 3, lane, [&](int bj, int ib, int ie, unsigned st) {
                    double fa[KC];
#pragma unroll
                    for (int h = 0; h < KC; ++h)
                        fa[h] = wip_neg(lds_f64(pf + 8u * (unsigned)(4 * h * ldp + 8 * bj)));
                    const int k = 8 * bj + r8;
                    for (int bi = ib >> 3; 8 * bi < ie; ++bi) {
                        const unsigned ta = st + 8u * (unsigned)(r8 * kLwmStageLd + 8 * bi - ib + 2 * c4);
                        double2 v = lds_v2f64(ta);
                        const double o0 = v.x, o1 = v.y;
                        const int i0 = 8 * bi + 2 * c4;
                        if (bi == bj) {   // the diagonal of T lives in shared memory
                            if (i0 == k) v.x = lds_f64(tdb + 8u * (unsigned)k);
                            if (i0 + 1 == k) v.y = lds_f64(tdb + 8u * (unsigned)k);
                        }
#pragma unroll
                        for (int h = 0; h < KC; ++h) {
                            const double fb = lds_f64(pf + 8u * (unsigned)(4 * h * ldp + 8 * bi));
                            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                                         : "+d"(v.x), "+d"(v.y) : "d"(fa[h]), "d"(fb));
                        }
                        if (bi == bj) {   // entries at and below the diagonal of the tile are not T
                            if (i0 == k) sts_f64(tdb + 8u * (unsigned)k, v.x);
                            if (i0 + 1 == k) sts_f64(tdb + 8u * (unsigned)k, v.y);
                            if (i0 <= k) v.x = o0;
                            if (i0 + 1 <= k) v.y = o1;
                        }
                        sts_v2f64(ta, v.x, v.y);
                    }
                });
            }
            c = 0;
        };
        auto clear_panel = [&]() {
            for (int q = lane; q < KB * ldp; q += 32)
                sts_f64(pb + 8u * (unsigned)q, 0.0);
            __syncwarp();
            j0 = j;
            c = 0;
        };

        if constexpr (kAlgo == kGMW81) {
            // ---- GMW81, symmetric form (see wsm_gmw.cuh): L entries are c_is / sqrt(d_s) ----------
            while (j < n) {
                clear_panel();
                while (c < KB && j < n) {
                    const int m = lwm_argmax<R, true>(cdb, j, n, lane);   // gmw81.rs:71-78
                    exchange_and_column(m);
                    const double cjj = dsub(lds_f64(tdb + 8u * (unsigned)j), sjj);   // from the same updates
                    double th = 0.0;
#pragma unroll
                    for (int q = 0; q < R; ++q) {
                        const int i = lane + 32 * q;
                        if (i > j && i < n)
                            th = fmax(th, fabs(col[q]));
                    }
                    const double theta = warp_extreme<true>(kFull, th, true, 0.0, win);   // gmw81.rs:87-100
                    const double tb = dmul(theta, rbeta);
                    const double dj = fmax(fmax(fabs(cjj), dmul(tb, tb)), delta);     // gmw81.rs:102
                    double y = rsqrt_normal(dj);
                    double sq = dmul(dj, y);
                    if (!fast_range_ok(dj))
                        wip_slow_gmw(dj, sq, y);
                    double cs[R];
#pragma unroll
                    for (int q = 0; q < R; ++q) {
                        const int i = lane + 32 * q;
                        cs[q] = dmul(col[q], y);
                        if (i > j && i < n) {                                         // gmw81.rs:104-109
                            const unsigned a = cdb + 8u * (unsigned)i;
                            sts_f64(a, __fma_rn(-cs[q], cs[q], lds_f64(a)));
                        }
                    }
                    append(cs, sq, dsub(dj, cjj));                                    // gmw81.rs:110
                }
                flush();
            }
        } else {
            // ---- SE90 / SE99 -------------------------------------------------------------------------
            double delta_prev = 0.0, dmin_carry = 0.0;
            bool have_dmin = false;
            int st = 0;   // 0 phase 1 running, 1 switch to phase 2 at kstart, 2 finished in phase 1
            // phase 1 (se90.rs:61-96, se99.rs:61-109)
            while (st == 0) {
                clear_panel();
                while (c < KB) {
                    if (j == n) {
                        st = 2;
                        break;
                    }
                    const int m = lwm_argmax<R, false>(cdb, j, n, lane);
                    const double piv = lds_f64(cdb + 8u * (unsigned)m);
                    if (kAlgo == kSE99) {   // se99.rs:62-73, before the pivot is applied
                        double dmin = dmin_carry;
                        if (!have_dmin) {
                            double lm = INFINITY;
#pragma unroll
                            for (int q = 0; q < R; ++q) {
                                const int i = lane + 32 * q;
                                if (i >= j && i < n) {
                                    const double v = lds_f64(cdb + 8u * (unsigned)i);
                                    lm = v < lm ? v : lm;
                                }
                            }
                            dmin = warp_extreme<false>(kFull, lm, true, INFINITY, win);
                        }
                        const double dmax = (piv == piv) ? piv : -INFINITY;
                        if (dmax < taugamma || dmin < dmul(-MODCHOL_MU, dmax)) {
                            kstart = j;
                            st = 1;
                            break;
                        }
                    }
                    exchange_and_column(m);
                    const bool fast = fast_range_ok(piv);
                    const double y = rsqrt_normal(piv);
                    double sq = fast ? dmul(piv, y) : dsqrt(piv);
                    double cs[R], dgn[R];
                    double lm = INFINITY;
#pragma unroll
                    for (int q = 0; q < R; ++q) {
                        const int i = lane + 32 * q;
                        cs[q] = dgn[q] = 0.0;
                        if (i > j && i < n) {
                            const double dg = lds_f64(cdb + 8u * (unsigned)i);
                            double nv;   // look-ahead value (se90.rs:70-77, se99.rs:82-89)
                            if (fast) {
                                cs[q] = dmul(col[q], y);
                                dgn[q] = __fma_rn(-cs[q], cs[q], dg);
                                nv = dgn[q];
                            } else {
                                cs[q] = ddiv(col[q], sq);
                                dgn[q] = __fma_rn(-cs[q], cs[q], dg);
                                nv = dsub(dg, ddiv(dmul(col[q], col[q]), piv));
                            }
                            lm = nv < lm ? nv : lm;
                        }
                    }
                    const double la = warp_extreme<false>(kFull, lm, true, INFINITY, win);
                    have_dmin = fast;
                    dmin_carry = la;
                    const double thresh = (kAlgo == kSE99) ? dmul(-MODCHOL_MU, gamma) : taugamma;
                    if (la < thresh) {   // the pivot at j stays applied (se99.rs:90-93)
                        kstart = j;
                        st = 1;
                        break;
                    }
#pragma unroll
                    for (int q = 0; q < R; ++q) {
                        const int i = lane + 32 * q;
                        if (i > j && i < n)
                            sts_f64(cdb + 8u * (unsigned)i, dgn[q]);
                    }
                    append(cs, sq, 0.0);
                }
                flush();
            }
            if (st == 1 && kAlgo == kSE99 && kstart == n - 1) {   // se99.rs:115-119
                const double ljj = lds_f64(cdb + 8u * (unsigned)(n - 1));
                const double ej = tail1x1_e(ljj, gamma);
                if (lane == 0) {
                    sts_f64(evb + 8u * (unsigned)(n - 1), ej);
                    Lb[(size_t)(n - 1) * n + (n - 1)] = dsqrt(dadd(ljj, ej));
                }
                __syncwarp();
            } else if (st == 1) {
                // ---- phase 2 (se90.rs:101-167, se99.rs:121-187) ------------------------------------
                // Gershgorin bounds of the trailing block as the reference holds it (the pending
                // columns were applied by flush()): g_i = a_ii - sum_{k >= j, k != i} |a_ik|
                for (int q = lane; q < npad; q += 32)
                    sts_f64(rsb + 8u * (unsigned)q, 0.0);
                __syncwarp();
                lwm_stream<false>(rg, Lb, Lb, n, j >> 3, lane, [&](int bj, int ib, int ie, unsigned stg) {
                    double cacc[kLwmChunk / 32];
#pragma unroll
                    for (int t = 0; t < kLwmChunk / 32; ++t)
                        cacc[t] = 0.0;
                    const int rows = min(8, n - 8 * bj);
                    for (int r = 0; r < rows; ++r) {
                        const int k = 8 * bj + r;
                        if (k < j)
                            continue;
                        double part = 0.0;
#pragma unroll
                        for (int t = 0; t < kLwmChunk / 32; ++t) {
                            const int i = ib + lane + 32 * t;
                            if (i < ie && i > k) {
                                const double v = fabs(lds_f64(stg + 8u * (unsigned)(r * kLwmStageLd + i - ib)));
                                part = dadd(part, v);
                                cacc[t] = dadd(cacc[t], v);
                            }
                        }
                        const double rowsum = wip_sum32(part);
                        if (lane == 0)
                            sts_f64(rsb + 8u * (unsigned)k, dadd(lds_f64(rsb + 8u * (unsigned)k), rowsum));
                    }
                    __syncwarp();
#pragma unroll
                    for (int t = 0; t < kLwmChunk / 32; ++t) {
                        const int i = ib + lane + 32 * t;
                        if (i < ie)
                            sts_f64(rsb + 8u * (unsigned)i, dadd(lds_f64(rsb + 8u * (unsigned)i), cacc[t]));
                    }
                });
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    const int i = lane + 32 * q;
                    if (i >= j && i < n)
                        sts_f64(gb + 8u * (unsigned)i,
                                dsub(lds_f64(cdb + 8u * (unsigned)i), lds_f64(rsb + 8u * (unsigned)i)));
                }
                __syncwarp();
                while (j + 2 < n) {
                    clear_panel();
                    while (c < KB && j + 2 < n) {
                        const int m = lwm_argmax<R, false>(gb, j, n, lane);   // se90.rs:115, se99.rs:135
                        exchange_and_column(m);
                        double piv = lds_f64(cdb + 8u * (unsigned)j);
                        double acol[R], part = 0.0;
#pragma unroll
                        for (int q = 0; q < R; ++q) {
                            const int i = lane + 32 * q;
                            acol[q] = (i > j && i < n) ? abs_bits(col[q]) : 0.0;
                            part = dadd(part, acol[q]);
                        }
                        // E_jj (se90.rs:124-131, se99.rs:144-151)
                        const double normj = wip_sum32(part);
                        const double ej = max_b_not_nan(dadd(-piv, max_b_not_nan(normj, taugamma)), delta_prev);
                        piv = dadd(piv, ej);
                        delta_prev = ej;
                        // bound update (se90.rs:134-139, se99.rs:154-159) and column scaling
                        const bool upd = fabs(dsub(piv, normj)) > MODCHOL_EPS;
                        const double y = rsqrt_normal(piv);
                        double sq = dmul(piv, y);
                        double t = __fma_rn(-normj, dmul(y, y), 1.0);
                        const bool fast = fast_range_ok(piv);
                        if (!fast) {
                            sq = dsqrt(piv);
                            t = dsub(1.0, ddiv(normj, piv));
                        }
                        double cs[R];
#pragma unroll
                        for (int q = 0; q < R; ++q) {
                            const int i = lane + 32 * q;
                            cs[q] = fast ? dmul(col[q], y) : ddiv(col[q], sq);
                            if (i > j && i < n) {
                                const unsigned a = cdb + 8u * (unsigned)i;
                                sts_f64(a, __fma_rn(-cs[q], cs[q], lds_f64(a)));
                                if (upd) {
                                    const unsigned ga = gb + 8u * (unsigned)i;
                                    sts_f64(ga, __fma_rn(acol[q], t, lds_f64(ga)));
                                }
                            }
                        }
                        append(cs, sq, ej);
                    }
                    flush();
                }
                // final 2x2 block, never pivoted (se90.rs:155-166, se99.rs:175-186); the trailing
                // block is up to date
                double a00 = lds_f64(cdb + 8u * (unsigned)(n - 2));
                double a11 = lds_f64(cdb + 8u * (unsigned)(n - 1));
                double a10 = __ldcg(Lb + (size_t)(n - 2) * n + (n - 1));
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
                __syncwarp();
                if (lane == 0) {
                    sts_f64(evb + 8u * (unsigned)(n - 2), en);
                    sts_f64(evb + 8u * (unsigned)(n - 1), en);
                    Lb[(size_t)(n - 2) * n + (n - 2)] = a00;
                    Lb[(size_t)(n - 1) * n + (n - 2)] = a10;
                    Lb[(size_t)(n - 1) * n + (n - 1)] = a11;
                }
                __syncwarp();
            }
        }

        // ---- out: strict upper triangle zero (gmw81.rs:112-125 / se99.rs:189-192), e in original
        // order (se99.rs:194-198), p --------------------------------------------------------------
        __syncwarp();
        if (rg.tma) {
            for (int q = lane; q < npad; q += 32)
                sts_f64(rsb + 8u * (unsigned)q, 0.0);
            fence_proxy_async();
            __syncwarp();
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const int k = lane + 32 * q;
                if (k < n - 1) {
                    int s0 = k + 1;
                    if (s0 & 1) {
                        Lb[(size_t)k * n + s0] = 0.0;
                        ++s0;
                    }
                    if (s0 < n)
                        bulk_s2g(Lb + (size_t)k * n + s0, rsb, 8u * (unsigned)(n - s0));
                }
            }
            bulk_commit();
            bulk_wait_read0();
        } else {
            for (int k = 0; k < n; ++k)
                for (int i = k + 1 + lane; i < n; i += 32)
                    Lb[(size_t)k * n + i] = 0.0;
        }
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const int i = lane + 32 * q;
            if (i < n) {
                unsigned pi;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(pi) : "r"(permb + 4u * (unsigned)i) : "memory");
                E[(size_t)b * n + pi] = lds_f64(evb + 8u * (unsigned)i);
                P[(size_t)b * n + i] = (unsigned long long)pi;
            }
        }
        if (K && lane == 0)
            K[b] = kstart;
        __syncwarp();
    }
    if (rg.tma)
        bulk_wait0();
}

}  // namespace modchol
