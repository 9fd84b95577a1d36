// wip_kernels.cuh -- FAST-mode GMW81 / SE90 / SE99 for 16 < n <= 32 with IMPLICIT pivoting:
// one warp per matrix, lane i owns ORIGINAL row i for the whole factorisation and nothing is
// ever swapped (gmw81.rs:39-134, se90.rs:38-181, se99.rs:35-201; the swap of utils.rs:30-49 is
// replaced by the per-lane `pos`, the permuted position the reference would hold the row at).
//
// Where the data lives (per warp):
//   * C, the symmetric trailing matrix in ORIGINAL coordinates, lives in REGISTERS as the
//     accumulator fragments of ten fixed 8x8 tiles (lower block triangle, 40 registers per
//     lane) and is brought up to date every FOUR steps by one mma.sync.m8n8k4.f64 (DMMA) per
//     tile -- no load or store of C around the tensor-pipe instruction.  After each update the
//     fragments are also DUMPED into a packed lower triangle in shared memory (element (i,k),
//     i >= k, at T(i)+k), which is all a step needs: lane i > r reads T(i)+r (the triangular
//     numbers are a permutation modulo 16 on aligned runs of 16 rows), lane i < r reads T(r)+i
//     (contiguous) -- one LDS per lane and step, no swap, no row scalars to exchange.  Between
//     updates a column needs at most three fused multiply-adds with the pending panel columns.
//   * the four pending columns go through a 32x4 panel in shared memory whose layout makes the
//     per-step column write and the fragment load both conflict free.
//   * L is written to TENSOR MEMORY: lane i stores L(i,t) at step t into its own TMEM lane,
//     columns 2t..2t+1 (tcgen05.st, one instruction per step, off the shared-memory pipe), so
//     finished entries never have to be protected from the dumps.  A finished row stores exact
//     zeros, so its row of L ends at the diagonal (se99.rs:189-192) without a pass over it.  In
//     the epilogue every lane reads its row back (tcgen05.ld) and writes it to row pos of the
//     output with 256-bit stores -- no un-permutation pass.
//   * pivot search: order-preserving integer keys + ONE REDUX when the high word of the
//     maximum is unique (practically always); ties fall back to the full first-in-position
//     search of utils.rs:52-91 on (hi, lo, pos).
//   * the Gershgorin bounds of a phase 2 that starts at column 0 (symmetric-uniform inputs)
//     come from absolute column sums accumulated while the matrix streams in.
// FAST arithmetic only (fused multiply-add, rsqrt scaling, tree-ordered norms): results agree
// with the oracle within the north-star tolerances, not bit for bit -- STRICT mode keeps the
// position-ordered kernels of wsm_se.cuh / wsm_gmw.cuh.
#pragma once
#include <type_traits>

#include "wsm_se.cuh"

#ifndef MODCHOL_WIP_DMMA_SUM
#define MODCHOL_WIP_DMMA_SUM 1   /* column norm: two DMMAs (1), one DMMA + two shuffle rounds (2) or five shuffle rounds (0) */
#endif

namespace modchol {

constexpr int kWipTri = 528;       // T(32) doubles of packed triangle
constexpr int kWipPanelLd = 36;    // panel column stride: == 4 (mod 16) doubles, see above
#ifndef MODCHOL_WIP_KB
#define MODCHOL_WIP_KB 8
#endif
constexpr int kWipKB = MODCHOL_WIP_KB;   // steps between two trailing updates (4 or 8)
// Two layouts of the panel of pending columns, chosen per size class (measured, profiles/r02_notes.md):
//   by columns (n > 24): column c at 8 c kWipPanelLd bytes, every lane keeps its own entries in
//       registers, one LDS + FMA per pending column;
//   by rows (n <= 24): row i at 8 i kWipRowLd bytes (80-byte rows: the 128-bit reads of a lane's own row
//       are conflict free), no register copies, four pending columns per pair of 128-bit reads.
#ifndef MODCHOL_WIP_ROWPANEL
#define MODCHOL_WIP_ROWPANEL 2   /* 0 by columns, 1 by rows, 2 by rows for n <= 24 */
#endif
constexpr int kWipRowLd = kWipKB + 2;
constexpr int kWipPanelDoubles = kWipKB * kWipPanelLd > 32 * kWipRowLd ? kWipKB * kWipPanelLd : 32 * kWipRowLd;
constexpr int kWipDoubles = kWipTri + kWipPanelDoubles;
constexpr int kWipWarps = 4;       // warp w of a block owns TMEM lanes [32 w, 32 w + 32)
constexpr int kWipTmemCols = 64;   // 32 doubles of L per lane

template <int I>
using IC = std::integral_constant<int, I>;

// f(IC<I>) for I = Begin .. End-1 until one call returns true (-> true)
template <int Begin, int End, typename F>
__device__ __forceinline__ bool wip_steps(F&& f)
{
    if constexpr (Begin < End) {
        if (f(IC<Begin>{}))
            return true;
        return wip_steps<Begin + 1, End>(static_cast<F&&>(f));
    } else {
        return false;
    }
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
    unsigned b1 = __ballot_sync(kFull, c1);
    if (__popc(b1) != 1) {
        const unsigned ml = __reduce_max_sync(kFull, c1 ? klo : 0u);
        const bool win = c1 && klo == ml;
        unsigned mpos = __reduce_min_sync(kFull, win ? (unsigned)pos : 1024u);
        if (mpos > 255u)
            mpos = (unsigned)j;
        b1 = __ballot_sync(kFull, !fin && pos == (int)mpos);
    }
    return 31 - __clz(b1);   // exactly one bit is set
}

__device__ __forceinline__ double wip_neg(double x)   // -x on the integer pipe
{
    return __hiloint2double(__double2hiint(x) ^ (int)0x80000000, __double2loint(x));
}

// sum of x over the 32 lanes on the FP64 tensor pipe, returned to every lane: with x as the B
// fragment (lane l holds B[l%4][l/4]) and A = 1, D[m][n] = sum_k x_{4n+k} is the sum of lane
// quad n, and lane l receives D[l/4][2(l%4)], D[l/4][2(l%4)+1] -- the four lanes of a quad end
// up holding all eight quad sums between them; their pairwise sums as the A fragment of a second
// DMMA against B = 1 give the total in every lane.  Three instructions instead of five shuffle
// rounds (15), and a shorter dependency chain.
__device__ __forceinline__ double wip_sum32(double x)
{
    double q0, q1, s0, s1;
    const double one = 1.0, zero = 0.0;
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%4, %4};"
                 : "=d"(q0), "=d"(q1) : "d"(one), "d"(x), "d"(zero));
    double h = dadd(q0, q1);
#if MODCHOL_WIP_DMMA_SUM == 2
    // the four lanes of a quad hold the pairwise sums of the eight quad sums: two shuffle rounds
    h = dadd(h, __shfl_xor_sync(kFull, h, 1));
    h = dadd(h, __shfl_xor_sync(kFull, h, 2));
    return h;
#else
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%4, %4};"
                 : "=d"(s0), "=d"(s1) : "d"(h), "d"(one), "d"(zero));
    return s0;
#endif
}

// pivots outside the safe range of the rsqrt formulas (<= 0, NaN, Inf, subnormal, huge): the
// STRICT formulas (rare; the fast values are computed unconditionally and overridden here)
__device__ __forceinline__ void wip_slow_p2(double piv, double normj, double col, double& sq, double& t, double& cs)
{
    sq = dsqrt(piv);
    t = dsub(1.0, ddiv(normj, piv));
    cs = ddiv(col, sq);
}
__device__ __forceinline__ void wip_slow_p1(double piv, double col, double dg, double& sq, double& cs, double& dgn,
                                         double& nv)
{
    sq = dsqrt(piv);
    cs = ddiv(col, sq);
    dgn = __fma_rn(-cs, cs, dg);
    nv = dsub(dg, ddiv(dmul(col, col), piv));
}
__device__ __forceinline__ void wip_slow_gmw(double dj, double& sq, double& y)
{
    sq = dsqrt(dj);
    y = ddiv(1.0, sq);
}

// L(lane, t) -> this lane's TMEM row (taddr: warp-uniform, lane quarter + column 2 t)
__device__ __forceinline__ void wip_tmem_store(unsigned taddr, double v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};"
                 :: "r"(taddr), "r"(__double2loint(v)), "r"(__double2hiint(v)) : "memory");
}

// eight doubles of this lane's TMEM row, starting at column taddr
__device__ __forceinline__ void wip_tmem_load8(unsigned taddr, double (&v)[8])
{
    unsigned w[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]),
                   "=r"(w[7]), "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]),
                   "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int u = 0; u < 8; ++u)
        v[u] = __hiloint2double((int)w[2 * u + 1], (int)w[2 * u]);
}

__device__ __forceinline__ void stg_v4f64(double* p, double a, double b, double c, double d)
{
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" :: "l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

// kN: compile-time n (0: run-time n <= 8 NB).  kRegC: C fragments resident in registers (40 per
// lane, 20 warps per SM) or in the shared-memory triangle only (one LDS/DMMA/STS round trip per
// tile and update, 64 registers, 32 warps per SM).
// kSolve: the consumer step (A + E) x = rhs fused behind the factorisation (Bm, X: batch x nrhs x n,
// X may alias Bm): the epilogue also lays the rows of L down in the shared-memory triangle in
// POSITION order (row pos of lane i at T(pos)), the lanes change role (lane = position) and run
// the substitutions of wsm_solve_in_place on it -- L does not make the round trip through HBM, and
// is not written at all when the caller passes L == NULL.
template <int kAlgo, int NB, int kN, bool kRegC, int kMinBlocks, bool kSolve = false>
__global__ void __launch_bounds__(kWipWarps * 32, kMinBlocks)
wip_kernel(int n_rt, long long batch, const double* __restrict__ A, double* __restrict__ L,
           double* __restrict__ E, unsigned long long* __restrict__ P, int* __restrict__ K,
           const double* Bm = nullptr, double* X = nullptr, int nrhs = 0)
{
    extern __shared__ double sm[];
    __shared__ unsigned tmem_base;
    constexpr int kTiles = NB * (NB + 1) / 2;
    constexpr bool kWipRowPanel = MODCHOL_WIP_ROWPANEL == 1 || (MODCHOL_WIP_ROWPANEL == 2 && NB <= 3);
    const int n = kN ? kN : n_rt;
    int lane = threadIdx.x & 31;
    asm volatile("mov.u32 %0, %0;" : "+r"(lane));   // pinned (see wsm_se.cuh)
    // warp index through a shuffle from lane 0: ptxas then treats it (and the batch loop below) as
    // warp-uniform, so the .sync operations inside need no divergence check (BRA.DIV)
    const int wib = __shfl_sync(kFull, (int)(threadIdx.x >> 5), 0);

    // ---- tensor memory: 64 columns per block, warp w owns lanes [32 w, 32 w + 32) -------------
    if (wib == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"((unsigned)__cvta_generic_to_shared(&tmem_base)), "n"(kWipTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tw = tmem_base + ((unsigned)(wib * 32) << 16);

    unsigned wb = (unsigned)__cvta_generic_to_shared(sm + (size_t)wib * kWipDoubles);
    asm volatile("mov.u32 %0, %0;" : "+r"(wb));
    const unsigned pnb = wb + 8u * kWipTri;                  // kWipKB panel columns of kWipPanelLd
    unsigned tib = wb + 8u * (unsigned)tri(lane);            // byte address of slot (lane, 0)
    asm volatile("mov.u32 %0, %0;" : "+r"(tib));
    const unsigned pcol = pnb + 8u * (unsigned)lane;         // this lane's entry of panel column 0
    const int r8 = lane >> 2, c4 = lane & 3;
    const unsigned pf = kWipRowPanel ? pnb + 8u * (unsigned)(kWipRowLd * r8 + c4)
                                     : pnb + 8u * (unsigned)(kWipPanelLd * c4 + r8);      // fragment source
    const unsigned prow = pnb + 8u * (unsigned)(kWipRowLd * lane);         // row-major panel: this lane's row
    const unsigned fr0 = wb + 8u * (unsigned)(tri(r8) + 2 * c4);           // slot (r8, 2 c4)
    const bool pd0 = 2 * c4 <= r8, pd1 = 2 * c4 + 1 <= r8;                 // diagonal tiles: lower part
    const size_t nn = (size_t)n * n;
    const long long nslots = (long long)gridDim.x * kWipWarps;
    const bool has = lane < n;
    // 128/256-bit global accesses need aligned bases and an even row length
    const bool wide = kN != 0 && kN % 8 == 0 && (((size_t)A | (size_t)L) & 31) == 0;

    for (long long b = (long long)blockIdx.x * kWipWarps + wib; b < batch; b += nslots) {
        const double* Ab = A + (size_t)b * nn;
        // ---- load -----------------------------------------------------------------------------
        // kRegC: the ten tiles straight into the accumulator fragments (lane l: rows 8 bi + l/4,
        // columns 8 bj + 2 (l%4) + {0,1}); otherwise row by row into the packed triangle (lane =
        // column).  SE: absolute off-diagonal column sum (== row sum, A is symmetric) from the
        // streamed rows; GMW81: off-diagonal maximum (gmw81.rs:52-58), kept as integers (|.| of
        // non-NaN doubles order like their bits).
        double cf[kRegC ? kTiles : 1][2];
        if constexpr (kRegC) {
            const double* fp = Ab + (size_t)r8 * n + 2 * c4;
#pragma unroll
            for (int bi = 0; bi < NB; ++bi)
#pragma unroll
                for (int bj = 0; bj <= bi; ++bj) {
                    const int t = bi * (bi + 1) / 2 + bj;
                    const double* p = fp + (size_t)(8 * bi) * n + 8 * bj;
                    if (wide) {
                        const double2 v = __ldg(reinterpret_cast<const double2*>(p));
                        cf[t][0] = v.x;
                        cf[t][1] = v.y;
                    } else {
                        const int i = 8 * bi + r8, k = 8 * bj + 2 * c4;
                        cf[t][0] = (i < n && k < n) ? __ldg(p) : 0.0;
                        cf[t][1] = (i < n && k + 1 < n) ? __ldg(p + 1) : 0.0;
                    }
                }
        }
        double dg = has ? __ldg(Ab + (size_t)lane * (n + 1)) : 0.0;   // this lane's diagonal element
        double asum = 0.0;
        unsigned long long om = 0ull;
        if constexpr (!kRegC)
            __syncwarp();   // the previous matrix's readers of the triangle are done
        if constexpr (kAlgo != kGMW81 || !kRegC) {
            const double* src = Ab + lane;
            unsigned dst = wb + 8u * (unsigned)lane;
            auto take = [&](int i, double v) {
                if constexpr (!kRegC) {
                    sts_f64_if(lane <= i, dst, v);
                    dst += 8u * (unsigned)(i + 1);
                }
                if constexpr (kAlgo == kGMW81) {
                    const unsigned long long a =
                        (unsigned long long)__double_as_longlong(v) & 0x7fffffffffffffffull;
                    om = (lane != i && a > om) ? a : om;
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

        // fragments -> packed triangle in shared memory (every slot, finished or not: L lives in
        // tensor memory, so the triangle only ever holds C as of the last update)
        auto dump = [&]() {
            if constexpr (!kRegC)
                return;
#pragma unroll
            for (int bi = 0; bi < NB; ++bi) {
                // T(8 bi + r8) = T(8 bi) + T(r8) + 8 bi r8
                const unsigned ra = fr0 + 8u * (unsigned)(tri(8 * bi) + 8 * bi * r8);
#pragma unroll
                for (int bj = 0; bj <= bi; ++bj) {
                    const int t = bi * (bi + 1) / 2 + bj;
                    if (bi == bj) {
                        sts_f64_if(pd0, ra + 64u * bj, cf[t][0]);
                        sts_f64_if(pd1, ra + 64u * bj + 8, cf[t][1]);
                    } else {
                        sts_f64(ra + 64u * bj, cf[t][0]);
                        sts_f64(ra + 64u * bj + 8, cf[t][1]);
                    }
                }
            }
            __syncwarp();
        };
        // C(i,k) -= sum_{c < cnt} P(i,c) P(k,c): one DMMA per tile (A fragment of block row bi:
        // lane l holds P(8 bi + l/4, l%4); the same pattern is the B fragment of block column bj)
        auto rank_update = [&](int cnt) {
            constexpr int KC = kWipKB / 4;   // k-chunks of four panel columns
            double frag[KC][NB], nfrag[KC][NB];
#pragma unroll
            for (int h = 0; h < KC; ++h)
#pragma unroll
                for (int q = 0; q < NB; ++q) {
                    if constexpr (kWipRowPanel) {   // columns not yet produced hold zeros
                        frag[h][q] = lds_f64(pf + 8u * (unsigned)(8 * q * kWipRowLd + 4 * h));
                    } else {
                        const unsigned a = pf + 8u * (unsigned)(4 * kWipPanelLd * h) + 64u * q;
                        frag[h][q] = cnt == kWipKB ? lds_f64(a) : lds_f64_or_zero(4 * h + c4 < cnt, a);
                    }
                    nfrag[h][q] = wip_neg(frag[h][q]);
                }
#pragma unroll
            for (int bi = 0; bi < NB; ++bi) {
                const unsigned ra = fr0 + 8u * (unsigned)(tri(8 * bi) + 8 * bi * r8);
#pragma unroll
                for (int bj = 0; bj <= bi; ++bj) {
                    if constexpr (kRegC) {
                        const int t = bi * (bi + 1) / 2 + bj;
#pragma unroll
                        for (int h = 0; h < KC; ++h)
                            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                                         : "+d"(cf[t][0]), "+d"(cf[t][1]) : "d"(nfrag[h][bi]), "d"(frag[h][bj]));
                    }
                }
                if constexpr (!kRegC) {
                    // (diagonal tiles: the entries above the diagonal are other rows' slots -- loaded,
                    // updated and not stored.)  Per block row: every load first, then the tensor-pipe
                    // operations, then the stores; the accessors are volatile asm, kept in program order
                    double c0[NB], c1[NB];
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
                                         : "+d"(c0[bj]), "+d"(c1[bj]) : "d"(nfrag[h][bi]), "d"(frag[h][bj]));
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
            }
            if constexpr (kRegC)
                dump();
            else
                __syncwarp();
            if constexpr (kWipRowPanel) {   // every lane clears its own panel row (read by no one else now)
#pragma unroll
                for (int c = 0; c < kWipKB; c += 2)
                    asm volatile("st.shared.v2.f64 [%0], {%1, %1};" :: "r"(prow + 8u * c), "d"(0.0) : "memory");
                __syncwarp();
            }
        };
        if constexpr (kRegC) {
            __syncwarp();   // the previous matrix's readers of the triangle are done
            dump();
        } else {
            __syncwarp();
        }
        if constexpr (kWipRowPanel) {
#pragma unroll
            for (int c = 0; c < kWipKB; c += 2)
                asm volatile("st.shared.v2.f64 [%0], {%1, %1};" :: "r"(prow + 8u * c), "d"(0.0) : "memory");
            __syncwarp();
        }

        double ev = 0.0;
        int pos = lane;                       // permuted position of this lane's row
        bool fin = !has;                      // row already eliminated (or no row)
        unsigned fmask = n >= 32 ? 0u : (0xffffffffu << n);
        int j = 0, cnt = 0, kstart = -1;
        double csr[kWipKB - 1];               // this row's entries of the pending panel columns
#pragma unroll
        for (int c = 0; c < kWipKB - 1; ++c)
            csr[c] = 0.0;
        bool win;

        // current value of C(lane, r): stored slot minus the C pending panel columns.  C is a run-time
        // count: one copy of every step in the instruction stream (eight unrolled copies, one per
        // count, made the hot loop 50 KB of code and cost 2 stall cycles per issue in instruction
        // fetch); the switch falls through the pending columns C-1 .. 0.
        auto column = [&](auto C, int r) -> double {
            const unsigned tr = wb + 4u * (unsigned)(r * (r + 1));
            double col = lds_f64(lane > r ? tib + 8u * (unsigned)r : tr + 8u * (unsigned)lane);
            if constexpr (kWipRowPanel) {
                // row-major panel: this row's and the pivot row's entries of the pending columns as
                // 128-bit reads (the pivot row's are broadcasts); columns not yet produced are zero,
                // the upper half of the panel is skipped while it is empty
                const unsigned pq = pnb + 8u * (unsigned)(kWipRowLd * r);
                auto half = [&](unsigned off) {
                    const double2 o0 = lds_v2f64(prow + off), o1 = lds_v2f64(prow + off + 16u);
                    const double2 q0 = lds_v2f64(pq + off), q1 = lds_v2f64(pq + off + 16u);
                    double s0 = dmul(o0.x, q0.x), s1 = dmul(o0.y, q0.y);
                    s0 = __fma_rn(o1.x, q1.x, s0);
                    s1 = __fma_rn(o1.y, q1.y, s1);
                    col = dsub(col, dadd(s0, s1));
                };
                int cv;
                if constexpr (std::is_integral_v<decltype(C)>) cv = C; else cv = decltype(C)::value;
                if (cv > 0)
                    half(0u);
                if constexpr (kWipKB > 4)
                    if (cv > 4)
                        half(32u);
                return col;
            }
            const unsigned pr = pnb + 8u * (unsigned)r;
            if constexpr (!std::is_integral_v<decltype(C)>) {   // compile-time count (unrolled callers)
#pragma unroll
                for (int c = 0; c < decltype(C)::value; ++c)
                    col = __fma_rn(-csr[c], lds_f64(pr + 8u * (unsigned)(kWipPanelLd * c)), col);
                return col;
            } else {
            auto corr = [&](auto ctag) {
                constexpr int c = decltype(ctag)::value;
                if constexpr (c < kWipKB - 1)
                    col = __fma_rn(-csr[c], lds_f64(pr + 8u * (unsigned)(kWipPanelLd * c)), col);
            };
            // a tree of warp-uniform branches on the count (no jump table: its constant load and
            // indirect branch sat on the step's critical path twice)
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
            return col;
            }
        };
        auto column_rt = [&](int c, int r) -> double { return column(c, r); };
        // the pivot row r takes position j; the row that sat there takes r's old position
        auto take_position = [&](int r) {
            const int pr = __shfl_sync(kFull, pos, r);
            pos = pos == j ? pr : pos;
            pos = lane == r ? j : pos;
        };
        // column j of L: cs (sq for the pivot row) into tensor memory and panel column C;
        // row r is finished
        auto retire = [&](auto C, int r, double cs, double sq) {
            wip_tmem_store(tw + 2u * (unsigned)j, lane == r ? sq : cs);
            if constexpr (kWipRowPanel) {
                int cv;
                if constexpr (std::is_integral_v<decltype(C)>) cv = C; else cv = decltype(C)::value;
                sts_f64(prow + 8u * (unsigned)cv, cs);
            } else if constexpr (!std::is_integral_v<decltype(C)>) {
                constexpr int kC = decltype(C)::value;
                sts_f64(pcol + 8u * (unsigned)(kWipPanelLd * kC), cs);
                if constexpr (kC < kWipKB - 1)
                    csr[kC] = cs;
            } else {
            sts_f64(pcol + 8u * (unsigned)(kWipPanelLd * C), cs);
            auto keep = [&](auto ctag) {
                constexpr int c = decltype(ctag)::value;
                if constexpr (c < kWipKB - 1)
                    csr[c] = cs;
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
            }
            fin = fin || lane == r;
            fmask |= 1u << r;
            __syncwarp();
            ++j;
        };

        if constexpr (kAlgo == kGMW81) {
            // ---- GMW81, symmetric form (see wsm_gmw.cuh): L entries are c_is / sqrt(d_s) ------
            // off-diagonal maximum (gmw81.rs:52-58) from the fragments, as integers (|.| of
            // non-NaN doubles order like their bits); diagonal-tile entries with row == column
            // are skipped
            if constexpr (kRegC)
#pragma unroll
            for (int bi = 0; bi < NB; ++bi)
#pragma unroll
                for (int bj = 0; bj <= bi; ++bj)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const unsigned long long a =
                            (unsigned long long)__double_as_longlong(cf[bi * (bi + 1) / 2 + bj][e]) & 0x7fffffffffffffffull;
                        const bool diag = bi == bj && 2 * c4 + e == r8;
                        om = (!diag && a > om) ? a : om;
                    }
            const double diag_max = warp_extreme<true>(kFull, fabs(dg), has, 0.0, win);
            const unsigned ohi = __reduce_max_sync(kFull, (unsigned)(om >> 32));
            const unsigned olo = __reduce_max_sync(kFull, (unsigned)(om >> 32) == ohi ? (unsigned)om : 0u);
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
                const double col = column(ctag, r);   // lane r: c_jj from the same updates
                const double cjj = __shfl_sync(kFull, col, r);
                const bool live = !fin && lane != r;
                // theta = max_{i>j} |c_ij| (gmw81.rs:87-100), d_j (gmw81.rs:102)
                const double theta = warp_extreme<true>(kFull, fabs(col), live, 0.0, win);
                const double tb = dmul(theta, rbeta);
                const double dj = fmax(fmax(fabs(cjj), dmul(tb, tb)), delta);
                double y = rsqrt_normal(dj);
                double sq = dmul(dj, y);
                if (!fast_range_ok(dj))
                    wip_slow_gmw(dj, sq, y);
                const double cs = fin ? 0.0 : dmul(col, y);         // finished rows receive exact zeros
                dg = __fma_rn(-cs, cs, dg);                         // gmw81.rs:104-109 (no mask needed)
                ev = lane == r ? dsub(dj, cjj) : ev;                // gmw81.rs:110
                retire(ctag, r, cs, sq);
            };
            // (one algorithm phase only: the eight unrolled copies of the step, one per count of
            // pending columns, fit the instruction cache and save the dispatch on the count)
            if constexpr (kWipRowPanel) {
                while (j < n) {
                    gmw_step(cnt);
                    ++cnt;
                    if (j < n && cnt == kWipKB) {
                        rank_update(kWipKB);
                        cnt = 0;
                    }
                }
            } else {
            while (j < n) {
                if (wip_steps<0, kWipKB>([&](auto ctag) {
                        gmw_step(ctag);
                        return j == n;
                    }))
                    break;
                rank_update(kWipKB);
            }
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
                const double col = column(ctag, r);
                const bool live = !fin && lane != r;
                const bool fast = fast_range_ok(piv);
                const double y = rsqrt_normal(piv);
                double sq = dmul(piv, y);
                double cs = dmul(col, y);
                double dgn = __fma_rn(-cs, cs, dg);
                // look-ahead value (se90.rs:70-77, se99.rs:82-89); FAST: the updated diagonal
                double nv = dgn;
                if (!fast)
                    wip_slow_p1(piv, col, dg, sq, cs, dgn, nv);
                const double la = warp_extreme<false>(kFull, nv, live, INFINITY, win);
                have_dmin = fast;
                dmin_carry = la;
                const double thresh = (kAlgo == kSE99) ? dmul(-MODCHOL_MU, gamma) : taugamma;
                if (la < thresh) {   // the pivot at j stays applied (se99.rs:90-93)
                    kstart = j;
                    return 1;
                }
                cs = fin ? 0.0 : cs;
                dg = dgn;   // (finished rows and the pivot row: never looked at again)
                retire(ctag, r, cs, sq);
                return j == n ? 2 : 0;
            };
            int st = 0;
            for (;;) {
                st = p1_step(cnt);   // cnt: pending panel columns at this step
                if (st != 0)
                    break;
                if (++cnt == kWipKB) {
                    rank_update(kWipKB);
                    cnt = 0;
                }
            }

            // ---- phase 2 -------------------------------------------------------------------
            if (st == 1 && kAlgo == kSE99 && kstart == n - 1) {   // se99.rs:115-119
                const int u = 31 - __clz(__ballot_sync(kFull, !fin));
                const double ljj = __shfl_sync(kFull, dg, u);
                const double ej = tail1x1_e(ljj, gamma);
                ev = lane == u ? ej : ev;
                wip_tmem_store(tw + 2u * (unsigned)(n - 1), lane == u ? dsqrt(dadd(ljj, ej)) : 0.0);
            } else if (st == 1) {
                if (j > 0) {
                    // the Gershgorin bounds need the trailing block as the reference holds it: apply
                    // the pending columns, then g_i = a_ii - sum_{k alive, k != i} |a_ik|
                    // (se90.rs:105-110, se99.rs:125-130)
                    if (cnt > 0) {
                        rank_update(cnt);
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
                    const double col = column(ctag, r);
                    const bool live = !fin && lane != r;
                    const double acol = live ? abs_bits(col) : 0.0;
                    // E_jj (se90.rs:124-131, se99.rs:144-151)
                    const double normj = MODCHOL_WIP_DMMA_SUM ? wip_sum32(acol) : warp_sum_tree<32>(kFull, acol);
                    const double ej = max_b_not_nan(dadd(-piv, max_b_not_nan(normj, taugamma)), delta_prev);
                    piv = dadd(piv, ej);
                    delta_prev = ej;
                    // bound update (se90.rs:134-139, se99.rs:154-159) and column scaling
                    const bool upd = fabs(dsub(piv, normj)) > MODCHOL_EPS;
                    const double y = rsqrt_normal(piv);
                    double sq = dmul(piv, y);
                    double t = __fma_rn(-normj, dmul(y, y), 1.0);
                    double cs = dmul(col, y);
                    if (!fast_range_ok(piv))
                        wip_slow_p2(piv, normj, col, sq, t, cs);
                    // finished rows receive exact zeros (their row of L in tensor memory ends at the
                    // diagonal, se99.rs:189-192); their diagonal and bound are never looked at again,
                    // so the updates below need no mask
                    cs = fin ? 0.0 : cs;
                    dg = __fma_rn(-cs, cs, dg);
                    g = upd ? __fma_rn(acol, t, g) : g;
                    ev = lane == r ? ej : ev;
                    retire(ctag, r, cs, sq);
                };
                while (j + 2 < n) {
                    p2_step(cnt);
                    if (++cnt == kWipKB) {
                        rank_update(kWipKB);
                        cnt = 0;
                    }
                }
                // final 2x2 block, never pivoted (se90.rs:155-166, se99.rs:175-186): the two rows
                // alive sit at positions n-2 and n-1
                const int u0 = 31 - __clz(__ballot_sync(kFull, !fin && pos == n - 2));
                const int u1 = 31 - __clz(__ballot_sync(kFull, !fin && pos == n - 1));
                double a00 = __shfl_sync(kFull, dg, u0);
                double a11 = __shfl_sync(kFull, dg, u1);
                const double c2 = column_rt(cnt, u0);
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
                ev = (lane == u0 || lane == u1) ? en : ev;
                wip_tmem_store(tw + 2u * (unsigned)(n - 2), lane == u0 ? a00 : (lane == u1 ? a10 : 0.0));
                wip_tmem_store(tw + 2u * (unsigned)(n - 1), lane == u1 ? a11 : 0.0);
            }
        }

        // ---- epilogue: this lane's row of L leaves tensor memory for row `pos` of the output,
        // zero beyond the diagonal (se99.rs:189-192); e in original order (se99.rs:194-198); p ---
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        if constexpr (kSolve)
            __syncwarp();   // every lane is done reading C from the triangle
        if (L || kSolve) {
            double* dstp = L + (size_t)b * nn + (size_t)pos * n;
            const unsigned rowb = wb + 8u * (unsigned)tri(pos);   // row pos of the position-ordered triangle
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                if (8 * c >= n)
                    break;
                double v[8];
                wip_tmem_load8(tw + 16u * c, v);
                // (entries beyond the diagonal were stored as zeros: a finished row receives 0)
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
                        sts_f64_if(has && 8 * c + u <= pos, rowb + 8u * (unsigned)(8 * c + u), v[u]);
                }
            }
        }
        if (has) {
            if (E)
                E[(size_t)b * n + lane] = ev;
            if (P)
                P[(size_t)b * n + pos] = (unsigned long long)lane;
        }
        if (K && lane == 0)
            K[b] = kstart;
        if constexpr (kSolve) {
            // lane -> position: the original index sitting at position `lane` through the panel
            // area (free now), then L z = P rhs, L^T w = z, x[p_i] = w_i on the triangle
            sts_f64(pnb + 8u * (unsigned)pos, (double)lane);
            __syncwarp();
            const unsigned tibs[1] = {tib};
            const int rows[1] = {lane};
            const bool rvs[1] = {has};
            const int perms[1] = {(int)lds_f64(pnb + 8u * (unsigned)lane)};
            wsm_solve_in_place<32, 1>(kFull, n, wb, tibs, rows, rvs, perms, Bm + (size_t)b * nrhs * n,
                                      X + (size_t)b * nrhs * n, nrhs);
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (wib == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                     :: "r"(tmem_base), "n"(kWipTmemCols) : "memory");
}

}  // namespace modchol
