// bsm_kernels.cuh -- FAST-mode blocked (panel) kernels for the LARGE size class, 128 < n <= 1024:
// one thread block per matrix, thread = permuted position (row), physical pivoting, the working
// matrix resident in L2 / HBM and the current column PANEL in shared memory.
//
//   for each panel of KB columns [j0, j0 + KB):
//       for j in the panel (left-looking inside the panel, gmw81.rs:70-111 / se9x phase loops):
//           pivot search on the per-row scalars (running diagonal / Gershgorin bound) -- the
//               scalars are brought up to date at every step, so every pivot, theta, norm and
//               E decision sees current values although the trailing matrix is only updated at
//               panel boundaries;
//           symmetric exchange j <-> m: finished rows of L, panel rows, trailing matrix, scalars;
//           column j = T(j.., j) - P(j.., 0..c) P(j, 0..c)^T   (T: one coalesced row of the
//               transposed trailing storage, brought in by ONE bulk-copy (TMA) instruction and an
//               mbarrier; P: shared memory, conflict free);
//           scale, append as panel column c, update the scalars
//       write the panel's rows of L (64 / 128 / 256-byte row segments)
//       trailing update T -= P P^T of the shrinking trailing block on the FP64 tensor pipe
//           (mma.sync.m8n8k4.f64, KB/4 per 8x8 tile), tiles streamed from L2 with 128-bit
//           accesses, every warp of the block on its own block rows.
//
// Storage: the working matrix lives IN PLACE in the caller's output buffer -- no scratch.  The
// strictly lower trailing part T(i,k), i > k, is kept TRANSPOSED in the strict upper triangle,
// l[k n + i] (the symmetric input delivers it as is: a[k][i] = a[i][k]), so that a pivot column
// is a contiguous row segment; its diagonal lives in shared memory; finished L entries go to
// the lower triangle (final place).  The strict upper triangle is zeroed at the end
// (gmw81.rs:112-125 / se99.rs:189-192).
//
// Flops: n^3/3 (the trailing block shrinks, unlike the implicit-pivoting small-n kernels whose
// fixed tiles would cost 3x here); L2/HBM traffic of the updates: ~ n^3 / (3 KB) * 16 bytes.
#pragma once
#include "wip_kernels.cuh"

namespace modchol {

constexpr int kBsmMaxWarps = 32;
// doubles of dynamic shared memory: scratch + diagonal + column line + panel
constexpr unsigned kBsmScratchBytes = 4u * 192u + 8u * 64u + 8u * 8u + 16u + 48u;   // reductions, scalars, mbarrier; 1 408

// panel column stride (doubles): >= n rounded up to 8, == 4 (mod 16): the per-step column
// write (thread = row) and the DMMA fragment loads are both conflict free
__host__ __device__ inline int bsm_ldp(int n)
{
    int l = (n + 7) & ~7;
    while ((l & 15) != 4)
        l += 4;
    return l;
}
// doubles of dynamic shared memory: panel + diagonal + column line + exchange scratch + mbarrier
__host__ __device__ inline size_t bsm_smem_doubles(int n, int kb)
{
    return (size_t)kb * bsm_ldp(n) + 2 * (size_t)((n + 7) & ~7) + kBsmScratchBytes / 8;
}

__device__ __forceinline__ void sts_u32(unsigned a, unsigned v)
{
    smem_check(a, 4u);
    asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned lds_u32(unsigned a)
{
    smem_check(a, 4u);
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}

struct BsmScratch {
    unsigned base;   // byte address: [2][3][32] words (hi, lo, pos), then [2][32] doubles, then 8 doubles
    int buf;
};

// first maximum by position of v over the candidate threads (utils.rs:52-91): position of the
// winner (`lo` when there is no candidate); ONE __syncthreads
__device__ __forceinline__ int bsm_argmax_first(BsmScratch& sc, int nwarps, double v, bool cand, int pos, int lo,
                                                double& best)
{
    int khi;
    unsigned klo;
    wip_key(v, khi, klo);
    cand = cand && v == v;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wh = __reduce_max_sync(kFull, cand ? khi : (int)0x80000000);
    const bool c1 = cand && khi == wh;
    const unsigned wl = __reduce_max_sync(kFull, c1 ? klo : 0u);
    const unsigned wp = __reduce_min_sync(kFull, (c1 && klo == wl) ? (unsigned)pos : 0xffffu);
    const unsigned a = sc.base + 4u * (unsigned)(sc.buf * 96);
    if (lane == 0) {
        sts_u32(a + 4u * warp, (unsigned)wh);
        sts_u32(a + 4u * (32 + warp), wl);
        sts_u32(a + 4u * (64 + warp), wp);
    }
    __syncthreads();
    const bool in = lane < nwarps;
    const int h = in ? (int)lds_u32(a + 4u * lane) : (int)0x80000000;
    const unsigned l = in ? lds_u32(a + 4u * (32 + lane)) : 0u;
    const unsigned p = in ? lds_u32(a + 4u * (64 + lane)) : 0xffffu;
    const int gh = __reduce_max_sync(kFull, h);
    const bool g1 = h == gh;
    const unsigned gl = __reduce_max_sync(kFull, g1 ? l : 0u);
    const unsigned gp = __reduce_min_sync(kFull, (g1 && l == gl) ? p : 0xffffu);
    sc.buf ^= 1;
    const int m = gh;
    best = (m == (int)0x80000000) ? -INFINITY
                                  : __hiloint2double(m ^ ((m >> 31) & 0x7fffffff), (int)(gl ^ (unsigned)(m >> 31)));
    return (gh == (int)0x80000000 || gp >= 0xffffu) ? lo : (int)gp;
}

// block-wide maximum (kMax) / minimum of v over the candidate threads, NaN never taken, `none`
// without candidates; ONE __syncthreads
template <bool kMax>
__device__ __forceinline__ double bsm_extreme(BsmScratch& sc, int nwarps, double v, bool cand, double none)
{
    int khi;
    unsigned klo;
    wip_key(v, khi, klo);
    cand = cand && v == v;
    if (!kMax) {
        khi = ~khi;
        klo = ~klo;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wh = __reduce_max_sync(kFull, cand ? khi : (int)0x80000000);
    const unsigned wl = __reduce_max_sync(kFull, (cand && khi == wh) ? klo : 0u);
    const unsigned a = sc.base + 4u * (unsigned)(sc.buf * 96);
    if (lane == 0) {
        sts_u32(a + 4u * warp, (unsigned)wh);
        sts_u32(a + 4u * (32 + warp), wl);
    }
    __syncthreads();
    const bool in = lane < nwarps;
    const int h = in ? (int)lds_u32(a + 4u * lane) : (int)0x80000000;
    const unsigned l = in ? lds_u32(a + 4u * (32 + lane)) : 0u;
    int gh = __reduce_max_sync(kFull, h);
    unsigned gl = __reduce_max_sync(kFull, h == gh ? l : 0u);
    sc.buf ^= 1;
    if (gh == (int)0x80000000)
        return none;
    if (!kMax) {
        gh = ~gh;
        gl = ~gl;
    }
    return __hiloint2double(gh ^ ((gh >> 31) & 0x7fffffff), (int)(gl ^ (unsigned)(gh >> 31)));
}

// block-wide sum (tree order inside the warps, DMMA across them); ONE __syncthreads
__device__ __forceinline__ double bsm_sum(BsmScratch& sc, int nwarps, double x)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    x = wip_sum32(x);
    const unsigned a = sc.base + 4u * 192u + 8u * (unsigned)(sc.buf * 32);
    if (lane == 0)
        sts_f64(a + 8u * warp, x);
    __syncthreads();
    const double y = lane < nwarps ? lds_f64(a + 8u * lane) : 0.0;
    sc.buf ^= 1;
    return wip_sum32(y);
}

// ---- mbarrier + bulk copy (TMA, non-tensor form) ------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned mbar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity)
{
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        " @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}"
        :: "r"(mbar), "r"(parity) : "memory");
}
// generic-proxy writes (st.global / st.shared) before, async-proxy (bulk copy) accesses after
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async;" ::: "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned mbar)
{
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}

__device__ __forceinline__ void bulk_s2g(void* dst, unsigned src, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// A -> working buffer by the bulk-copy engine (TMA): warp w stages storage row k (columns
// c0 = (k+1) & ~1 .. n: the strict upper triangle, from an even column so that the copy is 16-byte
// aligned) in its slot of the idle panel -- cp.async.bulk global -> shared, completion on the
// warp's mbarrier --, takes the off-diagonal maximum from it (kMax, gmw81.rs:52-58) and sends it on
// with cp.async.bulk shared -> global.  `slots` rows are in flight per block.  The extra element
// copied when k + 1 is odd is the diagonal slot l[k n + k], which is written later anyway.
template <bool kMax>
__device__ __forceinline__ double bsm_copy_in_tma(const double* __restrict__ Ab, double* __restrict__ Lb, int n, int npad,
                                                  unsigned pb, int slots, unsigned mbars, unsigned& tphase)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double om = 0.0;
    if (warp < slots) {
        const unsigned slot = pb + 8u * (unsigned)(warp * npad);
        const unsigned mb = mbars + 8u * (unsigned)warp;
        for (int k = warp; k < n - 1; k += slots) {
            const int c0 = (k + 1) & ~1;
            const unsigned bytes = 8u * (unsigned)(n - c0);
            if (lane == 0) {
                mbar_expect_tx(mb, bytes);
                bulk_g2s(slot, Ab + (size_t)k * n + c0, bytes, mb);
            }
            mbar_wait(mb, tphase);
            tphase ^= 1u;
            if (kMax)
                for (int i = k + 1 + lane; i < n; i += 32)
                    om = fmax(om, fabs(lds_f64(slot + 8u * (unsigned)(i - c0))));
            __syncwarp();
            if (lane == 0) {
                bulk_s2g(Lb + (size_t)k * n + c0, slot, bytes);
                bulk_commit();
                bulk_wait_read0();   // the slot may be overwritten
            }
            __syncwarp();
        }
        if (lane == 0)
            bulk_wait0();            // written: the block's barrier that follows publishes it
    }
    return om;
}

// strict upper triangle := 0 by bulk stores of a zero line (the first n doubles of the panel)
__device__ __forceinline__ void bsm_zero_upper_tma(double* __restrict__ Lb, int n, int npad, unsigned pb)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int q = threadIdx.x; q < npad; q += blockDim.x)
        sts_f64(pb + 8u * (unsigned)q, 0.0);
    fence_proxy_async();
    __syncthreads();
    if (lane == 0) {
        for (int k = warp; k < n - 1; k += nwarps) {
            int s0 = k + 1;
            if (s0 & 1) {
                Lb[(size_t)k * n + s0] = 0.0;
                ++s0;
            }
            if (s0 < n)
                bulk_s2g(Lb + (size_t)k * n + s0, pb, 8u * (unsigned)(n - s0));
        }
        bulk_commit();
        bulk_wait_read0();   // the zero line may be overwritten (the stores complete on their own)
    }
}

// Trailing update T(i,k) -= sum_c P(i,c) P(k,c) for j1 <= k < i < n, T(i,k) stored at Lb[k n + i],
// and the same for the diagonal td[i] (shared memory).  Tile (bj, bi), bi >= bj, is held
// transposed: fragment rows = k in block bj (A fragment: lane l holds P(8 bj + l/4, 4 h + l%4)),
// fragment columns = i in block bi (B fragment: P(8 bi + l/4, 4 h + l%4)), accumulator
// C[l/4][2 (l%4) + {0,1}] = T(8 bi + 2 (l%4) + {0,1}, 8 bj + l/4): two doubles contiguous in i.
// Each warp takes the block rows bi = b0 + warp, + nwarps, ... (cyclic: balanced to within one
// row), keeps the row's B fragments in registers and streams the tiles with 128-bit accesses.
template <int KB>
__device__ __forceinline__ void bsm_trailing_update(double* __restrict__ Lb, int n, int j1, unsigned pb, int ldp,
                                                    unsigned tdb, bool wide)
{
    constexpr int KC = KB / 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int r8 = lane >> 2, c4 = lane & 3;
    const int b0 = j1 >> 3, nb = (n + 7) >> 3;
    const unsigned pf = pb + 8u * (unsigned)(c4 * ldp + r8);   // P(r8, c4)
    for (int bi = b0 + warp; bi < nb; bi += nwarps) {
        double fb[KC];
#pragma unroll
        for (int h = 0; h < KC; ++h)
            fb[h] = lds_f64(pf + 8u * (unsigned)(4 * h * ldp + 8 * bi));
        const int i0 = 8 * bi + 2 * c4;
        const bool full_i = 8 * bi + 7 < n;
        int bj = b0;
        // interior tiles (bj < bi, all rows and columns inside the matrix), four at a time: the
        // four 128-bit loads are in flight together, so a warp pays one L2 round trip per group
        if (full_i && wide) {
            for (; bj + 4 <= bi; bj += 4) {
                double2 v[4];
                double* tp = Lb + (size_t)(8 * bj + r8) * n + i0;
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    v[u] = *reinterpret_cast<const double2*>(tp + (size_t)(8 * u) * n);
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int h = 0; h < KC; ++h) {
                        const double fa = wip_neg(lds_f64(pf + 8u * (unsigned)(4 * h * ldp + 8 * (bj + u))));
                        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                                     : "+d"(v[u].x), "+d"(v[u].y) : "d"(fa), "d"(fb[h]));
                    }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    *reinterpret_cast<double2*>(tp + (size_t)(8 * u) * n) = v[u];
            }
        }
        for (; bj <= bi; ++bj) {
            const int k = 8 * bj + r8;
            double* tp = Lb + (size_t)k * n + i0;
            const bool interior = bj < bi && full_i;   // warp-uniform; (bj < bi => k-block is full too)
            double c0, c1;
            if (interior && wide) {
                const double2 v = *reinterpret_cast<const double2*>(tp);
                c0 = v.x;
                c1 = v.y;
            } else {
                c0 = (k < n && i0 < n && i0 > k) ? tp[0] : 0.0;
                c1 = (k < n && i0 + 1 < n && i0 + 1 > k) ? tp[1] : 0.0;
                if (bj == bi) {   // the diagonal entries of T live in shared memory
                    if (i0 == k && k < n) c0 = lds_f64(tdb + 8u * (unsigned)k);
                    if (i0 + 1 == k && k < n) c1 = lds_f64(tdb + 8u * (unsigned)k);
                }
            }
#pragma unroll
            for (int h = 0; h < KC; ++h) {
                const double fa = wip_neg(lds_f64(pf + 8u * (unsigned)(4 * h * ldp + 8 * bj)));
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                             : "+d"(c0), "+d"(c1) : "d"(fa), "d"(fb[h]));
            }
            if (interior && wide) {
                *reinterpret_cast<double2*>(tp) = make_double2(c0, c1);
            } else {
                if (k < n && i0 < n && i0 > k) tp[0] = c0;
                if (k < n && i0 + 1 < n && i0 + 1 > k) tp[1] = c1;
                if (bj == bi) {
                    if (i0 == k && k < n) sts_f64(tdb + 8u * (unsigned)k, c0);
                    if (i0 + 1 == k && k < n) sts_f64(tdb + 8u * (unsigned)k, c1);
                }
            }
        }
    }
}

// GMW81 (gmw81.rs:39-134), symmetric FAST form (see wsm_gmw.cuh): L entries are c_is / sqrt(d_s).
template <int kThreads, int KB, int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
bsm_gmw_kernel(int n, long long batch, const double* __restrict__ A, double* __restrict__ L,
               double* __restrict__ E, unsigned long long* __restrict__ P, int* __restrict__ K)
{
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int nwarps = kThreads >> 5;
    const int ldp = bsm_ldp(n);
    const int npad = (n + 7) & ~7;
    const unsigned smb = (unsigned)__cvta_generic_to_shared(sm);
    // fixed-size scratch first (its addresses are the block's base plus constants: nothing to
    // recompute inside the loops), then the diagonal, the column line and the panel
    BsmScratch sc;
    sc.base = smb;
    sc.buf = 0;
    const unsigned scalb = smb + 4u * 192u + 8u * 64u;              // 8 doubles of scalars
    const unsigned mbar = scalb + 8u * 8u;                          // one mbarrier
    const unsigned tdb = smb + kBsmScratchBytes;                    // diagonal of T, by position
    const unsigned lineb = tdb + 8u * (unsigned)npad;               // pivot column line
    const unsigned pb = lineb + 8u * (unsigned)npad;                // panel: column c at pb + 8 c ldp
    const size_t nn = (size_t)n * n;
    const int row = tid;
    const bool rv = row < n;
    // 128-bit tile accesses / bulk copies need 16-byte aligned rows
    const bool wide = (n & 1) == 0 && (((size_t)L) & 15) == 0;
    // bulk copies (TMA) need 16-byte aligned rows of both buffers; one mbarrier per warp in the
    // (otherwise unused) column line
    const bool tma = wide && (((size_t)A) & 15) == 0;
    const int slots = min(nwarps, (KB * ldp) / npad);
    if (lane == 0)
        mbar_init(lineb + 8u * (unsigned)(tid >> 5), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    unsigned tphase = 0;
    __syncthreads();

    for (long long b = blockIdx.x; b < batch; b += gridDim.x) {
        const double* Ab = A + (size_t)b * nn;
        double* Lb = L + (size_t)b * nn;
        // ---- in: the strict upper triangle of A is T transposed; diagonal to shared memory;
        // off-diagonal maximum on the fly (gmw81.rs:52-58) ------------------------------------
        double om = 0.0;
        if (tma) {
            om = bsm_copy_in_tma<true>(Ab, Lb, n, npad, pb, slots, lineb, tphase);
        } else {
            for (int k = tid >> 5; k < n; k += nwarps) {   // one row per warp: coalesced
                for (int i = k + 1 + lane; i < n; i += 32) {
                    const double v = __ldg(Ab + (size_t)k * n + i);
                    Lb[(size_t)k * n + i] = v;
                    om = fmax(om, fabs(v));
                }
            }
        }
        double cd = rv ? __ldg(Ab + (size_t)row * (n + 1)) : 0.0;   // running diagonal (gmw81.rs:66-67)
        if (row < npad)
            sts_f64(tdb + 8u * (unsigned)row, cd);
        double ev = 0.0;
        int perm = row;
        const double diag_max = bsm_extreme<true>(sc, nwarps, fabs(cd), rv, 0.0);
        const double off_max = bsm_extreme<true>(sc, nwarps, om, true, 0.0);
        const double nf = (double)n;
        const double delta = dmul(MODCHOL_EPS, fmax(1.0, dadd(diag_max, off_max)));
        const double beta =
            dsqrt(fmax(fmax(diag_max, ddiv(off_max, dsqrt(dsub(dmul(nf, nf), 1.0)))), MODCHOL_EPS));
        const double rbeta = ddiv(1.0, beta);
        fence_proxy_async();
        __syncthreads();   // T is in place (block-scope visibility of the global stores)

        for (int j0 = 0; j0 < n; j0 += KB) {
            const int jend = min(n, j0 + KB);
            // panel rows of finished positions and columns not yet produced must read as zero in
            // the trailing update: clear the panel
            for (int q = tid; q < KB * ldp; q += kThreads)
                sts_f64(pb + 8u * (unsigned)q, 0.0);
            __syncthreads();
            for (int j = j0; j < jend; ++j) {
                const int c = j - j0;
                // ---- pivot on max |c_ii| (gmw81.rs:71-78) ----------------------------------------
                double best;
                const int m = bsm_argmax_first(sc, nwarps, fabs(cd), rv && row >= j, row, j, best);
                // ---- symmetric exchange j <-> m (utils.rs:30-49) fused with the read of column j:
                // the exchange loads T(i,j) and T(i,m) anyway and the new column j IS the old column
                // m, so a step costs ONE round trip to L2; the stores are not waited for (the next
                // reader of those locations sits behind a barrier).  Nothing in shared memory moves
                // before the column is evaluated -- rows j and m of the panel and of the diagonal
                // are read through `prow` / `pj` -- so the exchange needs no barrier of its own.
                const bool swp = m != j;
                const int pj = m;                                   // old place of the new row j
                const int prow = (swp && row == j) ? m : ((swp && row == m) ? j : row);
                double col = 0.0, x = 0.0, y = 0.0, lx = 0.0, ly = 0.0;
                double *a1 = nullptr, *a2 = nullptr;
                const bool tsw = swp && rv && row > j && row != m;
                const bool lsw = swp && tid < j0;                   // L(j, 0..j0) <-> L(m, 0..j0)
                if (tsw) {                                          // T(i,j) <-> T(m,i) | T(i,m)
                    a1 = Lb + (size_t)j * n + row;
                    a2 = row < m ? Lb + (size_t)row * n + m : Lb + (size_t)m * n + row;
                    x = *a1;
                    y = *a2;
                    col = y;
                } else if (rv && row > j) {
                    col = Lb[(size_t)j * n + row];                  // no exchange, or T(m,j), which stays
                }
                if (lsw) {
                    lx = Lb[(size_t)j * n + tid];
                    ly = Lb[(size_t)m * n + tid];
                }
                if (swp && (row == j || row == m)) {                // scalars through shared memory
                    const unsigned q = scalb + (row == j ? 0u : 16u);
                    sts_f64(q, cd);
                    sts_f64(q + 8, (double)perm);
                }
                if (row == j)
                    col = lds_f64(tdb + 8u * (unsigned)pj);         // c_jj from the same updates
                // left-looking inside the panel: minus P(i, 0..c) . P(j, 0..c), accumulated apart
                // from the loaded element so that the products overlap the round trip to L2
                if (rv && row >= j) {
                    unsigned pr = pb + 8u * (unsigned)prow;         // panel column t at pb + 8 t ldp
                    unsigned pq = pb + 8u * (unsigned)pj;
                    const unsigned st8 = 8u * (unsigned)ldp;
                    double s0 = 0.0, s1 = 0.0;
                    int t = 0;
                    for (; t + 4 <= c; t += 4, pr += 4u * st8, pq += 4u * st8) {
                        const double r0 = lds_f64(pr), r1 = lds_f64(pr + st8);
                        const double r2 = lds_f64(pr + 2u * st8), r3 = lds_f64(pr + 3u * st8);
                        const double q0 = lds_f64(pq), q1 = lds_f64(pq + st8);
                        const double q2 = lds_f64(pq + 2u * st8), q3 = lds_f64(pq + 3u * st8);
                        s0 = __fma_rn(r0, q0, s0);
                        s1 = __fma_rn(r1, q1, s1);
                        s0 = __fma_rn(r2, q2, s0);
                        s1 = __fma_rn(r3, q3, s1);
                    }
                    for (; t < c; ++t, pr += st8, pq += st8)
                        s0 = __fma_rn(lds_f64(pr), lds_f64(pq), s0);
                    col = dsub(col, dadd(s0, s1));
                }
                if (row == j)
                    sts_f64(scalb + 32u, col);
                if (tsw) {
                    *a1 = y;
                    *a2 = x;
                }
                if (lsw) {
                    Lb[(size_t)j * n + tid] = ly;
                    Lb[(size_t)m * n + tid] = lx;
                }
                // theta = max_{i>j} |c_ij| (gmw81.rs:87-100); c_jj rides on the same barrier
                const double theta = bsm_extreme<true>(sc, nwarps, fabs(col), rv && row > j, 0.0);
                const double cjj = lds_f64(scalb + 32u);
                if (swp) {
                    // now the shared-memory side of the exchange (every reader of the old places is
                    // past the barrier): scalars, panel rows j and m, diagonal of T
                    if (row == j || row == m) {
                        const unsigned q = scalb + (row == j ? 16u : 0u);
                        cd = lds_f64(q);
                        perm = (int)lds_f64(q + 8);
                    }
                    if (tid < c) {
                        const unsigned q = pb + 8u * (unsigned)(tid * ldp);
                        const double u = lds_f64(q + 8u * j), w = lds_f64(q + 8u * m);
                        sts_f64(q + 8u * j, w);
                        sts_f64(q + 8u * m, u);
                    }
                    if (tid == 32) {
                        const double u = lds_f64(tdb + 8u * j), w = lds_f64(tdb + 8u * m);
                        sts_f64(tdb + 8u * j, w);
                        sts_f64(tdb + 8u * m, u);
                    }
                }
                const double tb = dmul(theta, rbeta);
                // gmw81.rs:102; delta and (theta/beta)^2 are never NaN here, so `a > b ? a : b` with the
                // possibly-NaN value on the left is f64::max at a third of the instructions
                const double dj = max_b_not_nan(max_b_not_nan(fabs(cjj), dmul(tb, tb)), delta);
                double yy = rsqrt_normal(dj);
                double sq = dmul(dj, yy);
                if (!fast_range_ok(dj))
                    wip_slow_gmw(dj, sq, yy);
                const double ls = dmul(col, yy);
                if (rv && row >= j)
                    sts_f64(pb + 8u * (unsigned)(c * ldp + row), row == j ? sq : ls);
                if (rv && row > j)
                    cd = __fma_rn(-ls, ls, cd);                     // gmw81.rs:104-109
                if (row == j)
                    ev = dsub(dj, cjj);                             // gmw81.rs:110
                // no barrier here: the next step's pivot search has one between these writes and
                // their first readers
            }
            __syncthreads();
            // ---- the panel's columns of L: row i gets L(i, j0 .. min(i, jend-1)) ----------------
            if (rv && row >= j0) {
                const int cnt = min(jend - j0, row - j0 + 1);
                double* dst = Lb + (size_t)row * n + j0;
                for (int t = 0; t < cnt; ++t)
                    dst[t] = lds_f64(pb + 8u * (unsigned)(t * ldp + row));
            }
            // (rows inside the panel are finished and lie outside the blocks the update touches:
            // jend is a multiple of 8)
            if (jend < n)
                bsm_trailing_update<KB>(Lb, n, jend, pb, ldp, tdb, wide);
            fence_proxy_async();
            __syncthreads();
        }

        // ---- out: strict upper triangle zero (gmw81.rs:112-125), e un-permuted, p ---------------
        if (tma) {
            bsm_zero_upper_tma(Lb, n, npad, pb);
        } else {
            for (int k = tid >> 5; k < n; k += nwarps)
                for (int i = k + 1 + lane; i < n; i += 32)
                    Lb[(size_t)k * n + i] = 0.0;
        }
        if (rv) {
            E[(size_t)b * n + perm] = ev;                               // gmw81.rs:127-131
            P[(size_t)b * n + row] = (unsigned long long)perm;
        }
        if (K && tid == 0)
            K[b] = -1;
        __syncthreads();
    }
    if (tma && lane == 0)
        bulk_wait0();   // the last matrix's zero-fill
}


// SE90 / SE99 (se90.rs:38-181, se99.rs:35-201) in the same blocked form: thread = permuted position,
// running diagonal and Gershgorin bound in registers, FAST arithmetic of wip_kernels.cuh.  A phase
// switch ends the current panel early (its columns are written out and applied to the trailing
// block), so that the Gershgorin pass (se90.rs:105-110, se99.rs:125-130) reads the trailing
// block as the reference holds it; later panels start at the switch column.
template <int kAlgo, int kThreads, int KB, int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
bsm_se_kernel(int n, long long batch, const double* __restrict__ A, double* __restrict__ L,
              double* __restrict__ E, unsigned long long* __restrict__ P, int* __restrict__ K)
{
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int nwarps = kThreads >> 5;
    const int ldp = bsm_ldp(n);
    const int npad = (n + 7) & ~7;
    const unsigned smb = (unsigned)__cvta_generic_to_shared(sm);
    BsmScratch sc;                                                  // layout: see bsm_gmw_kernel
    sc.base = smb;
    sc.buf = 0;
    const unsigned scalb = smb + 4u * 192u + 8u * 64u;              // 8 doubles of scalars
    const unsigned tdb = smb + kBsmScratchBytes;                    // diagonal of T (trailing-update bookkeeping)
    const unsigned lineb = tdb + 8u * (unsigned)npad;
    const unsigned pb = lineb + 8u * (unsigned)npad;                // panel: column c at pb + 8 c ldp
    const size_t nn = (size_t)n * n;
    const int row = tid;
    const bool rv = row < n;
    const bool wide = (n & 1) == 0 && (((size_t)L) & 15) == 0;
    const bool tma = wide && (((size_t)A) & 15) == 0;   // see bsm_gmw_kernel
    const int slots = min(nwarps, (KB * ldp) / npad);
    if (lane == 0)
        mbar_init(lineb + 8u * (unsigned)(tid >> 5), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    unsigned tphase = 0;
    __syncthreads();

    for (long long b = blockIdx.x; b < batch; b += gridDim.x) {
        const double* Ab = A + (size_t)b * nn;
        double* Lb = L + (size_t)b * nn;
        // ---- in: the strict upper triangle of A is T transposed; diagonal to registers -----------
        if (tma) {
            bsm_copy_in_tma<false>(Ab, Lb, n, npad, pb, slots, lineb, tphase);
        } else {
            for (int k = tid >> 5; k < n; k += nwarps)
                for (int i = k + 1 + lane; i < n; i += 32)
                    Lb[(size_t)k * n + i] = __ldg(Ab + (size_t)k * n + i);
        }
        double cd = rv ? __ldg(Ab + (size_t)row * (n + 1)) : 0.0;   // running diagonal
        if (row < npad)
            sts_f64(tdb + 8u * (unsigned)row, cd);
        double ev = 0.0, g = 0.0;
        int perm = row;
        // gamma = max |a_ii| folded from 0.0 (se90.rs:55-57, se99.rs:54-56)
        const double gamma = bsm_extreme<true>(sc, nwarps, fabs(cd), rv, 0.0);
        const double taugamma = dmul(MODCHOL_TAU, gamma);
        __syncthreads();   // T is in place (block-scope visibility of the global stores)

        int j = 0, j0 = 0, c = 0, kstart = -1, st = 0;
        double delta_prev = 0.0, dmin_carry = 0.0;
        bool have_dmin = false;

        // symmetric exchange j <-> m (utils.rs:30-49), global side, fused with the read of column j,
        // then minus the pending panel columns; the shared-memory side follows the step's reduction
        // barrier (finish_exchange), see bsm_gmw_kernel
        auto exchange_and_column = [&](int m) -> double {
            const bool swp = m != j;
            const int pj = m;
            const int prow = (swp && row == j) ? m : ((swp && row == m) ? j : row);
            double col = 0.0, x = 0.0, y = 0.0, lx = 0.0, ly = 0.0;
            double *a1 = nullptr, *a2 = nullptr;
            const bool tsw = swp && rv && row > j && row != m;
            const bool lsw = swp && tid < j0;
            if (tsw) {
                a1 = Lb + (size_t)j * n + row;
                a2 = row < m ? Lb + (size_t)row * n + m : Lb + (size_t)m * n + row;
                x = *a1;
                y = *a2;
                col = y;
            } else if (rv && row > j) {
                col = Lb[(size_t)j * n + row];
            }
            if (lsw) {
                lx = Lb[(size_t)j * n + tid];
                ly = Lb[(size_t)m * n + tid];
            }
            if (swp && (row == j || row == m)) {                // scalars through shared memory
                const unsigned q = scalb + (row == j ? 0u : 24u);
                sts_f64(q, cd);
                sts_f64(q + 8, (double)perm);
                sts_f64(q + 16, g);
            }
            if (rv && row > j) {
                unsigned pr = pb + 8u * (unsigned)prow;
                unsigned pq = pb + 8u * (unsigned)pj;
                const unsigned st8 = 8u * (unsigned)ldp;
                double s0 = 0.0, s1 = 0.0;
                int t = 0;
                for (; t + 4 <= c; t += 4, pr += 4u * st8, pq += 4u * st8) {
                    const double r0 = lds_f64(pr), r1 = lds_f64(pr + st8);
                    const double r2 = lds_f64(pr + 2u * st8), r3 = lds_f64(pr + 3u * st8);
                    const double q0 = lds_f64(pq), q1 = lds_f64(pq + st8);
                    const double q2 = lds_f64(pq + 2u * st8), q3 = lds_f64(pq + 3u * st8);
                    s0 = __fma_rn(r0, q0, s0);
                    s1 = __fma_rn(r1, q1, s1);
                    s0 = __fma_rn(r2, q2, s0);
                    s1 = __fma_rn(r3, q3, s1);
                }
                for (; t < c; ++t, pr += st8, pq += st8)
                    s0 = __fma_rn(lds_f64(pr), lds_f64(pq), s0);
                col = dsub(col, dadd(s0, s1));
            }
            if (tsw) {
                *a1 = y;
                *a2 = x;
            }
            if (lsw) {
                Lb[(size_t)j * n + tid] = ly;
                Lb[(size_t)m * n + tid] = lx;
            }
            return col;
        };
        auto finish_exchange = [&](int m) {
            if (m == j)
                return;
            if (row == j || row == m) {
                const unsigned q = scalb + (row == j ? 24u : 0u);
                cd = lds_f64(q);
                perm = (int)lds_f64(q + 8);
                g = lds_f64(q + 16);
            }
            if (tid < c) {
                const unsigned q = pb + 8u * (unsigned)(tid * ldp);
                const double u = lds_f64(q + 8u * j), w = lds_f64(q + 8u * m);
                sts_f64(q + 8u * j, w);
                sts_f64(q + 8u * m, u);
            }
        };
        auto clear_panel = [&]() {
            for (int q = tid; q < KB * ldp; q += kThreads)
                sts_f64(pb + 8u * (unsigned)q, 0.0);
            __syncthreads();
            j0 = j;
            c = 0;
        };
        // the panel's columns of L go to their final place; T -= P P^T on the trailing block
        auto flush = [&]() {
            __syncthreads();
            if (c > 0) {
                if (rv && row >= j0) {
                    const int cnt = min(c, row - j0 + 1);
                    double* dst = Lb + (size_t)row * n + j0;
                    for (int t = 0; t < cnt; ++t)
                        dst[t] = lds_f64(pb + 8u * (unsigned)(t * ldp + row));
                }
                if (j < n)
                    bsm_trailing_update<KB>(Lb, n, j, pb, ldp, tdb, wide);
            }
            __syncthreads();
        };

        // ---- phase 1 (se90.rs:61-96, se99.rs:61-109) ----------------------------------------------
        while (st == 0) {
            clear_panel();
            while (c < KB) {
                if (j == n) {
                    st = 2;
                    break;
                }
                double piv;
                if (row == j)
                    sts_f64(scalb + 56u, cd);   // diagonal at position j: row m takes it over below
                const int m = bsm_argmax_first(sc, nwarps, cd, rv && row >= j, row, j, piv);
                if (kAlgo == kSE99) {   // se99.rs:62-73, before the pivot is applied
                    const double dmin = have_dmin ? dmin_carry
                                                  : bsm_extreme<false>(sc, nwarps, cd, rv && row >= j, INFINITY);
                    if (piv < taugamma || dmin < dmul(-MODCHOL_MU, piv)) {
                        kstart = j;
                        st = 1;
                        break;
                    }
                }
                const double col = exchange_and_column(m);
                const bool live = rv && row > j;
                const bool fast = fast_range_ok(piv);
                const double y = rsqrt_normal(piv);
                double sq = dmul(piv, y);
                double cs = dmul(col, y);
                // the scalars of positions j and m trade places after the reduction barrier
                // (finish_exchange); row m already works with the diagonal it is about to receive
                const double cdo = (m != j && row == m) ? lds_f64(scalb + 56u) : cd;
                double dgn = __fma_rn(-cs, cs, cdo);
                double nv = dgn;   // look-ahead value (se90.rs:70-77, se99.rs:82-89); FAST: the updated diagonal
                if (!fast)
                    wip_slow_p1(piv, col, cdo, sq, cs, dgn, nv);
                const double la = bsm_extreme<false>(sc, nwarps, nv, live, INFINITY);
                finish_exchange(m);
                have_dmin = fast;
                dmin_carry = la;
                const double thresh = (kAlgo == kSE99) ? dmul(-MODCHOL_MU, gamma) : taugamma;
                if (la < thresh) {   // the pivot at j stays applied (se99.rs:90-93)
                    kstart = j;
                    st = 1;
                    break;
                }
                if (live)
                    cd = dgn;
                if (rv && row >= j)
                    sts_f64(pb + 8u * (unsigned)(c * ldp + row), row == j ? sq : cs);
                ++j;
                ++c;
            }
            flush();
        }

        if (st == 1 && kAlgo == kSE99 && kstart == n - 1) {   // se99.rs:115-119
            if (row == n - 1) {
                const double ej = tail1x1_e(cd, gamma);
                ev = ej;
                Lb[(size_t)(n - 1) * n + (n - 1)] = dsqrt(dadd(cd, ej));
            }
        } else if (st == 1) {
            // ---- phase 2 (se90.rs:101-167, se99.rs:121-187) ----------------------------------------
            // Gershgorin bounds of the trailing block (the pending columns were applied by flush()):
            // g_i = a_ii - sum_{k >= j, k != i} |a_ik|, T(i,k) at l[k n + i] for i > k
            if (rv && row >= j) {
                double s0 = 0.0, s1 = 0.0;
                int k = j;
                for (; k + 2 <= row; k += 2) {
                    s0 = dadd(s0, fabs(Lb[(size_t)k * n + row]));
                    s1 = dadd(s1, fabs(Lb[(size_t)(k + 1) * n + row]));
                }
                for (; k < row; ++k)
                    s0 = dadd(s0, fabs(Lb[(size_t)k * n + row]));
                const double* rp = Lb + (size_t)row * n;
                int i = row + 1;
                for (; i + 2 <= n; i += 2) {
                    s0 = dadd(s0, fabs(rp[i]));
                    s1 = dadd(s1, fabs(rp[i + 1]));
                }
                for (; i < n; ++i)
                    s0 = dadd(s0, fabs(rp[i]));
                g = dsub(cd, dadd(s0, s1));
            }
            while (j + 2 < n) {
                clear_panel();
                while (c < KB && j + 2 < n) {
                    double best;
                    const int m = bsm_argmax_first(sc, nwarps, g, rv && row >= j, row, j, best);   // se90.rs:115
                    if (row == m)
                        sts_f64(scalb + 48u, cd);   // the pivot's diagonal, read after the next barrier
                    const double col = exchange_and_column(m);
                    const bool live = rv && row > j;
                    const double acol = live ? abs_bits(col) : 0.0;
                    const double normj = bsm_sum(sc, nwarps, acol);   // se90.rs:124, se99.rs:144
                    double piv = lds_f64(scalb + 48u);
                    finish_exchange(m);
                    // E_jj (se90.rs:124-131, se99.rs:144-151)
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
                    if (live) {
                        cd = __fma_rn(-cs, cs, cd);
                        if (upd)
                            g = __fma_rn(acol, t, g);
                    }
                    if (row == j)
                        ev = ej;
                    if (rv && row >= j)
                        sts_f64(pb + 8u * (unsigned)(c * ldp + row), row == j ? sq : cs);
                    ++j;
                    ++c;
                }
                flush();
            }
            // final 2x2 block, never pivoted (se90.rs:155-166, se99.rs:175-186); the trailing block is
            // up to date
            if (row == n - 2)
                sts_f64(scalb + 48u, cd);
            if (row == n - 1)
                sts_f64(scalb + 56u, cd);
            __syncthreads();
            if (row == n - 2 || row == n - 1) {
                double a00 = lds_f64(scalb + 48u), a11 = lds_f64(scalb + 56u);
                double a10 = Lb[(size_t)(n - 2) * n + (n - 1)];
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
                ev = en;
                if (row == n - 2) {
                    Lb[(size_t)(n - 2) * n + (n - 2)] = a00;
                } else {
                    Lb[(size_t)(n - 1) * n + (n - 2)] = a10;
                    Lb[(size_t)(n - 1) * n + (n - 1)] = a11;
                }
            }
        }
        __syncthreads();

        // ---- out: strict upper triangle zero (se99.rs:189-192), e un-permuted (se99.rs:194-198), p
        if (tma) {
            bsm_zero_upper_tma(Lb, n, npad, pb);
        } else {
            for (int k = tid >> 5; k < n; k += nwarps)
                for (int i = k + 1 + lane; i < n; i += 32)
                    Lb[(size_t)k * n + i] = 0.0;
        }
        if (rv) {
            E[(size_t)b * n + perm] = ev;
            P[(size_t)b * n + row] = (unsigned long long)perm;
        }
        if (K && tid == 0)
            K[b] = kstart;
        __syncthreads();
    }
    if (tma && lane == 0)
        bulk_wait0();   // the last matrix's zero-fill
}

}  // namespace modchol
