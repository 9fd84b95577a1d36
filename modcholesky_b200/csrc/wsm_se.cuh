// wsm_se.cuh -- SE90 / SE99 (se90.rs:38-181, se99.rs:35-201) with the matrix resident in SHARED
// memory: G = 4/8/16/32 lanes per matrix (32 / G matrices per warp), packed lower triangle,
// lane = permuted position (row), R rows per lane (n <= G R).  The hot path for every n <= 64.
//
//   * packed row-major lower storage, element (i,k) at T(i)+k with T(i) = i(i+1)/2: 4.2 KB at
//     n = 32, 16.6 KB at n = 64 and ~60 registers per lane, so an SM holds 32 warps (the
//     register-resident family, warp_se.cuh, holds 16 and is latency bound).  The triangular
//     numbers are a permutation modulo 16 on every aligned run of 16 rows, so a column access
//     by 32 lanes (8-byte elements, 16 lanes per wavefront) is conflict free;
//   * rows live at their permuted POSITION, so the symmetric row/column swap (utils.rs:30-49)
//     is two loads and two predicated stores per lane and all indexing is run-time (no per-j
//     code specialisation);
//   * LEFT-looking: columns >= j keep the permuted input and column j is brought up to date when
//     it becomes the pivot column (wsm_left_column) -- the same roundings in the same order as
//     the reference's right-looking update; FAST mode with one matrix per warp additionally
//     applies the finished columns in blocks of four on the FP64 tensor pipe
//     (wsm_rank4_update_dmma);
//   * explicit 32-bit shared addressing, predicated single-instruction stores.
// Only the lower triangle of A is read (A symmetric), so DRAM reads drop below the
// algorithmic 8 n^2 bytes.  Measurements and the optimisation log: profiles/r01_notes.md.
#pragma once
#include <cstdio>
#include <type_traits>

#include "warp_common.cuh"

#ifndef MODCHOL_WSM_DMMA
#define MODCHOL_WSM_DMMA 1   /* FAST mode: blocked rank-4 trailing update with DMMA (0 = off) */
#endif

namespace modchol {

__host__ __device__ constexpr int tri(int i) { return (i * (i + 1)) >> 1; }

// st.shared.f64 [p], v  executed only where pred -- as ONE predicated instruction (written as
// `if (pred) *p = v;` the compiler wraps every element update in a BSSY/BRA/BSYNC diamond).
__device__ __forceinline__ void st_shared_if(bool pred, double* p, double v)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("{\n .reg .pred q;\n setp.ne.s32 q, %2, 0;\n @q st.shared.f64 [%0], %1;\n}"
                 :: "r"(a), "d"(v), "r"((int)pred) : "memory");
}

// doubles of shared memory per matrix: packed triangle + 16 scratch + a column line of n entries
// (+8: the catch-up sweep reads it four at a time), kept 16-byte aligned
__host__ __device__ inline int wsm_warp_doubles(int n)
{
    const int d = ((tri(n) + 1) & ~1) + 16 + ((n + 9) & ~1);
    // n <= 8: four matrices per warp, two per shared-memory wavefront -- a stride of 8 (mod 16)
    // doubles puts the second matrix on the banks the first one leaves free; n <= 4: eight per
    // warp, four per wavefront, stride 4 (mod 16): T(0..3) = {0,1,3,6} + {0,4,8,12} are distinct
    if (n <= 4)
        return ((d + 11) & ~15) + 4;
    return n <= 8 ? ((d + 7) & ~15) + 8 : d;
}

// local (per lane, over its R rows) then warp-wide extreme of v over candidate rows.
template <bool kMax, int R>
__device__ __forceinline__ double wsm_extreme(unsigned mask, const double (&v)[R], const bool (&cand)[R],
                                              double none, bool (&win)[R])
{
    if constexpr (R == 1)
        return warp_extreme<kMax>(mask, v[0], cand[0], none, win[0]);
    int khi[R];
    unsigned klo[R];
    bool c[R];
    int bh = (int)0x80000000;
    unsigned bl = 0u;
#pragma unroll
    for (int s = 0; s < R; ++s) {
        make_key(v[s], khi[s], klo[s]);
        c[s] = cand[s] && (v[s] == v[s]);
        if (!kMax) {
            khi[s] = ~khi[s];
            klo[s] = ~klo[s];
        }
        if (c[s] && (khi[s] > bh || (khi[s] == bh && klo[s] > bl))) {
            bh = khi[s];
            bl = klo[s];
        }
    }
    const int INT_MIN_ = (int)0x80000000;
    const int mh = __reduce_max_sync(mask, bh);
    const unsigned ml = __reduce_max_sync(mask, bh == mh ? bl : 0u);
#pragma unroll
    for (int s = 0; s < R; ++s)
        win[s] = c[s] && khi[s] == mh && klo[s] == ml;
    if (mh == INT_MIN_)
        return none;
    return kMax ? key_value(mh, ml) : key_value(~mh, ~ml);
}

// first-strict-maximum pivot search over rows [lo, n) (utils.rs:52-91): position of the winner
template <int R>
__device__ __forceinline__ int wsm_argmax_first(unsigned mask, const double (&v)[R],
                                                const bool (&active)[R], const int (&row)[R], int lo,
                                                double& best)
{
    if constexpr (R == 1)
        return warp_argmax_first(mask, v[0], active[0], row[0], lo, best);
    bool win[R];
    best = wsm_extreme<true, R>(mask, v, active, -INFINITY, win);
    unsigned mine = 1024u;
    bool head_nan = false;
#pragma unroll
    for (int s = 0; s < R; ++s) {
        if (win[s] && (unsigned)row[s] < mine)
            mine = (unsigned)row[s];
        head_nan = head_nan || (active[s] && row[s] == lo && (v[s] != v[s]));
    }
    int m = (int)__reduce_min_sync(mask, mine);
    if (__any_sync(mask, head_nan) || m > 255)
        m = lo;
    return m;
}

// Explicit 32-bit shared-window addressing.  With generic `double*` accesses the compiler keeps
// re-deriving the shared window base (S2R / S2UR / ULEA) under register pressure; one `unsigned`
// byte address per warp plus these accessors removes ~60 instructions per step.
// -DMODCHOL_DEBUG_BOUNDS=1 (the "dbg" build of modcholesky_b200/build.py, run by
// tests/test_parity_gpu.py::test_debug_bounds_build): every access through these accessors is
// checked against the block's shared-memory window and its natural alignment and traps
// otherwise -- the stand-in for compute-sanitizer, which is closed on the GPU pool.
#ifndef MODCHOL_DEBUG_BOUNDS
#define MODCHOL_DEBUG_BOUNDS 0
#endif
__device__ __forceinline__ void smem_check(unsigned a, unsigned bytes, int line = 0, const char* file = "")
{
#if MODCHOL_DEBUG_BOUNDS
    // the accessors only ever address the block's DYNAMIC shared memory (it starts behind the
    // system-reserved kilobyte and the kernel's static variables)
    extern __shared__ double modchol_dyn_smem_base[];
    const unsigned lo = (unsigned)__cvta_generic_to_shared(modchol_dyn_smem_base);
    unsigned dyn;
    asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
    if ((a & (bytes - 1u)) != 0u || a < lo || a + bytes > lo + dyn) {
        printf("modchol: shared-memory access out of bounds: address %u, %u bytes, window [%u, %u) (block %d thread %d, %s:%d)\n",
               a, bytes, lo, lo + dyn, (int)blockIdx.x, (int)threadIdx.x, file, line);
        asm volatile("trap;");
    }
#endif
}
__device__ __forceinline__ double lds_f64(unsigned a, int line = __builtin_LINE(), const char* file = __builtin_FILE())
{
    smem_check(a, 8u, line, file);
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ double2 lds_v2f64(unsigned a, int line = __builtin_LINE(), const char* file = __builtin_FILE())
{
    smem_check(a, 16u, line, file);
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f64(unsigned a, double v, int line = __builtin_LINE(), const char* file = __builtin_FILE())
{
    smem_check(a, 8u, line, file);
    asm volatile("st.shared.f64 [%0], %1;" :: "r"(a), "d"(v) : "memory");
}
__device__ __forceinline__ void sts_f64_if(bool pred, unsigned a, double v, int line = __builtin_LINE(), const char* file = __builtin_FILE())
{
    if (pred)
        smem_check(a, 8u, line, file);
    asm volatile("{\n .reg .pred q;\n setp.ne.s32 q, %2, 0;\n @q st.shared.f64 [%0], %1;\n}"
                 :: "r"(a), "d"(v), "r"((int)pred) : "memory");
}

// The same accessors on 64-bit GENERIC byte addresses (shared or global window), for the
// block-level kernels whose first rows are spilled to global memory (csm_se.cuh, n > 236).
__device__ __forceinline__ double lds_f64(unsigned long long a)
{
    double v;
    asm volatile("ld.f64 %0, [%1];" : "=d"(v) : "l"(a) : "memory");
    return v;
}
__device__ __forceinline__ double2 lds_v2f64(unsigned long long a)
{
    double2 v;
    asm volatile("ld.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f64(unsigned long long a, double v)
{
    asm volatile("st.f64 [%0], %1;" :: "l"(a), "d"(v) : "memory");
}
__device__ __forceinline__ void sts_f64_if(bool pred, unsigned long long a, double v)
{
    asm volatile("{\n .reg .pred q;\n setp.ne.s32 q, %2, 0;\n @q st.f64 [%0], %1;\n}"
                 :: "l"(a), "d"(v), "r"((int)pred) : "memory");
}

// ld.shared.f64 where pred, else 0.0 -- one predicated load into a zeroed register
__device__ __forceinline__ double lds_f64_or_zero(bool pred, unsigned a, int line = __builtin_LINE(), const char* file = __builtin_FILE())
{
    if (pred)
        smem_check(a, 8u, line, file);
    double v;
    asm volatile("{\n .reg .pred q;\n setp.ne.s32 q, %2, 0;\n mov.f64 %0, 0d0000000000000000;\n"
                 " @q ld.shared.f64 %0, [%1];\n}"
                 : "=d"(v) : "r"(a), "r"((int)pred) : "memory");
    return v;
}

// Epilogue shared by the shared-memory warp kernels: the packed lower triangle at byte address wb
// goes out as a dense row-major n x n matrix with an explicitly zero strict upper triangle
// (se99.rs:189-192); lane = column, one coalesced row per store, running addresses.
template <int G, int R>
__device__ __forceinline__ void wsm_store_lower(double* __restrict__ Lb, int n, unsigned wb, int sub)
{
    double* dst[R];
    unsigned src[R];
    bool kv[R];
#pragma unroll
    for (int s = 0; s < R; ++s) {
        dst[s] = Lb + sub + G * s;
        src[s] = wb + 8u * (unsigned)(sub + G * s);
        kv[s] = sub + G * s < n;
    }
#pragma unroll 4
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int s = 0; s < R; ++s) {
            const double v = lds_f64_or_zero(kv[s] && sub + G * s <= i, src[s]);
            if (kv[s])
                *dst[s] = v;
            dst[s] += n;
            src[s] += 8u * (unsigned)(i + 1);
        }
    }
}

// Prologue shared by the shared-memory warp kernels: the lower triangle of the dense row-major
// input goes into the packed triangle at byte address wb (lane = column, eight rows of loads in
// flight per trip, running addresses).  kOffMax: also return this lane's max |a_ik|, k < i
// (gmw81.rs:52-58; A symmetric, so the lower off-diagonals carry the off-diagonal maximum).
template <int G, int R, bool kOffMax>
__device__ __forceinline__ double wsm_load_lower(const double* __restrict__ Ab, int n, unsigned wb, int sub)
{
    const double* src[R];
    unsigned dst[R];
#pragma unroll
    for (int s = 0; s < R; ++s) {
        src[s] = Ab + sub + G * s;
        dst[s] = wb + 8u * (unsigned)(sub + G * s);
    }
    double om = 0.0;
    int i = 0;
    for (; i + 8 <= n; i += 8) {
        double v[8][R];
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int s = 0; s < R; ++s)
                v[u][s] = (sub + G * s <= i + u) ? __ldg(src[s] + (size_t)u * n) : 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int s = 0; s < R; ++s) {
                sts_f64_if(sub + G * s <= i + u, dst[s], v[u][s]);
                dst[s] += 8u * (unsigned)(i + u + 1);
                if (kOffMax)
                    om = fmax(om, (sub + G * s == i + u) ? 0.0 : fabs(v[u][s]));
            }
#pragma unroll
        for (int s = 0; s < R; ++s)
            src[s] += (size_t)8 * n;
    }
    for (; i < n; ++i) {
#pragma unroll
        for (int s = 0; s < R; ++s) {
            const bool in = sub + G * s <= i;
            const double v = in ? __ldg(src[s]) : 0.0;
            sts_f64_if(in, dst[s], v);
            dst[s] += 8u * (unsigned)(i + 1);
            src[s] += n;
            if (kOffMax)
                om = fmax(om, (sub + G * s == i) ? 0.0 : fabs(v));
        }
    }
    return om;
}

// ndarray-order sum of line[lo .. lo+len) (values already |.|), any len, returned to every lane:
// lanes 0..7 own the eight partials, then the pairwise combine, then the tail.
template <int G>
__device__ __forceinline__ double wsm_nd_sum_line(unsigned mask, unsigned lineb, int lo, int len, int sub)
{
    const int nch = len >> 3;
    double p = 0.0;
    if (sub < 8)
        for (int c = 0; c < nch; ++c)
            p = dadd(p, lds_f64(lineb + 8u * (unsigned)(lo + 8 * c + sub)));
    const double pair = dadd(p, __shfl_down_sync(mask, p, 4, G));
    double acc = dadd(0.0, __shfl_sync(mask, pair, 0, G));
    acc = dadd(acc, __shfl_sync(mask, pair, 1, G));
    acc = dadd(acc, __shfl_sync(mask, pair, 2, G));
    acc = dadd(acc, __shfl_sync(mask, pair, 3, G));
    for (int i = 8 * nch; i < len; ++i)
        acc = dadd(acc, lds_f64(lineb + 8u * (unsigned)(lo + i)));
    return acc;
}

__device__ __forceinline__ double wsm_nd_sum_line(unsigned lineb, int lo, int len, int lane)
{
    return wsm_nd_sum_line<32>(kFull, lineb, lo, len, lane);
}

// ndarray-order sum of |val(0)|, ..., |val(len-1)|, executed by ONE lane.
template <typename ValFn>
__device__ __forceinline__ double wsm_nd_sum_abs_fn(int len, ValFn val)
{
    double p[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
        p[u] = 0.0;
    int t = 0;
    for (; t + 8 <= len; t += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
            p[u] = dadd(p[u], fabs(val(t + u)));
    }
    double acc = 0.0;
    acc = dadd(acc, dadd(p[0], p[4]));
    acc = dadd(acc, dadd(p[1], p[5]));
    acc = dadd(acc, dadd(p[2], p[6]));
    acc = dadd(acc, dadd(p[3], p[7]));
    for (; t < len; ++t)
        acc = dadd(acc, fabs(val(t)));
    return acc;
}

// same with element t at W[addr(t)]
template <typename AddrFn>
__device__ __forceinline__ double wsm_nd_sum_abs_lane(const double* W, int len, AddrFn addr)
{
    return wsm_nd_sum_abs_fn(len, [&](int t) { return W[addr(t)]; });
}

// Symmetric exchange of positions j < m in the packed lower triangle (utils.rs:30-49):
//   i < j     : (j,i) <-> (m,i)        i == j : (j,j) <-> (m,m)
//   j < i < m : (i,j) <-> (m,i)        i == m : (m,j) stays        i > m : (i,j) <-> (i,m)
// wb = byte address of W(0,0), tjb = byte address of W(j,0), tib[s] = byte address of W(row,0).
// Branch free: two address selects, two loads, two predicated stores per row.  The per-row
// scalars (x, y, perm) of positions j and m are exchanged by shuffle (R == 1) or through the
// scratch line at scrb (R == 2).
template <int G, int R>
__device__ __forceinline__ void wsm_swap_positions(unsigned mask, unsigned wb, unsigned scrb,
                                                   const unsigned (&tib)[R], const int (&row)[R],
                                                   const bool (&rv)[R], int sub,
                                                   int j, int m, unsigned tjb, double (&x)[R],
                                                   double (&y)[R], int (&perm)[R])
{
    const unsigned tmb = wb + 8u * (unsigned)tri(m);
    unsigned a1[R], a2[R];
    double t1[R], t2[R];
    bool act[R];
#pragma unroll
    for (int s = 0; s < R; ++s) {
        const int i = row[s];
        act[s] = rv[s] && i != m;
        a1[s] = (i <= j) ? tjb + 8u * i : tib[s] + 8u * j;
        a2[s] = (i == j) ? tmb + 8u * m : ((i < m) ? tmb + 8u * i : tib[s] + 8u * m);
        if (!act[s]) {
            a1[s] = wb;
            a2[s] = wb;
        }
        t1[s] = lds_f64(a1[s]);
        t2[s] = lds_f64(a2[s]);
    }
    if constexpr (R == 1) {
        const int partner = (sub == j) ? m : ((sub == m) ? j : sub);
        x[0] = __shfl_sync(mask, x[0], partner, G);
        y[0] = __shfl_sync(mask, y[0], partner, G);
        perm[0] = __shfl_sync(mask, perm[0], partner, G);
    } else {
#pragma unroll
        for (int s = 0; s < R; ++s) {
            const int i = row[s];
            if (i == j || i == m) {
                const unsigned q = scrb + (i == j ? 0u : 32u);
                sts_f64(q, x[s]);
                sts_f64(q + 8, y[s]);
                sts_f64(q + 16, (double)perm[s]);
            }
        }
    }
    if constexpr (R > 1)
        __syncwarp(mask);   // scratch written above is read below (R == 1: every lane exchanges
                            // its own disjoint pair of elements, no barrier needed before the stores)
#pragma unroll
    for (int s = 0; s < R; ++s) {
        sts_f64_if(act[s], a1[s], t2[s]);
        sts_f64_if(act[s], a2[s], t1[s]);
        if constexpr (R > 1) {
            const int i = row[s];
            if (i == j || i == m) {
                const unsigned q = scrb + (i == j ? 32u : 0u);
                x[s] = lds_f64(q);
                y[s] = lds_f64(q + 8);
                perm[s] = (int)lds_f64(q + 16);
            }
        }
    }
    __syncwarp(mask);
}

// value v[s] of the row at position j, broadcast to the warp
template <int G, int R>
__device__ __forceinline__ double wsm_bcast_row(unsigned mask, const double (&v)[R], int j)
{
    double x = v[0];
#pragma unroll
    for (int s = 1; s < R; ++s)
        if (j / G == s)
            x = v[s];
    return __shfl_sync(mask, x, j % G, G);
}

// lanes [grp G, grp G + G) of the warp
template <int G>
__device__ __forceinline__ unsigned wsm_group_mask(int grp)
{
    if constexpr (G == 32)
        return kFull;
    else
        return ((1u << G) - 1u) << (grp * G);
}

// LEFT-looking evaluation of the pivot column.  The reference updates the whole trailing
// block at every step (l_ik -= l_ij l_kj, se90.rs:88-92); here columns >= j keep the
// permuted INPUT and column j is brought up to date only now:
//     col_i = a_ij - l_i,sbase l_j,sbase - ... - l_i,j-1 l_j,j-1      (in this order)
// which applies exactly the same roundings in the same order as the right-looking
// updates would have, with read-only shared-memory traffic (no store per element).
// sbase = first column not yet applied to the stored trailing block (0, or the phase-2
// start once the block has been materialised for the Gershgorin bounds).
// kDiag: row j evaluates its own element too (col of row j = c_jj), for GMW81.
template <int G, int R, bool kStrict, bool kDiag = false>
__device__ __forceinline__ void wsm_left_column(int j, unsigned tjb, int sbase, const unsigned (&tib)[R],
                                                const int (&row)[R], const bool (&rv)[R],
                                                double (&col)[R])
{
    const int lo = kDiag ? j - 1 : j;   // rows > lo take part
    // row j of L is contiguous at tjb (W(j, 0..j-1)) and is read by warp-wide broadcast
    // loads, two columns per LDS.128 once (tjb + 8 t) is 16-byte aligned.  Finished rows
    // (<= j) stay out of it: a lane whose rows are all finished skips the dot product
    // (an idle half warp halves the wavefronts of every LDS.64 of the other half), and a
    // slice s of rows that is finished in every lane is not evaluated at all.
#pragma unroll
    for (int s = 0; s < R; ++s)
        col[s] = 0.0;
    auto dot = [&](auto s0tag) {
        constexpr int S0 = decltype(s0tag)::value;
#pragma unroll
        for (int s = S0; s < R; ++s)
            col[s] = lds_f64(tib[s] + 8u * j);
        unsigned rj = tjb + 8u * (unsigned)sbase;
        unsigned ri[R];
#pragma unroll
        for (int s = S0; s < R; ++s)
            ri[s] = tib[s] + 8u * (unsigned)sbase;
        int left = j - sbase;
        auto step1 = [&]() {
            const double lj = lds_f64(rj);
#pragma unroll
            for (int s = S0; s < R; ++s)
                col[s] = nmul_add<kStrict>(col[s], lds_f64(ri[s]), lj);
        };
        auto step2 = [&](unsigned o) {
            const double2 l01 = lds_v2f64(rj + o);
#pragma unroll
            for (int s = S0; s < R; ++s) {
                const double w0 = lds_f64(ri[s] + o), w1 = lds_f64(ri[s] + o + 8);
                col[s] = nmul_add<kStrict>(col[s], w0, l01.x);
                col[s] = nmul_add<kStrict>(col[s], w1, l01.y);
            }
        };
        auto advance = [&](unsigned bytes) {
            rj += bytes;
#pragma unroll
            for (int s = S0; s < R; ++s)
                ri[s] += bytes;
        };
        if ((rj & 8u) && left > 0) {
            step1();
            advance(8);
            --left;
        }
        for (; left >= 8; left -= 8) {
            const double2 l01 = lds_v2f64(rj), l23 = lds_v2f64(rj + 16);
            const double2 l45 = lds_v2f64(rj + 32), l67 = lds_v2f64(rj + 48);
            const double lj[8] = {l01.x, l01.y, l23.x, l23.y, l45.x, l45.y, l67.x, l67.y};
#pragma unroll
            for (int s = S0; s < R; ++s) {
                double w[8];
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    w[u] = lds_f64(ri[s] + 8u * u);
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    col[s] = nmul_add<kStrict>(col[s], w[u], lj[u]);
            }
            advance(64);
        }
        if (left & 4) {
            step2(0);
            step2(16);
            advance(32);
        }
        if (left & 2) {
            step2(0);
            advance(16);
        }
        if (left & 1)
            step1();
    };
    if constexpr (R == 1) {
        if (rv[0] && row[0] > lo)
            dot(std::integral_constant<int, 0>{});
    } else {
        if (lo >= G - 1) {   // slice 0 (rows < G) is finished in every lane
            if (rv[R - 1] && row[R - 1] > lo)
                dot(std::integral_constant<int, R - 1>{});
        } else if (row[R - 1] > lo) {
            dot(std::integral_constant<int, 0>{});
        }
    }
#pragma unroll
    for (int s = 0; s < R; ++s)
        if (!(rv[s] && row[s] > lo))
            col[s] = 0.0;
}

// FAST mode, one matrix per warp: blocked trailing update on the FP64 tensor pipe.  Every four
// steps the finished columns [sbase, j) (at most four) are applied to the stored trailing block,
//     W(i,k) -= sum_t W(i,t) W(k,t),   j <= k <= i < n,
// as 8x8 tiles of one DMMA (mma.sync.m8n8k4.f64) each, so that the left-looking dot products
// never span more than three columns: 8x fewer own-row loads and FMA issue slots for the same
// flops.  Fragments: lane l holds A[l/4][l%4] = W(8 bi + l/4, sbase + l%4) (and the same element
// pattern serves as B for column block bj), C[l/4][2 (l%4) + {0,1}].  The roundings differ from
// the sequential order, hence FAST only.
template <int kBlocks>
__device__ __noinline__ void wsm_rank4_update_dmma(unsigned wb, int n, int j, int sbase, int lane)
{
    const int r8 = lane >> 2, c4 = lane & 3;
    const int b0 = j >> 3, nb = (n + 7) >> 3, cnt = j - sbase;
    double frag[kBlocks];
#pragma unroll
    for (int b = 0; b < kBlocks; ++b) {
        const int rr = 8 * b + r8;
        const bool in = b >= b0 && rr >= j && rr < n && c4 < cnt;
        frag[b] = lds_f64_or_zero(in, wb + 8u * (unsigned)(tri(in ? rr : 0) + sbase + c4));
    }
#pragma unroll
    for (int bi = 0; bi < kBlocks; ++bi) {
        if (bi < b0 || bi >= nb)
            continue;
        const int i = 8 * bi + r8;
        const unsigned ra = wb + 8u * (unsigned)tri(i < n ? i : 0);
        const double na = -frag[bi];
#pragma unroll
        for (int bj = 0; bj <= bi; ++bj) {
            if (bj < b0)
                continue;
            const int k0 = 8 * bj + 2 * c4, k1 = k0 + 1;
            const bool p0 = i < n && k0 >= j && k0 <= i, p1 = i < n && k1 >= j && k1 <= i;
            const double c0 = lds_f64_or_zero(p0, ra + 8u * (unsigned)k0);
            const double c1 = lds_f64_or_zero(p1, ra + 8u * (unsigned)k1);
            double d0, d1;
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%4, %5};"
                         : "=d"(d0), "=d"(d1) : "d"(na), "d"(frag[bj]), "d"(c0), "d"(c1));
            sts_f64_if(p0, ra + 8u * (unsigned)k0, d0);
            sts_f64_if(p1, ra + 8u * (unsigned)k1, d1);
        }
    }
}

// Fused consumer step: with the factor still in shared memory, solve
// (A + E) x = b for nrhs right-hand sides:  P (A+E) P^T = L L^T with P[i, p_i] = 1, hence
// L L^T (P x) = P b.  Lane = position: y_i = b[p_i]; forward substitution in axpy form over the
// COLUMNS of L (conflict-free column reads), backward substitution in axpy form over its ROWS
// (contiguous); x[p_i] = w_i.  Each step broadcasts one value and costs one LDS + one FMA per row.
template <int G, int R>
__device__ __forceinline__ void wsm_solve_in_place(unsigned mask, int n, unsigned wb,
                                                   const unsigned (&tib)[R], const int (&row)[R],
                                                   const bool (&rv)[R], const int (&perm)[R],
                                                   const double* rhs, double* x, int nrhs)
{
    double rd[R];   // 1 / L_ii of this lane's positions
#pragma unroll
    for (int s = 0; s < R; ++s)
        rd[s] = rv[s] ? ddiv(1.0, lds_f64(tib[s] + 8u * row[s])) : 0.0;
    for (int r = 0; r < nrhs; ++r) {
        double z[R], t[R];
#pragma unroll
        for (int s = 0; s < R; ++s)
            z[s] = rv[s] ? rhs[(size_t)r * n + perm[s]] : 0.0;
        for (int j = 0; j < n; ++j) {   // L z = y
#pragma unroll
            for (int s = 0; s < R; ++s)
                t[s] = dmul(z[s], rd[s]);
            const double zj = wsm_bcast_row<G, R>(mask, t, j);
#pragma unroll
            for (int s = 0; s < R; ++s) {
                if (row[s] == j)
                    z[s] = zj;
                else if (rv[s] && row[s] > j)
                    z[s] = __fma_rn(-lds_f64(tib[s] + 8u * j), zj, z[s]);
            }
        }
        unsigned tjb = wb + 8u * (unsigned)tri(n - 1);
        for (int j = n - 1; j >= 0; tjb -= 8u * (unsigned)j, --j) {   // L^T w = z
#pragma unroll
            for (int s = 0; s < R; ++s)
                t[s] = dmul(z[s], rd[s]);
            const double wj = wsm_bcast_row<G, R>(mask, t, j);
#pragma unroll
            for (int s = 0; s < R; ++s) {
                if (row[s] == j)
                    z[s] = wj;
                else if (row[s] < j)
                    z[s] = __fma_rn(-lds_f64(tjb + 8u * row[s]), wj, z[s]);
            }
        }
#pragma unroll
        for (int s = 0; s < R; ++s)
            if (rv[s])
                x[(size_t)r * n + perm[s]] = z[s];
    }
}

template <int kAlgo, int G, int R, bool kStrict, int kWarps, int kMinBlocks>
__global__ void __launch_bounds__(kWarps * 32, kMinBlocks)
wsm_se_kernel(int n, int wstride, long long batch, const double* __restrict__ A,
              double* __restrict__ L, double* __restrict__ E, unsigned long long* __restrict__ P,
              int* __restrict__ K, const double* Bm, double* X, int nrhs)   // X may alias Bm
{
    extern __shared__ double sm[];
    // G lanes per matrix, 32 / G matrices per warp; every collective below is confined to the
    // group (member mask + shuffle width), so the groups of a warp may diverge freely
    constexpr int kPack = 32 / G;
    int lane = threadIdx.x & 31;
    asm volatile("mov.u32 %0, %0;" : "+r"(lane));   // pinned, see wb below
    const int wib = threadIdx.x >> 5;
    const int sub = lane & (G - 1), grp = lane / G;
    const unsigned mask = wsm_group_mask<G>(grp);
    double* W = sm + (size_t)(wib * kPack + grp) * wstride;
    // byte address of W(0,0); pinned in a register (a volatile move stops the compiler from
    // re-deriving it -- thread index arithmetic, cvta -- at every use under register pressure)
    unsigned wb = (unsigned)__cvta_generic_to_shared(W);
    asm volatile("mov.u32 %0, %0;" : "+r"(wb));
    const unsigned scrb = wb + 8u * (unsigned)((tri(n) + 1) & ~1);      // 16 doubles of scratch
    const unsigned lineb = scrb + 128u;                                  // 72-entry column line
    const size_t nn = (size_t)n * n;
    const long long nslots = (long long)gridDim.x * kWarps * kPack;
    // blocked trailing update on the tensor pipe: FAST, one matrix per warp (mma.sync is warp-wide),
    // one row per lane (measured: n=32 uniform +3 %, mixed-phase +16 % -- the phase switch then
    // sweeps at most three columns; with two rows per lane, 32 < n <= 64, it loses 5-20 %)
    constexpr bool kDmma = !kStrict && G == 32 && R == 1 && MODCHOL_WSM_DMMA;

    int row[R];
    unsigned tib[R];   // byte address of W(row, 0)
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
        // ---- load the lower triangle, eight rows per trip (coalesced, loads batched) --------
        __syncwarp(mask);
        wsm_load_lower<G, R, false>(Ab, n, wb, sub);
        __syncwarp(mask);

        double dg[R], g[R], ev[R];
        int perm[R];
        double av[R];
#pragma unroll
        for (int s = 0; s < R; ++s) {
            dg[s] = rv[s] ? lds_f64(tib[s] + 8u * row[s]) : 0.0;
            g[s] = 0.0;
            ev[s] = 0.0;
            perm[s] = row[s];
            av[s] = fabs(dg[s]);
        }
        bool win[R];
        // gamma = max |a_ii| folded from 0.0 (se90.rs:55-57, se99.rs:54-56)
        const double gamma = wsm_extreme<true, R>(mask, av, rv, 0.0, win);
        const double taugamma = dmul(MODCHOL_TAU, gamma);
        double delta_prev = 0.0;
        int kstart = -1;

        // exchange of positions j < m: packed triangle + the row scalars dg, g, perm
        auto swap_positions = [&](int j, int m, unsigned tjb, bool /*with_g*/) {
            wsm_swap_positions<G, R>(mask, wb, scrb, tib, row, rv, sub, j, m, tjb, dg, g, perm);
        };
        auto bcast_row = [&](const double (&v)[R], int j) -> double { return wsm_bcast_row<G, R>(mask, v, j); };
        // D: s = sqrt(pivot), y such that cs = col*y (FAST) -- see warp_se.cuh
        auto scale = [&](double piv, double& sq, double& y, double& inv_piv) -> bool {
            const bool fast = !kStrict && fast_range_ok(piv);
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
        auto left_column = [&](int j, unsigned tjb, int sbase, double (&col)[R]) {
            wsm_left_column<G, R, kStrict>(j, tjb, sbase, tib, row, rv, col);
        };
        // write column j of L (se90.rs:84-87, se99.rs:96-99); the diagonal update of the
        // reference's trailing update (l_ii -= l_ij^2) is carried in dg
        auto eliminate = [&](int j, unsigned tjb, double sq, const double (&cs)[R],
                             const double (&dg_new)[R]) {
#pragma unroll
            for (int s = 0; s < R; ++s) {
                const bool below = rv[s] && row[s] > j;
                sts_f64_if(below || row[s] == j, below ? tib[s] + 8u * j : tjb + 8u * j,
                           below ? cs[s] : sq);
                if (below)
                    dg[s] = dg_new[s];
            }
            __syncwarp(mask);
        };
        int sbase = 0;

        // ---- phase 1 (se90.rs:61-96, se99.rs:61-109) --------------------------------------
        int j = 0;
        unsigned tjb = wb;   // byte address of W(j,0), advanced with j
        bool have_dmin = false;
        double dmin_carry = 0.0;
        for (; j < n; tjb += 8u * (unsigned)(j + 1), ++j) {
            if constexpr (kDmma) {
                if (j - sbase == 4) {
                    wsm_rank4_update_dmma<4 * R>(wb, n, j, sbase, lane);
                    sbase = j;
                    __syncwarp(mask);
                }
            }
            bool active[R];
#pragma unroll
            for (int s = 0; s < R; ++s)
                active[s] = rv[s] && row[s] >= j;
            double dmax;
            const int m = wsm_argmax_first<R>(mask, dg, active, row, j, dmax);
            if (kAlgo == kSE99) {  // se99.rs:62-73, before the pivot is applied
                // FAST: the minimum of the diagonal over rows >= j IS the previous step's
                // look-ahead minimum (same rows, same values) -- no second reduction
                const double dmin = (!kStrict && have_dmin)
                                        ? dmin_carry
                                        : wsm_extreme<false, R>(mask, dg, active, INFINITY, win);
                if (dmax < taugamma || dmin < dmul(-MODCHOL_MU, dmax)) {
                    kstart = j;
                    break;
                }
            }
            if (m != j)
                swap_positions(j, m, tjb, false);
            // FAST: a winner other than the head is a non-NaN maximum and its value is the pivot
            const double piv = (!kStrict && m != j) ? dmax : bcast_row(dg, j);
            double colv[R];
            left_column(j, tjb, sbase, colv);
            double sq, y, inv_piv;
            const bool fast = scale(piv, sq, y, inv_piv);
            double cs[R], dg_new[R], nv[R];
            bool below[R];
#pragma unroll
            for (int s = 0; s < R; ++s) {
                below[s] = rv[s] && row[s] > j;
                const double col = colv[s];
                cs[s] = fast ? dmul(col, y) : ddiv(col, sq);
                dg_new[s] = nmul_add<kStrict>(dg[s], cs[s], cs[s]);
                // look-ahead value (se90.rs:70-77, se99.rs:82-89); FAST: the updated diagonal
                nv[s] = fast ? dg_new[s] : dsub(dg[s], ddiv(dmul(col, col), piv));
            }
            const double la = wsm_extreme<false, R>(mask, nv, below, INFINITY, win);
            have_dmin = fast;
            dmin_carry = la;
            const double thresh = (kAlgo == kSE99) ? dmul(-MODCHOL_MU, gamma) : taugamma;
            if (la < thresh) {  // the pivot at j stays applied (se99.rs:90-93)
                kstart = j;
                break;
            }
            eliminate(j, tjb, sq, cs, dg_new);
        }

        // ---- phase 2 ----------------------------------------------------------------------
        if (kstart >= 0) {
            if (kAlgo == kSE99 && kstart == n - 1) {  // se99.rs:115-119
                const double ljj = bcast_row(dg, n - 1);
                const double ej = tail1x1_e(ljj, gamma);
#pragma unroll
                for (int s = 0; s < R; ++s)
                    if (row[s] == n - 1) {
                        ev[s] = ej;
                        sts_f64(tib[s] + 8u * row[s], dsqrt(dadd(ljj, ej)));
                    }
            } else {
                // The Gershgorin bounds need the whole trailing block as the reference holds it:
                // apply the k0 finished columns to it once (same order, same roundings).
                const int k0 = kstart;
                if (k0 > sbase) {
                    // one right-looking sweep per finished column t (in order): W(i,k) -= l_it l_kt
                    // for k0 <= k <= i, four columns per trip, LDS.128 broadcasts of column t from
                    // the line, unconditional load + FMA, predicated store
                    const int kk0 = k0 & ~1;
                    unsigned cnt[R];
#pragma unroll
                    for (int s = 0; s < R; ++s) {
                        const int c = rv[s] ? row[s] - k0 + 1 : 0;
                        cnt[s] = c > 0 ? (unsigned)c : 0u;      // columns k0 .. row[s]
                    }
                    for (int t = sbase; t < k0; ++t) {
                        double lt[R];
#pragma unroll
                        for (int s = 0; s < R; ++s) {
                            lt[s] = lds_f64(tib[s] + 8u * t);   // l_it (rows >= k0 > t)
                            sts_f64_if(rv[s], lineb + 8u * row[s], lt[s]);
                        }
                        __syncwarp(mask);
                        unsigned wr[R];
#pragma unroll
                        for (int s = 0; s < R; ++s)
                            wr[s] = tib[s] + 8u * kk0;
                        unsigned lk = lineb + 8u * kk0;
                        for (int k = kk0; k < n; k += 4) {
                            const double2 c01 = lds_v2f64(lk);
                            const double2 c23 = lds_v2f64(lk + 16);
                            const double ck[4] = {c01.x, c01.y, c23.x, c23.y};
                            const unsigned off = (unsigned)(k - k0);   // wraps for k == k0 - 1
#pragma unroll
                            for (int s = 0; s < R; ++s) {
                                double w[4];
#pragma unroll
                                for (int u = 0; u < 4; ++u)
                                    w[u] = lds_f64(wr[s] + 8u * u);
#pragma unroll
                                for (int u = 0; u < 4; ++u)
                                    sts_f64_if(off + (unsigned)u < cnt[s], wr[s] + 8u * u,
                                               nmul_add<kStrict>(w[u], lt[s], ck[u]));
                                wr[s] += 32;
                            }
                            lk += 32;
                        }
                        __syncwarp(mask);
                    }
                    sbase = k0;
                }
                // lower Gershgorin bounds (se90.rs:105-110, se99.rs:125-130)
#pragma unroll
                for (int s = 0; s < R; ++s) {
                    const int i = row[s];
                    if (rv[s] && i >= k0) {
                        double s1, s2;
                        const int tis = (int)((tib[s] - wb) >> 3);
                        if constexpr (kStrict) {
                            s1 = wsm_nd_sum_abs_lane(W, i - k0, [&](int t) { return tis + k0 + t; });
                            s2 = wsm_nd_sum_abs_lane(W, n - 1 - i, [&](int t) { return tri(i + 1 + t) + i; });
                            g[s] = dsub(dsub(dg[s], s1), s2);
                        } else {
                            s1 = 0.0;
                            s2 = 0.0;
                            for (int c = k0; c < i; ++c)
                                s1 = dadd(s1, fabs(lds_f64(tib[s] + 8u * c)));
                            unsigned tr = wb + 8u * (unsigned)(tri(i + 1) + i);
                            for (int r = i + 1; r < n; ++r) {
                                s2 = dadd(s2, fabs(lds_f64(tr)));
                                tr += 8u * (unsigned)(r + 1);
                            }
                            g[s] = dsub(dg[s], dadd(s1, s2));
                        }
                    }
                }
                for (j = k0; j + 2 < n; tjb += 8u * (unsigned)(j + 1), ++j) {
                    if constexpr (kDmma) {
                        if (j - sbase == 4) {
                            wsm_rank4_update_dmma<4 * R>(wb, n, j, sbase, lane);
                            sbase = j;
                            __syncwarp(mask);
                        }
                    }
                    bool active[R];
#pragma unroll
                    for (int s = 0; s < R; ++s)
                        active[s] = rv[s] && row[s] >= j;
                    double gmax;  // se90.rs:115, se99.rs:135
                    const int m = wsm_argmax_first<R>(mask, g, active, row, j, gmax);
                    if (m != j)
                        swap_positions(j, m, tjb, true);
                    double piv = bcast_row(dg, j);
                    double col[R], acol[R];
                    bool below[R];
                    double part = 0.0;
                    left_column(j, tjb, sbase, col);
#pragma unroll
                    for (int s = 0; s < R; ++s) {
                        below[s] = rv[s] && row[s] > j;
                        acol[s] = abs_bits(col[s]);
                        part = (s == 0) ? acol[s] : dadd(part, acol[s]);
                    }
                    // E_jj (se90.rs:124-131, se99.rs:144-151)
                    double normj;
                    if constexpr (kStrict) {
#pragma unroll
                        for (int s = 0; s < R; ++s)
                            sts_f64_if(rv[s], lineb + 8u * row[s], acol[s]);
                        __syncwarp(mask);
                        normj = wsm_nd_sum_line<G>(mask, lineb, j + 1, n - j - 1, sub);
                        __syncwarp(mask);
                    } else {
                        normj = warp_sum_tree<G>(mask, part);
                    }
                    const double ej =
                        max_b_not_nan(dadd(-piv, max_b_not_nan(normj, taugamma)), delta_prev);
                    if (kStrict) {
                        if (ej > 0.0) {
                            piv = dadd(piv, ej);
                            delta_prev = ej;
                        }
                    } else {
                        piv = dadd(piv, ej);
                        delta_prev = ej;
                    }
                    double sq, y, inv_piv;
                    const bool fast = scale(piv, sq, y, inv_piv);
                    // bound update (se90.rs:134-139, se99.rs:154-159)
                    const bool upd = fabs(dsub(piv, normj)) > MODCHOL_EPS;
                    double t = 0.0;
                    if (upd)
                        t = fast ? __fma_rn(-normj, inv_piv, 1.0) : dsub(1.0, ddiv(normj, piv));
                    double cs[R], dg_new[R];
#pragma unroll
                    for (int s = 0; s < R; ++s) {
                        if (row[s] == j)
                            ev[s] = ej;
                        cs[s] = fast ? dmul(col[s], y) : ddiv(col[s], sq);
                        dg_new[s] = nmul_add<kStrict>(dg[s], cs[s], cs[s]);
                        if (upd && below[s])
                            g[s] = mul_add<kStrict>(g[s], acol[s], t);
                    }
                    eliminate(j, tjb, sq, cs, dg_new);
                }
                // final 2x2 block, never pivoted (se90.rs:155-166, se99.rs:175-186)
                double a00 = bcast_row(dg, n - 2);
                double a11 = bcast_row(dg, n - 1);
                const unsigned a10b = wb + 8u * (unsigned)(tri(n - 1) + n - 2);
                double c2[R];
                left_column(n - 2, wb + 8u * (unsigned)tri(n - 2), sbase, c2);
                double a10 = bcast_row(c2, n - 1);
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
                __syncwarp(mask);
#pragma unroll
                for (int s = 0; s < R; ++s) {
                    if (row[s] == n - 2) {
                        ev[s] = en;
                        sts_f64(tib[s] + 8u * row[s], a00);
                    }
                    if (row[s] == n - 1) {
                        ev[s] = en;
                        sts_f64(a10b, a10);
                        sts_f64(tib[s] + 8u * row[s], a11);
                    }
                }
            }
        }
        __syncwarp(mask);

        // ---- epilogue: L (strict upper explicitly zero), e un-permuted, p -------------------
        // (every output is optional: the fused solve may be all the caller wants)
        if (L)
            wsm_store_lower<G, R>(L + (size_t)b * nn, n, wb, sub);
#pragma unroll
        for (int s = 0; s < R; ++s)
            if (rv[s]) {
                if (E)
                    E[(size_t)b * n + perm[s]] = ev[s];                   // se99.rs:194-198
                if (P)
                    P[(size_t)b * n + row[s]] = (unsigned long long)perm[s];
            }
        if (K && sub == 0)
            K[b] = kstart;
        if (X)
            wsm_solve_in_place<G, R>(mask, n, wb, tib, row, rv, perm, Bm + (size_t)b * nrhs * n,
                                     X + (size_t)b * nrhs * n, nrhs);
    }
}

}  // namespace modchol
