// warp_common.cuh -- building blocks of the warp-per-matrix, register-resident kernels.
//
// Layout (one warp = one matrix, NMAX <= 32):
//   lane i owns ORIGINAL row i for the whole factorisation (rows are never moved: the
//   symmetric row swap of utils.rs:41-49 is implicit in `pos`, the permuted position of the
//   lane's row);   r[c] = current element (row of this lane, permuted COLUMN position c),
//   so the column swap of utils.rs:30-38 is a physical exchange of two registers in every
//   lane, selected by a warp-uniform switch (no dynamic register indexing, no local memory).
//   dg = current diagonal element of the lane's row (r[pos] would need a dynamic index).
// Measured on B200 (tools/microbench.cu): FP64 pipe 2 warp-instr/clk/SM, fmax/DSETP on
// doubles 0.43-0.63, SHFL and REDUX 1.0, broadcast LDS.64 ~1.0.  Hence: pivot searches run
// on order-preserving integer keys with REDUX instead of FP64 compares + shuffles, and the
// pivot column is broadcast through a 256-byte shared-memory line instead of shuffles.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace modchol {

constexpr unsigned kFull = 0xffffffffu;

template <int I, int End, typename F>
__device__ __forceinline__ void static_for(F&& f)
{
    if constexpr (I < End) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, End>(static_cast<F&&>(f));
    }
}

// ---- order-preserving 64-bit keys -------------------------------------------------------
// (khi signed, klo unsigned) compares lexicographically like the double it came from;
// -0 is folded onto +0 (IEEE `>` treats them as equal, utils.rs:63).
__device__ __forceinline__ void make_key(double v, int& khi, unsigned& klo)
{
    v = dadd(v, 0.0);   // -0.0 -> +0.0 (the two compare equal in the reference); one FP64 add
    const int hi = __double2hiint(v);
    const unsigned lo = (unsigned)__double2loint(v);
    const int mask = hi >> 31;
    khi = hi ^ (mask & 0x7fffffff);
    klo = lo ^ (unsigned)mask;
}
__device__ __forceinline__ double key_value(int khi, unsigned klo)
{
    const int mask = khi >> 31;
    return __hiloint2double(khi ^ (mask & 0x7fffffff), (int)(klo ^ (unsigned)mask));
}

// Extreme (max or min) of v over the lanes with `cand`, NaN lanes never taken -- the
// semantics of the reference's `if x > acc {x} else {acc}` folds and of utils.rs:52-70.
// Returns the extreme value (`none` if no lane qualifies); `win` = this lane attains it.
// `mask` is the member mask of the lane group that owns one matrix (the whole warp, or an
// aligned half / quarter when several small matrices share a warp).
template <bool kMax>
__device__ __forceinline__ double warp_extreme(unsigned mask, double v, bool cand, double none,
                                               bool& win)
{
    int khi;
    unsigned klo;
    make_key(v, khi, klo);
    cand = cand && (v == v);
    if (!kMax) {
        khi = ~khi;
        klo = ~klo;
    }
    const int INT_MIN_ = (int)0x80000000;
    const int mh = __reduce_max_sync(mask, cand ? khi : INT_MIN_);
    const bool c1 = cand && (khi == mh);
    const unsigned ml = __reduce_max_sync(mask, c1 ? klo : 0u);
    win = c1 && (klo == ml);
    if (mh == INT_MIN_)
        return none;
    return kMax ? key_value(mh, ml) : key_value(~mh, ~ml);
}

// First-strict-maximum pivot search (utils.rs:52-91) over the lanes with `active`, in
// permuted-position order: returns the POSITION of the winner (lowest position among equal
// maxima; a NaN at position `lo` sticks; all-NaN -> lo).  `best` receives the maximum.
__device__ __forceinline__ int warp_argmax_first(unsigned mask, double v, bool active, int pos,
                                                 int lo, double& best)
{
    bool win;
    best = warp_extreme<true>(mask, v, active, -INFINITY, win);
    int m = (int)__reduce_min_sync(mask, win ? (unsigned)pos : 1024u);
    const unsigned head_nan = __ballot_sync(mask, active && pos == lo && (v != v));
    if (head_nan || m > 255)
        m = lo;
    return m;
}

// ---- column exchange j <-> m with a register copy of column j ---------------------------
// The kernels keep `cj`, a register copy of the CURRENT content of column j (it falls out of
// the previous step's update for free).  Exchanging columns j and m (m > j, group-uniform) is
// then a single-level switch on m:   col = r[m];  r[m] = cj;   and the physical r[j] is not
// touched -- the elimination step overwrites it anyway.  (A first version dispatched on the
// pair (j, m): 496 cases, 63 KB of code, two instruction-cache misses per step.)
// The overwrite is an in-place asm: written as a plain assignment, every case of the switch
// makes ptxas materialise a fresh copy of the WHOLE register-resident matrix at the join.
__device__ __forceinline__ void overwrite_in_place(double& dst, double v)
{
    asm volatile("mov.f64 %0, %1;" : "+d"(dst) : "d"(v));
}
// old = slot; slot = v -- in ONE asm: a separate C++ read of the slot next to the asm write
// brings the whole-matrix copies back (toy kernel: 6552 vs 1480 SASS instructions).
__device__ __forceinline__ double exchange_in_place(double& slot, double v)
{
    double old;
    asm volatile("{\n mov.f64 %0, %1;\n mov.f64 %1, %2;\n}" : "=d"(old), "+d"(slot) : "d"(v));
    return old;
}

#define MODCHOL_ALLC(X)                                                                        \
    X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15)       \
    X(16) X(17) X(18) X(19) X(20) X(21) X(22) X(23) X(24) X(25) X(26) X(27) X(28) X(29) X(30)  \
    X(31)

// returns the old r[m] and stores cj there
template <int NMAX>
__device__ __forceinline__ double take_column(double (&r)[NMAX], int m, double cj)
{
    double col = 0.0;
#define MODCHOL_X(C)                            \
    case C:                                     \
        if constexpr (C < NMAX)                 \
            col = exchange_in_place(r[C], cj);  \
        break;
    switch (m) {
        MODCHOL_ALLC(MODCHOL_X)
    default:
        break;
    }
#undef MODCHOL_X
    return col;
}

// r[c] = v for a group-uniform run-time column c (in place, see overwrite_in_place)
template <int NMAX>
__device__ __forceinline__ void put_column(double (&r)[NMAX], int c, double v)
{
#define MODCHOL_X(C)                            \
    case C:                                     \
        if constexpr (C < NMAX)                 \
            overwrite_in_place(r[C], v);        \
        break;
    switch (c) {
        MODCHOL_ALLC(MODCHOL_X)
    default:
        break;
    }
#undef MODCHOL_X
}

// v = r[c] for a group-uniform run-time column c, by switch (12 instructions; the select
// chain below costs 3 per column)
template <int NMAX>
__device__ __forceinline__ double read_column(const double (&r)[NMAX], int c)
{
    double v = 0.0;
#define MODCHOL_X(C)                \
    case C:                         \
        if constexpr (C < NMAX)     \
            v = r[C];               \
        break;
    switch (c) {
        MODCHOL_ALLC(MODCHOL_X)
    default:
        break;
    }
#undef MODCHOL_X
    return v;
}
// r[c] = v in the lanes with pred, for a group-uniform run-time column c
template <int NMAX>
__device__ __forceinline__ void put_column_if(double (&r)[NMAX], int c, bool pred, double v)
{
    const double old = read_column<NMAX>(r, c);
    put_column<NMAX>(r, c, pred ? v : old);
}

// r[c] for a per-lane run-time column c (select chain).
template <int NMAX>
__device__ __forceinline__ double get_col(const double (&r)[NMAX], int c)
{
    double v = 0.0;
#pragma unroll
    for (int i = 0; i < NMAX; ++i)
        if (i == c)
            v = r[i];
    return v;
}
template <int NMAX>
__device__ __forceinline__ void set_col(double (&r)[NMAX], int c, bool pred, double v)
{
#pragma unroll
    for (int i = 0; i < NMAX; ++i)
        if (pred && i == c)
            r[i] = v;
}

// ndarray-order (numeric_util::unrolled_fold) sum of x_0..x_{len-1}, lane t holding x_t
// (lanes >= len must hold 0); len <= G - 1, lanes counted within the group of G.  Returned to
// every lane of the group.
template <int G>
__device__ __forceinline__ double nd_sum_lanes(unsigned mask, double x, int len)
{
    const int nch = len >> 3, tail = len & 7;
    double p = (nch >= 1) ? x : 0.0;  // lanes 0..7: 0 + x_u == x_u (x_u >= +0 or any finite)
    if (nch >= 2)
        p = dadd(p, __shfl_down_sync(mask, x, 8, G));
    if (nch >= 3)
        p = dadd(p, __shfl_down_sync(mask, x, 16, G));
    const double pair = dadd(p, __shfl_down_sync(mask, p, 4, G));  // lanes 0..3: p_u + p_{u+4}
    double acc = dadd(0.0, __shfl_sync(mask, pair, 0, G));
    acc = dadd(acc, __shfl_sync(mask, pair, 1, G));
    acc = dadd(acc, __shfl_sync(mask, pair, 2, G));
    acc = dadd(acc, __shfl_sync(mask, pair, 3, G));
    const int base = nch << 3;
    for (int i = 0; i < tail; ++i)
        acc = dadd(acc, __shfl_sync(mask, x, base + i, G));
    return acc;
}

// |x| on the integer pipe (fabs() of a select compiles to an FP64-pipe DADD).
__device__ __forceinline__ double abs_bits(double x)
{
    return __hiloint2double(__double2hiint(x) & 0x7fffffff, __double2loint(x));
}

// max(a, b) for b that is never NaN: `a > b ? a : b` equals f64::max / fmax (a NaN -> b) at
// a third of the instructions of the NaN-symmetric fmax().
__device__ __forceinline__ double max_b_not_nan(double a, double b) { return a > b ? a : b; }

// 1/sqrt(v) for v already known to be a positive normal number in [2^-767, 2^768]
// (fast_range_ok): MUFU.RSQ64H seed + one cubically convergent correction, i.e. CUDA's
// rsqrt() without its range checks and slow path.  Error < 1 ulp.
__device__ __forceinline__ double rsqrt_normal(double v)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(v));
    const double t = __dmul_rn(y, y);
    const double e = __fma_rn(-t, v, 1.0);                 // 1 - v*y^2
    const double p = __fma_rn(e, 0.375, 0.5);              // 1/2 + 3/8 e
    const double ey = __dmul_rn(e, y);
    return __fma_rn(p, ey, y);                             // y (1 + e/2 + 3/8 e^2)
}

template <int G>
__device__ __forceinline__ double warp_sum_tree(unsigned mask, double x)
{
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1)
        x = dadd(x, __shfl_xor_sync(mask, x, o, G));
    return x;
}

// True when v is a positive normal number comfortably inside the range where
// 1/sqrt(v), v*rsqrt(v) and their squares neither overflow nor lose precision; everything
// else (<= 0, NaN, Inf, subnormal, huge) takes the STRICT formulas so that the special-value
// behaviour of FAST mode is the reference's.
__device__ __forceinline__ bool fast_range_ok(double v)
{
    const unsigned e = ((unsigned)__double2hiint(v) >> 20);  // sign + exponent
    return e >= 0x100u && e <= 0x6ffu;                       // 2^-767 .. 2^+768, sign clear
}

}  // namespace modchol
