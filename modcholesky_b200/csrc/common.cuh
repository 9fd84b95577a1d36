// common.cuh -- arithmetic primitives shared by every modified-Cholesky kernel.
//
// The reference (argmin-rs/modcholesky, pure Rust) never contracts a - b*c into an FMA,
// divides with true IEEE division and sums norms with ndarray's eight-partial
// `unrolled_fold`.  `Arith<true>` (STRICT) reproduces every one of those roundings with
// the never-contracted __d*_rn intrinsics; `Arith<false>` (FAST) is free to fuse.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace modchol {

// cbrt(f64::EPSILON), se90.rs:51 / se99.rs:48-49.  CUDA's cbrt() is not correctly
// rounded, so the bits are hard-coded (they equal glibc's and mpmath's result).
__device__ __constant__ const double kTauDev = 0x1.965fea53d6e3dp-18;
#define MODCHOL_TAU 0x1.965fea53d6e3dp-18
#define MODCHOL_MU 0.1                      /* se99.rs:50 */
#define MODCHOL_EPS 2.220446049250313e-16   /* f64::EPSILON */

enum : int { kGMW81 = 0, kSE90 = 1, kSE99 = 2 };

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double dsqrt(double a) { return __dsqrt_rn(a); }

// a - b*c with the reference's two roundings (STRICT) or one (FAST).
template <bool kStrict>
__device__ __forceinline__ double nmul_add(double a, double b, double c)
{
    if constexpr (kStrict)
        return __dsub_rn(a, __dmul_rn(b, c));
    else
        return __fma_rn(-b, c, a);
}
// a + b*c
template <bool kStrict>
__device__ __forceinline__ double mul_add(double a, double b, double c)
{
    if constexpr (kStrict)
        return __dadd_rn(a, __dmul_rn(b, c));
    else
        return __fma_rn(b, c, a);
}

// utils.rs:14-27 eigenvalues_2x2 -> (hi, lo); reads both off-diagonals b=m[0,1], c=m[1,0].
__device__ __forceinline__ void eigenvalues_2x2(double a, double b, double c, double d, double& hi,
                                                double& lo)
{
    const double h = ddiv(-dadd(a, d), 2.0);
    const double t = dsqrt(dadd(dsub(dmul(h, h), dmul(a, d)), dmul(b, c)));
    const double half = ddiv(dadd(a, d), 2.0);
    const double l1 = dadd(half, t);
    const double l2 = dsub(half, t);
    if (l1 > l2) {
        hi = l1;
        lo = l2;
    } else {
        hi = l2;
        lo = l1;
    }
}

// E rule of the trailing 2x2 block (se90.rs:156-158 / se99.rs:176-178).
__device__ __forceinline__ double tail2x2_e(int variant, double lhi, double llo, double gamma,
                                            double delta_prev)
{
    const double tau = MODCHOL_TAU;
    const double spread = dsub(lhi, llo);
    double x;
    if (variant == kSE90)
        x = dadd(-llo, dmul(tau, fmax(gamma, dmul(ddiv(1.0, dsub(1.0, tau)), spread))));
    else
        x = dadd(-llo, fmax(dmul(tau, gamma), ddiv(dmul(tau, spread), dsub(1.0, tau))));
    return fmax(fmax(0.0, x), delta_prev);
}

// E rule of a phase-2 step (se90.rs:125-127 / se99.rs:145-147; tau*gamma == tau_bar*gamma).
__device__ __forceinline__ double phase2_e(double ljj, double normj, double taugamma,
                                           double delta_prev)
{
    return fmax(fmax(0.0, delta_prev), dadd(-ljj, fmax(normj, taugamma)));
}

// SE99's single-element tail (se99.rs:115-119).
__device__ __forceinline__ double tail1x1_e(double ljj, double gamma)
{
    const double tau = MODCHOL_TAU;
    return dadd(-ljj, fmax(dmul(tau, gamma), ddiv(dmul(tau, -ljj), dsub(1.0, tau))));
}

// ndarray 0.16 `Array1::sum()` order (numeric_util::unrolled_fold) executed by ONE thread
// over |x[i*stride]| (kAbs) or x[i*stride]: eight partials over chunks of eight,
// ((((0+(p0+p4))+(p1+p5))+(p2+p6))+(p3+p7)), then the (<8) tail in order.
template <bool kAbs>
__device__ __forceinline__ double nd_sum_thread(const double* x, int stride, int len)
{
    double p[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
        p[u] = 0.0;
    int i = 0;
    for (; i + 8 <= len; i += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            double v = x[(size_t)(i + u) * stride];
            p[u] = dadd(p[u], kAbs ? fabs(v) : v);
        }
    }
    double acc = 0.0;
    acc = dadd(acc, dadd(p[0], p[4]));
    acc = dadd(acc, dadd(p[1], p[5]));
    acc = dadd(acc, dadd(p[2], p[6]));
    acc = dadd(acc, dadd(p[3], p[7]));
    for (; i < len; ++i) {
        double v = x[(size_t)i * stride];
        acc = dadd(acc, kAbs ? fabs(v) : v);
    }
    return acc;
}

// ---- block-wide reductions (CTA-per-matrix kernels) -------------------------------------
// `red` is shared scratch of >= 66 doubles.  All threads of the block must call; the result
// is returned to every thread.  fmax/fmin ignore NaN exactly like the reference's
// `if x > acc {x} else {acc}` folds (gmw81.rs:49-58, se99.rs:62-70).
template <bool kIsMax>
__device__ __forceinline__ double block_reduce_minmax(double v, double* red)
{
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double w = __shfl_xor_sync(full, v, o);
        v = kIsMax ? fmax(v, w) : fmin(v, w);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();  // protect `red` from the previous reduction's readers
    if (lane == 0)
        red[warp] = v;
    __syncthreads();
    double r = red[0];
    for (int w = 1; w < nw; ++w)
        r = kIsMax ? fmax(r, red[w]) : fmin(r, red[w]);
    return r;
}

// (value, index) pair ordering for the reference's first-strict-maximum scan
// (utils.rs:52-91): larger value wins, equal values -> lower index wins, NaN never wins.
__device__ __forceinline__ void argmax_combine(double& v, int& i, double w, int wi)
{
    // wi < 0 marks "no candidate"
    if (wi >= 0 && (i < 0 || w > v || (w == v && wi < i))) {
        v = w;
        i = wi;
    }
}

// index in [lo, n) of the first maximum of key(t) (t strided over the block), with the
// reference's NaN rule: NaN never replaces the running maximum, a NaN at `lo` sticks.
template <typename KeyFn>
__device__ __forceinline__ int block_argmax_first(int lo, int n, KeyFn key, double* red)
{
    const unsigned full = 0xffffffffu;
    double bv = 0.0;
    int bi = -1;
    bool head_nan = false;
    for (int t = lo + threadIdx.x; t < n; t += blockDim.x) {
        double c = key(t);
        if (c != c) {
            if (t == lo)
                head_nan = true;
            continue;
        }
        argmax_combine(bv, bi, c, t);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double w = __shfl_xor_sync(full, bv, o);
        int wi = __shfl_xor_sync(full, bi, o);
        argmax_combine(bv, bi, w, wi);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    int* redi = reinterpret_cast<int*>(red + 32);
    __syncthreads();
    if (lane == 0) {
        red[warp] = bv;
        redi[warp] = bi;
    }
    if (threadIdx.x == 0)
        redi[32] = head_nan ? 1 : 0;  // thread 0 owns element `lo`
    __syncthreads();
    double rv = red[0];
    int ri = redi[0];
    for (int w = 1; w < nw; ++w)
        argmax_combine(rv, ri, red[w], redi[w]);
    if (redi[32] || ri < 0)
        ri = lo;
    return ri;
}

}  // namespace modchol
