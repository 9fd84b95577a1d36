// cta_kernels.cuh -- CTA-per-matrix kernels: any 1 <= n <= MODCHOL_MAX_N, STRICT arithmetic.
//
// One thread block factorises one matrix at a time (persistent, grid-stride over the
// batch).  The working matrix W (the reference's `l = self.clone()`) lives in shared
// memory when it fits ("cta_smem": n <= ~160) and otherwise in the caller's L buffer in
// global memory / L2 ("cta_gmem").  These kernels follow the reference's access pattern
// literally -- full symmetric storage, physical row+column swaps, lower-triangle update
// mirrored to the upper triangle -- and are bit-identical to the CPU oracle for every
// input, including NaN/Inf propagation and asymmetric garbage.  They are the coverage path
// for sizes that have no specialised kernel yet; the warp-per-matrix kernels
// (warp_kernels.cuh) are the throughput path for n <= 32.
#pragma once
#include "common.cuh"

namespace modchol {

struct CtaShared {
    double* W;      // n x ld working matrix (shared or global)
    double* g;      // n   Gershgorin bounds (SE) / running diagonal of c (GMW81)
    double* ev;     // n   E in permuted order
    double* colv;   // n   current column
    double* dv;     // n   d (GMW81) / scratch
    double* red;    // 66  reduction scratch
    int* perm;      // n
};

__device__ __forceinline__ CtaShared carve_shared(double* sm, int n, int ld, bool use_smem,
                                                  double* gmemW)
{
    CtaShared s;
    s.g = sm;
    s.ev = s.g + n;
    s.colv = s.ev + n;
    s.dv = s.colv + n;
    s.red = s.dv + n;
    s.perm = reinterpret_cast<int*>(s.red + 66);
    double* after = s.red + 66 + (n + 1) / 2 + 1;
    s.W = use_smem ? after : gmemW;
    (void)ld;
    return s;
}

__host__ __device__ inline size_t cta_shared_bytes(int n, int ld, bool use_smem)
{
    size_t d = (size_t)4 * n + 66 + (n + 1) / 2 + 1;
    if (use_smem)
        d += (size_t)n * ld;
    return d * sizeof(double);
}

// utils.rs:41-49 swap_rows then utils.rs:30-38 swap_columns on the full matrix.
__device__ __forceinline__ void cta_swap(double* W, int n, int ld, int r1, int r2)
{
    for (int c = threadIdx.x; c < n; c += blockDim.x) {
        double t = W[(size_t)r1 * ld + c];
        W[(size_t)r1 * ld + c] = W[(size_t)r2 * ld + c];
        W[(size_t)r2 * ld + c] = t;
    }
    __syncthreads();
    for (int r = threadIdx.x; r < n; r += blockDim.x) {
        double t = W[(size_t)r * ld + r1];
        W[(size_t)r * ld + r1] = W[(size_t)r * ld + r2];
        W[(size_t)r * ld + r2] = t;
    }
    __syncthreads();
}

// se90.rs:84-93 == se90.rs:142-151 == se99.rs:96-105 == se99.rs:162-171.
__device__ __forceinline__ void cta_se_eliminate(const CtaShared& s, int n, int ld, int j)
{
    double* W = s.W;
    const double ljj = dsqrt(W[(size_t)j * ld + j]);
    __syncthreads();
    if (threadIdx.x == 0)
        W[(size_t)j * ld + j] = ljj;
    for (int i = j + 1 + threadIdx.x; i < n; i += blockDim.x) {
        const double v = ddiv(W[(size_t)i * ld + j], ljj);
        W[(size_t)i * ld + j] = v;
        W[(size_t)j * ld + i] = ddiv(W[(size_t)j * ld + i], ljj);
        s.colv[i] = v;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int i = j + 1 + warp; i < n; i += nw) {
        const double ci = s.colv[i];
        for (int k = j + 1 + lane; k <= i; k += 32) {
            const double v = dsub(W[(size_t)i * ld + k], dmul(ci, s.colv[k]));
            W[(size_t)i * ld + k] = v;
            W[(size_t)k * ld + i] = v;
        }
    }
    __syncthreads();
}

// ndarray-order sum of |colv[lo..n)| computed by warp 0 (8 lanes own the 8 partials),
// broadcast to the whole block through red[64].
__device__ __forceinline__ double cta_nd_sum_abs_colv(const CtaShared& s, int lo, int n)
{
    const int len = n - lo;
    const int nchunk = len >> 3;
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        double p = 0.0;
        if (lane < 8)
            for (int c = 0; c < nchunk; ++c)
                p = dadd(p, fabs(s.colv[lo + 8 * c + lane]));
        const double q = __shfl_down_sync(0xffffffffu, p, 4);
        const double pair = dadd(p, q);  // lanes 0..3: p_u + p_{u+4}
        double acc = 0.0;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            acc = dadd(acc, __shfl_sync(0xffffffffu, pair, u));
        if (lane == 0) {
            for (int i = 8 * nchunk; i < len; ++i)
                acc = dadd(acc, fabs(s.colv[lo + i]));
            s.red[64] = acc;
        }
    }
    __syncthreads();
    const double r = s.red[64];
    __syncthreads();
    return r;
}

// ---------------------------------------------------------------------------------------
// SE90 / SE99   (se90.rs:38-181, se99.rs:35-201)
// ---------------------------------------------------------------------------------------
template <int kThreads>
__global__ void __launch_bounds__(kThreads)
se_cta_kernel(int variant, int n, int ld, int use_smem, long long batch,
              const double* __restrict__ A, double* __restrict__ L, double* __restrict__ E,
              unsigned long long* __restrict__ P, int* __restrict__ K)
{
    extern __shared__ double sm[];
    const double tau = MODCHOL_TAU, mu = MODCHOL_MU;
    const size_t nn = (size_t)n * n;

    for (long long b = blockIdx.x; b < batch; b += gridDim.x) {
        const double* Ab = A + (size_t)b * nn;
        double* Lb = L + (size_t)b * nn;
        CtaShared s = carve_shared(sm, n, ld, use_smem != 0, Lb);
        double* W = s.W;

        __syncthreads();  // previous matrix fully written out before smem is reused
        for (size_t idx = threadIdx.x; idx < nn; idx += kThreads) {
            const int r = (int)(idx / n), c = (int)(idx - (size_t)r * n);
            W[(size_t)r * ld + c] = Ab[idx];
        }
        for (int i = threadIdx.x; i < n; i += kThreads) {
            s.ev[i] = 0.0;
            s.perm[i] = i;
            s.g[i] = 0.0;
        }
        __syncthreads();

        // gamma = max |a_ii|, fold from 0.0 (se90.rs:55-57, se99.rs:54-56)
        double v = 0.0;
        for (int i = threadIdx.x; i < n; i += kThreads)
            v = fmax(v, fabs(W[(size_t)i * ld + i]));
        const double gamma = block_reduce_minmax<true>(v, s.red);
        const double taugamma = dmul(tau, gamma);

        bool phaseone = true;
        int j = 0;
        while (j < n && phaseone) {
            if (variant == kSE99) {  // se99.rs:62-73
                double mx = -INFINITY, mn = INFINITY;
                for (int i = j + threadIdx.x; i < n; i += kThreads) {
                    const double x = W[(size_t)i * ld + i];
                    mx = fmax(mx, x);
                    mn = fmin(mn, x);
                }
                mx = block_reduce_minmax<true>(mx, s.red);
                mn = block_reduce_minmax<false>(mn, s.red);
                if (mx < taugamma || mn < dmul(-mu, mx)) {
                    phaseone = false;
                    break;
                }
            }
            // pivot on the largest remaining diagonal (se90.rs:63-68, se99.rs:76-81)
            const int m = block_argmax_first(
                j, n, [&](int t) { return W[(size_t)t * ld + t]; }, s.red);
            if (m != j) {
                __syncthreads();
                cta_swap(W, n, ld, j, m);
                if (threadIdx.x == 0) {
                    int t = s.perm[j];
                    s.perm[j] = s.perm[m];
                    s.perm[m] = t;
                }
            }
            __syncthreads();
            // look-ahead (se90.rs:70-77, se99.rs:82-89): min with strict <, NaN never taken
            double la = INFINITY;
            {
                const double ljj = W[(size_t)j * ld + j];
                for (int i = j + 1 + threadIdx.x; i < n; i += kThreads) {
                    const double lij = W[(size_t)i * ld + j];
                    la = fmin(la, dsub(W[(size_t)i * ld + i], ddiv(dmul(lij, lij), ljj)));
                }
            }
            la = block_reduce_minmax<false>(la, s.red);
            const double thresh = (variant == kSE99) ? dmul(-mu, gamma) : taugamma;
            if (la < thresh) {
                phaseone = false;
                break;
            }
            cta_se_eliminate(s, n, ld, j);
            ++j;
        }

        const int kstart = phaseone ? -1 : j;
        double delta_prev = 0.0;
        __syncthreads();

        if (variant == kSE99 && !phaseone && j == n - 1) {  // se99.rs:115-119
            if (threadIdx.x == 0) {
                const double ljj = W[(size_t)j * ld + j];
                const double ej = tail1x1_e(ljj, gamma);
                s.ev[j] = ej;
                W[(size_t)j * ld + j] = dsqrt(dadd(ljj, ej));
            }
        }

        if (!phaseone && (variant == kSE90 || j < n - 1)) {
            const int k = j;
            // lower Gershgorin bounds of the trailing block (se90.rs:105-110, se99.rs:125-130)
            for (int i = k + threadIdx.x; i < n; i += kThreads) {
                const double s1 = nd_sum_thread<true>(&W[(size_t)i * ld + k], 1, i - k);
                const double s2 = nd_sum_thread<true>(&W[(size_t)(i + 1) * ld + i], ld, n - 1 - i);
                s.g[i] = dsub(dsub(W[(size_t)i * ld + i], s1), s2);
            }
            __syncthreads();
            for (int jj = k; jj + 2 < n; ++jj) {
                // pivot on the largest bound (se90.rs:115-121, se99.rs:135-141)
                const int m = block_argmax_first(jj, n, [&](int t) { return s.g[t]; }, s.red);
                if (m != jj) {
                    __syncthreads();
                    cta_swap(W, n, ld, jj, m);
                    if (threadIdx.x == 0) {
                        int t = s.perm[jj];
                        s.perm[jj] = s.perm[m];
                        s.perm[m] = t;
                        double tg = s.g[jj];
                        s.g[jj] = s.g[m];
                        s.g[m] = tg;
                    }
                }
                __syncthreads();
                for (int r = jj + 1 + threadIdx.x; r < n; r += kThreads)
                    s.colv[r] = W[(size_t)r * ld + jj];
                __syncthreads();
                // E_jj (se90.rs:124-131, se99.rs:144-151)
                const double normj = cta_nd_sum_abs_colv(s, jj + 1, n);
                double ljj = W[(size_t)jj * ld + jj];
                const double ej = phase2_e(ljj, normj, taugamma, delta_prev);
                if (ej > 0.0) {
                    ljj = dadd(ljj, ej);
                    delta_prev = ej;
                }
                __syncthreads();
                if (threadIdx.x == 0) {
                    s.ev[jj] = ej;
                    W[(size_t)jj * ld + jj] = ljj;
                }
                // bound update (se90.rs:134-139, se99.rs:154-159)
                if (fabs(dsub(ljj, normj)) > MODCHOL_EPS) {
                    const double t = dsub(1.0, ddiv(normj, ljj));
                    for (int i = jj + 1 + threadIdx.x; i < n; i += kThreads)
                        s.g[i] = dadd(s.g[i], dmul(fabs(s.colv[i]), t));
                }
                __syncthreads();
                cta_se_eliminate(s, n, ld, jj);
            }
            // final 2x2 (se90.rs:155-166, se99.rs:175-186)
            if (threadIdx.x == 0) {
                double& a00 = W[(size_t)(n - 2) * ld + (n - 2)];
                double& a01 = W[(size_t)(n - 2) * ld + (n - 1)];
                double& a10 = W[(size_t)(n - 1) * ld + (n - 2)];
                double& a11 = W[(size_t)(n - 1) * ld + (n - 1)];
                double lhi, llo;
                eigenvalues_2x2(a00, a01, a10, a11, lhi, llo);
                const double en = tail2x2_e(variant, lhi, llo, gamma, delta_prev);
                s.ev[n - 2] = en;
                s.ev[n - 1] = en;
                if (en > 0.0) {
                    a00 = dadd(a00, en);
                    a11 = dadd(a11, en);
                }
                a00 = dsqrt(a00);
                a10 = ddiv(a10, a00);
                a11 = dsqrt(dsub(a11, dmul(a10, a10)));
            }
        }
        __syncthreads();

        // epilogue: zero the strict upper triangle, un-permute E, emit p
        // (se90.rs:169-178, se99.rs:189-198)
        for (size_t idx = threadIdx.x; idx < nn; idx += kThreads) {
            const int r = (int)(idx / n), c = (int)(idx - (size_t)r * n);
            Lb[idx] = (c <= r) ? W[(size_t)r * ld + c] : 0.0;
        }
        for (int i = threadIdx.x; i < n; i += kThreads) {
            E[(size_t)b * n + s.perm[i]] = s.ev[i];
            P[(size_t)b * n + i] = (unsigned long long)s.perm[i];
        }
        if (K && threadIdx.x == 0)
            K[b] = kstart;
    }
}

// ---------------------------------------------------------------------------------------
// GMW81   (gmw81.rs:39-134)
// ---------------------------------------------------------------------------------------
// Single working array W:  trailing columns (>= j) hold the reference's `l` (the permuted
// input, never modified there); for a finished column s the values c_is (i > s) are kept
// TRANSPOSED in the dead upper triangle, W[s][i], so that the thread-per-row dot products of
// gmw81.rs:84 read consecutive addresses; W[i][s] (lower) receives the final
// l_is = c_is / d_s when row i becomes the pivot row (gmw81.rs:79-81).  The running
// diagonal of c (gmw81.rs:104-109) lives in s.g.  Full row+column swaps keep both copies
// consistent with the reference's swaps of `c` and `l` (gmw81.rs:73-76).
template <int kThreads>
__global__ void __launch_bounds__(kThreads)
gmw81_cta_kernel(int n, int ld, int use_smem, long long batch, const double* __restrict__ A,
                 double* __restrict__ L, double* __restrict__ E,
                 unsigned long long* __restrict__ P, int* __restrict__ K)
{
    extern __shared__ double sm[];
    const size_t nn = (size_t)n * n;

    for (long long b = blockIdx.x; b < batch; b += gridDim.x) {
        const double* Ab = A + (size_t)b * nn;
        double* Lb = L + (size_t)b * nn;
        CtaShared s = carve_shared(sm, n, ld, use_smem != 0, Lb);
        double* W = s.W;
        double* cd = s.g;   // running diag(c)
        double* ljs = s.colv;  // l_{j,0..j-1}

        __syncthreads();
        for (size_t idx = threadIdx.x; idx < nn; idx += kThreads) {
            const int r = (int)(idx / n), c = (int)(idx - (size_t)r * n);
            W[(size_t)r * ld + c] = Ab[idx];
        }
        for (int i = threadIdx.x; i < n; i += kThreads) {
            s.ev[i] = 0.0;
            s.perm[i] = i;
            s.dv[i] = 0.0;
        }
        __syncthreads();

        // gmw81.rs:49-64
        double dm = 0.0, om = 0.0;
        for (size_t idx = threadIdx.x; idx < nn; idx += kThreads) {
            const int r = (int)(idx / n), c = (int)(idx - (size_t)r * n);
            const double x = fabs(W[(size_t)r * ld + c]);
            if (r == c)
                dm = fmax(dm, x);
            else
                om = fmax(om, x);
        }
        const double diag_max = block_reduce_minmax<true>(dm, s.red);
        const double off_max = block_reduce_minmax<true>(om, s.red);
        const double nf = (double)n;
        const double delta = dmul(MODCHOL_EPS, fmax(1.0, dadd(diag_max, off_max)));
        const double beta = dsqrt(
            fmax(fmax(diag_max, ddiv(off_max, dsqrt(dsub(dmul(nf, nf), 1.0)))), MODCHOL_EPS));
        for (int i = threadIdx.x; i < n; i += kThreads)
            cd[i] = W[(size_t)i * ld + i];  // gmw81.rs:66-67
        __syncthreads();

        for (int j = 0; j < n; ++j) {
            // gmw81.rs:71-78
            const int m = block_argmax_first(j, n, [&](int t) { return fabs(cd[t]); }, s.red);
            if (m != j) {
                __syncthreads();
                cta_swap(W, n, ld, j, m);
                if (threadIdx.x == 0) {
                    int t = s.perm[j];
                    s.perm[j] = s.perm[m];
                    s.perm[m] = t;
                    double tc = cd[j];
                    cd[j] = cd[m];
                    cd[m] = tc;
                }
            }
            __syncthreads();
            // gmw81.rs:79-81: l_js = c_js / d_s
            for (int t = threadIdx.x; t < j; t += kThreads) {
                const double v = ddiv(W[(size_t)t * ld + j], s.dv[t]);
                ljs[t] = v;
                W[(size_t)j * ld + t] = v;
            }
            __syncthreads();
            // gmw81.rs:83-85: c_ij = l_ij - sum_s(l_js * c_is), ndarray summation order
            double theta = 0.0;
            for (int i = j + threadIdx.x; i < n; i += kThreads) {
                double p[8];
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    p[u] = 0.0;
                int t = 0;
                for (; t + 8 <= j; t += 8) {
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        p[u] = dadd(p[u], dmul(ljs[t + u], W[(size_t)(t + u) * ld + i]));
                }
                double acc = 0.0;
                acc = dadd(acc, dadd(p[0], p[4]));
                acc = dadd(acc, dadd(p[1], p[5]));
                acc = dadd(acc, dadd(p[2], p[6]));
                acc = dadd(acc, dadd(p[3], p[7]));
                for (; t < j; ++t)
                    acc = dadd(acc, dmul(ljs[t], W[(size_t)t * ld + i]));
                const double cij = dsub(W[(size_t)i * ld + j], acc);
                if (i == j) {
                    cd[j] = cij;  // the recomputed c_jj replaces the running value
                } else {
                    W[(size_t)j * ld + i] = cij;  // transposed store (see header)
                    theta = fmax(theta, fabs(cij));  // gmw81.rs:87-100
                }
            }
            theta = block_reduce_minmax<true>(theta, s.red);
            const double cjj = cd[j];
            const double tb = ddiv(theta, beta);
            const double dj = fmax(fmax(fabs(cjj), dmul(tb, tb)), delta);  // gmw81.rs:102
            // gmw81.rs:104-109
            for (int i = j + 1 + threadIdx.x; i < n; i += kThreads) {
                const double cij = W[(size_t)j * ld + i];
                cd[i] = dsub(cd[i], ddiv(dmul(cij, cij), dj));
            }
            if (threadIdx.x == 0) {
                s.dv[j] = dj;
                s.ev[j] = dsub(dj, cjj);  // gmw81.rs:110
            }
            __syncthreads();
        }

        // gmw81.rs:112-125: L = tril(l, unit diagonal) . diag(sqrt(d)); the dense dot adds
        // exact zeros to the single product, which only turns a -0 into +0.
        for (int i = threadIdx.x; i < n; i += kThreads)
            s.colv[i] = dsqrt(s.dv[i]);
        __syncthreads();
        for (size_t idx = threadIdx.x; idx < nn; idx += kThreads) {
            const int r = (int)(idx / n), c = (int)(idx - (size_t)r * n);
            double v = 0.0;
            if (c < r)
                v = dmul(W[(size_t)r * ld + c], s.colv[c]);
            else if (c == r)
                v = dmul(1.0, s.colv[c]);
            Lb[idx] = dadd(0.0, v);
        }
        for (int i = threadIdx.x; i < n; i += kThreads) {
            E[(size_t)b * n + s.perm[i]] = s.ev[i];  // gmw81.rs:127-131
            P[(size_t)b * n + i] = (unsigned long long)s.perm[i];
        }
        if (K && threadIdx.x == 0)
            K[b] = -1;
    }
}

}  // namespace modchol
