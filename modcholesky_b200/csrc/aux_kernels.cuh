// aux_kernels.cuh -- the kernels either side of the factorisation:
//   * batched GershgorinCircles (gershgorin.rs:20-41)
//   * batched acceptance check  P (A + diag(e)) P^T == L L^T  (se99.rs:221-231,
//     utils.rs:94-113), SURVEY.md section 8(f) row 1.
#pragma once
#include "common.cuh"

namespace modchol {

// One thread per (matrix, row).  The reference sums sequentially over j != i in index
// order (gershgorin.rs:31-37); so does each thread, which keeps the result bit-identical.
// Column reads (a[j][i]) are coalesced across the threads of a matrix; row reads walk
// 128-byte lines that stay in L1 for the sixteen iterations that consume them.
__global__ void __launch_bounds__(256)
gershgorin_circles_kernel(long long batch, int n, const double* __restrict__ A,
                          double* __restrict__ out)
{
    const long long total = batch * (long long)n;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long b = t / n;
        const int i = (int)(t - b * n);
        const double* a = A + (size_t)b * n * n;
        double ri = 0.0, ci = 0.0;
        for (int j = 0; j < n; ++j) {
            if (j == i)
                continue;
            ri = dadd(ri, fabs(a[(size_t)i * n + j]));
            ci = dadd(ci, fabs(a[(size_t)j * n + i]));
        }
        out[2 * t] = a[(size_t)i * n + i];
        out[2 * t + 1] = fmin(ri, ci);
    }
}

// One CTA per matrix.  resid = max_ij |(L L^T)_ij - A[p_i][p_j] - [i==j] e[p_i]| / max(1, max|A|).
__global__ void __launch_bounds__(256)
verify_kernel(long long batch, int n, const double* __restrict__ A, const double* __restrict__ L,
              const double* __restrict__ E, const unsigned long long* __restrict__ P,
              double* __restrict__ resid, int* __restrict__ flags)
{
    __shared__ double red[66];
    __shared__ int sflag;
    const size_t nn = (size_t)n * n;
    for (long long b = blockIdx.x; b < batch; b += gridDim.x) {
        const double* a = A + (size_t)b * nn;
        const double* l = L + (size_t)b * nn;
        const double* e = E + (size_t)b * n;
        const unsigned long long* p = P + (size_t)b * n;
        if (threadIdx.x == 0)
            sflag = 0;
        __syncthreads();
        int f = 0;
        double amax = 0.0, err = 0.0;
        // permutation validity: every index hit exactly once (O(n^2/T), fine for a checker)
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            int hits = 0;
            for (int q = 0; q < n; ++q)
                hits += (p[q] == (unsigned long long)i);
            if (hits != 1)
                f |= 1;
            const double ei = e[i];
            if (ei < 0.0)
                f |= 2;
            if (!(fabs(ei) <= 1.79769313486231570815e308))
                f |= 8;
        }
        if (f)
            atomicOr(&sflag, f);
        __syncthreads();
        const bool perm_ok = (sflag & 1) == 0;
        f = 0;
        for (size_t idx = threadIdx.x; idx < nn; idx += blockDim.x) {
            const int i = (int)(idx / n), j = (int)(idx - (size_t)i * n);
            const double lij = l[idx];
            if (j > i && lij != 0.0)
                f |= 4;
            if (!(fabs(lij) <= 1.79769313486231570815e308))
                f |= 8;
            amax = fmax(amax, fabs(a[idx]));
            if (perm_ok) {
                const int m = i < j ? i : j;
                double acc = 0.0;
                for (int k = 0; k <= m; ++k)
                    acc = __fma_rn(l[(size_t)i * n + k], l[(size_t)j * n + k], acc);
                const size_t pi = (size_t)p[i], pj = (size_t)p[j];
                double rhs = a[pi * n + pj];
                if (i == j)
                    rhs += e[pi];
                const double d = fabs(acc - rhs);
                err = (d != d) ? INFINITY : fmax(err, d);
            }
        }
        if (f)
            atomicOr(&sflag, f);
        amax = block_reduce_minmax<true>(amax, red);
        err = block_reduce_minmax<true>(err, red);
        __syncthreads();
        if (threadIdx.x == 0) {
            if (resid)
                resid[b] = perm_ok ? err / fmax(1.0, amax) : INFINITY;
            if (flags)
                flags[b] = sflag;
        }
        __syncthreads();
    }
}

// ---- warp-per-matrix verifier, n <= 8 NB <= 64 -------------------------------------------
// Same contract as verify_kernel (flags: 1 permutation invalid, 2 e < 0, 4 strict upper of L not
// zero, 8 non-finite entry of L or e).  L L^T is formed on the FP64 tensor pipe, 8x8 tile by
// 8x8 tile (mma.sync.m8n8k4.f64): lane l of the warp holds L(8 bi + l/4, 4 h + l%4), which is
// both the A fragment of block row bi and the B fragment of block column bi, so every element
// of L is fetched from HBM once (the block rows above are re-read from L1).  A is streamed once
// for max|A| and then gathered through p (L1/L2 hits); one pass over HBM per matrix:
// 16 n^2 + 16 n bytes, like the factorisation it checks.
__device__ __forceinline__ double warp_max_nonneg(double v)   // v >= 0 or +Inf, never NaN
{
    const unsigned hi = __reduce_max_sync(0xffffffffu, (unsigned)__double2hiint(v));
    const unsigned lo = __reduce_max_sync(0xffffffffu, (unsigned)__double2hiint(v) == hi ? (unsigned)__double2loint(v) : 0u);
    return __hiloint2double((int)hi, (int)lo);
}

template <int NB>
__global__ void __launch_bounds__(128)
verify_warp_kernel(long long batch, int n, const double* __restrict__ A, const double* __restrict__ L,
                   const double* __restrict__ E, const unsigned long long* __restrict__ P,
                   double* __restrict__ resid, int* __restrict__ flags)
{
    constexpr unsigned kFullMask = 0xffffffffu;
    constexpr double kDblMax = 1.79769313486231570815e308;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int r8 = lane >> 2, c4 = lane & 3;
    const size_t nn = (size_t)n * n;
    const bool vec2 = (nn & 1) == 0 && ((size_t)A & 15) == 0;
    for (long long b = (long long)blockIdx.x * wpb + wib; b < batch; b += (long long)gridDim.x * wpb) {
        const double* a = A + (size_t)b * nn;
        const double* l = L + (size_t)b * nn;
        const double* e = E + (size_t)b * n;
        const unsigned long long* p = P + (size_t)b * n;
        int f = 0;
        // ---- permutation: n indices below n that set n distinct bits; e >= 0 and finite ------
        unsigned mlo = 0u, mhi = 0u;
        bool inrange = true;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int i = lane + 32 * t;
            if (i < n) {
                const unsigned long long q = p[i];
                inrange = inrange && q < (unsigned long long)n;
                if (q < 32ull) mlo |= 1u << (unsigned)q;
                else if (q < 64ull) mhi |= 1u << (unsigned)(q - 32ull);
                const double ei = e[i];
                if (ei < 0.0) f |= 2;
                if (!(fabs(ei) <= kDblMax)) f |= 8;
            }
        }
        mlo = __reduce_or_sync(kFullMask, mlo);
        mhi = __reduce_or_sync(kFullMask, mhi);
        const unsigned wlo = n >= 32 ? 0xffffffffu : ((1u << n) - 1u);
        const unsigned whi = n <= 32 ? 0u : (n >= 64 ? 0xffffffffu : ((1u << (n - 32)) - 1u));
        const bool perm_ok = __all_sync(kFullMask, inrange) && mlo == wlo && mhi == whi;
        if (!perm_ok) f |= 1;
        // ---- max |A| (streamed, coalesced) -------------------------------------------------
        double amax = 0.0;
        if (vec2) {
            const double2* a2 = reinterpret_cast<const double2*>(a);
            for (size_t idx = lane; idx < nn / 2; idx += 32) {
                const double2 v = __ldg(a2 + idx);
                amax = fmax(amax, fmax(fabs(v.x), fabs(v.y)));
            }
        } else {
            for (size_t idx = lane; idx < nn; idx += 32)
                amax = fmax(amax, fabs(__ldg(a + idx)));
        }
        // ---- L L^T against P (A + E) P^T ---------------------------------------------------
        double err = 0.0;
#pragma unroll
        for (int bi = 0; bi < NB; ++bi) {
            if (8 * bi >= n)
                break;
            const int ri = 8 * bi + r8;                    // row of this lane's fragments
            double rowf[2 * NB];
#pragma unroll
            for (int h = 0; h < 2 * NB; ++h) {
                const int k = 4 * h + c4;
                double v = 0.0;
                if (ri < n && k < n) {
                    v = __ldg(l + (size_t)ri * n + k);
                    if (!(fabs(v) <= kDblMax)) f |= 8;
                    if (k > ri) {
                        if (v != 0.0) f |= 4;
                        v = 0.0;
                    }
                }
                rowf[h] = v;
            }
            const unsigned long long pi = (perm_ok && ri < n) ? p[ri] : 0ull;
#pragma unroll
            for (int bj = 0; bj <= bi; ++bj) {
                double c0 = 0.0, c1 = 0.0;
#pragma unroll
                for (int h = 0; h < 2 * bj + 2; ++h) {
                    double colf;
                    if (bj == bi) {
                        colf = rowf[h];
                    } else {
                        const int rj = 8 * bj + r8, k = 4 * h + c4;
                        colf = (k <= rj && k < n) ? __ldg(l + (size_t)rj * n + k) : 0.0;   // rj < ri < n or padded rows
                        if (rj >= n) colf = 0.0;
                    }
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                                 : "+d"(c0), "+d"(c1) : "d"(rowf[h]), "d"(colf));
                }
                if (perm_ok && ri < n) {
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        const int cj = 8 * bj + 2 * c4 + t;
                        if (cj <= ri) {
                            const unsigned long long pj = p[cj];
                            double r1 = __ldg(a + (size_t)pi * n + pj), r2 = __ldg(a + (size_t)pj * n + pi);
                            if (cj == ri) {
                                const double ee = e[pi];
                                r1 += ee;
                                r2 += ee;
                            }
                            const double acc = t ? c1 : c0;
                            const double d = fmax(fabs(acc - r1), fabs(acc - r2));
                            err = (d != d || acc != acc || r1 != r1 || r2 != r2) ? INFINITY : fmax(err, d);
                        }
                    }
                }
            }
        }
        f = __reduce_or_sync(kFullMask, f);
        amax = warp_max_nonneg(amax);
        err = warp_max_nonneg(err);
        if (lane == 0) {
            if (resid)
                resid[b] = perm_ok ? err / fmax(1.0, amax) : INFINITY;
            if (flags)
                flags[b] = f;
        }
    }
}

// ---- device-side assembly of the reference's bench inputs --------------------------------
// A = Q D Q^T with Q = H_0 H_1 ... H_{m-1}, H_r = I - 2 v_r v_r^T / (v_r^T v_r)
// (utils.rs:116-121 random_householder, utils.rs:126-147 random_diagonal, bench.rs:48-53).
// The reference multiplies dense matrices (O(n^3), last bits depend on the CPU's matrixmultiply
// micro-kernel); here each reflector is applied from both sides as a symmetric rank-2 update
//   H M H = M - tau (v u^T + u v^T) + tau^2 (v^T u) v v^T,   u = M v,  tau = 2 / v^T v,
// innermost reflector first (M = D is diagonal there, so the first update writes the matrix):
// O(m n^2) work, one CTA per matrix, M in place in the output.  The update is evaluated with a
// formula that is symmetric in (i, j) operation by operation, so A is exactly symmetric.
// Shared memory: 3 n + 66 doubles.
__global__ void __launch_bounds__(256)
assemble_qdqt_kernel(long long batch, int n, int m, const double* __restrict__ V,
                     const double* __restrict__ D, double* __restrict__ A)
{
    extern __shared__ double sm[];
    double* v = sm;
    double* u = sm + n;
    double* dd = sm + 2 * n;
    double* red = sm + 3 * n;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    const size_t nn = (size_t)n * n;
    auto block_sum = [&](double x) -> double {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            x += __shfl_xor_sync(0xffffffffu, x, o);
        __syncthreads();
        if (lane == 0)
            red[warp] = x;
        __syncthreads();
        double s = 0.0;
        for (int w = 0; w < nw; ++w)
            s += red[w];
        return s;
    };
    for (long long b = blockIdx.x; b < batch; b += gridDim.x) {
        double* M = A + (size_t)b * nn;
        __syncthreads();
        for (int i = tid; i < n; i += nt)
            dd[i] = D[(size_t)b * n + i];
        if (m == 0) {
            __syncthreads();
            for (size_t idx = tid; idx < nn; idx += nt) {
                const int i = (int)(idx / n), j = (int)(idx - (size_t)i * n);
                M[idx] = i == j ? dd[i] : 0.0;
            }
            continue;
        }
        for (int r = m - 1; r >= 0; --r) {
            const bool first = r == m - 1;
            __syncthreads();
            for (int i = tid; i < n; i += nt)
                v[i] = V[((size_t)b * m + r) * n + i];
            __syncthreads();
            if (first) {
                for (int i = tid; i < n; i += nt)
                    u[i] = dmul(dd[i], v[i]);
            } else {   // u = M v, one warp per row (coalesced)
                for (int i = warp; i < n; i += nw) {
                    double acc = 0.0;
                    for (int k = lane; k < n; k += 32)
                        acc = __fma_rn(M[(size_t)i * n + k], v[k], acc);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1)
                        acc += __shfl_xor_sync(0xffffffffu, acc, o);
                    if (lane == 0)
                        u[i] = acc;
                }
            }
            __syncthreads();
            double pv = 0.0, pu = 0.0;
            for (int i = tid; i < n; i += nt) {
                pv = __fma_rn(v[i], v[i], pv);
                pu = __fma_rn(v[i], u[i], pu);
            }
            const double vv = block_sum(pv);
            const double vu = block_sum(pu);
            const double tau = ddiv(2.0, vv);
            const double t2s = dmul(dmul(tau, tau), vu);
            for (size_t idx = tid; idx < nn; idx += nt) {
                const int i = (int)(idx / n), j = (int)(idx - (size_t)i * n);
                const double base = first ? (i == j ? dd[i] : 0.0) : M[idx];
                const double cross = dadd(dmul(v[i], u[j]), dmul(u[i], v[j]));
                const double upd = dsub(dmul(t2s, dmul(v[i], v[j])), dmul(tau, cross));
                M[idx] = dadd(base, upd);
            }
        }
    }
}

// ---- FP64 issue-rate probes (bench.py reports the large-n kernels against these) ------------
// Dependent chains of fused multiply-adds (8 per thread) / of m8n8k4 tensor-pipe operations
// (4 per warp); the host entry point times them with CUDA events and converts to TFLOP/s.
__global__ void __launch_bounds__(512)
fp64_dfma_probe_kernel(int iters, double x, double* __restrict__ sink)
{
    double acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
        acc[u] = (double)(threadIdx.x + u);
    const double y = x * 0.5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
            acc[u] = __fma_rn(acc[u], x, y);
    }
    double s = 0.0;
#pragma unroll
    for (int u = 0; u < 8; ++u)
        s += acc[u];
    if (s == 12345.678)
        sink[0] = s;
}

__global__ void __launch_bounds__(512)
fp64_dmma_probe_kernel(int iters, double x, double* __restrict__ sink)
{
    double c[4][2];
#pragma unroll
    for (int u = 0; u < 4; ++u)
        c[u][0] = c[u][1] = (double)u;
    const double a = x, bb = x * 0.25;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                         : "+d"(c[u][0]), "+d"(c[u][1]) : "d"(a), "d"(bb));
    }
    double s = 0.0;
#pragma unroll
    for (int u = 0; u < 4; ++u)
        s += c[u][0] + c[u][1];
    if (s == 12345.678)
        sink[0] = s;
}

// Batched solve (A + E) x = b from a factorisation {l, p}:  P (A+E) P^T = L L^T with
// P[i, p[i]] = 1, so  L L^T (P x) = P b.  One warp per (matrix, right-hand side):
//   y_i = b[p_i];  forward  L z = y  in dot form (row i of L: contiguous, coalesced);
//   backward  L^T w = z  in axpy form (again row i of L);  x[p_i] = w_i.
// The working vector lives in shared memory (n doubles per warp).
__global__ void __launch_bounds__(128)
solve_kernel(long long batch, int n, int nrhs, const double* __restrict__ L,
             const unsigned long long* __restrict__ P, const double* B, double* X)   // X may alias B
{
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    double* z = sm + (size_t)wib * n;
    const long long total = batch * (long long)nrhs;
    for (long long t = (long long)blockIdx.x * wpb + wib; t < total; t += (long long)gridDim.x * wpb) {
        const long long b = t / nrhs;
        const double* l = L + (size_t)b * n * n;
        const unsigned long long* p = P + (size_t)b * n;
        const double* rhs = B + (size_t)t * n;
        double* x = X + (size_t)t * n;
        __syncwarp();
        for (int i = lane; i < n; i += 32)
            z[i] = rhs[p[i]];
        __syncwarp();
        for (int i = 0; i < n; ++i) {  // forward substitution
            double acc = 0.0;
            for (int k = lane; k < i; k += 32)
                acc = __fma_rn(l[(size_t)i * n + k], z[k], acc);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
                acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0)
                z[i] = (z[i] - acc) / l[(size_t)i * n + i];
            __syncwarp();
        }
        for (int i = n - 1; i >= 0; --i) {  // backward substitution
            const double wi = z[i] / l[(size_t)i * n + i];
            __syncwarp();
            for (int k = lane; k < i; k += 32)
                z[k] = __fma_rn(-l[(size_t)i * n + k], wi, z[k]);
            if (lane == 0)
                z[i] = wi;
            __syncwarp();
        }
        for (int i = lane; i < n; i += 32)
            x[p[i]] = z[i];
    }
}

}  // namespace modchol
