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
