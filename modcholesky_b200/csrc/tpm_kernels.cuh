// tpm_kernels.cuh -- ONE THREAD PER MATRIX for the smallest sizes (n <= kTpmMaxN): GMW81, SE90
// and SE99 exactly as the reference executes them (gmw81.rs:39-134, se90.rs:38-181,
// se99.rs:35-201, utils.rs:14-91), one sequential instruction stream per matrix, no collective
// anywhere.  At these sizes a warp per matrix (or per 2 / 4 / 8 matrices) spends its time in
// shuffles and reductions over a handful of elements; here every lane does useful arithmetic and
// the batch streams through HBM.
//
// Where the data lives: a block of kTpmThreads threads takes a TILE of kTpmThreads consecutive
// matrices.  The tile is one contiguous range of A, so it is read with fully coalesced loads and
// scattered into shared memory matrix by matrix, thread t's matrix at doubles [t S, t S + n n)
// with S = n n | 1: an odd stride, so (a) every thread touching the same element (i, k) of ITS
// matrix -- the common case -- is conflict free, and (b) the block-wide copies are too (word
// q of the tile lands at q + q / (n n)).  Each thread then works on the full n x n matrix in
// place like the reference does on its clone (both triangles kept, swaps of full rows and columns:
// the arithmetic is the reference's element for element, hence bit-identical results in STRICT
// and FAST mode alike), with its vectors (g or d, e, p) in three more regions of stride n | 1.
// The finished tile of L, e (un-permuted) and p leaves through the same coalesced copies.
//
// All sums are ndarray's `Array1::sum()`: for fewer than eight terms that is the plain
// left-to-right sum from 0.0 (numeric_util::unrolled_fold, see common.cuh) -- n <= 8 never reaches
// eight terms, which is what bounds kTpmMaxN.
#pragma once
#include "common.cuh"

namespace modchol {

constexpr int kTpmThreads = 128;
constexpr int kTpmMaxN = 8;

__host__ __device__ constexpr int tpm_mat_stride(int n) { return (n * n) | 1; }
__host__ __device__ constexpr int tpm_vec_stride(int n) { return n | 1; }
// doubles of shared memory per thread: the matrix (GMW81: l and c), then g / d, e, p
__host__ __device__ constexpr int tpm_thread_doubles(int algo, int n)
{
    return (algo == kGMW81 ? 2 : 1) * tpm_mat_stride(n) + 3 * tpm_vec_stride(n);
}

// utils.rs:41-49 swap_rows followed by utils.rs:30-38 swap_columns on the full matrix
__device__ __forceinline__ void tpm_swap(double* m, int n, int r1, int r2)
{
    for (int c = 0; c < n; ++c) {
        const double t = m[r1 * n + c];
        m[r1 * n + c] = m[r2 * n + c];
        m[r2 * n + c] = t;
    }
    for (int r = 0; r < n; ++r) {
        const double t = m[r * n + r1];
        m[r * n + r1] = m[r * n + r2];
        m[r * n + r2] = t;
    }
}

// one right-looking elimination step, se90.rs:84-93 == :142-151 == se99.rs:96-105 == :162-171
__device__ __forceinline__ void tpm_eliminate(double* l, int n, int j)
{
    const double ljj = dsqrt(l[j * n + j]);
    l[j * n + j] = ljj;
    for (int i = j + 1; i < n; ++i) {
        const double lij = ddiv(l[i * n + j], ljj);
        l[i * n + j] = lij;
        l[j * n + i] = ddiv(l[j * n + i], ljj);
        for (int k = j + 1; k <= i; ++k) {
            const double v = dsub(l[i * n + k], dmul(lij, l[k * n + j]));
            l[i * n + k] = v;
            l[k * n + i] = v;
        }
    }
}

// SE90 / SE99 on one matrix; returns the column phase 2 started at (-1: never)
template <int kAlgo>
__device__ __forceinline__ int tpm_se(int n, double* l, double* g, double* e, unsigned long long* p)
{
    double gamma = 0.0;   // se90.rs:55-57, se99.rs:54-56
    for (int i = 0; i < n; ++i) {
        e[i] = 0.0;
        p[i] = (unsigned long long)i;
        const double d = fabs(l[i * n + i]);
        if (d > gamma)
            gamma = d;
    }
    const double taugamma = dmul(MODCHOL_TAU, gamma);
    bool phaseone = true;
    int j = 0;
    while (j < n) {   // phase 1: se90.rs:61-96, se99.rs:61-109
        if (kAlgo == kSE99) {   // se99.rs:62-73
            double dmax = -INFINITY, dmin = INFINITY;
            for (int i = j; i < n; ++i) {
                const double x = l[i * n + i];
                if (x > dmax) dmax = x;
                if (x < dmin) dmin = x;
            }
            if (dmax < taugamma || dmin < dmul(-MODCHOL_MU, dmax)) {
                phaseone = false;
                break;
            }
        }
        int m = j;   // utils.rs:52-70: first strict maximum of the diagonal, a NaN head sticks
        double best = l[j * n + j];
        for (int i = j + 1; i < n; ++i) {
            const double c = l[i * n + i];
            if (c > best) {
                best = c;
                m = i;
            }
        }
        if (m != j) {   // se90.rs:63-68, se99.rs:76-81
            tpm_swap(l, n, j, m);
            const unsigned long long t = p[j];
            p[j] = p[m];
            p[m] = t;
        }
        double la = INFINITY;   // se90.rs:70-77, se99.rs:82-89
        const double ljj = l[j * n + j];
        for (int i = j + 1; i < n; ++i) {
            const double lij = l[i * n + j];
            const double nv = dsub(l[i * n + i], ddiv(dmul(lij, lij), ljj));
            if (nv < la)
                la = nv;
        }
        const double thresh = kAlgo == kSE99 ? dmul(-MODCHOL_MU, gamma) : taugamma;   // se99.rs:90 / se90.rs:78
        if (la < thresh) {
            phaseone = false;
            break;
        }
        tpm_eliminate(l, n, j);
        ++j;
    }
    const int kstart = phaseone ? -1 : j;
    double delta_prev = 0.0;
    if (kAlgo == kSE99 && !phaseone && j == n - 1) {   // se99.rs:115-119
        const double ljj = l[j * n + j];
        const double ej = tail1x1_e(ljj, gamma);
        e[j] = ej;
        l[j * n + j] = dsqrt(dadd(ljj, ej));
    }
    if (!phaseone && (kAlgo == kSE90 || j + 1 < n)) {   // se90.rs:101-167, se99.rs:121-187
        const int k = j;
        for (int i = 0; i < n; ++i)
            g[i] = 0.0;
        for (int i = k; i < n; ++i) {   // lower Gershgorin bounds, se90.rs:105-110, se99.rs:125-130
            double s1 = 0.0, s2 = 0.0;
            for (int q = k; q < i; ++q)
                s1 = dadd(s1, fabs(l[i * n + q]));
            for (int r = i + 1; r < n; ++r)
                s2 = dadd(s2, fabs(l[r * n + i]));
            g[i] = dsub(dsub(l[i * n + i], s1), s2);
        }
        for (int jj = k; jj + 2 < n; ++jj) {
            int m = jj;   // se90.rs:115-121, se99.rs:135-141
            double best = g[jj];
            for (int i = jj + 1; i < n; ++i) {
                const double c = g[i];
                if (c > best) {
                    best = c;
                    m = i;
                }
            }
            if (m != jj) {
                tpm_swap(l, n, jj, m);
                const unsigned long long t = p[jj];
                p[jj] = p[m];
                p[m] = t;
                const double tg = g[jj];
                g[jj] = g[m];
                g[m] = tg;
            }
            double normj = 0.0;   // se90.rs:124-131, se99.rs:144-151
            for (int r = jj + 1; r < n; ++r)
                normj = dadd(normj, fabs(l[r * n + jj]));
            double ljj = l[jj * n + jj];
            const double ej = phase2_e(ljj, normj, taugamma, delta_prev);
            e[jj] = ej;
            if (ej > 0.0) {
                ljj = dadd(ljj, ej);
                l[jj * n + jj] = ljj;
                delta_prev = ej;
            }
            if (fabs(dsub(ljj, normj)) > MODCHOL_EPS) {   // se90.rs:134-139, se99.rs:154-159
                const double t = dsub(1.0, ddiv(normj, ljj));
                for (int i = jj + 1; i < n; ++i)
                    g[i] = dadd(g[i], dmul(fabs(l[i * n + jj]), t));
            }
            tpm_eliminate(l, n, jj);
        }
        // final 2x2 block, never pivoted: se90.rs:155-166, se99.rs:175-186
        const int a = (n - 2) * n + (n - 2), b = (n - 2) * n + (n - 1), c = (n - 1) * n + (n - 2),
                  d = (n - 1) * n + (n - 1);
        double lhi, llo;
        eigenvalues_2x2(l[a], l[b], l[c], l[d], lhi, llo);
        const double en = tail2x2_e(kAlgo, lhi, llo, gamma, delta_prev);
        e[n - 2] = en;
        e[n - 1] = en;
        if (en > 0.0) {
            l[a] = dadd(l[a], en);
            l[d] = dadd(l[d], en);
        }
        const double l00 = dsqrt(l[a]);
        l[a] = l00;
        const double l10 = ddiv(l[c], l00);
        l[c] = l10;
        l[d] = dsqrt(dsub(l[d], dmul(l10, l10)));
    }
    for (int i = 0; i + 1 < n; ++i)   // se90.rs:169-173, se99.rs:189-192
        for (int c = i + 1; c < n; ++c)
            l[i * n + c] = 0.0;
    return kstart;
}

// GMW81 on one matrix (gmw81.rs:39-134); c: the second matrix, d: n doubles
__device__ __forceinline__ void tpm_gmw(int n, double* l, double* c, double* d, double* e,
                                        unsigned long long* p)
{
    double diag_max = 0.0, off_max = 0.0;   // gmw81.rs:49-58
    for (int i = 0; i < n; ++i) {
        e[i] = 0.0;
        p[i] = (unsigned long long)i;
        for (int q = 0; q < n; ++q) {
            const double v = fabs(l[i * n + q]);
            if (i == q) {
                if (v > diag_max) diag_max = v;
            } else if (v > off_max) {
                off_max = v;
            }
            c[i * n + q] = i == q ? l[i * n + q] : 0.0;   // gmw81.rs:66-68
        }
        d[i] = 0.0;
    }
    const double nf = (double)n;
    const double delta = dmul(MODCHOL_EPS, fmax(1.0, dadd(diag_max, off_max)));   // gmw81.rs:60
    const double beta =
        dsqrt(fmax(fmax(diag_max, ddiv(off_max, dsqrt(dsub(dmul(nf, nf), 1.0)))), MODCHOL_EPS));   // :61-64
    for (int j = 0; j < n; ++j) {
        int m = j;   // utils.rs:73-91 on the diagonal of c (gmw81.rs:71-78)
        double best = fabs(c[j * n + j]);
        for (int i = j + 1; i < n; ++i) {
            const double v = fabs(c[i * n + i]);
            if (v > best) {
                best = v;
                m = i;
            }
        }
        if (m != j) {
            tpm_swap(c, n, j, m);
            tpm_swap(l, n, j, m);
            const unsigned long long t = p[j];
            p[j] = p[m];
            p[m] = t;
        }
        for (int s = 0; s < j; ++s)   // gmw81.rs:79-81
            l[j * n + s] = ddiv(c[j * n + s], d[s]);
        for (int i = j; i < n; ++i) {   // gmw81.rs:83-85
            double acc = 0.0;
            for (int s = 0; s < j; ++s)
                acc = dadd(acc, dmul(l[j * n + s], c[i * n + s]));
            c[i * n + j] = dsub(l[i * n + j], acc);
        }
        double theta = 0.0;   // gmw81.rs:87-100
        for (int i = j + 1; i < n; ++i) {
            const double v = fabs(c[i * n + j]);
            if (v > theta)
                theta = v;
        }
        const double tb = ddiv(theta, beta);
        const double cjj = c[j * n + j];
        const double dj = fmax(fmax(fabs(cjj), dmul(tb, tb)), delta);   // gmw81.rs:102
        d[j] = dj;
        for (int i = j + 1; i < n; ++i) {   // gmw81.rs:104-109
            const double cij = c[i * n + j];
            c[i * n + i] = dsub(c[i * n + i], ddiv(dmul(cij, cij), dj));
        }
        e[j] = dsub(dj, cjj);   // gmw81.rs:110
    }
    // gmw81.rs:112-125: unit diagonal, zero upper triangle, l . diag(sqrt(d)) -- one non-zero
    // term per element of the dense product; its accumulation from 0.0 turns a -0 into +0
    for (int i = 0; i < n; ++i)
        for (int q = 0; q < n; ++q) {
            double v = 0.0;
            if (q == i)
                v = dsqrt(d[q]);
            else if (q < i)
                v = dmul(l[i * n + q], dsqrt(d[q]));
            l[i * n + q] = dadd(0.0, v);
        }
}

template <int kAlgo>
__global__ void __launch_bounds__(kTpmThreads)
tpm_kernel(int n, long long batch, const double* __restrict__ A, double* __restrict__ L,
           double* __restrict__ E, unsigned long long* __restrict__ P, int* __restrict__ K)
{
    extern __shared__ double sm[];
    const int tid = threadIdx.x;
    const int nn = n * n, S = tpm_mat_stride(n), V = tpm_vec_stride(n);
    double* mats = sm;                                   // kTpmThreads x S
    double* cmat = mats + kTpmThreads * S;               // GMW81 only
    double* gvec = cmat + (kAlgo == kGMW81 ? kTpmThreads * S : 0);
    double* evec = gvec + kTpmThreads * V;
    unsigned long long* pvec = reinterpret_cast<unsigned long long*>(evec + kTpmThreads * V);
    // word q of a tile region with `w` words per matrix lives at q + q / w (odd stride w | 1 ...
    // for even w; for odd w the stride is w itself): advance (matrix, element) by kTpmThreads words
    const int step_m = kTpmThreads / nn, step_e = kTpmThreads % nn;
    const int step_mv = kTpmThreads / n, step_ev = kTpmThreads % n;

    for (long long t0 = (long long)blockIdx.x * kTpmThreads; t0 < batch; t0 += (long long)gridDim.x * kTpmThreads) {
        const int cnt = (int)min((long long)kTpmThreads, batch - t0);
        const int words = cnt * nn;
        // ---- in: coalesced over the tile, matrix by matrix into shared memory ------------------
        {
            const double* src = A + (size_t)t0 * nn;
            int mi = tid / nn, ei = tid % nn;
            for (int q = tid; q < words; q += kTpmThreads) {
                mats[mi * S + ei] = __ldg(src + q);
                mi += step_m;
                ei += step_e;
                if (ei >= nn) {
                    ei -= nn;
                    ++mi;
                }
            }
        }
        __syncthreads();
        int kstart = -1;
        if (tid < cnt) {
            double* l = mats + tid * S;
            double* g = gvec + tid * V;
            double* e = evec + tid * V;
            unsigned long long* p = pvec + tid * V;
            if constexpr (kAlgo == kGMW81)
                tpm_gmw(n, l, cmat + tid * S, g, e, p);
            else
                kstart = tpm_se<kAlgo>(n, l, g, e, p);
            // e in original order (gmw81.rs:127-131, se90.rs:174-178, se99.rs:194-198): through g
            for (int i = 0; i < n; ++i)
                g[(int)p[i]] = e[i];
        }
        __syncthreads();
        // ---- out -------------------------------------------------------------------------------
        if (L) {
            double* dst = L + (size_t)t0 * nn;
            int mi = tid / nn, ei = tid % nn;
            for (int q = tid; q < words; q += kTpmThreads) {
                dst[q] = mats[mi * S + ei];
                mi += step_m;
                ei += step_e;
                if (ei >= nn) {
                    ei -= nn;
                    ++mi;
                }
            }
        }
        {
            int mi = tid / n, ei = tid % n;
            const int vwords = cnt * n;
            for (int q = tid; q < vwords; q += kTpmThreads) {
                if (E)
                    E[(size_t)t0 * n + q] = gvec[mi * V + ei];
                if (P)
                    P[(size_t)t0 * n + q] = pvec[mi * V + ei];
                mi += step_mv;
                ei += step_ev;
                if (ei >= n) {
                    ei -= n;
                    ++mi;
                }
            }
        }
        if (K && tid < cnt)
            K[t0 + tid] = kstart;
        __syncthreads();   // the tile's shared memory is free again
    }
}


// ================================================================================================
// The same thread-per-matrix scheme with COMPILE-TIME n (tpr_kernel<kAlgo, N>, N <= kTprMaxN):
// every loop unrolled, the lower triangle, g / d, e and p in REGISTERS.  The run-time-n kernel
// above spends four of five instructions on index arithmetic, loop control and shared-memory
// traffic (profiles/r02_tpm_se99_n4_summary.txt: 106 warp-instructions per 4x4 matrix, 12 of them
// FP64); here the only run-time index is the pivot row m, and the exchange j <-> m becomes a chain
// of predicated register swaps over the candidates r = j+1 .. N-1.
// Storage is the LOWER triangle only: the reference keeps both triangles of its working matrix
// bit-identical (mirror stores se90.rs:91, se99.rs:103,169; the two divisions by l_jj act on equal
// values), so for the symmetric inputs the crate is specified for, every value it reads from the
// upper triangle is the lower one's -- the arithmetic below is still the reference's, element for
// element, and the results are bit-identical (tests/test_parity_gpu.py).
// ================================================================================================
constexpr int kTprMaxN = 8;

__device__ __forceinline__ void tpr_cswap(bool c, double& a, double& b)
{
    const double t = a;
    a = c ? b : a;
    b = c ? t : b;
}
__device__ __forceinline__ void tpr_cswap(bool c, int& a, int& b)
{
    const int t = a;
    a = c ? b : a;
    b = c ? t : b;
}

template <int N>
struct TprTri {
    double v[N * (N + 1) / 2];
    __device__ __forceinline__ double& operator()(int i, int k) { return v[i * (i + 1) / 2 + k]; }   // i >= k
};

// symmetric exchange of rows / columns J and m (J < m): utils.rs:41-49 + :30-38 on lower storage,
// as predicated swaps over the candidates r
template <int N, int J>
__device__ __forceinline__ void tpr_sym_cswap(TprTri<N>& a, int m)
{
#pragma unroll
    for (int r = J + 1; r < N; ++r) {
        const bool c = m == r;
        tpr_cswap(c, a(J, J), a(r, r));
#pragma unroll
        for (int k = 0; k < J; ++k)
            tpr_cswap(c, a(J, k), a(r, k));
#pragma unroll
        for (int k = J + 1; k < r; ++k)
            tpr_cswap(c, a(k, J), a(r, k));
#pragma unroll
        for (int k = r + 1; k < N; ++k)
            tpr_cswap(c, a(k, J), a(k, r));
    }
}
template <int N, int J, typename T>
__device__ __forceinline__ void tpr_vec_cswap(T (&v)[N], int m)
{
#pragma unroll
    for (int r = J + 1; r < N; ++r)
        tpr_cswap(m == r, v[J], v[r]);
}

// se90.rs:84-93 == :142-151 == se99.rs:96-105 == :162-171 on the lower triangle
template <int N, int J>
__device__ __forceinline__ void tpr_eliminate(TprTri<N>& l)
{
    const double ljj = dsqrt(l(J, J));
    l(J, J) = ljj;
#pragma unroll
    for (int i = J + 1; i < N; ++i) {
        const double lij = ddiv(l(i, J), ljj);
        l(i, J) = lij;
#pragma unroll
        for (int k = J + 1; k <= i; ++k)
            l(i, k) = dsub(l(i, k), dmul(lij, l(k, J)));
    }
}

template <int N>
struct TprSe {
    TprTri<N> l;
    double g[N], e[N];
    int p[N];
    double gamma, taugamma, delta_prev;
    bool ph1;
    int kstart;
};

// column J of SE90 / SE99: the phase-1 step (or the switch to phase 2), then the phase-2 step
template <int kAlgo, int N, int J>
__device__ __forceinline__ void tpr_se_step(TprSe<N>& s)
{
    TprTri<N>& l = s.l;
    bool elim = false;   // ONE elimination site per column: in a warp whose matrices are in different
                         // phases the two pivot paths run one after the other, the elimination once
    if (s.ph1) {   // phase 1: se90.rs:61-96, se99.rs:61-109
        if (kAlgo == kSE99) {   // se99.rs:62-73
            double dmax = -INFINITY, dmin = INFINITY;
#pragma unroll
            for (int i = J; i < N; ++i) {
                const double x = l(i, i);
                if (x > dmax) dmax = x;
                if (x < dmin) dmin = x;
            }
            if (dmax < s.taugamma || dmin < dmul(-MODCHOL_MU, dmax))
                s.ph1 = false;
        }
        if (s.ph1) {
            int m = J;   // utils.rs:52-70: first strict maximum, a NaN head sticks
            double best = l(J, J);
#pragma unroll
            for (int i = J + 1; i < N; ++i) {
                const double c = l(i, i);
                if (c > best) {
                    best = c;
                    m = i;
                }
            }
            tpr_sym_cswap<N, J>(l, m);   // se90.rs:63-68, se99.rs:76-81
            tpr_vec_cswap<N, J>(s.p, m);
            double la = INFINITY;   // se90.rs:70-77, se99.rs:82-89
            const double ljj = l(J, J);
#pragma unroll
            for (int i = J + 1; i < N; ++i) {
                const double lij = l(i, J);
                const double nv = dsub(l(i, i), ddiv(dmul(lij, lij), ljj));
                if (nv < la)
                    la = nv;
            }
            const double thresh = kAlgo == kSE99 ? dmul(-MODCHOL_MU, s.gamma) : s.taugamma;   // se99.rs:90 / se90.rs:78
            if (la < thresh)
                s.ph1 = false;
        }
        if (s.ph1) {
            elim = true;
        } else {
            s.kstart = J;   // phase 2 starts at this column
            if (kAlgo == kSE99 && J == N - 1) {   // se99.rs:115-119
                const double ljj = l(J, J);
                const double ej = tail1x1_e(ljj, s.gamma);
                s.e[J] = ej;
                l(J, J) = dsqrt(dadd(ljj, ej));
            } else if (kAlgo == kSE90 || J + 1 < N) {
                // lower Gershgorin bounds of the trailing block, se90.rs:105-110, se99.rs:125-130
#pragma unroll
                for (int i = J; i < N; ++i) {
                    double s1 = 0.0, s2 = 0.0;
#pragma unroll
                    for (int q = J; q < i; ++q)
                        s1 = dadd(s1, fabs(l(i, q)));
#pragma unroll
                    for (int r = i + 1; r < N; ++r)
                        s2 = dadd(s2, fabs(l(r, i)));
                    s.g[i] = dsub(dsub(l(i, i), s1), s2);
                }
            }
        }
    }
    if constexpr (J + 2 < N) {
        if (!s.ph1) {   // phase-2 step, se90.rs:113-152, se99.rs:133-172
            int m = J;
            double best = s.g[J];
#pragma unroll
            for (int i = J + 1; i < N; ++i) {
                const double c = s.g[i];
                if (c > best) {
                    best = c;
                    m = i;
                }
            }
            tpr_sym_cswap<N, J>(l, m);
            tpr_vec_cswap<N, J>(s.p, m);
            tpr_vec_cswap<N, J>(s.g, m);
            double normj = 0.0;   // se90.rs:124-131, se99.rs:144-151
#pragma unroll
            for (int r = J + 1; r < N; ++r)
                normj = dadd(normj, fabs(l(r, J)));
            double ljj = l(J, J);
            const double ej = phase2_e(ljj, normj, s.taugamma, s.delta_prev);
            s.e[J] = ej;
            if (ej > 0.0) {
                ljj = dadd(ljj, ej);
                l(J, J) = ljj;
                s.delta_prev = ej;
            }
            if (fabs(dsub(ljj, normj)) > MODCHOL_EPS) {   // se90.rs:134-139, se99.rs:154-159
                const double t = dsub(1.0, ddiv(normj, ljj));
#pragma unroll
                for (int i = J + 1; i < N; ++i)
                    s.g[i] = dadd(s.g[i], dmul(fabs(l(i, J)), t));
            }
            elim = true;
        }
    }
    if (elim)
        tpr_eliminate<N, J>(l);
}
template <int kAlgo, int N, int J>
__device__ __forceinline__ void tpr_se_steps(TprSe<N>& s)
{
    if constexpr (J < N) {
        tpr_se_step<kAlgo, N, J>(s);
        tpr_se_steps<kAlgo, N, J + 1>(s);
    }
}

// SE90 / SE99 on the lower triangle in s.l; leaves L there, e by position, p; returns kstart
template <int kAlgo, int N>
__device__ __forceinline__ void tpr_se(TprSe<N>& s)
{
    TprTri<N>& l = s.l;
    s.gamma = 0.0;   // se90.rs:55-57, se99.rs:54-56
#pragma unroll
    for (int i = 0; i < N; ++i) {
        s.e[i] = 0.0;
        s.p[i] = i;
        s.g[i] = 0.0;
        const double d = fabs(l(i, i));
        if (d > s.gamma)
            s.gamma = d;
    }
    s.taugamma = dmul(MODCHOL_TAU, s.gamma);
    s.ph1 = true;
    s.kstart = -1;
    s.delta_prev = 0.0;
    tpr_se_steps<kAlgo, N, 0>(s);
    if constexpr (N >= 2) {
        if (!s.ph1 && (kAlgo == kSE90 || s.kstart + 1 < N)) {   // final 2x2: se90.rs:155-166, se99.rs:175-186
            double lhi, llo;
            const double off = l(N - 1, N - 2);   // m[0,1] == m[1,0] (symmetric working matrix)
            eigenvalues_2x2(l(N - 2, N - 2), off, off, l(N - 1, N - 1), lhi, llo);
            const double en = tail2x2_e(kAlgo, lhi, llo, s.gamma, s.delta_prev);
            s.e[N - 2] = en;
            s.e[N - 1] = en;
            if (en > 0.0) {
                l(N - 2, N - 2) = dadd(l(N - 2, N - 2), en);
                l(N - 1, N - 1) = dadd(l(N - 1, N - 1), en);
            }
            const double l00 = dsqrt(l(N - 2, N - 2));
            l(N - 2, N - 2) = l00;
            const double l10 = ddiv(off, l00);
            l(N - 1, N - 2) = l10;
            l(N - 1, N - 1) = dsqrt(dsub(l(N - 1, N - 1), dmul(l10, l10)));
        }
    }
}

template <int N>
struct TprGmw {
    TprTri<N> l, c;
    double d[N], e[N];
    int p[N];
    double beta, delta;
};

// column J of GMW81 (gmw81.rs:70-111)
template <int N, int J>
__device__ __forceinline__ void tpr_gmw_step(TprGmw<N>& s)
{
    TprTri<N>&l = s.l, &c = s.c;
    int m = J;   // utils.rs:73-91 on the diagonal of c (gmw81.rs:71-78)
    double best = fabs(c(J, J));
#pragma unroll
    for (int i = J + 1; i < N; ++i) {
        const double v = fabs(c(i, i));
        if (v > best) {
            best = v;
            m = i;
        }
    }
    tpr_sym_cswap<N, J>(c, m);
    tpr_sym_cswap<N, J>(l, m);
    tpr_vec_cswap<N, J>(s.p, m);
#pragma unroll
    for (int q = 0; q < J; ++q)   // gmw81.rs:79-81
        l(J, q) = ddiv(c(J, q), s.d[q]);
#pragma unroll
    for (int i = J; i < N; ++i) {   // gmw81.rs:83-85
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < J; ++q)
            acc = dadd(acc, dmul(l(J, q), c(i, q)));
        c(i, J) = dsub(l(i, J), acc);
    }
    double theta = 0.0;   // gmw81.rs:87-100
#pragma unroll
    for (int i = J + 1; i < N; ++i) {
        const double v = fabs(c(i, J));
        if (v > theta)
            theta = v;
    }
    const double tb = ddiv(theta, s.beta);
    const double cjj = c(J, J);
    const double dj = fmax(fmax(fabs(cjj), dmul(tb, tb)), s.delta);   // gmw81.rs:102
    s.d[J] = dj;
#pragma unroll
    for (int i = J + 1; i < N; ++i) {   // gmw81.rs:104-109
        const double cij = c(i, J);
        c(i, i) = dsub(c(i, i), ddiv(dmul(cij, cij), dj));
    }
    s.e[J] = dsub(dj, cjj);   // gmw81.rs:110
}
template <int N, int J>
__device__ __forceinline__ void tpr_gmw_steps(TprGmw<N>& s)
{
    if constexpr (J < N) {
        tpr_gmw_step<N, J>(s);
        tpr_gmw_steps<N, J + 1>(s);
    }
}

// GMW81 on the lower triangle in s.l (gmw81.rs:39-134); leaves L there
template <int N>
__device__ __forceinline__ void tpr_gmw(TprGmw<N>& s)
{
    TprTri<N>&l = s.l, &c = s.c;
    double diag_max = 0.0, off_max = 0.0;   // gmw81.rs:49-58 (the upper triangle mirrors the lower)
#pragma unroll
    for (int i = 0; i < N; ++i) {
        s.e[i] = 0.0;
        s.d[i] = 0.0;
        s.p[i] = i;
#pragma unroll
        for (int q = 0; q <= i; ++q) {
            const double v = fabs(l(i, q));
            if (i == q) {
                if (v > diag_max) diag_max = v;
            } else if (v > off_max) {
                off_max = v;
            }
            c(i, q) = i == q ? l(i, q) : 0.0;   // gmw81.rs:66-68
        }
    }
    const double nf = (double)N;
    s.delta = dmul(MODCHOL_EPS, fmax(1.0, dadd(diag_max, off_max)));   // gmw81.rs:60
    s.beta = dsqrt(fmax(fmax(diag_max, ddiv(off_max, dsqrt(dsub(dmul(nf, nf), 1.0)))), MODCHOL_EPS));   // :61-64
    tpr_gmw_steps<N, 0>(s);
    // gmw81.rs:112-125: l . diag(sqrt(d)) with a unit diagonal; the dense product's accumulation
    // from 0.0 turns a -0 into +0
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int q = 0; q <= i; ++q)
            l(i, q) = dadd(0.0, q == i ? dsqrt(s.d[q]) : dmul(l(i, q), dsqrt(s.d[q])));
}

// The consumer step fused behind the factorisation: (A + E) x = rhs for nrhs right-hand sides of
// this thread's matrix, L L^T (P x) = P rhs with L in registers (position order) and p.  x may alias rhs
// (every entry of a right-hand side is read before the first is written).
template <int N>
__device__ __forceinline__ void tpr_solve(TprTri<N>& l, const int (&p)[N], const double* rhs, double* x, int nrhs)
{
    double rd[N];
#pragma unroll
    for (int i = 0; i < N; ++i)
        rd[i] = ddiv(1.0, l(i, i));
    for (int r = 0; r < nrhs; ++r) {
        double z[N];
#pragma unroll
        for (int i = 0; i < N; ++i)
            z[i] = rhs[(size_t)r * N + p[i]];
#pragma unroll
        for (int j = 0; j < N; ++j) {   // L z = P rhs
            z[j] = dmul(z[j], rd[j]);
#pragma unroll
            for (int i = j + 1; i < N; ++i)
                z[i] = __fma_rn(-l(i, j), z[j], z[i]);
        }
#pragma unroll
        for (int j = N - 1; j >= 0; --j) {   // L^T w = z
            z[j] = dmul(z[j], rd[j]);
#pragma unroll
            for (int i = 0; i < j; ++i)
                z[i] = __fma_rn(-l(j, i), z[j], z[i]);
        }
#pragma unroll
        for (int i = 0; i < N; ++i)
            x[(size_t)r * N + p[i]] = z[i];
    }
}

// ---- direct global access of one thread to ITS matrix ------------------------------------------
// The threads of a warp hold consecutive matrices, so one warp-wide access touches 32 different
// sectors; every byte of a sector is consumed by the same thread within the next instructions
// (L1 serves the rest of the sector), so DRAM sees each needed sector once -- and nothing has to be
// staged, indexed or synchronised.  Only the lower triangle is read.  128-bit loads / 256-bit
// stores where the row length and the base alignment allow.
__device__ __forceinline__ void tp_stg_v4(double* p, double a, double b, double c, double d)
{
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" :: "l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void tp_ldg_v4(const double* p, double& a, double& b, double& c, double& d)
{
    asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
// put(i, k, value) for every k <= i
template <int N, typename F>
__device__ __forceinline__ void tp_load_lower(const double* __restrict__ a, bool al16, F&& put, bool al32 = false)
{
    if (N % 4 == 0 && al32) {   // one whole sector per thread and instruction
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int k = 0; k <= i; k += 4) {
                double v0, v1, v2, v3;
                tp_ldg_v4(a + i * N + k, v0, v1, v2, v3);
                put(i, k, v0);
                if (k + 1 <= i) put(i, k + 1, v1);
                if (k + 2 <= i) put(i, k + 2, v2);
                if (k + 3 <= i) put(i, k + 3, v3);
            }
    } else if (N % 2 == 0 && al16) {
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int k = 0; k <= i; k += 2) {
                const double2 v = __ldg(reinterpret_cast<const double2*>(a + i * N + k));
                put(i, k, v.x);
                if (k + 1 <= i)
                    put(i, k + 1, v.y);
            }
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int k = 0; k <= i; ++k)
                put(i, k, __ldg(a + i * N + k));
    }
}
// the full row-major L: get(i, k) for k <= i, zero above the diagonal
template <int N, typename F>
__device__ __forceinline__ void tp_store_full(double* __restrict__ dst, bool al32, bool al16, F&& get)
{
    auto val = [&](int i, int k) { return k <= i ? get(i, k) : 0.0; };
    if (N % 4 == 0 && al32) {
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int k = 0; k < N; k += 4)
                tp_stg_v4(dst + i * N + k, val(i, k), val(i, k + 1), val(i, k + 2), val(i, k + 3));
    } else if (N % 2 == 0 && al16) {
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int k = 0; k < N; k += 2)
                *reinterpret_cast<double2*>(dst + i * N + k) = make_double2(val(i, k), val(i, k + 1));
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int k = 0; k < N; ++k)
                dst[i * N + k] = val(i, k);
    }
}

// e (un-permuted: gmw81.rs:127-131, se90.rs:174-178, se99.rs:194-198) and p of one matrix.  With
// N a multiple of 4 and aligned bases both vectors leave as 256-bit stores -- e is brought into
// original order in registers first (a select chain over p) -- instead of 2 N scattered 8-byte
// stores, 32 sectors per warp instruction each.
template <int N>
__device__ __forceinline__ void tp_store_e_p(double* E, unsigned long long* P, long long b,
                                             const double (&e)[N], const int (&p)[N])
{
    const bool vec = N % 4 == 0 && ((((size_t)E) | ((size_t)P)) & 31) == 0;
    if (vec) {
        if (E) {
            double eo[N];
#pragma unroll
            for (int k = 0; k < N; ++k) {
                eo[k] = 0.0;
#pragma unroll
                for (int i = 0; i < N; ++i)
                    eo[k] = p[i] == k ? e[i] : eo[k];
            }
#pragma unroll
            for (int k = 0; k < N; k += 4)
                tp_stg_v4(E + (size_t)b * N + k, eo[k], eo[k + 1], eo[k + 2], eo[k + 3]);
        }
        if (P) {
#pragma unroll
            for (int k = 0; k < N; k += 4)
                tp_stg_v4(reinterpret_cast<double*>(P + (size_t)b * N + k), __longlong_as_double((long long)p[k]),
                          __longlong_as_double((long long)p[k + 1]), __longlong_as_double((long long)p[k + 2]),
                          __longlong_as_double((long long)p[k + 3]));
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (E) E[(size_t)b * N + p[i]] = e[i];
            if (P) P[(size_t)b * N + i] = (unsigned long long)p[i];
        }
    }
}

template <int kAlgo, int N, bool kSolve = false>
__global__ void __launch_bounds__(kTpmThreads)
tpr_kernel(long long batch, const double* __restrict__ A, double* __restrict__ L,
           double* __restrict__ E, unsigned long long* __restrict__ P, int* __restrict__ K,
           const double* Bm = nullptr, double* X = nullptr, int nrhs = 0)   // X may alias Bm
{
    constexpr int NN = N * N;
    const bool al16 = (((size_t)A | (size_t)L) & 15) == 0, al32 = (((size_t)L) & 31) == 0;
    const bool al32a = (((size_t)A) & 31) == 0;
    const long long stride = (long long)gridDim.x * kTpmThreads;
    for (long long b = (long long)blockIdx.x * kTpmThreads + threadIdx.x; b < batch; b += stride) {
        const double* a = A + (size_t)b * NN;
        int kstart = -1;
        if constexpr (kAlgo == kGMW81) {
            TprGmw<N> s;
            tp_load_lower<N>(a, al16, [&](int i, int k, double v) { s.l(i, k) = v; }, al32a);
            tpr_gmw<N>(s);
            if (L)   // gmw81.rs:119-123
                tp_store_full<N>(L + (size_t)b * NN, al32, al16, [&](int i, int k) { return s.l(i, k); });
            tp_store_e_p<N>(E, P, b, s.e, s.p);
            if constexpr (kSolve)
                tpr_solve<N>(s.l, s.p, Bm + (size_t)b * nrhs * N, X + (size_t)b * nrhs * N, nrhs);
        } else {
            TprSe<N> s;
            tp_load_lower<N>(a, al16, [&](int i, int k, double v) { s.l(i, k) = v; }, al32a);
            tpr_se<kAlgo, N>(s);
            kstart = s.kstart;
            if (L)   // se90.rs:169-173, se99.rs:189-192
                tp_store_full<N>(L + (size_t)b * NN, al32, al16, [&](int i, int k) { return s.l(i, k); });
            tp_store_e_p<N>(E, P, b, s.e, s.p);
            if constexpr (kSolve)
                tpr_solve<N>(s.l, s.p, Bm + (size_t)b * nrhs * N, X + (size_t)b * nrhs * N, nrhs);
        }
        if (K)
            K[b] = kstart;
    }
}

// ---- the same kernel with the tile STAGED through shared memory: fully coalesced copies of a tile of
// kTpmThreads matrices (word q of the tile at q + q / (N N) for even N: odd stride between the
// threads' matrices), a barrier either side of the arithmetic.  Ahead of the direct accesses
// wherever these are narrower than a sector's worth per thread and instruction (odd n, n = 4, 6).
// doubles of shared memory per thread: the matrix (stride N N | 1), then e and p (stride N | 1)
__host__ __device__ constexpr int tpr_thread_doubles(int n) { return tpm_mat_stride(n) + 2 * tpm_vec_stride(n); }

template <int kAlgo, int N, bool kSolve = false>
__global__ void __launch_bounds__(kTpmThreads)
tpr_kernel_staged(long long batch, const double* __restrict__ A, double* __restrict__ L,
                  double* __restrict__ E, unsigned long long* __restrict__ P, int* __restrict__ K,
                  const double* Bm = nullptr, double* X = nullptr, int nrhs = 0)   // X may alias Bm
{
    extern __shared__ double sm[];
    constexpr int NN = N * N, S = tpm_mat_stride(N), V = tpm_vec_stride(N);
    const int tid = threadIdx.x;
    double* mats = sm;                       // kTpmThreads x S
    double* evec = mats + kTpmThreads * S;   // e, original order
    unsigned long long* pvec = reinterpret_cast<unsigned long long*>(evec + kTpmThreads * V);

    for (long long t0 = (long long)blockIdx.x * kTpmThreads; t0 < batch; t0 += (long long)gridDim.x * kTpmThreads) {
        const int cnt = (int)min((long long)kTpmThreads, batch - t0);
        const int words = cnt * NN;
        {   // ---- in: coalesced over the tile, matrix by matrix into shared memory ---------------
            const double* src = A + (size_t)t0 * NN;
            for (int q = tid; q < words; q += kTpmThreads)
                mats[q + (S - NN) * (q / NN)] = __ldg(src + q);
        }
        __syncthreads();
        int kstart = -1;
        if (tid < cnt) {
            double* mine = mats + tid * S;
            double* eo = evec + tid * V;
            unsigned long long* po = pvec + tid * V;
            if constexpr (kAlgo == kGMW81) {
                TprGmw<N> s;
#pragma unroll
                for (int i = 0; i < N; ++i)
#pragma unroll
                    for (int q = 0; q <= i; ++q)
                        s.l(i, q) = mine[i * N + q];
                tpr_gmw<N>(s);
#pragma unroll
                for (int i = 0; i < N; ++i) {
#pragma unroll
                    for (int q = 0; q < N; ++q)
                        mine[i * N + q] = q <= i ? s.l(i, q) : 0.0;   // gmw81.rs:119-123
                    eo[s.p[i]] = s.e[i];                                // gmw81.rs:127-131
                    po[i] = (unsigned long long)s.p[i];
                }
                if constexpr (kSolve)
                    tpr_solve<N>(s.l, s.p, Bm + (size_t)(t0 + tid) * nrhs * N, X + (size_t)(t0 + tid) * nrhs * N, nrhs);
            } else {
                TprSe<N> s;
#pragma unroll
                for (int i = 0; i < N; ++i)
#pragma unroll
                    for (int q = 0; q <= i; ++q)
                        s.l(i, q) = mine[i * N + q];
                tpr_se<kAlgo, N>(s);
                kstart = s.kstart;
#pragma unroll
                for (int i = 0; i < N; ++i) {
#pragma unroll
                    for (int q = 0; q < N; ++q)
                        mine[i * N + q] = q <= i ? s.l(i, q) : 0.0;   // se90.rs:169-173, se99.rs:189-192
                    eo[s.p[i]] = s.e[i];                                // se90.rs:174-178, se99.rs:194-198
                    po[i] = (unsigned long long)s.p[i];
                }
                if constexpr (kSolve)
                    tpr_solve<N>(s.l, s.p, Bm + (size_t)(t0 + tid) * nrhs * N, X + (size_t)(t0 + tid) * nrhs * N, nrhs);
            }
        }
        __syncthreads();
        // ---- out -------------------------------------------------------------------------------
        if (L) {
            double* dst = L + (size_t)t0 * NN;
            for (int q = tid; q < words; q += kTpmThreads)
                dst[q] = mats[q + (S - NN) * (q / NN)];
        }
        {
            const int vwords = cnt * N;
            for (int q = tid; q < vwords; q += kTpmThreads) {
                const int at = q + (V - N) * (q / N);
                if (E)
                    E[(size_t)t0 * N + q] = evec[at];
                if (P)
                    P[(size_t)t0 * N + q] = pvec[at];
            }
        }
        if (K && tid < cnt)
            K[t0 + tid] = kstart;
        __syncthreads();   // the tile's shared memory is free again
    }
}


// ================================================================================================
// 8 < n <= kTpsMaxN: still one thread per matrix, but the lower triangle no longer fits the
// register file -- it lives in a thread-private range of SHARED memory (packed, T(n) doubles, odd
// stride between threads: conflict free whenever the threads of a warp touch the same element),
// together with g, e and p.  n is a compile-time constant; the O(n^2)-per-step trailing update is
// instantiated per column (tps_update<N, J>), everything that is O(n) per step (pivot search,
// exchange, column scaling, norms, bound updates) runs as ordinary loops over run-time indices, so
// the code stays inside the instruction cache.  One warp per block (32 matrices per tile), no
// block-wide barrier.
// Sums of eight or more terms follow ndarray's unrolled_fold (eight partials, see common.cuh).
// GMW81 shares one triangle between the reference's l and c (tps_gmw below).
// ================================================================================================
constexpr int kTpsMaxN = 16;
constexpr int kTpsThreads = 32;
// per thread: the triangle, then SE90 / SE99: one vector shared by g (positions not yet eliminated)
// and e (positions already eliminated); GMW81: lrow, d and one vector shared by the running
// diagonal and e.  The permutation is sixteen 4-bit fields of one register.
__host__ __device__ constexpr int tps_thread_doubles(int n, int algo = kSE99) { return (n * (n + 1) / 2 + (algo == kGMW81 ? 3 : 1) * n) | 1; }

struct TpsPerm {
    unsigned long long v = 0xFEDCBA9876543210ull;   // p[i] = i
    __device__ __forceinline__ int get(int i) const { return (int)((v >> (4 * i)) & 15ull); }
    __device__ __forceinline__ void swap(int j, int m)
    {
        const unsigned long long x = (unsigned long long)(get(j) ^ get(m));
        v ^= (x << (4 * j)) | (x << (4 * m));
    }
};

// ndarray `Array1::sum()` over len terms x(0) .. x(len-1)
template <typename F>
__device__ __forceinline__ double tps_nd_sum(int len, F x)
{
    double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0, p4 = 0.0, p5 = 0.0, p6 = 0.0, p7 = 0.0;
    int i = 0;
    for (; i + 8 <= len; i += 8) {
        p0 = dadd(p0, x(i));
        p1 = dadd(p1, x(i + 1));
        p2 = dadd(p2, x(i + 2));
        p3 = dadd(p3, x(i + 3));
        p4 = dadd(p4, x(i + 4));
        p5 = dadd(p5, x(i + 5));
        p6 = dadd(p6, x(i + 6));
        p7 = dadd(p7, x(i + 7));
    }
    double acc = 0.0;
    acc = dadd(acc, dadd(p0, p4));
    acc = dadd(acc, dadd(p1, p5));
    acc = dadd(acc, dadd(p2, p6));
    acc = dadd(acc, dadd(p3, p7));
    for (; i < len; ++i)
        acc = dadd(acc, x(i));
    return acc;
}

__device__ __forceinline__ int tps_tri(int i) { return (i * (i + 1)) >> 1; }

// symmetric exchange j <-> m (j < m) on the packed lower triangle: utils.rs:41-49 + :30-38
__device__ __forceinline__ void tps_swap(double* l, int n, int j, int m)
{
    const int tj = tps_tri(j), tm = tps_tri(m);
    auto sw = [&](int a, int b) {
        const double t = l[a];
        l[a] = l[b];
        l[b] = t;
    };
    sw(tj + j, tm + m);
    for (int k = 0; k < j; ++k)
        sw(tj + k, tm + k);
    for (int k = j + 1; k < m; ++k)
        sw(tps_tri(k) + j, tm + k);
    for (int k = m + 1; k < n; ++k) {
        const int tk = tps_tri(k);
        sw(tk + j, tk + m);
    }
}

// se90.rs:84-93 == :142-151 == se99.rs:96-105 == :162-171 in two parts.  The column scaling
// (one square root, n-j-1 divisions -- each a twenty-instruction sequence) is an ordinary loop
// over the run-time column j: one copy in the instruction stream ...
__device__ __forceinline__ void tps_scale_column(double* l, int n, int j)
{
    const int tj = tps_tri(j);
    const double ljj = dsqrt(l[tj + j]);
    l[tj + j] = ljj;
#pragma unroll 4
    for (int i = j + 1; i < n; ++i) {
        const int a = tps_tri(i) + j;
        l[a] = ddiv(l[a], ljj);
    }
}
// ... and the O(n^2) trailing update is instantiated per column J: the scaled column in registers,
// every trailing element one LDS + DMUL + DSUB + STS at a constant offset
template <int N, int J>
__device__ __forceinline__ void tps_update(double* l)
{
    double col[N];
#pragma unroll
    for (int i = J + 1; i < N; ++i)
        col[i] = l[i * (i + 1) / 2 + J];
#pragma unroll
    for (int i = J + 1; i < N; ++i)
#pragma unroll
        for (int k = J + 1; k <= i; ++k)
            l[i * (i + 1) / 2 + k] = dsub(l[i * (i + 1) / 2 + k], dmul(col[i], col[k]));
}
template <int N, int J>
__device__ __forceinline__ void tps_eliminate_at(double* l, int j)
{
    if constexpr (J < N) {
        if (j == J)
            tps_update<N, J>(l);
        else
            tps_eliminate_at<N, J + 1>(l, j);
    }
}

// SE90 / SE99 on the packed lower triangle at l; g, e, p: N doubles each.  Returns kstart.
template <int kAlgo, int N>
__device__ __forceinline__ int tps_se(double* l, double* g, TpsPerm& p)
{
    double gamma = 0.0;   // se90.rs:55-57, se99.rs:54-56
    for (int i = 0; i < N; ++i) {
        g[i] = 0.0;   // e of a column that phase 1 eliminates
        const double d = fabs(l[tps_tri(i) + i]);
        if (d > gamma)
            gamma = d;
    }
    const double taugamma = dmul(MODCHOL_TAU, gamma);
    bool ph1 = true;
    int kstart = -1;
    double delta_prev = 0.0;
    for (int j = 0; j < N; ++j) {
        const int tj = tps_tri(j);
        bool elim = false;
        if (ph1) {   // phase 1: se90.rs:61-96, se99.rs:61-109
            if (kAlgo == kSE99) {   // se99.rs:62-73
                double dmax = -INFINITY, dmin = INFINITY;
                for (int i = j; i < N; ++i) {
                    const double x = l[tps_tri(i) + i];
                    if (x > dmax) dmax = x;
                    if (x < dmin) dmin = x;
                }
                if (dmax < taugamma || dmin < dmul(-MODCHOL_MU, dmax))
                    ph1 = false;
            }
            if (ph1) {
                int m = j;   // utils.rs:52-70
                double best = l[tj + j];
                for (int i = j + 1; i < N; ++i) {
                    const double c = l[tps_tri(i) + i];
                    if (c > best) {
                        best = c;
                        m = i;
                    }
                }
                if (m != j) {   // se90.rs:63-68, se99.rs:76-81
                    tps_swap(l, N, j, m);
                    p.swap(j, m);
                }
                double la = INFINITY;   // se90.rs:70-77, se99.rs:82-89
                const double ljj = l[tj + j];
                for (int i = j + 1; i < N; ++i) {
                    const int ti = tps_tri(i);
                    const double lij = l[ti + j];
                    const double nv = dsub(l[ti + i], ddiv(dmul(lij, lij), ljj));
                    if (nv < la)
                        la = nv;
                }
                const double thresh = kAlgo == kSE99 ? dmul(-MODCHOL_MU, gamma) : taugamma;   // se99.rs:90 / se90.rs:78
                if (la < thresh)
                    ph1 = false;
                else
                    elim = true;
            }
            if (!ph1) {
                kstart = j;   // phase 2 starts at this column
                if (kAlgo == kSE99 && j == N - 1) {   // se99.rs:115-119
                    const double ljj = l[tj + j];
                    const double ej = tail1x1_e(ljj, gamma);
                    g[j] = ej;
                    l[tj + j] = dsqrt(dadd(ljj, ej));
                } else if (kAlgo == kSE90 || j + 1 < N) {
                    // lower Gershgorin bounds of the trailing block, se90.rs:105-110, se99.rs:125-130
                    for (int i = j; i < N; ++i) {
                        const int ti = tps_tri(i);
                        const double s1 = tps_nd_sum(i - j, [&](int q) { return fabs(l[ti + j + q]); });
                        const double s2 = tps_nd_sum(N - 1 - i, [&](int r) { return fabs(l[tps_tri(i + 1 + r) + i]); });
                        g[i] = dsub(dsub(l[ti + i], s1), s2);
                    }
                }
            }
        }
        if (!ph1 && j + 2 < N) {   // phase-2 step, se90.rs:113-152, se99.rs:133-172
            int m = j;
            double best = g[j];
            for (int i = j + 1; i < N; ++i) {
                const double c = g[i];
                if (c > best) {
                    best = c;
                    m = i;
                }
            }
            if (m != j) {
                tps_swap(l, N, j, m);
                p.swap(j, m);
                const double t = g[j];
                g[j] = g[m];
                g[m] = t;
            }
            // se90.rs:124-131, se99.rs:144-151
            const double normj = tps_nd_sum(N - 1 - j, [&](int r) { return fabs(l[tps_tri(j + 1 + r) + j]); });
            double ljj = l[tj + j];
            const double ej = phase2_e(ljj, normj, taugamma, delta_prev);
            g[j] = ej;   // the bound of position j is dead: its slot takes e_j
            if (ej > 0.0) {
                ljj = dadd(ljj, ej);
                l[tj + j] = ljj;
                delta_prev = ej;
            }
            if (fabs(dsub(ljj, normj)) > MODCHOL_EPS) {   // se90.rs:134-139, se99.rs:154-159
                const double t = dsub(1.0, ddiv(normj, ljj));
                for (int i = j + 1; i < N; ++i)
                    g[i] = dadd(g[i], dmul(fabs(l[tps_tri(i) + j]), t));
            }
            elim = true;
        }
        if (elim) {
            tps_scale_column(l, N, j);
            tps_eliminate_at<N, 0>(l, j);
        }
    }
    if (!ph1 && (kAlgo == kSE90 || kstart + 1 < N)) {   // final 2x2: se90.rs:155-166, se99.rs:175-186
        constexpr int A00 = (N - 2) * (N - 1) / 2 + (N - 2), A10 = (N - 1) * N / 2 + (N - 2), A11 = A10 + 1;
        double lhi, llo;
        const double off = l[A10];   // m[0,1] == m[1,0] (symmetric working matrix)
        eigenvalues_2x2(l[A00], off, off, l[A11], lhi, llo);
        const double en = tail2x2_e(kAlgo, lhi, llo, gamma, delta_prev);
        g[N - 2] = en;
        g[N - 1] = en;
        double a00 = l[A00], a11 = l[A11];
        if (en > 0.0) {
            a00 = dadd(a00, en);
            a11 = dadd(a11, en);
        }
        const double l00 = dsqrt(a00);
        l[A00] = l00;
        const double l10 = ddiv(off, l00);
        l[A10] = l10;
        l[A11] = dsqrt(dsub(a11, dmul(l10, l10)));
    }
    return kstart;
}

// ---- GMW81 for 8 < n <= kTpsMaxN (gmw81.rs:39-134) on ONE triangle ------------------------------
// The reference keeps two matrices, l (a clone of A that turns into the unit-lower factor) and c.
// Per thread they share one packed triangle W: W(i, s), s < i, holds the (permuted) a_is until
// column s is formed at step s, c_is from then on, and l_is = c_is / d_s once row i has been the
// pivot row (step i: the quotients go through the buffer lrow while the dot products of the step
// still need c_js).  The diagonal of W keeps a_ii (gmw81.rs:83-85 rebuilds c_jj from it); the
// running diagonal c_ii that the pivot search and gmw81.rs:104-109 use lives in cd.
// dot_J = sum_{s<J} lrow[s] * W(i, s) in ndarray's order, lrow in registers, J a compile-time constant
template <int N, int J>
__device__ __forceinline__ void tps_gmw_column(double* w, const double* lrow_s)
{
    double lr[J > 0 ? J : 1];
#pragma unroll
    for (int s = 0; s < J; ++s)
        lr[s] = lrow_s[s];
    for (int i = J; i < N; ++i) {
        double* wi = w + tps_tri(i);
        double acc = 0.0;
        if constexpr (J >= 8) {
            double pp[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                pp[u] = 0.0;
#pragma unroll
            for (int s = 0; s < (J / 8) * 8; ++s)
                pp[s % 8] = dadd(pp[s % 8], dmul(lr[s], wi[s]));
            acc = dadd(acc, dadd(pp[0], pp[4]));
            acc = dadd(acc, dadd(pp[1], pp[5]));
            acc = dadd(acc, dadd(pp[2], pp[6]));
            acc = dadd(acc, dadd(pp[3], pp[7]));
        }
#pragma unroll
        for (int s = (J / 8) * 8; s < J; ++s)
            acc = dadd(acc, dmul(lr[s], wi[s]));
        wi[J] = dsub(wi[J], acc);   // gmw81.rs:83-85: c_iJ = l_iJ - sum
    }
}
template <int N, int J>
__device__ __forceinline__ void tps_gmw_column_at(double* w, const double* lrow_s, int j)
{
    if constexpr (J < N) {
        if (j == J)
            tps_gmw_column<N, J>(w, lrow_s);
        else
            tps_gmw_column_at<N, J + 1>(w, lrow_s, j);
    }
}

// w: packed triangle (A's lower triangle on entry, L on exit); lrow, cd, d: N doubles each; e_j
// takes the slot of the running diagonal of position j, which is dead once j has been the pivot
template <int N>
__device__ __forceinline__ void tps_gmw(double* w, double* lrow, double* cd, double* d, TpsPerm& p)
{
    double diag_max = 0.0, off_max = 0.0;   // gmw81.rs:49-58 (the upper triangle mirrors the lower)
    for (int i = 0; i < N; ++i) {
        const int ti = tps_tri(i);
        d[i] = 0.0;
        cd[i] = w[ti + i];   // gmw81.rs:66-68
        const double v = fabs(w[ti + i]);
        if (v > diag_max) diag_max = v;
        for (int q = 0; q < i; ++q) {
            const double o = fabs(w[ti + q]);
            if (o > off_max) off_max = o;
        }
    }
    const double nf = (double)N;
    const double delta = dmul(MODCHOL_EPS, fmax(1.0, dadd(diag_max, off_max)));   // gmw81.rs:60
    const double beta = dsqrt(fmax(fmax(diag_max, ddiv(off_max, dsqrt(dsub(dmul(nf, nf), 1.0)))), MODCHOL_EPS));   // :61-64
    for (int j = 0; j < N; ++j) {
        const int tj = tps_tri(j);
        int m = j;   // utils.rs:73-91 on the running diagonal (gmw81.rs:71-78)
        double best = fabs(cd[j]);
        for (int i = j + 1; i < N; ++i) {
            const double v = fabs(cd[i]);
            if (v > best) {
                best = v;
                m = i;
            }
        }
        if (m != j) {
            tps_swap(w, N, j, m);
            p.swap(j, m);
            const double t = cd[j];
            cd[j] = cd[m];
            cd[m] = t;
        }
        for (int s = 0; s < j; ++s)   // gmw81.rs:79-81
            lrow[s] = ddiv(w[tj + s], d[s]);
        tps_gmw_column_at<N, 0>(w, lrow, j);   // gmw81.rs:83-85 for i = j .. n-1
        for (int s = 0; s < j; ++s)   // row j of the factor is final
            w[tj + s] = lrow[s];
        double theta = 0.0;   // gmw81.rs:87-100
        for (int i = j + 1; i < N; ++i) {
            const double v = fabs(w[tps_tri(i) + j]);
            if (v > theta)
                theta = v;
        }
        const double tb = ddiv(theta, beta);
        const double cjj = w[tj + j];
        const double dj = fmax(fmax(fabs(cjj), dmul(tb, tb)), delta);   // gmw81.rs:102
        d[j] = dj;
#pragma unroll 4
        for (int i = j + 1; i < N; ++i) {   // gmw81.rs:104-109
            const double cij = w[tps_tri(i) + j];
            cd[i] = dsub(cd[i], ddiv(dmul(cij, cij), dj));
        }
        cd[j] = dsub(dj, cjj);   // gmw81.rs:110: e_j
    }
    // gmw81.rs:112-125: l . diag(sqrt(d)) with a unit diagonal; the dense product's accumulation
    // from 0.0 turns a -0 into +0
    for (int q = 0; q < N; ++q)
        lrow[q] = dsqrt(d[q]);
    for (int i = 0; i < N; ++i) {
        const int ti = tps_tri(i);
        for (int q = 0; q < i; ++q)
            w[ti + q] = dadd(0.0, dmul(w[ti + q], lrow[q]));
        w[ti + i] = dadd(0.0, lrow[i]);
    }
}

template <int kAlgo, int N>
__global__ void __launch_bounds__(kTpsThreads)
tps_kernel(long long batch, const double* __restrict__ A, double* __restrict__ L,
           double* __restrict__ E, unsigned long long* __restrict__ P, int* __restrict__ K)
{
    extern __shared__ double sm[];
    constexpr int NN = N * N, T = N * (N + 1) / 2, S = tps_thread_doubles(N, kAlgo);
    double* mine = sm + threadIdx.x * S;
    double* g = mine + T;   // SE: g / e; GMW81: the running diagonal / e, then lrow and d
    const bool al16 = (((size_t)A | (size_t)L) & 15) == 0, al32 = (((size_t)L) & 31) == 0;
    const bool al32a = (((size_t)A) & 31) == 0;
    const long long stride = (long long)gridDim.x * kTpsThreads;
    for (long long b = (long long)blockIdx.x * kTpsThreads + threadIdx.x; b < batch; b += stride) {
        // in: the lower triangle straight from this thread's matrix (see tp_load_lower)
        tp_load_lower<N>(A + (size_t)b * NN, al16, [&](int i, int k, double v) { mine[i * (i + 1) / 2 + k] = v; }, al32a);
        int kstart = -1;
        TpsPerm p;
        if constexpr (kAlgo == kGMW81)
            tps_gmw<N>(mine, g + N, g, g + 2 * N, p);
        else
            kstart = tps_se<kAlgo, N>(mine, g, p);
        if (L)   // se90.rs:169-173, se99.rs:189-192, gmw81.rs:119-123
            tp_store_full<N>(L + (size_t)b * NN, al32, al16, [&](int i, int k) { return mine[i * (i + 1) / 2 + k]; });
        for (int i = 0; i < N; ++i) {   // se90.rs:174-178, se99.rs:194-198, gmw81.rs:127-131
            const int pi = p.get(i);
            if (E) E[(size_t)b * N + pi] = g[i];
            if (P) P[(size_t)b * N + i] = (unsigned long long)pi;
        }
        if (K)
            K[b] = kstart;
    }
}

// ---- staged variant: the warp copies its tile of 32 matrices with coalesced accesses (only the
// lower triangle is fetched and kept); ahead of the direct accesses for n <= 11
template <int kAlgo, int N>
__global__ void __launch_bounds__(kTpsThreads)
tps_kernel_staged(long long batch, const double* __restrict__ A, double* __restrict__ L,
                  double* __restrict__ E, unsigned long long* __restrict__ P, int* __restrict__ K)
{
    extern __shared__ double sm[];
    constexpr int NN = N * N, T = N * (N + 1) / 2, S = tps_thread_doubles(N, kAlgo);
    const int lane = threadIdx.x;
    double* mine = sm + lane * S;

    for (long long t0 = (long long)blockIdx.x * kTpsThreads; t0 < batch; t0 += (long long)gridDim.x * kTpsThreads) {
        const int cnt = (int)min((long long)kTpsThreads, batch - t0);
        const int words = cnt * NN;
        {
            const double* src = A + (size_t)t0 * NN;
            for (int q = lane; q < words; q += kTpsThreads) {
                const int mat = q / NN, r = q - mat * NN, i = r / N, k = r - i * N;
                if (k <= i)
                    sm[mat * S + i * (i + 1) / 2 + k] = __ldg(src + q);
            }
        }
        __syncwarp();
        int kstart = -1;
        if (lane < cnt) {
            double* g = mine + T;
            TpsPerm p;
            if constexpr (kAlgo == kGMW81)
                tps_gmw<N>(mine, g + N, g, g + 2 * N, p);
            else
                kstart = tps_se<kAlgo, N>(mine, g, p);
            for (int i = 0; i < N; ++i) {   // se90.rs:174-178, se99.rs:194-198, gmw81.rs:127-131
                const int pi = p.get(i);
                if (E) E[(size_t)(t0 + lane) * N + pi] = g[i];
                if (P) P[(size_t)(t0 + lane) * N + i] = (unsigned long long)pi;
            }
        }
        __syncwarp();
        if (L) {   // L with its strict upper triangle zero (se90.rs:169-173, se99.rs:189-192)
            double* dst = L + (size_t)t0 * NN;
            for (int q = lane; q < words; q += kTpsThreads) {
                const int mat = q / NN, r = q - mat * NN, i = r / N, k = r - i * N;
                dst[q] = k <= i ? sm[mat * S + i * (i + 1) / 2 + k] : 0.0;
            }
        }
        if (K && lane < cnt)
            K[t0 + lane] = kstart;
        __syncwarp();   // the tile's shared memory is free again
    }
}

}  // namespace modchol
