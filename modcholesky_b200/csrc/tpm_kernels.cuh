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

}  // namespace modchol
