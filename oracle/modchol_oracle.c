/*
 * modchol_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C restatement of the arithmetic of the reference crate
 * argmin-rs/modcholesky v0.2.0 (pure Rust; cannot be compiled in this image: no
 * rustc/cargo, no network).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg may load this file's shared object, and there
 * only as the checker or the timed CPU baseline -- never on the product path.
 *
 * PARITY STATUS: pinned against every literal the reference's own in-file tests hold
 * (lib.rs:27-60, gmw81.rs:143-289, se90.rs:190-314, se99.rs:210-336,
 * gershgorin.rs:47-67, utils.rs:153-215) -- see tests/test_oracle_golden.py.  Those
 * tests only reach n <= 6; for n >= 8 the summation order of ndarray 0.16's
 * `Array1::sum()` (numeric_util::unrolled_fold, eight partial sums) is restated from
 * the published ndarray source, which is NOT vendored under /root/reference, so the
 * last-bit behaviour of norms for n >= 8 is "parity unpinned" (DESIGN.md section 3).
 *
 * Build: see oracle/Makefile.  Flags that matter: -O2 -ffp-contract=off (Rust never
 * contracts a - b*c into an FMA), no -ffast-math.
 *
 * Every function cites the reference file:line it follows.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

#define ORACLE_GMW81 0
#define ORACLE_SE90 1
#define ORACLE_SE99 2

/* cbrt(f64::EPSILON) (se90.rs:51, se99.rs:48-49).  glibc cbrt and mpmath agree on
 * these bits; hard-coded so that every build of the oracle and the CUDA kernels uses
 * exactly one value. */
static const double ORACLE_TAU = 0x1.965fea53d6e3dp-18;
static const double ORACLE_MU = 0.1; /* se99.rs:50 */

/* ------------------------------------------------------------------------------ */
/* third-party arithmetic restated (ndarray 0.16, not vendored in /root/reference)  */
/* ------------------------------------------------------------------------------ */

/* ndarray `Array1::<f64>::sum()` on contiguous data == numeric_util::unrolled_fold
 * with init 0.0 and `+`: eight running partials over chunks of eight, combined as
 * ((((0+(p0+p4))+(p1+p5))+(p2+p6))+(p3+p7)), then the (<8) tail added in order.
 * Call sites: gmw81.rs:84, se90.rs:108-109,124, se99.rs:128-129,144. */
static double nd_sum(const double *x, size_t len)
{
    double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0, p4 = 0.0, p5 = 0.0, p6 = 0.0, p7 = 0.0;
    while (len >= 8) {
        p0 = p0 + x[0];
        p1 = p1 + x[1];
        p2 = p2 + x[2];
        p3 = p3 + x[3];
        p4 = p4 + x[4];
        p5 = p5 + x[5];
        p6 = p6 + x[6];
        p7 = p7 + x[7];
        x += 8;
        len -= 8;
    }
    double acc = 0.0;
    acc = acc + (p0 + p4);
    acc = acc + (p1 + p5);
    acc = acc + (p2 + p6);
    acc = acc + (p3 + p7);
    for (size_t i = 0; i < len && i < 7; ++i)
        acc = acc + x[i];
    return acc;
}

/* ------------------------------------------------------------------------------ */
/* decision-margin bookkeeping (oracle-only diagnostic, not in the reference)       */
/* ------------------------------------------------------------------------------ */
/* The north-star asks for p to be bit-exact "wherever the reference's pivot choice is
 * not a rounding tie".  To make that testable the oracle records, per matrix, the
 * smallest relative gap of any discrete decision it took (pivot argmax: best vs
 * runner-up; phase switches: tested value vs threshold).  A matrix whose margin is
 * far above rounding noise must reproduce p exactly under any arithmetic that agrees
 * to ~1e-13; one whose margin is ~1e-16 is a rounding tie. */
typedef struct {
    double margin;
    double scale;
} margin_t;

static void margin_note(margin_t *m, double a, double b)
{
    if (!m)
        return;
    if (isnan(a) || isnan(b) || isinf(a) || isinf(b))
        return;
    double s = fmax(fmax(fabs(a), fabs(b)), m->scale);
    double g = (s > 0.0) ? fabs(a - b) / s : 0.0;
    if (g < m->margin)
        m->margin = g;
}

/* ------------------------------------------------------------------------------ */
/* utils.rs                                                                         */
/* ------------------------------------------------------------------------------ */

/* utils.rs:52-70 index_of_largest: first strict maximum; the running max starts at
 * element 0, `ci > max` is false for NaN so NaN never replaces it (and a NaN at
 * element 0 sticks).  `v` has stride `inc`. */
static size_t argmax_first(const double *v, size_t inc, size_t len, margin_t *mg)
{
    double best = v[0], second = -INFINITY;
    size_t bi = 0;
    for (size_t i = 0; i < len; ++i) {
        double c = v[i * inc];
        if (c > best) {
            second = best;
            best = c;
            bi = i;
        } else if (i != 0 && c > second) {
            second = c;
        }
    }
    if (len > 1)
        margin_note(mg, best, second);
    return bi;
}

/* utils.rs:73-91 index_of_largest_abs: same scan over |v|. */
static size_t argmax_abs_first(const double *v, size_t inc, size_t len, margin_t *mg)
{
    double best = fabs(v[0]), second = -INFINITY;
    size_t bi = 0;
    for (size_t i = 0; i < len; ++i) {
        double c = fabs(v[i * inc]);
        if (c > best) {
            second = best;
            best = c;
            bi = i;
        } else if (i != 0 && c > second) {
            second = c;
        }
    }
    if (len > 1)
        margin_note(mg, best, second);
    return bi;
}

/* utils.rs:41-49 swap_rows followed by utils.rs:30-38 swap_columns on a full n x n
 * row-major matrix (call sites gmw81.rs:73-76, se90.rs:65-66,117-118,
 * se99.rs:78-79,137-138). */
static void swap_rows_then_cols(double *m, size_t n, size_t r1, size_t r2)
{
    for (size_t c = 0; c < n; ++c) {
        double t = m[r1 * n + c];
        m[r1 * n + c] = m[r2 * n + c];
        m[r2 * n + c] = t;
    }
    for (size_t r = 0; r < n; ++r) {
        double t = m[r * n + r1];
        m[r * n + r1] = m[r * n + r2];
        m[r * n + r2] = t;
    }
}

/* utils.rs:14-27 eigenvalues_2x2 -> (hi, lo).  Reads both off-diagonals. */
void oracle_eigenvalues_2x2(double a, double b, double c, double d, double *hi, double *lo)
{
    double h = -(a + d) / 2.0;
    double t = sqrt(h * h - a * d + b * c);
    double l1 = (a + d) / 2.0 + t;
    double l2 = (a + d) / 2.0 - t;
    if (l1 > l2) {
        *hi = l1;
        *lo = l2;
    } else {
        *hi = l2;
        *lo = l1;
    }
}

/* gershgorin.rs:20-41: (a_ii, min(sum_{j!=i}|a_ij|, sum_{j!=i}|a_ji|)), plain
 * sequential sums.  out is n pairs (centre, radius). */
void oracle_gershgorin_circles(int n, const double *a, double *out)
{
    for (int i = 0; i < n; ++i) {
        double ri = 0.0, ci = 0.0;
        for (int j = 0; j < n; ++j) {
            if (i == j)
                continue;
            ri += fabs(a[(size_t)i * n + j]);
            ci += fabs(a[(size_t)j * n + i]);
        }
        out[2 * i] = a[(size_t)i * n + i];
        out[2 * i + 1] = fmin(ri, ci); /* f64::min */
    }
}

/* ------------------------------------------------------------------------------ */
/* shared tail: zero upper triangle, un-permute E                                   */
/* gmw81.rs:119-131, se90.rs:169-178, se99.rs:189-198                               */
/* ------------------------------------------------------------------------------ */
static void finish_lower_and_e(size_t n, double *l, double *e, const uint64_t *p, double *tmp)
{
    for (size_t i = 0; i + 1 < n; ++i)
        for (size_t c = i + 1; c < n; ++c)
            l[i * n + c] = 0.0;
    memcpy(tmp, e, n * sizeof(double));
    for (size_t i = 0; i < n; ++i)
        e[p[i]] = tmp[i];
}

/* ------------------------------------------------------------------------------ */
/* SE90 / SE99                                                                      */
/* ------------------------------------------------------------------------------ */

/* One right-looking elimination step on the full symmetric working matrix:
 * se90.rs:84-93 == se90.rs:142-151 == se99.rs:96-105 == se99.rs:162-171. */
static void se_eliminate(double *l, size_t n, size_t j)
{
    l[j * n + j] = sqrt(l[j * n + j]);
    const double ljj = l[j * n + j];
    for (size_t i = j + 1; i < n; ++i) {
        l[i * n + j] /= ljj;
        l[j * n + i] /= ljj;
        const double lij = l[i * n + j];
        for (size_t k = j + 1; k <= i; ++k) {
            l[i * n + k] -= lij * l[k * n + j];
            l[k * n + i] = l[i * n + k];
        }
    }
}

/* look-ahead value min_{i>j}(l_ii - l_ij^2 / l_jj): se90.rs:70-77, se99.rs:82-89. */
static double se_lookahead(const double *l, size_t n, size_t j)
{
    double acc = INFINITY;
    for (size_t i = j + 1; i < n; ++i) {
        double lij = l[i * n + j];
        double nv = l[i * n + i] - (lij * lij) / l[j * n + j];
        if (nv < acc)
            acc = nv;
    }
    return acc;
}

/* ws: workspace of at least 3n doubles.  Returns phase-2 start index k, or -1 if the
 * factorisation never left phase 1. */
static int se_factor(int variant, size_t n, const double *a, double *l, double *e, uint64_t *p,
                     double *ws, margin_t *mg)
{
    const double tau = ORACLE_TAU, tau_bar = ORACLE_TAU, mu = ORACLE_MU;
    double *g = ws, *scratch = ws + n, *etmp = ws + 2 * n;
    int k_start = -1;

    memcpy(l, a, n * n * sizeof(double)); /* se90.rs:46, se99.rs:43 */
    for (size_t i = 0; i < n; ++i) {
        e[i] = 0.0;
        p[i] = i;
    }
    double gamma = 0.0; /* se90.rs:55-57, se99.rs:54-56 */
    for (size_t i = 0; i < n; ++i)
        if (fabs(l[i * n + i]) > gamma)
            gamma = fabs(l[i * n + i]);
    if (mg)
        mg->scale = gamma;

    int phaseone = 1;
    size_t j = 0;
    while (j < n && phaseone) {
        if (variant == ORACLE_SE99) { /* se99.rs:62-73 */
            double dmax = -INFINITY, dmin = INFINITY;
            for (size_t i = j; i < n; ++i) {
                double x = l[i * n + i];
                if (x > dmax)
                    dmax = x;
                if (x < dmin)
                    dmin = x;
            }
            margin_note(mg, dmax, tau_bar * gamma);
            margin_note(mg, dmin, -mu * dmax);
            if (dmax < tau_bar * gamma || dmin < -mu * dmax) {
                phaseone = 0;
                break;
            }
        }
        /* se90.rs:63-68, se99.rs:76-81 */
        size_t m = argmax_first(&l[j * n + j], n + 1, n - j, mg);
        if (m != 0) {
            swap_rows_then_cols(l, n, j, j + m);
            uint64_t t = p[j];
            p[j] = p[j + m];
            p[j + m] = t;
        }
        double la = se_lookahead(l, n, j);
        double thresh = (variant == ORACLE_SE99) ? -mu * gamma : tau * gamma; /* se99.rs:90 / se90.rs:78 */
        margin_note(mg, la, thresh);
        if (la < thresh) {
            phaseone = 0;
            break;
        }
        se_eliminate(l, n, j);
        ++j;
    }

    double delta_prev = 0.0;

    if (!phaseone)
        k_start = (int)j;

    /* se99.rs:115-119: a single trailing element is left */
    if (variant == ORACLE_SE99 && !phaseone && j == n - 1) {
        e[j] = -l[j * n + j] + fmax(tau_bar * gamma, tau * (-l[j * n + j]) / (1.0 - tau));
        l[j * n + j] += e[j];
        l[j * n + j] = sqrt(l[j * n + j]);
    }

    /* se90.rs:101-167, se99.rs:121-187 */
    if (!phaseone && (variant == ORACLE_SE90 || j + 1 < n)) {
        const size_t k = j;
        for (size_t i = 0; i < n; ++i)
            g[i] = 0.0;
        /* lower Gershgorin bounds: se90.rs:105-110, se99.rs:125-130 */
        for (size_t i = k; i < n; ++i) {
            size_t c = 0;
            for (size_t q = k; q < i; ++q)
                scratch[c++] = fabs(l[i * n + q]);
            double s1 = nd_sum(scratch, c);
            c = 0;
            for (size_t r = i + 1; r < n; ++r)
                scratch[c++] = fabs(l[r * n + i]);
            double s2 = nd_sum(scratch, c);
            g[i] = l[i * n + i] - s1 - s2;
        }
        for (size_t jj = k; jj + 2 < n; ++jj) {
            /* se90.rs:115-121, se99.rs:135-141 */
            size_t m = argmax_first(&g[jj], 1, n - jj, mg);
            if (m != 0) {
                swap_rows_then_cols(l, n, jj, jj + m);
                uint64_t t = p[jj];
                p[jj] = p[jj + m];
                p[jj + m] = t;
                double tg = g[jj];
                g[jj] = g[jj + m];
                g[jj + m] = tg;
            }
            /* se90.rs:124-131, se99.rs:144-151 */
            size_t c = 0;
            for (size_t r = jj + 1; r < n; ++r)
                scratch[c++] = fabs(l[r * n + jj]);
            double normj = nd_sum(scratch, c);
            /* tau*gamma (se90.rs:127) and tau_bar*gamma (se99.rs:147) are the same value */
            e[jj] = fmax(fmax(0.0, delta_prev), -l[jj * n + jj] + fmax(normj, tau_bar * gamma));
            if (e[jj] > 0.0) {
                l[jj * n + jj] += e[jj];
                delta_prev = e[jj];
            }
            /* se90.rs:134-139, se99.rs:154-159 */
            if (fabs(l[jj * n + jj] - normj) > 1.0 * DBL_EPSILON) {
                double t = 1.0 - normj / l[jj * n + jj];
                for (size_t i = jj + 1; i < n; ++i)
                    g[i] += fabs(l[i * n + jj]) * t;
            }
            se_eliminate(l, n, jj);
        }
        /* final 2x2: se90.rs:155-166, se99.rs:175-186 */
        double lhi, llo;
        oracle_eigenvalues_2x2(l[(n - 2) * n + (n - 2)], l[(n - 2) * n + (n - 1)],
                               l[(n - 1) * n + (n - 2)], l[(n - 1) * n + (n - 1)], &lhi, &llo);
        double en;
        if (variant == ORACLE_SE90)
            en = fmax(fmax(0.0, -llo + tau * fmax(gamma, 1.0 / (1.0 - tau) * (lhi - llo))), delta_prev);
        else
            en = fmax(fmax(0.0, -llo + fmax(tau_bar * gamma, tau * (lhi - llo) / (1.0 - tau))), delta_prev);
        e[n - 2] = en;
        e[n - 1] = en;
        if (en > 0.0) {
            l[(n - 2) * n + (n - 2)] += en;
            l[(n - 1) * n + (n - 1)] += en;
        }
        l[(n - 2) * n + (n - 2)] = sqrt(l[(n - 2) * n + (n - 2)]);
        l[(n - 1) * n + (n - 2)] /= l[(n - 2) * n + (n - 2)];
        double t = l[(n - 1) * n + (n - 2)];
        l[(n - 1) * n + (n - 1)] = sqrt(l[(n - 1) * n + (n - 1)] - t * t);
    }

    finish_lower_and_e(n, l, e, p, etmp);
    return k_start;
}

/* ------------------------------------------------------------------------------ */
/* GMW81  (gmw81.rs:39-134)                                                         */
/* ------------------------------------------------------------------------------ */
/* ws: n*n + 3n doubles. */
static int gmw81_factor(size_t n, const double *a, double *l, double *e, uint64_t *p, double *ws,
                        margin_t *mg)
{
    double *c = ws, *d = ws + n * n, *scratch = d + n, *etmp = scratch + n;

    memcpy(l, a, n * n * sizeof(double)); /* gmw81.rs:45 */
    for (size_t i = 0; i < n; ++i) {
        e[i] = 0.0;
        p[i] = i;
    }
    double diag_max = 0.0, off_max = 0.0; /* gmw81.rs:49-58 */
    for (size_t i = 0; i < n; ++i)
        if (fabs(l[i * n + i]) > diag_max)
            diag_max = fabs(l[i * n + i]);
    for (size_t i = 0; i < n; ++i)
        for (size_t q = 0; q < n; ++q)
            if (i != q && fabs(l[i * n + q]) > off_max)
                off_max = fabs(l[i * n + q]);
    if (mg)
        mg->scale = diag_max;

    const double nn = (double)n;
    const double delta = DBL_EPSILON * fmax(1.0, diag_max + off_max);                        /* :60 */
    const double beta = sqrt(fmax(fmax(diag_max, off_max / sqrt(nn * nn - 1.0)), DBL_EPSILON)); /* :61-64 */

    memset(c, 0, n * n * sizeof(double)); /* gmw81.rs:66-68 */
    for (size_t i = 0; i < n; ++i) {
        c[i * n + i] = l[i * n + i];
        d[i] = 0.0;
    }

    for (size_t j = 0; j < n; ++j) {
        size_t m = argmax_abs_first(&c[j * n + j], n + 1, n - j, mg); /* :71-78 */
        if (m != 0) {
            swap_rows_then_cols(c, n, j, j + m);
            swap_rows_then_cols(l, n, j, j + m);
            uint64_t t = p[j];
            p[j] = p[j + m];
            p[j + m] = t;
        }
        for (size_t s = 0; s < j; ++s) /* :79-81 */
            l[j * n + s] = c[j * n + s] / d[s];
        for (size_t i = j; i < n; ++i) { /* :83-85 */
            for (size_t s = 0; s < j; ++s)
                scratch[s] = l[j * n + s] * c[i * n + s];
            c[i * n + j] = l[i * n + j] - nd_sum(scratch, j);
        }
        double theta = 0.0; /* :87-100 */
        if (j + 1 < n)
            for (size_t i = j + 1; i < n; ++i)
                if (fabs(c[i * n + j]) > theta)
                    theta = fabs(c[i * n + j]);
        double tb = theta / beta;
        d[j] = fmax(fmax(fabs(c[j * n + j]), tb * tb), delta); /* :102 */
        if (j + 1 < n)                                            /* :104-109 */
            for (size_t i = j + 1; i < n; ++i) {
                double c2 = c[i * n + j] * c[i * n + j];
                c[i * n + i] -= c2 / d[j];
            }
        e[j] = d[j] - c[j * n + j]; /* :110 */
    }

    /* gmw81.rs:112-125: unit diagonal, zero upper, l = l . diag(sqrt(d)).  The dense
     * `dot` has exactly one non-zero term per output element; the accumulation of the
     * remaining exact zeros only normalises a -0 result to +0 (0.0 + x).
     * KNOWN DIVERGENCE for NON-FINITE factors: in the reference's dense product a NaN or
     * Inf entry l_ik meets the zeros of dout's other columns (NaN * 0, Inf * 0) and turns
     * the WHOLE row i of L into NaN, and an infinite sqrt(d_q) meets the zeros of the upper
     * triangle (0 * Inf); this restatement (and the CUDA kernels) propagate non-finite
     * values only through the one meaningful term.  Finite factorisations are unaffected. */
    for (size_t i = 0; i < n; ++i) {
        for (size_t q = 0; q < n; ++q) {
            double v;
            if (q > i)
                v = 0.0;
            else if (q == i)
                v = 1.0 * sqrt(d[q]);
            else
                v = l[i * n + q] * sqrt(d[q]);
            l[i * n + q] = 0.0 + v;
        }
    }
    memcpy(etmp, e, n * sizeof(double)); /* :127-131 */
    for (size_t i = 0; i < n; ++i)
        e[p[i]] = etmp[i];
    return -1;
}

/* ------------------------------------------------------------------------------ */
/* public entry points (called through ctypes by tests/ and bench.py only)          */
/* ------------------------------------------------------------------------------ */

size_t oracle_workspace_doubles(int n)
{
    size_t nn = (size_t)n;
    return nn * nn + 4 * nn + 8;
}

/* Factorise ONE n x n row-major matrix.  l: n*n, e: n, p: n (u64, = Rust usize).
 * phase2_start (nullable): k at which SE90/SE99 entered phase 2, -1 if never / GMW81.
 * margin (nullable): smallest relative decision gap (see margin_note).
 * Returns 0, or -1 on bad arguments (the reference panics for n == 0:
 * `0..(n-1)` underflows, gmw81.rs:120, se90.rs:170, se99.rs:190). */
int oracle_modchol(int algo, int n, const double *a, double *l, double *e, uint64_t *p,
                   int32_t *phase2_start, double *margin, double *ws)
{
    if (n < 1 || !a || !l || !e || !p || algo < 0 || algo > 2)
        return -1;
    int own = 0;
    if (!ws) {
        ws = (double *)malloc(oracle_workspace_doubles(n) * sizeof(double));
        if (!ws)
            return -2;
        own = 1;
    }
    margin_t mg = {INFINITY, 0.0};
    int k;
    if (algo == ORACLE_GMW81)
        k = gmw81_factor((size_t)n, a, l, e, p, ws, &mg);
    else
        k = se_factor(algo, (size_t)n, a, l, e, p, ws, &mg);
    if (phase2_start)
        *phase2_start = k;
    if (margin)
        *margin = mg.margin;
    if (own)
        free(ws);
    return 0;
}

/* The reference's path looped over a batch (the north-star's "ndarray path looped
 * with rayon" stand-in), on plain pthreads with a shared atomic work counter.
 * nthreads <= 0 means all online cores.  Returns the number of threads used, or a
 * negative error. */
typedef struct {
    int algo, n;
    int64_t batch;
    const double *a;
    double *l, *e;
    uint64_t *p;
    int32_t *k;
    double *margin;
    atomic_llong *next;
    atomic_int *err;
} batch_job_t;

#define ORACLE_CHUNK 64

static void *batch_worker(void *arg)
{
    batch_job_t *jb = (batch_job_t *)arg;
    const size_t n = (size_t)jb->n, nn = n * n;
    double *ws = (double *)malloc(oracle_workspace_doubles(jb->n) * sizeof(double));
    if (!ws) {
        atomic_store(jb->err, 1);
        return NULL;
    }
    for (;;) {
        long long b0 = atomic_fetch_add(jb->next, ORACLE_CHUNK);
        if (b0 >= jb->batch)
            break;
        long long b1 = b0 + ORACLE_CHUNK < jb->batch ? b0 + ORACLE_CHUNK : jb->batch;
        for (long long b = b0; b < b1; ++b)
            oracle_modchol(jb->algo, jb->n, jb->a + (size_t)b * nn, jb->l + (size_t)b * nn,
                           jb->e + (size_t)b * n, jb->p + (size_t)b * n, jb->k ? jb->k + b : NULL,
                           jb->margin ? jb->margin + b : NULL, ws);
    }
    free(ws);
    return NULL;
}

int oracle_max_threads(void)
{
    long c = sysconf(_SC_NPROCESSORS_ONLN);
    return c > 0 ? (int)c : 1;
}

int oracle_modchol_batched(int algo, int64_t batch, int n, const double *a, double *l, double *e,
                           uint64_t *p, int32_t *phase2_start, double *margin, int nthreads)
{
    if (batch < 0 || n < 1 || algo < 0 || algo > 2)
        return -1;
    if (nthreads <= 0)
        nthreads = oracle_max_threads();
    if (nthreads > 1024)
        nthreads = 1024;
    if ((int64_t)nthreads > (batch + ORACLE_CHUNK - 1) / ORACLE_CHUNK)
        nthreads = (int)((batch + ORACLE_CHUNK - 1) / ORACLE_CHUNK);
    if (nthreads < 1)
        nthreads = 1;
    atomic_llong next = 0;
    atomic_int err = 0;
    batch_job_t jb = {algo, n, batch, a, l, e, p, phase2_start, margin, &next, &err};
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    if (!th)
        return -2;
    int started = 0;
    for (int t = 1; t < nthreads; ++t) {
        if (pthread_create(&th[started], NULL, batch_worker, &jb) != 0)
            break;
        ++started;
    }
    batch_worker(&jb); /* the calling thread works too */
    for (int t = 0; t < started; ++t)
        pthread_join(th[t], NULL);
    free(th);
    return atomic_load(&err) ? -2 : started + 1;
}

double oracle_tau(void) { return ORACLE_TAU; }
