/*
 * modchol_b200.h -- C-ABI of the B200-native batched modified-Cholesky engine.
 *
 * This is the drop-in boundary for the hot path of argmin-rs/modcholesky v0.2.0.  The
 * reference has no FFI of its own: its "operator API" is three Rust trait methods on
 * ndarray::Array2<f64>,
 *
 *     ModCholeskyGMW81::mod_cholesky_gmw81(&self) -> Decomposition<L,E,P>   gmw81.rs:24-41
 *     ModCholeskySE90 ::mod_cholesky_se90 (&self) -> Decomposition<L,E,P>   se90.rs:24-40
 *     ModCholeskySE99 ::mod_cholesky_se99 (&self) -> Decomposition<L,E,P>   se99.rs:21-37
 *     struct Decomposition<L,E,P> { l, e, p }                               lib.rs:109-119
 *     GershgorinCircles::gershgorin_circles(&self) -> Vec<(f64,f64)>        gershgorin.rs:15-42
 *
 * so the entry points below are what a Rust `extern "C"` block for that path binds
 * (rust/src/ffi.rs; INTEGRATION.md shows the binding a maintainer would add).  Plain
 * pointers and sizes only; no torch / CUDA types in any signature (`stream` is an
 * opaque cudaStream_t passed as void*, NULL = the legacy default stream).
 *
 * Data layout (= ndarray standard layout, so Array2/Array3 buffers are passed as is):
 *   a, l : batch x n x n  f64, row-major, contiguous.  `a` must be symmetric; it is
 *          never written.  `l` is lower triangular with the strict upper triangle
 *          explicitly zero (se99.rs:189-192).
 *   e    : batch x n      f64, diagonal of E in ORIGINAL index order (se99.rs:194-198)
 *   p    : batch x n      u64 (Rust usize), p[i] = original index at permuted position i
 * such that  P (A + diag(e)) P^T = L L^T  with  P[i, p[i]] = 1  (utils.rs:94-101).
 *
 * Errors: the reference panics (assert!(is_square), se99.rs:38) or silently
 * propagates NaN/Inf; here every call returns a status (0 = ok, negative = error,
 * text via modchol_last_error()), NaN/Inf are propagated exactly like the reference (one
 * exception: GMW81's final dense product l.dot(dout), gmw81.rs:125, spreads a non-finite
 * entry of L over its whole row through NaN * 0 terms; here it stays in its element),
 * and there is NO CPU fallback: without a usable CUDA device every compute call fails
 * with MODCHOL_ERR_NO_DEVICE.
 *
 * Thread safety: all entry points are re-entrant; the only shared state is a
 * mutex-guarded per-device cache of streams and staging buffers.
 */
#ifndef MODCHOL_B200_H
#define MODCHOL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MODCHOL_VERSION 100 /* 0.1.0 */
#define MODCHOL_MAX_N 2048

/* algorithm selector: which reference trait method the call replaces */
#define MODCHOL_GMW81 0 /* gmw81.rs:39-134 */
#define MODCHOL_SE90 1  /* se90.rs:38-181  */
#define MODCHOL_SE99 2  /* se99.rs:35-201  */

/* where a, l, e, p (and the optional info arrays) live */
#define MODCHOL_MEM_HOST 0
#define MODCHOL_MEM_DEVICE 1

/* arithmetic mode.
 * STRICT: every rounding the reference performs is performed, in the same order
 *   (no FMA contraction, true IEEE division and square root, ndarray's eight-partial
 *   summation order for the norms) -- results are bit-identical to the CPU oracle.
 * FAST:  a - b*c contracted to FMA, column scaling by a correctly rounded reciprocal
 *   square root product, tree-ordered norms.  Agrees with STRICT within the
 *   north-star tolerances (p identical off rounding ties, |dE| <= 1e-10 ||E||,
 *   reconstruction <= 1e-12 ||A||). */
#define MODCHOL_MODE_STRICT 0
#define MODCHOL_MODE_FAST 1

/* status codes */
#define MODCHOL_OK 0
#define MODCHOL_ERR_BAD_ARG (-1)     /* null pointer, n < 1, batch < 0, unknown algo/mode/mem */
#define MODCHOL_ERR_CUDA (-2)        /* a CUDA runtime call or kernel failed */
#define MODCHOL_ERR_UNSUPPORTED (-3) /* n > MODCHOL_MAX_N */
#define MODCHOL_ERR_NO_DEVICE (-4)   /* no CUDA device: there is no CPU fallback */
#define MODCHOL_ERR_ALLOC (-5)

/* Batched factorisation; replaces a loop of `a_b.mod_cholesky_{gmw81,se90,se99}()`
 * (gmw81.rs:39, se90.rs:38, se99.rs:35) over b = 0..batch.
 *   phase2_start : optional (NULL ok) batch x i32; index k at which SE90/SE99 left
 *                  phase 1 (se90.rs:102, se99.rs:122), -1 if it never did / for GMW81.
 *   mem_kind     : MODCHOL_MEM_DEVICE - all pointers are device pointers on the current
 *                  device, the work is enqueued on `stream` and the call returns without
 *                  synchronising;  MODCHOL_MEM_HOST - host pointers, the library stages
 *                  the batch through the current device in overlapped chunks and returns
 *                  when the results are in host memory (`stream` is ignored). */
int modchol_factor_batched_f64(int algo, int mode, int64_t batch, int32_t n, const double *a,
                               double *l, double *e, uint64_t *p, int32_t *phase2_start,
                               int mem_kind, void *stream);

/* Same, host pointers only, with the batch split into `ngpus` contiguous slices, one per
 * device 0..ngpus-1 (ngpus <= 0: all visible devices).  The matrices are independent, so
 * there is no collective: each device factorises its slice. */
int modchol_factor_batched_multi_gpu_f64(int algo, int mode, int64_t batch, int32_t n,
                                         const double *a, double *l, double *e, uint64_t *p,
                                         int32_t *phase2_start, int ngpus);

/* Batched GershgorinCircles::gershgorin_circles (gershgorin.rs:20-41):
 * circles[b][i] = (a_ii, min(sum_{j!=i}|a_ij|, sum_{j!=i}|a_ji|)), batch x n x 2 f64.
 * `a` need not be symmetric here. */
int modchol_gershgorin_circles_batched_f64(int64_t batch, int32_t n, const double *a,
                                           double *circles, int mem_kind, void *stream);

/* Batched acceptance check -- the identity every reference test asserts
 * (e.g. se99.rs:221-231 with utils.rs:94-113):
 *   resid[b]   = max_ij |(L L^T)_ij - (P (A + diag(e)) P^T)_ij| / max(1, max_ij |A_ij|)
 *   flags[b]   = bit0: p is not a permutation, bit1: min(e) < 0, bit2: non-zero strict
 *                upper triangle in L, bit3: NaN/Inf in L or e.
 * Either output may be NULL. */
int modchol_verify_batched_f64(int64_t batch, int32_t n, const double *a, const double *l,
                               const double *e, const uint64_t *p, double *resid,
                               int32_t *flags, int mem_kind, void *stream);

/* The step after the path (the caller's Newton step): solve (A + E) x = b for `nrhs` right-hand
 * sides per matrix from a factorisation {l, p} produced above:  L L^T (P x) = P b.
 *   b, x : batch x nrhs x n f64 (each right-hand side contiguous); x may alias b. */
int modchol_solve_batched_f64(int64_t batch, int32_t n, int32_t nrhs, const double *l,
                              const uint64_t *p, const double *b, double *x, int mem_kind,
                              void *stream);

/* The same consumer step FUSED with the factorisation: factorise a_b with `algo` and solve
 * (A_b + E_b) x = b for nrhs right-hand sides per matrix in one call.  For n <= 64 this is ONE
 * kernel and the factor never leaves the SM (shared / tensor memory), so L does not round-trip through HBM
 * (n = 32: 6.4 KB read + 0.25 KB written per matrix instead of 16.9 KB + the solve's 8.5 KB);
 * larger n run the factor kernel into scratch followed by the solve kernel.
 *   b, x : batch x nrhs x n f64; x may alias b.
 *   l, e, p, phase2_start : optional outputs as in modchol_factor_batched_f64 (NULL = not wanted). */
int modchol_factor_solve_batched_f64(int algo, int mode, int64_t batch, int32_t n, int32_t nrhs,
                                     const double *a, const double *b, double *x, double *l,
                                     double *e, uint64_t *p, int32_t *phase2_start, int mem_kind,
                                     void *stream);

/* Device-side assembly of the reference's bench inputs (benches/bench.rs:48-53):
 *   A_b = Q_b D_b Q_b^T,  Q_b = H_0 H_1 ... H_{nrefl-1},  H_r = I - 2 v_r v_r^T / (v_r^T v_r)
 * (utils.rs:116-121 random_householder is H for one seeded v; utils.rs:126-147 random_diagonal
 * is D).  v: batch x nrefl x n Householder vectors, d: batch x n diagonals, a: batch x n x n
 * output, exactly symmetric.  Each reflector is applied as a symmetric rank-2 update
 * (O(nrefl n^2) per matrix); the result equals the reference's dense products to rounding. */
int modchol_assemble_qdqt_batched_f64(int64_t batch, int32_t n, int32_t nrefl, const double *v,
                                      const double *d, double *a, int mem_kind, void *stream);

/* Measured FP64 issue peaks of the current device in TFLOP/s: dependent fused-multiply-add
 * chains on the FP64 pipe and m8n8k4 tensor-pipe operations (either pointer may be NULL).
 * bench.py quotes the large-n kernels against them.  Synchronous, a few milliseconds. */
int modchol_fp64_peak_tflops(double *dfma_tflops, double *dmma_tflops);

/* Name of the kernel class the dispatcher uses for (algo, mode, n) -- e.g. "warp32",
 * "cta_smem", "cta_gmem" -- for logging and tests.  Never NULL. */
const char *modchol_kernel_class(int algo, int mode, int32_t n);

/* Tuning knob: which small-n kernel family serves n <= 32 -- 0 = automatic (one thread per
 * matrix, "thread", for n <= 8 and up to n = 12 (GMW81: 10) in FAST and n = 16 in STRICT mode,
 * batches of 512 (n <= 8) / 4096 (n > 8) matrices and more;
 * the shared-memory-resident "warp_smem" family above), 1 = register-resident "warp32", 2 =
 * shared-memory-resident ("warp_smem" always serves 32 < n <= 64), 3 = "thread" wherever it
 * exists (n <= 16).
 * Returns the previous setting.  Both families produce identical STRICT results; the knob
 * exists for benchmarking and tests. */
int modchol_set_small_kernel_family(int family);

/* Number of kernel launches issued by this library since load (all threads). */
int64_t modchol_launch_count(void);

/* Thread-local text of the last error of the calling thread ("" if none). */
const char *modchol_last_error(void);

int modchol_version(void);

/* Visible CUDA devices (0 if none / no driver). */
int modchol_device_count(void);

/* Frees the cached per-device streams and staging buffers (optional; safe to call
 * when no call is in flight). */
void modchol_shutdown(void);

#ifdef __cplusplus
}
#endif
#endif /* MODCHOL_B200_H */
