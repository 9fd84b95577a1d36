// warp_kernels.cuh -- dispatch of the warp-per-matrix, register-resident kernels (n <= 32).
// The kernels are instantiated in separate translation units (warp_inst.cu compiled once per
// (algorithm, mode) so that the build parallelises); this header only declares the launchers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace modchol {

constexpr int kWarpsPerBlock = 4;

inline int warp_nmax(int n) { return n <= 8 ? 8 : (n <= 16 ? 16 : (n <= 24 ? 24 : 32)); }

inline bool warp_kernel_available(int algo, int mode, int n)
{
    (void)mode;
    return n >= 1 && n <= 32 && algo >= 0 && algo <= 2;
}

// defined in warp_inst.cu (one object per algorithm/mode pair): launch_warp_a<algo>_s<strict>
#define MODCHOL_DECL_WARP(A, S)                                                                  \
    cudaError_t launch_warp_a##A##_s##S(int n, long long batch, const double* a, double* l,      \
                                        double* e, unsigned long long* p, int* k, int sms,       \
                                        cudaStream_t st);
MODCHOL_DECL_WARP(0, 1)
MODCHOL_DECL_WARP(0, 0)
MODCHOL_DECL_WARP(1, 1)
MODCHOL_DECL_WARP(1, 0)
MODCHOL_DECL_WARP(2, 1)
MODCHOL_DECL_WARP(2, 0)

// shared-memory-resident warp kernels (wsm_se.cuh, wsm_gmw.cuh), n <= 64; defined in wsm_inst.cu
#define MODCHOL_DECL_WSM(A, S)                                                                   \
    cudaError_t launch_wsm_a##A##_s##S(int n, long long batch, const double* a, double* l,       \
                                       double* e, unsigned long long* p, int* k,                 \
                                       const double* rhs, double* x, int nrhs, int sms,          \
                                       cudaStream_t st);
MODCHOL_DECL_WSM(0, 1)
MODCHOL_DECL_WSM(0, 0)
MODCHOL_DECL_WSM(1, 1)
MODCHOL_DECL_WSM(1, 0)
MODCHOL_DECL_WSM(2, 1)
MODCHOL_DECL_WSM(2, 0)

inline bool wsm_kernel_available(int algo, int mode, int n)
{
    (void)mode;
    return n >= 1 && n <= 64 && algo >= 0 && algo <= 2;
}

// l, e, p, k may each be NULL; x != NULL adds the fused solve of (A + E) x = rhs (nrhs per matrix)
inline cudaError_t launch_wsm_kernel(int algo, int mode, int n, long long batch, const double* a,
                                     double* l, double* e, unsigned long long* p, int* k, int sms,
                                     cudaStream_t st, const double* rhs = nullptr,
                                     double* x = nullptr, int nrhs = 0)
{
    const bool strict = mode == 0;
    if (algo == 0)
        return strict ? launch_wsm_a0_s1(n, batch, a, l, e, p, k, rhs, x, nrhs, sms, st)
                      : launch_wsm_a0_s0(n, batch, a, l, e, p, k, rhs, x, nrhs, sms, st);
    if (algo == 1)
        return strict ? launch_wsm_a1_s1(n, batch, a, l, e, p, k, rhs, x, nrhs, sms, st)
                      : launch_wsm_a1_s0(n, batch, a, l, e, p, k, rhs, x, nrhs, sms, st);
    if (algo == 2)
        return strict ? launch_wsm_a2_s1(n, batch, a, l, e, p, k, rhs, x, nrhs, sms, st)
                      : launch_wsm_a2_s0(n, batch, a, l, e, p, k, rhs, x, nrhs, sms, st);
    return cudaErrorNotSupported;
}

// block-per-matrix packed shared-memory kernels (csm_se.cuh, csm_gmw.cuh), 64 < n <= 256;
// defined in csm_inst.cu
#define MODCHOL_DECL_CSM(A, S)                                                                   \
    cudaError_t launch_csm_a##A##_s##S(int n, long long batch, const double* a, double* l,       \
                                       double* e, unsigned long long* p, int* k, int sms,        \
                                       cudaStream_t st);
MODCHOL_DECL_CSM(0, 1)
MODCHOL_DECL_CSM(0, 0)
MODCHOL_DECL_CSM(1, 1)
MODCHOL_DECL_CSM(1, 0)
MODCHOL_DECL_CSM(2, 1)
MODCHOL_DECL_CSM(2, 0)

constexpr int kCsmMaxN = 236;        // packed triangle + line + scratch <= 227 KB
constexpr int kCsmSpillMaxN = 256;   // one thread per row, 256 threads; first rows spilled to L2

inline bool csm_kernel_available(int algo, int mode, int n)
{
    (void)mode;
    return n > 64 && n <= kCsmSpillMaxN && algo >= 0 && algo <= 2;
}

inline cudaError_t launch_csm_kernel(int algo, int mode, int n, long long batch, const double* a,
                                     double* l, double* e, unsigned long long* p, int* k, int sms,
                                     cudaStream_t st)
{
    const bool strict = mode == 0;
    if (algo == 0)
        return strict ? launch_csm_a0_s1(n, batch, a, l, e, p, k, sms, st)
                      : launch_csm_a0_s0(n, batch, a, l, e, p, k, sms, st);
    if (algo == 1)
        return strict ? launch_csm_a1_s1(n, batch, a, l, e, p, k, sms, st)
                      : launch_csm_a1_s0(n, batch, a, l, e, p, k, sms, st);
    if (algo == 2)
        return strict ? launch_csm_a2_s1(n, batch, a, l, e, p, k, sms, st)
                      : launch_csm_a2_s0(n, batch, a, l, e, p, k, sms, st);
    return cudaErrorNotSupported;
}

// blocked large-n kernels (bsm_kernels.cuh), FAST mode, 128 < n <= 1024; defined in bsm_inst.cu
constexpr int kBsmMaxN = 1024;
cudaError_t launch_bsm_kernel(int algo, int n, long long batch, const double* a, double* l, double* e,
                              unsigned long long* p, int* k, int sms, cudaStream_t st);
inline bool bsm_kernel_available(int algo, int mode, int n)
{
    return mode == 1 && algo >= 0 && algo <= 2 && n > 128 && n <= kBsmMaxN;
}

// warp-per-matrix large-n kernels (lwm_kernels.cuh), FAST mode, 64 < n <= 256; defined in lwm_inst.cu
constexpr int kLwmMaxN = 256;
#define MODCHOL_DECL_LWM(a)                                                                             \
    cudaError_t launch_lwm_a##a(int n, long long batch, const double* a_, double* l, double* e,        \
                                unsigned long long* p, int* k, int sms, cudaStream_t st);
MODCHOL_DECL_LWM(0)
MODCHOL_DECL_LWM(1)
MODCHOL_DECL_LWM(2)
#undef MODCHOL_DECL_LWM
inline bool lwm_kernel_available(int algo, int mode, int n)
{
    return mode == 1 && algo >= 0 && algo <= 2 && n > 64 && n <= kLwmMaxN;
}
inline cudaError_t launch_lwm_kernel(int algo, int n, long long batch, const double* a, double* l, double* e,
                                     unsigned long long* p, int* k, int sms, cudaStream_t st)
{
    if (algo == 0) return launch_lwm_a0(n, batch, a, l, e, p, k, sms, st);
    if (algo == 1) return launch_lwm_a1(n, batch, a, l, e, p, k, sms, st);
    if (algo == 2) return launch_lwm_a2(n, batch, a, l, e, p, k, sms, st);
    return cudaErrorNotSupported;
}

// thread-per-matrix kernels (tpm_kernels.cuh), both modes (the reference's arithmetic element for
// element), n <= 16; defined in tpm_inst.cu
constexpr int kTpmMaxNHost = 8;    // registers / run-time n, all three algorithms
constexpr int kTpsMaxNHost = 16;   // triangle in thread-private shared memory
cudaError_t launch_tpm_kernel(int algo, int n, long long batch, const double* a, double* l, double* e,
                              unsigned long long* p, int* k, int sms, cudaStream_t st);
cudaError_t launch_tpm_solve_kernel(int algo, int n, long long batch, const double* a, double* l, double* e,
                                    unsigned long long* p, int* k, const double* rhs, double* x, int nrhs,
                                    int sms, cudaStream_t st);
inline bool tpm_kernel_available(int algo, int n)
{
    return algo >= 0 && algo <= 2 && n >= 1 && n <= kTpsMaxNHost;
}

inline cudaError_t launch_warp_kernel(int algo, int mode, int n, long long batch, const double* a,
                                      double* l, double* e, unsigned long long* p, int* k, int sms,
                                      cudaStream_t st)
{
    const bool strict = mode == 0;
    if (algo == 0)
        return strict ? launch_warp_a0_s1(n, batch, a, l, e, p, k, sms, st)
                      : launch_warp_a0_s0(n, batch, a, l, e, p, k, sms, st);
    if (algo == 1)
        return strict ? launch_warp_a1_s1(n, batch, a, l, e, p, k, sms, st)
                      : launch_warp_a1_s0(n, batch, a, l, e, p, k, sms, st);
    if (algo == 2)
        return strict ? launch_warp_a2_s1(n, batch, a, l, e, p, k, sms, st)
                      : launch_warp_a2_s0(n, batch, a, l, e, p, k, sms, st);
    return cudaErrorNotSupported;
}

}  // namespace modchol
