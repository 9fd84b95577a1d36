// bsm_inst.cu -- instantiates the blocked large-n kernels (bsm_kernels.cuh) and their launcher.
#include <cstdlib>

#include "bsm_kernels.cuh"
#include "launch_cache.cuh"
#include "warp_kernels.cuh"

namespace modchol {

template <int kThreads, int KB, int kMinBlocks>
static cudaError_t launch_bsm_gmw(int n, long long batch, const double* a, double* l, double* e,
                                  unsigned long long* p, int* k, int sms, cudaStream_t st)
{
    auto kern = bsm_gmw_kernel<kThreads, KB, kMinBlocks>;
    const size_t smem = bsm_smem_doubles(n, KB) * sizeof(double);
    const size_t smem_max = bsm_smem_doubles(kThreads, KB) * sizeof(double);
    int occ = 1;
    cudaError_t ce = cached_occupancy(kern, kThreads, smem, smem_max, &occ);
    if (ce != cudaSuccess) return ce;
    if (const char* f = getenv("MODCHOL_BSM_OCC"))   // tuning knob: resident blocks per SM
        occ = atoi(f) > 0 && atoi(f) < occ ? atoi(f) : occ;
    long long blocks = batch;
    const long long resident = (long long)sms * occ;
    if (blocks > resident) blocks = resident;   // persistent: every block strides over the batch
    kern<<<(unsigned)blocks, kThreads, smem, st>>>(n, batch, a, l, e, p, k);
    return cudaGetLastError();
}

template <int kAlgo, int kThreads, int KB, int kMinBlocks>
static cudaError_t launch_bsm_se(int n, long long batch, const double* a, double* l, double* e,
                                 unsigned long long* p, int* k, int sms, cudaStream_t st)
{
    auto kern = bsm_se_kernel<kAlgo, kThreads, KB, kMinBlocks>;
    const size_t smem = bsm_smem_doubles(n, KB) * sizeof(double);
    const size_t smem_max = bsm_smem_doubles(kThreads, KB) * sizeof(double);
    int occ = 1;
    cudaError_t ce = cached_occupancy(kern, kThreads, smem, smem_max, &occ);
    if (ce != cudaSuccess) return ce;
    if (const char* f = getenv("MODCHOL_BSM_OCC"))   // tuning knob: resident blocks per SM
        occ = atoi(f) > 0 && atoi(f) < occ ? atoi(f) : occ;
    long long blocks = batch;
    const long long resident = (long long)sms * occ;
    if (blocks > resident) blocks = resident;   // persistent: every block strides over the batch
    kern<<<(unsigned)blocks, kThreads, smem, st>>>(n, batch, a, l, e, p, k);
    return cudaGetLastError();
}

template <int kAlgo>
static cudaError_t launch_bsm_se_n(int n, long long batch, const double* a, double* l, double* e,
                                   unsigned long long* p, int* k, int sms, cudaStream_t st)
{
    if (n <= 256)
        return launch_bsm_se<kAlgo, 256, 32, 3>(n, batch, a, l, e, p, k, sms, st);
    if (n <= 512)
        return launch_bsm_se<kAlgo, 512, 32, 1>(n, batch, a, l, e, p, k, sms, st);
    return launch_bsm_se<kAlgo, 1024, 16, 1>(n, batch, a, l, e, p, k, sms, st);
}

cudaError_t launch_bsm_kernel(int algo, int n, long long batch, const double* a, double* l, double* e,
                              unsigned long long* p, int* k, int sms, cudaStream_t st)
{
    if (algo == kSE90)
        return launch_bsm_se_n<kSE90>(n, batch, a, l, e, p, k, sms, st);
    if (algo == kSE99)
        return launch_bsm_se_n<kSE99>(n, batch, a, l, e, p, k, sms, st);
    if (algo != kGMW81)
        return cudaErrorNotSupported;
    const char* kbe = getenv("MODCHOL_BSM_KB");      // tuning knob: panel width
    const int kb = kbe ? atoi(kbe) : 0;
    if (n <= 256)   // measured at n = 256 x 4096: KB 32 / 3 blocks per SM 0.58 M/s, KB 16 / 4 blocks 0.46
        return kb == 16 ? launch_bsm_gmw<256, 16, 4>(n, batch, a, l, e, p, k, sms, st)
                        : launch_bsm_gmw<256, 32, 3>(n, batch, a, l, e, p, k, sms, st);
    if (n <= 512)
        return kb == 16 ? launch_bsm_gmw<512, 16, 2>(n, batch, a, l, e, p, k, sms, st)
                        : launch_bsm_gmw<512, 32, 1>(n, batch, a, l, e, p, k, sms, st);
    return launch_bsm_gmw<1024, 16, 1>(n, batch, a, l, e, p, k, sms, st);
}

}  // namespace modchol
