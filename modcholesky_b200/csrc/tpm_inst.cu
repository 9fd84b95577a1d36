// tpm_inst.cu -- instantiates the thread-per-matrix kernels (tpm_kernels.cuh) and their launcher.
#include "launch_cache.cuh"
#include "tpm_kernels.cuh"
#include "warp_kernels.cuh"

namespace modchol {

template <int kAlgo>
static cudaError_t launch_tpm_algo(int n, long long batch, const double* a, double* l, double* e,
                                   unsigned long long* p, int* k, int sms, cudaStream_t st)
{
    auto kern = tpm_kernel<kAlgo>;
    const size_t smem = (size_t)kTpmThreads * tpm_thread_doubles(kAlgo, n) * sizeof(double);
    const size_t smem_max = (size_t)kTpmThreads * tpm_thread_doubles(kAlgo, kTpmMaxN) * sizeof(double);
    int occ = 1;
    cudaError_t ce = cached_occupancy(kern, kTpmThreads, smem, smem_max, &occ);
    if (ce != cudaSuccess) return ce;
    long long blocks = (batch + kTpmThreads - 1) / kTpmThreads;
    const long long resident = (long long)sms * occ;
    if (blocks > resident) blocks = resident;   // persistent: every block strides over the tiles
    kern<<<(unsigned)blocks, kTpmThreads, smem, st>>>(n, batch, a, l, e, p, k);
    return cudaGetLastError();
}

cudaError_t launch_tpm_kernel(int algo, int n, long long batch, const double* a, double* l, double* e,
                              unsigned long long* p, int* k, int sms, cudaStream_t st)
{
    if (n < 1 || n > kTpmMaxN)
        return cudaErrorNotSupported;
    if (algo == kGMW81) return launch_tpm_algo<kGMW81>(n, batch, a, l, e, p, k, sms, st);
    if (algo == kSE90) return launch_tpm_algo<kSE90>(n, batch, a, l, e, p, k, sms, st);
    if (algo == kSE99) return launch_tpm_algo<kSE99>(n, batch, a, l, e, p, k, sms, st);
    return cudaErrorNotSupported;
}

}  // namespace modchol
