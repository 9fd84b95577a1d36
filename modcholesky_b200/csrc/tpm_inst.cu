// tpm_inst.cu -- instantiates the thread-per-matrix kernels (tpm_kernels.cuh) and their launcher.
#include <cstdlib>

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

// compile-time n, everything in registers (n <= kTprMaxN)
template <int kAlgo, int N>
static cudaError_t launch_tpr_n(long long batch, const double* a, double* l, double* e,
                                unsigned long long* p, int* k, int sms, cudaStream_t st)
{
    auto kern = tpr_kernel<kAlgo, N>;
    constexpr size_t smem = (size_t)kTpmThreads * tpr_thread_doubles(N) * sizeof(double);
    int occ = 1;
    cudaError_t ce = cached_occupancy(kern, kTpmThreads, smem, smem, &occ);
    if (ce != cudaSuccess) return ce;
    long long blocks = (batch + kTpmThreads - 1) / kTpmThreads;
    const long long resident = (long long)sms * occ;
    if (blocks > resident) blocks = resident;
    kern<<<(unsigned)blocks, kTpmThreads, smem, st>>>(batch, a, l, e, p, k);
    return cudaGetLastError();
}
template <int kAlgo>
static cudaError_t launch_tpr_algo(int n, long long batch, const double* a, double* l, double* e,
                                   unsigned long long* p, int* k, int sms, cudaStream_t st)
{
    switch (n) {
    case 1: return launch_tpr_n<kAlgo, 1>(batch, a, l, e, p, k, sms, st);
    case 2: return launch_tpr_n<kAlgo, 2>(batch, a, l, e, p, k, sms, st);
    case 3: return launch_tpr_n<kAlgo, 3>(batch, a, l, e, p, k, sms, st);
    case 4: return launch_tpr_n<kAlgo, 4>(batch, a, l, e, p, k, sms, st);
    case 5: return launch_tpr_n<kAlgo, 5>(batch, a, l, e, p, k, sms, st);
    case 6: return launch_tpr_n<kAlgo, 6>(batch, a, l, e, p, k, sms, st);
    case 7: return launch_tpr_n<kAlgo, 7>(batch, a, l, e, p, k, sms, st);
    case 8: return launch_tpr_n<kAlgo, 8>(batch, a, l, e, p, k, sms, st);
    default: return cudaErrorNotSupported;
    }
}
static_assert(kTprMaxN == 8, "launch_tpr_algo lists the instantiations");

static bool tpr_enabled()   // MODCHOL_TPR=0: the run-time-n kernels for every n <= 8 (A/B knob)
{
    const char* v = getenv("MODCHOL_TPR");
    return !(v && v[0] == '0');
}

cudaError_t launch_tpm_kernel(int algo, int n, long long batch, const double* a, double* l, double* e,
                              unsigned long long* p, int* k, int sms, cudaStream_t st)
{
    if (n < 1 || n > kTpmMaxN)
        return cudaErrorNotSupported;
    if (n <= kTprMaxN && tpr_enabled()) {
        if (algo == kGMW81) return launch_tpr_algo<kGMW81>(n, batch, a, l, e, p, k, sms, st);
        if (algo == kSE90) return launch_tpr_algo<kSE90>(n, batch, a, l, e, p, k, sms, st);
        if (algo == kSE99) return launch_tpr_algo<kSE99>(n, batch, a, l, e, p, k, sms, st);
    }
    if (algo == kGMW81) return launch_tpm_algo<kGMW81>(n, batch, a, l, e, p, k, sms, st);
    if (algo == kSE90) return launch_tpm_algo<kSE90>(n, batch, a, l, e, p, k, sms, st);
    if (algo == kSE99) return launch_tpm_algo<kSE99>(n, batch, a, l, e, p, k, sms, st);
    return cudaErrorNotSupported;
}

}  // namespace modchol
