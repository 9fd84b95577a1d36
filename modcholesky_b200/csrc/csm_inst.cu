// csm_inst.cu -- instantiates the block-per-matrix packed shared-memory kernels (csm_se.cuh) for
// ONE (algorithm, mode) pair: -DMODCHOL_INST_ALGO={1,2} -DMODCHOL_INST_STRICT={0,1}.
#include "launch_cache.cuh"
#include "warp_kernels.cuh"
#include "csm_gmw.cuh"
#include "csm_se.cuh"

#ifndef MODCHOL_INST_ALGO
#error "compile with -DMODCHOL_INST_ALGO=.. -DMODCHOL_INST_STRICT=.."
#endif

namespace modchol {

template <int kAlgo, bool kStrict, int kThreads, int kMinBlocks, bool kDmma = false>
static cudaError_t launch_csm(int n, long long batch, const double* a, double* l, double* e,
                              unsigned long long* p, int* k, int sms, cudaStream_t st)
{
    auto kern = [] {
        if constexpr (kAlgo == kGMW81)
            return csm_gmw_kernel<kStrict, kThreads, kMinBlocks, false, kDmma && !kStrict>;
        else
            return csm_se_kernel<kAlgo, kStrict, kThreads, kMinBlocks, false, kDmma && !kStrict>;
    }();
    const size_t smem = csm_smem_bytes(n, 0, kDmma && !kStrict);
    int occ = 1;
    cudaError_t ce = cached_occupancy(kern, kThreads, smem, (size_t)227 * 1024, &occ);
    if (ce != cudaSuccess) return ce;
    long long blocks = batch;
    const long long resident = (long long)sms * occ;
    if (blocks > resident) blocks = resident;   // persistent: every block strides over the batch
    kern<<<(unsigned)blocks, kThreads, smem, st>>>(n, batch, a, l, e, p, k, nullptr, 0);
    return cudaGetLastError();
}

// n above the all-in-shared-memory limit: the first S rows of the packed triangle go to a
// stream-ordered global buffer (one slice per resident block; it stays in L1/L2)
template <int kAlgo, bool kStrict, int kThreads>
static cudaError_t launch_csm_spill(int n, long long batch, const double* a, double* l, double* e,
                                    unsigned long long* p, int* k, int sms, cudaStream_t st)
{
    auto kern = [] {
        if constexpr (kAlgo == kGMW81)
            return csm_gmw_kernel<kStrict, kThreads, 1, true, !kStrict>;
        else
            return csm_se_kernel<kAlgo, kStrict, kThreads, 1, true, !kStrict>;
    }();
    int dev = 0, optin = 0;
    cudaError_t ce = cudaGetDevice(&dev);
    if (ce != cudaSuccess) return ce;
    ce = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (ce != cudaSuccess) return ce;
    constexpr bool kPanel = !kStrict;   // the DMMA update's column panel
    int S = 1;
    while (S < n && csm_smem_bytes(n, S, kPanel) > (size_t)optin)
        ++S;
    const size_t smem = csm_smem_bytes(n, S, kPanel);
    if (smem > (size_t)optin) return cudaErrorInvalidValue;
    int occ_unused = 1;
    ce = cached_occupancy(kern, kThreads, smem, (size_t)optin, &occ_unused);
    if (ce != cudaSuccess) return ce;
    long long blocks = batch < sms ? batch : sms;   // one block per SM, persistent
    double* spill = nullptr;
    ce = cudaMallocAsync(&spill, (size_t)blocks * tri(S) * sizeof(double), st);
    if (ce != cudaSuccess) return ce;
    kern<<<(unsigned)blocks, kThreads, smem, st>>>(n, batch, a, l, e, p, k, spill, S);
    ce = cudaGetLastError();
    const cudaError_t cf = cudaFreeAsync(spill, st);
    return ce != cudaSuccess ? ce : cf;
}

#define MODCHOL_CAT2(a, b, c, d) a##b##c##d
#define MODCHOL_CAT(a, b, c, d) MODCHOL_CAT2(a, b, c, d)
cudaError_t MODCHOL_CAT(launch_csm_a, MODCHOL_INST_ALGO, _s, MODCHOL_INST_STRICT)(
    int n, long long batch, const double* a, double* l, double* e, unsigned long long* p, int* k,
    int sms, cudaStream_t st)
{
    constexpr int A = MODCHOL_INST_ALGO;
    constexpr bool S = MODCHOL_INST_STRICT != 0;
    if (n <= 112)
        return launch_csm<A, S, 128, 3>(n, batch, a, l, e, p, k, sms, st);
    if (n <= 128)   // FAST: blocked DMMA trailing update from here on
        return launch_csm<A, S, 128, 3, true>(n, batch, a, l, e, p, k, sms, st);
    // (the FAST kernels carry an 8 KB panel: they spill from n = 233 instead of 237)
    const bool fits = csm_smem_bytes(n, 0, !S) <= (size_t)227 * 1024;
    if (n <= kCsmMaxN && fits)
        return launch_csm<A, S, 256, 1, true>(n, batch, a, l, e, p, k, sms, st);
    return launch_csm_spill<A, S, 256>(n, batch, a, l, e, p, k, sms, st);
}

}  // namespace modchol
