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

// compile-time n, everything in registers (n <= kTprMaxN).  Direct per-thread global accesses
// where one thread moves whole sectors per instruction (n = 2, 4, 8: rows of 32-byte multiples, e
// and p as 256-bit stores too), the tile
// staged through shared memory otherwise (measured: profiles/r02_thread_per_matrix.txt).
template <int kAlgo, int N, bool kSolve>
static cudaError_t launch_tpr_n(long long batch, const double* a, double* l, double* e,
                                unsigned long long* p, int* k, int sms, cudaStream_t st,
                                const double* rhs, double* x, int nrhs)
{
    constexpr bool kDirect = N == 2 || N == 4 || N == 8;
    long long blocks = (batch + kTpmThreads - 1) / kTpmThreads;
    if constexpr (kDirect) {
        auto kern = tpr_kernel<kAlgo, N, kSolve>;
        // no shared memory: the whole carve-out goes to L1, which serves the rest of each sector a
        // thread has touched (queried once per device)
        static int occ_dev[64];
        int dev = 0;
        cudaError_t ce = cudaGetDevice(&dev);
        if (ce != cudaSuccess) return ce;
        int occ = dev < 64 ? occ_dev[dev] : 0;
        if (occ == 0) {
            ce = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxL1);
            if (ce != cudaSuccess) return ce;
            ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kTpmThreads, 0);
            if (ce != cudaSuccess) return ce;
            if (occ < 1) occ = 1;
            if (dev < 64) occ_dev[dev] = occ;   // benign race: every writer stores the same value
        }
        const long long resident = (long long)sms * occ;
        if (blocks > resident) blocks = resident;
        kern<<<(unsigned)blocks, kTpmThreads, 0, st>>>(batch, a, l, e, p, k, rhs, x, nrhs);
    } else {
        auto kern = tpr_kernel_staged<kAlgo, N, kSolve>;
        constexpr size_t smem = (size_t)kTpmThreads * tpr_thread_doubles(N) * sizeof(double);
        int occ = 1;
        cudaError_t ce = cached_occupancy(kern, kTpmThreads, smem, smem, &occ);
        if (ce != cudaSuccess) return ce;
        const long long resident = (long long)sms * occ;
        if (blocks > resident) blocks = resident;
        kern<<<(unsigned)blocks, kTpmThreads, smem, st>>>(batch, a, l, e, p, k, rhs, x, nrhs);
    }
    return cudaGetLastError();
}
template <int kAlgo, bool kSolve>
static cudaError_t launch_tpr_algo(int n, long long batch, const double* a, double* l, double* e,
                                   unsigned long long* p, int* k, int sms, cudaStream_t st,
                                   const double* rhs, double* x, int nrhs)
{
    switch (n) {
    case 1: return launch_tpr_n<kAlgo, 1, kSolve>(batch, a, l, e, p, k, sms, st, rhs, x, nrhs);
    case 2: return launch_tpr_n<kAlgo, 2, kSolve>(batch, a, l, e, p, k, sms, st, rhs, x, nrhs);
    case 3: return launch_tpr_n<kAlgo, 3, kSolve>(batch, a, l, e, p, k, sms, st, rhs, x, nrhs);
    case 4: return launch_tpr_n<kAlgo, 4, kSolve>(batch, a, l, e, p, k, sms, st, rhs, x, nrhs);
    case 5: return launch_tpr_n<kAlgo, 5, kSolve>(batch, a, l, e, p, k, sms, st, rhs, x, nrhs);
    case 6: return launch_tpr_n<kAlgo, 6, kSolve>(batch, a, l, e, p, k, sms, st, rhs, x, nrhs);
    case 7: return launch_tpr_n<kAlgo, 7, kSolve>(batch, a, l, e, p, k, sms, st, rhs, x, nrhs);
    case 8: return launch_tpr_n<kAlgo, 8, kSolve>(batch, a, l, e, p, k, sms, st, rhs, x, nrhs);
    default: return cudaErrorNotSupported;
    }
}
static_assert(kTprMaxN == 8, "launch_tpr_algo lists the instantiations");

// 8 < n <= kTpsMaxN: compile-time n, the triangle in thread-private shared memory
template <int kAlgo, int N>
static cudaError_t launch_tps_n(long long batch, const double* a, double* l, double* e,
                                unsigned long long* p, int* k, int sms, cudaStream_t st)
{
    // staged (coalesced) tile copies up to n = 11, direct per-thread accesses above
    auto kern = [] {
        if constexpr (N >= 12)
            return tps_kernel<kAlgo, N>;
        else
            return tps_kernel_staged<kAlgo, N>;
    }();
    constexpr size_t smem = (size_t)kTpsThreads * tps_thread_doubles(N, kAlgo) * sizeof(double);
    int occ = 1;
    cudaError_t ce = cached_occupancy(kern, kTpsThreads, smem, smem, &occ);
    if (ce != cudaSuccess) return ce;
    long long blocks = (batch + kTpsThreads - 1) / kTpsThreads;
    const long long resident = (long long)sms * occ;
    if (blocks > resident) blocks = resident;
    kern<<<(unsigned)blocks, kTpsThreads, smem, st>>>(batch, a, l, e, p, k);
    return cudaGetLastError();
}
template <int kAlgo>
static cudaError_t launch_tps_algo(int n, long long batch, const double* a, double* l, double* e,
                                   unsigned long long* p, int* k, int sms, cudaStream_t st)
{
    switch (n) {
    case 9: return launch_tps_n<kAlgo, 9>(batch, a, l, e, p, k, sms, st);
    case 10: return launch_tps_n<kAlgo, 10>(batch, a, l, e, p, k, sms, st);
    case 11: return launch_tps_n<kAlgo, 11>(batch, a, l, e, p, k, sms, st);
    case 12: return launch_tps_n<kAlgo, 12>(batch, a, l, e, p, k, sms, st);
    case 13: return launch_tps_n<kAlgo, 13>(batch, a, l, e, p, k, sms, st);
    case 14: return launch_tps_n<kAlgo, 14>(batch, a, l, e, p, k, sms, st);
    case 15: return launch_tps_n<kAlgo, 15>(batch, a, l, e, p, k, sms, st);
    case 16: return launch_tps_n<kAlgo, 16>(batch, a, l, e, p, k, sms, st);
    default: return cudaErrorNotSupported;
    }
}
static_assert(kTpsMaxN == 16, "launch_tps_algo lists the instantiations");

static bool tpr_enabled()   // MODCHOL_TPR=0: the run-time-n kernels for every n <= 8 (A/B knob)
{
    const char* v = getenv("MODCHOL_TPR");
    return !(v && v[0] == '0');
}

cudaError_t launch_tpm_kernel(int algo, int n, long long batch, const double* a, double* l, double* e,
                              unsigned long long* p, int* k, int sms, cudaStream_t st)
{
    if (n > kTpmMaxN && n <= kTpsMaxN) {
        if (algo == kGMW81) return launch_tps_algo<kGMW81>(n, batch, a, l, e, p, k, sms, st);
        if (algo == kSE90) return launch_tps_algo<kSE90>(n, batch, a, l, e, p, k, sms, st);
        if (algo == kSE99) return launch_tps_algo<kSE99>(n, batch, a, l, e, p, k, sms, st);
        return cudaErrorNotSupported;
    }
    if (n < 1 || n > kTpmMaxN)
        return cudaErrorNotSupported;
    if (n <= kTprMaxN && tpr_enabled()) {
        if (algo == kGMW81) return launch_tpr_algo<kGMW81, false>(n, batch, a, l, e, p, k, sms, st, nullptr, nullptr, 0);
        if (algo == kSE90) return launch_tpr_algo<kSE90, false>(n, batch, a, l, e, p, k, sms, st, nullptr, nullptr, 0);
        if (algo == kSE99) return launch_tpr_algo<kSE99, false>(n, batch, a, l, e, p, k, sms, st, nullptr, nullptr, 0);
    }
    if (algo == kGMW81) return launch_tpm_algo<kGMW81>(n, batch, a, l, e, p, k, sms, st);
    if (algo == kSE90) return launch_tpm_algo<kSE90>(n, batch, a, l, e, p, k, sms, st);
    if (algo == kSE99) return launch_tpm_algo<kSE99>(n, batch, a, l, e, p, k, sms, st);
    return cudaErrorNotSupported;
}

// factorisation + solve in one kernel (n <= kTprMaxN): l, e, p, k may each be NULL
cudaError_t launch_tpm_solve_kernel(int algo, int n, long long batch, const double* a, double* l, double* e,
                                    unsigned long long* p, int* k, const double* rhs, double* x, int nrhs,
                                    int sms, cudaStream_t st)
{
    if (n < 1 || n > kTprMaxN)
        return cudaErrorNotSupported;
    if (algo == kGMW81) return launch_tpr_algo<kGMW81, true>(n, batch, a, l, e, p, k, sms, st, rhs, x, nrhs);
    if (algo == kSE90) return launch_tpr_algo<kSE90, true>(n, batch, a, l, e, p, k, sms, st, rhs, x, nrhs);
    if (algo == kSE99) return launch_tpr_algo<kSE99, true>(n, batch, a, l, e, p, k, sms, st, rhs, x, nrhs);
    return cudaErrorNotSupported;
}

}  // namespace modchol
