// warp_inst.cu -- instantiates the warp-per-matrix kernels for ONE (algorithm, mode) pair,
// selected with -DMODCHOL_INST_ALGO={0,1,2} -DMODCHOL_INST_STRICT={0,1}.  build.py compiles
// this file once per pair, in parallel.
#include "warp_kernels.cuh"
#include "warp_gmw.cuh"
#include "warp_se.cuh"

#ifndef MODCHOL_MINB32
#define MODCHOL_MINB32 4   /* resident 4-warp blocks per SM asked of ptxas for NMAX >= 24 */
#endif
#ifndef MODCHOL_INST_ALGO
#error "compile with -DMODCHOL_INST_ALGO=.. -DMODCHOL_INST_STRICT=.."
#endif

namespace modchol {

template <int kAlgo, int NMAX, bool kStrict>
static cudaError_t launch_one(int n, long long batch, const double* a, double* l, double* e,
                              unsigned long long* p, int* k, int sms, cudaStream_t st)
{
    // resident blocks per SM the register allocation is asked to allow (4 warps per block)
    constexpr int kMinBlocks = NMAX >= 24 ? MODCHOL_MINB32 : (NMAX >= 16 ? 5 : 6);
    constexpr int M = 32 / se_group_lanes<NMAX>();   // matrices per warp
    const long long slots = (batch + M - 1) / M;
    long long blocks = (slots + kWarpsPerBlock - 1) / kWarpsPerBlock;
    int occ = 1;
    cudaError_t ce;
    if constexpr (kAlgo == kGMW81) {
        auto kern = gmw_warp_kernel<NMAX, kStrict, kWarpsPerBlock, kMinBlocks>;
        const size_t smem = (size_t)kWarpsPerBlock * gmw_warp_smem_doubles<NMAX>() * sizeof(double);
        ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kWarpsPerBlock * 32, smem);
        if (ce != cudaSuccess) return ce;
        const long long resident = (long long)sms * (occ < 1 ? 1 : occ);
        if (blocks > resident) blocks = resident;  // persistent: every warp strides over the batch
        kern<<<(unsigned)blocks, kWarpsPerBlock * 32, smem, st>>>(n, batch, a, l, e, p, k);
    } else {
        auto kern = se_warp_kernel<kAlgo, NMAX, kStrict, kWarpsPerBlock, kMinBlocks>;
        const size_t smem = (size_t)kWarpsPerBlock * se_warp_smem_doubles<NMAX, kStrict>() * sizeof(double);
        ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (ce != cudaSuccess) return ce;
        ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kWarpsPerBlock * 32, smem);
        if (ce != cudaSuccess) return ce;
        const long long resident = (long long)sms * (occ < 1 ? 1 : occ);
        if (blocks > resident) blocks = resident;
        const int vec_ok = ((n & 1) == 0) && ((reinterpret_cast<uintptr_t>(l) & 15) == 0);
        kern<<<(unsigned)blocks, kWarpsPerBlock * 32, smem, st>>>(n, vec_ok, batch, a, l, e, p, k);
    }
    return cudaGetLastError();
}

#define MODCHOL_CAT2(a, b, c, d) a##b##c##d
#define MODCHOL_CAT(a, b, c, d) MODCHOL_CAT2(a, b, c, d)
cudaError_t MODCHOL_CAT(launch_warp_a, MODCHOL_INST_ALGO, _s, MODCHOL_INST_STRICT)(
    int n, long long batch, const double* a, double* l, double* e, unsigned long long* p, int* k,
    int sms, cudaStream_t st)
{
    constexpr int A = MODCHOL_INST_ALGO;
    constexpr bool S = MODCHOL_INST_STRICT != 0;
    switch (warp_nmax(n)) {
    case 8: return launch_one<A, 8, S>(n, batch, a, l, e, p, k, sms, st);
    case 16: return launch_one<A, 16, S>(n, batch, a, l, e, p, k, sms, st);
    case 24: return launch_one<A, 24, S>(n, batch, a, l, e, p, k, sms, st);
    default: return launch_one<A, 32, S>(n, batch, a, l, e, p, k, sms, st);
    }
}

}  // namespace modchol
