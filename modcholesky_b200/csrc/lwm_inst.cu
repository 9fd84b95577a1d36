// lwm_inst.cu -- instantiates the warp-per-matrix large-n kernels (lwm_kernels.cuh, FAST mode,
// 64 < n <= 256) for ONE algorithm: -DMODCHOL_INST_ALGO={0,1,2}.
#include <cstdlib>

#include "launch_cache.cuh"
#include "lwm_kernels.cuh"
#include "warp_kernels.cuh"

#ifndef MODCHOL_INST_ALGO
#error "compile with -DMODCHOL_INST_ALGO=.."
#endif

namespace modchol {

template <int kAlgo, int R, int KB>
static cudaError_t launch_lwm(int n, long long batch, const double* a, double* l, double* e,
                              unsigned long long* p, int* k, int sms, cudaStream_t st)
{
    auto kern = lwm_kernel<kAlgo, R, KB>;
    const size_t smem = lwm_smem_bytes(n, KB);
    const size_t smem_max = lwm_smem_bytes(32 * R, KB);
    int occ = 1;
    cudaError_t ce = cached_occupancy(kern, 32, smem, smem_max, &occ);
    if (ce != cudaSuccess) return ce;
    // matrices in flight per SM: bounded by shared memory (occ) and by the L2 -- the working
    // triangles of all resident matrices (4 n^2 bytes each) should stay inside ~100 MB
    int want = (int)((100ull << 20) / ((size_t)4 * n * n * (size_t)sms));
    if (want < 1) want = 1;
    if (const char* f = getenv("MODCHOL_LWM_OCC"))   // tuning knob
        want = atoi(f) > 0 ? atoi(f) : want;
    if (occ > want) occ = want;
    long long blocks = batch;
    const long long resident = (long long)sms * occ;
    if (blocks > resident) blocks = resident;   // persistent: every warp strides over the batch
    kern<<<(unsigned)blocks, 32, smem, st>>>(n, batch, a, l, e, p, k);
    return cudaGetLastError();
}

#define MODCHOL_CAT2(a, b) a##b
#define MODCHOL_CAT(a, b) MODCHOL_CAT2(a, b)
cudaError_t MODCHOL_CAT(launch_lwm_a, MODCHOL_INST_ALGO)(int n, long long batch, const double* a, double* l,
                                                          double* e, unsigned long long* p, int* k, int sms,
                                                          cudaStream_t st)
{
    constexpr int A = MODCHOL_INST_ALGO;
    const char* kbe = getenv("MODCHOL_LWM_KB");      // tuning knob: panel width
    const int kb = kbe ? atoi(kbe) : 16;
    if (n <= 128)
        return kb == 8 ? launch_lwm<A, 4, 8>(n, batch, a, l, e, p, k, sms, st)
                       : launch_lwm<A, 4, 16>(n, batch, a, l, e, p, k, sms, st);
    return kb == 8    ? launch_lwm<A, 8, 8>(n, batch, a, l, e, p, k, sms, st)
           : kb == 32 ? launch_lwm<A, 8, 32>(n, batch, a, l, e, p, k, sms, st)
                      : launch_lwm<A, 8, 16>(n, batch, a, l, e, p, k, sms, st);
}

}  // namespace modchol
