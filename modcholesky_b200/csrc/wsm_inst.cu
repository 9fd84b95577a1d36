// wsm_inst.cu -- instantiates the shared-memory-resident warp kernels (wsm_se.cuh) for ONE
// (algorithm, mode) pair: -DMODCHOL_INST_ALGO={1,2} -DMODCHOL_INST_STRICT={0,1}.
#include <cstdio>
#include <cstdlib>

#include "launch_cache.cuh"
#include "warp_kernels.cuh"
#include "wip2_kernels.cuh"
#include "wip_kernels.cuh"
#include "wipp_kernels.cuh"
#include "wsm_gmw.cuh"
#include "wsm_se.cuh"

#ifndef MODCHOL_WSM_MINB
#define MODCHOL_WSM_MINB 8   /* resident 4-warp blocks per SM asked of ptxas for n <= 32 */
#endif
#ifndef MODCHOL_WSM_MINB_PACKED
#define MODCHOL_WSM_MINB_PACKED 6   /* same, two matrices per warp with two rows per lane */
#endif
#ifndef MODCHOL_INST_ALGO
#error "compile with -DMODCHOL_INST_ALGO=.. -DMODCHOL_INST_STRICT=.."
#endif

namespace modchol {

template <int kAlgo, int G, int R, bool kStrict>
static cudaError_t launch_wsm(int n, long long batch, const double* a, double* l, double* e,
                              unsigned long long* p, int* k, const double* rhs, double* x, int nrhs,
                              int sms, cudaStream_t st)
{
    constexpr int kWarps = 4;
    constexpr int kPack = 32 / G;   // matrices per warp
    // resident 4-warp blocks asked of ptxas: one matrix per warp and one row per lane fits 64
    // registers (8 blocks); two rows per lane wants ~80 (6 blocks packed, 3 at n <= 64)
    constexpr int kMinBlocks = R == 1 ? MODCHOL_WSM_MINB : (G == 16 ? MODCHOL_WSM_MINB_PACKED : 3);
    auto kern = [] {
        if constexpr (kAlgo == kGMW81)
            return wsm_gmw_kernel<G, R, kStrict, kWarps, kMinBlocks>;
        else
            return wsm_se_kernel<kAlgo, G, R, kStrict, kWarps, kMinBlocks>;
    }();
    const int wstride = wsm_warp_doubles(n);
    const size_t smem = (size_t)kWarps * kPack * wstride * sizeof(double);
    // the matrices live in shared memory: largest carve-out; attributes and occupancy are cached
    constexpr int kNmax = G * R;   // largest n this instantiation serves
    const size_t smem_max = (size_t)kWarps * kPack * wsm_warp_doubles(kNmax) * sizeof(double);
    int occ = 1;
    cudaError_t ce = cached_occupancy(kern, kWarps * 32, smem, smem_max, &occ);
    if (ce != cudaSuccess) return ce;
    long long blocks = (batch + kWarps * kPack - 1) / (kWarps * kPack);
    const long long resident = (long long)sms * occ;
    if (blocks > resident) blocks = resident;   // persistent: every lane group strides over the batch
    kern<<<(unsigned)blocks, kWarps * 32, smem, st>>>(n, wstride, batch, a, l, e, p, k, rhs, x, nrhs);
    return cudaGetLastError();
}

// FAST mode, 16 < n <= 32: the implicit-pivoting kernels (wip_kernels.cuh), kSolve: with the fused solve.
// The launch attributes and the occupancy are queried once per instantiation and device.
#ifndef MODCHOL_WIP_REGC
#define MODCHOL_WIP_REGC 0   /* C fragments in registers (1) or in the shared-memory triangle (0) */
#endif
#ifndef MODCHOL_WIP_MINB
#define MODCHOL_WIP_MINB (MODCHOL_WIP_REGC ? 5 : 8)
#endif
template <int kAlgo, int NB, int kN, bool kSolve = false>
static cudaError_t launch_wip(int n, long long batch, const double* a, double* l, double* e,
                              unsigned long long* p, int* k, int sms, cudaStream_t st,
                              const double* rhs = nullptr, double* x = nullptr, int nrhs = 0)
{
    auto kern = wip_kernel<kAlgo, NB, kN, MODCHOL_WIP_REGC != 0, MODCHOL_WIP_MINB, kSolve>;
    constexpr size_t smem = (size_t)kWipWarps * kWipDoubles * sizeof(double);
    // The occupancy query answers 1 for any kernel that allocates tensor memory (it assumes all
    // 512 columns); this kernel takes kWipTmemCols per block, so the real limit is the register
    // file / shared memory one, capped by 512 / kWipTmemCols.  Computed once per device.
    static int occ_dev[64];   // resident blocks per SM, 0 = not yet computed on that device
    int dev = 0;
    cudaError_t ce = cudaGetDevice(&dev);
    if (ce != cudaSuccess) return ce;
    int occ = dev < 64 ? occ_dev[dev] : 0;
    if (occ == 0) {
        int ignored = 0;
        ce = cached_occupancy(kern, kWipWarps * 32, smem, smem, &ignored);   // sets the attributes
        if (ce != cudaSuccess) return ce;
        cudaFuncAttributes fa;
        ce = cudaFuncGetAttributes(&fa, kern);
        if (ce != cudaSuccess) return ce;
        const int by_regs = 65536 / (fa.numRegs * kWipWarps * 32);
        const int by_smem = (int)((227 * 1024) / (smem + fa.sharedSizeBytes + 1024));
        occ = by_regs < by_smem ? by_regs : by_smem;
        if (occ < 1) occ = 1;
        if (occ > 512 / kWipTmemCols) occ = 512 / kWipTmemCols;
        if (dev < 64) occ_dev[dev] = occ;   // benign race: every writer stores the same value
    }
    long long blocks = (batch + kWipWarps - 1) / kWipWarps;
    const long long resident = (long long)sms * occ;
    if (blocks > resident) blocks = resident;   // persistent: every warp strides over the batch
    kern<<<(unsigned)blocks, kWipWarps * 32, smem, st>>>(n, batch, a, l, e, p, k, rhs, x, nrhs);
    return cudaGetLastError();
}

// FAST mode, 32 < n <= 64: the two-rows-per-lane implicit-pivoting kernels
// (wip2_kernels.cuh); 256 tensor-memory columns and 84 KB of shared memory per block.
template <int kAlgo, int NB, int kN, bool kSolve = false>
static cudaError_t launch_wip2(int n, long long batch, const double* a, double* l, double* e,
                               unsigned long long* p, int* k, int sms, cudaStream_t st,
                               const double* rhs = nullptr, double* x = nullptr, int nrhs = 0)
{
    auto kern = wip2_kernel<kAlgo, NB, kN, 2, kSolve>;
    constexpr size_t smem = (size_t)kWipWarps * kWip2Doubles * sizeof(double);
    static int occ_dev[64];
    int dev = 0;
    cudaError_t ce = cudaGetDevice(&dev);
    if (ce != cudaSuccess) return ce;
    int occ = dev < 64 ? occ_dev[dev] : 0;
    if (occ == 0) {
        int ignored = 0;
        ce = cached_occupancy(kern, kWipWarps * 32, smem, smem, &ignored);   // sets the attributes
        if (ce != cudaSuccess) return ce;
        cudaFuncAttributes fa;
        ce = cudaFuncGetAttributes(&fa, kern);
        if (ce != cudaSuccess) return ce;
        const int by_regs = 65536 / (fa.numRegs * kWipWarps * 32);
        const int by_smem = (int)((227 * 1024) / (smem + fa.sharedSizeBytes + 1024));
        occ = by_regs < by_smem ? by_regs : by_smem;
        if (occ < 1) occ = 1;
        if (occ > 512 / kWip2TmemCols) occ = 512 / kWip2TmemCols;
        if (dev < 64) occ_dev[dev] = occ;
    }
    long long blocks = (batch + kWipWarps - 1) / kWipWarps;
    const long long resident = (long long)sms * occ;
    if (blocks > resident) blocks = resident;   // persistent: every warp strides over the batch
    kern<<<(unsigned)blocks, kWipWarps * 32, smem, st>>>(n, batch, a, l, e, p, k, rhs, x, nrhs);
    return cudaGetLastError();
}

// FAST mode, 4 < n <= 16: several matrices per warp (wipp_kernels.cuh)
template <int kAlgo, int G, int kN, bool kSolve = false>
static cudaError_t launch_wipp(int n, long long batch, const double* a, double* l, double* e,
                               unsigned long long* p, int* k, int sms, cudaStream_t st,
                               const double* rhs = nullptr, double* x = nullptr, int nrhs = 0)
{
    auto kern = wipp_kernel<kAlgo, G, kN, 8, kSolve>;
    constexpr int kGroups = 32 / G;
    constexpr size_t smem = (size_t)kWipWarps * kGroups * wpp_mat_doubles(G) * sizeof(double);
    static int occ_dev[64];
    int dev = 0;
    cudaError_t ce = cudaGetDevice(&dev);
    if (ce != cudaSuccess) return ce;
    int occ = dev < 64 ? occ_dev[dev] : 0;
    if (occ == 0) {
        int ignored = 0;
        ce = cached_occupancy(kern, kWipWarps * 32, smem, smem, &ignored);   // sets the attributes
        if (ce != cudaSuccess) return ce;
        cudaFuncAttributes fa;
        ce = cudaFuncGetAttributes(&fa, kern);
        if (ce != cudaSuccess) return ce;
        const int by_regs = 65536 / (fa.numRegs * kWipWarps * 32);
        const int by_smem = (int)((227 * 1024) / (smem + fa.sharedSizeBytes + 1024));
        occ = by_regs < by_smem ? by_regs : by_smem;
        if (occ < 1) occ = 1;
        if (occ > 16) occ = 16;   // 64 warps per SM; 32 tensor-memory columns per block
        if (dev < 64) occ_dev[dev] = occ;
    }
    long long blocks = (batch + kWipWarps * kGroups - 1) / (kWipWarps * kGroups);
    const long long resident = (long long)sms * occ;
    if (blocks > resident) blocks = resident;   // persistent: every warp strides over the batch
    kern<<<(unsigned)blocks, kWipWarps * 32, smem, st>>>(n, batch, a, l, e, p, k, rhs, x, nrhs);
    return cudaGetLastError();
}

static bool wipp_enabled()
{
    const char* v = getenv("MODCHOL_WIPP");
    return !(v && v[0] == '0');
}

static bool wip2_enabled()
{
    const char* v = getenv("MODCHOL_WIP2");
    return !(v && v[0] == '0');
}

static bool wip_enabled()
{
    const char* v = getenv("MODCHOL_WIP");
    return !(v && v[0] == '0');
}

// 16 < n <= 32 can also run two matrices per warp (16 lanes x 2 rows each): 11 % fewer
// instructions per matrix and +4 % (FAST) / +20 % (STRICT) when every matrix takes the same
// path, but -27 % when the two matrices of a warp leave phase 1 at different columns and the
// halves serialise (profiles/r01_notes.md) -- opt-in with MODCHOL_WSM_PACK=1, read per call.
static bool wsm_pack_pref()
{
    const char* v = getenv("MODCHOL_WSM_PACK");
    return v && v[0] == '1';
}

#define MODCHOL_CAT2(a, b, c, d) a##b##c##d
#define MODCHOL_CAT(a, b, c, d) MODCHOL_CAT2(a, b, c, d)
cudaError_t MODCHOL_CAT(launch_wsm_a, MODCHOL_INST_ALGO, _s, MODCHOL_INST_STRICT)(
    int n, long long batch, const double* a, double* l, double* e, unsigned long long* p, int* k,
    const double* rhs, double* x, int nrhs, int sms, cudaStream_t st)
{
    constexpr int A = MODCHOL_INST_ALGO;
    constexpr bool S = MODCHOL_INST_STRICT != 0;
    // FAST, 4 < n <= 64: the implicit-pivoting kernels, with the fused solve when x is given
    if constexpr (!S) {
        const bool sv = x != nullptr;
#define MODCHOL_WIPX(fn, ...)                                                                      \
    (sv ? fn<A, __VA_ARGS__, true>(n, batch, a, l, e, p, k, sms, st, rhs, x, nrhs)                 \
        : fn<A, __VA_ARGS__, false>(n, batch, a, l, e, p, k, sms, st))
        if (n > 8 && n <= 16 && wipp_enabled())
            return n < 16 ? MODCHOL_WIPX(launch_wipp, 16, 0) : MODCHOL_WIPX(launch_wipp, 16, 16);
        if (n > 4 && n <= 8 && wipp_enabled())
            return n < 8 ? MODCHOL_WIPX(launch_wipp, 8, 0) : MODCHOL_WIPX(launch_wipp, 8, 8);
        if (n > 16 && n <= 32 && wip_enabled() && (!sv || MODCHOL_WIP_REGC == 0))
            return n <= 24 ? MODCHOL_WIPX(launch_wip, 3, 0)
                   : n < 32 ? MODCHOL_WIPX(launch_wip, 4, 0) : MODCHOL_WIPX(launch_wip, 4, 32);
        if (n > 32 && n <= 64 && wip2_enabled())
            return n <= 48 ? MODCHOL_WIPX(launch_wip2, 6, 0)
                   : n < 64 ? MODCHOL_WIPX(launch_wip2, 8, 0) : MODCHOL_WIPX(launch_wip2, 8, 64);
#undef MODCHOL_WIPX
    }
    if (n <= 4)
        return launch_wsm<A, 4, 1, S>(n, batch, a, l, e, p, k, rhs, x, nrhs, sms, st);
    if (n <= 8)
        return launch_wsm<A, 8, 1, S>(n, batch, a, l, e, p, k, rhs, x, nrhs, sms, st);
    if (n <= 16)
        return launch_wsm<A, 16, 1, S>(n, batch, a, l, e, p, k, rhs, x, nrhs, sms, st);
    if (n <= 32)
        return wsm_pack_pref() ? launch_wsm<A, 16, 2, S>(n, batch, a, l, e, p, k, rhs, x, nrhs, sms, st)
                               : launch_wsm<A, 32, 1, S>(n, batch, a, l, e, p, k, rhs, x, nrhs, sms, st);
    return launch_wsm<A, 32, 2, S>(n, batch, a, l, e, p, k, rhs, x, nrhs, sms, st);
}

}  // namespace modchol
