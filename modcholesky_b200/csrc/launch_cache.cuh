// launch_cache.cuh -- per-kernel launch configuration, queried once.
// cudaFuncSetAttribute (shared-memory opt-in, carve-out) and the occupancy query cost several
// microseconds each; the single-matrix trait path (batch = 1) would pay them on every call.  The
// attributes are set once per (kernel, device) for the largest dynamic shared-memory size the
// kernel may be launched with, and the occupancy is cached per (kernel, device, shared bytes).
#pragma once
#include <cuda_runtime.h>

#include <mutex>
#include <vector>

namespace modchol {

struct LaunchCacheEntry {
    const void* fn;
    int dev;
    size_t smem;   // (size_t)-1: the attribute record of (fn, dev)
    int occ;
};

inline std::mutex& launch_cache_mutex()
{
    static std::mutex mu;
    return mu;
}
inline std::vector<LaunchCacheEntry>& launch_cache_entries()
{
    static std::vector<LaunchCacheEntry> v;
    return v;
}

// Resident blocks per SM of `kern` launched with `threads` threads and `smem` dynamic shared
// bytes on the current device; `smem_max` = the largest `smem` this kernel is ever launched with
// (the opt-in attribute is set to it once).
template <typename Kern>
cudaError_t cached_occupancy(Kern kern, int threads, size_t smem, size_t smem_max, int* occ)
{
    const void* fn = reinterpret_cast<const void*>(kern);
    int dev = 0;
    cudaError_t ce = cudaGetDevice(&dev);
    if (ce != cudaSuccess) return ce;
    std::lock_guard<std::mutex> lock(launch_cache_mutex());
    auto& v = launch_cache_entries();
    bool have_attr = false;
    for (const auto& en : v) {
        if (en.fn != fn || en.dev != dev) continue;
        if (en.smem == (size_t)-1) have_attr = true;
        if (en.smem == smem) {
            *occ = en.occ;
            return cudaSuccess;
        }
    }
    if (!have_attr) {
        ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
        if (ce != cudaSuccess) return ce;
        ce = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  (int)cudaSharedmemCarveoutMaxShared);
        if (ce != cudaSuccess) return ce;
        v.push_back({fn, dev, (size_t)-1, 0});
    }
    int o = 1;
    ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, threads, smem);
    if (ce != cudaSuccess) return ce;
    if (o < 1) o = 1;
    v.push_back({fn, dev, smem, o});
    *occ = o;
    return cudaSuccess;
}

}  // namespace modchol
