// warp_kernels.cuh -- warp-per-matrix, register-resident kernels for n <= 32 (stub: filled
// in by the next commit).
#pragma once
#include "common.cuh"
namespace modchol {
inline bool warp_kernel_available(int, int, int) { return false; }
inline cudaError_t launch_warp_kernel(int, int, int, long long, const double*, double*, double*,
                                      unsigned long long*, int*, int, cudaStream_t)
{
    return cudaErrorNotSupported;
}
}  // namespace modchol
