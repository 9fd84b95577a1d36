// tmem_probe.cu -- does TMEM work as a warp-private scratchpad?  Each CTA of 4 warps allocates
// 64 columns; warp w owns lanes [32w, 32w+32); every lane stores 32 doubles (2 columns each) at a
// run-time column, then reads its row back in chunks of 16 columns.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(128, 8) probe(int reps, unsigned long long* bad, double* out)
{
    __shared__ uint32_t tbase;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;"
                     :: "r"((uint32_t)__cvta_generic_to_shared(&tbase)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tw = tbase + ((uint32_t)(warp * 32) << 16);
    unsigned long long nbad = 0;
    for (int r = 0; r < reps; ++r) {
        for (int t = 0; t < 32; ++t) {
            const double v = (double)(blockIdx.x * 1000003 + warp * 10007 + lane * 101 + t + r) * 0.5;
            const uint32_t lo = (uint32_t)__double2loint(v), hi = (uint32_t)__double2hiint(v);
            asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};"
                         :: "r"(tw + 2u * t), "r"(lo), "r"(hi) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        for (int c = 0; c < 4; ++c) {
            uint32_t w[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                         "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                         : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]),
                           "=r"(w[7]), "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]),
                           "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
                         : "r"(tw + 16u * c) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int u = 0; u < 8; ++u) {
                const int t = 8 * c + u;
                const double v = __hiloint2double((int)w[2 * u + 1], (int)w[2 * u]);
                const double e = (double)(blockIdx.x * 1000003 + warp * 10007 + lane * 101 + t + r) * 0.5;
                if (v != e) ++nbad;
                if (r == 0 && blockIdx.x == 0 && warp == 1 && lane == 3) out[t] = v;
            }
        }
    }
    if (nbad) atomicAdd(bad, nbad);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" :: "r"(tbase) : "memory");
}

int main()
{
    unsigned long long* bad;
    double* out;
    cudaMalloc(&bad, 8);
    cudaMalloc(&out, 32 * 8);
    cudaMemset(bad, 0, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int it = 0; it < 3; ++it) {
        cudaEventRecord(e0);
        probe<<<148 * 8 * 4, 128>>>(200, bad, out);
        cudaEventRecord(e1);
        cudaError_t ce = cudaDeviceSynchronize();
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        unsigned long long h = 0;
        cudaMemcpy(&h, bad, 8, cudaMemcpyDeviceToHost);
        double ho[32];
        cudaMemcpy(ho, out, 256, cudaMemcpyDeviceToHost);
        printf("run %d: %s, mismatches %llu, %.3f ms, out[5]=%g (expect %g)\n", it, cudaGetErrorString(ce), h, ms,
               ho[5], (double)(1 * 10007 + 3 * 101 + 5) * 0.5);
    }
    return 0;
}
