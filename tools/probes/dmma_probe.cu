// dmma_probe.cu -- throughput and latency of mma.sync.m8n8k4.f64 (DMMA) against DFMA on one SM.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 1024

template <int ILP, bool kDmma>
__global__ void __launch_bounds__(1024) k(double* out, long long* cyc, double seed)
{
    double a = seed + threadIdx.x * 1e-3, b = seed * 0.5;
    double c0[ILP], c1[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c0[i] = i; c1[i] = -i; }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (kDmma)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                             : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
            else
                c0[i] = __fma_rn(a, b, c0[i]);
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int ILP, bool kDmma>
void run(const char* name, int threads)
{
    double* out; long long* cyc;
    cudaMalloc(&out, 8 * 148 * threads); cudaMalloc(&cyc, 8 * 148);
    k<ILP, kDmma><<<148, threads>>>(out, cyc, 1.5);
    k<ILP, kDmma><<<148, threads>>>(out, cyc, 1.5);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, 8 * 148, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    double instr = (double)ITERS * ILP * (threads / 32);
    printf("%-6s ILP=%d thr=%4d: %8lld cycles, %.3f warp-instr/clk/SM, %.1f cycles per instr per warp (latency if ILP=1)\n",
           name, ILP, threads, mx, instr / mx, (double)mx / (ITERS * ILP));
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    run<1, true>("DMMA", 32); run<1, false>("DFMA", 32);
    run<8, true>("DMMA", 32); run<8, false>("DFMA", 32);
    run<8, true>("DMMA", 128); run<8, false>("DFMA", 128);
    run<8, true>("DMMA", 512); run<8, false>("DFMA", 512);
    run<8, true>("DMMA", 1024); run<8, false>("DFMA", 1024);
    return 0;
}
