// microbench.cu -- per-SM issue rates of the instructions the warp-per-matrix kernels are
// made of (FP64 pipe, shuffles, redux, broadcast LDS, double div/sqrt sequences).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
// Output: warp-instructions per clock per SM (cycles from clock64 inside the kernel).
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048
#define ILP 8

template <int OP>
__global__ void __launch_bounds__(1024) k(double* out, long long* cyc, double seed)
{
    __shared__ double sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = seed + threadIdx.x;
    __syncthreads();
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = seed + threadIdx.x * 1e-3 + i;
    double c = seed * 1.0000001, d = seed * 0.37;
    unsigned u[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) u[i] = threadIdx.x * 2654435761u + i;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == 0) x[i] = __fma_rn(x[i], c, d);
            if (OP == 1) x[i] = __dadd_rn(x[i], d);
            if (OP == 2) x[i] = __dmul_rn(x[i], c);
            if (OP == 3) x[i] = fmax(x[i], d + i);
            if (OP == 4) x[i] = __ddiv_rn(d, x[i]);
            if (OP == 5) x[i] = __dsqrt_rn(x[i]) + 2.0;
            if (OP == 6) x[i] = __shfl_xor_sync(0xffffffffu, x[i], 1 + (i & 15));
            if (OP == 7) u[i] = __reduce_max_sync(0xffffffffu, u[i]) + threadIdx.x;
            if (OP == 8) { double2 v = *reinterpret_cast<double2*>(&sm[(2 * i + (it & 1) * 16) & 62]); x[i] = __fma_rn(v.x, v.y, x[i]); }
            if (OP == 9) x[i] = (x[i] > d + i) ? c : x[i] + 1.0;   // DSETP + select
            if (OP == 10) x[i] = rsqrt(x[i]) + 2.0;
            if (OP == 11) { double v = sm[(i + it) & 63]; x[i] = __dadd_rn(v, x[i]); }
            if (OP == 12) u[i] = __shfl_xor_sync(0xffffffffu, u[i], 1 + (i & 15)) + 1;
            if (OP == 13) u[i] = __ballot_sync(0xffffffffu, u[i] & 1) + u[i];
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i] + u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int threads, int blocks_per_sm, int sms)
{
    double* out; long long* cyc;
    int blocks = sms * blocks_per_sm;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaMalloc(&cyc, sizeof(long long) * blocks);
    k<OP><<<blocks, threads>>>(out, cyc, 1.5);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<blocks, threads>>>(out, cyc, 1.5);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long* h = new long long[blocks];
    cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
    double warp_instr_per_sm = (double)ITERS * ILP * (threads / 32) * blocks_per_sm;
    printf("%-28s thr=%4d blk/SM=%d  cycles=%9.0f  warp-instr/clk/SM=%6.3f  (%.3f ms, %.2f G warp-instr/s chip)\n",
           name, threads, blocks_per_sm, avg, warp_instr_per_sm / avg, ms,
           warp_instr_per_sm * sms / (ms * 1e-3) / 1e9);
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("  CUDA error %s\n", cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc); delete[] h;
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("device %s, %d SMs, clock %d kHz\n", p.name, sms, p.clockRate);
    for (int thr : {128, 512, 1024}) {
        run<0>("DFMA", thr, 1, sms);
        run<1>("DADD", thr, 1, sms);
        run<2>("DMUL", thr, 1, sms);
        run<3>("fmax(double)", thr, 1, sms);
        run<9>("DSETP+sel", thr, 1, sms);
        run<4>("ddiv_rn (sequence)", thr, 1, sms);
        run<5>("dsqrt_rn (sequence)+add", thr, 1, sms);
        run<10>("rsqrt(double)+add", thr, 1, sms);
        run<6>("SHFL.64 (2x32)", thr, 1, sms);
        run<12>("SHFL.32", thr, 1, sms);
        run<7>("REDUX.MAX u32", thr, 1, sms);
        run<13>("VOTE.ballot", thr, 1, sms);
        run<8>("LDS.128 bcast + DFMA", thr, 1, sms);
        run<11>("LDS.64 bcast + DADD", thr, 1, sms);
    }
    return 0;
}
