// Microbenchmark: FP64 FMA throughput vs independent chains per thread and warps per SM sub-partition.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dfma_latency dfma_latency.cu && ./dfma_latency
#include <cstdio>
#include <cuda_runtime.h>
template <int C>
__global__ void k(double *out, int iters, double a, double b)
{
    double r[C];
#pragma unroll
    for (int i = 0; i < C; ++i) r[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < C; ++i) r[i] = fma(r[i], a, b);
    double s = 0;
#pragma unroll
    for (int i = 0; i < C; ++i) s += r[i];
    if (s == 12345.678) out[threadIdx.x] = s;
}
template <int C>
void run(int warps_per_smsp, int sms)
{
    double *out; cudaMalloc(&out, 1 << 20);
    int threads = 32 * 4 * warps_per_smsp; // one block per SM, 4 SMSPs
    int blocks = sms, iters = 1 << 15;
    if (threads > 1024) { blocks = sms * (threads / 1024); threads = 1024; }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<C><<<blocks, threads>>>(out, 1000, 0.999, 1e-9);
    cudaEventRecord(e0);
    k<C><<<blocks, threads>>>(out, iters, 0.999, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double inst = (double)C * iters * blocks * threads / 32;           // warp-instructions
    double per_smsp_per_clk = inst / (sms * 4) / (ms * 1e-3 * 1.965e9); // at 1965 MHz
    printf("chains %d warps/SMSP %2d : %.3f warp-DFMA/clk/SMSP  (%.1f TFLOP/s)\n", C, warps_per_smsp, per_smsp_per_clk,
           inst * 64 / (ms * 1e-3) / 1e12);
    cudaFree(out);
}
int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int w : {1, 2, 4, 6, 8, 16}) { run<1>(w, sms); run<2>(w, sms); run<4>(w, sms); run<8>(w, sms); }
    return 0;
}
