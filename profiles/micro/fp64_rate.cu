// fp64 issue rate and dependent latency on B200: nvcc -gencode arch=compute_100a,code=sm_100a -O3 fp64_rate.cu -o fp64_rate
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k(double* out, int iters, double a, double b)
{
    double x[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) x[c] = threadIdx.x * 1e-3 + c;
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int c = 0; c < CH; ++c) x[c] = fma(x[c], a, b);
    double s = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
void run(int warps_per_sm)
{
    int sms = 148, threads = 32 * warps_per_sm, iters = 20000;
    double* d; cudaMalloc(&d, sizeof(double) * sms * threads);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<CH><<<sms, threads>>>(d, 100, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    k<CH><<<sms, threads>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma_total = (double)sms * threads * iters * CH;
    double cyc = ms * 1e-3 * 1.965e9;
    printf("chains %2d warps/SM %2d: %.3f ms  %.2f TFLOP/s  %.1f DFMA/clk/SM  cycles per dependent step %.1f\n", CH, warps_per_sm, ms,
        2 * fma_total / ms / 1e9, fma_total / sms / cyc, cyc / iters);
    cudaFree(d);
}
int main()
{
    run<1>(1); run<1>(4); run<2>(4); run<4>(4); run<8>(4); run<1>(8); run<1>(16); run<4>(16); run<8>(16); run<8>(32); run<4>(64);
    return 0;
}
