// throughput of DMUL / DADD / DFMA on B200 (independent chains, 16 warps/SM)
#include <cstdio>
#include <cuda_runtime.h>
template <int OP, int CH>
__global__ void k(double* out, int iters, double a, double b)
{
    double x[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) x[c] = 1.0 + threadIdx.x * 1e-9 + c * 1e-7;
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            if (OP == 0) x[c] = fma(x[c], a, b);
            else if (OP == 1) x[c] = x[c] * a;
            else if (OP == 2) x[c] = x[c] + b;
            else x[c] = __fma_rn(x[c], a, -0.0);
        }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP, int CH>
void run(const char* name, int warps)
{
    int sms = 148, threads = 32 * warps, iters = 20000;
    double* d; cudaMalloc(&d, sizeof(double) * sms * threads);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP, CH><<<sms, threads>>>(d, 100, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    k<OP, CH><<<sms, threads>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)sms * threads * iters * CH, cyc = ms * 1e-3 * 1.965e9;
    printf("%-10s chains %d warps/SM %2d: %.3f ms  %.1f ops/clk/SM  %.1f cycles per dependent step\n", name, CH, warps, ms, ops / sms / cyc, cyc / iters);
    cudaFree(d);
}
int main()
{
    run<0, 4>("DFMA", 16); run<1, 4>("DMUL", 16); run<2, 4>("DADD", 16); run<3, 4>("fma(x,a,-0)", 16);
    run<0, 1>("DFMA", 1); run<1, 1>("DMUL", 1); run<2, 1>("DADD", 1);
    return 0;
}
