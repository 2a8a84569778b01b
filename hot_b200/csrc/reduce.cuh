// Deterministic two-stage sum reductions over DOF / particle arrays (K simultaneous sums).
//
// The reference reduces with tbb::parallel_reduce (grain 256, ImplicitSolver.h:158-171,192-197,265-273) or serial
// Eigen sums (LinearSolver.h:34-37); the summation tree there depends on TBB's splitting.  Here the tree is fixed
// (grid-stride per thread -> warp shuffle -> CTA -> one final CTA over the CTA partials in index order), so results
// are bit-reproducible from run to run, which the solver convergence tests rely on.
#pragma once
#include "sim.h"

namespace hot {

constexpr int RED_BLOCKS = 148 * 4; // 4 CTAs per SM
constexpr int RED_THREADS = 256;

template <int K>
__device__ __forceinline__ void block_sum(double (&acc)[K], double* out /* K values, written by thread 0 */)
{
    __shared__ double sh[K][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) sh[k][warp] = v;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double v = lane < nw ? sh[k][lane] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
            if (lane == 0) out[k] = v;
        }
    }
    __syncthreads();
}

template <int K, class F>
__global__ void __launch_bounds__(RED_THREADS) k_reduce_partial(long n, F f, double* __restrict__ partial)
{
    double acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.0;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) f(i, acc);
    __shared__ double res[K];
    block_sum<K>(acc, res);
    if (threadIdx.x == 0)
#pragma unroll
        for (int k = 0; k < K; ++k) partial[(size_t)k * gridDim.x + blockIdx.x] = res[k];
}

template <int K>
__global__ void __launch_bounds__(RED_THREADS) k_reduce_final(int nb, const double* __restrict__ partial, double* __restrict__ out)
{
    double acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        acc[k] = 0.0;
        for (int i = threadIdx.x; i < nb; i += blockDim.x) acc[k] += partial[(size_t)k * nb + i];
    }
    __shared__ double res[K];
    block_sum<K>(acc, res);
    if (threadIdx.x == 0)
#pragma unroll
        for (int k = 0; k < K; ++k) out[k] = res[k];
}

// K sums of f over [0, n) into dev_out[0..K) (device, s->red_out when null); optional synchronous host copy.
template <int K, class F>
int reduce_to(Sim* s, long n, F f, double* dev_out, double* host_out)
{
    HOT_CUDA(s->red_partial.reserve((size_t)8 * RED_BLOCKS));
    HOT_CUDA(s->red_out.reserve(64));
    if (!s->h_red) HOT_CUDA(cudaMallocHost((void**)&s->h_red, 64 * sizeof(double)));
    if (!dev_out) dev_out = s->red_out.p;
    int nb = (int)((n + RED_THREADS - 1) / RED_THREADS);
    if (nb > RED_BLOCKS) nb = RED_BLOCKS;
    if (nb < 1) nb = 1;
    k_reduce_partial<K, F><<<nb, RED_THREADS, 0, s->stream>>>(n, f, s->red_partial.p);
    HOT_LAUNCHED(s);
    k_reduce_final<K><<<1, RED_THREADS, 0, s->stream>>>(nb, s->red_partial.p, dev_out);
    HOT_LAUNCHED(s);
    if (host_out) {
        HOT_CUDA(cudaMemcpyAsync(s->h_red, dev_out, K * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
        HOT_CUDA(cudaStreamSynchronize(s->stream));
        for (int k = 0; k < K; ++k) host_out[k] = s->h_red[k];
    }
    return 0;
}

struct DotF { // sum a_i b_i
    const double *a, *b;
    __device__ void operator()(long i, double (&acc)[1]) const { acc[0] += a[i] * b[i]; }
};

struct OwnDotF { // the same over the nodes this rank counts (partitioned object)
    const double *a, *b;
    const unsigned char* own;
    __device__ void operator()(long i, double (&acc)[1]) const
    {
        if (own[i / 3]) acc[0] += a[i] * b[i];
    }
};

} // namespace hot
