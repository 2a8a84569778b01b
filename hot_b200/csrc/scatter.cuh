// Shared skeleton of the particle->grid scatters of the hot path (a6 P2G, a12 force rasterisation, a13 matrix-free
// Hessian apply, a18 CN tolerance): one CTA per page group, thread = (cell of the page, x-plane of the 3x3x3 stencil).
//
// The reference serialises these scatters into 8 colour passes (MpmSimulationBase.h:251-264) and does a
// read-modify-write of a 128-byte GridState per (particle, node).  Here (see transfer.cu for the roofline argument):
//   stage      the particle's SoA attributes (coalesced loads of one contiguous run) are reduced to a small RAW record
//              (position + the policy's per-particle payload, e.g. m, m v, m C for P2G or the 3x3 matrix T for the vector
//              scatters) in shared memory, field-major and index-swizzled so that the strided reads below are conflict-free;
//              up to SC_CHUNK = 3 x CTA-size particles per pass, so a typical page group (~8-12 particles per cell) is ONE
//              pass and every (cell, plane) thread has work;
//   accumulate thread (c, pl) walks the particles of cell c (adjacent thanks to the sort key), re-derives the B-spline
//              weights from the position (cheaper than staging them: 12 extra doubles per particle would halve the chunk)
//              and keeps the 9 nodes x NCH channels of x-plane pl in registers - no atomics, no shared-memory traffic but
//              the record reads;
//   combine    per-warp (B+2)^3 tiles in shared memory, 9 (j,k) steps; inside a step the lanes of a warp hit distinct nodes
//              (a warp holds < 16 consecutive cells => distinct (cy,cz); the 3 planes of a cell are distinct x), so plain
//              RMW + __syncwarp is race free;
//   flush      warp tiles are summed and written with one fp64 RED per touched node and channel.
#pragma once
#include "sim.h"

namespace hot {

constexpr int SC_THREADS = 3 * Geo::E; // 96 for the 2x4x4 fp64 page
constexpr int SC_WARPS = SC_THREADS / 32;
constexpr int SC_CHUNK = 3 * SC_THREADS; // particles staged per pass
constexpr int SC_PAD = SC_CHUNK + SC_CHUNK / 8 + 2; // swizzled row length
static_assert(SC_THREADS % 32 == 0, "whole warps");
// neighbouring cells read records ~ppc apart: p + p/8 spreads them over the banks for the usual 4..16 particles per cell
__device__ __forceinline__ int sc_swz(int p) { return p + (p >> 3); }

// quadratic B-spline weights of one axis in the reference's operation order (BSplines.h:55-81)
__device__ __forceinline__ void bspline_axis(double d0, double* w, double* dw)
{
    double z = 1.5 - d0;
    w[0] = 0.5 * (z * z);
    double d1 = d0 - 1.0;
    w[1] = 0.75 - d1 * d1;
    double d2 = 1.0 - d1;
    double zz = 1.5 - d2;
    w[2] = 0.5 * (zz * zz);
    if (dw) {
        dw[0] = -z;
        dw[1] = -2.0 * d1;
        dw[2] = zz;
    }
}

// tile node -> grid array index (page neighbour q, in-page element e); -1 if the page is absent
__device__ __forceinline__ long tile_to_grid(int n, const int* __restrict__ nbr)
{
    int tz = n % Geo::TZ, ty = (n / Geo::TZ) % Geo::TY, tx = n / (Geo::TZ * Geo::TY);
    int q = ((tx >= Geo::BX) << 2) | ((ty >= Geo::BY) << 1) | (tz >= Geo::BZ);
    int e = (((tx & (Geo::BX - 1)) << Geo::yb | (ty & (Geo::BY - 1))) << Geo::zb) | (tz & (Geo::BZ - 1));
    int slot = nbr[q];
    return slot < 0 ? -1 : (long)slot * Geo::E + e;
}

// in-page cell coordinates of a base node
__device__ __forceinline__ int tile_base(int bx, int by, int bz)
{
    return (((bx & (Geo::BX - 1)) * Geo::TY) + (by & (Geo::BY - 1))) * Geo::TZ + (bz & (Geo::BZ - 1));
}

// Policy interface:
//   static constexpr int NCH, RAW (doubles per staged particle; fields 0..2 are the position), GATHER (0/1: stage needs a
//                                  gathered DOF field tile)
//   struct Args { ... }                              kernel arguments (trivially copyable)
//   __device__ static void stage(const Args&, size_t s, double* rec /* field f at rec[f * SC_PAD] */, const double* gtile)
//   __device__ static void accumulate(const Args&, const double* rec, int pl, double (&acc)[9][NCH])
//   __device__ static void flush(const Args&, long a, const double (&v)[NCH])     a = grid array index
//   __device__ static void gather_node(const Args&, long a, double (&v)[3])       (GATHER only)
template <class Policy, int MINB = 5, int UNROLL = 1>
__global__ void __launch_bounds__(SC_THREADS, MINB) k_plane_scatter(typename Policy::Args args, const int* __restrict__ cell_start,
    const int* __restrict__ group_slot, const int* __restrict__ nbr8)
{
    constexpr int NCH = Policy::NCH, RAW = Policy::RAW, TILE = Geo::TILE, E = Geo::E;
    constexpr int REC_DOUBLES = RAW * SC_PAD, TILE_DOUBLES = SC_WARPS * NCH * TILE;
    // the warp tiles alias the record buffer: records are dead once the last chunk has been accumulated
    __shared__ __align__(16) double smem[REC_DOUBLES > TILE_DOUBLES ? REC_DOUBLES : TILE_DOUBLES];
    __shared__ double gtile[Policy::GATHER ? 3 * TILE : 1];
    __shared__ int s_cs[E + 1];
    __shared__ int s_nbr[8];

    const int g = blockIdx.x, tid = threadIdx.x;
    if (tid <= E) s_cs[tid] = cell_start[(size_t)g * (E + 1) + tid];
    if (tid < 8) s_nbr[tid] = nbr8[(size_t)group_slot[g] * 8 + tid];
    __syncthreads();
    if (Policy::GATHER) {
        for (int n = tid; n < TILE; n += SC_THREADS) {
            double v[3] = {0.0, 0.0, 0.0};
            long a = tile_to_grid(n, s_nbr);
            if (a >= 0) Policy::gather_node(args, a, v);
            gtile[n] = v[0]; gtile[TILE + n] = v[1]; gtile[2 * TILE + n] = v[2];
        }
    }
    const int first = s_cs[0], end = s_cs[E];
    const int c = tid / 3, pl = tid - 3 * c;
    const int my_b = s_cs[c], my_e = s_cs[c + 1];
    double acc[9][NCH];
#pragma unroll
    for (int a = 0; a < 9; ++a)
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) acc[a][ch] = 0.0;

    for (int cb = first; cb < end; cb += SC_CHUNK) {
        const int cn = min(SC_CHUNK, end - cb);
        __syncthreads(); // previous chunk consumed (and gtile ready on the first pass)
        for (int k = tid; k < cn; k += SC_THREADS) Policy::stage(args, (size_t)cb + k, smem + sc_swz(k), gtile);
        __syncthreads();
        const int pb = max(my_b, cb) - cb, pe = min(my_e, cb + cn) - cb;
        int p = pb;
        if (UNROLL == 2) { // two particles per trip: their weight evaluations are independent instruction streams (ILP)
            for (; p + 1 < pe; p += 2) {
                Policy::accumulate(args, smem + sc_swz(p), pl, acc);
                Policy::accumulate(args, smem + sc_swz(p + 1), pl, acc);
            }
        }
        for (; p < pe; ++p) Policy::accumulate(args, smem + sc_swz(p), pl, acc);
    }
    __syncthreads(); // records dead -> reuse as warp tiles
    for (int a = tid; a < TILE_DOUBLES; a += SC_THREADS) smem[a] = 0.0;
    __syncthreads();
    {
        const int cz = c & (Geo::BZ - 1), cy = (c >> Geo::zb) & (Geo::BY - 1), cx = c >> (Geo::zb + Geo::yb);
        double* wt = smem + (size_t)(tid >> 5) * NCH * TILE;
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int n = ((cx + pl) * Geo::TY + (cy + j)) * Geo::TZ + (cz + k);
                if (my_e > my_b) {
#pragma unroll
                    for (int ch = 0; ch < NCH; ++ch) wt[ch * TILE + n] += acc[j * 3 + k][ch];
                }
                __syncwarp();
            }
    }
    __syncthreads();
    for (int n = tid; n < TILE; n += SC_THREADS) {
        double v[NCH];
        bool any = false;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            double sum = 0.0;
#pragma unroll
            for (int w = 0; w < SC_WARPS; ++w) sum += smem[(size_t)w * NCH * TILE + ch * TILE + n];
            v[ch] = sum;
            any |= sum != 0.0;
        }
        if (any) {
            long a = tile_to_grid(n, s_nbr);
            if (a >= 0) Policy::flush(args, a, v);
        }
    }
}

// B-spline evaluation of one particle
struct SplineEval {
    double w[3][3], dw[3][3], d0n[3]; // weights, weight derivatives (not yet / dx), x_node(base) - x_p
    int base[3];
    // from a staged record (fields 0..2 = position)
    __device__ __forceinline__ void eval_rec(const double* __restrict__ rec, double dx, double one_over_dx, bool want_dw)
    {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            double Xd = rec[d * SC_PAD], xi;
            base[d] = base_node_of(Xd, one_over_dx, &xi);
            bspline_axis(xi - (double)base[d], w[d], want_dw ? dw[d] : nullptr);
            d0n[d] = (double)base[d] * dx - Xd;
        }
    }
    __device__ __forceinline__ void eval(const double* __restrict__ X, size_t ps, size_t s, double dx, double one_over_dx, bool want_dw)
    {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            double Xd = X[d * ps + s], xi;
            base[d] = base_node_of(Xd, one_over_dx, &xi);
            bspline_axis(xi - (double)base[d], w[d], want_dw ? dw[d] : nullptr);
            d0n[d] = (double)base[d] * dx - Xd;
        }
    }
};

} // namespace hot
