// Particle->grid scatters of the hot path (a6 P2G, a12 force rasterisation, a13 matrix-free Hessian apply, a18 CN tolerance):
// one CTA per (half) page group, no colour passes, no atomics in the particle loop.
//
// The reference serialises these scatters into 8 colour passes (MpmSimulationBase.h:251-264) and does a read-modify-write of
// a 128-byte GridState per (particle, node).  Here the particles of a page group (one contiguous, coalesced run of every
// sorted SoA row) go through three phases:
//   prep       thread per particle: everything that does not depend on the stencil node (B-spline weights of the three axes
//              in the reference's operation order, the policy's payload) is computed once and parked in shared memory as a
//              particle-major record;
//   accumulate thread per (cell, part of the 3x3x3 stencil) walks the particles of ITS cell - adjacent thanks to the sort key -
//              and keeps its nodes x channels in registers;
//   combine    every thread parks its sums, then thread (tile node) adds the contributions of its node in a fixed order read
//              (a closed-form walk, no index table) and issues one fp64 RED per channel.
// Two thread mappings share this structure: the COLUMN form (9 threads per cell, 3 nodes each; half-page CTAs) and the PLANE
// form (3 threads per cell, 9 nodes each; cells handed to lanes by decreasing particle count).  Which one a policy uses is
// a measured choice (Policy::PLANE), see DESIGN.md.
#pragma once
#include "sim.h"
#include <cstdlib>

namespace hot {

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

// ---- column scatter -----------------------------------------------------------------
// One CTA per HALF page group (CS_SPLIT = 2: the 16 cells of one x-layer of the 2x4x4 page), thread = (cell, (i, j) column of
// the 3x3x3 stencil): 9 threads per cell, each keeping the 3 nodes x NCH channels of its z-column in registers.
// History (ncu): the first skeleton of this file had thread = (cell, x-plane) re-deriving the weights per thread
// (16.4 of 32 lanes active in its accumulate loop, 29 % of the stalls on the CTA barrier behind the slowest cell, 128
// registers -> 15 warps per SM).  Against that the column form has
//   * a warp holding 3.6 cells instead of 10.7, so the per-cell particle loops of its lanes diverge far less;
//   * everything that does not depend on the stencil node computed ONCE per particle by the prep pass, so the (cell, column)
//     threads issue ~30 fp64 instructions per particle instead of ~145;
//   * 12 accumulators instead of 36: <= 72-96 registers; a half page of the usual 8-12 particles per cell is ONE prep pass of
//     CS_CHUNK = 192 particles, 34-46 KB of shared memory, 4-5 CTAs per SM at different phases;
//   * a gather combine: every thread parks its 3 x NCH sums, then thread (node) adds the contributions of its node in a fixed
//     order (closed-form walk over the (cell, column, k) triples that feed the node) - no shared-memory atomics, no warp
//     tiles, deterministic inside the CTA - and issues one RED per channel.
// Its weakness is shared-memory bandwidth (9 lanes fetch the same per-particle words), which is why the plane form below, with
// compact records, is the default of the P2G and force scatters; the column form serves the CN-tolerance scatter and A/B runs.
// HBM traffic is the algorithmic minimum: every particle attribute is read once (coalesced runs of the sorted SoA rows),
// every touched node receives one RED per channel and half page.
template <int SPLIT>
struct ColGeo {
    static constexpr int CELLS = Geo::E / SPLIT; // SPLIT 2: one x-layer of cells (cell index = (cx << 4) | (cy << 2) | cz)
    static constexpr int THREADS = 9 * CELLS;
    static constexpr int TXH = Geo::BX / SPLIT + 2;
    static constexpr int NT = TXH * Geo::TY * Geo::TZ; // tile nodes of the half page
    static constexpr int NSRC = 27 * CELLS;
};
static_assert(Geo::BX == 2, "SPLIT = 2 halves the page along x");
constexpr int CS_SPLIT = 2;
constexpr int CS_CHUNK = 192; // particles per prep pass
constexpr int CS_THREADS = ColGeo<CS_SPLIT>::THREADS;
// Prepared records are PARTICLE-major, Policy::REC doubles each (REC even, REC / 2 odd: 16-byte aligned records whose
// starts walk over all 8 groups of 4 banks), so that the (cell, column) threads fetch them as 128-bit words: the accumulate
// loop is bound by shared-memory wavefronts (ncu: 1.7 wavefronts per LDS.64 even with 9 lanes reading the same word), and
// an LDS.128 moves twice the data per wavefront.
template <class Policy>
constexpr size_t cs_smem_bytes()
{
    static_assert(Policy::REC % 2 == 0 && (Policy::REC / 2) % 2 == 1, "record stride: odd multiple of 16 bytes");
    constexpr size_t rec = (size_t)Policy::REC * CS_CHUNK, con = (size_t)3 * Policy::NCH * CS_THREADS;
    return (rec > con ? rec : con) * sizeof(double);
}
__device__ __forceinline__ void lds2(const double* p, double& a, double& b)
{
    const double2 v = *reinterpret_cast<const double2*>(p);
    a = v.x; b = v.y;
}
__device__ __forceinline__ void sts2(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }
// Policy interface (column form):
//   static constexpr int NCH, REC (doubles per prepared particle), MINB (CTAs per SM the shared-memory footprint allows)
//   struct Args
//   __device__ static void prep(const Args&, size_t s, double* rec /* this particle's record */)
//   __device__ static void accumulate_col(const double* rec, int i, int j, double di, double dj, double (&acc)[3][NCH])
//   __device__ static void prefetch(const Args&, int first, int end, int tid, int nthreads)   L2 prefetch of the particle rows
//   static constexpr bool DOF                                                   target is a DOF vector (a = DOF id) or grid channels
//   __device__ static void flush1(const Args&, long a, int ch, double v)        a = DOF id / grid array index
// DBG: clock64 stamps of every CTA's phases are averaged into dbg[0..7] (HOT_CS_DEBUG; profiling aid, never the bench path)
template <class Policy, int MINB = Policy::MINB, bool DBG = false>
__global__ void __launch_bounds__(CS_THREADS, MINB) k_column_scatter(typename Policy::Args args, const int* __restrict__ cell_start,
    const int* __restrict__ group_slot, const int* __restrict__ nbr8, const int* __restrict__ tile_dof,
    int pf_dist, unsigned long long* dbg = nullptr)
{
    long long t_[8];
#define CS_STAMP(k) do { if (DBG) t_[k] = clock64(); } while (0)
    CS_STAMP(0);
    using G = ColGeo<CS_SPLIT>;
    constexpr int NCH = Policy::NCH, E = Geo::E, THREADS = G::THREADS;
    extern __shared__ __align__(16) double cs_smem[];
    __shared__ int s_cs[G::CELLS + 1];
    __shared__ int s_nbr[8];

    const int g = blockIdx.x / CS_SPLIT, h = blockIdx.x - g * CS_SPLIT, tid = threadIdx.x;
    const int* csg = cell_start + (size_t)g * (E + 1) + h * G::CELLS;
    // every thread reads the run bounds itself (one broadcast load) so that the prep loads below start without a CTA barrier
    const int first = csg[0], end = csg[G::CELLS];
    // the half page that will take over this CTA's slot (pf_dist CTAs ahead): its run bounds are requested now, used at the end
    int pf_first = 0, pf_end = 0;
    if (pf_dist > 0 && blockIdx.x + pf_dist < gridDim.x) {
        const int bp = blockIdx.x + pf_dist, gp = bp / CS_SPLIT, hp = bp - gp * CS_SPLIT;
        pf_first = cell_start[(size_t)gp * (E + 1) + hp * G::CELLS];
        pf_end = cell_start[(size_t)gp * (E + 1) + hp * G::CELLS + G::CELLS];
    }
    if (tid <= G::CELLS) s_cs[tid] = csg[tid];
    if (tid < 8) s_nbr[tid] = nbr8[(size_t)group_slot[g] * 8 + tid];
    if (first == end) return; // empty half
    const int c = tid / 9, ij = tid - 9 * c, i = ij / 3, j = ij - 3 * i;
    const double di = (double)i, dj = (double)j;
    double acc[3][NCH];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) acc[k][ch] = 0.0;

    for (int cb = first; cb < end; cb += CS_CHUNK) {
        const int cn = min(CS_CHUNK, end - cb);
        if (cb != first) __syncthreads(); // previous pass consumed
        if (cb == first) CS_STAMP(1);
        for (int k = tid; k < cn; k += THREADS) Policy::prep(args, (size_t)cb + k, cs_smem + (size_t)k * Policy::REC);
        if (cb == first) CS_STAMP(2);
        __syncthreads(); // records (and, on the first pass, s_cs / s_nbr) visible
        if (cb == first) CS_STAMP(3);
        const int my_b = s_cs[c], my_e = s_cs[c + 1];
        const int pb = max(my_b, cb) - cb, pe = min(my_e, cb + cn) - cb;
        for (int p = pb; p < pe; ++p) Policy::accumulate_col(cs_smem + (size_t)p * Policy::REC, i, j, di, dj, acc);
    }
    CS_STAMP(4);
    Policy::prefetch(args, pf_first, pf_end, tid, THREADS);
    __syncthreads(); // records dead -> reuse as the contribution array [ch][k][tid]
    CS_STAMP(5);
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) cs_smem[(ch * 3 + k) * THREADS + tid] = acc[k][ch];
    __syncthreads();
    // gather combine: tile node (txl, ty, tz) of the half page receives column (i = txl, j) / node k of cell (0, ty - j, tz - k);
    // the source index moves by a constant per loop step, so the walk needs no index table and no dependent loads
    if (tid < G::NT) {
        const int tz = tid % Geo::TZ, ty = (tid / Geo::TZ) % Geo::TY, txl = tid / (Geo::TZ * Geo::TY);
        const int j0 = max(0, ty - (Geo::BY - 1)), j1 = min(2, ty), k0 = max(0, tz - (Geo::BZ - 1)), k1 = min(2, tz);
        double sum[NCH];
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) sum[ch] = 0.0;
        for (int j = j0; j <= j1; ++j) {
            int src = k0 * THREADS + ((((ty - j) << Geo::zb) | (tz - k0)) * 9) + txl * 3 + j;
            for (int k = k0; k <= k1; ++k, src += THREADS - 9) {
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) sum[ch] += cs_smem[ch * 3 * THREADS + src];
            }
        }
        // node tid of the half tile = node (h + txl, ty, tz) of the page tile; DOF-vector targets take the node's DOF id from the
        // per-step table, grid-channel targets (P2G runs before the numbering) the grid slot
        const int nfull = tid + h * (Geo::TY * Geo::TZ);
        const long a = Policy::DOF ? (long)tile_dof[(size_t)g * Geo::TILE + nfull] : tile_to_grid(nfull, s_nbr);
        if (a >= 0) {
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch)
                if (sum[ch] != 0.0) Policy::flush1(args, a, ch, sum[ch]);
        }
    }
    if (DBG && (tid & 31) == 0 && end - first <= CS_CHUNK) { // per warp: start->loads issued, prep, barrier wait, accumulate, ..., total
        CS_STAMP(7);
        for (int k = 1; k < 8; ++k) atomicAdd(dbg + k, (unsigned long long)(t_[k] - t_[k - 1]));
        atomicAdd(dbg, 1ull);
    }
#undef CS_STAMP
}

// prefetch distance in CTAs: HOT_PF_DIST if set, else one wave of resident CTAs (148 SMs x ctas_per_sm)
inline int pf_distance(Sim* s, int ctas_per_sm)
{
    if (s->pf_dist < 0) {
        const char* e = getenv("HOT_PF_DIST");
        s->pf_dist = e ? atoi(e) : 148;
    }
    return s->pf_dist * ctas_per_sm;
}
// launch over this rank's page groups [g0, g1)
template <class Policy>
int launch_column_scatter(Sim* s, const typename Policy::Args& a)
{
    if (s->g1 <= s->g0) return 0;
    HOT_FUNC_ATTR_ONCE(s, (k_column_scatter<Policy>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cs_smem_bytes<Policy>());
    static int dbg_runs = getenv("HOT_CS_DEBUG") ? atoi(getenv("HOT_CS_DEBUG")) : 0;
    if (dbg_runs > 0) {
        --dbg_runs;
        HOT_FUNC_ATTR_ONCE(s, (k_column_scatter<Policy, Policy::MINB, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cs_smem_bytes<Policy>());
        unsigned long long* d = nullptr;
        unsigned long long h[8] = {0};
        HOT_CUDA(cudaMalloc((void**)&d, sizeof h));
        HOT_CUDA(cudaMemcpy(d, h, sizeof h, cudaMemcpyHostToDevice));
        k_column_scatter<Policy, Policy::MINB, true><<<(unsigned)(CS_SPLIT * (s->g1 - s->g0)), CS_THREADS, cs_smem_bytes<Policy>(), s->stream>>>(a,
            s->cell_start.p + s->g0 * (Geo::E + 1), s->group_slot.p + s->g0, s->nbr8.p,
            Policy::DOF ? s->tile_dof.p + (size_t)s->g0 * Geo::TILE : nullptr, pf_distance(s, Policy::MINB), d);
        HOT_CUDA(cudaStreamSynchronize(s->stream));
        HOT_CUDA(cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost));
        cudaFree(d);
        const double n = h[0] ? (double)h[0] : 1.0;
        fprintf(stderr, "[cs dbg] NCH %d: %llu warps; mean cycles: bounds+issue %.0f | prep %.0f | barrier %.0f | accumulate %.0f | prefetch+barrier %.0f | park+barrier %.0f | combine+flush %.0f\n",
            Policy::NCH, h[0], h[1] / n, h[2] / n, h[3] / n, h[4] / n, h[5] / n, h[6] / n, h[7] / n);
        s->launches++;
        return 0;
    }
    k_column_scatter<Policy><<<(unsigned)(CS_SPLIT * (s->g1 - s->g0)), CS_THREADS, cs_smem_bytes<Policy>(), s->stream>>>(a,
        s->cell_start.p + s->g0 * (Geo::E + 1), s->group_slot.p + s->g0, s->nbr8.p,
        Policy::DOF ? s->tile_dof.p + (size_t)s->g0 * Geo::TILE : nullptr, pf_distance(s, Policy::MINB));
    HOT_LAUNCHED(s);
    return 0;
}

// ---- plane scatter, second form ---------------------------------------------------------------------------------------
// The column skeleton above turned out to be bound by shared-memory wavefronts: its 9 threads per cell each fetch the same
// ~16 per-particle doubles (ncu: 19 wavefronts per particle, 66 us of the 113 us kernel, LDS.128 did not help).  Going back
// to thread = (cell, x-plane) - 3 threads per cell, 9 nodes x NCH accumulators each - cuts that traffic 3x per particle while
// keeping what the column form established: per-particle prep pass (weights once per particle), particle-major records,
// gather combine.  The plane form's own weakness - a warp holds 10.7 cells whose particle counts differ, 16 of
// 32 lanes active in the first plane kernel - is removed by handing the cells to the lanes in order of DECREASING particle
// count, which the gather combine permits (any thread may own any cell).  What then bounds the kernel is the serial phase
// chain of a page group at few resident CTAs: the records are COMPACT (the three B-spline arguments instead of nine weights,
// Policy::RECP doubles; the thread re-derives the weights) so that 4-5 CTAs fit per SM (P2G 104 -> 96 us, force scatter
// 103 -> 89 us at C2).
constexpr int PS_THREADS = 3 * Geo::E; // 96
constexpr int PS_NCHP = 4; // parked channels per (node, thread), padded
// Straight-line gather combine of the plane form.  Parked sums lie as [node (j,k) of the thread's plane][cell * 3 + plane][channel];
// lane (tz, ch) of a warp adds up the tile nodes (tx = 0..3, TY, tz): node (tx, ty, tz) receives node (j, k) of plane i of cell
// (tx - i, ty - j, tz - k).  Every shared-memory offset is a compile-time constant plus one per-lane term (12 tz + ch), the
// three z-sources are predicated per lane: no loops, no index arithmetic (the nested-loop form this replaces was 20 % of the
// kernel's instructions).  Lanes (tz, ch) of a half warp hit 16 different 8-byte banks.
template <int TY>
__device__ __forceinline__ void plane_combine(const double* __restrict__ park, int lane_off, bool pk0, bool pk1, bool pk2, double* __restrict__ out)
{
#pragma unroll
    for (int tx = 0; tx < 4; ++tx) {
        double sum = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int cx = tx - i;
            if (cx < 0 || cx >= Geo::BX) continue;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int cy = TY - j;
                if (cy < 0 || cy >= Geo::BY) continue;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int off = ((j * 3 + k) * PS_THREADS + (cx * Geo::BY + cy) * Geo::BZ * 3 + i) * PS_NCHP - k * 3 * PS_NCHP;
                    const bool pk = k == 0 ? pk0 : (k == 1 ? pk1 : pk2);
                    if (pk) sum += park[off + lane_off];
                }
            }
        }
        out[tx] = sum;
    }
}
constexpr int PS_CAP = 384; // particles per prep pass: a full page at 12 particles per cell
template <class Policy>
constexpr size_t ps_smem_bytes()
{
    constexpr size_t rec = (size_t)Policy::RECP * PS_CAP, con = (size_t)9 * Policy::NCH * PS_THREADS;
    return (rec > con ? rec : con) * sizeof(double);
}
// Policy interface (plane form): NCH, DOF, prefetch, flush1 as for the column form and
//   static constexpr int RECP (doubles per plane-form record), PMINB (CTAs per SM its shared-memory footprint allows)
//   __device__ static void prep_plane(const Args&, size_t s, double* rec)
//   __device__ static void accumulate_plane(const Args&, const double* rec, int i, double di, double (&acc)[9][NCH])
template <class Policy>
__global__ void __launch_bounds__(PS_THREADS, Policy::PMINB) k_plane2_scatter(typename Policy::Args args, const int* __restrict__ cell_start,
    const int* __restrict__ group_slot, const int* __restrict__ nbr8, const int* __restrict__ tile_dof,
    int pf_dist)
{
    constexpr int NCH = Policy::NCH, E = Geo::E, THREADS = PS_THREADS;
    extern __shared__ __align__(16) double cs_smem[];
    __shared__ int s_cs[E + 1];
    __shared__ int s_order[E];
    __shared__ int s_nbr[8];

    const int g = blockIdx.x, tid = threadIdx.x;
    const int* csg = cell_start + (size_t)g * (E + 1);
    const int first = csg[0], end = csg[E]; // every thread: the prep loads below start without waiting for a CTA barrier
    int pf_first = 0, pf_end = 0;
    if (pf_dist > 0 && g + pf_dist < (int)gridDim.x) {
        pf_first = cell_start[(size_t)(g + pf_dist) * (E + 1)];
        pf_end = cell_start[(size_t)(g + pf_dist) * (E + 1) + E];
    }
    if (tid <= E) s_cs[tid] = csg[tid];
    if (tid < 8) s_nbr[tid] = nbr8[(size_t)group_slot[g] * 8 + tid];
    double acc[9][NCH];
#pragma unroll
    for (int a = 0; a < 9; ++a)
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) acc[a][ch] = 0.0;
    int c = 0;
    const int i = tid % 3;
    const double di = (double)i;
    for (int cb = first; cb < end; cb += PS_CAP) {
        const int cn = min(PS_CAP, end - cb);
        if (cb != first) __syncthreads(); // previous pass consumed
        for (int k = tid; k < cn; k += THREADS) Policy::prep_plane(args, (size_t)cb + k, cs_smem + (size_t)k * Policy::RECP);
        __syncthreads(); // records (and, on the first pass, s_cs / s_nbr) visible
        if (cb == first) {
            // cells to lanes in order of decreasing particle count: the lanes of a warp then run loops of similar length
            if (tid < E) {
                const int cnt = s_cs[tid + 1] - s_cs[tid];
                int rank = 0;
                for (int o = 0; o < E; ++o) {
                    const int co = s_cs[o + 1] - s_cs[o];
                    rank += (co > cnt) || (co == cnt && o < tid);
                }
                s_order[rank] = tid;
            }
            __syncthreads();
            c = s_order[tid / 3];
        }
        const int my_b = s_cs[c], my_e = s_cs[c + 1];
        const int pb = max(my_b, cb) - cb, pe = min(my_e, cb + cn) - cb;
        for (int p = pb; p < pe; ++p) Policy::accumulate_plane(args, cs_smem + (size_t)p * Policy::RECP, i, di, acc);
    }
    Policy::prefetch(args, pf_first, pf_end, tid, THREADS);
    __syncthreads(); // records dead -> reuse as the parked sums [jk][cell * 3 + i][ch]
    {
        double* park = cs_smem + (size_t)(c * 3 + i) * PS_NCHP;
#pragma unroll
        for (int a = 0; a < 9; ++a) {
            double* d = park + (size_t)a * THREADS * PS_NCHP;
            sts2(d, acc[a][0], NCH > 1 ? acc[a][1 % NCH] : 0.0);
            if (NCH > 2) sts2(d + 2, acc[a][2 % NCH], NCH > 3 ? acc[a][3 % NCH] : 0.0);
        }
    }
    __syncthreads();
    // static gather combine: warp w owns the tile rows ty in {2,0} / {3,5} / {1,4} (24 source pairs each), lane = (tz, channel)
    static_assert(Geo::BX == 2 && Geo::BY == 4 && Geo::BZ == 4 && NCH <= PS_NCHP, "the combine is written for the 2x4x4 page");
    const int w = tid >> 5, lane = tid & 31;
    const int ctz = lane / PS_NCHP, cch = lane - ctz * PS_NCHP;
    const bool c_on = ctz < Geo::TZ && cch < NCH;
    const int lane_off = 3 * PS_NCHP * ctz + cch;
    const bool pk0 = c_on && ctz < Geo::BZ, pk1 = c_on && ctz >= 1 && ctz - 1 < Geo::BZ, pk2 = c_on && ctz >= 2 && ctz - 2 < Geo::BZ;
    double out[8];
    int tya, tyb;
    if (w == 0) { plane_combine<2>(cs_smem, lane_off, pk0, pk1, pk2, out); plane_combine<0>(cs_smem, lane_off, pk0, pk1, pk2, out + 4); tya = 2; tyb = 0; }
    else if (w == 1) { plane_combine<3>(cs_smem, lane_off, pk0, pk1, pk2, out); plane_combine<5>(cs_smem, lane_off, pk0, pk1, pk2, out + 4); tya = 3; tyb = 5; }
    else { plane_combine<1>(cs_smem, lane_off, pk0, pk1, pk2, out); plane_combine<4>(cs_smem, lane_off, pk0, pk1, pk2, out + 4); tya = 1; tyb = 4; }
    if (!c_on) return;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int ty = h == 0 ? tya : tyb;
#pragma unroll
        for (int tx = 0; tx < 4; ++tx) {
            const double v = out[h * 4 + tx];
            if (v == 0.0) continue;
            const int n = (tx * Geo::TY + ty) * Geo::TZ + ctz;
            const long a = Policy::DOF ? (long)tile_dof[(size_t)g * Geo::TILE + n] : tile_to_grid(n, s_nbr);
            if (a >= 0) Policy::flush1(args, a, cch, v);
        }
    }
}
template <class Policy>
int launch_plane2_scatter(Sim* s, const typename Policy::Args& a)
{
    if (s->g1 <= s->g0) return 0;
    HOT_FUNC_ATTR_ONCE(s, (k_plane2_scatter<Policy>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ps_smem_bytes<Policy>());
    k_plane2_scatter<Policy><<<(unsigned)(s->g1 - s->g0), PS_THREADS, ps_smem_bytes<Policy>(), s->stream>>>(a,
        s->cell_start.p + s->g0 * (Geo::E + 1), s->group_slot.p + s->g0, s->nbr8.p,
        Policy::DOF ? s->tile_dof.p + (size_t)s->g0 * Geo::TILE : nullptr, pf_distance(s, Policy::PMINB));
    HOT_LAUNCHED(s);
    return 0;
}
// which skeleton a scatter uses: the policy's measured default, or HOT_SCATTER = plane | column for A/B runs
template <class Policy>
int launch_scatter(Sim* s, const typename Policy::Args& a)
{
    static const int forced = [] {
        const char* e = getenv("HOT_SCATTER");
        return !e ? 0 : (e[0] == 'c' ? 1 : 2);
    }();
    const bool plane = forced ? forced == 2 : Policy::PLANE;
    return plane ? launch_plane2_scatter<Policy>(s, a) : launch_column_scatter<Policy>(s, a);
}

// B-spline weights (and, with GRAD, weight derivatives / dx) of the three axes of one particle in the reference's operation
// order; d0n = x_node(base) - x_p per axis
template <bool GRAD>
__device__ __forceinline__ void prep_weights(const double (&Xp)[3], double dx, double one_over_dx, double (&w)[3][3], double (&g)[3][3], double (&d0n)[3])
{
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        double xi, dw[3];
        const int b = base_node_of(Xp[d], one_over_dx, &xi);
        bspline_axis(xi - (double)b, w[d], GRAD ? dw : nullptr);
        d0n[d] = (double)b * dx - Xp[d];
        if (GRAD) {
#pragma unroll
            for (int t = 0; t < 3; ++t) g[d][t] = one_over_dx * dw[t];
        }
    }
}

// B-spline evaluation of one particle
struct SplineEval {
    double w[3][3], dw[3][3], d0n[3]; // weights, weight derivatives (not yet / dx), x_node(base) - x_p
    int base[3];
    __device__ __forceinline__ void eval3(const double (&Xp)[3], double dx, double one_over_dx, bool want_dw)
    {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            double xi;
            base[d] = base_node_of(Xp[d], one_over_dx, &xi);
            bspline_axis(xi - (double)base[d], w[d], want_dw ? dw[d] : nullptr);
            d0n[d] = (double)base[d] * dx - Xp[d];
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
