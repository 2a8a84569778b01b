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
#include <cstdlib>

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

// ---- column scatter (the version the hot path uses) -----------------------------------------------------------------
// One CTA per HALF page group (SPLIT = 2: the 16 cells of one x-layer of the 2x4x4 page), thread = (cell, (i, j) column of
// the 3x3x3 stencil): 9 threads per cell, each keeping the 3 nodes x NCH channels of its z-column in registers.
// Compared with the plane skeleton above (ncu: 16.4 of 32 lanes active in its accumulate loop, 29 % of the stalls on the CTA
// barrier behind the slowest cell, 128 registers -> 15 warps per SM):
//   * a warp holds 3.6 cells instead of 10.7, so the per-cell particle loops of its lanes diverge far less;
//   * everything that does not depend on the stencil node is computed ONCE per particle by a thread-per-particle prep
//     pass (B-spline weights of the three axes in the reference's operation order, the policy's payload) and parked in
//     shared memory field-major, so the (cell, column) threads issue ~30 fp64 instructions per particle instead of ~145;
//   * 12 accumulators instead of 36: <= 72 registers; a half page of the usual 8-12 particles per cell is ONE prep pass of
//     CHUNK = 192 particles (a second pass serialises a DRAM round trip behind a quarter-full accumulate loop), 34-42 KB of
//     shared memory, 5 CTAs per SM at different phases;
//   * combine is a gather: every thread parks its 3 x NCH sums, then thread (node) adds the contributions of its node in a
//     fixed order read from a tiny index table (csr_start / csr_src: which (cell, column, k) feed which tile node) - no
//     shared-memory atomics, no warp tiles, deterministic inside the CTA - and issues one RED per channel.
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
constexpr int CS_PAD = CS_CHUNK + 1; // odd row length: field f of particle p sits in bank 2 (f + p) mod 32
constexpr int CS_THREADS = ColGeo<CS_SPLIT>::THREADS;
template <class Policy>
constexpr size_t cs_smem_bytes()
{
    constexpr size_t rec = (size_t)Policy::REC * CS_PAD, con = (size_t)3 * Policy::NCH * CS_THREADS;
    return (rec > con ? rec : con) * sizeof(double);
}
// host side of the combine table: start[NT + 1] then src[NSRC]; src = k * THREADS + cell_local * 9 + i * 3 + j
inline void cs_build_table(short* tab)
{
    using G = ColGeo<CS_SPLIT>;
    short* start = tab;
    short* src = tab + G::NT + 1;
    int fill[G::NT + 1] = {0};
    for (int pass = 0; pass < 2; ++pass) {
        for (int c = 0; c < G::CELLS; ++c) {
            const int cz = c & (Geo::BZ - 1), cy = (c >> Geo::zb) & (Geo::BY - 1), cx = c >> (Geo::zb + Geo::yb);
            for (int ij = 0; ij < 9; ++ij)
                for (int k = 0; k < 3; ++k) {
                    const int n = ((cx + ij / 3) * Geo::TY + (cy + ij % 3)) * Geo::TZ + (cz + k);
                    if (pass == 0) fill[n + 1]++;
                    else src[fill[n]++] = (short)(k * G::THREADS + c * 9 + ij);
                }
        }
        if (pass == 0) {
            for (int n = 0; n < G::NT; ++n) fill[n + 1] += fill[n];
            for (int n = 0; n <= G::NT; ++n) start[n] = (short)fill[n];
        }
    }
}
constexpr int CS_TABLE_LEN = ColGeo<CS_SPLIT>::NT + 1 + ColGeo<CS_SPLIT>::NSRC;

// Policy interface (column form):
//   static constexpr int NCH, REC (doubles per prepared particle)
//   struct Args
//   __device__ static void prep(const Args&, size_t s, double* rec /* field f at rec[f * CS_PAD] */)
//   __device__ static void accumulate_col(const double* rec, int i, int j, double di, double dj, double (&acc)[3][NCH])
//   __device__ static void prefetch(const Args&, int first, int end, int tid, int nthreads)   L2 prefetch of the particle rows
//   static constexpr bool DOF                                                   target is a DOF vector (a = DOF id) or grid channels
//   __device__ static void flush1(const Args&, long a, int ch, double v)        a = DOF id / grid array index
template <class Policy, int MINB = 5>
__global__ void __launch_bounds__(CS_THREADS, MINB) k_column_scatter(typename Policy::Args args, const int* __restrict__ cell_start,
    const int* __restrict__ group_slot, const int* __restrict__ nbr8, const short* __restrict__ table, const int* __restrict__ tile_dof,
    int pf_dist)
{
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
        for (int k = tid; k < cn; k += THREADS) Policy::prep(args, (size_t)cb + k, cs_smem + k);
        __syncthreads(); // records (and, on the first pass, s_cs / s_nbr) visible
        const int my_b = s_cs[c], my_e = s_cs[c + 1];
        const int pb = max(my_b, cb) - cb, pe = min(my_e, cb + cn) - cb;
        for (int p = pb; p < pe; ++p) Policy::accumulate_col(cs_smem + p, i, j, di, dj, acc);
    }
    Policy::prefetch(args, pf_first, pf_end, tid, THREADS);
    __syncthreads(); // records dead -> reuse as the contribution array [ch][k][tid]
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) cs_smem[(ch * 3 + k) * THREADS + tid] = acc[k][ch];
    __syncthreads();
    if (tid < G::NT) {
        const int e0 = table[tid], e1 = table[tid + 1];
        double sum[NCH];
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) sum[ch] = 0.0;
        for (int e = e0; e < e1; ++e) {
            const int src = table[G::NT + 1 + e];
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) sum[ch] += cs_smem[ch * 3 * THREADS + src];
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
    if (!s->cs_table_ready) {
        short tab[CS_TABLE_LEN];
        cs_build_table(tab);
        HOT_CUDA(s->cs_table.reserve(CS_TABLE_LEN));
        HOT_CUDA(cudaMemcpyAsync(s->cs_table.p, tab, sizeof tab, cudaMemcpyHostToDevice, s->stream));
        HOT_CUDA(cudaStreamSynchronize(s->stream)); // tab is a stack array
        s->cs_table_ready = true;
    }
    static const cudaError_t attr = cudaFuncSetAttribute(k_column_scatter<Policy>, cudaFuncAttributeMaxDynamicSharedMemorySize,
        (int)cs_smem_bytes<Policy>());
    HOT_CUDA(attr);
    k_column_scatter<Policy><<<(unsigned)(CS_SPLIT * (s->g1 - s->g0)), CS_THREADS, cs_smem_bytes<Policy>(), s->stream>>>(a,
        s->cell_start.p + s->g0 * (Geo::E + 1), s->group_slot.p + s->g0, s->nbr8.p, s->cs_table.p,
        Policy::DOF ? s->tile_dof.p + (size_t)s->g0 * Geo::TILE : nullptr, pf_distance(s, 5));
    HOT_LAUNCHED(s);
    return 0;
}

// B-spline weights of the three axes of one particle into a prepared record: fields 0..8 = w[axis][t]; with GRAD also
// fields 9..17 = dw[axis][t] / dx.  Returns x_node(base) - x_p per axis.
template <bool GRAD>
__device__ __forceinline__ void prep_weights(const double (&Xp)[3], double dx, double one_over_dx, double* __restrict__ rec, double (&d0n)[3])
{
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        double xi, w[3], dw[3];
        const int b = base_node_of(Xp[d], one_over_dx, &xi);
        bspline_axis(xi - (double)b, w, GRAD ? dw : nullptr);
        d0n[d] = (double)b * dx - Xp[d];
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            rec[(3 * d + t) * CS_PAD] = w[t];
            if (GRAD) rec[(9 + 3 * d + t) * CS_PAD] = one_over_dx * dw[t];
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
