// a15: assembly of the 125-slot block-row system matrix, its BC projection and diagonal; the matrix-free block-Jacobi
// diagonal of --matfree.
//
// Reference: ImplicitSolverObjective::buildMatrix<projectSystem> (Projects/multigrid/ImplicitSolver.h:470-603),
// buildDiagonal (:605-665), linearOffset (:465-468), FBasedMpmForceHelper::runLambdaWithDifferential
// (Lib/MPM/Force/FBasedMpmForceHelper.h:63-121), SquareMatrix::buildDiagonal (SquareMatrix.h:301-324).
//
// Re-design for the GPU.  The reference walks the particles in 8 colour passes and does, per particle, 378 node pairs x
// (9 dPdF block products + 2 read-modify-writes of 72-byte blocks).  Here
//  * the per-particle dense dPdF and the Fn^T grad w products are replaced by the contracted Hessian H~ (force.cu), so a
//    node pair costs  U_a = H~ . grad w_a  (81 FMA, once per a)  and  U_a . grad w_b  (27 FMA);
//  * all particles of one SPGrid cell share their 27 stencil nodes, so the 27x27 pair blocks are summed over the
//    particles of the cell in registers (thread = node a x one x-plane of 9 nodes b) and only the per-cell sums are
//    added to the matrix: ~ppc times fewer read-modify-writes, no colour passes.
#include "scatter.cuh"
#include "reduce.cuh"

namespace hot {

Sim::~Sim()
{
    for (MGLevel* l : levels) delete l;
}

namespace {

constexpr int TPB = 256;
inline int nblk(long n) { return (int)((n + TPB - 1) / TPB); }
constexpr int W = MGLevel::W;
constexpr int TILE = Geo::TILE;
constexpr int E = Geo::E;

__device__ __forceinline__ constexpr int tri(int i, int j) { return i <= j ? j * (j + 1) / 2 + i : i * (i + 1) / 2 + j; }

__global__ void k_id2coord(int n_nodes, const int* __restrict__ dof_slot, const uint32_t* __restrict__ page_id, int* __restrict__ coord)
{
    int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n_nodes) return;
    int a = dof_slot[id];
    uint64_t off = ((uint64_t)page_id[a / Geo::E] << 12) | ((uint64_t)(a % Geo::E) << Geo::data_bits);
    coord[3 * id] = (int)bit_pack(off, Geo::xmask);
    coord[3 * id + 1] = (int)bit_pack(off, Geo::ymask);
    coord[3 * id + 2] = (int)bit_pack(off, Geo::zmask);
}

// rows start as: every slot -> self with a zero block; slot 62 (offset 0) = m_i I   (ImplicitSolver.h:485-493)
__global__ void k_matrix_init(int n, const double* __restrict__ mass, int* __restrict__ col, double* __restrict__ val)
{
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)n * W) return;
    const int i = (int)(t / W), s = (int)(t - (long)i * W);
    col[t] = i;
    const double m = (s == 62 && mass) ? mass[i] : 0.0; // (partitioned: the inertia term is added after the rows have been summed)
#pragma unroll
    for (int q = 0; q < 9; ++q) val[((size_t)i * 9 + q) * W + s] = (q == 0 || q == 4 || q == 8) ? m : 0.0;
}

constexpr int AS_THREADS = 96;
constexpr int AS_CHUNK = 8; // particles of one cell staged per pass
constexpr int AS_REC = 126; // 27 x 3 weight gradients + 45 Hessian entries

__global__ void __launch_bounds__(AS_THREADS) k_assemble(const int* __restrict__ cell_start, const int* __restrict__ group_slot,
    const int* __restrict__ nbr8, size_t ps, const double* __restrict__ X, const double* __restrict__ H, double dx, double one_over_dx,
    double dt2, const int* __restrict__ g_idx, int* __restrict__ col, double* __restrict__ val)
{
    __shared__ double rec[AS_CHUNK][AS_REC];
    __shared__ int s_cs[E + 1];
    __shared__ int s_nbr[8];
    __shared__ int s_id[TILE];
    const int g = blockIdx.x, tid = threadIdx.x;
    if (tid <= E) s_cs[tid] = cell_start[(size_t)g * (E + 1) + tid];
    if (tid < 8) s_nbr[tid] = nbr8[(size_t)group_slot[g] * 8 + tid];
    __syncthreads();
    for (int n = tid; n < TILE; n += AS_THREADS) {
        long a = tile_to_grid(n, s_nbr);
        s_id[n] = a >= 0 ? g_idx[a] : -1;
    }
    // work item: node a of the cell's 27-stencil, x-plane bp of the partner nodes b
    const int a = tid / 3, bp = tid - 3 * a; // a < 27 for tid < 81
    const int ai = a / 9, aj = (a / 3) % 3, ak = a % 3;
    // Only the blocks of the upper triangle (offset a - b lexicographically >= 0, slot >= 62) are computed and added; k_mirror fills the
    // rest with the transposes (the reference, too, visits each node pair once and writes both blocks, ImplicitSolver.h:516-547):
    // half the REDs and half the block products.  bmask: which of the thread's 9 partner nodes (bj, bk) are on the upper side.
    unsigned bmask = 0;
    if (tid < 81)
        for (int b = 0; b < 9; ++b) {
            const int slot = (ai - bp + 2) * 25 + (aj - b / 3 + 2) * 5 + (ak - b % 3 + 2);
            if (slot >= 62) bmask |= 1u << b;
        }
    for (int c = 0; c < E; ++c) {
        const int cb = s_cs[c], ce = s_cs[c + 1];
        if (ce == cb) continue; // uniform over the CTA
        double acc[9][9];
#pragma unroll
        for (int b = 0; b < 9; ++b)
#pragma unroll
            for (int q = 0; q < 9; ++q) acc[b][q] = 0.0;
        for (int p0 = cb; p0 < ce; p0 += AS_CHUNK) {
            const int pn = min(AS_CHUNK, ce - p0);
            __syncthreads();
            // stage: weight gradients of the 27 nodes (reference association, MpmGrid.h:272-291) and H~
            for (int it = tid; it < pn * 27; it += AS_THREADS) {
                const int p = it / 27, nd = it - 27 * p;
                const int i = nd / 9, j = (nd / 3) % 3, k = nd % 3;
                SplineEval sp;
                sp.eval(X, ps, (size_t)p0 + p, dx, one_over_dx, true);
                const double wi = sp.w[0][i], wj = sp.w[1][j], wk = sp.w[2][k];
                const double dwidxi = one_over_dx * sp.dw[0][i];
                const double wij = wi * wj;
                rec[p][3 * nd] = (dwidxi * wj) * wk;
                rec[p][3 * nd + 1] = (wi * one_over_dx * sp.dw[1][j]) * wk;
                rec[p][3 * nd + 2] = wij * one_over_dx * sp.dw[2][k];
            }
            for (int it = tid; it < pn * 45; it += AS_THREADS) {
                const int p = it / 45, e = it - 45 * p;
                rec[p][81 + e] = H[(size_t)e * ps + p0 + p];
            }
            __syncthreads();
            if (bmask) {
                for (int p = 0; p < pn; ++p) {
                    const double* r = rec[p];
                    const double* Hh = r + 81;
                    const double ga[3] = {r[3 * a], r[3 * a + 1], r[3 * a + 2]};
                    // U(rr, s + 3 d) = sum_c H~(rr + 3 c, s + 3 d) ga[c]
                    double U[27];
#pragma unroll
                    for (int cc = 0; cc < 9; ++cc)
#pragma unroll
                        for (int rr = 0; rr < 3; ++rr)
                            U[rr + 3 * cc] = Hh[tri(rr, cc)] * ga[0] + Hh[tri(rr + 3, cc)] * ga[1] + Hh[tri(rr + 6, cc)] * ga[2];
#pragma unroll
                    for (int b = 0; b < 9; ++b) {
                        if (!(bmask >> b & 1)) continue;
                        const double* gb = r + 3 * (bp * 9 + b);
                        const double g0 = gb[0], g1 = gb[1], g2 = gb[2];
#pragma unroll
                        for (int ss = 0; ss < 3; ++ss)
#pragma unroll
                            for (int rr = 0; rr < 3; ++rr)
                                acc[b][rr + 3 * ss] += U[rr + 3 * ss] * g0 + U[rr + 3 * (ss + 3)] * g1 + U[rr + 3 * (ss + 6)] * g2;
                    }
                }
            }
        }
        if (tid < 81) {
            const int cz = c & (Geo::BZ - 1), cy = (c >> Geo::zb) & (Geo::BY - 1), cx = c >> (Geo::zb + Geo::yb);
            const int ida = s_id[((cx + ai) * Geo::TY + (cy + aj)) * Geo::TZ + (cz + ak)];
            if (ida >= 0) {
#pragma unroll
                for (int b = 0; b < 9; ++b) {
                    const int bj = b / 3, bk = b % 3;
                    const int idb = s_id[((cx + bp) * Geo::TY + (cy + bj)) * Geo::TZ + (cz + bk)];
                    if (idb < 0) continue;
                    const int slot = (ai - bp + 2) * 25 + (aj - bj + 2) * 5 + (ak - bk + 2);
                    col[(size_t)ida * W + slot] = idb;
                    if (!(bmask >> b & 1)) continue;
#pragma unroll
                    for (int q = 0; q < 9; ++q) atomicAdd(val + ((size_t)ida * 9 + q) * W + slot, dt2 * acc[b][q]);
                }
            }
        }
    }
}


// ---- assembly, default form: upper triangle, balanced threads, double-buffered staging ------------------------------------------
// Same arithmetic as k_assemble (per cell: U_a = H~ . grad w_a once per thread and particle, 27 FMA per node pair, per-cell sums in
// registers, REDs per cell), but
//  * only the 405 blocks per cell with offset a - b >= 0 (slot >= 62) are computed; k_mirror writes the transposes.  Threads own
//    (a, x-plane bp) with ai >= bp only: 54 busy threads instead of 81 of which a third would idle;
//  * the records of the NEXT chunk of particles (H~ rows by 8-byte cp.async, weight gradients computed by the staging threads) land
//    in the second buffer while the current chunk is contracted: one barrier per chunk, no exposed DRAM round trip;
//  * 64-thread CTAs: 4 per SM at 252 registers.
constexpr int AU_THREADS = 64;
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr(smem)), "l"(gmem) : "memory");
}
__global__ void __launch_bounds__(AU_THREADS) k_assemble_upper(const int* __restrict__ cell_start, const int* __restrict__ group_slot,
    const int* __restrict__ nbr8, size_t ps, const double* __restrict__ X, const double* __restrict__ H, double dx, double one_over_dx,
    double dt2, const int* __restrict__ g_idx, int* __restrict__ col, double* __restrict__ val)
{
    __shared__ __align__(16) double rec[2][AS_CHUNK][AS_REC];
    __shared__ int s_cs[E + 1];
    __shared__ int s_nbr[8];
    __shared__ int s_id[TILE];
    const int g = blockIdx.x, tid = threadIdx.x;
    if (tid <= E) s_cs[tid] = cell_start[(size_t)g * (E + 1) + tid];
    if (tid < 8) s_nbr[tid] = nbr8[(size_t)group_slot[g] * 8 + tid];
    __syncthreads();
    for (int n = tid; n < TILE; n += AU_THREADS) {
        long a = tile_to_grid(n, s_nbr);
        s_id[n] = a >= 0 ? g_idx[a] : -1;
    }
    // work item: node a = (ai, aj, ak) of the cell's stencil x x-plane bp <= ai of the partner nodes
    const bool worker = tid < 54;
    const int u = worker ? tid / 9 : 0, rj = tid % 9;
    const int ai = u == 0 ? 1 : (u == 1 || u == 2 ? 2 : u - 3), bp = u == 0 || u == 1 ? 0 : (u == 2 ? 1 : u - 3);
    const int aj = rj / 3, ak = rj % 3, a = ai * 9 + aj * 3 + ak;
    unsigned bmask = 0;
    if (worker)
        for (int b = 0; b < 9; ++b) {
            const int slot = (ai - bp + 2) * 25 + (aj - b / 3 + 2) * 5 + (ak - b % 3 + 2);
            if (slot >= 62) bmask |= 1u << b;
        }
    auto stage = [&](int buf, int p0, int pn) {
        for (int it = tid; it < pn * 45; it += AU_THREADS) {
            const int p = it / 45, e = it - 45 * p;
            cp_async8(&rec[buf][p][81 + e], H + (size_t)e * ps + p0 + p);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        for (int it = tid; it < pn * 27; it += AU_THREADS) {
            const int p = it / 27, nd = it - 27 * p;
            const int ijk[3] = {nd / 9, (nd / 3) % 3, nd % 3};
            double w[3], dws[3]; // weight and weight derivative / dx of this node per axis (few live registers next to acc / U)
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                double xi, wa[3], dwa[3];
                const int bn = base_node_of(X[d * ps + p0 + p], one_over_dx, &xi);
                bspline_axis(xi - (double)bn, wa, dwa);
                const int t = ijk[d];
                w[d] = t == 0 ? wa[0] : (t == 1 ? wa[1] : wa[2]);
                dws[d] = t == 0 ? dwa[0] : (t == 1 ? dwa[1] : dwa[2]);
            }
            rec[buf][p][3 * nd] = ((one_over_dx * dws[0]) * w[1]) * w[2];
            rec[buf][p][3 * nd + 1] = (w[0] * one_over_dx * dws[1]) * w[2];
            rec[buf][p][3 * nd + 2] = (w[0] * w[1]) * one_over_dx * dws[2];
        }
    };
    int c = 0;
    while (c < E && s_cs[c + 1] == s_cs[c]) ++c;
    if (c >= E) return; // (a group has particles; uniform)
    int p0 = s_cs[c], buf = 0;
    __syncthreads(); // s_id
    stage(0, p0, min(AS_CHUNK, s_cs[c + 1] - p0));
    double acc[9][9];
#pragma unroll
    for (int b = 0; b < 9; ++b)
#pragma unroll
        for (int q = 0; q < 9; ++q) acc[b][q] = 0.0;
    while (c < E) {
        const int ce = s_cs[c + 1], pn = min(AS_CHUNK, ce - p0);
        int nc = c, np0 = p0 + pn;
        if (np0 >= ce) {
            ++nc;
            while (nc < E && s_cs[nc + 1] == s_cs[nc]) ++nc;
            np0 = nc < E ? s_cs[nc] : 0;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads(); // this chunk's records visible; everybody is done with the other buffer
        if (nc < E) stage(buf ^ 1, np0, min(AS_CHUNK, s_cs[nc + 1] - np0));
        if (bmask) {
            for (int p = 0; p < pn; ++p) {
                const double* r = rec[buf][p];
                const double* Hh = r + 81;
                const double ga[3] = {r[3 * a], r[3 * a + 1], r[3 * a + 2]};
                double U[27];
#pragma unroll
                for (int cc = 0; cc < 9; ++cc)
#pragma unroll
                    for (int rr = 0; rr < 3; ++rr)
                        U[rr + 3 * cc] = Hh[tri(rr, cc)] * ga[0] + Hh[tri(rr + 3, cc)] * ga[1] + Hh[tri(rr + 6, cc)] * ga[2];
#pragma unroll
                for (int b = 0; b < 9; ++b) {
                    if (!(bmask >> b & 1)) continue;
                    const double* gb = r + 3 * (bp * 9 + b);
                    const double g0 = gb[0], g1 = gb[1], g2 = gb[2];
#pragma unroll
                    for (int ss = 0; ss < 3; ++ss)
#pragma unroll
                        for (int rr = 0; rr < 3; ++rr)
                            acc[b][rr + 3 * ss] += U[rr + 3 * ss] * g0 + U[rr + 3 * (ss + 3)] * g1 + U[rr + 3 * (ss + 6)] * g2;
                }
            }
            if (nc != c) { // the cell is complete: its sums go to the matrix
                const int cz = c & (Geo::BZ - 1), cy = (c >> Geo::zb) & (Geo::BY - 1), cx = c >> (Geo::zb + Geo::yb);
                const int ida = s_id[((cx + ai) * Geo::TY + (cy + aj)) * Geo::TZ + (cz + ak)];
#pragma unroll
                for (int b = 0; b < 9; ++b) {
                    if (!(bmask >> b & 1)) continue;
                    const int bj = b / 3, bk = b % 3;
                    const int idb = s_id[((cx + bp) * Geo::TY + (cy + bj)) * Geo::TZ + (cz + bk)];
                    if (ida >= 0 && idb >= 0) {
                        const int slot = (ai - bp + 2) * 25 + (aj - bj + 2) * 5 + (ak - bk + 2);
                        col[(size_t)ida * W + slot] = idb;
                        col[(size_t)idb * W + 124 - slot] = ida; // the mirrored entry (k_mirror reads it)
#pragma unroll
                        for (int q = 0; q < 9; ++q) atomicAdd(val + ((size_t)ida * 9 + q) * W + slot, dt2 * acc[b][q]);
                    }
#pragma unroll
                    for (int q = 0; q < 9; ++q) acc[b][q] = 0.0;
                }
            }
        }
        c = nc; p0 = np0; buf ^= 1;
    }
}

// ---- row-gather assembly (HOT_ASSEMBLE=rows) -----------------------------------------------------------------------------------------
// The scatter form above adds every (cell, node pair) block with 9 global REDs: 840 M atomics at C2, 3.8 GB of DRAM writes for a
// 0.97 GB matrix (ncu, round 1).  Here a WARP owns a block row: it walks the 27 cells whose particles reach its node, lane b < 27
// evaluates the 3x3 block towards stencil node b of every particle (sum over the particles of a cell in registers), the row is
// built in shared memory (slot = (a - b) + 2 per axis) and written ONCE, coalesced, together with its column ids - no atomics,
// DRAM writes = the matrix.  Per visit (particle, a): U = H~ . grad w_a by 27 lanes (3 DFMA each, H~ rows read through L1),
// broadcast through shared memory, then 27 DFMA per lane.
constexpr int AR_ROWS = 4; // rows (warps) per CTA
constexpr int AR_REC = 128; // doubles per particle record: H~ (45, packed upper triangle), grad w of the 27 stencil nodes (81), pad
struct ArWarp {
    double tile[9 * W]; // entry q of slot s at tile[q * W + s]
    double rec[2][AR_REC]; // the particle record of this and the next visit (cp.async)
    double ubuf[2][28];
    int col[W];
    int cfirst[28], cend[28]; // particle range of the 27 cells around the row's node
};
__device__ inline int find_slot_m(uint32_t pid, long n_pages, const uint32_t* __restrict__ pid_sorted, const int* __restrict__ slot_sorted)
{
    long lo = 0, hi = n_pages;
    while (lo < hi) {
        long mid = (lo + hi) >> 1;
        if (pid_sorted[mid] < pid) lo = mid + 1;
        else hi = mid;
    }
    return (lo < n_pages && pid_sorted[lo] == pid) ? slot_sorted[lo] : -1;
}
// page slot of page + (dx, dy, dz) pages, d in {-1, 0, 1}: nbr27[slot * 27 + (dx + 1) * 9 + (dy + 1) * 3 + dz + 1]
__global__ void k_nbr27(long n_pages, const uint32_t* __restrict__ page_id, const uint32_t* __restrict__ pid_sorted, const int* __restrict__ slot_sorted,
    int* __restrict__ nbr27)
{
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_pages * 27) return;
    const long slot = t / 27;
    const int q = (int)(t - slot * 27);
    const uint64_t off = (uint64_t)page_id[slot] << 12;
    const int x = (int)bit_pack(off, Geo::xmask) + Geo::BX * (q / 9 - 1), y = (int)bit_pack(off, Geo::ymask) + Geo::BY * ((q / 3) % 3 - 1),
              z = (int)bit_pack(off, Geo::zmask) + Geo::BZ * (q % 3 - 1);
    int r = -1;
    if (x >= 0 && y >= 0 && z >= 0 && x < 4096 && y < 4096 && z < 4096) r = find_slot_m((uint32_t)(linear_offset(x, y, z) >> 12), n_pages, pid_sorted, slot_sorted);
    nbr27[t] = r;
}
__global__ void k_group_of_slot(long n_groups, const int* __restrict__ group_slot, int* __restrict__ group_of_slot)
{
    const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n_groups) group_of_slot[group_slot[g]] = (int)g;
}
// per particle one 1 KB record: H~ (45 entries, packed upper triangle) and the weight gradients of its 27 stencil nodes (reference
// association, MpmGrid.h:272-291) - a warp stages it with two 16-byte cp.async per lane
__global__ void k_assemble_prep(long n, size_t ps, const double* __restrict__ X, const double* __restrict__ H, double dx, double one_over_dx,
    double* __restrict__ rec)
{
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * AR_REC) return;
    const long p = t / AR_REC;
    const int e = (int)(t - p * AR_REC);
    double v = 0.0;
    if (e < 45) v = H[(size_t)e * ps + p];
    else if (e < 126) {
        const int nd = (e - 45) / 3, c = (e - 45) - 3 * nd;
        const int ijk[3] = {nd / 9, (nd / 3) % 3, nd % 3};
        double w[3], dws[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            double xi, wa[3], dwa[3];
            const int bn = base_node_of(X[d * ps + p], one_over_dx, &xi);
            bspline_axis(xi - (double)bn, wa, dwa);
            w[d] = wa[ijk[d]];
            dws[d] = one_over_dx * dwa[ijk[d]];
        }
        v = c == 0 ? (dws[0] * w[1]) * w[2] : (c == 1 ? (w[0] * dws[1]) * w[2] : (w[0] * w[1]) * dws[2]);
    }
    rec[t] = v;
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(smem)), "l"(gmem) : "memory");
}
__global__ void __launch_bounds__(32 * AR_ROWS) k_assemble_rows(int n_nodes, const int* __restrict__ dof_slot, const int* __restrict__ nbr27,
    const int* __restrict__ group_of_slot, const int* __restrict__ cell_start, const double* __restrict__ rec,
    double dt2, const double* __restrict__ mass, const int* __restrict__ g_idx, int* __restrict__ col, double* __restrict__ val)
{
    extern __shared__ __align__(16) unsigned char ar_smem[];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int id = blockIdx.x * AR_ROWS + w;
    if (id >= n_nodes) return; // (whole warp)
    ArWarp& S = reinterpret_cast<ArWarp*>(ar_smem)[w];
    for (int t = lane; t < 9 * W; t += 32) S.tile[t] = 0.0;
    for (int t = lane; t < W; t += 32) S.col[t] = id;
    const int at = dof_slot[id];
    const int pslot = at / E, e = at - pslot * E;
    const int ex = e >> (Geo::yb + Geo::zb), ey = (e >> Geo::zb) & (Geo::BY - 1), ez = e & (Geo::BZ - 1);
    // lane = stencil index: as `a` the cell whose stencil node a is this row's node, as `b` the partner node of a particle's stencil
    const int ln = lane < 27 ? lane : 26; // lanes 27..31 shadow lane 26 (no divergence in the particle loop; they never flush)
    const int bi = ln / 9, bj = (ln / 3) % 3, bk = ln % 3;
    {
        int cx = ex - bi, cy = ey - bj, cz = ez - bk;
        const int px = cx < 0 ? -1 : 0, py = cy < 0 ? -1 : 0, pz = cz < 0 ? -1 : 0;
        cx -= px * Geo::BX; cy -= py * Geo::BY; cz -= pz * Geo::BZ;
        const int cslot = nbr27[(size_t)pslot * 27 + (px + 1) * 9 + (py + 1) * 3 + pz + 1];
        const int g = cslot >= 0 ? group_of_slot[cslot] : -1;
        int first = 0, end = 0;
        if (g >= 0) {
            const int ce = (cx * Geo::BY + cy) * Geo::BZ + cz;
            first = cell_start[(size_t)g * (E + 1) + ce];
            end = cell_start[(size_t)g * (E + 1) + ce + 1];
        }
        if (lane < 27) { S.cfirst[lane] = first; S.cend[lane] = end; }
    }
    // the column id of slot (a - b): node at coord - (a - b), looked up per (a, lane b) at flush time through the 27 neighbour pages
    const int rr = ln % 3, cc = ln / 3;
    const int t0 = tri(rr, cc), t1 = tri(rr + 3, cc), t2 = tri(rr + 6, cc);
    __syncwarp();
    // flat visit list over the non-empty cells; the record of the next visit is always in flight
    int a = 0;
    while (a < 27 && S.cend[a] <= S.cfirst[a]) ++a;
    int p = a < 27 ? S.cfirst[a] : 0;
    int ub = 0;
    auto stage = [&](int buf, int pp) {
        const double* src = rec + (size_t)pp * AR_REC + 4 * lane; // 2 x 16 bytes per lane and half record
        cp_async16(&S.rec[buf][4 * lane], src);
        cp_async16(&S.rec[buf][4 * lane + 2], src + 2);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (a < 27) stage(0, p);
    double acc[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) acc[q] = 0.0;
    while (a < 27) {
        // next visit
        int na = a, np = p + 1;
        if (np >= S.cend[a]) {
            ++na;
            while (na < 27 && S.cend[na] <= S.cfirst[na]) ++na;
            np = na < 27 ? S.cfirst[na] : 0;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        if (na < 27) stage(ub ^ 1, np);
        const double* R = S.rec[ub];
        const double g0 = R[45 + 3 * ln], g1 = R[46 + 3 * ln], g2 = R[47 + 3 * ln];
        const double ga0 = R[45 + 3 * a], ga1 = R[46 + 3 * a], ga2 = R[47 + 3 * a];
        S.ubuf[ub][ln] = fma(R[t2], ga2, fma(R[t1], ga1, R[t0] * ga0)); // U(rr, cc) = sum_c H~(rr + 3 c, cc) ga[c]  (lanes >= 27 rewrite entry 26 with the same value)
        __syncwarp();
        {
            const double2* u2 = reinterpret_cast<const double2*>(S.ubuf[ub]);
            double U[28];
#pragma unroll
            for (int t = 0; t < 14; ++t) {
                const double2 v = u2[t];
                U[2 * t] = v.x; U[2 * t + 1] = v.y;
            }
#pragma unroll
            for (int ss = 0; ss < 3; ++ss)
#pragma unroll
                for (int r3 = 0; r3 < 3; ++r3) acc[r3 + 3 * ss] = fma(U[r3 + 3 * (ss + 6)], g2, fma(U[r3 + 3 * (ss + 3)], g1, fma(U[r3 + 3 * ss], g0, acc[r3 + 3 * ss])));
        }
        ub ^= 1;
        if (na != a) {
            // cell a done -> into the row: slot (a - b) + 2 per axis; the column id of that slot is the node at coord - (a - b)
            const int ai = a / 9, aj = (a / 3) % 3, ak = a % 3;
            if (lane < 27) {
                const int dx_ = ai - bi, dy_ = aj - bj, dz_ = ak - bk;
                int nx = ex - dx_, ny = ey - dy_, nz = ez - dz_;
                const int qx = nx < 0 ? -1 : (nx >= Geo::BX ? 1 : 0), qy = ny < 0 ? -1 : (ny >= Geo::BY ? 1 : 0), qz = nz < 0 ? -1 : (nz >= Geo::BZ ? 1 : 0);
                nx -= qx * Geo::BX; ny -= qy * Geo::BY; nz -= qz * Geo::BZ;
                const int nslot = nbr27[(size_t)pslot * 27 + (qx + 1) * 9 + (qy + 1) * 3 + qz + 1];
                const int idb = nslot >= 0 ? g_idx[(size_t)nslot * E + (nx * Geo::BY + ny) * Geo::BZ + nz] : -1;
                if (idb >= 0) {
                    const int slot = (dx_ + 2) * 25 + (dy_ + 2) * 5 + dz_ + 2;
                    S.col[slot] = idb;
#pragma unroll
                    for (int q = 0; q < 9; ++q) S.tile[q * W + slot] += dt2 * acc[q];
                }
            }
#pragma unroll
            for (int q = 0; q < 9; ++q) acc[q] = 0.0;
            __syncwarp();
        }
        a = na; p = np;
    }
    // inertia term m_i I on the self slot (ImplicitSolver.h:485-493), then the row goes out once
    if (lane == 0) {
        const double m = mass[id];
        S.tile[0 * W + 62] += m; S.tile[4 * W + 62] += m; S.tile[8 * W + 62] += m;
    }
    __syncwarp();
    for (int t = lane; t < W; t += 32) col[(size_t)id * W + t] = S.col[t];
    double* vrow = val + (size_t)id * 9 * W;
    for (int t = lane; t < 9 * W; t += 32) vrow[t] = S.tile[t];
}

// lower triangle of the assembled matrix: block (j, i) = block (i, j)^T; slot s' < 62 of row j mirrors slot 124 - s' of row col[j][s']
__global__ void k_mirror(int n, const int* __restrict__ col, double* __restrict__ val)
{
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)n * 62) return;
    const int j = (int)(t / 62), s = (int)(t - (long)j * 62);
    const int i = col[(size_t)j * W + s];
    if (i == j) return; // empty slot (slot 62 is the only one that addresses the node itself)
    const double* src = val + (size_t)i * 9 * W + (124 - s);
    double* dst = val + (size_t)j * 9 * W + s;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) dst[(r + 3 * c) * W] = src[(c + 3 * r) * W];
}

__global__ void k_add_mass(int n, const double* __restrict__ mass, double* __restrict__ val)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double m = mass[i];
    double* v = val + (size_t)i * 9 * W + 62;
    v[0] += m; v[4 * W] += m; v[8 * W] += m;
}

__global__ void k_bc_of(int n_bc, const int* __restrict__ node, int* __restrict__ bc_of)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < n_bc) bc_of[node[b]] = b;
}

// BC projection of the system, ImplicitSolver.h:554-593
__global__ void k_bc_matrix(int n, const int* __restrict__ bc_of, const int* __restrict__ slip, const double* __restrict__ R,
    const double* __restrict__ Rinv, const int* __restrict__ col, double* __restrict__ val)
{
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)n * 125) return;
    const int i = (int)(t / 125), s = (int)(t - (long)i * 125);
    const int j = col[(size_t)i * W + s];
    const int bi = bc_of[i], bj = bc_of[j];
    if (bi < 0 && bj < 0) return;
    const bool iSlip = bi >= 0 && slip[bi] != 0, jSlip = bj >= 0 && slip[bj] != 0;
    double* v = val + (size_t)i * 9 * W + s; // entry q at v[q * W]
    if ((bi >= 0 && !iSlip) || (bj >= 0 && !jSlip)) {
        const bool self = (j == i) && s == 62;
#pragma unroll
        for (int q = 0; q < 9; ++q) v[q * W] = (self && (q == 0 || q == 4 || q == 8)) ? 1.0 : 0.0;
        return;
    }
    double a[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) a[q] = v[q * W];
    if (iSlip) {
        const double* Rm = R + 9 * (size_t)bi;
        double t9[9];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int r = 0; r < 3; ++r) t9[r + 3 * c] = Rm[r] * a[3 * c] + Rm[r + 3] * a[3 * c + 1] + Rm[r + 6] * a[3 * c + 2];
#pragma unroll
        for (int q = 0; q < 9; ++q) a[q] = t9[q];
    }
    if (jSlip) {
        const double* Rm = Rinv + 9 * (size_t)bj;
        double t9[9];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int r = 0; r < 3; ++r) t9[r + 3 * c] = a[r] * Rm[3 * c] + a[r + 3] * Rm[3 * c + 1] + a[r + 6] * Rm[3 * c + 2];
#pragma unroll
        for (int q = 0; q < 9; ++q) a[q] = t9[q];
    }
    if (iSlip) a[0] = a[3] = a[6] = 0.0;
    if (jSlip) a[0] = a[1] = a[2] = 0.0;
    if (i == j && s == 62) a[0] = 1.0;
#pragma unroll
    for (int q = 0; q < 9; ++q) v[q * W] = a[q];
}

__device__ __forceinline__ void inv3(const double* A, double* B)
{
    const double c0 = A[4] * A[8] - A[7] * A[5], c1 = A[7] * A[2] - A[1] * A[8], c2 = A[1] * A[5] - A[4] * A[2];
    const double det = A[0] * c0 + A[3] * c1 + A[6] * c2;
    B[0] = c0 / det; B[1] = c1 / det; B[2] = c2 / det;
    B[3] = (A[6] * A[5] - A[3] * A[8]) / det; B[4] = (A[0] * A[8] - A[6] * A[2]) / det; B[5] = (A[3] * A[2] - A[0] * A[5]) / det;
    B[6] = (A[3] * A[7] - A[6] * A[4]) / det; B[7] = (A[6] * A[1] - A[0] * A[7]) / det; B[8] = (A[0] * A[4] - A[3] * A[1]) / det;
}

// matrix-free diagonal (ImplicitSolver.h:605-665): D_a += dt^2 sum_{c,d} H~(r+3c, s+3d) gw_a,c gw_a,d
__global__ void k_diag_mf(long n, size_t ps, const double* __restrict__ X, const double* __restrict__ H, double dx, double one_over_dx,
    double dt2, const int* __restrict__ g_idx, const uint32_t* __restrict__ pid_sorted, const int* __restrict__ slot_sorted, long n_pages, double* __restrict__ D)
{
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * 27) return;
    const long s = t / 27;
    const int nd = (int)(t - 27 * s), i = nd / 9, j = (nd / 3) % 3, k = nd % 3;
    SplineEval sp;
    sp.eval(X, ps, (size_t)s, dx, one_over_dx, true);
    const double wi = sp.w[0][i], wj = sp.w[1][j], wk = sp.w[2][k];
    const double g[3] = {(one_over_dx * sp.dw[0][i] * wj) * wk, (wi * one_over_dx * sp.dw[1][j]) * wk, (wi * wj) * one_over_dx * sp.dw[2][k]};
    // node -> DOF id through the page table
    const uint64_t off = linear_offset(sp.base[0] + i, sp.base[1] + j, sp.base[2] + k);
    const uint32_t pid = (uint32_t)(off >> 12);
    long lo = 0, hi = n_pages;
    while (lo < hi) {
        long mid = (lo + hi) >> 1;
        if (pid_sorted[mid] < pid) lo = mid + 1;
        else hi = mid;
    }
    if (lo >= n_pages || pid_sorted[lo] != pid) return;
    const int id = g_idx[(size_t)slot_sorted[lo] * E + (int)((off & 0xfff) >> Geo::data_bits)];
    if (id < 0) return;
#pragma unroll
    for (int ss = 0; ss < 3; ++ss)
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) {
            double v = 0.0;
#pragma unroll
            for (int d = 0; d < 3; ++d)
#pragma unroll
                for (int c = 0; c < 3; ++c) v += H[(size_t)tri(rr + 3 * c, ss + 3 * d) * ps + s] * g[c] * g[d];
            atomicAdd(D + 9 * (size_t)id + rr + 3 * ss, dt2 * v);
        }
}
__global__ void k_add9(long n, const double* __restrict__ x, double* __restrict__ y)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] += x[i];
}
__global__ void k_diag_mf_init(int n, const double* __restrict__ mass, double* __restrict__ D)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 9 * n) return;
    const int q = t % 9;
    D[t] = (q == 0 || q == 4 || q == 8) ? mass[t / 9] : 0.0;
}
__global__ void k_diag_invert(int n, int Ainv, double* __restrict__ D)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a[9], b[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) a[q] = D[9 * (size_t)i + q];
    if (Ainv == 0) {
#pragma unroll
        for (int q = 0; q < 9; ++q) b[q] = (q == 0 || q == 4 || q == 8) ? 1.0 / a[q] : 0.0;
    }
    else if (a[0] == 0.0 && a[4] == 0.0 && a[8] == 0.0) { // a node only other ranks touch (partitioned run): keep it finite
#pragma unroll
        for (int q = 0; q < 9; ++q) b[q] = (q == 0 || q == 4 || q == 8) ? 1.0 : 0.0;
    }
    else inv3(a, b);
#pragma unroll
    for (int q = 0; q < 9; ++q) D[9 * (size_t)i + q] = b[q];
}
__global__ void k_block_diag_apply(int n, const double* __restrict__ D, const double* __restrict__ x, double* __restrict__ y)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* d = D + 9 * (size_t)i;
    const double x0 = x[3 * (size_t)i], x1 = x[3 * (size_t)i + 1], x2 = x[3 * (size_t)i + 2];
#pragma unroll
    for (int r = 0; r < 3; ++r) y[3 * (size_t)i + r] = d[r] * x0 + d[r + 3] * x1 + d[r + 6] * x2;
}

} // namespace

int fill_id2coord(Sim* s, int* coord_dev)
{
    if (s->num_nodes == 0) return 0;
    k_id2coord<<<nblk(s->num_nodes), TPB, 0, s->stream>>>(s->num_nodes, s->dof_slot.p, s->page_id.p, coord_dev);
    HOT_LAUNCHED(s);
    return 0;
}

int build_matrix(Sim* s, bool bcproject)
{
    if (s->world > 1 && !s->ghost_ring)
        return fail(s, "buildMatrix on a partitioned object needs the ghost ring: hot_set_ghost_ring(h, 1) before hot_sort_and_activate");
    int rc = ensure_hessian(s);
    if (rc) return rc;
    cudaStream_t st = s->stream;
    const int nn = s->num_nodes;
    if (nn <= 0) return fail(s, "buildMatrix: no grid nodes");
    if (s->levels.empty()) s->levels.push_back(new MGLevel);
    MGLevel& L = *s->levels[0];
    L.n = nn;
    HOT_CUDA(L.coord.reserve(3 * (size_t)nn));
    HOT_CUDA(L.col.reserve((size_t)nn * W));
    HOT_CUDA(L.val.reserve((size_t)nn * 9 * W));
    KTime t(s, KC_ASSEMBLE);
    rc = fill_id2coord(s, L.coord.p);
    if (rc) return rc;
    // default: cell-aggregated scatter of the upper triangle + mirror pass; HOT_ASSEMBLE=rows: the atomics-free row-gather form
    // (DRAM writes = the matrix, but bound by the shared-memory broadcast of U: 6.6 + 1.8 ms against 3.x ms at C2, profiles/r2_assemble.md)
    static const bool scatter_form = !(getenv("HOT_ASSEMBLE") && !strcmp(getenv("HOT_ASSEMBLE"), "rows"));
    static const bool round1_form = getenv("HOT_ASSEMBLE") && !strcmp(getenv("HOT_ASSEMBLE"), "scatter81"); // the round-1 kernel (+ upper-only)
    if (scatter_form) {
        k_matrix_init<<<nblk((long)nn * W), TPB, 0, st>>>(nn, s->world > 1 ? nullptr : s->mass_matrix.p, L.col.p, L.val.p);
        HOT_LAUNCHED(s);
        if (round1_form)
            k_assemble<<<(unsigned)s->n_groups, AS_THREADS, 0, st>>>(s->cell_start.p, s->group_slot.p, s->nbr8.p, s->P.stride, s->P.X.p, s->f_H.p,
                s->dx, 1.0 / s->dx, s->dt * s->dt, s->g_idx.p, L.col.p, L.val.p);
        else
            k_assemble_upper<<<(unsigned)s->n_groups, AU_THREADS, 0, st>>>(s->cell_start.p, s->group_slot.p, s->nbr8.p, s->P.stride, s->P.X.p, s->f_H.p,
                s->dx, 1.0 / s->dx, s->dt * s->dt, s->g_idx.p, L.col.p, L.val.p);
        HOT_LAUNCHED(s);
        k_mirror<<<nblk((long)nn * 62), TPB, 0, st>>>(nn, L.col.p, L.val.p);
        HOT_LAUNCHED(s);
    }
    else {
        const long NP = s->n_pages;
        HOT_CUDA(s->asm_nbr27.reserve(27 * (size_t)NP));
        HOT_CUDA(s->asm_group_of_slot.reserve((size_t)NP));
        HOT_CUDA(s->asm_wrec.reserve(AR_REC * (size_t)s->N));
        k_nbr27<<<nblk(NP * 27), TPB, 0, st>>>(NP, s->page_id.p, s->pid_sorted.p, s->slot_sorted.p, s->asm_nbr27.p);
        HOT_LAUNCHED(s);
        HOT_CUDA(cudaMemsetAsync(s->asm_group_of_slot.p, 0xff, (size_t)NP * sizeof(int), st));
        k_group_of_slot<<<nblk(s->n_groups), TPB, 0, st>>>(s->n_groups, s->group_slot.p, s->asm_group_of_slot.p);
        HOT_LAUNCHED(s);
        k_assemble_prep<<<nblk(s->N * AR_REC), TPB, 0, st>>>(s->N, s->P.stride, s->P.X.p, s->f_H.p, s->dx, 1.0 / s->dx, s->asm_wrec.p);
        HOT_LAUNCHED(s);
        HOT_FUNC_ATTR_ONCE(s, k_assemble_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(AR_ROWS * sizeof(ArWarp)));
        k_assemble_rows<<<(unsigned)((nn + AR_ROWS - 1) / AR_ROWS), 32 * AR_ROWS, AR_ROWS * sizeof(ArWarp), st>>>(nn, s->dof_slot.p, s->asm_nbr27.p,
            s->asm_group_of_slot.p, s->cell_start.p, s->asm_wrec.p, s->dt * s->dt, s->mass_matrix.p, s->g_idx.p, L.col.p, L.val.p);
        HOT_LAUNCHED(s);
    }
    if (s->world > 1) {
        // Partitioned object: the rows hold this rank's particles only.  Sum them over the holders of the shared pages (slot-wise:
        // a slot is a coordinate offset, the same on every rank), 128 of the 1152 doubles of a row per exchange; then the inertia
        // term, once; then the column ids from the node coordinates (another rank's particles couple nodes this rank's do not).
        if (!scatter_form) return fail(s, "buildMatrix on a partitioned object uses the scatter form (unset HOT_ASSEMBLE)");
        for (int q = 0; q < 9; ++q) {
            rc = dist_exchange_rows(s, L.val.p, 9 * W, q * W, W);
            if (rc) return rc;
        }
        k_add_mass<<<nblk(nn), TPB, 0, st>>>(nn, s->mass_matrix.p, L.val.p);
        HOT_LAUNCHED(s);
        rc = build_coord_map(s, L);
        if (rc) return rc;
        rc = columns_from_coords(s, L);
        if (rc) return rc;
    }
    if (bcproject && s->n_bc > 0) {
        HOT_CUDA(s->bc_of.reserve(nn));
        HOT_CUDA(cudaMemsetAsync(s->bc_of.p, 0xff, (size_t)nn * sizeof(int), st));
        k_bc_of<<<nblk(s->n_bc), TPB, 0, st>>>(s->n_bc, s->bc_node.p, s->bc_of.p);
        HOT_LAUNCHED(s);
        // like the reference, the system projection keys on CollisionNode::shouldRotate (the slip flags), not on the mode
        k_bc_matrix<<<nblk((long)nn * 125), TPB, 0, st>>>(nn, s->bc_of.p, s->bc_slip.p, s->bc_R.p, s->bc_Rinv.p, L.col.p, L.val.p);
        HOT_LAUNCHED(s);
    }
    s->matrix_bcproject = bcproject;
    s->matrix_built = true;
    s->mg_built = false;
    return 0;
}

// ImplicitSolverObjective::buildDiagonal: inverse diagonal blocks of M + dt^2 K without assembling the matrix
int build_diagonal_mf(Sim* s, int Ainv)
{
    int rc = ensure_hessian(s);
    if (rc) return rc;
    cudaStream_t st = s->stream;
    const int nn = s->num_nodes;
    HOT_CUDA(s->diag_mf.reserve(9 * (size_t)nn));
    KTime t(s, KC_ASSEMBLE);
    k_diag_mf_init<<<nblk(9 * (long)nn), TPB, 0, st>>>(nn, s->mass_matrix.p, s->diag_mf.p);
    HOT_LAUNCHED(s);
    double* dst = s->diag_mf.p;
    if (s->world > 1) { // own particles into a zeroed scratch, summed on the interface nodes, added to the mass term
        HOT_CUDA(s->scat_tmp.reserve(9 * (size_t)nn));
        HOT_CUDA(cudaMemsetAsync(s->scat_tmp.p, 0, 9 * (size_t)nn * sizeof(double), st));
        dst = s->scat_tmp.p;
    }
    const long np = s->p1 - s->p0;
    if (np > 0) {
        k_diag_mf<<<nblk(np * 27), TPB, 0, st>>>(np, s->P.stride, s->P.X.p + s->p0, s->f_H.p + s->p0, s->dx, 1.0 / s->dx, s->dt * s->dt, s->g_idx.p,
            s->pid_sorted.p, s->slot_sorted.p, s->n_pages, dst);
        HOT_LAUNCHED(s);
    }
    if (s->world > 1) {
        rc = dist_exchange_shared(s, dst, 9);
        if (rc) return rc;
        k_add9<<<nblk(9 * (long)nn), TPB, 0, st>>>(9 * (long)nn, dst, s->diag_mf.p);
        HOT_LAUNCHED(s);
    }
    k_diag_invert<<<nblk(nn), TPB, 0, st>>>(nn, Ainv, s->diag_mf.p);
    HOT_LAUNCHED(s);
    return 0;
}

int apply_block_diag(Sim* s, int n, const double* D9, const double* x, double* y)
{
    k_block_diag_apply<<<nblk(n), TPB, 0, s->stream>>>(n, D9, x, y);
    HOT_LAUNCHED(s);
    return 0;
}

} // namespace hot
