// a16-a20: Galerkin multigrid on the device - hierarchy build, block-row SpMV, restriction / prolongation,
// 8-colour block Gauss-Seidel, Jacobi / optimal-Jacobi / PCG smoothers, V-cycle.
//
// Reference: MultigridBuilder::build (Projects/multigrid/MultigridPreconditioner.h:553-703), linear_weight_template
// (:445-466), SquareMatrix::{multiply, buildDiagonal, buildCoarseMatrix, buildTransposeMatrix, comp}
// (SquareMatrix.h:39-46,301-324,477-487,526-607), MultigridOperator::{jacobi_smooth, optimal_jacobi_smooth, cg_smooth,
// gs_smooth, operator()} (MultigridPreconditioner.h:160-226,266-318,362-421), setup_logic / setup_parameters (:480-551).
//
// Re-design for the GPU
//  * rows are coordinate-addressed on every level (slot = 5^3 stencil offset), so the Galerkin triple product needs no
//    hash maps: thread (coarse row I, slot) sums  w_a w_b A_f[2I+a][slot(2(I-J)+a-b)]  over the <= 27 x 27 children pairs;
//  * the serial first-touch numbering of coarse nodes and the (colour, first-seen block, first-seen node) sweep order are
//    reproduced with stable radix sorts + run heads (same trick as the page list in sort.cu), so coarse DOF ids and the
//    Gauss-Seidel sequence are identical to the reference's;
//  * SpMV / GS are warp-per-row over the padded row layout documented in sim.h (coalesced 256-byte requests).
#include "sim.h"
#include "reduce.cuh"
#include <cub/cub.cuh>
#include <cooperative_groups.h>
#include <algorithm>
#include <cstdlib>

namespace hot {
namespace cg = cooperative_groups;
namespace {

constexpr int TPB = 256;
inline int nblk(long n) { return (int)((n + TPB - 1) / TPB); }
constexpr int W = MGLevel::W;
constexpr uint64_t NOKEY = ~0ull;

template <class F>
int with_tmp(Sim* s, F f)
{
    size_t bytes = 0;
    cudaError_t e = f((void*)nullptr, bytes);
    if (e != cudaSuccess) return cuda_fail(s, e, "cub size query");
    e = s->cub_tmp.reserve(bytes + 16);
    if (e != cudaSuccess) return cuda_fail(s, e, "cub temp alloc");
    e = f((void*)s->cub_tmp.p, bytes);
    if (e != cudaSuccess) return cuda_fail(s, e, "cub run");
    s->launches += 1;
    return 0;
}

__host__ __device__ inline uint64_t coord_key(int x, int y, int z) { return ((uint64_t)x << 24) | ((uint64_t)y << 12) | (uint64_t)z; }

__device__ inline int find_node(const uint64_t* __restrict__ keys, const int* __restrict__ ids, int n, uint64_t key)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (keys[mid] < key) lo = mid + 1;
        else hi = mid;
    }
    return (lo < n && keys[lo] == key) ? ids[lo] : -1;
}

__global__ void k_coord_keys(int n, const int* __restrict__ coord, uint64_t* __restrict__ key, int* __restrict__ id)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    key[i] = coord_key(coord[3 * i], coord[3 * i + 1], coord[3 * i + 2]);
    id[i] = i;
}

// candidates of the coarse node set, MultigridPreconditioner.h:630-668: position i*8 + q reproduces the serial visiting
// order (fine ids ascending, new_x outer / new_z inner); zero-weight candidates get NOKEY
__global__ void k_coarse_candidates(int n, const int* __restrict__ coord, uint64_t* __restrict__ key, int* __restrict__ pos)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * 8) return;
    const int i = t >> 3, q = t & 7;
    const int x = coord[3 * i], y = coord[3 * i + 1], z = coord[3 * i + 2];
    const int qx = q >> 2, qy = (q >> 1) & 1, qz = q & 1;
    const bool nonzero = (qx == 0 || (x & 1)) && (qy == 0 || (y & 1)) && (qz == 0 || (z & 1));
    key[t] = nonzero ? coord_key(x / 2 + qx, y / 2 + qy, z / 2 + qz) : NOKEY;
    pos[t] = t;
}
__global__ void k_head_flags64(long n, const uint64_t* __restrict__ k, int* __restrict__ flag)
{
    const long s = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    flag[s] = (k[s] != NOKEY) && ((s == 0) || (k[s] != k[s - 1]));
}
__global__ void k_iota(long n, int* __restrict__ v)
{
    const long s = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) v[s] = (int)s;
}
// order[c] = index into the ascending key list of the c-th coarse node (first-touch order)
__global__ void k_coarse_finish(int nc, const int* __restrict__ order, const uint64_t* __restrict__ key_asc, int* __restrict__ coord,
    int* __restrict__ id_sorted)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc) return;
    const int j = order[c];
    const uint64_t k = key_asc[j];
    coord[3 * c] = (int)(k >> 24);
    coord[3 * c + 1] = (int)((k >> 12) & 0xfff);
    coord[3 * c + 2] = (int)(k & 0xfff);
    id_sorted[j] = c;
}

// P rows: 8 slots (x/2 + {0,1})^3, trilinear weights w * I3; zero-weight slots alias slot 0's column (:640-647)
__global__ void k_build_P(int n, const int* __restrict__ coord, const uint64_t* __restrict__ ckeys, const int* __restrict__ cids, int nc,
    int* __restrict__ pcol, double* __restrict__ pw)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int x = coord[3 * i], y = coord[3 * i + 1], z = coord[3 * i + 2];
    int c0 = 0;
    for (int q = 0; q < 8; ++q) {
        const int qx = q >> 2, qy = (q >> 1) & 1, qz = q & 1;
        const bool nonzero = (qx == 0 || (x & 1)) && (qy == 0 || (y & 1)) && (qz == 0 || (z & 1));
        int c = c0;
        double w = 0.0;
        if (nonzero) {
            c = find_node(ckeys, cids, nc, coord_key(x / 2 + qx, y / 2 + qy, z / 2 + qz));
            w = ((x & 1) ? 0.5 : 1.0) * ((y & 1) ? 0.5 : 1.0) * ((z & 1) ? 0.5 : 1.0);
        }
        if (q == 0) c0 = c;
        pcol[(size_t)i * 8 + q] = c;
        pw[(size_t)i * 8 + q] = w;
    }
}
// R = P^T rows: children 2I + a, a in {-1,0,1}^3, weight prod (a == 0 ? 1 : 1/2); 27 entries padded to 32
__global__ void k_build_R(int nc, const int* __restrict__ ccoord, const uint64_t* __restrict__ fkeys, const int* __restrict__ fids, int nf,
    int* __restrict__ rcol, double* __restrict__ rw)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nc * 32) return;
    const int I = t >> 5, k = t & 31;
    int c = 0;
    double w = 0.0;
    if (k < 27) {
        const int ax = k / 9 - 1, ay = (k / 3) % 3 - 1, az = k % 3 - 1;
        const int x = 2 * ccoord[3 * I] + ax, y = 2 * ccoord[3 * I + 1] + ay, z = 2 * ccoord[3 * I + 2] + az;
        if (x >= 0 && y >= 0 && z >= 0) {
            const int f = find_node(fkeys, fids, nf, coord_key(x, y, z));
            if (f >= 0) {
                c = f;
                w = (ax ? 0.5 : 1.0) * (ay ? 0.5 : 1.0) * (az ? 0.5 : 1.0);
            }
        }
    }
    rcol[t] = c;
    rw[t] = w;
}

// Galerkin product A_c = R (A_f P), one CTA (128 threads = padded slots) per coarse row
__global__ void __launch_bounds__(W) k_galerkin(int nc, const int* __restrict__ ccoord, const uint64_t* __restrict__ ckeys,
    const int* __restrict__ cids, const int* __restrict__ rcol, const double* __restrict__ rw, const double* __restrict__ fval,
    int* __restrict__ ccol, double* __restrict__ cval)
{
    const int I = blockIdx.x, s = threadIdx.x;
    __shared__ int s_child[27];
    __shared__ double s_w[27];
    if (s < 27) {
        s_child[s] = rcol[(size_t)I * 32 + s];
        s_w[s] = rw[(size_t)I * 32 + s];
    }
    __syncthreads();
    double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int J = I;
    if (s < 125) {
        const int Dx = s / 25 - 2, Dy = (s / 5) % 5 - 2, Dz = s % 5 - 2; // coord_I - coord_J
        const int jx = ccoord[3 * I] - Dx, jy = ccoord[3 * I + 1] - Dy, jz = ccoord[3 * I + 2] - Dz;
        int found = -1;
        if (jx >= 0 && jy >= 0 && jz >= 0) found = find_node(ckeys, cids, nc, coord_key(jx, jy, jz));
        if (found >= 0) {
            J = found;
            for (int k = 0; k < 27; ++k) {
                const double wa = s_w[k];
                if (wa == 0.0) continue;
                const int ax = k / 9 - 1, ay = (k / 3) % 3 - 1, az = k % 3 - 1;
                const double* row = fval + (size_t)s_child[k] * 9 * W;
                for (int bx = -1; bx <= 1; ++bx) {
                    const int ex = 2 * Dx + ax - bx;
                    if (ex < -2 || ex > 2) continue;
                    for (int by = -1; by <= 1; ++by) {
                        const int ey = 2 * Dy + ay - by;
                        if (ey < -2 || ey > 2) continue;
                        for (int bz = -1; bz <= 1; ++bz) {
                            const int ez = 2 * Dz + az - bz;
                            if (ez < -2 || ez > 2) continue;
                            const double w = wa * (bx ? 0.5 : 1.0) * (by ? 0.5 : 1.0) * (bz ? 0.5 : 1.0);
                            const int fs = (ex + 2) * 25 + (ey + 2) * 5 + (ez + 2);
#pragma unroll
                            for (int q = 0; q < 9; ++q) acc[q] += w * row[q * W + fs];
                        }
                    }
                }
            }
        }
    }
    ccol[(size_t)I * W + s] = J;
#pragma unroll
    for (int q = 0; q < 9; ++q) cval[((size_t)I * 9 + q) * W + s] = acc[q];
}

__device__ __forceinline__ void inv3(const double* A, double* B)
{
    const double c0 = A[4] * A[8] - A[7] * A[5], c1 = A[7] * A[2] - A[1] * A[8], c2 = A[1] * A[5] - A[4] * A[2];
    const double det = A[0] * c0 + A[3] * c1 + A[6] * c2;
    B[0] = c0 / det; B[1] = c1 / det; B[2] = c2 / det;
    B[3] = (A[6] * A[5] - A[3] * A[8]) / det; B[4] = (A[0] * A[8] - A[6] * A[2]) / det; B[5] = (A[3] * A[2] - A[0] * A[5]) / det;
    B[6] = (A[3] * A[7] - A[6] * A[4]) / det; B[7] = (A[6] * A[1] - A[0] * A[7]) / det; B[8] = (A[0] * A[4] - A[3] * A[1]) / det;
}
// SquareMatrix::buildDiagonal: D_i = the self block (slot 62), its inverse by -Ainv
__global__ void k_level_diagonal(int n, int Ainv, const double* __restrict__ val, double* __restrict__ diag, double* __restrict__ dinv)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a[9], b[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) a[q] = val[((size_t)i * 9 + q) * W + 62];
    if (Ainv == 0) {
#pragma unroll
        for (int q = 0; q < 9; ++q) b[q] = (q == 0 || q == 4 || q == 8) ? 1.0 / a[q] : 0.0;
    }
    else inv3(a, b);
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        diag[9 * (size_t)i + q] = a[q];
        dinv[9 * (size_t)i + q] = b[q];
    }
}


// partitioned assembly (matrix.cu): after the rows have been summed over the ranks, a non-zero block may sit in a slot whose column
// this rank's own particles never set - the column of slot s is the node at coord_i - offset(s)
__global__ void k_cols_from_coords(int n, const int* __restrict__ coord, const uint64_t* __restrict__ keys, const int* __restrict__ ids,
    int* __restrict__ col, double* __restrict__ val)
{
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)n * W) return;
    const int i = (int)(t / W), s = (int)(t - (long)i * W);
    int j = i;
    if (s < 125 && s != 62) {
        double* v = val + (size_t)i * 9 * W + s;
        bool nz = false;
#pragma unroll
        for (int q = 0; q < 9; ++q) nz = nz || v[q * W] != 0.0;
        if (nz) {
            const int x = coord[3 * i] - (s / 25 - 2), y = coord[3 * i + 1] - ((s / 5) % 5 - 2), z = coord[3 * i + 2] - (s % 5 - 2);
            const int f = (x >= 0 && y >= 0 && z >= 0) ? find_node(keys, ids, n, coord_key(x, y, z)) : -1;
            if (f >= 0) j = f;
            else { // a column this rank does not hold: only on ghost pages, whose rows are never used (their values are taken over)
#pragma unroll
                for (int q = 0; q < 9; ++q) v[q * W] = 0.0;
            }
        }
    }
    col[t] = j;
}
// replicated coarse level of a partitioned object: node c = the c-th smallest coordinate key of the union over the ranks
__global__ void k_coarse_from_keys(int nc, const uint64_t* __restrict__ key_asc, int* __restrict__ coord, int* __restrict__ id_sorted)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc) return;
    const uint64_t k = key_asc[c];
    coord[3 * c] = (int)(k >> 24);
    coord[3 * c + 1] = (int)((k >> 12) & 0xfff);
    coord[3 * c + 2] = (int)(k & 0xfff);
    id_sorted[c] = c;
}
// R rows restricted to the fine nodes this rank counts: partial sums over the ranks add up to the whole restriction / Galerkin product
__global__ void k_mask_R(long n, const int* __restrict__ rcol, const unsigned char* __restrict__ own, double* __restrict__ rw)
{
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n && rw[t] != 0.0 && !own[rcol[t]]) rw[t] = 0.0;
}
__global__ void k_fill_u64(long n, uint64_t v, uint64_t* __restrict__ p)
{
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) p[t] = v;
}

// ---- Gauss-Seidel schedule ---------------------------------------------------------------------------------------------
__global__ void k_block_keys(int n, const int* __restrict__ coord, uint32_t* __restrict__ key, int* __restrict__ id)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    key[i] = ((uint32_t)(coord[3 * i] >> 2) << 20) | ((uint32_t)(coord[3 * i + 1] >> 2) << 10) | (uint32_t)(coord[3 * i + 2] >> 2);
    id[i] = i;
}
__global__ void k_head_flags32(int n, const uint32_t* __restrict__ k, int* __restrict__ flag)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) flag[s] = (s == 0) || (k[s] != k[s - 1]);
}
// key2 = colour << 32 | first (smallest) node id of the 4^3 block: sorting by it gives (colour, first-seen block) order
__global__ void k_sweep_keys(int n, const uint32_t* __restrict__ bkey_sorted, const int* __restrict__ id_sorted, const int* __restrict__ seg_incl,
    const int* __restrict__ head_pos, uint64_t* __restrict__ key2)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t b = bkey_sorted[p];
    const uint32_t color = (((b >> 20) & 1) << 2) | (((b >> 10) & 1) << 1) | (b & 1);
    const int first_id = id_sorted[head_pos[seg_incl[p] - 1]];
    key2[p] = ((uint64_t)color << 32) | (uint32_t)first_id;
}
__global__ void k_head_flags_key2(int n, const uint64_t* __restrict__ k, int* __restrict__ flag, int* __restrict__ color_count)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const bool head = (s == 0) || (k[s] != k[s - 1]);
    flag[s] = head;
    if (head) atomicAdd(color_count + (int)(k[s] >> 32), 1);
}
__global__ void k_rank(int n, const int* __restrict__ seq, int* __restrict__ rank)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) rank[seq[p]] = p;
}
// sweep rank of every stored neighbour, in the row layout of col: saves the dependent rank[col] gather in the sweeps
__global__ void k_colrank(long n_entries, const int* __restrict__ col, const int* __restrict__ rank, int* __restrict__ colrank)
{
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_entries) colrank[t] = rank[col[t]];
}

// ---- Gauss-Seidel row stream --------------------------------------------------------------------------------------------
// A sweep only uses the couplings of a row to nodes that come EARLIER (forward) or LATER (backward) in the sweep order - about
// half of the 125 slots each - but which slots those are depends on the rank of the neighbours, so the fixed-slot row has to be
// read whole (all of its sectors are touched).  The stream stores, per sweep position p and direction, only the entries that
// direction uses, compacted into chunks of 32 (code[32] + 9 x value[32], the coalesced layout of the fixed rows):
//   code >= 0   the neighbour is final when the row's block is swept (other block): its DOF id
//   code <  0   the neighbour belongs to the row's own 4^3 block: -(sweep-local index in that direction) - 1
// Padding entries carry code = GS_PAD (skipped) and zero blocks.  ~2.3 chunks per row and direction instead of 4.
constexpr int GS_PAD = -2147483647 - 1;
__global__ void k_gs_pblock(int n_blocks, const int* __restrict__ block_start, int* __restrict__ pblock)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    for (int p = block_start[b]; p < block_start[b + 1]; ++p) pblock[p] = b;
}
// pass 0 (fill == false): chunk counts per position; pass 1: the entries.  One warp per sweep position.
template <bool FILL>
__global__ void __launch_bounds__(TPB) k_gs_stream(int n, const int* __restrict__ seq, const int* __restrict__ rank, const int* __restrict__ pblock,
    const int* __restrict__ block_start, const int* __restrict__ col, const double* __restrict__ val, int* __restrict__ cntF,
    int* __restrict__ cntB, const int* __restrict__ offF, const int* __restrict__ offB, int* __restrict__ codeF, double* __restrict__ svalF,
    int* __restrict__ codeB, double* __restrict__ svalB)
{
    const int p = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (p >= n) return;
    const int i = seq[p], b = pblock[p], ps = block_start[b], pe = block_start[b + 1];
    int nF = 0, nB = 0;
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int t = 0; t < W / 32; ++t) {
        const int sl = lane + 32 * t;
        const int j = col[(size_t)i * W + sl];
        const int rj = (sl < 125 && j != i) ? rank[j] : p; // self / absent / padding slots: in neither set
        const bool f = rj < p, bk = rj > p;
        const unsigned mf = __ballot_sync(0xffffffffu, f), mb = __ballot_sync(0xffffffffu, bk);
        if (FILL) {
            if (f || (bk && codeB)) { // (codeB == nullptr: forward stream only - the residual update of the block-inverse form)
                const int e = f ? nF + __popc(mf & lt) : nB + __popc(mb & lt);
                const size_t c = (size_t)(f ? offF[p] : offB[p]) + (e >> 5);
                const int l = e & 31;
                const bool inblock = rj >= ps && rj < pe;
                const int code = !inblock ? j : -((f ? rj - ps : pe - 1 - rj) + 1);
                (f ? codeF : codeB)[c * 32 + l] = code;
                double* dst = (f ? svalF : svalB) + c * 9 * 32 + l;
#pragma unroll
                for (int q = 0; q < 9; ++q) dst[q * 32] = val[((size_t)i * 9 + q) * W + sl];
            }
        }
        nF += __popc(mf);
        nB += __popc(mb);
    }
    if (!FILL) {
        if (lane == 0) {
            cntF[p] = (nF + 31) >> 5;
            cntB[p] = (nB + 31) >> 5;
        }
        return;
    }
    // pad the last chunk of either direction
    for (int d = 0; d < (codeB ? 2 : 1); ++d) {
        const int cnt = d ? nB : nF, e = cnt + lane;
        if ((cnt & 31) != 0 && (e >> 5) == (cnt >> 5)) {
            const size_t c = (size_t)(d ? offB[p] : offF[p]) + (e >> 5);
            const int l = e & 31;
            (d ? codeB : codeF)[c * 32 + l] = GS_PAD;
            double* dst = (d ? svalB : svalF) + c * 9 * 32 + l;
#pragma unroll
            for (int q = 0; q < 9; ++q) dst[q * 32] = 0.0;
        }
    }
}

// ---- operators ---------------------------------------------------------------------------------------------------------
// SquareMatrix::multiply: one warp per block row.  mode 0: b = A x; mode 1: b -= A x
template <int MODE>
__global__ void __launch_bounds__(TPB) k_spmv(int n, const int* __restrict__ col, const double* __restrict__ val, const double* __restrict__ x,
    double* __restrict__ b)
{
    const int row = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (row >= n) return;
    const int* c = col + (size_t)row * W;
    const double* v = val + (size_t)row * 9 * W;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll
    for (int t = 0; t < W / 32; ++t) {
        const int s = lane + 32 * t;
        const int j = c[s];
        const double x0 = x[3 * (size_t)j], x1 = x[3 * (size_t)j + 1], x2 = x[3 * (size_t)j + 2];
        a0 += v[s] * x0 + v[3 * W + s] * x1 + v[6 * W + s] * x2;
        a1 += v[W + s] * x0 + v[4 * W + s] * x1 + v[7 * W + s] * x2;
        a2 += v[2 * W + s] * x0 + v[5 * W + s] * x1 + v[8 * W + s] * x2;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_down_sync(0xffffffffu, a0, o);
        a1 += __shfl_down_sync(0xffffffffu, a1, o);
        a2 += __shfl_down_sync(0xffffffffu, a2, o);
    }
    if (lane == 0) {
        double* o = b + 3 * (size_t)row;
        if (MODE == 0) { o[0] = a0; o[1] = a1; o[2] = a2; }
        else { o[0] -= a0; o[1] -= a1; o[2] -= a2; }
    }
}

// restriction: coarse_I = sum_k rw[I,k] fine[rcol[I,k]], one warp per coarse node
__global__ void __launch_bounds__(TPB) k_restrict(int nc, const int* __restrict__ rcol, const double* __restrict__ rw,
    const double* __restrict__ fine, double* __restrict__ coarse)
{
    const int I = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (I >= nc) return;
    const int j = rcol[(size_t)I * 32 + lane];
    const double w = rw[(size_t)I * 32 + lane];
    double a0 = w * fine[3 * (size_t)j], a1 = w * fine[3 * (size_t)j + 1], a2 = w * fine[3 * (size_t)j + 2];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_down_sync(0xffffffffu, a0, o);
        a1 += __shfl_down_sync(0xffffffffu, a1, o);
        a2 += __shfl_down_sync(0xffffffffu, a2, o);
    }
    if (lane == 0) {
        coarse[3 * (size_t)I] = a0; coarse[3 * (size_t)I + 1] = a1; coarse[3 * (size_t)I + 2] = a2;
    }
}
// prolongation: fine_i = sum_q pw[i,q] coarse[pcol[i,q]]
__global__ void k_prolong(int nf, const int* __restrict__ pcol, const double* __restrict__ pw, const double* __restrict__ coarse,
    double* __restrict__ fine)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf) return;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int j = pcol[(size_t)i * 8 + q];
        const double w = pw[(size_t)i * 8 + q];
        a0 += w * coarse[3 * (size_t)j]; a1 += w * coarse[3 * (size_t)j + 1]; a2 += w * coarse[3 * (size_t)j + 2];
    }
    fine[3 * (size_t)i] = a0; fine[3 * (size_t)i + 1] = a1; fine[3 * (size_t)i + 2] = a2;
}

// One colour phase of gs_smooth, two-phase form (the version the V-cycle uses).  One CTA per 4^3 block, the block's
// <= 64 nodes processed as two halves of <= 32 in sweep order; per half
//   phase A (all warps, bandwidth-bound): every row streams its 125 slots once.  Couplings to nodes that are already final
//     (earlier colours from HBM, the first half of this block from shared memory) are multiplied with their values and folded
//     into the right-hand side; couplings to earlier rows of the SAME half are parked in shared memory as a dense strictly
//     lower-triangular array of 3x3 blocks, stored per pivot column and per matrix entry so that phase B reads are
//     conflict-free (32*31/2 blocks = 36 KB, so several CTAs share an SM and overlap each other's phases);
//   phase B (one warp, on-chip): right-looking block forward substitution, one row per lane, pivot broadcast by shuffle.
// Same-colour blocks are >= 5 nodes apart, beyond the stencil radius 2, so this equals the reference's node-serial sweep
// (MultigridPreconditioner.h:276-310) up to the order of additions.  FWD: out_i = Dinv_i (rhs_i - sum_{rank j < rank i} A_ij out_j),
// optionally out_scaled_i = D_i out_i (the "hdu = D hdu" pass of :292-293 fused); BWD: rank j > rank i.
constexpr int GS_THREADS = 256;
constexpr int GS_HALF = 32;
constexpr int GS_PAIRS = GS_HALF * (GS_HALF - 1) / 2; // 496
__device__ __forceinline__ int gs_pair(int il, int kl) { return kl * (GS_HALF - 1) - kl * (kl - 1) / 2 + (il - kl - 1); } // il > kl

__device__ long long* g_gs_dbg = nullptr; // debug: clock64 stamps of one block (HOT_GS_DEBUG)
#define GS_STAMP(k)                                                                  \
    do {                                                                             \
        if (g_gs_dbg && b == g_gs_dbg[31] && threadIdx.x == 0) g_gs_dbg[k] = clock64(); \
    } while (0)

struct GSShared {
    double Lt[9][GS_PAIRS]; // entry q of the coupling (il, kl) at Lt[q][gs_pair(il, kl)]
    double s_rhs[GS_HALF][3];
    double s_x[2 * GS_HALF][3];
};

template <bool FWD, int THREADS, bool STREAM>
__device__ __forceinline__ void gs_block_body(GSShared& sh, int b, const int* __restrict__ block_start, const int* __restrict__ seq,
    const int* __restrict__ colrank, const int* __restrict__ col, const double* __restrict__ val, const double* __restrict__ dinv,
    const double* __restrict__ diag, const double* rhs, double* out, double* out_scaled, const int* __restrict__ soff,
    const int* __restrict__ scode, const double* __restrict__ sval)
{
    // STREAM: the rows come from the per-direction stream (k_gs_stream: only the entries this sweep direction uses,
    // chunks of 32), otherwise from the fixed 125-slot rows with the rank test per slot
    double (&Lt)[9][GS_PAIRS] = sh.Lt;
    double (&s_rhs)[GS_HALF][3] = sh.s_rhs;
    double (&s_x)[2 * GS_HALF][3] = sh.s_x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ps = block_start[b], pe = block_start[b + 1], nb = pe - ps;
    GS_STAMP(0);
    for (int h0 = 0; h0 < nb; h0 += GS_HALF) {
        const int hn = min(GS_HALF, nb - h0);
        __syncthreads(); // previous half fully consumed (Lt, s_rhs) and its s_x visible
        GS_STAMP(h0 ? 5 : 1);
        for (int e = tid; e < 9 * GS_PAIRS; e += THREADS) (&Lt[0][0])[e] = 0.0;
        __syncthreads();
        GS_STAMP(h0 ? 6 : 2);
        // sweep-local index of a node: FWD rank - ps, BWD pe - 1 - rank.
        // Rows are taken two at a time per warp so that the index loads of both rows, then the value loads of both rows, are
        // in flight together (fp64 dependent-issue latency on this part is ~45 cycles and a cold row costs two DRAM round
        // trips: memory-level parallelism per warp is what bounds this phase on the coarse levels).
        constexpr int NW = THREADS / 32;
        for (int il0 = warp; il0 < hn; il0 += 2 * NW) {
            int il_[2], i_[2], gl_[2], jj[2][W / 32], rr[2][W / 32], c0_[2], nc_[2];
            bool on[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                il_[u] = il0 + u * NW;
                on[u] = il_[u] < hn;
                gl_[u] = h0 + il_[u];
                const int p = FWD ? ps + gl_[u] : pe - 1 - gl_[u];
                i_[u] = on[u] ? seq[p] : 0;
                c0_[u] = (STREAM && on[u]) ? soff[p] : 0;
                nc_[u] = (STREAM && on[u]) ? soff[p + 1] - c0_[u] : 0;
            }
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int t = 0; t < W / 32; ++t) {
                    if (STREAM) {
                        // jj = code of the entry (>= 0: DOF id of a final neighbour, < 0: -(sweep-local index) - 1); rr unused
                        jj[u][t] = t < nc_[u] ? scode[((size_t)c0_[u] + t) * 32 + lane] : 0;
                        rr[u][t] = 0;
                    }
                    else {
                        jj[u][t] = on[u] ? col[(size_t)i_[u] * W + lane + 32 * t] : 0;
                        rr[u][t] = on[u] ? colrank[(size_t)i_[u] * W + lane + 32 * t] : 0;
                    }
                }
            double Dm[2][9];
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int q = 0; q < 9; ++q) Dm[u][q] = on[u] ? dinv[9 * (size_t)i_[u] + q] : 0.0;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (!on[u]) continue; // warp-uniform
                const int il = il_[u], gl = gl_[u], i = i_[u];
                const double* v = val + (size_t)i * 9 * W;
                double acc[W / 32][3];
#pragma unroll
                for (int t = 0; t < W / 32; ++t) {
                    acc[t][0] = acc[t][1] = acc[t][2] = 0.0;
                    const int sl = lane + 32 * t;
                    const int j = jj[u][t];
                    // kl < 0: final before this block; [0, gl): earlier in this block
                    const int kl = STREAM ? (j >= 0 ? -1 : -j - 1) : (FWD ? rr[u][t] - ps : pe - 1 - rr[u][t]);
                    if (STREAM ? (t < nc_[u] && j != GS_PAD) : kl < gl) {
                        // entry q of this lane's block: stream chunk (c0 + t), lane-major per entry, or slot sl of the fixed row
                        const double* vp = STREAM ? sval + ((size_t)c0_[u] + t) * 9 * 32 + lane : v + sl;
                        const int vs = STREAM ? 32 : W;
                        const double v0 = vp[0], v1 = vp[vs], v2 = vp[2 * vs], v3 = vp[3 * vs], v4 = vp[4 * vs], v5 = vp[5 * vs],
                                     v6 = vp[6 * vs], v7 = vp[7 * vs], v8 = vp[8 * vs];
                        if (kl < h0) {
                            double x0, x1, x2;
                            if (kl < 0) {
                                x0 = out[3 * (size_t)j]; x1 = out[3 * (size_t)j + 1]; x2 = out[3 * (size_t)j + 2];
                            }
                            else {
                                x0 = s_x[kl][0]; x1 = s_x[kl][1]; x2 = s_x[kl][2];
                            }
                            acc[t][0] = v0 * x0 + v3 * x1 + v6 * x2;
                            acc[t][1] = v1 * x0 + v4 * x1 + v7 * x2;
                            acc[t][2] = v2 * x0 + v5 * x1 + v8 * x2;
                        }
                        else { // same half, earlier row: park Dinv_i * A_ik for the on-chip substitution
                            const int e = gs_pair(il, kl - h0);
                            const double* D = Dm[u];
                            Lt[0][e] = D[0] * v0 + D[3] * v1 + D[6] * v2; Lt[1][e] = D[1] * v0 + D[4] * v1 + D[7] * v2;
                            Lt[2][e] = D[2] * v0 + D[5] * v1 + D[8] * v2; Lt[3][e] = D[0] * v3 + D[3] * v4 + D[6] * v5;
                            Lt[4][e] = D[1] * v3 + D[4] * v4 + D[7] * v5; Lt[5][e] = D[2] * v3 + D[5] * v4 + D[8] * v5;
                            Lt[6][e] = D[0] * v6 + D[3] * v7 + D[6] * v8; Lt[7][e] = D[1] * v6 + D[4] * v7 + D[7] * v8;
                            Lt[8][e] = D[2] * v6 + D[5] * v7 + D[8] * v8;
                        }
                    }
                }
                double a0 = (acc[0][0] + acc[1][0]) + (acc[2][0] + acc[3][0]);
                double a1 = (acc[0][1] + acc[1][1]) + (acc[2][1] + acc[3][1]);
                double a2 = (acc[0][2] + acc[1][2]) + (acc[2][2] + acc[3][2]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    a0 += __shfl_down_sync(0xffffffffu, a0, o);
                    a1 += __shfl_down_sync(0xffffffffu, a1, o);
                    a2 += __shfl_down_sync(0xffffffffu, a2, o);
                }
                if (lane == 0) { // r~ = Dinv (rhs - external couplings)
                    const double* D = Dm[u];
                    const double r0 = rhs[3 * (size_t)i] - a0, r1 = rhs[3 * (size_t)i + 1] - a1, r2 = rhs[3 * (size_t)i + 2] - a2;
                    s_rhs[il][0] = D[0] * r0 + D[3] * r1 + D[6] * r2;
                    s_rhs[il][1] = D[1] * r0 + D[4] * r1 + D[7] * r2;
                    s_rhs[il][2] = D[2] * r0 + D[5] * r1 + D[8] * r2;
                }
            }
        }
        __syncthreads();
        GS_STAMP(h0 ? 7 : 3);
        if (warp == 0) {
            // x_k = r~_k ; r~_i -= (Dinv_i A_ik) x_k for the later rows i: the pivot needs no multiply, the update is one
            // 3-deep FMA chain per component
            double r0 = 0.0, r1 = 0.0, r2 = 0.0;
            int node = -1;
            if (lane < hn) {
                node = seq[FWD ? ps + h0 + lane : pe - 1 - (h0 + lane)];
                r0 = s_rhs[lane][0]; r1 = s_rhs[lane][1]; r2 = s_rhs[lane][2];
            }
            // the coupling blocks of pivot k + 1 are fetched from shared memory while pivot k's update runs (they do not depend
            // on x), so a step's critical path is shuffle -> 3-deep FMA chain only
            double l[9], ln[9];
            const bool row = lane < hn;
#pragma unroll
            for (int q = 0; q < 9; ++q) l[q] = (row && lane > 0) ? Lt[q][gs_pair(lane, 0)] : 0.0;
            int en = lane - 1; // gs_pair(lane, k + 1) kept incrementally: gs_pair(l, k + 1) - gs_pair(l, k) = GS_HALF - 2 - k
            for (int k = 0; k < hn; ++k) {
                const bool nxt = row && lane > k + 1 && k + 1 < hn;
                en += GS_HALF - 2 - k;
#pragma unroll
                for (int q = 0; q < 9; ++q) ln[q] = nxt ? Lt[q][en] : 0.0; // (en is only in range where nxt holds)
                const double x0 = __shfl_sync(0xffffffffu, r0, k);
                const double x1 = __shfl_sync(0xffffffffu, r1, k);
                const double x2 = __shfl_sync(0xffffffffu, r2, k);
                if (lane > k) { // (l = 0 on lanes >= hn)
                    r0 = fma(-l[6], x2, fma(-l[3], x1, fma(-l[0], x0, r0)));
                    r1 = fma(-l[7], x2, fma(-l[4], x1, fma(-l[1], x0, r1)));
                    r2 = fma(-l[8], x2, fma(-l[5], x1, fma(-l[2], x0, r2)));
                }
#pragma unroll
                for (int q = 0; q < 9; ++q) l[q] = ln[q];
            }
            const double mx0 = r0, mx1 = r1, mx2 = r2;
            if (lane < hn) {
                s_x[h0 + lane][0] = mx0; s_x[h0 + lane][1] = mx1; s_x[h0 + lane][2] = mx2;
                out[3 * (size_t)node] = mx0; out[3 * (size_t)node + 1] = mx1; out[3 * (size_t)node + 2] = mx2;
                if (FWD && out_scaled) {
                    const double* d = diag + 9 * (size_t)node;
                    out_scaled[3 * (size_t)node] = d[0] * mx0 + d[3] * mx1 + d[6] * mx2;
                    out_scaled[3 * (size_t)node + 1] = d[1] * mx0 + d[4] * mx1 + d[7] * mx2;
                    out_scaled[3 * (size_t)node + 2] = d[2] * mx0 + d[5] * mx1 + d[8] * mx2;
                }
            }
            GS_STAMP(h0 ? 8 : 4);
        }
    }
}


struct GSArgs {
    int n;
    int cfb[9];
    const int *block_start, *seq, *colrank, *col;
    const double *val, *dinv, *diag;
    double *r, *hdu, *dhdu, *du, *u;
    int fuse_update; // u += du, r -= A du in the same launch (no BC projection needed on this level)
    const int *soff[2], *scode[2]; // per-direction row stream (null: fixed rows); [0] forward, [1] backward
    const double* sval[2];
    const int* pblock; // block of every sweep position
    int stream_update; // the tail uses r_new = L (hdu - du) from the forward stream
    // block-inverse form (k_gx_*): per-direction stream [external rows | inverse section] per half block, indexed by sweep
    // position in DIRECTION order (backward: n - 1 - p); the residual update reads the full forward row stream (soff[0] ...)
    const int* xoff[2];
    const double* xdata[2]; // chunk records of GX_REC doubles
};

// Tail of gs_smooth (u += du, r -= A du, MultigridPreconditioner.h:311-314) from the forward stream.  With (D + L) hdu = r and
// (D + U) du = D hdu (the two sweeps, L / U = couplings to earlier / later nodes of the sweep order):
//   r - A du = (D + L) hdu - L du - (D + U) du = L (hdu - du),
// so the new residual needs only the LOWER couplings - exactly the forward stream, about half of the bytes of the full rows.
// Holds when Dinv is the inverse of the diagonal block (-Ainv 1) and no BC projection sits between A du and r.
__device__ __forceinline__ void gs_stream_update_row(const GSArgs& a, int p, int lane)
{
    const int i = a.seq[p], ps = a.block_start[a.pblock[p]];
    const int c0 = a.soff[0][p], c1 = a.soff[0][p + 1];
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    // two chunks per trip, codes AND values requested together (padding entries are zero blocks in valid memory, so the value
    // loads need not wait for the code): the row's dependent chain is offsets -> (codes, values) -> gathers
    for (int c = c0; c < c1; c += 2) {
        int code[2];
        double v[2][9];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const bool on = c + u < c1;
            code[u] = on ? a.scode[0][(size_t)(c + u) * 32 + lane] : GS_PAD;
            const double* vp = a.sval[0] + (size_t)(c + u) * 9 * 32 + lane;
#pragma unroll
            for (int q = 0; q < 9; ++q) v[u][q] = on ? vp[q * 32] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (code[u] != GS_PAD) {
                const int j = code[u] >= 0 ? code[u] : a.seq[ps - code[u] - 1];
                const double x0 = a.hdu[3 * (size_t)j] - a.du[3 * (size_t)j], x1 = a.hdu[3 * (size_t)j + 1] - a.du[3 * (size_t)j + 1],
                             x2 = a.hdu[3 * (size_t)j + 2] - a.du[3 * (size_t)j + 2];
                a0 += v[u][0] * x0 + v[u][3] * x1 + v[u][6] * x2;
                a1 += v[u][1] * x0 + v[u][4] * x1 + v[u][7] * x2;
                a2 += v[u][2] * x0 + v[u][5] * x1 + v[u][8] * x2;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_down_sync(0xffffffffu, a0, o);
        a1 += __shfl_down_sync(0xffffffffu, a1, o);
        a2 += __shfl_down_sync(0xffffffffu, a2, o);
    }
    if (lane == 0) {
        const size_t o = 3 * (size_t)i;
        a.r[o] = a0; a.r[o + 1] = a1; a.r[o + 2] = a2;
        a.u[o] += a.du[o]; a.u[o + 1] += a.du[o + 1]; a.u[o + 2] += a.du[o + 2];
    }
}
__global__ void __launch_bounds__(TPB) k_gs_stream_update(GSArgs a)
{
    const int p = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (p < a.n) gs_stream_update_row(a, p, threadIdx.x & 31);
}

// One colour phase as its own launch (fallback when a cooperative launch is not possible)
template <bool FWD, bool STREAM>
__global__ void __launch_bounds__(GS_THREADS, 3) k_gs_block(int b0, GSArgs a)
{
    __shared__ GSShared sh;
    if (FWD)
        gs_block_body<true, GS_THREADS, STREAM>(sh, b0 + blockIdx.x, a.block_start, a.seq, a.colrank, a.col, a.val, a.dinv, a.diag, a.r, a.hdu, a.dhdu,
            a.soff[0], a.scode[0], a.sval[0]);
    else
        gs_block_body<false, GS_THREADS, STREAM>(sh, b0 + blockIdx.x, a.block_start, a.seq, a.colrank, a.col, a.val, a.dinv, a.diag, a.dhdu, a.du, nullptr,
            a.soff[1], a.scode[1], a.sval[1]);
}

// The whole symmetric sweep of gs_smooth in ONE cooperative launch: 8 forward colour phases, 8 backward ones, then
// u += du, r -= A du, separated by grid barriers instead of 17 dependent launches (on the coarse levels a phase is a
// handful of blocks, so launch gaps and ramp-up dominated the smoother).
template <int THREADS, bool STREAM>
__global__ void __launch_bounds__(THREADS) k_gs_sweep(GSArgs a)
{
    __shared__ GSShared sh;
    cg::grid_group grid = cg::this_grid();
    for (int c = 0; c < 8; ++c) {
        for (int b = a.cfb[c] + blockIdx.x; b < a.cfb[c + 1]; b += gridDim.x)
            gs_block_body<true, THREADS, STREAM>(sh, b, a.block_start, a.seq, a.colrank, a.col, a.val, a.dinv, a.diag, a.r, a.hdu, a.dhdu, a.soff[0], a.scode[0], a.sval[0]);
        grid.sync();
    }
    for (int c = 7; c >= 0; --c) {
        for (int b = a.cfb[c] + blockIdx.x; b < a.cfb[c + 1]; b += gridDim.x)
            gs_block_body<false, THREADS, STREAM>(sh, b, a.block_start, a.seq, a.colrank, a.col, a.val, a.dinv, a.diag, a.dhdu, a.du, nullptr, a.soff[1], a.scode[1], a.sval[1]);
        grid.sync();
    }
    if (!a.fuse_update) return;
    const int lane = threadIdx.x & 31;
    const long nwarps = (long)gridDim.x * (THREADS / 32);
    if (STREAM && a.stream_update) {
        for (long p = (long)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5); p < a.n; p += nwarps) gs_stream_update_row(a, (int)p, lane);
        return;
    }
    for (long row = (long)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5); row < a.n; row += nwarps) {
        const int* c = a.col + (size_t)row * W;
        const double* v = a.val + (size_t)row * 9 * W;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll
        for (int t = 0; t < W / 32; ++t) {
            const int s = lane + 32 * t;
            const int j = c[s];
            const double x0 = a.du[3 * (size_t)j], x1 = a.du[3 * (size_t)j + 1], x2 = a.du[3 * (size_t)j + 2];
            a0 += v[s] * x0 + v[3 * W + s] * x1 + v[6 * W + s] * x2;
            a1 += v[W + s] * x0 + v[4 * W + s] * x1 + v[7 * W + s] * x2;
            a2 += v[2 * W + s] * x0 + v[5 * W + s] * x1 + v[8 * W + s] * x2;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a0 += __shfl_down_sync(0xffffffffu, a0, o);
            a1 += __shfl_down_sync(0xffffffffu, a1, o);
            a2 += __shfl_down_sync(0xffffffffu, a2, o);
        }
        if (lane == 0) {
            const size_t o = 3 * (size_t)row;
            a.r[o] -= a0; a.r[o + 1] -= a1; a.r[o + 2] -= a2;
            a.u[o] += a.du[o]; a.u[o + 1] += a.du[o + 1]; a.u[o + 2] += a.du[o + 2];
        }
    }
}

// ---- ring form of a colour phase (default) ------------------------------------------------------------------------------
// Same arithmetic as gs_block_body<STREAM>, but the row stream of a block reaches shared memory through a ring of TMA bulk
// copies (cp.async.bulk onto an mbarrier per stage) instead of per-lane global loads, so that no warp ever waits for a DRAM
// round trip of stream data:
//   * a block's chunks are contiguous in the stream (chunk offsets are a prefix sum over the sweep positions), so a stage of
//     GS_KC chunks is two bulk copies (GS_KC x 2 304 B of values, GS_KC x 128 B of codes);
//   * warp 0 is the producer during phase A (one lane: wait for the slot's "empty" barrier, expect_tx, two bulk copies) and the
//     on-chip substitution warp during phase B; it runs up to NST stages ahead, ACROSS the half boundary, so the first stages of
//     the second half arrive while the first half's substitution runs;
//   * the other warps consume whole chunks (one warp per chunk, round robin): values and codes from shared memory, x of final
//     neighbours from L2, in-block couplings parked in Lt, one partial sum per chunk; the row sums are formed in chunk order by
//     the substitution warp (deterministic);
//   * per block the only dependent global round trips left are block_start -> (chunk offsets, node ids) -> stream / Dinv / rhs.
constexpr int GS_KC = 4;                               // chunks per ring stage
constexpr int GS_MAX_HALF_CHUNKS = GS_HALF * (W / 32); // a row has at most W / 32 chunks per direction
template <int NST>
struct __align__(16) GSRingShared {
    double Lt[9][GS_PAIRS];
    double val[NST][GS_KC][9][32];
    int code[NST][GS_KC][32];
    double s_x[2 * GS_HALF][3];
    double s_dinv[GS_HALF][9];
    double s_part[GS_MAX_HALF_CHUNKS][3];
    int s_off[2 * GS_HALF + 1];
    int s_seq[2 * GS_HALF];
    unsigned char s_crow[GS_MAX_HALF_CHUNKS];
    unsigned long long full[NST], empty[NST];
};

template <int THREADS, int NST>
__device__ __forceinline__ void gs_ring_init(GSRingShared<NST>& sh)
{
    if (threadIdx.x == 0)
        for (int k = 0; k < NST; ++k) {
            mbar_init(&sh.full[k], 1);
            mbar_init(&sh.empty[k], THREADS / 32 - 1);
        }
    __syncthreads();
}

// `gn`: ring stages this CTA has used so far (identical in every thread; slot = gn % NST, barrier parity from gn / NST)
template <bool FWD, int THREADS, int NST>
__device__ __forceinline__ void gs_block_ring(GSRingShared<NST>& sh, unsigned& gn, int b, const GSArgs& a)
{
    constexpr int NCW = THREADS / 32 - 1;
    const int* __restrict__ soff = a.soff[FWD ? 0 : 1];
    const int* __restrict__ scode = a.scode[FWD ? 0 : 1];
    const double* __restrict__ sval = a.sval[FWD ? 0 : 1];
    const double* rhs = FWD ? a.r : a.dhdu;
    double* out = FWD ? a.hdu : a.du;
    double* out_scaled = FWD ? a.dhdu : nullptr;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ps = a.block_start[b], pe = a.block_start[b + 1], nb = pe - ps;
    __syncthreads(); // (coop form: the previous block of this CTA is done with s_off / s_seq / s_x)
    for (int t = tid; t <= nb; t += THREADS) sh.s_off[t] = soff[ps + t];
    for (int t = tid; t < nb; t += THREADS) sh.s_seq[t] = a.seq[FWD ? ps + t : pe - 1 - t]; // node of sweep-local index t
    __syncthreads();
    // chunk ranges of the two halves.  FWD: half h covers positions ps + h0 ..; BWD: positions pe - 1 - h0 downwards
    const int nhalf = nb > GS_HALF ? 2 : 1;
    int cb[2], nch[2], nst[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int h0 = h * GS_HALF, hn = min(GS_HALF, nb - h0);
        if (h < nhalf) {
            cb[h] = FWD ? sh.s_off[h0] : sh.s_off[nb - h0 - hn];
            nch[h] = (FWD ? sh.s_off[h0 + hn] : sh.s_off[nb - h0]) - cb[h];
        }
        else { cb[h] = 0; nch[h] = 0; }
        nst[h] = (nch[h] + GS_KC - 1) / GS_KC;
    }
    const unsigned base = gn;
    const int total = nst[0] + nst[1];
    int pi = 0; // producer: stages of this block issued so far
    auto advance = [&](int limit) { // producer lane only
        for (; pi < limit; ++pi) {
            const int h = pi < nst[0] ? 0 : 1, n = pi - (h ? nst[0] : 0);
            const unsigned g = base + (unsigned)pi, slot = g % NST, use = g / NST;
            if (use > 0) mbar_wait(&sh.empty[slot], (use - 1) & 1);
            const int k = min(GS_KC, nch[h] - GS_KC * n);
            const size_t c0 = (size_t)cb[h] + (size_t)GS_KC * n;
            mbar_expect_tx(&sh.full[slot], (unsigned)k * (9 * 32 * 8 + 32 * 4));
            bulk_load(&sh.val[slot][0][0][0], sval + c0 * 9 * 32, (unsigned)k * 9 * 32 * 8, &sh.full[slot]);
            bulk_load(&sh.code[slot][0][0], scode + c0 * 32, (unsigned)k * 32 * 4, &sh.full[slot]);
        }
    };
    if (tid == 0) advance(min(total, NST));
    for (int h = 0; h < nhalf; ++h) {
        const int h0 = h * GS_HALF, hn = min(GS_HALF, nb - h0);
        const int hbase = h ? nst[0] : 0;
        __syncthreads(); // previous half's substitution is done: Lt free, its s_x visible
        for (int e = tid; e < 9 * GS_PAIRS; e += THREADS) (&sh.Lt[0][0])[e] = 0.0;
        for (int e = tid; e < hn * 9; e += THREADS) {
            const int il = e / 9;
            sh.s_dinv[il][e - 9 * il] = a.dinv[9 * (size_t)sh.s_seq[h0 + il] + (e - 9 * il)];
        }
        // position (relative to ps) of local row il, its chunk range relative to the half's first chunk
        int my_o0 = 0, my_o1 = 0;
        if (tid < hn) {
            const int p = FWD ? h0 + tid : nb - 1 - h0 - tid;
            my_o0 = sh.s_off[p] - cb[h];
            my_o1 = sh.s_off[p + 1] - cb[h];
            for (int c = my_o0; c < my_o1; ++c) sh.s_crow[c] = (unsigned char)tid;
        }
        double g0 = 0.0, g1 = 0.0, g2 = 0.0; // right-hand side of this lane's row (substitution warp)
        int node = -1;
        if (warp == 0 && lane < hn) {
            node = sh.s_seq[h0 + lane];
            g0 = rhs[3 * (size_t)node]; g1 = rhs[3 * (size_t)node + 1]; g2 = rhs[3 * (size_t)node + 2];
        }
        __syncthreads();
        if (warp == 0) {
            if (lane == 0) advance(min(total, hbase + nst[h] + NST));
            __syncwarp();
        }
        else {
            const int cw = warp - 1;
            for (int n = 0; n < nst[h]; ++n) {
                const unsigned g = base + (unsigned)(hbase + n), slot = g % NST;
                mbar_wait(&sh.full[slot], (g / NST) & 1);
                const int kn = min(GS_KC, nch[h] - GS_KC * n);
                for (int kk = 0; kk < kn; ++kk) {
                    const int ci = GS_KC * n + kk;
                    if (ci % NCW != cw) continue; // warp-uniform
                    const int il = sh.s_crow[ci];
                    const int code = sh.code[slot][kk][lane];
                    const double* vp = &sh.val[slot][kk][0][lane];
                    const double v0 = vp[0], v1 = vp[32], v2 = vp[64], v3 = vp[96], v4 = vp[128], v5 = vp[160], v6 = vp[192], v7 = vp[224],
                                 v8 = vp[256];
                    double x0 = 0.0, x1 = 0.0, x2 = 0.0;
                    if (code != GS_PAD) {
                        if (code >= 0) {
                            x0 = out[3 * (size_t)code]; x1 = out[3 * (size_t)code + 1]; x2 = out[3 * (size_t)code + 2];
                        }
                        else {
                            const int kl = -code - 1; // sweep-local index of an in-block neighbour (earlier in this direction)
                            if (kl < h0) {
                                x0 = sh.s_x[kl][0]; x1 = sh.s_x[kl][1]; x2 = sh.s_x[kl][2];
                            }
                            else { // same half: park Dinv_i * A_ik for the on-chip substitution
                                const int e = gs_pair(il, kl - h0);
                                const double* D = sh.s_dinv[il];
                                const double D0 = D[0], D1 = D[1], D2 = D[2], D3 = D[3], D4 = D[4], D5 = D[5], D6 = D[6], D7 = D[7], D8 = D[8];
                                sh.Lt[0][e] = D0 * v0 + D3 * v1 + D6 * v2; sh.Lt[1][e] = D1 * v0 + D4 * v1 + D7 * v2;
                                sh.Lt[2][e] = D2 * v0 + D5 * v1 + D8 * v2; sh.Lt[3][e] = D0 * v3 + D3 * v4 + D6 * v5;
                                sh.Lt[4][e] = D1 * v3 + D4 * v4 + D7 * v5; sh.Lt[5][e] = D2 * v3 + D5 * v4 + D8 * v5;
                                sh.Lt[6][e] = D0 * v6 + D3 * v7 + D6 * v8; sh.Lt[7][e] = D1 * v6 + D4 * v7 + D7 * v8;
                                sh.Lt[8][e] = D2 * v6 + D5 * v7 + D8 * v8;
                            }
                        }
                    }
                    double a0 = v0 * x0 + v3 * x1 + v6 * x2;
                    double a1 = v1 * x0 + v4 * x1 + v7 * x2;
                    double a2 = v2 * x0 + v5 * x1 + v8 * x2;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        a0 += __shfl_down_sync(0xffffffffu, a0, o);
                        a1 += __shfl_down_sync(0xffffffffu, a1, o);
                        a2 += __shfl_down_sync(0xffffffffu, a2, o);
                    }
                    if (lane == 0) { sh.s_part[ci][0] = a0; sh.s_part[ci][1] = a1; sh.s_part[ci][2] = a2; }
                }
                __syncwarp(); // every lane has read its slot entries
                if (lane == 0) mbar_arrive(&sh.empty[slot]);
            }
        }
        __syncthreads();
        if (warp == 0) {
            // r~ = Dinv (rhs - external couplings), the row's chunk partials added in chunk order
            double r0 = 0.0, r1 = 0.0, r2 = 0.0;
            const bool row = lane < hn;
            if (row) {
                double e0 = 0.0, e1 = 0.0, e2 = 0.0;
                for (int c = my_o0; c < my_o1; ++c) { e0 += sh.s_part[c][0]; e1 += sh.s_part[c][1]; e2 += sh.s_part[c][2]; }
                const double* D = sh.s_dinv[lane];
                const double q0 = g0 - e0, q1 = g1 - e1, q2 = g2 - e2;
                r0 = D[0] * q0 + D[3] * q1 + D[6] * q2;
                r1 = D[1] * q0 + D[4] * q1 + D[7] * q2;
                r2 = D[2] * q0 + D[5] * q1 + D[8] * q2;
            }
            // right-looking block forward substitution, one row per lane (as in gs_block_body)
            double l[9], ln[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) l[q] = (row && lane > 0) ? sh.Lt[q][gs_pair(lane, 0)] : 0.0;
            int en = lane - 1;
            for (int k = 0; k < hn; ++k) {
                const bool nxt = row && lane > k + 1 && k + 1 < hn;
                en += GS_HALF - 2 - k;
#pragma unroll
                for (int q = 0; q < 9; ++q) ln[q] = nxt ? sh.Lt[q][en] : 0.0;
                const double x0 = __shfl_sync(0xffffffffu, r0, k);
                const double x1 = __shfl_sync(0xffffffffu, r1, k);
                const double x2 = __shfl_sync(0xffffffffu, r2, k);
                if (lane > k) {
                    r0 = fma(-l[6], x2, fma(-l[3], x1, fma(-l[0], x0, r0)));
                    r1 = fma(-l[7], x2, fma(-l[4], x1, fma(-l[1], x0, r1)));
                    r2 = fma(-l[8], x2, fma(-l[5], x1, fma(-l[2], x0, r2)));
                }
#pragma unroll
                for (int q = 0; q < 9; ++q) l[q] = ln[q];
            }
            if (row) {
                sh.s_x[h0 + lane][0] = r0; sh.s_x[h0 + lane][1] = r1; sh.s_x[h0 + lane][2] = r2;
                out[3 * (size_t)node] = r0; out[3 * (size_t)node + 1] = r1; out[3 * (size_t)node + 2] = r2;
                if (FWD && out_scaled) {
                    const double* d = a.diag + 9 * (size_t)node;
                    out_scaled[3 * (size_t)node] = d[0] * r0 + d[3] * r1 + d[6] * r2;
                    out_scaled[3 * (size_t)node + 1] = d[1] * r0 + d[4] * r1 + d[7] * r2;
                    out_scaled[3 * (size_t)node + 2] = d[2] * r0 + d[5] * r1 + d[8] * r2;
                }
            }
        }
    }
    gn = base + (unsigned)total;
}

extern __shared__ __align__(16) unsigned char gs_dyn_smem[];
constexpr int GS_RING_NST = 3;      // per-phase launches: 3 CTAs of 8 warps per SM, 3 x 9.5 KB of stream in flight each
constexpr int GS_RING_NST_COOP = 6; // cooperative form (one 16-warp CTA per SM)

template <bool FWD>
__global__ void __launch_bounds__(GS_THREADS, 3) k_gs_block_ring(int b0, GSArgs a)
{
    GSRingShared<GS_RING_NST>& sh = *reinterpret_cast<GSRingShared<GS_RING_NST>*>(gs_dyn_smem);
    gs_ring_init<GS_THREADS, GS_RING_NST>(sh);
    unsigned gn = 0;
    gs_block_ring<FWD, GS_THREADS, GS_RING_NST>(sh, gn, b0 + blockIdx.x, a);
}

// the whole symmetric sweep in one cooperative launch (k_gs_sweep with the ring form of the colour phases)
template <int THREADS, int NST>
__global__ void __launch_bounds__(THREADS) k_gs_sweep_ring(GSArgs a)
{
    GSRingShared<NST>& sh = *reinterpret_cast<GSRingShared<NST>*>(gs_dyn_smem);
    cg::grid_group grid = cg::this_grid();
    gs_ring_init<THREADS, NST>(sh);
    unsigned gn = 0;
    for (int c = 0; c < 8; ++c) {
        for (int b = a.cfb[c] + blockIdx.x; b < a.cfb[c + 1]; b += gridDim.x) gs_block_ring<true, THREADS, NST>(sh, gn, b, a);
        grid.sync();
    }
    for (int c = 7; c >= 0; --c) {
        for (int b = a.cfb[c] + blockIdx.x; b < a.cfb[c + 1]; b += gridDim.x) gs_block_ring<false, THREADS, NST>(sh, gn, b, a);
        grid.sync();
    }
    if (!a.fuse_update) return;
    const int lane = threadIdx.x & 31;
    const long nwarps = (long)gridDim.x * (THREADS / 32);
    if (a.stream_update) {
        for (long p = (long)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5); p < a.n; p += nwarps) gs_stream_update_row(a, (int)p, lane);
        return;
    }
    for (long row = (long)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5); row < a.n; row += nwarps) {
        const int* c = a.col + (size_t)row * W;
        const double* v = a.val + (size_t)row * 9 * W;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll
        for (int t = 0; t < W / 32; ++t) {
            const int s = lane + 32 * t;
            const int j = c[s];
            const double x0 = a.du[3 * (size_t)j], x1 = a.du[3 * (size_t)j + 1], x2 = a.du[3 * (size_t)j + 2];
            a0 += v[s] * x0 + v[3 * W + s] * x1 + v[6 * W + s] * x2;
            a1 += v[W + s] * x0 + v[4 * W + s] * x1 + v[7 * W + s] * x2;
            a2 += v[2 * W + s] * x0 + v[5 * W + s] * x1 + v[8 * W + s] * x2;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a0 += __shfl_down_sync(0xffffffffu, a0, o);
            a1 += __shfl_down_sync(0xffffffffu, a1, o);
            a2 += __shfl_down_sync(0xffffffffu, a2, o);
        }
        if (lane == 0) {
            const size_t o = 3 * (size_t)row;
            a.r[o] -= a0; a.r[o + 1] -= a1; a.r[o + 2] -= a2;
            a.u[o] += a.du[o]; a.u[o + 1] += a.du[o + 1]; a.u[o + 2] += a.du[o + 2];
        }
    }
}

// ---- block-inverse form of the colour phases (default) --------------------------------------------------------------------
// The on-chip substitution of a half block is a serial chain: 32 steps x (shuffle + 3 dependent fp64 FMA) ~ 7-8 k cycles per half,
// ~16 k cycles (8 us) per block and colour phase that no amount of bandwidth removes (profiles/r1_gs_phase_stamps.txt) - more
// than the whole stream of a block takes at the SM's share of HBM bandwidth.  But the matrix of that chain does not depend on the
// iterate: with q = rhs - (couplings to nodes that are final before this half), the half's solution is
//     x = (I + L~)^-1 Dinv q,   L~_ik = Dinv_i A_ik  (i, k in the half, k earlier),
// so N = (I + L~)^-1 Dinv (lower block-triangular, diagonal blocks Dinv_i) is computed ONCE per hierarchy build
// (k_gx_inverse) and a sweep applies it as a dense product: x_i = Dinv_i q_i + sum_{k<i} N_ik q_k.  No dependent chain is left in
// a colour phase; it is a stream like the SpMV.  Rows i and hn-1-i of the strict lower triangle have i + (hn-1-i) = hn-1 <= 31
// blocks together, so the inverse of a half packs into ceil(hn/2) stream chunks (chunk m: lanes [0,m) row m, lanes [m,hn-1)
// row hn-1-m) - the same bytes as the in-half couplings it replaces.
// Stream of one block, contiguous, in processing order:  [ext rows of half 0][inverse of half 0][ext rows of half 1][inverse 1]
// ("ext" = couplings to other blocks, code = DOF id, and to the other half, code = -(local index) - 1).  A chunk is ONE record of
// GX_REC doubles (9 x 32 values, then 32 codes), so it moves with one TMA bulk copy.  Both directions index their offsets by
// sweep position in DIRECTION order (backward: n-1-p), so a block's chunks are consumed in ascending order.
// Kernel structure (k_gx_block, one launch per colour phase and CTA per block; measured alternatives in profiles/r2_gs_experiments.md):
//  * consumer warps own whole rows (row il -> warp il % NCW, inverse chunk m -> warp m % NCW): the x gather of a row's chunk k+1 is in
//    flight while chunk k is multiplied, lane partials are summed over the row's <= 4 chunks and reduced once per row;
//  * every consumer warp has its own ring of D chunk slots with a full / empty mbarrier pair per slot; lane w of warp 0 is the
//    producer of consumer warp w (empty barrier TESTED, expect_tx, one cp.async.bulk per chunk) in one converged loop over all
//    lanes - cp.async.bulk is a uniform-datapath instruction, lanes on their own paths would be issued one after the other - and
//    never joins the consumers' named barriers, so the stream keeps flowing through the section boundaries of a block;
//  * Dinv and D of the block's nodes are prefetched into shared memory once per block;
//  * everything up to that point reads data that is static between hierarchy builds, so colour phases 2..16 of a sweep are launched
//    as programmatic dependents of the phase before them (griddepcontrol.launch_dependents / .wait): their static prologue and
//    first TMA requests overlap the previous colour;
//  * k_gx_sweep is the cooperative single-launch form (A/B), k_gx_block_cl sweeps a block with a thread-block cluster (A/B).
// The residual update r - A du = L (hdu - du) reads the full forward row stream of k_gs_stream (k_gs_stream_update).
constexpr int GX_REC = 9 * 32 + 16;       // doubles per chunk record (2 432 bytes)
__device__ __forceinline__ const int* gx_codes(const double* rec) { return reinterpret_cast<const int*>(rec + 9 * 32); }
__device__ __forceinline__ int* gx_codes(double* rec) { return reinterpret_cast<int*>(rec + 9 * 32); }
template <int NCW, int D>
struct __align__(16) GXShared {
    double ring[NCW][D][GX_REC];
    double s_x[2 * GS_HALF][3];
    double s_q[GS_HALF][3];
    double s_dinv[2 * GS_HALF][9]; // Dinv and D of the block's nodes (sweep-local order), loaded once per block
    double s_diag[2 * GS_HALF][9];
    int s_off[2 * GS_HALF + 1];
    int s_seq[2 * GS_HALF];
    int p_off[2 * GS_HALF + 1]; // the producer warp's copy of its current block's chunk offsets (it runs ahead of the consumers)
    unsigned long long full[NCW][D], empty[NCW][D];
    unsigned long long cbar; // cluster form: barrier of the consumer warps of ALL CTAs of the cluster (count CL x NCW)
};

// pass 0 (FILL == false): chunk counts per direction-order position t (dir 0: p = t, dir 1: p = n-1-t); pass 1: the entries.
// One warp per position.  In-half couplings are left out (the inverse sections, filled by k_gx_inverse, stand for them).
template <bool FILL>
__global__ void __launch_bounds__(TPB) k_gx_stream(int n, int dir, const int* __restrict__ seq, const int* __restrict__ rank,
    const int* __restrict__ pblock, const int* __restrict__ block_start, const int* __restrict__ col, const double* __restrict__ val,
    int* __restrict__ cntX, const int* __restrict__ offX, double* __restrict__ dataX)
{
    const int t = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (t >= n) return;
    const int p = dir ? n - 1 - t : t;
    const int i = seq[p], b = pblock[p], ps = block_start[b], pe = block_start[b + 1];
    const int gl = dir ? pe - 1 - p : p - ps, half = gl >> 5, h0 = half << 5, hn = min(GS_HALF, pe - ps - h0);
    const bool last = gl == h0 + hn - 1;
    int nX = 0;
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int tt = 0; tt < W / 32; ++tt) {
        const int sl = lane + 32 * tt;
        const int j = col[(size_t)i * W + sl];
        const int rj = (sl < 125 && j != i) ? rank[j] : p; // self / absent / padding slots: in neither set
        const bool earlier = dir ? rj > p : rj < p;
        const bool inblock = rj >= ps && rj < pe;
        const int kl = dir ? pe - 1 - rj : rj - ps;
        const bool x = earlier && !(inblock && (kl >> 5) == half);
        const unsigned mx = __ballot_sync(0xffffffffu, x);
        if (FILL && x) {
            const int e = nX + __popc(mx & lt);
            double* rec = dataX + ((size_t)offX[t] + (e >> 5)) * GX_REC;
            gx_codes(rec)[e & 31] = inblock ? -(kl + 1) : j;
#pragma unroll
            for (int q = 0; q < 9; ++q) rec[q * 32 + (e & 31)] = val[((size_t)i * 9 + q) * W + sl];
        }
        nX += __popc(mx);
    }
    if (!FILL) {
        if (lane == 0) cntX[t] = ((nX + 31) >> 5) + (last ? (hn + 1) >> 1 : 0);
        return;
    }
    const int e = nX + lane; // pad the last chunk of the row
    if ((nX & 31) != 0 && (e >> 5) == (nX >> 5)) {
        double* rec = dataX + ((size_t)offX[t] + (e >> 5)) * GX_REC;
        gx_codes(rec)[e & 31] = GS_PAD;
#pragma unroll
        for (int q = 0; q < 9; ++q) rec[q * 32 + (e & 31)] = 0.0;
    }
}

// N = (I + L~)^-1 Dinv of one half block and direction, written into the half's inverse section of the stream.
// CTA = (block, half), 128 threads: the in-half couplings Dinv_i A_ik are gathered from the fixed rows into shared memory (same
// layout as gs_block_body's Lt), then thread (k, c) carries column (k, c) of N through a right-looking substitution: it starts
// as column c of Dinv_k in block row k, and block row i > k receives -L~_ij t_j from every finished row j.
constexpr int GXI_THREADS = 128;
struct GXInvShared {
    double Lt[9][GS_PAIRS];
    double t[3 * GS_HALF][3 * GS_HALF]; // t[3 i + r][column]
    double dinv[GS_HALF][9];
};
__global__ void __launch_bounds__(GXI_THREADS) k_gx_inverse(int dir, int n, const int* __restrict__ block_start, const int* __restrict__ seq,
    const int* __restrict__ colrank, const int* __restrict__ col, const double* __restrict__ val, const double* __restrict__ dinv,
    const int* __restrict__ xoff, double* __restrict__ xdata)
{
    GXInvShared& sh = *reinterpret_cast<GXInvShared*>(gs_dyn_smem);
    const int b = blockIdx.x >> 1, half = blockIdx.x & 1, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ps = block_start[b], pe = block_start[b + 1], nb = pe - ps, h0 = half * GS_HALF;
    if (h0 >= nb) return;
    const int hn = min(GS_HALF, nb - h0);
    for (int e = tid; e < 9 * GS_PAIRS; e += GXI_THREADS) (&sh.Lt[0][0])[e] = 0.0;
    for (int e = tid; e < hn * 9; e += GXI_THREADS) {
        const int il = e / 9, gl = h0 + il;
        sh.dinv[il][e - 9 * il] = dinv[9 * (size_t)seq[dir ? pe - 1 - gl : ps + gl] + (e - 9 * il)];
    }
    __syncthreads();
    for (int il = warp; il < hn; il += GXI_THREADS / 32) {
        const int gl = h0 + il, p = dir ? pe - 1 - gl : ps + gl, i = seq[p];
        const double* D = sh.dinv[il];
#pragma unroll
        for (int tt = 0; tt < W / 32; ++tt) {
            const int sl = lane + 32 * tt;
            const int j = col[(size_t)i * W + sl];
            const int rj = (sl < 125 && j != i) ? colrank[(size_t)i * W + sl] : p;
            const bool earlier = dir ? rj > p : rj < p;
            const bool inblock = rj >= ps && rj < pe;
            const int kl = dir ? pe - 1 - rj : rj - ps;
            if (earlier && inblock && (kl >> 5) == half) {
                const double* v = val + (size_t)i * 9 * W + sl;
                const double v0 = v[0], v1 = v[W], v2 = v[2 * W], v3 = v[3 * W], v4 = v[4 * W], v5 = v[5 * W], v6 = v[6 * W], v7 = v[7 * W],
                             v8 = v[8 * W];
                const int e = gs_pair(il, kl - h0);
                sh.Lt[0][e] = D[0] * v0 + D[3] * v1 + D[6] * v2; sh.Lt[1][e] = D[1] * v0 + D[4] * v1 + D[7] * v2;
                sh.Lt[2][e] = D[2] * v0 + D[5] * v1 + D[8] * v2; sh.Lt[3][e] = D[0] * v3 + D[3] * v4 + D[6] * v5;
                sh.Lt[4][e] = D[1] * v3 + D[4] * v4 + D[7] * v5; sh.Lt[5][e] = D[2] * v3 + D[5] * v4 + D[8] * v5;
                sh.Lt[6][e] = D[0] * v6 + D[3] * v7 + D[6] * v8; sh.Lt[7][e] = D[1] * v6 + D[4] * v7 + D[7] * v8;
                sh.Lt[8][e] = D[2] * v6 + D[5] * v7 + D[8] * v8;
            }
        }
    }
    // the inverse section of this half: the last ceil(hn / 2) chunks before the next half / block starts
    const int T0 = dir ? n - pe : ps, mch = (hn + 1) >> 1;
    double* M = xdata + ((size_t)xoff[T0 + h0 + hn] - mch) * GX_REC;
    for (int e = tid; e < mch * GX_REC; e += GXI_THREADS) M[e] = 0.0;
    __syncthreads();
    for (int e = tid; e < mch * 32; e += GXI_THREADS) gx_codes(M + (size_t)(e >> 5) * GX_REC)[e & 31] = GS_PAD; // (unused by the sweeps)
    const int tcol = tid, k = tcol / 3, c = tcol - 3 * k;
    const bool valid = tcol < 3 * hn;
    if (tcol < 3 * GS_HALF) {
        for (int i = 0; i < hn; ++i)
#pragma unroll
            for (int r = 0; r < 3; ++r) sh.t[3 * i + r][tcol] = (valid && i == k) ? sh.dinv[k][r + 3 * c] : 0.0;
        // (all columns of a warp run the same j range, so the Lt reads are broadcasts; rows j < k of a column are zero)
        const int j0 = (warp * 32) / 3;
        for (int j = j0; j < hn - 1; ++j) {
            const double t0 = sh.t[3 * j][tcol], t1 = sh.t[3 * j + 1][tcol], t2 = sh.t[3 * j + 2][tcol];
            for (int i = j + 1; i < hn; ++i) {
                const int e = gs_pair(i, j);
                sh.t[3 * i][tcol] -= sh.Lt[0][e] * t0 + sh.Lt[3][e] * t1 + sh.Lt[6][e] * t2;
                sh.t[3 * i + 1][tcol] -= sh.Lt[1][e] * t0 + sh.Lt[4][e] * t1 + sh.Lt[7][e] * t2;
                sh.t[3 * i + 2][tcol] -= sh.Lt[2][e] * t0 + sh.Lt[5][e] * t1 + sh.Lt[8][e] * t2;
            }
        }
        if (valid)
            for (int i = k + 1; i < hn; ++i) {
                const int m = i < mch ? i : hn - 1 - i, l = i < mch ? k : m + k;
#pragma unroll
                for (int r = 0; r < 3; ++r) M[(size_t)m * GX_REC + (r + 3 * c) * 32 + l] = sh.t[3 * i + r][tcol];
            }
    }
}

// producer lane w (warp 0): the chunks of block b that consumer warp w will ask for, in its order of consumption - per half the
// ext chunks of rows w, w + NCW, ..., then the inverse chunks w, w + NCW, ... - from its `first`-th on, at most up to its
// `limit`-th.  pn = chunks this lane has requested so far (slot = pn % D).  Returns how many of the block's chunks it passed.
// The producer lanes share a warp: they walk one converged loop and only TEST their slot's empty barrier in it (a lane that
// blocked on its consumer would hold up the requests of all the others).
template <int NCW>
struct GXIter {
    const int* off;
    int nb, w, stride, h0, hn, mch, M0, il, c, c1, m; // w: first row / inverse chunk of this worker, stride: number of workers
    bool inv;
    __device__ __forceinline__ void half()
    {
        hn = min(GS_HALF, nb - h0); mch = (hn + 1) >> 1; M0 = off[h0 + hn] - mch;
        il = w - stride; c = c1 = 0; inv = false; m = w;
    }
    __device__ __forceinline__ void init(const int* off_, int nb_, int w_, int stride_ = NCW)
    {
        off = off_; nb = nb_; w = w_; stride = stride_; h0 = 0;
        half();
    }
    __device__ __forceinline__ int next() // absolute chunk index, -1 when the block is exhausted
    {
        for (;;) {
            if (!inv) {
                if (c < c1) return c++;
                il += stride;
                if (il < hn) { c = off[h0 + il]; c1 = il == hn - 1 ? M0 : off[h0 + il + 1]; }
                else inv = true;
            }
            else {
                if (m < mch) { const int r = M0 + m; m += stride; return r; }
                h0 += GS_HALF;
                if (h0 >= nb) return -1;
                half();
            }
        }
    }
};
// (whole warp 0 calls this; lanes >= NCW only help loading the block's offsets into shared memory)
template <bool FWD, int NCW, int D, int CL = 1>
__device__ __forceinline__ int gx_produce(GXShared<NCW, D>& sh, unsigned& pn, int w, int b, const GSArgs& a, int first, int limit, int crank = 0)
{
    const int d = FWD ? 0 : 1;
    const int ps = a.block_start[b], pe = a.block_start[b + 1], nb = pe - ps;
    __syncwarp();
    for (int t = w; t <= nb; t += 32) sh.p_off[t] = a.xoff[d][(FWD ? ps : a.n - pe) + t];
    __syncwarp();
    GXIter<NCW> it;
    int k = 0, c = -1;
    if (w < NCW) {
        it.init(sh.p_off, nb, w * CL + crank, NCW * CL);
        c = it.next();
    }
    while (__any_sync(0xffffffffu, c >= 0 && k < limit)) {
#pragma unroll 1
        for (int rep = 0; rep < D; ++rep)
            if (c >= 0 && k < limit) {
                if (k < first) { ++k; c = it.next(); } // requested before (ahead of a grid barrier)
                else {
                    const unsigned slot = pn % D, use = pn / D;
                    if (use == 0 || mbar_test(&sh.empty[w][slot], (use - 1) & 1)) {
                        mbar_expect_tx(&sh.full[w][slot], (unsigned)(GX_REC * 8));
                        bulk_load(&sh.ring[w][slot][0], a.xdata[d] + (size_t)c * GX_REC, (unsigned)(GX_REC * 8), &sh.full[w][slot]);
                        ++pn; ++k; c = it.next();
                    }
                }
            }
    }
    return k;
}

template <int NCT>
__device__ __forceinline__ void gx_cbar() { asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory"); }

// ---- cluster form: one 4^3 block is swept by the CL CTAs of a thread-block cluster (rows il -> CTA il % CL), so CL SMs stream it.
// q and x are replicated into every CTA's shared memory through distributed shared memory; the consumer warps of the whole cluster
// meet on an mbarrier per CTA (every warp arrives remotely on all CL of them, release / acquire at cluster scope).
__device__ __forceinline__ unsigned gx_cluster_rank()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned gx_mapa(const void* p, unsigned cta)
{
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr(p)), "r"(cta));
    return r;
}
__device__ __forceinline__ void gx_st_cluster(unsigned addr, double v) { asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }
__device__ __forceinline__ void gx_cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int NCT, int CL>
__device__ __forceinline__ void gx_csync(unsigned long long* cbar, unsigned& phase, int lane)
{
    if (CL == 1) { gx_cbar<NCT>(); return; }
    __syncwarp();
    if (lane == 0)
        for (unsigned c = 0; c < (unsigned)CL; ++c)
            asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(gx_mapa(cbar, c)) : "memory");
    unsigned ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok)
                     : "r"(smem_addr(cbar)), "r"(phase & 1u)
                     : "memory");
    } while (!ok);
    ++phase;
}

#ifdef HOT_GX_PROFILE // clock stamps of one watched block (HOT_GS_DEBUG=n), compiled in only for profiling builds
#define GX_STAMP(k)                                  \
    do {                                             \
        if (dbg_st) g_gs_dbg[k] = clock64();         \
    } while (0)
#else
#define GX_STAMP(k) do { } while (0)
#endif
// consumer warps (all warps but warp 0): one block of one colour phase.  wn: chunks this warp has consumed so far.
// A warp owns whole rows: the (<= 4) ext chunks of a row are acquired together, their x gathers are in flight together, the lane
// partials of all of them are summed in registers and reduced ONCE per row; q = rhs - sum goes straight to shared memory.
template <bool FWD, int THREADS, int D, int CL = 1>
__device__ __forceinline__ void gx_consume(GXShared<THREADS / 32 - 1, D>& sh, unsigned& wn, int b, const GSArgs& a, int crank = 0)
{
    constexpr int NCW = THREADS / 32 - 1, NCT = NCW * 32;
    unsigned cphase = 0; // (cluster form: one block per kernel, so the barrier phases start at 0)
    const int d = FWD ? 0 : 1;
    const double* rhs = FWD ? a.r : a.dhdu;
    double* out = FWD ? a.hdu : a.du;
    const int ct = threadIdx.x - 32, lane = threadIdx.x & 31, cw = (threadIdx.x >> 5) - 1;
    const int ps = a.block_start[b], pe = a.block_start[b + 1], nb = pe - ps, T0 = FWD ? ps : a.n - pe;
#ifdef HOT_GX_PROFILE
    const bool dbg_st = g_gs_dbg && b == g_gs_dbg[31] && threadIdx.x == 32;
#endif
    GX_STAMP(0);
    gx_cbar<NCT>(); // the previous block of this CTA is done with the block tables
    for (int t = ct; t <= nb; t += NCT) sh.s_off[t] = a.xoff[d][T0 + t];
    for (int t = ct; t < nb; t += NCT) sh.s_seq[t] = a.seq[FWD ? ps + t : pe - 1 - t]; // node of sweep-local index t
    gx_cbar<NCT>();
    GX_STAMP(1);
    // Dinv (and D for the fused dhdu = D hdu of the forward sweep) of every node of the block: requested now, first used after the
    // first barrier of the half, so the round trip hides behind the ext rows
    for (int e = ct; e < nb * 9; e += NCT) {
        const int gl = e / 9, q = e - 9 * gl;
        sh.s_dinv[gl][q] = a.dinv[9 * (size_t)sh.s_seq[gl] + q];
        if (FWD) sh.s_diag[gl][q] = a.diag[9 * (size_t)sh.s_seq[gl] + q];
    }
    // Everything above (and the whole producer warp) touches only data that is static between hierarchy builds.  A colour phase
    // launched as a programmatic dependent of the previous one (smooth_gs) gets this far while the previous colour still runs;
    // the iterate (rhs, out) is read below.  Without a programmatic dependency this returns at once.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    auto acquire = [&](unsigned k) -> const double* { // the k-th record from now on
        const unsigned g = wn + k, slot = g % D;
        mbar_wait(&sh.full[cw][slot], (g / D) & 1);
        return &sh.ring[cw][slot][0];
    };
    auto release = [&](int n) {
        __syncwarp(); // every lane has read its entries of the records
        if (lane == 0)
            for (int k = 0; k < n; ++k) mbar_arrive(&sh.empty[cw][(wn + (unsigned)k) % D]);
        wn += (unsigned)n;
    };
    const int nhalf = nb > GS_HALF ? 2 : 1;
    for (int h = 0; h < nhalf; ++h) {
        const int h0 = h * GS_HALF, hn = min(GS_HALF, nb - h0), mch = (hn + 1) >> 1;
        const int M0 = sh.s_off[h0 + hn] - mch; // first inverse chunk of the half (absolute chunk index)
        GX_STAMP(h ? 7 : 2);
        for (int il = cw * CL + crank; il < hn; il += NCW * CL) {
            const int n = (il == hn - 1 ? M0 : sh.s_off[h0 + il + 1]) - sh.s_off[h0 + il]; // ext chunks of the row, <= W / 32
            double g0 = 0.0, g1 = 0.0, g2 = 0.0;
            if (lane == 0) { // (a block-wide prefetch of the right-hand sides into shared memory costs a barrier: measured slower)
                const int node = sh.s_seq[h0 + il];
                g0 = rhs[3 * (size_t)node]; g1 = rhs[3 * (size_t)node + 1]; g2 = rhs[3 * (size_t)node + 2];
            }
            // software pipeline over the row's chunks: the x gather of chunk k + 1 is in flight while chunk k is multiplied; a
            // slot is released as soon as its values are in registers, so the producer stays D - 2 chunks ahead
            auto fetch_x = [&](const double* rec, double& x0, double& x1, double& x2) {
                const int code = gx_codes(rec)[lane];
                x0 = x1 = x2 = 0.0;
                if (code != GS_PAD) {
                    if (code >= 0) {
                        x0 = out[3 * (size_t)code]; x1 = out[3 * (size_t)code + 1]; x2 = out[3 * (size_t)code + 2];
                    }
                    else { // a node of the first half (final)
                        const int kl = -code - 1;
                        x0 = sh.s_x[kl][0]; x1 = sh.s_x[kl][1]; x2 = sh.s_x[kl][2];
                    }
                }
            };
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, b0 = 0.0, b1 = 0.0, b2 = 0.0; // two accumulator sets (alternating chunks)
            const double* rc = nullptr;
            double x0 = 0.0, x1 = 0.0, x2 = 0.0;
            if (n > 0) {
                rc = acquire(0u);
                fetch_x(rc, x0, x1, x2);
            }
            for (int k = 0; k < n; ++k) {
                const double* rn = nullptr;
                double y0 = 0.0, y1 = 0.0, y2 = 0.0;
                if (k + 1 < n) {
                    rn = acquire(1u);
                    fetch_x(rn, y0, y1, y2);
                }
                const double* vp = rc + lane;
                const double v0 = vp[0], v1 = vp[32], v2 = vp[64], v3 = vp[96], v4 = vp[128], v5 = vp[160], v6 = vp[192], v7 = vp[224], v8 = vp[256];
                release(1);
                if (k & 1) {
                    b0 += v0 * x0 + v3 * x1 + v6 * x2; b1 += v1 * x0 + v4 * x1 + v7 * x2; b2 += v2 * x0 + v5 * x1 + v8 * x2;
                }
                else {
                    a0 += v0 * x0 + v3 * x1 + v6 * x2; a1 += v1 * x0 + v4 * x1 + v7 * x2; a2 += v2 * x0 + v5 * x1 + v8 * x2;
                }
                rc = rn; x0 = y0; x1 = y1; x2 = y2;
            }
            a0 += b0; a1 += b1; a2 += b2;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a0 += __shfl_down_sync(0xffffffffu, a0, o);
                a1 += __shfl_down_sync(0xffffffffu, a1, o);
                a2 += __shfl_down_sync(0xffffffffu, a2, o);
            }
            if (lane == 0) { // q = rhs - external couplings
                if (CL == 1) { sh.s_q[il][0] = g0 - a0; sh.s_q[il][1] = g1 - a1; sh.s_q[il][2] = g2 - a2; }
                else
                    for (unsigned c = 0; c < (unsigned)CL; ++c) {
                        const unsigned qa = gx_mapa(&sh.s_q[il][0], c);
                        gx_st_cluster(qa, g0 - a0); gx_st_cluster(qa + 8, g1 - a1); gx_st_cluster(qa + 16, g2 - a2);
                    }
            }
        }
        GX_STAMP(h ? 8 : 3);
        gx_csync<NCT, CL>(&sh.cbar, cphase, lane); // q of the half complete (and s_dinv)
        GX_STAMP(h ? 9 : 4);
        for (int m = cw * CL + crank; m < mch; m += NCW * CL) {
            const double* r0 = acquire(0u);
            const int rowB = hn - 1 - m;
            const bool bvalid = rowB != m;
            const bool inA = lane < m, inB = bvalid && lane >= m && lane < hn - 1;
            const int k = inA ? lane : lane - m;
            const double* vp = r0 + lane;
            const double v0 = vp[0], v1 = vp[32], v2 = vp[64], v3 = vp[96], v4 = vp[128], v5 = vp[160], v6 = vp[192], v7 = vp[224], v8 = vp[256];
            double q0 = 0.0, q1 = 0.0, q2 = 0.0;
            if (inA || inB) { q0 = sh.s_q[k][0]; q1 = sh.s_q[k][1]; q2 = sh.s_q[k][2]; }
            const double p0 = v0 * q0 + v3 * q1 + v6 * q2;
            const double p1 = v1 * q0 + v4 * q1 + v7 * q2;
            const double p2 = v2 * q0 + v5 * q1 + v8 * q2;
            double a0 = inA ? p0 : 0.0, a1 = inA ? p1 : 0.0, a2 = inA ? p2 : 0.0;
            double b0 = inB ? p0 : 0.0, b1 = inB ? p1 : 0.0, b2 = inB ? p2 : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); a2 += __shfl_xor_sync(0xffffffffu, a2, o);
                b0 += __shfl_xor_sync(0xffffffffu, b0, o); b1 += __shfl_xor_sync(0xffffffffu, b1, o); b2 += __shfl_xor_sync(0xffffffffu, b2, o);
            }
            if (lane == 0 || (lane == 1 && bvalid)) { // lane 0 finishes row m, lane 1 row hn-1-m: x = Dinv q + N q
                const int il = lane == 0 ? m : rowB;
                const double* Di = sh.s_dinv[h0 + il];
                const double u0 = sh.s_q[il][0], u1 = sh.s_q[il][1], u2 = sh.s_q[il][2];
                const double x0 = (Di[0] * u0 + Di[3] * u1 + Di[6] * u2) + (lane == 0 ? a0 : b0);
                const double x1 = (Di[1] * u0 + Di[4] * u1 + Di[7] * u2) + (lane == 0 ? a1 : b1);
                const double x2 = (Di[2] * u0 + Di[5] * u1 + Di[8] * u2) + (lane == 0 ? a2 : b2);
                const int node = sh.s_seq[h0 + il];
                if (CL == 1) { sh.s_x[h0 + il][0] = x0; sh.s_x[h0 + il][1] = x1; sh.s_x[h0 + il][2] = x2; }
                else
                    for (unsigned c = 0; c < (unsigned)CL; ++c) {
                        const unsigned xa = gx_mapa(&sh.s_x[h0 + il][0], c);
                        gx_st_cluster(xa, x0); gx_st_cluster(xa + 8, x1); gx_st_cluster(xa + 16, x2);
                    }
                out[3 * (size_t)node] = x0; out[3 * (size_t)node + 1] = x1; out[3 * (size_t)node + 2] = x2;
                if (FWD) { // dhdu = D hdu (the "hdu = D hdu" pass of :292-293 fused)
                    const double* dg = sh.s_diag[h0 + il];
                    a.dhdu[3 * (size_t)node] = dg[0] * x0 + dg[3] * x1 + dg[6] * x2;
                    a.dhdu[3 * (size_t)node + 1] = dg[1] * x0 + dg[4] * x1 + dg[7] * x2;
                    a.dhdu[3 * (size_t)node + 2] = dg[2] * x0 + dg[5] * x1 + dg[8] * x2;
                }
            }
            release(1);
        }
        GX_STAMP(h ? 10 : 5);
        gx_csync<NCT, CL>(&sh.cbar, cphase, lane); // s_x of this half visible to the next half's cross-half reads; s_q / s_dinv free
        GX_STAMP(h ? 11 : 6);
    }
}

template <int THREADS, int D>
__device__ __forceinline__ void gx_init(GXShared<THREADS / 32 - 1, D>& sh)
{
    constexpr int NCW = THREADS / 32 - 1;
    for (int e = threadIdx.x; e < NCW * D; e += THREADS) {
        mbar_init(&sh.full[e / D][e % D], 1);
        mbar_init(&sh.empty[e / D][e % D], 1);
    }
    __syncthreads();
}

// one colour phase per launch: (256 threads, depth 4) runs 3 CTAs per SM, (512, 4) one 15-consumer-warp CTA per SM - a phase lasts as
// long as its largest block, so more warps per block beat more blocks per SM unless a colour has many waves of blocks
template <bool FWD, int THREADS, int D, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_gx_block(int b0, GSArgs a)
{
    constexpr int NCW = THREADS / 32 - 1;
    GXShared<NCW, D>& sh = *reinterpret_cast<GXShared<NCW, D>*>(gs_dyn_smem);
    gx_init<THREADS, D>(sh);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); // the next colour's CTAs may start their static prologue as slots free up
    const int b = b0 + blockIdx.x;
    if (threadIdx.x < 32) {
        unsigned pn = 0;
        gx_produce<FWD, NCW, D>(sh, pn, (int)threadIdx.x, b, a, 0, 1 << 30);
    }
    else {
        unsigned wn = 0;
        gx_consume<FWD, THREADS, D>(sh, wn, b, a);
    }
}

// cluster form of one colour phase: block b0 + blockIdx.x / CL is swept by the CL CTAs of a cluster
template <bool FWD, int THREADS, int D, int MINB, int CL>
__global__ void __launch_bounds__(THREADS, MINB) k_gx_block_cl(int b0, GSArgs a)
{
    constexpr int NCW = THREADS / 32 - 1;
    GXShared<NCW, D>& sh = *reinterpret_cast<GXShared<NCW, D>*>(gs_dyn_smem);
    if (threadIdx.x == 0) mbar_init(&sh.cbar, CL * NCW);
    gx_init<THREADS, D>(sh);
    gx_cluster_sync_all(); // every CTA's barriers exist before anybody arrives on them remotely
    const int b = b0 + (int)blockIdx.x / CL, crank = (int)gx_cluster_rank();
    if (threadIdx.x < 32) {
        unsigned pn = 0;
        gx_produce<FWD, NCW, D, CL>(sh, pn, (int)threadIdx.x, b, a, 0, 1 << 30, crank);
    }
    else {
        unsigned wn = 0;
        gx_consume<FWD, THREADS, D, CL>(sh, wn, b, a, crank);
    }
}

// the whole symmetric sweep in one cooperative launch.  The producer lanes run ahead of the grid barriers: the stream is static
// data, so the first chunks of the next colour's first block are requested BEFORE the barrier that ends the current colour.
template <int THREADS, int D, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_gx_sweep(GSArgs a)
{
    constexpr int NCW = THREADS / 32 - 1;
    GXShared<NCW, D>& sh = *reinterpret_cast<GXShared<NCW, D>*>(gs_dyn_smem);
    cg::grid_group grid = cg::this_grid();
    gx_init<THREADS, D>(sh);
    unsigned cnt = 0; // consumer warp: chunks consumed; producer lane: chunks requested
    int pre = 0;      // producer lane: chunks of the NEXT block already requested
    for (int phase = 0; phase < 16; ++phase) {
        const bool fwd = phase < 8;
        const int c = fwd ? phase : 15 - phase;
        if (threadIdx.x >= 32) {
            for (int b = a.cfb[c] + blockIdx.x; b < a.cfb[c + 1]; b += gridDim.x) {
                if (fwd) gx_consume<true, THREADS, D>(sh, cnt, b, a);
                else gx_consume<false, THREADS, D>(sh, cnt, b, a);
            }
        }
        else {
            {
                const int w = (int)threadIdx.x;
                for (int b = a.cfb[c] + blockIdx.x; b < a.cfb[c + 1]; b += gridDim.x) {
                    if (fwd) gx_produce<true, NCW, D>(sh, cnt, w, b, a, pre, 1 << 30);
                    else gx_produce<false, NCW, D>(sh, cnt, w, b, a, pre, 1 << 30);
                    pre = 0;
                }
                if (phase < 15) { // first block of the next phase
                    const bool nf = phase + 1 < 8;
                    const int nc = nf ? phase + 1 : 15 - (phase + 1);
                    const int b = a.cfb[nc] + blockIdx.x;
                    if (b < a.cfb[nc + 1]) {
                        const int mine = nf ? gx_produce<true, NCW, D>(sh, cnt, w, b, a, 0, D) : gx_produce<false, NCW, D>(sh, cnt, w, b, a, 0, D);
                        pre = min(mine, D);
                    }
                }
            }
            __syncwarp();
        }
        grid.sync();
    }
    if (!a.fuse_update) return;
    const int lane = threadIdx.x & 31;
    const long nwarps = (long)gridDim.x * (THREADS / 32);
    if (a.stream_update) {
        for (long p = (long)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5); p < a.n; p += nwarps) gs_stream_update_row(a, (int)p, lane);
        return;
    }
    for (long row = (long)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5); row < a.n; row += nwarps) {
        const int* c = a.col + (size_t)row * W;
        const double* v = a.val + (size_t)row * 9 * W;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll
        for (int t = 0; t < W / 32; ++t) {
            const int s = lane + 32 * t;
            const int j = c[s];
            const double x0 = a.du[3 * (size_t)j], x1 = a.du[3 * (size_t)j + 1], x2 = a.du[3 * (size_t)j + 2];
            a0 += v[s] * x0 + v[3 * W + s] * x1 + v[6 * W + s] * x2;
            a1 += v[W + s] * x0 + v[4 * W + s] * x1 + v[7 * W + s] * x2;
            a2 += v[2 * W + s] * x0 + v[5 * W + s] * x1 + v[8 * W + s] * x2;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a0 += __shfl_down_sync(0xffffffffu, a0, o);
            a1 += __shfl_down_sync(0xffffffffu, a1, o);
            a2 += __shfl_down_sync(0xffffffffu, a2, o);
        }
        if (lane == 0) {
            const size_t o = 3 * (size_t)row;
            a.r[o] -= a0; a.r[o + 1] -= a1; a.r[o + 2] -= a2;
            a.u[o] += a.du[o]; a.u[o + 1] += a.du[o + 1]; a.u[o + 2] += a.du[o + 2];
        }
    }
}

// u += du and r -= A du in one pass over A (the tail of gs_smooth, :311-314); one warp per row
__global__ void __launch_bounds__(TPB) k_spmv_update(int n, const int* __restrict__ col, const double* __restrict__ val,
    const double* __restrict__ du, double* __restrict__ u, double* __restrict__ r)
{
    const int row = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (row >= n) return;
    const int* c = col + (size_t)row * W;
    const double* v = val + (size_t)row * 9 * W;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll
    for (int t = 0; t < W / 32; ++t) {
        const int s = lane + 32 * t;
        const int j = c[s];
        const double x0 = du[3 * (size_t)j], x1 = du[3 * (size_t)j + 1], x2 = du[3 * (size_t)j + 2];
        a0 += v[s] * x0 + v[3 * W + s] * x1 + v[6 * W + s] * x2;
        a1 += v[W + s] * x0 + v[4 * W + s] * x1 + v[7 * W + s] * x2;
        a2 += v[2 * W + s] * x0 + v[5 * W + s] * x1 + v[8 * W + s] * x2;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_down_sync(0xffffffffu, a0, o);
        a1 += __shfl_down_sync(0xffffffffu, a1, o);
        a2 += __shfl_down_sync(0xffffffffu, a2, o);
    }
    if (lane == 0) {
        const size_t o = 3 * (size_t)row;
        r[o] -= a0; r[o + 1] -= a1; r[o + 2] -= a2;
        u[o] += du[o]; u[o + 1] += du[o + 1]; u[o + 2] += du[o + 2];
    }
}

__global__ void k_block_diag_inplace(int n, const double* __restrict__ D, double* __restrict__ x)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* d = D + 9 * (size_t)i;
    const double x0 = x[3 * (size_t)i], x1 = x[3 * (size_t)i + 1], x2 = x[3 * (size_t)i + 2];
#pragma unroll
    for (int r = 0; r < 3; ++r) x[3 * (size_t)i + r] = d[r] * x0 + d[r + 3] * x1 + d[r + 6] * x2;
}
__global__ void k_block_diag(int n, const double* __restrict__ D, const double* __restrict__ x, double* __restrict__ y, double scale)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* d = D + 9 * (size_t)i;
    const double x0 = x[3 * (size_t)i], x1 = x[3 * (size_t)i + 1], x2 = x[3 * (size_t)i + 2];
#pragma unroll
    for (int r = 0; r < 3; ++r) y[3 * (size_t)i + r] = scale * (d[r] * x0 + d[r + 3] * x1 + d[r + 6] * x2);
}

// ---- BLAS-1 (K16) --------------------------------------------------------------------------------------------------------
__global__ void k_axpy(long n, double a, const double* __restrict__ x, double* __restrict__ y)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] += a * x[i];
}
__global__ void k_axpy_dev(long n, const double* __restrict__ num, const double* __restrict__ den, double sign, const double* __restrict__ x,
    double* __restrict__ y)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] += sign * (num[0] / den[0]) * x[i];
}
__global__ void k_xpay_dev(long n, const double* __restrict__ x, const double* __restrict__ num, const double* __restrict__ den,
    double* __restrict__ y)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = x[i] + (num[0] / den[0]) * y[i];
}
__global__ void k_xpay(long n, const double* __restrict__ x, double b, double* __restrict__ y)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = x[i] + b * y[i];
}
__global__ void k_abs(long n, double* __restrict__ y)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = fabs(y[i]);
}
// +-1 start vector of estimate2norm: a fixed hash of the entry index (the reference seeds rand() with the wall clock)
__global__ void k_sign_pattern(long n, double* __restrict__ y)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = ((uint32_t)((size_t)i * 2654435761u) >> 16) & 1u ? 1.0 : -1.0;
}
__global__ void k_scale(long n, double a, double* __restrict__ y)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] *= a;
}

int reserve_level_vectors(Sim* s, MGLevel& L)
{
    const size_t m = 3 * (size_t)L.n;
    HOT_CUDA(L.residual.reserve(m));
    HOT_CUDA(L.initial_residual.reserve(m));
    HOT_CUDA(L.sol.reserve(m));
    HOT_CUDA(L.du.reserve(m));
    HOT_CUDA(L.dAu.reserve(m));
    HOT_CUDA(L.tmp.reserve(m));
    return 0;
}

// (colour, first-seen block, node id) sweep order of markColors, MultigridPreconditioner.h:582-605
int gx_cluster_cfg()
{
    static const int v = getenv("HOT_GX_CLUSTER") ? atoi(getenv("HOT_GX_CLUSTER")) : 0;
    return (v == 2 || v == 4) ? v : 0;
}
// A/B switch, default on: Gauss-Seidel colour phases in block-inverse form (k_gx_*); 0: the substitution forms (k_gs_*)
bool gs_inv_mode()
{
    static const bool on = !(getenv("HOT_GS_INV") && atoi(getenv("HOT_GS_INV")) == 0);
    return on;
}
// streams of the block-inverse form: counts -> offsets -> entries -> inverse sections (once per hierarchy build)
int build_gx_streams(Sim* s, MGLevel& L)
{
    cudaStream_t st = s->stream;
    const int n = L.n;
    // cnt[0 / 1]: chunks per direction-order position of the block-inverse streams; cnt[2] / cnt[3]: chunks per sweep position of
    // the FULL forward / backward row stream (k_gs_stream) - only the forward one is stored: the residual update r = L (hdu - du)
    // reads every strictly-lower coupling of a row, and one stream with one partly filled last chunk per row (597 MB at C2 level 0)
    // beats "ext rows + in-half rows" with two of them (837 MB measured)
    HOT_CUDA(s->scratch_i.reserve(4 * (size_t)n + 8));
    int* cnt[4] = {s->scratch_i.p, s->scratch_i.p + (n + 1), s->scratch_i.p + 2 * ((size_t)n + 1), s->scratch_i.p + 3 * ((size_t)n + 1)};
    HOT_CUDA(cudaMemsetAsync(s->scratch_i.p, 0, (4 * (size_t)n + 4) * sizeof(int), st));
    for (int d = 0; d < 2; ++d) {
        HOT_CUDA(L.gx_off[d].reserve((size_t)n + 1));
        k_gx_stream<false><<<nblk(32L * n), TPB, 0, st>>>(n, d, L.gs_seq.p, L.gs_rank.p, L.gs_pblock.p, L.gs_block_start.p, L.col.p, L.val.p, cnt[d],
            nullptr, nullptr);
        HOT_LAUNCHED(s);
    }
    k_gs_stream<false><<<nblk(32L * n), TPB, 0, st>>>(n, L.gs_seq.p, L.gs_rank.p, L.gs_pblock.p, L.gs_block_start.p, L.col.p, L.val.p, cnt[2], cnt[3],
        nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    HOT_LAUNCHED(s);
    HOT_CUDA(L.gs_off[0].reserve((size_t)n + 1));
    int* off[3] = {L.gx_off[0].p, L.gx_off[1].p, L.gs_off[0].p};
    for (int k = 0; k < 3; ++k) {
        int rc = with_tmp(s, [&](void* t, size_t& b) { return cub::DeviceScan::ExclusiveSum(t, b, cnt[k], off[k], n + 1, st); });
        if (rc) return rc;
        HOT_CUDA(cudaMemcpyAsync(s->hcount + 12 + k, off[k] + n, sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    HOT_CUDA(cudaStreamSynchronize(st));
    for (int d = 0; d < 2; ++d) {
        L.gx_chunks[d] = s->hcount[12 + d];
        HOT_CUDA(L.gx_data[d].reserve((size_t)L.gx_chunks[d] * GX_REC + GX_REC));
    }
    L.gs_chunks[0] = s->hcount[14];
    L.gs_chunks[1] = 0;
    HOT_CUDA(L.gs_code[0].reserve((size_t)L.gs_chunks[0] * 32 + 32));
    HOT_CUDA(L.gs_sval[0].reserve((size_t)L.gs_chunks[0] * 9 * 32 + 32));
    HOT_FUNC_ATTR_ONCE(s, k_gx_inverse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GXInvShared));
    for (int d = 0; d < 2; ++d) {
        k_gx_stream<true><<<nblk(32L * n), TPB, 0, st>>>(n, d, L.gs_seq.p, L.gs_rank.p, L.gs_pblock.p, L.gs_block_start.p, L.col.p, L.val.p, nullptr,
            L.gx_off[d].p, L.gx_data[d].p);
        HOT_LAUNCHED(s);
        k_gx_inverse<<<2 * L.n_blocks, GXI_THREADS, sizeof(GXInvShared), st>>>(d, n, L.gs_block_start.p, L.gs_seq.p, L.gs_colrank.p, L.col.p, L.val.p,
            L.dinv.p, L.gx_off[d].p, L.gx_data[d].p);
        HOT_LAUNCHED(s);
    }
    k_gs_stream<true><<<nblk(32L * n), TPB, 0, st>>>(n, L.gs_seq.p, L.gs_rank.p, L.gs_pblock.p, L.gs_block_start.p, L.col.p, L.val.p, nullptr, nullptr,
        L.gs_off[0].p, nullptr, L.gs_code[0].p, L.gs_sval[0].p, nullptr, nullptr);
    HOT_LAUNCHED(s);
    return 0;
}

int build_gs_schedule(Sim* s, MGLevel& L)
{
    cudaStream_t st = s->stream;
    const int n = L.n;
    HOT_CUDA(s->cand_key.reserve(n));
    HOT_CUDA(s->cand_key_alt.reserve(n));
    HOT_CUDA(s->cand_val.reserve(n));
    HOT_CUDA(s->cand_val_alt.reserve(n));
    HOT_CUDA(s->head_flag.reserve(n));
    HOT_CUDA(s->scratch_i.reserve(3 * (size_t)n + 8));
    HOT_CUDA(s->keys_alt.reserve(2 * (size_t)n));
    HOT_CUDA(s->dcount.reserve(16));
    HOT_CUDA(L.gs_seq.reserve(n));
    HOT_CUDA(L.gs_rank.reserve(n));
    HOT_CUDA(L.gs_block_start.reserve((size_t)n + 1));
    if (!s->hcount) HOT_CUDA(cudaMallocHost((void**)&s->hcount, 32 * sizeof(int)));
    k_block_keys<<<nblk(n), TPB, 0, st>>>(n, L.coord.p, s->cand_key.p, s->cand_val.p);
    HOT_LAUNCHED(s);
    int rc = with_tmp(s, [&](void* t, size_t& b) {
        return cub::DeviceRadixSort::SortPairs(t, b, s->cand_key.p, s->cand_key_alt.p, s->cand_val.p, s->cand_val_alt.p, n, 0, 30, st);
    });
    if (rc) return rc;
    // segments of equal block key: inclusive scan of the head flags, head positions
    k_head_flags32<<<nblk(n), TPB, 0, st>>>(n, s->cand_key_alt.p, s->head_flag.p);
    HOT_LAUNCHED(s);
    int* seg_incl = s->scratch_i.p;
    int* head_pos = s->scratch_i.p + n;
    rc = with_tmp(s, [&](void* t, size_t& b) { return cub::DeviceScan::InclusiveSum(t, b, s->head_flag.p, seg_incl, n, st); });
    if (rc) return rc;
    rc = with_tmp(s, [&](void* t, size_t& b) {
        return cub::DeviceSelect::Flagged(t, b, cub::CountingInputIterator<int>(0), s->head_flag.p, head_pos, s->dcount.p, n, st);
    });
    if (rc) return rc;
    uint64_t* key2 = s->keys_alt.p;
    uint64_t* key2_sorted = s->keys_alt.p + n;
    k_sweep_keys<<<nblk(n), TPB, 0, st>>>(n, s->cand_key_alt.p, s->cand_val_alt.p, seg_incl, head_pos, key2);
    HOT_LAUNCHED(s);
    rc = with_tmp(s, [&](void* t, size_t& b) {
        return cub::DeviceRadixSort::SortPairs(t, b, key2, key2_sorted, s->cand_val_alt.p, L.gs_seq.p, n, 0, 35, st);
    });
    if (rc) return rc;
    HOT_CUDA(cudaMemsetAsync(s->dcount.p, 0, 16 * sizeof(int), st));
    k_head_flags_key2<<<nblk(n), TPB, 0, st>>>(n, key2_sorted, s->head_flag.p, s->dcount.p + 1);
    HOT_LAUNCHED(s);
    rc = with_tmp(s, [&](void* t, size_t& b) {
        return cub::DeviceSelect::Flagged(t, b, cub::CountingInputIterator<int>(0), s->head_flag.p, L.gs_block_start.p, s->dcount.p, n, st);
    });
    if (rc) return rc;
    k_rank<<<nblk(n), TPB, 0, st>>>(n, L.gs_seq.p, L.gs_rank.p);
    HOT_LAUNCHED(s);
    HOT_CUDA(L.gs_colrank.reserve((size_t)n * W));
    k_colrank<<<nblk((long)n * W), TPB, 0, st>>>((long)n * W, L.col.p, L.gs_rank.p, L.gs_colrank.p);
    HOT_LAUNCHED(s);
    HOT_CUDA(cudaMemcpyAsync(s->hcount, s->dcount.p, 9 * sizeof(int), cudaMemcpyDeviceToHost, st));
    HOT_CUDA(cudaStreamSynchronize(st));
    L.n_blocks = s->hcount[0];
    HOT_CUDA(cudaMemcpyAsync(L.gs_block_start.p + L.n_blocks, &n, sizeof(int), cudaMemcpyHostToDevice, st));
    HOT_CUDA(cudaStreamSynchronize(st)); // &n is a stack variable
    L.color_first_block[0] = 0;
    for (int c = 0; c < 8; ++c) L.color_first_block[c + 1] = L.color_first_block[c] + s->hcount[1 + c];
    HOT_CUDA(L.gs_pblock.reserve(n));
    k_gs_pblock<<<nblk(L.n_blocks), TPB, 0, st>>>(L.n_blocks, L.gs_block_start.p, L.gs_pblock.p);
    HOT_LAUNCHED(s);
    if (gs_inv_mode()) return build_gx_streams(s, L);
    // the per-direction row stream of the sweeps
    HOT_CUDA(L.gs_off[0].reserve((size_t)n + 1));
    HOT_CUDA(L.gs_off[1].reserve((size_t)n + 1));
    int* cntF = s->scratch_i.p;
    int* cntB = s->scratch_i.p + n + 1;
    HOT_CUDA(cudaMemsetAsync(s->scratch_i.p, 0, (2 * (size_t)n + 2) * sizeof(int), st));
    k_gs_stream<false><<<nblk(32L * n), TPB, 0, st>>>(n, L.gs_seq.p, L.gs_rank.p, L.gs_pblock.p, L.gs_block_start.p, L.col.p, L.val.p, cntF, cntB,
        nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    HOT_LAUNCHED(s);
    for (int d = 0; d < 2; ++d) {
        int* cnt = d ? cntB : cntF;
        rc = with_tmp(s, [&](void* t, size_t& b) { return cub::DeviceScan::ExclusiveSum(t, b, cnt, L.gs_off[d].p, n + 1, st); });
        if (rc) return rc;
        HOT_CUDA(cudaMemcpyAsync(s->hcount + 10 + d, L.gs_off[d].p + n, sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    HOT_CUDA(cudaStreamSynchronize(st));
    for (int d = 0; d < 2; ++d) {
        L.gs_chunks[d] = s->hcount[10 + d];
        HOT_CUDA(L.gs_code[d].reserve((size_t)L.gs_chunks[d] * 32 + 32));
        HOT_CUDA(L.gs_sval[d].reserve((size_t)L.gs_chunks[d] * 9 * 32 + 32));
    }
    k_gs_stream<true><<<nblk(32L * n), TPB, 0, st>>>(n, L.gs_seq.p, L.gs_rank.p, L.gs_pblock.p, L.gs_block_start.p, L.col.p, L.val.p, nullptr, nullptr,
        L.gs_off[0].p, L.gs_off[1].p, L.gs_code[0].p, L.gs_sval[0].p, L.gs_code[1].p, L.gs_sval[1].p);
    HOT_LAUNCHED(s);
    return 0;
}

// candidates of the coarse node set of level F: ascending unique keys in s->mg_heads_key, first candidate position of each in
// s->mg_heads_pos; the count in *nc_out
int coarse_candidates(Sim* s, MGLevel& F, int* nc_out)
{
    cudaStream_t st = s->stream;
    const int nf = F.n;
    const long nc8 = (long)nf * 8;
    // persistent scratch (a cudaMalloc / cudaFree pair per buffer and level used to dominate the hierarchy build)
    DevBuf<uint64_t>&ckey = s->mg_ckey, &ckey_sorted = s->mg_ckey_sorted, &heads_key = s->mg_heads_key;
    DevBuf<int>&cpos = s->mg_cpos, &cpos_sorted = s->mg_cpos_sorted, &heads_pos = s->mg_heads_pos, &order = s->mg_order;
    HOT_CUDA(ckey.reserve(nc8));
    HOT_CUDA(ckey_sorted.reserve(nc8));
    HOT_CUDA(cpos.reserve(nc8));
    HOT_CUDA(cpos_sorted.reserve(nc8));
    HOT_CUDA(heads_pos.reserve(nc8));
    HOT_CUDA(heads_key.reserve(nc8));
    HOT_CUDA(order.reserve(2 * nc8));
    HOT_CUDA(s->head_flag.reserve(nc8));
    HOT_CUDA(s->dcount.reserve(16));
    if (!s->hcount) HOT_CUDA(cudaMallocHost((void**)&s->hcount, 32 * sizeof(int)));
    k_coarse_candidates<<<nblk(nc8), TPB, 0, st>>>(nf, F.coord.p, ckey.p, cpos.p);
    HOT_LAUNCHED(s);
    int rc = with_tmp(s, [&](void* t, size_t& b) {
        return cub::DeviceRadixSort::SortPairs(t, b, ckey.p, ckey_sorted.p, cpos.p, cpos_sorted.p, (int)nc8, 0, 64, st);
    });
    if (rc) return rc;
    k_head_flags64<<<nblk(nc8), TPB, 0, st>>>(nc8, ckey_sorted.p, s->head_flag.p);
    HOT_LAUNCHED(s);
    rc = with_tmp(s, [&](void* t, size_t& b) {
        return cub::DeviceSelect::Flagged(t, b, ckey_sorted.p, s->head_flag.p, heads_key.p, s->dcount.p, (int)nc8, st);
    });
    if (rc) return rc;
    rc = with_tmp(s, [&](void* t, size_t& b) {
        return cub::DeviceSelect::Flagged(t, b, cpos_sorted.p, s->head_flag.p, heads_pos.p, s->dcount.p, (int)nc8, st);
    });
    if (rc) return rc;
    HOT_CUDA(cudaMemcpyAsync(s->hcount, s->dcount.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    HOT_CUDA(cudaStreamSynchronize(st));
    *nc_out = s->hcount[0];
    return 0;
}

int coarsen(Sim* s, MGLevel& F, MGLevel& C)
{
    cudaStream_t st = s->stream;
    const int nf = F.n;
    const long nc8 = (long)nf * 8;
    DevBuf<uint64_t>& heads_key = s->mg_heads_key;
    DevBuf<int>&cpos = s->mg_cpos, &heads_pos = s->mg_heads_pos, &order = s->mg_order;
    int nc_cand = 0;
    int rc = coarse_candidates(s, F, &nc_cand);
    if (rc) return rc;
    s->hcount[0] = nc_cand;
    const int nc = s->hcount[0];
    if (nc <= 0) return fail(s, "multigrid: empty coarse level");
    C.n = nc;
    // first-touch order = heads sorted by their first candidate position
    int* iota = order.p;
    int* ord = order.p + nc8;
    k_iota<<<nblk(nc), TPB, 0, st>>>(nc, iota);
    HOT_LAUNCHED(s);
    rc = with_tmp(s, [&](void* t, size_t& b) {
        return cub::DeviceRadixSort::SortPairs(t, b, (const uint32_t*)heads_pos.p, (uint32_t*)cpos.p, iota, ord, nc, 0, 32, st);
    });
    if (rc) return rc;
    HOT_CUDA(C.coord.reserve(3 * (size_t)nc));
    HOT_CUDA(C.key_sorted.reserve(nc));
    HOT_CUDA(C.id_sorted.reserve(nc));
    HOT_CUDA(cudaMemcpyAsync(C.key_sorted.p, heads_key.p, (size_t)nc * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
    k_coarse_finish<<<nblk(nc), TPB, 0, st>>>(nc, ord, heads_key.p, C.coord.p, C.id_sorted.p);
    HOT_LAUNCHED(s);
    // transfer operators
    HOT_CUDA(F.pcol.reserve((size_t)nf * 8));
    HOT_CUDA(F.pw.reserve((size_t)nf * 8));
    HOT_CUDA(F.rcol.reserve((size_t)nc * 32));
    HOT_CUDA(F.rw.reserve((size_t)nc * 32));
    k_build_P<<<nblk(nf), TPB, 0, st>>>(nf, F.coord.p, C.key_sorted.p, C.id_sorted.p, nc, F.pcol.p, F.pw.p);
    HOT_LAUNCHED(s);
    k_build_R<<<nblk((long)nc * 32), TPB, 0, st>>>(nc, C.coord.p, F.key_sorted.p, F.id_sorted.p, nf, F.rcol.p, F.rw.p);
    HOT_LAUNCHED(s);
    // Galerkin product
    HOT_CUDA(C.col.reserve((size_t)nc * W));
    HOT_CUDA(C.val.reserve((size_t)nc * 9 * W));
    k_galerkin<<<nc, W, 0, st>>>(nc, C.coord.p, C.key_sorted.p, C.id_sorted.p, F.rcol.p, F.rw.p, F.val.p, C.col.p, C.val.p);
    HOT_LAUNCHED(s);
    return 0;
}

// Level 0 -> 1 of a PARTITIONED object: level 0 stays distributed (local nodes incl. the ghost ring), level 1 and everything coarser
// is REPLICATED on every rank (1/7 of the fine level; its smoothers then need no exchange at all):
//  * coarse node set = union over the ranks of the local candidate sets (all-gather of the coordinate keys), node c = c-th smallest key
//    on every rank alike;
//  * R rows over this rank's fine nodes, masked to the nodes it counts (own_node = page authority): restriction and the Galerkin
//    product R A P become partial sums, completed by an all-reduce (the rows of authority nodes are complete and have all their columns);
//  * P rows of the local fine nodes point at the replicated coarse ids: prolongation needs no exchange.
int coarsen_dist(Sim* s, MGLevel& F, MGLevel& C)
{
    cudaStream_t st = s->stream;
    const int nf = F.n, Wd = s->world;
    int nc_local = 0;
    int rc = coarse_candidates(s, F, &nc_local);
    if (rc) return rc;
    std::vector<int> counts(Wd);
    rc = dist_all_gather_host(s, &nc_local, counts.data(), sizeof(int));
    if (rc) return rc;
    long maxc = 1;
    for (int c : counts) maxc = std::max<long>(maxc, c);
    // all-gather of the padded key lists, sort, unique
    DevBuf<uint64_t>&ckey = s->mg_ckey, &ckey_sorted = s->mg_ckey_sorted;
    const long tot = maxc * Wd;
    HOT_CUDA(ckey.reserve((size_t)tot + maxc));
    HOT_CUDA(ckey_sorted.reserve((size_t)tot));
    HOT_CUDA(s->head_flag.reserve((size_t)tot));
    uint64_t* mine = ckey.p + tot;
    k_fill_u64<<<nblk(maxc), TPB, 0, st>>>(maxc, NOKEY, mine);
    HOT_LAUNCHED(s);
    HOT_CUDA(cudaMemcpyAsync(mine, s->mg_heads_key.p, (size_t)nc_local * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
    rc = dist_all_gather_dev(s, mine, ckey.p, maxc * (long)sizeof(uint64_t));
    if (rc) return rc;
    rc = with_tmp(s, [&](void* t, size_t& b) { return cub::DeviceRadixSort::SortKeys(t, b, ckey.p, ckey_sorted.p, (int)tot, 0, 64, st); });
    if (rc) return rc;
    k_head_flags64<<<nblk(tot), TPB, 0, st>>>(tot, ckey_sorted.p, s->head_flag.p);
    HOT_LAUNCHED(s);
    HOT_CUDA(C.key_sorted.reserve((size_t)tot));
    rc = with_tmp(s, [&](void* t, size_t& b) {
        return cub::DeviceSelect::Flagged(t, b, ckey_sorted.p, s->head_flag.p, C.key_sorted.p, s->dcount.p, (int)tot, st);
    });
    if (rc) return rc;
    HOT_CUDA(cudaMemcpyAsync(s->hcount, s->dcount.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    HOT_CUDA(cudaStreamSynchronize(st));
    const int nc = s->hcount[0];
    if (nc <= 0) return fail(s, "multigrid: empty coarse level");
    C.n = nc;
    HOT_CUDA(C.coord.reserve(3 * (size_t)nc));
    HOT_CUDA(C.id_sorted.reserve(nc));
    k_coarse_from_keys<<<nblk(nc), TPB, 0, st>>>(nc, C.key_sorted.p, C.coord.p, C.id_sorted.p);
    HOT_LAUNCHED(s);
    HOT_CUDA(F.pcol.reserve((size_t)nf * 8));
    HOT_CUDA(F.pw.reserve((size_t)nf * 8));
    HOT_CUDA(F.rcol.reserve((size_t)nc * 32));
    HOT_CUDA(F.rw.reserve((size_t)nc * 32));
    k_build_P<<<nblk(nf), TPB, 0, st>>>(nf, F.coord.p, C.key_sorted.p, C.id_sorted.p, nc, F.pcol.p, F.pw.p);
    HOT_LAUNCHED(s);
    k_build_R<<<nblk((long)nc * 32), TPB, 0, st>>>(nc, C.coord.p, F.key_sorted.p, F.id_sorted.p, nf, F.rcol.p, F.rw.p);
    HOT_LAUNCHED(s);
    k_mask_R<<<nblk((long)nc * 32), TPB, 0, st>>>((long)nc * 32, F.rcol.p, s->own_node.p, F.rw.p);
    HOT_LAUNCHED(s);
    HOT_CUDA(C.col.reserve((size_t)nc * W));
    HOT_CUDA(C.val.reserve((size_t)nc * 9 * W));
    k_galerkin<<<nc, W, 0, st>>>(nc, C.coord.p, C.key_sorted.p, C.id_sorted.p, F.rcol.p, F.rw.p, F.val.p, C.col.p, C.val.p);
    HOT_LAUNCHED(s);
    return dist_allreduce_buffer(s, C.val.p, (long)nc * 9 * W, 0);
}

int finish_level(Sim* s, MGLevel& L, int Ainv, bool colors)
{
    HOT_CUDA(L.diag.reserve(9 * (size_t)L.n));
    HOT_CUDA(L.dinv.reserve(9 * (size_t)L.n));
    k_level_diagonal<<<nblk(L.n), TPB, 0, s->stream>>>(L.n, Ainv, L.val.p, L.diag.p, L.dinv.p);
    HOT_LAUNCHED(s);
    int rc = reserve_level_vectors(s, L);
    if (rc) return rc;
    if (colors) return build_gs_schedule(s, L);
    return 0;
}

inline int regular_iters(const Sim* s, int level) { return s->mg_times + level * s->mg_levelscale; }
inline int top_iters(const Sim* s, int level)
{ // setup_parameters, MultigridPreconditioner.h:524-551 (topDownMGS = false)
    if (s->mg_levels == 1) return regular_iters(s, level);
    if (!(s->mg_coarse == 2 || s->mg_coarse == 6)) return regular_iters(s, level) * 3;
    return 10000;
}

int level_project(Sim* s, int level, double* v)
{ // A.project: only level 0 of a system that was not BC-projected carries the projection (:695-699)
    if (level == 0 && !s->matrix_bcproject) return bc_project(s, v);
    return 0;
}

int mg_scale(Sim* s, MGLevel& L, const double* r, double* mr, double scale = 1.0)
{ // scaler_func: D^-1 r with the -Ainv flavour baked into dinv
    k_block_diag<<<nblk(L.n), TPB, 0, s->stream>>>(L.n, L.dinv.p, r, mr, scale);
    HOT_LAUNCHED(s);
    return 0;
}

#define RC(x)                  \
    do {                       \
        int rc__ = (x);        \
        if (rc__) return rc__; \
    } while (0)

int smooth_jacobi(Sim* s, int level, double* u, double* r, int iterations)
{
    MGLevel& L = *s->levels[level];
    const long m = 3L * L.n;
    for (; iterations--;) {
        RC(mg_scale(s, L, r, L.du.p, s->mg_topomega));
        RC(vec_axpy(s, m, 1.0, L.du.p, u));
        RC(level_spmv(s, level, L.du.p, L.dAu.p));
        RC(level_project(s, level, L.dAu.p));
        RC(vec_axpy(s, m, -1.0, L.dAu.p, r));
    }
    return 0;
}
int smooth_optimal_jacobi(Sim* s, int level, double* u, double* r, int iterations, double tolerance)
{
    MGLevel& L = *s->levels[level];
    const long m = 3L * L.n;
    double* sc = s->red_out.p; // device scalars
    for (; iterations--;) {
        double rr;
        RC(vec_dot(s, m, r, r, sc, &rr));
        if (sqrt(rr) < tolerance) break;
        RC(mg_scale(s, L, r, L.du.p));
        RC(level_spmv(s, level, L.du.p, L.dAu.p));
        RC(level_project(s, level, L.dAu.p));
        RC(vec_dot(s, m, L.du.p, r, sc + 1, nullptr));
        RC(vec_dot(s, m, L.du.p, L.dAu.p, sc + 2, nullptr));
        RC(vec_axpy_dev(s, m, sc + 1, sc + 2, 1.0, L.du.p, u));
        RC(vec_axpy_dev(s, m, sc + 1, sc + 2, -1.0, L.dAu.p, r));
    }
    return 0;
}
// cg_smooth, MultigridPreconditioner.h:190-226: Jacobi-PCG until z.r < 0.25 z0.r0 of the restricted INITIAL residual
int smooth_cg(Sim* s, int level, double* u, double* r, int iterations)
{
    MGLevel& L = *s->levels[level];
    const long m = 3L * L.n;
    double* sc = s->red_out.p;
    double* z = L.tmp.p;
    double zTrk0, zTrk;
    RC(mg_scale(s, L, L.initial_residual.p, z));
    RC(vec_dot(s, m, z, L.initial_residual.p, sc, &zTrk0));
    RC(mg_scale(s, L, r, z));
    RC(vec_copy(s, m, z, L.du.p));
    RC(vec_dot(s, m, z, r, sc + 1, &zTrk)); // sc[1] = z.r (current)
    const double tolerance = zTrk0 * 0.25;
    int cnt = 0;
    for (; iterations--;) {
        if (zTrk < tolerance) break;
        RC(level_spmv(s, level, L.du.p, L.dAu.p));
        RC(level_project(s, level, L.dAu.p));
        RC(vec_dot(s, m, L.dAu.p, L.du.p, sc + 2, nullptr));
        RC(vec_axpy_dev(s, m, sc + 1, sc + 2, 1.0, L.du.p, u));
        RC(vec_axpy_dev(s, m, sc + 1, sc + 2, -1.0, L.dAu.p, r));
        RC(mg_scale(s, L, r, z));
        HOT_CUDA(cudaMemcpyAsync(sc + 3, sc + 1, sizeof(double), cudaMemcpyDeviceToDevice, s->stream)); // zTrkPre
        RC(vec_dot(s, m, z, r, sc + 1, &zTrk));
        RC(vec_xpay_dev(s, m, z, sc + 1, sc + 3, L.du.p)); // du = z + beta du
        ++cnt;
    }
    s->last_cg_iters = cnt;
    return 0;
}
// chebyshev_smooth, MultigridPreconditioner.h:227-264
int smooth_chebyshev(Sim* s, int level, double* u, double* r, int iterations)
{
    MGLevel& L = *s->levels[level];
    const long m = 3L * L.n;
    double* p = L.tmp.p;
    const double d = (L.lMax + L.lMin) / 2, c = (L.lMax - L.lMin) / 2;
    int cnt = 1;
    iterations--;
    RC(mg_scale(s, L, r, p));
    double alpha = 1 / d, beta;
    RC(vec_copy(s, m, p, L.du.p));
    RC(level_spmv(s, level, L.du.p, L.dAu.p));
    RC(level_project(s, level, L.dAu.p));
    RC(vec_axpy(s, m, alpha, L.du.p, u));
    RC(vec_axpy(s, m, -alpha, L.dAu.p, r));
    for (; iterations-- > 0; ++cnt) {
        RC(mg_scale(s, L, r, p));
        beta = 0.5 * c * c * alpha * alpha;
        if (cnt > 1) beta *= 0.5;
        alpha = 1 / (d - beta / alpha);
        k_xpay<<<nblk(m), TPB, 0, s->stream>>>(m, p, beta, L.du.p);
        HOT_LAUNCHED(s);
        RC(level_spmv(s, level, L.du.p, L.dAu.p));
        RC(level_project(s, level, L.dAu.p));
        RC(vec_axpy(s, m, alpha, L.du.p, u));
        RC(vec_axpy(s, m, -alpha, L.dAu.p, r));
    }
    return 0;
}
// SquareMatrix::estimate2norm, SquareMatrix.h:375-475 (power iteration on A A; lMin = lMax / 30 "experience")
int estimate2norm(Sim* s, int level, double tol = 1e-6)
{
    MGLevel& L = *s->levels[level];
    const long m = 3L * L.n;
    double *v = L.du.p, *x = L.dAu.p, *sc = s->red_out.p;
    const bool plain = s->dot_plain;
    s->dot_plain = level > 0;
    k_sign_pattern<<<nblk(m), TPB, 0, s->stream>>>(m, v);
    HOT_LAUNCHED(s);
    RC(level_spmv(s, level, v, x));
    k_abs<<<nblk(m), TPB, 0, s->stream>>>(m, x);
    HOT_LAUNCHED(s);
    double xx, vv;
    RC(vec_dot(s, m, x, x, sc, &xx));
    double e = sqrt(xx), e0 = 0;
    if (e == 0) {
        L.lMin = L.lMax = 0;
        s->dot_plain = plain;
        return 0;
    }
    RC(vec_scale(s, m, 1.0 / e, x));
    for (int iter = 0; iter < 512 && fabs(e - e0) > tol * e; ++iter) {
        e0 = e;
        RC(level_spmv(s, level, x, v));
        RC(level_spmv(s, level, v, x));
        RC(vec_dot(s, m, x, x, sc, &xx));
        RC(vec_dot(s, m, v, v, sc + 1, &vv));
        const double normx = sqrt(xx);
        e = normx / sqrt(vv);
        RC(vec_scale(s, m, 1.0 / normx, x));
    }
    L.lMax = e;
    L.lMin = e / 30;
    s->dot_plain = plain;
    return 0;
}
template <int THREADS, bool STREAM>
int launch_gs_sweep(Sim* s, GSArgs& a, int max_blocks_per_color, bool* launched)
{
    static int per_sm_dev[64], n_sm_dev[64];
    static bool init_dev[64] = {false};
    const int dslot = s->device & 63; // occupancy and the cooperative-launch capability are per device
    if (!init_dev[dslot]) { per_sm_dev[dslot] = -1; n_sm_dev[dslot] = 0; init_dev[dslot] = true; }
    int &per_sm = per_sm_dev[dslot], &n_sm = n_sm_dev[dslot];
    if (per_sm < 0) {
        int dev = s->device, coop = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gs_sweep<THREADS, STREAM>, THREADS, 0) != cudaSuccess || !coop) per_sm = 0;
    }
    *launched = false;
    if (per_sm <= 0) return 0;
    int grid = std::min(max_blocks_per_color, per_sm * n_sm);
    if (a.fuse_update) grid = std::max(grid, std::min(per_sm * n_sm, (a.n + THREADS / 32 - 1) / (THREADS / 32)));
    if (grid < 1) grid = 1;
    void* params[] = {&a};
    cudaError_t e = cudaLaunchCooperativeKernel((void*)k_gs_sweep<THREADS, STREAM>, dim3(grid), dim3(THREADS), params, 0, s->stream);
    if (e != cudaSuccess) {
        cudaGetLastError(); // clear; fall back to per-phase launches
        per_sm = 0;
        return 0;
    }
    s->launches++;
    *launched = true;
    return 0;
}

template <int THREADS, int NST>
int launch_gs_sweep_ring(Sim* s, GSArgs& a, int max_blocks_per_color, bool* launched)
{
    static int per_sm_dev[64], n_sm_dev[64];
    static bool init_dev[64] = {false};
    const int dslot = s->device & 63;
    if (!init_dev[dslot]) { per_sm_dev[dslot] = -1; n_sm_dev[dslot] = 0; init_dev[dslot] = true; }
    int &per_sm = per_sm_dev[dslot], &n_sm = n_sm_dev[dslot];
    constexpr size_t smem = sizeof(GSRingShared<NST>);
    if (per_sm < 0) {
        int dev = s->device, coop = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        if (cudaFuncSetAttribute(k_gs_sweep_ring<THREADS, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess
            || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gs_sweep_ring<THREADS, NST>, THREADS, smem) != cudaSuccess || !coop) {
            cudaGetLastError();
            per_sm = 0;
        }
    }
    *launched = false;
    if (per_sm <= 0) return 0;
    int grid = std::min(max_blocks_per_color, per_sm * n_sm);
    if (a.fuse_update) grid = std::max(grid, std::min(per_sm * n_sm, (a.n + THREADS / 32 - 1) / (THREADS / 32)));
    if (grid < 1) grid = 1;
    void* params[] = {&a};
    cudaError_t e = cudaLaunchCooperativeKernel((void*)k_gs_sweep_ring<THREADS, NST>, dim3(grid), dim3(THREADS), params, smem, s->stream);
    if (e != cudaSuccess) {
        cudaGetLastError(); // clear; fall back to per-phase launches
        per_sm = 0;
        return 0;
    }
    s->launches++;
    *launched = true;
    return 0;
}

template <int THREADS, int D, int MINB>
int launch_gx_sweep(Sim* s, GSArgs& a, int max_blocks_per_color, bool* launched)
{
    static int per_sm_dev[64], n_sm_dev[64];
    static bool init_dev[64] = {false};
    const int dslot = s->device & 63;
    if (!init_dev[dslot]) { per_sm_dev[dslot] = -1; n_sm_dev[dslot] = 0; init_dev[dslot] = true; }
    int &per_sm = per_sm_dev[dslot], &n_sm = n_sm_dev[dslot];
    constexpr size_t smem = sizeof(GXShared<THREADS / 32 - 1, D>);
    if (per_sm < 0) {
        int dev = s->device, coop = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        if (cudaFuncSetAttribute(k_gx_sweep<THREADS, D, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess
            || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gx_sweep<THREADS, D, MINB>, THREADS, smem) != cudaSuccess || !coop) {
            cudaGetLastError();
            per_sm = 0;
        }
    }
    *launched = false;
    if (per_sm <= 0 || max_blocks_per_color > per_sm * n_sm) return 0; // (every block of a colour needs its own resident CTA slot or a loop; keep one wave)
    int grid = std::min(max_blocks_per_color, per_sm * n_sm);
    if (a.fuse_update) grid = std::max(grid, std::min(per_sm * n_sm, (a.n + THREADS / 32 - 1) / (THREADS / 32)));
    if (grid < 1) grid = 1;
    void* params[] = {&a};
    cudaError_t e = cudaLaunchCooperativeKernel((void*)k_gx_sweep<THREADS, D, MINB>, dim3(grid), dim3(THREADS), params, smem, s->stream);
    if (e != cudaSuccess) {
        cudaGetLastError(); // clear; fall back to per-phase launches
        per_sm = 0;
        return 0;
    }
    s->launches++;
    *launched = true;
    return 0;
}

// one colour phase of the default shape; `dependent`: launched with programmatic stream serialisation, i.e. its CTAs may start (and run
// up to their griddepcontrol.wait) before the previous kernel of the stream has finished
template <bool FWD>
int launch_gx_block(Sim* s, int b0, int n_blocks, const GSArgs& a, bool dependent)
{
    constexpr int THREADS = 288, D = 3;
    constexpr size_t smem = sizeof(GXShared<THREADS / 32 - 1, D>);
    HOT_FUNC_ATTR_ONCE(s, (k_gx_block<FWD, THREADS, D, 3>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)n_blocks);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = dependent ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    HOT_CUDA(cudaLaunchKernelEx(&cfg, k_gx_block<FWD, THREADS, D, 3>, b0, a));
    return 0;
}

// one colour phase in cluster form: n_blocks clusters of CL CTAs
template <bool FWD, int CL>
int launch_gx_block_cl(Sim* s, int b0, int n_blocks, const GSArgs& a)
{
    constexpr int THREADS = 256, D = 4;
    constexpr size_t smem = sizeof(GXShared<THREADS / 32 - 1, D>);
    HOT_FUNC_ATTR_ONCE(s, (k_gx_block_cl<FWD, THREADS, D, 3, CL>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_blocks * CL));
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    HOT_CUDA(cudaLaunchKernelEx(&cfg, k_gx_block_cl<FWD, THREADS, D, 3, CL>, b0, a));
    return 0;
}

// gs_smooth, MultigridPreconditioner.h:266-318
int smooth_gs(Sim* s, int level, double* u, double* r, int iterations)
{
    MGLevel& L = *s->levels[level];
    if (L.n_blocks <= 0) return fail(s, "gs_smooth: the hierarchy was built without the colour schedule (smoother / coarseSolver 5)");
    cudaStream_t st = s->stream;
    const bool project = level == 0 && !s->matrix_bcproject;
    GSArgs a;
    a.n = L.n;
    int max_blocks = 0;
    for (int c = 0; c < 9; ++c) a.cfb[c] = L.color_first_block[c];
    for (int c = 0; c < 8; ++c) max_blocks = std::max(max_blocks, a.cfb[c + 1] - a.cfb[c]);
    a.block_start = L.gs_block_start.p; a.seq = L.gs_seq.p; a.colrank = L.gs_colrank.p; a.col = L.col.p;
    a.val = L.val.p; a.dinv = L.dinv.p; a.diag = L.diag.p;
    a.r = r; a.hdu = L.tmp.p; a.dhdu = L.dAu.p; a.du = L.du.p; a.u = u; // hdu: unscaled forward solution; dhdu = D hdu
    a.fuse_update = project ? 0 : 1;
    const bool inv = gs_inv_mode(); // block-inverse form: its own streams (the substitution forms' streams are not built)
    static const bool stream_env = !(getenv("HOT_GS_STREAM") && atoi(getenv("HOT_GS_STREAM")) == 0); // A/B switch, default on
    const bool use_stream = stream_env && !inv;
    for (int d = 0; d < 2; ++d) {
        a.xoff[d] = inv ? L.gx_off[d].p : nullptr;
        a.xdata[d] = inv ? L.gx_data[d].p : nullptr;
    }
    for (int d = 0; d < 2; ++d) { // (block-inverse form: the full forward stream only, for the residual update)
        const bool have = use_stream || (inv && d == 0);
        a.soff[d] = have ? L.gs_off[d].p : nullptr;
        a.scode[d] = have ? L.gs_code[d].p : nullptr;
        a.sval[d] = have ? L.gs_sval[d].p : nullptr;
    }
    a.pblock = L.gs_pblock.p;
    // A/B switch, default on: the stream reaches shared memory through a ring of TMA bulk copies (gs_block_ring)
    static const bool ring_env = !(getenv("HOT_GS_RING") && atoi(getenv("HOT_GS_RING")) == 0);
    const bool use_ring = use_stream && ring_env;
    static const bool no_ident = getenv("HOT_GS_STREAM_UPDATE") && atoi(getenv("HOT_GS_STREAM_UPDATE")) == 0; // A/B switch
    a.stream_update = ((use_stream || inv) && !project && s->mg_Ainv == 1 && !no_ident) ? 1 : 0;
    iterations = (iterations + 1) >> 1;
    static long long* dbg_dev = nullptr;
    const char* dbg_env = getenv("HOT_GS_DEBUG");
    if (dbg_env && !dbg_dev) {
        cudaMalloc((void**)&dbg_dev, 32 * sizeof(long long));
        cudaMemcpyToSymbol(g_gs_dbg, &dbg_dev, sizeof(dbg_dev));
    }
    if (dbg_dev) {
        long long init[32] = {0};
        init[31] = a.cfb[0] + atoi(dbg_env); // watched block: the n-th of colour 0
        cudaMemcpy(dbg_dev, init, sizeof init, cudaMemcpyHostToDevice);
    }
    for (; iterations--;) {
        bool launched = false;
        // few blocks per colour: a big CTA per block (16 warps stream the rows) on one SM each; many: 3 CTAs of 8 warps per SM
        // (many blocks per colour: the per-phase launches below keep 3 CTAs per SM busy, which measures faster than the
        //  cooperative form whose register budget allows only 2)
        static const bool force_coop = getenv("HOT_GS_COOP") && atoi(getenv("HOT_GS_COOP")) != 0;
        static const bool no_coop = getenv("HOT_GS_COOP") && atoi(getenv("HOT_GS_COOP")) == 0; // per-phase launches on every level
        const bool dist0 = level == 0 && s->world > 1; // partitioned level 0: a take-over exchange follows every colour phase
        // block-inverse form: one launch per colour phase on EVERY level, phases 2..16 as programmatic dependents (measured at C2, GS level
        // 1 / 2: 0.29 / 0.23 ms against 0.34 / 0.27 ms for the cooperative single-launch form, which stays behind HOT_GS_COOP=1)
        if (dist0 || no_coop || (inv && (gx_cluster_cfg() > 1 || !force_coop))) {}
        else if (inv) {
            // the cooperative single-launch form: 15 consumer warps x 4 slots, one CTA per SM, levels with at most 2 blocks per SM and colour
            // (other shapes were measured and dropped: 31 warps x 2 slots, 23 x 3, and the 3-CTAs-per-SM shape for level 0 -
            //  profiles/r2_gs_experiments.md)
            if (max_blocks <= 2 * 148) RC((launch_gx_sweep<512, 4, 1>(s, a, max_blocks, &launched)));
        }
        else if (max_blocks <= 2 * 148) {
            if (use_ring) RC((launch_gs_sweep_ring<512, GS_RING_NST_COOP>(s, a, max_blocks, &launched)));
            else RC((use_stream ? launch_gs_sweep<512, true>(s, a, max_blocks, &launched) : launch_gs_sweep<512, false>(s, a, max_blocks, &launched)));
        }
        else if (force_coop)
            RC((use_stream ? launch_gs_sweep<GS_THREADS, true>(s, a, max_blocks, &launched) : launch_gs_sweep<GS_THREADS, false>(s, a, max_blocks, &launched)));
        if (!launched) {
            constexpr size_t ring_smem = sizeof(GSRingShared<GS_RING_NST>);
            constexpr size_t gx_smem7 = sizeof(GXShared<7, 3>);
            // A/B: 0 = 3 CTAs of 288 threads (8 consumer warps) per SM, 2 = of 256 threads (7 consumer warps: 0.68 against 0.66 ms at C2 level 0;
            // one 512-thread CTA per SM measured 0.78 ms and was dropped)
            static const int block_cfg = getenv("HOT_GX_BLOCK") ? atoi(getenv("HOT_GX_BLOCK")) : 0;
            if (inv && block_cfg == 2) {
                HOT_FUNC_ATTR_ONCE(s, (k_gx_block<true, 256, 3, 3>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gx_smem7);
                HOT_FUNC_ATTR_ONCE(s, (k_gx_block<false, 256, 3, 3>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gx_smem7);
            }
            // colour phases 2..16 of a sweep as programmatic dependents of the phase before them (A/B: HOT_GX_PDL=0); not with the
            // take-over exchanges of a partitioned level 0 in between
            static const bool pdl_env = !(getenv("HOT_GX_PDL") && atoi(getenv("HOT_GX_PDL")) == 0);
            const bool pdl = pdl_env && !dist0;
            bool prev_phase = false; // the previous launch of this stream was a colour phase of this sweep
            // A/B: sweep every block with a thread-block cluster of 2 / 4 CTAs (k_gx_block_cl); with it every level runs per-phase launches
            const int cluster = gx_cluster_cfg();
            if (use_ring) {
                HOT_FUNC_ATTR_ONCE(s, k_gs_block_ring<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_smem);
                HOT_FUNC_ATTR_ONCE(s, k_gs_block_ring<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_smem);
            }
            for (int c = 0; c < 8; ++c) {
                const int b0 = a.cfb[c], b1 = a.cfb[c + 1];
                if (b1 == b0) continue;
                if (inv && cluster == 2) RC((launch_gx_block_cl<true, 2>(s, b0, b1 - b0, a)));
                else if (inv && cluster == 4) RC((launch_gx_block_cl<true, 4>(s, b0, b1 - b0, a)));
                else if (inv && block_cfg == 2) k_gx_block<true, 256, 3, 3><<<b1 - b0, 256, gx_smem7, st>>>(b0, a);
                else if (inv) { RC((launch_gx_block<true>(s, b0, b1 - b0, a, pdl && prev_phase))); prev_phase = true; }
                else if (use_ring) k_gs_block_ring<true><<<b1 - b0, GS_THREADS, ring_smem, st>>>(b0, a);
                else (use_stream ? k_gs_block<true, true> : k_gs_block<true, false>)<<<b1 - b0, GS_THREADS, 0, st>>>(b0, a);
                HOT_LAUNCHED(s);
                if (dist0) RC(dist_takeover_shared(s, a.hdu, 3)); // later colours read this colour's values on pages other ranks own
            }
            for (int c = 7; c >= 0; --c) {
                const int b0 = a.cfb[c], b1 = a.cfb[c + 1];
                if (b1 == b0) continue;
                if (inv && cluster == 2) RC((launch_gx_block_cl<false, 2>(s, b0, b1 - b0, a)));
                else if (inv && cluster == 4) RC((launch_gx_block_cl<false, 4>(s, b0, b1 - b0, a)));
                else if (inv && block_cfg == 2) k_gx_block<false, 256, 3, 3><<<b1 - b0, 256, gx_smem7, st>>>(b0, a);
                else if (inv) { RC((launch_gx_block<false>(s, b0, b1 - b0, a, pdl && prev_phase))); prev_phase = true; }
                else if (use_ring) k_gs_block_ring<false><<<b1 - b0, GS_THREADS, ring_smem, st>>>(b0, a);
                else (use_stream ? k_gs_block<false, true> : k_gs_block<false, false>)<<<b1 - b0, GS_THREADS, 0, st>>>(b0, a);
                HOT_LAUNCHED(s);
                if (dist0) RC(dist_takeover_shared(s, a.du, 3));
            }
            if (!project) {
                if (a.stream_update) k_gs_stream_update<<<nblk(32L * L.n), TPB, 0, st>>>(a);
                else k_spmv_update<<<nblk(32L * L.n), TPB, 0, st>>>(L.n, L.col.p, L.val.p, L.du.p, u, r);
                HOT_LAUNCHED(s);
                if (dist0) RC(dist_takeover_shared(s, r, 3)); // (u += du is pointwise on consistent vectors)
            }
        }
        if (project) {
            RC(vec_axpy(s, 3L * L.n, 1.0, L.du.p, u));
            RC(level_spmv(s, level, L.du.p, L.dAu.p));
            RC(level_project(s, level, L.dAu.p));
            RC(vec_axpy(s, 3L * L.n, -1.0, L.dAu.p, r));
        }
    }
    if (dbg_dev) {
        long long h[32];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, dbg_dev, sizeof h, cudaMemcpyDeviceToHost);
        if (inv) {
            fprintf(stderr, "[gx dbg] level %d n %d blocks/colour<=%d: entry->tables %lld | half0: prologue %lld ext %lld q %lld inverse %lld sync %lld | half1: prologue %lld ext %lld q %lld inverse %lld sync %lld |",
                level, L.n, max_blocks, h[1] - h[0], h[2] - h[1], h[3] - h[2], h[4] - h[3], h[5] - h[4], h[6] - h[5], h[7] - h[6], h[8] - h[7], h[9] - h[8], h[10] - h[9], h[11] - h[10]);
            fprintf(stderr, "\n");
        }
        else
        fprintf(stderr, "[gs dbg] level %d n %d blocks/colour<=%d: half0 wait %lld zero %lld phaseA %lld phaseB %lld | half1 wait %lld zero %lld phaseA %lld phaseB %lld cycles\n",
            level, L.n, max_blocks, h[1] - h[0], h[2] - h[1], h[3] - h[2], h[4] - h[3], h[5] - h[4], h[6] - h[5], h[7] - h[6], h[8] - h[7]);
    }
    return 0;
}

} // namespace

// ---- vector ops -----------------------------------------------------------------------------------------------------------
int vec_axpy(Sim* s, long n, double a, const double* x, double* y)
{
    if (n <= 0) return 0;
    k_axpy<<<nblk(n), TPB, 0, s->stream>>>(n, a, x, y);
    HOT_LAUNCHED(s);
    return 0;
}
int vec_axpy_dev(Sim* s, long n, const double* num, const double* den, double sign, const double* x, double* y)
{
    if (n <= 0) return 0;
    k_axpy_dev<<<nblk(n), TPB, 0, s->stream>>>(n, num, den, sign, x, y);
    HOT_LAUNCHED(s);
    return 0;
}
int vec_xpay_dev(Sim* s, long n, const double* x, const double* num, const double* den, double* y)
{
    if (n <= 0) return 0;
    k_xpay_dev<<<nblk(n), TPB, 0, s->stream>>>(n, x, num, den, y);
    HOT_LAUNCHED(s);
    return 0;
}
int vec_copy(Sim* s, long n, const double* x, double* y)
{
    HOT_CUDA(cudaMemcpyAsync(y, x, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
    return 0;
}
int vec_zero(Sim* s, long n, double* y)
{
    HOT_CUDA(cudaMemsetAsync(y, 0, (size_t)n * sizeof(double), s->stream));
    return 0;
}
int vec_scale(Sim* s, long n, double a, double* y)
{
    if (n <= 0) return 0;
    k_scale<<<nblk(n), TPB, 0, s->stream>>>(n, a, y);
    HOT_LAUNCHED(s);
    return 0;
}
// dot over the DOF vectors of level 0 (n = 3 num_nodes): own nodes + all-reduce when the object is partitioned
int vec_dot(Sim* s, long n, const double* a, const double* b, double* dev_out, double* host_out)
{
    if (s->world <= 1 || s->dot_plain || n != 3L * s->num_nodes) return reduce_to<1>(s, n, DotF{a, b}, dev_out, host_out);
    HOT_CUDA(s->red_out.reserve(64));
    if (!dev_out) dev_out = s->red_out.p;
    int rc = reduce_to<1>(s, n, OwnDotF{a, b, s->own_node.p}, dev_out, nullptr);
    if (!rc) rc = dist_allreduce_buffer(s, dev_out, 1, 0);
    if (rc || !host_out) return rc;
    HOT_CUDA(cudaMemcpyAsync(s->h_red, dev_out, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    HOT_CUDA(cudaStreamSynchronize(s->stream));
    *host_out = s->h_red[0];
    return 0;
}

// ---- hierarchy --------------------------------------------------------------------------------------------------------------
int build_coord_map(Sim* s, MGLevel& L)
{
    const int n = L.n;
    HOT_CUDA(L.key_sorted.reserve(n));
    HOT_CUDA(L.id_sorted.reserve(n));
    HOT_CUDA(s->keys_alt.reserve(n));
    HOT_CUDA(s->cand_val.reserve(n));
    k_coord_keys<<<nblk(n), TPB, 0, s->stream>>>(n, L.coord.p, s->keys_alt.p, s->cand_val.p);
    HOT_LAUNCHED(s);
    return with_tmp(s, [&](void* t, size_t& b) {
        return cub::DeviceRadixSort::SortPairs(t, b, s->keys_alt.p, L.key_sorted.p, s->cand_val.p, L.id_sorted.p, n, 0, 36, s->stream);
    });
}

// MultigridBuilder::build
int columns_from_coords(Sim* s, MGLevel& L)
{
    k_cols_from_coords<<<nblk((long)L.n * W), TPB, 0, s->stream>>>(L.n, L.coord.p, L.key_sorted.p, L.id_sorted.p, L.col.p, L.val.p);
    HOT_LAUNCHED(s);
    return 0;
}

int level_estimate_2norm(Sim* s, int level, double* lmax_lmin)
{
    if (!s->mg_built || level < 0 || level >= s->mg_levels) return fail(s, "estimate2norm: bad level (call hot_build_mg first)");
    int rc = estimate2norm(s, level);
    if (rc) return rc;
    lmax_lmin[0] = s->levels[level]->lMax;
    lmax_lmin[1] = s->levels[level]->lMin;
    return 0;
}

int build_mg(Sim* s, int levels, int smoother, int coarse_solver, int Ainv, int times, int levelscale, double topomega)
{
    if (!s->matrix_built) return fail(s, "buildMultigrid: call hot_build_matrix first");
    if (levels < 1 || levels > 10) return fail(s, "Level depth exceeds 10! Too Deep!");
    if (!s->matrix_bcproject && levels > 1) return fail(s, "multigrid needs the BC-projected system (ImplicitSolver.h:339)");
    for (int k : {smoother, coarse_solver})
        if (!(k == 0 || k == 1 || k == 2 || k == 5 || k == 6))
            return fail(s, "No proper smoother is selected! (supported: 0 Jacobi, 1 optimal Jacobi, 2 PCG, 5 GS, 6 Chebyshev)");
    if (!(Ainv == 0 || Ainv == 1)) return fail(s, "The Dinv function picked doesn't exist.");
    KTime t(s, KC_TRANSFER);
    s->mg_levels = levels; s->mg_smoother = smoother; s->mg_coarse = coarse_solver; s->mg_Ainv = Ainv; s->mg_times = times;
    s->mg_levelscale = levelscale; s->mg_topomega = topomega;
    while ((int)s->levels.size() < levels) s->levels.push_back(new MGLevel);
    HOT_CUDA(s->red_out.reserve(64));
    const bool colors = smoother == 5 || coarse_solver == 5;
    RC(build_coord_map(s, *s->levels[0]));
    RC(finish_level(s, *s->levels[0], Ainv, colors));
    // estimate2norm where the Chebyshev smoother will run (MultigridPreconditioner.h:610-611,682-683)
    if ((coarse_solver == 6 && levels == 1) || (smoother == 6 && levels > 1)) RC(estimate2norm(s, 0));
    if (s->world > 1 && !s->ghost_ring) return fail(s, "buildMultigrid on a partitioned object needs the ghost ring (hot_set_ghost_ring)");
    for (int l = 0; l + 1 < levels; ++l) {
        if (l == 0 && s->world > 1) RC(coarsen_dist(s, *s->levels[0], *s->levels[1]));
        else RC(coarsen(s, *s->levels[l], *s->levels[l + 1]));
        RC(finish_level(s, *s->levels[l + 1], Ainv, colors));
        if ((coarse_solver == 6 && l + 2 == levels) || (smoother == 6 && l + 2 < levels)) RC(estimate2norm(s, l + 1));
    }
    for (int l = 0; l < levels; ++l)
        if (!colors) s->levels[l]->n_blocks = 0;
    s->mg_built = true;
    return 0;
}

int level_spmv(Sim* s, int level, const double* x, double* b)
{
    MGLevel& L = *s->levels[level];
    KTime t(s, KC_SPMV);
    k_spmv<0><<<nblk(32L * L.n), TPB, 0, s->stream>>>(L.n, L.col.p, L.val.p, x, b);
    HOT_LAUNCHED(s);
    // partitioned level 0: rows on ghost pages lack columns - every holder takes the authority's result (coarser levels are replicated)
    if (level == 0 && s->world > 1) return dist_takeover_shared(s, b, 3);
    return 0;
}
static int level_spmv_sub(Sim* s, int level, const double* x, double* b)
{
    MGLevel& L = *s->levels[level];
    KTime t(s, KC_SPMV);
    k_spmv<1><<<nblk(32L * L.n), TPB, 0, s->stream>>>(L.n, L.col.p, L.val.p, x, b);
    HOT_LAUNCHED(s);
    if (level == 0 && s->world > 1) return dist_takeover_shared(s, b, 3);
    return 0;
}
int level_restrict(Sim* s, int level, const double* fine, double* coarse)
{
    MGLevel& F = *s->levels[level];
    const int nc = s->levels[level + 1]->n;
    KTime t(s, KC_TRANSFER);
    k_restrict<<<nblk(32L * nc), TPB, 0, s->stream>>>(nc, F.rcol.p, F.rw.p, fine, coarse);
    HOT_LAUNCHED(s);
    // partitioned level 0: the R rows are masked to the fine nodes this rank counts, the replicated coarse vector is the sum over the ranks
    if (level == 0 && s->world > 1) return dist_allreduce_buffer(s, coarse, 3L * nc, 0);
    return 0;
}
int level_prolong(Sim* s, int level, const double* coarse, double* fine)
{
    MGLevel& F = *s->levels[level];
    KTime t(s, KC_TRANSFER);
    k_prolong<<<nblk(F.n), TPB, 0, s->stream>>>(F.n, F.pcol.p, F.pw.p, coarse, fine);
    HOT_LAUNCHED(s);
    return 0;
}

// smoothFunc(u, r, du, dAu, A, iterations, tolerance) with the integer codes of -smoother / -coarseSolver
int level_smooth(Sim* s, int level, int kind, double* u, double* r, int iterations, double tolerance)
{
    KTime t(s, KC_GS);
    s->dot_plain = level > 0; // partitioned object: the coarse levels are replicated, their dots need no mask / all-reduce
    int rc;
    switch (kind) {
    case 0: rc = smooth_jacobi(s, level, u, r, iterations); break;
    case 1: rc = smooth_optimal_jacobi(s, level, u, r, iterations, tolerance); break;
    case 2: rc = smooth_cg(s, level, u, r, iterations); break;
    case 5: rc = smooth_gs(s, level, u, r, iterations); break;
    case 6: rc = smooth_chebyshev(s, level, u, r, iterations); break;
    default: rc = fail(s, "No proper smoother is selected! (supported: 0 Jacobi, 1 optimal Jacobi, 2 PCG, 5 GS, 6 Chebyshev)");
    }
    s->dot_plain = false;
    return rc;
}

// MultigridOperator::operator(), MultigridPreconditioner.h:362-421.  `out` must not alias `in`.
int vcycle(Sim* s, const double* in, double* out, bool timed)
{
    if (!s->mg_built) return fail(s, "vcycle: call hot_build_mg first");
    cudaStream_t st = s->stream;
    const int Lc = s->mg_levels;
    std::vector<MGLevel*>& lv = s->levels;
    struct Seg { int level, what; cudaEvent_t a, b; };
    std::vector<Seg> segs;
    auto begin = [&](int level, int what) {
        if (!timed) return;
        Seg g{level, what, s->timers.get(), s->timers.get()};
        cudaEventRecord(g.a, st);
        segs.push_back(g);
    };
    auto end = [&]() {
        if (timed) cudaEventRecord(segs.back().b, st);
    };
    RC(vec_copy(s, 3L * lv[0]->n, in, lv[0]->residual.p)); // correctResidualProjection adds dRhs == 0
    RC(vec_zero(s, 3L * lv[0]->n, out));
    if (Lc > 1) RC(level_restrict(s, 0, lv[0]->residual.p, lv[1]->initial_residual.p));
    else RC(vec_copy(s, 3L * lv[0]->n, lv[0]->residual.p, lv[0]->initial_residual.p));
    for (int l = 1; l < Lc - 1; ++l) RC(level_restrict(s, l, lv[l]->initial_residual.p, lv[l + 1]->initial_residual.p));
    int level;
    for (level = 0; level < Lc - 1; ++level) {
        double* sol = level == 0 ? out : lv[level]->sol.p;
        begin(level, 0);
        RC(level_smooth(s, level, s->mg_smoother, sol, lv[level]->residual.p, regular_iters(s, level), 0.0));
        end();
        begin(level, 1);
        RC(level_restrict(s, level, lv[level]->residual.p, lv[level + 1]->residual.p));
        end();
        RC(vec_zero(s, 3L * lv[level + 1]->n, lv[level + 1]->sol.p));
    }
    begin(level, 0);
    // top.tolFunc(level) = cneps * cneps (MultigridPreconditioner.h:529,399): optimal Jacobi stops early on it, PCG derives its own
    RC(level_smooth(s, level, s->mg_coarse, level == 0 ? out : lv[level]->sol.p, lv[level]->residual.p, top_iters(s, level), s->mg_cneps * s->mg_cneps));
    end();
    for (--level; level >= 0; --level) {
        double* sol = level == 0 ? out : lv[level]->sol.p;
        begin(level, 2);
        RC(level_prolong(s, level, lv[level + 1]->sol.p, lv[level]->du.p));
        end();
        begin(level, 3);
        RC(vec_axpy(s, 3L * lv[level]->n, 1.0, lv[level]->du.p, sol));
        RC(level_spmv_sub(s, level, lv[level]->du.p, lv[level]->residual.p)); // residual -= A du (dAu fused away)
        end();
        begin(level, 0);
        RC(level_smooth(s, level, s->mg_smoother, sol, lv[level]->residual.p, regular_iters(s, level), 0.0));
        end();
    }
    if (timed) {
        for (int i = 0; i < 10; ++i)
            for (int j = 0; j < 4; ++j) s->vc_ms[i][j] = 0.0;
        HOT_CUDA(cudaStreamSynchronize(st));
        for (Seg& g : segs) {
            float ms = 0;
            cudaEventElapsedTime(&ms, g.a, g.b);
            if (g.level < 10) s->vc_ms[g.level][g.what] += ms;
            s->timers.pool.push_back(g.a);
            s->timers.pool.push_back(g.b);
        }
    }
    return 0;
}

} // namespace hot
