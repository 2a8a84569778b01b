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
    const double m = s == 62 ? mass[i] : 0.0;
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
            if (tid < 81) {
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
#pragma unroll
                    for (int q = 0; q < 9; ++q) atomicAdd(val + ((size_t)ida * 9 + q) * W + slot, dt2 * acc[b][q]);
                }
            }
        }
    }
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
    if (s->world > 1) return fail(s, "buildMatrix: the assembled-matrix / multigrid path is single-GPU in this version; partitioned runs use --matfree (lsolver 2)");
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
    k_matrix_init<<<nblk((long)nn * W), TPB, 0, st>>>(nn, s->mass_matrix.p, L.col.p, L.val.p);
    HOT_LAUNCHED(s);
    k_assemble<<<(unsigned)s->n_groups, AS_THREADS, 0, st>>>(s->cell_start.p, s->group_slot.p, s->nbr8.p, s->P.stride, s->P.X.p, s->f_H.p,
        s->dx, 1.0 / s->dx, s->dt * s->dt, s->g_idx.p, L.col.p, L.val.p);
    HOT_LAUNCHED(s);
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
