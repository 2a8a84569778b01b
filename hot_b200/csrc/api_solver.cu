// extern "C" entry points of the assembled-matrix / multigrid / solver part of include/hot_b200.h: host-buffer
// marshalling around the device-resident operators of matrix.cu, multigrid.cu and solver.cu.
#include "api_internal.h"
#include <algorithm>
#include "reduce.cuh"

using namespace hot;

namespace {

constexpr int TPB = 256;
inline int nblk(long n) { return (int)((n + TPB - 1) / TPB); }
constexpr int W = MGLevel::W;

template <class T>
int d2h(hot_sim* s, T* host, const T* dev, size_t n)
{
    if (n == 0) return 0;
    HOT_CUDA(cudaMemcpyAsync(host, dev, n * sizeof(T), cudaMemcpyDeviceToHost, s->stream));
    HOT_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}
template <class T>
int h2d(hot_sim* s, T* dev, const T* host, size_t n)
{
    if (n == 0) return 0;
    HOT_CUDA(cudaMemcpyAsync(dev, host, n * sizeof(T), cudaMemcpyHostToDevice, s->stream));
    return 0;
}

// padded device rows -> the reference's entryCol / entryVal arrays (n x 125, blocks column-major)
__global__ void k_export_rows(int n, const int* __restrict__ col, const double* __restrict__ val, int* __restrict__ ecol, double* __restrict__ eval)
{
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)n * 125) return;
    const int i = (int)(t / 125), s = (int)(t - (long)i * 125);
    if (ecol) ecol[t] = col[(size_t)i * W + s];
    if (eval)
#pragma unroll
        for (int q = 0; q < 9; ++q) eval[9 * (size_t)t + q] = val[((size_t)i * 9 + q) * W + s];
}

struct NnzF { // entries whose column is not the row itself (empty slots alias the row, the diagonal is counted apart)
    const int* col;
    __device__ void operator()(long t, double (&acc)[1]) const { acc[0] += col[t] != (int)(t / W) ? 1.0 : 0.0; }
};

int check_level(hot_sim* s, int level, const char* who, bool need_coarser = false)
{
    if (!s->mg_built) return fail(s, std::string(who) + ": call hot_build_mg first");
    if (level < 0 || level + (need_coarser ? 1 : 0) >= s->mg_levels) return fail(s, std::string(who) + ": bad level");
    return 0;
}

} // namespace

extern "C" {

int hot_build_matrix(hot_sim* s, int bcproject) { return build_matrix(s, bcproject != 0); }

int hot_get_matrix(hot_sim* s, int* entryCol, double* entryVal)
{
    if (!s->matrix_built) return fail(s, "hot_get_matrix: call hot_build_matrix first");
    MGLevel& L = *s->levels[0];
    const size_t ne = (size_t)L.n * 125;
    HOT_CUDA(s->stage_i.reserve(ne));
    HOT_CUDA(s->stage.reserve(9 * ne));
    k_export_rows<<<nblk((long)ne), TPB, 0, s->stream>>>(L.n, L.col.p, L.val.p, s->stage_i.p, s->stage.p);
    HOT_LAUNCHED(s);
    int rc = 0;
    if (entryCol) rc = d2h(s, entryCol, s->stage_i.p, ne);
    if (!rc && entryVal) rc = d2h(s, entryVal, s->stage.p, 9 * ne);
    return rc;
}

int hot_build_diagonal(hot_sim* s, int Ainv, double* diag_inv)
{
    int rc = build_diagonal_mf(s, Ainv);
    if (rc) return rc;
    if (diag_inv) return d2h(s, diag_inv, s->diag_mf.p, 9 * (size_t)s->num_nodes);
    return 0;
}

int hot_build_mg(hot_sim* s, int levels, int smoother, int coarse_solver, int Ainv, int times, int levelscale, double topomega)
{
    return build_mg(s, levels, smoother, coarse_solver, Ainv, times, levelscale, topomega);
}
int hot_mg_levels(hot_sim* s) { return s->mg_built ? s->mg_levels : 0; }
int hot_estimate_2norm(hot_sim* s, int level, double* lmax_lmin) { return level_estimate_2norm(s, level, lmax_lmin); }
int hot_get_level_dofs(hot_sim* s, int* dofs)
{
    if (!s->mg_built) return fail(s, "hot_get_level_dofs: call hot_build_mg first");
    for (int l = 0; l < s->mg_levels; ++l) dofs[l] = s->levels[l]->n;
    return 0;
}
// structurally non-zero 3x3 blocks of A_level (the nnzb of the SpMV roofline): stored neighbours + the diagonal
int hot_level_nnz_blocks(hot_sim* s, int level, long long* nnzb)
{
    if (!s->matrix_built || level < 0 || level >= (int)s->levels.size() || (level > 0 && !s->mg_built)) return fail(s, "hot_level_nnz_blocks: bad level");
    MGLevel& L = *s->levels[level];
    double h = 0;
    int rc = reduce_to<1>(s, (long)L.n * W, NnzF{L.col.p}, nullptr, &h);
    if (rc) return rc;
    *nnzb = (long long)(h + 0.5) + L.n;
    return 0;
}
int hot_get_level_coords(hot_sim* s, int level, int* coord)
{
    int rc = check_level(s, level, "hot_get_level_coords");
    if (rc) return rc;
    return d2h(s, coord, s->levels[level]->coord.p, 3 * (size_t)s->levels[level]->n);
}
int hot_get_level_matrix(hot_sim* s, int level, int kind, int* colsize, int* col, double* val)
{
    int rc = check_level(s, level, "hot_get_level_matrix", kind != 0);
    if (rc) return rc;
    MGLevel& L = *s->levels[level];
    if (kind == 0) {
        if (colsize) *colsize = 125;
        if (!col && !val) return 0;
        const size_t ne = (size_t)L.n * 125;
        HOT_CUDA(s->stage_i.reserve(ne));
        HOT_CUDA(s->stage.reserve(9 * ne));
        k_export_rows<<<nblk((long)ne), TPB, 0, s->stream>>>(L.n, L.col.p, L.val.p, s->stage_i.p, s->stage.p);
        HOT_LAUNCHED(s);
        if (col) rc = d2h(s, col, s->stage_i.p, ne);
        if (!rc && val) rc = d2h(s, val, s->stage.p, 9 * ne);
        return rc;
    }
    // transfer operators carry scalar weights (w * I3 in the reference): val = one weight per entry
    if (kind == 1) {
        if (colsize) *colsize = 8;
        if (col) rc = d2h(s, col, L.pcol.p, (size_t)L.n * 8);
        if (!rc && val) rc = d2h(s, val, L.pw.p, (size_t)L.n * 8);
        return rc;
    }
    if (kind == 2) {
        const size_t nc = s->levels[level + 1]->n;
        if (colsize) *colsize = 32;
        if (col) rc = d2h(s, col, L.rcol.p, nc * 32);
        if (!rc && val) rc = d2h(s, val, L.rw.p, nc * 32);
        return rc;
    }
    return fail(s, "hot_get_level_matrix: kind must be 0 (system), 1 (prolongation) or 2 (restriction)");
}
int hot_get_level_diagonal(hot_sim* s, int level, double* diagonalVal, double* diagonalInv)
{
    int rc = check_level(s, level, "hot_get_level_diagonal");
    if (rc) return rc;
    MGLevel& L = *s->levels[level];
    if (diagonalVal) rc = d2h(s, diagonalVal, L.diag.p, 9 * (size_t)L.n);
    if (!rc && diagonalInv) rc = d2h(s, diagonalInv, L.dinv.p, 9 * (size_t)L.n);
    return rc;
}
int hot_get_gs_schedule(hot_sim* s, int level, int* n_blocks, int* color_first_block9, int* seq, int* block_start)
{
    int rc = check_level(s, level, "hot_get_gs_schedule");
    if (rc) return rc;
    MGLevel& L = *s->levels[level];
    if (n_blocks) *n_blocks = L.n_blocks;
    if (color_first_block9)
        for (int c = 0; c < 9; ++c) color_first_block9[c] = L.color_first_block[c];
    if (L.n_blocks <= 0) return 0;
    if (seq) rc = d2h(s, seq, L.gs_seq.p, (size_t)L.n);
    if (!rc && block_start) rc = d2h(s, block_start, L.gs_block_start.p, (size_t)L.n_blocks + 1);
    return rc;
}

int hot_spmv(hot_sim* s, int level, const double* x, double* b)
{
    if (!s->matrix_built) return fail(s, "hot_spmv: call hot_build_matrix first");
    if (level != 0 || s->mg_built) {
        int rc = check_level(s, level, "hot_spmv");
        if (rc) return rc;
    }
    const size_t m = 3 * (size_t)s->levels[level]->n;
    HOT_CUDA(s->work[1].reserve(m));
    HOT_CUDA(s->work[2].reserve(m));
    int rc = h2d(s, s->work[1].p, x, m);
    if (!rc) rc = level_spmv(s, level, s->work[1].p, s->work[2].p);
    if (rc) return rc;
    return d2h(s, b, s->work[2].p, m);
}
int hot_restrict(hot_sim* s, int level, const double* fine, double* coarse)
{
    int rc = check_level(s, level, "hot_restrict", true);
    if (rc) return rc;
    const size_t mf = 3 * (size_t)s->levels[level]->n, mc = 3 * (size_t)s->levels[level + 1]->n;
    HOT_CUDA(s->work[1].reserve(mf));
    HOT_CUDA(s->work[2].reserve(mc));
    rc = h2d(s, s->work[1].p, fine, mf);
    if (!rc) rc = level_restrict(s, level, s->work[1].p, s->work[2].p);
    if (rc) return rc;
    return d2h(s, coarse, s->work[2].p, mc);
}
int hot_prolong(hot_sim* s, int level, const double* coarse, double* fine)
{
    int rc = check_level(s, level, "hot_prolong", true);
    if (rc) return rc;
    const size_t mf = 3 * (size_t)s->levels[level]->n, mc = 3 * (size_t)s->levels[level + 1]->n;
    HOT_CUDA(s->work[1].reserve(mc));
    HOT_CUDA(s->work[2].reserve(mf));
    rc = h2d(s, s->work[1].p, coarse, mc);
    if (!rc) rc = level_prolong(s, level, s->work[1].p, s->work[2].p);
    if (rc) return rc;
    return d2h(s, fine, s->work[2].p, mf);
}
int hot_smooth(hot_sim* s, int level, int kind, double* u, double* r, int iterations, double tolerance, const double* initial_residual)
{
    int rc = check_level(s, level, "hot_smooth");
    if (rc) return rc;
    MGLevel& L = *s->levels[level];
    const size_t m = 3 * (size_t)L.n;
    HOT_CUDA(s->work[1].reserve(m));
    HOT_CUDA(s->work[2].reserve(m));
    rc = h2d(s, s->work[1].p, (const double*)u, m);
    if (!rc) rc = h2d(s, s->work[2].p, (const double*)r, m);
    if (!rc && initial_residual) rc = h2d(s, L.initial_residual.p, initial_residual, m);
    if (!rc) rc = level_smooth(s, level, kind, s->work[1].p, s->work[2].p, iterations, tolerance);
    if (rc) return rc;
    rc = d2h(s, u, s->work[1].p, m);
    if (!rc) rc = d2h(s, r, s->work[2].p, m);
    return rc;
}
int hot_vcycle(hot_sim* s, const double* in, double* out)
{
    if (!s->mg_built) return fail(s, "hot_vcycle: call hot_build_mg first");
    const size_t m = 3 * (size_t)s->levels[0]->n;
    HOT_CUDA(s->work[1].reserve(m));
    HOT_CUDA(s->work[2].reserve(m));
    int rc = h2d(s, s->work[1].p, in, m);
    if (!rc) rc = vcycle(s, s->work[1].p, s->work[2].p, true);
    if (rc) return rc;
    return d2h(s, out, s->work[2].p, m);
}
int hot_vcycle_timing(hot_sim* s, double* ms40, int* coarse_cg_iters)
{
    for (int i = 0; i < 10; ++i)
        for (int j = 0; j < 4; ++j) ms40[4 * i + j] = s->vc_ms[i][j];
    if (coarse_cg_iters) *coarse_cg_iters = s->last_cg_iters;
    return 0;
}
// device-resident V-cycles for benchmarking: `reps` applications on a fixed right-hand side already on the device
int hot_vcycle_bench(hot_sim* s, int reps, double* ms_total)
{
    if (!s->mg_built) return fail(s, "hot_vcycle_bench: call hot_build_mg first");
    const size_t m = 3 * (size_t)s->levels[0]->n;
    HOT_CUDA(s->work[1].reserve(m));
    HOT_CUDA(s->work[2].reserve(m));
    // L2 flushed before every cycle (256 MiB memset outside the cycle's event pair)
    HOT_CUDA(s->l2_flush.reserve((size_t)256 << 20));
    std::vector<cudaEvent_t> ev;
    for (int i = 0; i < reps; ++i) {
        HOT_CUDA(cudaMemsetAsync(s->l2_flush.p, i & 0xff, (size_t)256 << 20, s->stream));
        ev.push_back(s->timers.get());
        cudaEventRecord(ev.back(), s->stream);
        int rc = vcycle(s, s->work[1].p, s->work[2].p, false);
        if (rc) return rc;
        ev.push_back(s->timers.get());
        cudaEventRecord(ev.back(), s->stream);
    }
    HOT_CUDA(cudaStreamSynchronize(s->stream));
    double total = 0;
    for (size_t k = 0; k + 1 < ev.size(); k += 2) {
        float ms = 0;
        cudaEventElapsedTime(&ms, ev[k], ev[k + 1]);
        total += ms;
    }
    for (cudaEvent_t e : ev) s->timers.pool.push_back(e);
    if (ms_total) *ms_total = total;
    return 0;
}

// `reps` device-resident applications of one operator of the path, timed with CUDA events on the handle's stream:
// op 0 matrix-free Hessian apply (a13), 1 block SpMV on `level` (a16), 2 updateState without energy (a9-a11),
// 3 computeResidual (a12), 4 one smoother call of the configured -smoother on `level` (a18), 5 hot_build_matrix (a15),
// 6 hot_build_mg with the current settings (a17), 7 one -coarseSolver call on `level` with the level's restricted initial
// residual of the last V-cycle as right-hand side (a19; the iteration count is returned by hot_vcycle_timing)
int hot_op_bench(hot_sim* s, int op, int level, int reps, double* ms_total)
{
    if (!s->state_valid) return fail(s, "hot_op_bench: call hot_update_state first");
    const bool on_level = op == 1 || op == 4 || op == 7;
    if (on_level && (level < 0 || level >= (int)s->levels.size() || !s->matrix_built || (op != 1 && !s->mg_built)))
        return fail(s, "hot_op_bench: matrix / hierarchy not built or bad level");
    if (op == 6 && !s->matrix_built) return fail(s, "hot_op_bench: matrix not built");
    const size_t m = 3 * (size_t)(on_level ? s->levels[level]->n : s->num_nodes);
    HOT_CUDA(s->work[3].reserve(m));
    HOT_CUDA(s->work[4].reserve(m));
    HOT_CUDA(cudaMemsetAsync(s->work[4].p, 0, m * sizeof(double), s->stream));
    if (op != 4 && op != 7) HOT_CUDA(cudaMemcpyAsync(s->work[3].p, s->dv.p, std::min(m, 3 * (size_t)s->num_nodes) * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
    if (op == 7) {
        // the -coarseSolver call on its own: the level's restricted initial residual of the last V-cycle is both the right-hand
        // side and cg_smooth's stopping reference, so z.r starts AT z0.r0 and the solver has to bring it below the 0.25 z0.r0
        // of MultigridPreconditioner.h:209 (inside a V-cycle the pre-smoothing usually achieves that already: 0 iterations)
        HOT_CUDA(s->work[5].reserve(m));
        HOT_CUDA(cudaMemcpyAsync(s->work[5].p, s->levels[level]->initial_residual.p, m * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
    }
    // iteration -1 is an untimed warm-up: first-use allocations (cudaMalloc blocks the host while the stream idles) and
    // once-per-linearisation work (the contracted particle Hessian of ensure_hessian) stay out of the per-application time.
    // L2 (126 MB) is flushed before every timed application (256 MiB memset, outside the event pair of the application): the
    // coarse levels' matrices would otherwise be served from L2 on every repetition after the first.
    HOT_CUDA(s->l2_flush.reserve((size_t)256 << 20));
    std::vector<cudaEvent_t> ev;
    for (int i = -1; i < reps; ++i) {
        if (op == 7) { // fresh right-hand side / zero solution for every application (outside the event pair)
            HOT_CUDA(cudaMemcpyAsync(s->work[3].p, s->work[5].p, m * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
            HOT_CUDA(cudaMemsetAsync(s->work[4].p, 0, m * sizeof(double), s->stream));
        }
        if (i >= 0) {
            HOT_CUDA(cudaMemsetAsync(s->l2_flush.p, i & 0xff, (size_t)256 << 20, s->stream));
            ev.push_back(s->timers.get());
            cudaEventRecord(ev.back(), s->stream);
        }
        int rc = 0;
        switch (op) {
        case 7: rc = level_smooth(s, level, s->mg_coarse, s->work[4].p, s->work[3].p, (s->mg_coarse == 2 || s->mg_coarse == 6) ? 10000 : 3 * s->mg_times,
                                  s->mg_cneps * s->mg_cneps); break;
        case 0: rc = hessian_apply_mf(s, s->work[3].p, s->work[4].p); break;
        case 1: rc = level_spmv(s, level, s->work[3].p, s->work[4].p); break;
        case 2: rc = update_state(s, false, nullptr); break;
        case 3: rc = compute_residual(s, s->work[4].p); break;
        case 4: rc = level_smooth(s, level, s->mg_smoother, s->work[4].p, s->work[3].p, s->mg_times, 0.0); break;
        case 5: rc = build_matrix(s, s->matrix_bcproject); break;
        case 6: rc = build_mg(s, s->mg_levels, s->mg_smoother, s->mg_coarse, s->mg_Ainv, s->mg_times, s->mg_levelscale, s->mg_topomega); break;
        default: rc = fail(s, "hot_op_bench: unknown op");
        }
        if (rc) return rc;
        if (i >= 0) {
            ev.push_back(s->timers.get());
            cudaEventRecord(ev.back(), s->stream);
        }
    }
    HOT_CUDA(cudaStreamSynchronize(s->stream));
    double total = 0;
    for (size_t k = 0; k + 1 < ev.size(); k += 2) {
        float ms = 0;
        cudaEventElapsedTime(&ms, ev[k], ev[k + 1]);
        total += ms;
    }
    for (cudaEvent_t e : ev) s->timers.pool.push_back(e);
    if (ms_total) *ms_total = total;
    return 0;
}

} // extern "C"
