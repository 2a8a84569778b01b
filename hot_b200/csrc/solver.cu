// a21, a22, a24: the outer solvers as host C++ driving device-resident vectors - inexact PCG, extended Newton, L-BFGS
// around the V-cycle, the objective's line search / CN exit test, and the backward-Euler glue.
//
// Reference: InexactConjugateGradient::solve (Lib/Ziran/Math/Linear/InexactConjugateGradient.h:49-103),
// ExtendedNewtonsMethod::solve (Lib/Ziran/Math/Nonlinear/ExtendedNewtonsMethod.h:39-66), LBFGS::solve
// (Lib/Ziran/Math/Nonlinear/LBFGS.h:300-437, RingBuffer :23-69), ImplicitSolverObjective::{shouldExitByCN, lineSearch,
// HinvApproxInit, computeStep} (Projects/multigrid/ImplicitSolver.h:174-211,312-432), MultigridSimulation::
// {computeCharacteristicNorm, startBackwardEuler, backwardEulerStep} (Projects/multigrid/MultigridSimulation.h:128-233).
//
// All vectors stay in HBM; the host sees one scalar per convergence test (pinned read-back), step lengths alpha / beta
// are consumed on the device (vec_axpy_dev / vec_xpay_dev), so a PCG iteration costs one stream synchronisation.
#include "api_internal.h"
#include "reduce.cuh"
#include "dense3.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>

namespace hot {
namespace {

constexpr int TPB = 256;
inline int nblk(long n) { return (int)((n + TPB - 1) / TPB); }

#define RC(x)                  \
    do {                       \
        int rc__ = (x);        \
        if (rc__) return rc__; \
    } while (0)

// sum_i |r_i|^2 / tol_i^2  (shouldExitByCN, ImplicitSolver.h:192-197) and sum |r_i|^2
struct CNNormF {
    const double *r, *tol;
    const unsigned char* own; // partitioned object: shared nodes count on one rank
    __device__ void operator()(long i, double (&acc)[2]) const
    {
        if (own && !own[i]) return;
        const double a = r[3 * i], b = r[3 * i + 1], c = r[3 * i + 2], n2 = a * a + b * b + c * c, t = tol ? tol[i] : 1.0;
        acc[0] += n2 / (t * t);
        acc[1] += n2;
    }
};
// max_p ||dPdF(F = I)||_F  (computeCharacteristicNorm, MultigridSimulation.h:136-152); positive doubles order like uint64
__global__ void k_max_dpdf_norm(long n, const double* __restrict__ mu, const double* __restrict__ lam, int project, unsigned long long* out)
{
    const long s = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const double one[3] = {1.0, 1.0, 1.0};
    HessBlocks hb;
    model_blocks(one, mu[s], lam[s], project, hb);
    double n2 = 0.0;
#pragma unroll
    for (int q = 0; q < 9; ++q) n2 += hb.A[q] * hb.A[q];
#pragma unroll
    for (int q = 0; q < 4; ++q) n2 += hb.B01[q] * hb.B01[q] + hb.B12[q] * hb.B12[q] + hb.B20[q] * hb.B20[q];
    atomicMax(out, (unsigned long long)__double_as_longlong(sqrt(n2)));
}
__global__ void k_waxpy(long n, const double* __restrict__ x, double a, const double* __restrict__ y, double* __restrict__ w)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) w[i] = x[i] + a * y[i];
}
__global__ void k_sub(long n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] - b[i];
}

// ImplicitSolverObjective state inside one backward-Euler step (ImplicitSolver.h:58-71)
struct Objective {
    hot_sim* s;
    hot_solver_options opt;
    hot_solve_log* log;
    long m; // 3 * num_nodes
    double Ek = 0;
    bool updated = false;
    int precond = 0; // 0 identity, 1 matrix-free block Jacobi, 2 level-0 diagonal, 3 V-cycle
    double* dv0; // accepted iterate
    double* dvnew;
    double* vec(int k) { return s->sv[k].p; }
};

enum { V_DV0 = 0, V_DVNEW, V_RES, V_STEP, V_R, V_P, V_Q, V_TEMP, V_RING0 /* 18 vectors */, V_COUNT = V_RING0 + 18 };

int waxpy(Objective& O, const double* x, double a, const double* y, double* w)
{
    Sim* s = O.s;
    k_waxpy<<<nblk(O.m), TPB, 0, s->stream>>>(O.m, x, a, y, w);
    HOT_LAUNCHED(s);
    return 0;
}

int residual_norms(Objective& O, const double* r, double* l2, double* scaled)
{
    Sim* s = O.s;
    double h[2];
    RC(reduce_to<2>(s, s->num_nodes, CNNormF{r, O.opt.usecn ? s->cn_tol.p : nullptr, s->world > 1 ? s->own_node.p : nullptr}, s->red_out.p + 16, h));
    RC(dist_allreduce_host(s, h, 2, 0));
    *l2 = sqrt(h[1]);
    *scaled = h[0];
    return 0;
}

// shouldExitByCN, ImplicitSolver.h:174-211
int should_exit_by_cn(Objective& O, const double* residual, bool* exit, double* l2_out)
{
    Sim* s = O.s;
    const long nn = s->world > 1 ? s->global_nodes : s->num_nodes; // N_n of the whole object
    double l2, scaled;
    RC(residual_norms(O, residual, &l2, &scaled));
    if (l2_out) *l2_out = l2;
    hot_solve_log* L = O.log;
    if (L && L->n_log < HOT_LOG_CAP) {
        L->residual_norm[L->n_log] = l2;
        L->scaled_norm[L->n_log] = O.opt.usecn ? (nn ? sqrt(scaled / nn) : 0.0) : l2;
        L->energy[L->n_log] = O.Ek;
        L->linear_iterations[L->n_log] = 0;
        L->n_log++;
    }
    if (!O.opt.usecn) *exit = l2 < O.opt.cneps;
    else *exit = nn == 0 || scaled < nn;
    return 0;
}

// updateState / computeResidual with the `updated` short-circuit (ImplicitSolver.h:128-155,237-252).  x == s->dv.p is the
// reference's aliasing case (moveNodes returns early, MpmSimulationBase.cpp:738-739).
int obj_update_state(Objective& O, const double* x, bool force = false)
{
    Sim* s = O.s;
    if (O.updated && !force) return 0;
    if (x != s->dv.p) RC(vec_copy(s, O.m, x, s->dv.p));
    double e = 0;
    RC(update_state(s, O.opt.linesearch != 0, &e));
    if (O.opt.linesearch) O.Ek = e;
    return 0;
}
int obj_compute_residual(Objective& O, double* r, bool force = false)
{
    if (O.updated && !force) return 0;
    return compute_residual(O.s, r);
}

// lineSearch, ImplicitSolver.h:312-333 (halving capped at 60 probes so a NaN energy cannot hang the caller)
int line_search(Objective& O, double* ddv, double* residual, double alpha)
{
    Sim* s = O.s;
    RC(bc_rotate(s, ddv, true)); // recoverSolution
    const double Ek0 = O.Ek;
    int probes = 0;
    do {
        RC(waxpy(O, O.dv0, alpha, ddv, O.dvnew));
        RC(obj_update_state(O, O.dvnew, true));
        alpha *= 0.5;
        if (O.log) O.log->total_linesearch_probes++;
    } while (O.Ek > Ek0 && ++probes < 60);
    alpha *= 2;
    RC(vec_scale(s, O.m, alpha, ddv));
    RC(bc_rotate(s, ddv, false)); // transformResidual
    RC(obj_compute_residual(O, residual, true));
    O.updated = true;
    std::swap(O.dv0, O.dvnew); // dv0 = dvnew
    return 0;
}

int obj_multiply(Objective& O, const double* x, double* b)
{
    if (O.opt.matfree) return hessian_apply_mf(O.s, x, b);
    return level_spmv(O.s, 0, x, b);
}
int obj_precondition(Objective& O, const double* in, double* out)
{
    Sim* s = O.s;
    switch (O.precond) {
    case 0: return vec_copy(s, O.m, in, out);
    case 1: return apply_block_diag(s, s->num_nodes, s->diag_mf.p, in, out);
    case 2: return apply_block_diag(s, s->num_nodes, s->levels[0]->dinv.p, in, out);
    default: return vcycle(s, in, out, false);
    }
}

// InexactConjugateGradient::solve
int inexact_pcg(Objective& O, double* x, const double* b, double tolerance, int max_iterations, int* iters)
{
    Sim* s = O.s;
    const long m = O.m;
    double *r = O.vec(V_R), *p = O.vec(V_P), *q = O.vec(V_Q), *temp = O.vec(V_TEMP);
    double* sc = s->red_out.p + 24; // device scalars: [0] zTrk, [1] p.Ap, [2] zTrk_last
    KTime t(s, KC_BLAS1);
    RC(obj_multiply(O, x, temp));
    k_sub<<<nblk(m), TPB, 0, s->stream>>>(m, b, temp, r);
    HOT_LAUNCHED(s);
    RC(bc_project(s, r));
    RC(obj_precondition(O, r, q));
    RC(vec_copy(s, m, q, p));
    double zTrk;
    RC(vec_dot(s, m, r, q, sc, &zTrk));
    double rpn = sqrt(zTrk);
    const double forcing = std::min(0.5, sqrt(std::max(rpn, tolerance)));
    const double local_tolerance = forcing * rpn;
    int cnt;
    for (cnt = 0; cnt < max_iterations; ++cnt) {
        if (rpn < local_tolerance) break;
        RC(obj_multiply(O, p, temp));
        RC(bc_project(s, temp));
        RC(vec_dot(s, m, temp, p, sc + 1, nullptr));
        RC(vec_axpy_dev(s, m, sc, sc + 1, 1.0, p, x));
        RC(vec_axpy_dev(s, m, sc, sc + 1, -1.0, temp, r));
        RC(obj_precondition(O, r, q));
        HOT_CUDA(cudaMemcpyAsync(sc + 2, sc, sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
        RC(vec_dot(s, m, q, r, sc, &zTrk));
        RC(vec_xpay_dev(s, m, q, sc, sc + 2, p)); // p = q + beta p
        rpn = sqrt(zTrk);
    }
    if (iters) *iters = cnt;
    return 0;
}


// Minres::solve, Lib/Ziran/Math/Linear/Minres.h:71-176 (-lsolver 1); Givens rotations of Givens.h:73-141 on the host, the seven
// Lanczos / direction vectors in HBM (the L-BFGS ring slots, unused by a Newton solve); two scalar read-backs per iteration
__global__ void k_minres_m(long n, double delta, double epsilon, double inv_gamma, const double* __restrict__ mkm1, const double* __restrict__ mkm2,
    double tk, double* __restrict__ mk, double* __restrict__ x)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = (mk[i] - delta * mkm1[i] - epsilon * mkm2[i]) * inv_gamma;
    mk[i] = v;
    x[i] += tk * v;
}
struct GivensRot {
    double c = 1, s = 0;
    void compute(double a, double b)
    {
        const double d = a * a + b * b, sq = std::sqrt(d);
        c = 1; s = 0;
        if (sq) { const double t = 1 / sq; c = a * t; s = -b * t; }
    }
    void row_rotation(double (&v)[2]) const
    {
        const double t1 = v[0], t2 = v[1];
        v[0] = c * t1 - s * t2;
        v[1] = s * t1 + c * t2;
    }
};
int minres_solve(Objective& O, double* x, const double* b, double relative_tolerance, double tolerance, int max_iterations, int* iters)
{
    Sim* s = O.s;
    const long m = O.m;
    double *mk = O.vec(V_RING0), *mkm1 = O.vec(V_RING0 + 1), *mkm2 = O.vec(V_RING0 + 2), *z = O.vec(V_RING0 + 3), *qkp1 = O.vec(V_RING0 + 4),
           *qk = O.vec(V_RING0 + 5), *qkm1 = O.vec(V_RING0 + 6);
    double* sc = s->red_out.p + 24;
    KTime t(s, KC_BLAS1);
    for (double* v : {mk, mkm1, mkm2, qk, qkm1}) RC(vec_zero(s, m, v));
    GivensRot Gk, Gkm1, Gkm2;
    double gamma = 0, delta = 0, epsilon = 0, beta_kp1 = 0, alpha_k = 0, beta_k = 0, tk = 0, d;
    RC(obj_multiply(O, x, qkp1));
    k_sub<<<nblk(m), TPB, 0, s->stream>>>(m, b, qkp1, qkp1);
    HOT_LAUNCHED(s);
    RC(bc_project(s, qkp1));
    RC(obj_precondition(O, qkp1, z));
    RC(vec_dot(s, m, z, qkp1, sc, &d));
    double rpn = std::sqrt(d);
    beta_kp1 = rpn;
    const double local_tolerance = std::min(relative_tolerance * rpn, tolerance);
    if (iters) *iters = 0;
    if (rpn < local_tolerance) return 0;
    if (rpn > 0) {
        RC(vec_scale(s, m, 1.0 / beta_kp1, qkp1));
        RC(vec_scale(s, m, 1.0 / beta_kp1, z));
    }
    double rhs2[2] = {rpn, 0};
    for (int k = 0; k < max_iterations; k++) {
        if (rpn < local_tolerance) {
            if (iters) *iters = k;
            return 0;
        }
        std::swap(mkm2, mkm1);
        std::swap(mkm1, mk);
        RC(vec_copy(s, m, z, mk));
        beta_k = beta_kp1;
        std::swap(qkm1, qkp1);
        std::swap(qkm1, qk);
        RC(obj_multiply(O, mk, qkp1));
        RC(bc_project(s, qkp1));
        RC(vec_dot(s, m, mk, qkp1, sc, &alpha_k));
        RC(vec_axpy(s, m, -alpha_k, qk, qkp1));
        RC(vec_axpy(s, m, -beta_k, qkm1, qkp1));
        RC(obj_precondition(O, qkp1, z));
        RC(vec_dot(s, m, z, qkp1, sc + 1, &d));
        beta_kp1 = std::sqrt(std::max(0.0, d));
        if (beta_kp1 > 0) {
            RC(vec_scale(s, m, 1.0 / beta_kp1, qkp1));
            RC(vec_scale(s, m, 1.0 / beta_kp1, z));
        }
        { // applyAllPreviousGivensRotationsAndDetermineNewGivens, :147-176
            Gkm2 = Gkm1;
            Gkm1 = Gk;
            double ep[2] = {0, beta_k};
            Gkm2.row_rotation(ep);
            epsilon = ep[0];
            double dz[2] = {ep[1], alpha_k};
            Gkm1.row_rotation(dz);
            delta = dz[0];
            double tmp[2] = {dz[1], beta_kp1};
            Gk.compute(tmp[0], tmp[1]);
            Gk.row_rotation(tmp);
            gamma = tmp[0];
            Gk.row_rotation(rhs2);
            tk = rhs2[0];
            const double res = rhs2[1];
            rhs2[0] = res; rhs2[1] = 0;
            rpn = res < 0 ? -res : res;
        }
        k_minres_m<<<nblk(m), TPB, 0, s->stream>>>(m, delta, epsilon, 1.0 / gamma, mkm1, mkm2, tk, mk, x);
        HOT_LAUNCHED(s);
    }
    if (iters) *iters = max_iterations;
    return 0;
}

// force_project: HinvApproxInit always assembles buildMatrix<true> (ImplicitSolver.h:337); --bcproject then only decides
// whether level 0 additionally carries objective.project (MultigridPreconditioner.h:695-699)
int rebuild_matrix_and_preconditioner(Objective& O, bool force_project = false)
{
    Sim* s = O.s;
    const hot_solver_options& o = O.opt;
    RC(build_matrix(s, force_project || o.bcproject));
    s->matrix_bcproject = o.bcproject != 0;
    RC(build_mg(s, o.mg_level, o.smoother, o.coarse_solver, o.Ainv, o.mg_times, o.mg_scale, o.topomega));
    if (O.log) O.log->matrix_builds++;
    return 0;
}

// computeStep, ImplicitSolver.h:355-432 (lsolver 2)
int compute_step(Objective& O, double* ddv, double* residual, double cg_tolerance, double rel_tol)
{
    Sim* s = O.s;
    const hot_solver_options& o = O.opt;
    RC(vec_zero(s, O.m, ddv));
    if (!o.matfree) {
        RC(rebuild_matrix_and_preconditioner(O));
        O.precond = (o.mg_level == 1 && o.mg_times == 1) ? 2 : 3; // "force diagonal entry preconditioner", :381-396
    }
    else {
        RC(build_diagonal_mf(s, o.Ainv));
        O.precond = 1;
    }
    int iters = 0;
    // -lsolver 1: MINRES (ImplicitSolver.h:406-411), 2: inexact PCG; both with the absolute tolerance backwardEulerStep gives them:
    // minres.setTolerance(maxcntol) / cg.setTolerance(maxcntol) with --usecn (MultigridSimulation.h:206-207), else the constructor's 1 (:86-88)
    if (o.lsolver == 1) RC(minres_solve(O, ddv, residual, rel_tol, cg_tolerance, o.max_cg_iterations, &iters));
    else RC(inexact_pcg(O, ddv, residual, cg_tolerance, o.max_cg_iterations, &iters));
    if (O.log) {
        O.log->total_linear_iterations += iters;
        if (O.log->n_log > 0) O.log->linear_iterations[O.log->n_log - 1] = iters;
    }
    if (o.linesearch) return line_search(O, ddv, residual, 1.0);
    return 0;
}

// ExtendedNewtonsMethod::solve; x is simulation.dv itself like in the reference
int newton_solve(Objective& O, double cg_tolerance, double tolerance)
{
    Sim* s = O.s;
    double* x = s->dv.p;
    double *step = O.vec(V_STEP), *residual = O.vec(V_RES);
    for (int it = 0; it < O.opt.max_newton_iterations; it++) {
        RC(obj_update_state(O, x));
        RC(obj_compute_residual(O, residual));
        if (O.log) O.log->iterations = it;
        bool exit = false;
        double l2 = 0;
        RC(should_exit_by_cn(O, residual, &exit, &l2));
        if (exit) {
            if (O.log) O.log->converged = 1;
            return 0;
        }
        // "gast15" suggested relative tolerance of the linear solve (ExtendedNewtonsMethod.h:58), used by MINRES
        RC(compute_step(O, step, residual, cg_tolerance, std::min(0.5, std::sqrt(std::max(l2, tolerance)))));
        RC(bc_rotate(s, step, true));
        RC(vec_axpy(s, O.m, 1.0, step, x));
        RC(bc_rotate(s, step, false));
        if (O.log) O.log->iterations = it + 1;
    }
    return 0;
}

// LBFGS::solve
int lbfgs_solve(Objective& O)
{
    Sim* s = O.s;
    const long m = O.m;
    double* x = s->dv.p;
    double* residual = O.vec(V_RES);
    RC(obj_update_state(O, x));
    RC(obj_compute_residual(O, residual));
    constexpr int H = 8, SZ = H + 1;
    struct Ring { // RingBuffer<_, 9>, LBFGS.h:23-69
        int head = 1, tail = 0, size = 0;
        void push_back() { ++tail; ++size; if (tail == SZ) tail = 0; if (size > SZ) inc_head(); }
        void pop_back() { if (size == 0) return; --tail; --size; if (tail < 0) tail = SZ - 1; }
        void inc_head() { if (size == 0) return; ++head; --size; if (head == SZ) head = 0; }
        int at(int index) const { return (index + head) % SZ; }
        int back() const { return tail; }
    } ring;
    auto dxx = [&](int k) { return O.vec(V_RING0 + k); };
    auto dg = [&](int k) { return O.vec(V_RING0 + SZ + k); };
    double dgTdx[SZ] = {0}, ksi[H] = {0};
    ring.push_back();
    for (int it = 0; it < O.opt.max_lbfgs_iterations; it++) {
        if (O.log) O.log->iterations = it;
        bool exit = false;
        RC(should_exit_by_cn(O, residual, &exit, nullptr));
        if (exit) {
            if (O.log) O.log->converged = 1;
            return 0;
        }
        if (O.opt.adaptive_h ? (it & 0xf) == 0 : it == 0) { // HinvApproxInit, ImplicitSolver.h:335-353
            RC(rebuild_matrix_and_preconditioner(O, true));
            while (ring.size > 0) ring.pop_back();
            ring.push_back();
        }
        RC(vec_copy(s, m, residual, dg(ring.back())));
        {
            KTime t(s, KC_BLAS1);
            for (int i = ring.size - 2; i >= 0; --i) {
                const int k = ring.at(i);
                double d;
                RC(vec_dot(s, m, dxx(k), residual, nullptr, &d));
                ksi[i] = d * dgTdx[k];
                RC(vec_axpy(s, m, -ksi[i], dg(k), residual));
            }
        }
        double* d = dxx(ring.back());
        RC(vcycle(s, residual, d, false));
        if (O.log) O.log->total_linear_iterations++;
        RC(bc_project(s, d));
        {
            KTime t(s, KC_BLAS1);
            for (int i = 0; i < ring.size - 1; ++i) {
                const int k = ring.at(i);
                double dot;
                RC(vec_dot(s, m, dg(k), d, nullptr, &dot));
                RC(vec_axpy(s, m, ksi[i] - dot * dgTdx[k], dxx(k), d));
            }
        }
        if (O.opt.linesearch) RC(line_search(O, d, residual, 1.0));
        RC(bc_rotate(s, d, true));
        RC(vec_axpy(s, m, 1.0, d, x));
        RC(bc_rotate(s, d, false));
        RC(obj_update_state(O, x));
        RC(obj_compute_residual(O, residual));
        double* y = dg(ring.back());
        RC(vec_axpy(s, m, -1.0, residual, y));
        double yd;
        RC(vec_dot(s, m, y, d, nullptr, &yd));
        dgTdx[ring.back()] = 1.0 / yd;
        if (dgTdx[ring.back()] <= 0.0) ring.pop_back();
        ring.push_back();
        if (O.log) O.log->iterations = it + 1;
    }
    return 0;
}

int reserve_vectors(hot_sim* s, long m, int count)
{
    for (int k = 0; k < count; ++k) HOT_CUDA(s->sv[k].reserve(m > 0 ? m : 1));
    HOT_CUDA(s->red_out.reserve(64));
    return 0;
}

} // namespace
} // namespace hot

using namespace hot;

extern "C" {

void hot_default_options(hot_solver_options* o)
{
    std::memset(o, 0, sizeof *o);
    o->lsolver = 3; o->project = 1; o->bcproject = 1; o->linesearch = 1; o->usecn = 1;
    o->mg_level = 3; o->mg_times = 1; o->smoother = 5; o->coarse_solver = 2; o->Ainv = 1;
    o->max_newton_iterations = 3; o->max_lbfgs_iterations = 10000; o->max_cg_iterations = 10000;
    o->cneps = 1e-7; o->topomega = 0.1;
}

int hot_pcg(hot_sim* s, const double* b, double* x, double tolerance, int max_iterations, int matfree, int preconditioner, int* iters)
{
    if (!s->state_valid) return fail(s, "hot_pcg: call hot_update_state first");
    if (!matfree && !s->matrix_built) return fail(s, "hot_pcg: call hot_build_matrix first");
    if (preconditioner == 2 && !s->mg_built) return fail(s, "hot_pcg: call hot_build_mg first");
    if (preconditioner == 1 && !matfree && !s->mg_built) return fail(s, "hot_pcg: the Jacobi preconditioner of the assembled matrix needs hot_build_mg (levels >= 1)");
    Objective O;
    O.s = s;
    hot_default_options(&O.opt);
    O.opt.matfree = matfree;
    O.log = nullptr;
    O.m = 3L * s->num_nodes;
    O.precond = preconditioner == 0 ? 0 : (preconditioner == 2 ? 3 : (matfree ? 1 : 2));
    int rc = reserve_vectors(s, O.m, V_RING0);
    if (rc) return rc;
    if (preconditioner == 1 && matfree) {
        rc = build_diagonal_mf(s, 1);
        if (rc) return rc;
    }
    double *xd = s->sv[V_STEP].p, *bd = s->sv[V_RES].p;
    HOT_CUDA(cudaMemcpyAsync(xd, x, O.m * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    HOT_CUDA(cudaMemcpyAsync(bd, b, O.m * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    rc = inexact_pcg(O, xd, bd, tolerance, max_iterations, iters);
    if (rc) return rc;
    HOT_CUDA(cudaMemcpyAsync(x, xd, O.m * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    HOT_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

int hot_backward_euler_step(hot_sim* s, const hot_solver_options* opt, hot_solve_log* log)
{
    if (!opt) return fail(s, "hot_backward_euler_step: null options");
    if (log) std::memset(log, 0, sizeof *log);
    if (!s->p2g_done) return fail(s, "backwardEulerStep: call hot_p2g (and hot_set_bc) first");
    if (opt->lsolver != 1 && opt->lsolver != 2 && opt->lsolver != 3) return fail(s, "lsolver must be 1 (Newton + MINRES), 2 (Newton + PCG) or 3 (L-BFGS)");
    if (opt->lsolver == 3 && opt->matfree) return fail(s, "LBFGS only works with project & with-matrix (Projects/multigrid/README:13-15)");
    if (s->world > 1 && !(opt->lsolver != 3 && opt->matfree) && !s->ghost_ring)
        return fail(s, "partitioned runs with an assembled matrix (-lsolver 3, -lsolver 2 without --matfree) need the ghost ring: hot_set_ghost_ring(h, 1) before the sort");
    Objective O;
    O.s = s;
    O.opt = *opt;
    O.log = log;
    O.m = 3L * s->num_nodes;
    int rc = reserve_vectors(s, O.m, V_COUNT);
    if (rc) return rc;
    O.dv0 = s->sv[V_DV0].p;
    O.dvnew = s->sv[V_DVNEW].p;
    s->project_pd = opt->project != 0;
    s->mg_cneps = opt->cneps; // top.tolFunc of the V-cycles of this solve
    rc = backup_strain(s); // startBackwardEuler, MultigridSimulation.h:167-186
    if (rc) return rc;
    // computeCharacteristicNorm :128-165 and the tolerances of :199-211
    double tol = opt->cneps;
    if (opt->usecn) {
        HOT_CUDA(s->cn_tol.reserve(s->num_nodes > 0 ? s->num_nodes : 1));
        rc = eval_cn_tolerance(s, opt->cneps, s->dt, s->cn_tol.p);
        if (rc) return rc;
        // dPdFNorm_max is a function-static in the reference (computed on the first step, reused afterwards even when hardening
        // changes mu / lambda): the cache lives until the next hot_set_particles
        if (s->dpdf_norm_max < 0.0) {
            unsigned long long* mx = (unsigned long long*)(s->red_out.p + 32);
            HOT_CUDA(cudaMemsetAsync(mx, 0, sizeof(*mx), s->stream));
            if (s->p1 > s->p0) {
                k_max_dpdf_norm<<<nblk(s->p1 - s->p0), TPB, 0, s->stream>>>(s->p1 - s->p0, s->P.mu.p + s->p0, s->P.lam.p + s->p0, model_flags(s), mx);
                HOT_LAUNCHED(s);
            }
            if (!s->h_red) HOT_CUDA(cudaMallocHost((void**)&s->h_red, 64 * sizeof(double)));
            HOT_CUDA(cudaMemcpyAsync(s->h_red, mx, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
            HOT_CUDA(cudaStreamSynchronize(s->stream));
            double nm = s->h_red[0];
            rc = dist_allreduce_host(s, &nm, 1, 1);
            if (rc) return rc;
            s->dpdf_norm_max = nm;
        }
        const double nmax = s->dpdf_norm_max;
        tol = opt->cneps * s->dt * 24 * std::sqrt((double)(s->world > 1 ? s->global_nodes : s->num_nodes)) * s->dx * s->dx * nmax;
    }
    if (log) log->tolerance = tol;
    const double cg_tol = opt->usecn ? tol : 1.0; // cg.setTolerance(1) in the objective ctor, maxcntol with --usecn
    // resetLSFlag, ImplicitSolver.h:277-282
    O.updated = false;
    rc = vec_copy(s, O.m, s->dv.p, O.dv0);
    if (rc) return rc;
    rc = opt->lsolver != 3 ? newton_solve(O, cg_tol, tol) : lbfgs_solve(O);
    if (rc) return rc;
    // keep ImplicitSolverObjective::dv0 readable (hot_get_dv0)
    HOT_CUDA(s->dv0_keep.reserve(O.m > 0 ? O.m : 1));
    rc = vec_copy(s, O.m, opt->linesearch ? O.dv0 : s->dv.p, s->dv0_keep.p);
    if (rc) return rc;
    s->dv0_valid = true;
    rc = restore_strain(s);
    if (rc) return rc;
    HOT_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

int hot_get_dv0(hot_sim* s, double* dv0)
{
    if (!s->dv0_valid) return fail(s, "hot_get_dv0: no solve yet");
    HOT_CUDA(cudaMemcpyAsync(dv0, s->dv0_keep.p, 3 * (size_t)s->num_nodes * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    HOT_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

} // extern "C"
