// TEST INFRASTRUCTURE ONLY — included by hot_oracle.cpp.  Restates the outer solvers of the hot path: a21 inexact PCG,
// a22 L-BFGS around the V-cycle, the extended Newton loop, the objective's line search / CN exit test / step
// computation (a14) and the backward-Euler glue of a24.  The option / log structs are the public ones of
// include/hot_b200.h (plain C data, shared so that one test harness drives both implementations).
#include "../include/hot_b200.h"

namespace {

// ImplicitSolverObjective state that lives across calls inside one backward-Euler step (ImplicitSolver.h:58-71)
struct ObjectiveState {
    Vd dv0, nodeCNTol;
    double Ek = 0;
    bool updated = false;
    hot_solver_options opt;
    hot_solve_log* log = nullptr;
    int precond = 0; // 0 identity, 1 diagVal (matrix-free block Jacobi), 2 level-0 diagonal, 3 V-cycle
};

// computeNorm, ImplicitSolver.h:158-171
double l2norm(const Vd& r) { return std::sqrt(vec_dot(r, r)); }

void log_iter(ObjectiveState& O, Sim* s, const Vd& residual, double scaled)
{
    hot_solve_log* L = O.log;
    if (!L || L->n_log >= HOT_LOG_CAP) return;
    L->residual_norm[L->n_log] = l2norm(residual);
    L->scaled_norm[L->n_log] = scaled;
    L->energy[L->n_log] = O.Ek;
    L->linear_iterations[L->n_log] = 0;
    L->n_log++;
}

// shouldExitByCN, ImplicitSolver.h:174-211
bool should_exit_by_cn(Sim* s, ObjectiveState& O, const Vd& residual)
{
    const int nn = s->num_nodes;
    if (!O.opt.usecn) {
        double res = l2norm(residual);
        log_iter(O, s, residual, res);
        return res < O.opt.cneps;
    }
    double scaled = 0;
    for (int i = 0; i < nn; ++i) {
        const double* r = &residual[3 * (size_t)i];
        scaled += (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]) / (O.nodeCNTol[i] * O.nodeCNTol[i]);
    }
    log_iter(O, s, residual, nn ? std::sqrt(scaled / nn) : 0.0);
    if (nn == 0) return true;
    return scaled < nn;
}

// updateState / computeResidual with the `updated` short-circuit (ImplicitSolver.h:128-155,237-252)
int obj_update_state(Sim* s, ObjectiveState& O, const Vd& dv, bool force = false)
{
    if (O.updated && !force) return 0;
    double e = 0;
    int rc = orc_update_state(s, dv.data(), O.opt.linesearch ? &e : nullptr);
    if (O.opt.linesearch) O.Ek = e;
    return rc;
}
int obj_compute_residual(Sim* s, ObjectiveState& O, Vd& r, bool force = false)
{
    if (O.updated && !force) return 0;
    return orc_compute_residual(s, r.data());
}

// lineSearch, ImplicitSolver.h:312-333 (the halving loop is capped at 60 probes so a NaN energy cannot hang the test)
int line_search(Sim* s, ObjectiveState& O, Vd& ddv, Vd& residual, double alpha)
{
    Vd dvnew(ddv.size());
    bc_rotate(s, ddv.data(), true); // recoverSolution
    const double Ek0 = O.Ek;
    int probes = 0;
    do {
        for (size_t q = 0; q < ddv.size(); ++q) dvnew[q] = O.dv0[q] + ddv[q] * alpha;
        int rc = obj_update_state(s, O, dvnew, true);
        if (rc) return rc;
        alpha *= 0.5;
        if (O.log) O.log->total_linesearch_probes++;
    } while (O.Ek > Ek0 && ++probes < 60);
    alpha *= 2;
    for (auto& v : ddv) v *= alpha;
    bc_rotate(s, ddv.data(), false); // transformResidual
    int rc = obj_compute_residual(s, O, residual, true);
    O.updated = true;
    O.dv0 = dvnew;
    return rc;
}

// objective.multiply / precondition / project as the Krylov solver sees them (ImplicitSolver.h:741-763)
int obj_multiply(Sim* s, ObjectiveState& O, const Vd& x, Vd& b)
{
    if (O.opt.matfree) return orc_hessian_apply_mf(s, x.data(), b.data());
    return orc_spmv(s, 0, x.data(), b.data());
}
int obj_precondition(Sim* s, ObjectiveState& O, const Vd& in, Vd& out)
{
    MatrixState& M = matrix_of(s);
    const int nn = s->num_nodes;
    switch (O.precond) {
    case 0: out = in; return 0;
    case 1:
    case 2: {
        const Vd& D = O.precond == 1 ? M.diagVal : (M.Ainv == 0 ? M.sysmats[0].diagonalEntry : M.sysmats[0].diagonalBlock);
        for (int i = 0; i < nn; ++i) {
            double y[3] = {0, 0, 0};
            m3_mulv_add(&D[9 * (size_t)i], &in[3 * (size_t)i], y);
            out[3 * (size_t)i] = y[0]; out[3 * (size_t)i + 1] = y[1]; out[3 * (size_t)i + 2] = y[2];
        }
        return 0;
    }
    default: return mg_vcycle(s, M, in.data(), out.data());
    }
}

// InexactConjugateGradient::solve, Lib/Ziran/Math/Linear/InexactConjugateGradient.h:49-103
int inexact_pcg(Sim* s, ObjectiveState& O, Vd& x, const Vd& b, double tolerance, int max_iterations, int* iters)
{
    const size_t m = b.size();
    Vd r(m), p(m), q(m), temp(m);
    int rc = obj_multiply(s, O, x, temp);
    if (rc) return rc;
    for (size_t i = 0; i < m; ++i) r[i] = b[i] - temp[i];
    bc_project(s, r.data());
    rc = obj_precondition(s, O, r, q);
    if (rc) return rc;
    p = q;
    double zTrk = vec_dot(r, q);
    double rpn = std::sqrt(zTrk);
    const double forcing = std::min(0.5, std::sqrt(std::max(rpn, tolerance)));
    const double local_tolerance = forcing * rpn;
    int cnt = 0;
    for (cnt = 0; cnt < max_iterations; ++cnt) {
        if (rpn < local_tolerance) break;
        rc = obj_multiply(s, O, p, temp);
        if (rc) return rc;
        bc_project(s, temp.data());
        const double alpha = zTrk / vec_dot(temp, p);
        for (size_t i = 0; i < m; ++i) {
            x[i] += p[i] * alpha;
            r[i] -= temp[i] * alpha;
        }
        rc = obj_precondition(s, O, r, q);
        if (rc) return rc;
        const double zTrk_last = zTrk;
        zTrk = vec_dot(q, r);
        const double beta = zTrk / zTrk_last;
        for (size_t i = 0; i < m; ++i) p[i] = q[i] + beta * p[i];
        rpn = std::sqrt(zTrk);
    }
    if (iters) *iters = cnt;
    return 0;
}


// Minres::solve, Lib/Ziran/Math/Linear/Minres.h:71-176 (Givens rotations: Givens.h:73-141); tolerance = min(relative * |r0|_Minv, absolute)
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
int minres_solve(Sim* s, ObjectiveState& O, Vd& x, const Vd& b, double relative_tolerance, double tolerance, int max_iterations, int* iters)
{
    const size_t m = b.size();
    Vd mk(m, 0.0), mkm1(m, 0.0), mkm2(m, 0.0), z(m, 0.0), qkp1(m, 0.0), qk(m, 0.0), qkm1(m, 0.0);
    GivensRot Gk, Gkm1, Gkm2;
    double gamma = 0, delta = 0, epsilon = 0, beta_kp1 = 0, alpha_k = 0, beta_k = 0, tk = 0;
    int rc = obj_multiply(s, O, x, qkp1);
    if (rc) return rc;
    for (size_t i = 0; i < m; ++i) qkp1[i] = b[i] - qkp1[i];
    bc_project(s, qkp1.data());
    rc = obj_precondition(s, O, qkp1, z);
    if (rc) return rc;
    double rpn = std::sqrt(vec_dot(z, qkp1));
    beta_kp1 = rpn;
    const double local_tolerance = std::min(relative_tolerance * rpn, tolerance);
    if (iters) *iters = 0;
    if (rpn < local_tolerance) return 0;
    if (rpn > 0)
        for (size_t i = 0; i < m; ++i) { qkp1[i] /= beta_kp1; z[i] /= beta_kp1; }
    double rhs2[2] = {rpn, 0};
    for (int k = 0; k < max_iterations; k++) {
        if (rpn < local_tolerance) {
            if (iters) *iters = k;
            return 0;
        }
        mkm2.swap(mkm1);
        mkm1.swap(mk);
        mk = z;
        beta_k = beta_kp1;
        qkm1.swap(qkp1);
        qkm1.swap(qk);
        rc = obj_multiply(s, O, mk, qkp1);
        if (rc) return rc;
        bc_project(s, qkp1.data());
        alpha_k = vec_dot(mk, qkp1);
        for (size_t i = 0; i < m; ++i) qkp1[i] -= alpha_k * qk[i];
        for (size_t i = 0; i < m; ++i) qkp1[i] -= beta_k * qkm1[i];
        rc = obj_precondition(s, O, qkp1, z);
        if (rc) return rc;
        beta_kp1 = std::sqrt(std::max(0.0, vec_dot(z, qkp1)));
        if (beta_kp1 > 0)
            for (size_t i = 0; i < m; ++i) { qkp1[i] /= beta_kp1; z[i] /= beta_kp1; }
        { // applyAllPreviousGivensRotationsAndDetermineNewGivens :147-176
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
        for (size_t i = 0; i < m; ++i) mk[i] = (mk[i] - delta * mkm1[i] - epsilon * mkm2[i]) / gamma;
        for (size_t i = 0; i < m; ++i) x[i] += tk * mk[i];
    }
    if (iters) *iters = max_iterations;
    return 0;
}

// force_project: HinvApproxInit always assembles buildMatrix<true> (ImplicitSolver.h:337); the --bcproject flag then only
// decides whether level 0 additionally carries objective.project (MultigridPreconditioner.h:695-699)
int rebuild_matrix_and_preconditioner(Sim* s, ObjectiveState& O, bool force_project = false)
{
    const hot_solver_options& o = O.opt;
    int rc = orc_build_matrix(s, force_project ? 1 : o.bcproject);
    if (rc) return rc;
    matrix_of(s).bcproject = o.bcproject != 0;
    rc = orc_build_mg(s, o.mg_level, o.smoother, o.coarse_solver, o.Ainv, o.mg_times, o.mg_scale, o.topomega);
    if (O.log) O.log->matrix_builds++;
    return rc;
}

// computeStep, ImplicitSolver.h:355-432 (lsolver 2)
int compute_step(Sim* s, ObjectiveState& O, Vd& ddv, Vd& residual, double rel_tol /* used by MINRES; the inexact CG derives its own */, double cg_tolerance)
{
    std::fill(ddv.begin(), ddv.end(), 0.0);
    const hot_solver_options& o = O.opt;
    if (!o.matfree) {
        int rc = rebuild_matrix_and_preconditioner(s, O);
        if (rc) return rc;
        O.precond = (o.mg_level == 1 && o.mg_times == 1) ? 2 : 3; // "force diagonal entry preconditioner" :381-396
    }
    else {
        int rc = orc_build_diagonal(s, o.Ainv, nullptr);
        if (rc) return rc;
        O.precond = 1;
    }
    int iters = 0;
    // -lsolver 1: MINRES (ImplicitSolver.h:406-411), -lsolver 2: inexact PCG; both with the absolute tolerance backwardEulerStep gives them:
    // minres.setTolerance(maxcntol) / cg.setTolerance(maxcntol) with --usecn (MultigridSimulation.h:206-207), else the constructor's 1 (:86-88)
    int rc = o.lsolver == 1 ? minres_solve(s, O, ddv, residual, rel_tol, cg_tolerance, o.max_cg_iterations, &iters)
                            : inexact_pcg(s, O, ddv, residual, cg_tolerance, o.max_cg_iterations, &iters);
    if (rc) return rc;
    if (O.log) {
        O.log->total_linear_iterations += iters;
        if (O.log->n_log > 0) O.log->linear_iterations[O.log->n_log - 1] = iters;
    }
    if (o.linesearch) return line_search(s, O, ddv, residual, 1.0);
    return 0;
}

// ExtendedNewtonsMethod::solve, Lib/Ziran/Math/Nonlinear/ExtendedNewtonsMethod.h:39-66
int newton_solve(Sim* s, ObjectiveState& O, double tolerance, double cg_tolerance)
{
    Vd& x = s->dv; // the reference passes Base::dv itself: x aliases the simulation's dv
    Vd step(x.size()), residual(x.size());
    for (int it = 0; it < O.opt.max_newton_iterations; it++) {
        int rc = obj_update_state(s, O, x);
        if (rc) return rc;
        rc = obj_compute_residual(s, O, residual);
        if (rc) return rc;
        const double residual_norm = l2norm(residual);
        if (O.log) O.log->iterations = it;
        if (should_exit_by_cn(s, O, residual)) {
            if (O.log) O.log->converged = 1;
            return 0;
        }
        const double rel = std::min(0.5, std::sqrt(std::max(residual_norm, tolerance)));
        rc = compute_step(s, O, step, residual, rel, cg_tolerance);
        if (rc) return rc;
        bc_rotate(s, step.data(), true);
        for (size_t q = 0; q < x.size(); ++q) x[q] += step[q];
        bc_rotate(s, step.data(), false);
        if (O.log) O.log->iterations = it + 1;
    }
    return 0;
}

// LBFGS::solve, Lib/Ziran/Math/Nonlinear/LBFGS.h:300-437
int lbfgs_solve(Sim* s, ObjectiveState& O)
{
    Vd& x = s->dv;
    const size_t m = x.size();
    Vd residual(m);
    int rc = obj_update_state(s, O, x);
    if (rc) return rc;
    rc = obj_compute_residual(s, O, residual);
    if (rc) return rc;
    constexpr int H = 8, SZ = H + 1;
    // RingBuffer<_, 9> (:23-69)
    struct Ring {
        int head = 1, tail = 0, size = 0;
        void push_back() { ++tail; ++size; if (tail == SZ) tail = 0; if (size > SZ) inc_head(); }
        void pop_back() { if (size == 0) return; --tail; --size; if (tail < 0) tail = SZ - 1; }
        void inc_head() { if (size == 0) return; ++head; --size; if (head == SZ) head = 0; }
        int at(int index) const { return (index + head) % SZ; }
        int back() const { return tail; }
    } ring;
    std::vector<Vd> dxx(SZ, Vd(m)), dg(SZ, Vd(m));
    double dgTdx[SZ] = {0};
    double ksi[H] = {0};
    ring.push_back();
    MatrixState& M = matrix_of(s);
    for (int it = 0; it < O.opt.max_lbfgs_iterations; it++) {
        if (O.log) O.log->iterations = it;
        if (should_exit_by_cn(s, O, residual)) {
            if (O.log) O.log->converged = 1;
            return 0;
        }
        if (O.opt.adaptive_h ? (it & 0xf) == 0 : it == 0) { // HinvApproxInit :335-353
            rc = rebuild_matrix_and_preconditioner(s, O, true);
            if (rc) return rc;
            while (ring.size > 0) ring.pop_back();
            ring.push_back();
        }
        dg[ring.back()] = residual;
        for (int i = ring.size - 2; i >= 0; --i) {
            const int k = ring.at(i);
            ksi[i] = vec_dot(dxx[k], residual) * dgTdx[k];
            for (size_t q = 0; q < m; ++q) residual[q] -= ksi[i] * dg[k][q];
        }
        Vd& d = dxx[ring.back()];
        rc = mg_vcycle(s, M, residual.data(), d.data());
        if (rc) return rc;
        if (O.log) O.log->total_linear_iterations++;
        bc_project(s, d.data());
        for (int i = 0; i < ring.size - 1; ++i) {
            const int k = ring.at(i);
            const double c = ksi[i] - vec_dot(dg[k], d) * dgTdx[k];
            for (size_t q = 0; q < m; ++q) d[q] += dxx[k][q] * c;
        }
        if (O.opt.linesearch) {
            rc = line_search(s, O, d, residual, 1.0);
            if (rc) return rc;
        }
        bc_rotate(s, d.data(), true);
        for (size_t q = 0; q < m; ++q) x[q] += d[q];
        bc_rotate(s, d.data(), false);
        rc = obj_update_state(s, O, x);
        if (rc) return rc;
        rc = obj_compute_residual(s, O, residual);
        if (rc) return rc;
        Vd& y = dg[ring.back()];
        for (size_t q = 0; q < m; ++q) y[q] -= residual[q];
        dgTdx[ring.back()] = 1.0 / vec_dot(y, d);
        if (dgTdx[ring.back()] <= 0.0) ring.pop_back();
        ring.push_back();
        if (O.log) O.log->iterations = it + 1;
    }
    return 0;
}

} // namespace

extern "C" {

void orc_default_options(hot_solver_options* o)
{
    std::memset(o, 0, sizeof *o);
    o->lsolver = 3; o->project = 1; o->bcproject = 1; o->linesearch = 1; o->usecn = 1;
    o->mg_level = 3; o->mg_times = 1; o->smoother = 5; o->coarse_solver = 2; o->Ainv = 1;
    o->max_newton_iterations = 3; o->max_lbfgs_iterations = 10000; o->max_cg_iterations = 10000;
    o->cneps = 1e-7; o->topomega = 0.1;
}

int orc_pcg(void* h, const double* b, double* x, double tolerance, int max_iterations, int matfree, int preconditioner, int* iters)
{
    Sim* s = (Sim*)h;
    ObjectiveState O;
    orc_default_options(&O.opt);
    O.opt.matfree = matfree;
    O.precond = preconditioner == 0 ? 0 : (preconditioner == 2 ? 3 : (matfree ? 1 : 2));
    const size_t m = 3 * (size_t)s->num_nodes;
    Vd xv(x, x + m), bv(b, b + m);
    int rc = 0;
    if (O.precond == 1) rc = orc_build_diagonal(s, 1, nullptr);
    if (O.precond == 2 && matrix_of(s).sysmats.empty()) return fail(s, "orc_pcg: the Jacobi preconditioner of the assembled matrix needs orc_build_mg");
    if (rc) return rc;
    rc = inexact_pcg(s, O, xv, bv, tolerance, max_iterations, iters);
    std::copy(xv.begin(), xv.end(), x);
    return rc;
}

// The operator callbacks of orc_pcg / orc_minres on their own (tests drive the REFERENCE's Krylov solver classes with them:
// oracle/ziran_krylov_shim.cpp): A.multiply, A.precondition with the same selection as orc_pcg; A.project is orc_project.
static void krylov_objective(ObjectiveState& O, int matfree, int preconditioner)
{
    orc_default_options(&O.opt);
    O.opt.matfree = matfree;
    O.precond = preconditioner == 0 ? 0 : (preconditioner == 2 ? 3 : (matfree ? 1 : 2));
}
int orc_obj_multiply(void* h, int matfree, const double* x, double* b)
{
    Sim* s = (Sim*)h;
    ObjectiveState O;
    krylov_objective(O, matfree, 0);
    const size_t m = 3 * (size_t)s->num_nodes;
    Vd xv(x, x + m), bv(m);
    int rc = obj_multiply(s, O, xv, bv);
    std::copy(bv.begin(), bv.end(), b);
    return rc;
}
int orc_obj_precondition(void* h, int matfree, int preconditioner, const double* in, double* out)
{
    Sim* s = (Sim*)h;
    ObjectiveState O;
    krylov_objective(O, matfree, preconditioner);
    const size_t m = 3 * (size_t)s->num_nodes;
    if (O.precond == 1 && matrix_of(s).diagVal.size() != 3 * m) { // (orc_pcg builds the matrix-free block diagonal before it iterates)
        int rc0 = orc_build_diagonal(s, 1, nullptr);
        if (rc0) return rc0;
    }
    if (O.precond == 2 && matrix_of(s).sysmats.empty()) return fail(s, "orc_obj_precondition: the Jacobi preconditioner of the assembled matrix needs orc_build_mg");
    Vd iv(in, in + m), ov(m);
    int rc = obj_precondition(s, O, iv, ov);
    std::copy(ov.begin(), ov.end(), out);
    return rc;
}

// Minres::solve on the current system as a stand-alone call (same operator / preconditioner selection as orc_pcg)
int orc_minres(void* h, const double* b, double* x, double relative_tolerance, double tolerance, int max_iterations, int matfree, int preconditioner,
    int* iters)
{
    Sim* s = (Sim*)h;
    ObjectiveState O;
    orc_default_options(&O.opt);
    O.opt.matfree = matfree;
    O.precond = preconditioner == 0 ? 0 : (preconditioner == 2 ? 3 : (matfree ? 1 : 2));
    const size_t m = 3 * (size_t)s->num_nodes;
    Vd xv(x, x + m), bv(b, b + m);
    int rc = 0;
    if (O.precond == 1) rc = orc_build_diagonal(s, 1, nullptr);
    if (O.precond == 2 && matrix_of(s).sysmats.empty()) return fail(s, "orc_minres: the Jacobi preconditioner of the assembled matrix needs orc_build_mg");
    if (rc) return rc;
    rc = minres_solve(s, O, xv, bv, relative_tolerance, tolerance, max_iterations, iters);
    std::copy(xv.begin(), xv.end(), x);
    return rc;
}

// MultigridSimulation::backwardEulerStep (MultigridSimulation.h:188-233) after the caller has set the BC table
// `lbfgs`: the L-BFGS loop to run for -lsolver 3 - lbfgs_solve below, or the reference's own ZIRAN::LBFGS::solve driven on this
// objective (oracle/lbfgs_ref_shim.cpp, test infrastructure that pins lbfgs_solve to the reference's code)
static int backward_euler_step_with(void* h, const hot_solver_options* opt, hot_solve_log* log, int (*lbfgs)(Sim*, ObjectiveState&))
{
    Sim* s = (Sim*)h;
    ForceState& f = force_of(s);
    ObjectiveState O;
    O.opt = *opt;
    O.log = log;
    if (log) std::memset(log, 0, sizeof *log);
    if (opt->lsolver != 1 && opt->lsolver != 2 && opt->lsolver != 3) return fail(s, "lsolver must be 1 (Newton + MINRES), 2 (Newton + PCG) or 3 (L-BFGS)");
    if (opt->lsolver == 3 && opt->matfree) return fail(s, "LBFGS only works with project & with-matrix (Projects/multigrid/README:13-15)");
    f.project = opt->project != 0;
    matrix_of(s).cneps = opt->cneps;
    orc_backup_strain(s); // startBackwardEuler :167-186
    // computeCharacteristicNorm :128-165
    double tol = opt->cneps;
    if (opt->usecn) {
        O.nodeCNTol.assign(s->num_nodes, 0.0);
        int rc = orc_eval_cn_tolerance(s, opt->cneps, s->dt, O.nodeCNTol.data());
        if (rc) return rc;
        // `static bool first` / `static double dPdFNorm_max` (:131-133): evaluated on the first step only, reused afterwards
        if (s->dpdf_norm_max < 0) {
            double nmax = -1;
            const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
            for (long i = 0; i < s->N; ++i) {
                Scratch sc;
                double Hd[81], nrm = 0;
                update_scratch(I, s->mu[i], s->lambda[i], f.project, sc);
                first_piola_derivative(sc, Hd);
                for (int q = 0; q < 81; ++q) nrm += Hd[q] * Hd[q];
                nmax = std::max(nmax, std::sqrt(nrm));
            }
            s->dpdf_norm_max = nmax;
        }
        const double nmax = s->dpdf_norm_max;
        tol = opt->cneps * s->dt * 24 * std::sqrt((double)s->num_nodes) * s->dx * s->dx * nmax;
    }
    if (log) log->tolerance = tol;
    const double cg_tol = opt->usecn ? tol : 1.0; // cg.setTolerance(1) in the objective ctor, maxcntol with --usecn
    // resetLSFlag :277-282
    O.updated = false;
    O.dv0 = s->dv;
    int rc = opt->lsolver != 3 ? newton_solve(s, O, tol, cg_tol) : lbfgs(s, O);
    if (rc) return rc;
    matrix_of(s).dv0 = opt->linesearch ? O.dv0 : s->dv;
    orc_restore_strain(s);
    return 0;
}
int orc_backward_euler_step(void* h, const hot_solver_options* opt, hot_solve_log* log) { return backward_euler_step_with(h, opt, log, lbfgs_solve); }

// ImplicitSolverObjective::dv0 (ImplicitSolver.h:58): the last iterate the line search accepted.  With --linesearch the
// reference leaves dv = dv0 + (one more copy of the last step) behind; without it dv0 == dv.
int orc_get_dv0(void* h, double* dv0)
{
    Sim* s = (Sim*)h;
    MatrixState& M = matrix_of(s);
    if (M.dv0.size() != 3 * (size_t)s->num_nodes) return fail(s, "orc_get_dv0: no solve yet");
    std::copy(M.dv0.begin(), M.dv0.end(), dv0);
    return 0;
}

} // extern "C"
