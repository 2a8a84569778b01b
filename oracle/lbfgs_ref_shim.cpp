// TEST INFRASTRUCTURE ONLY - builds oracle/_ref/libhot_oracle_lbfgsref.so: the whole CPU oracle (this translation unit includes
// hot_oracle.cpp, so the library has the complete orc_* API) plus ONE extra entry point, zr_lbfgs_backward_euler_step, which is
// orc_backward_euler_step with the L-BFGS loop replaced by the reference's OWN ZIRAN::LBFGS<Objective>::solve
// (Lib/Ziran/Math/Nonlinear/LBFGS.h:300-437, compiled where it lies) driven on the oracle's objective:
//   updateState / computeResidual / shouldExitByCN / HinvApproxInit / precondition / project / lineSearch / recoverSolution /
//   transformResidual  ->  the oracle's obj_update_state / obj_compute_residual / should_exit_by_cn / rebuild_matrix_and_preconditioner /
//   mg_vcycle / bc_project / line_search / bc_rotate.
// Same objective, the reference's loop: ring-buffer history, two-loop recursion, the dgTdx <= 0 drop rule, rebuild schedule
// (HOTSettings::useAdaptiveHessian) and the order of the objective calls are the reference's; tests/test_oracle_lbfgs_ref.py compares
// the oracle's restatement (lbfgs_solve, row a22) with it step by step.
// Stand-ins (oracle/ref_shim): Eigen / Tick / TBB / Partio / ARPACK headers, logging and timer macros (no-ops).
#include "hot_oracle.cpp"

#include <array>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <Ziran/Math/Linear/DenseExt.h>
#include <Ziran/CS/Util/Logging.h>
#include <Ziran/CS/Util/ErrorContext.h>
namespace ZIRAN {
template <class T, int dim> class GridState;                     // named by debugging helpers of LBFGS.h that are never instantiated
template <class T, int dim, int degree = 2> class BSplineWeights;
} // namespace ZIRAN
#include <tbb/tbb.h>
// (LBFGS::solve has debugging branches, HOTSettings::debugMode > 0, never taken here, that assemble the preconditioner into an Eigen sparse
// matrix: the stand-in Eigen's inert SparseMatrix / Triplet let them compile)
#include <Ziran/Math/Nonlinear/LBFGS.h>

namespace {
struct RArr {
    std::vector<double> v;
    double sum() const { double s = 0; for (double x : v) s += x; return s; }
};
inline RArr operator*(const RArr& a, const RArr& b)
{
    RArr r; r.v.resize(a.v.size());
    for (size_t i = 0; i < a.v.size(); ++i) r.v[i] = a.v[i] * b.v[i];
    return r;
}
// the NewtonVector the reference's template is instantiated with: a plain dynamic vector with the operations LBFGS::solve uses.
// It can ALIAS external storage: the reference calls lbfgs.solve(Base::dv) with the simulation's own dv, which moveNodes (and with it
// every updateState, including the probes of the line search) overwrites - that aliasing is where the reference's
// "dv = dv0 + one more copy of the last step" behaviour with --linesearch comes from, so it has to be reproduced.
struct RVec {
    std::vector<double> own;
    std::vector<double>* p = &own;
    RVec() {}
    RVec(const RVec& o) : own(*o.p), p(&own) {}
    RVec& operator=(const RVec& o) { if (this != &o) *p = *o.p; return *this; } // (assigns the VALUES, an alias stays an alias)
    void alias(std::vector<double>& ext) { p = &ext; }
    std::vector<double>& vec() { return *p; }
    const std::vector<double>& vec() const { return *p; }
    void resizeLike(const RVec& o) { p->resize(o.p->size()); }
    RArr array() const { return RArr{*p}; }
    // (used by the debugging branches only)
    int cols() const { return (int)(p->size() / 3); }
    void setZero() { for (double& x : *p) x = 0; }
    double& operator()(int d, int i) { return (*p)[3 * (size_t)i + d]; }
    double operator()(int d, int i) const { return (*p)[3 * (size_t)i + d]; }
    double norm() const { double s = 0; for (double x : *p) s += x * x; return std::sqrt(s); }
    RVec& operator+=(const RVec& o) { for (size_t i = 0; i < p->size(); ++i) (*p)[i] += (*o.p)[i]; return *this; }
    RVec& operator-=(const RVec& o) { for (size_t i = 0; i < p->size(); ++i) (*p)[i] -= (*o.p)[i]; return *this; }
};
inline RVec operator*(const RVec& a, double s) { RVec r(a); for (double& x : r.vec()) x *= s; return r; }
inline RVec operator*(double s, const RVec& a) { return a * s; }

struct OracleObjective {
    using NewtonVector = RVec;
    using Scalar = double;
    Sim* s;
    ObjectiveState& O;
    int rc = 0;          // first error of an oracle call (the reference's interface has no error channel)
    int cn_calls = 0;    // shouldExitByCN is called once per iteration: the iteration counter of the log
    void keep(int r) { if (r && !rc) rc = r; }
    void updateState(const RVec& x)
    { // moveNodes (MpmSimulationBase.cpp:736-747): dv = x unless x IS dv
        if (&x.vec() != &s->dv) s->dv = x.vec();
        keep(obj_update_state(s, O, s->dv));
    }
    void computeResidual(RVec& r) { r.vec().resize(s->dv.size()); keep(obj_compute_residual(s, O, r.vec())); }
    bool shouldExitByCN(const RVec& r)
    {
        if (O.log) O.log->iterations = cn_calls;
        ++cn_calls;
        return rc != 0 || should_exit_by_cn(s, O, r.vec()); // (an oracle error ends the loop)
    }
    void HinvApproxInit() { keep(rebuild_matrix_and_preconditioner(s, O, true)); }
    void precondition(const RVec& in, RVec& out)
    {
        out.vec().resize(in.vec().size());
        keep(mg_vcycle(s, matrix_of(s), in.vec().data(), out.vec().data()));
        if (O.log) O.log->total_linear_iterations++;
    }
    void project(RVec& v) { bc_project(s, v.vec().data()); }
    double lineSearch(RVec& d, RVec& residual, double alpha) { keep(line_search(s, O, d.vec(), residual.vec(), alpha)); return alpha; }
    void recoverSolution(RVec& d) { bc_rotate(s, d.vec().data(), true); }
    void transformResidual(RVec& d) { bc_rotate(s, d.vec().data(), false); }
    // named by debugging branches of LBFGS::solve that are compiled but not taken (HOTSettings::debugMode == 0)
    double angle(const RVec&, const RVec&) { return 1.0; }
    double innerProduct(const RVec&, const RVec&) { return 0.0; }
    void multiply(const RVec&, RVec&) {}
    void BCprojectionSanityCheck(const RVec&, const RVec&) {}
    void checkMultigridSystemMatrix() {}
};

int reference_lbfgs(Sim* s, ObjectiveState& O)
{
    HOTSettings::useAdaptiveHessian = O.opt.adaptive_h != 0;
    HOTSettings::debugMode = 0;
    OracleObjective obj{s, O};
    ZIRAN::LBFGS<OracleObjective> lbfgs(obj, O.opt.cneps, O.opt.max_lbfgs_iterations);
    RVec x;
    x.alias(s->dv); // lbfgs.solve(Base::dv, ...): MultigridSimulation.h:221
    const bool converged = lbfgs.solve(x, false, O.opt.linesearch != 0);
    if (O.log) {
        O.log->converged = converged && !obj.rc ? 1 : 0;
        if (!converged) O.log->iterations = O.opt.max_lbfgs_iterations;
    }
    return obj.rc;
}
} // namespace

extern "C" int zr_lbfgs_backward_euler_step(void* h, const hot_solver_options* opt, hot_solve_log* log)
{
    if (opt->lsolver != 3) return -1;
    return backward_euler_step_with(h, opt, log, reference_lbfgs);
}
