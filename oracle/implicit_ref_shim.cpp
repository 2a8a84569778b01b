// TEST INFRASTRUCTURE ONLY — not part of the product path.
//
// The REFERENCE'S OWN implicit-solver objective — ZIRAN::ImplicitSolverObjective<Simulation> (Projects/multigrid/ImplicitSolver.h: a class
// template over the simulation type) — compiled where it lies and instantiated on a small stand-in simulation, so that its member functions
//   buildMatrix<projectSystem>      ImplicitSolver.h:470-603   (row a15: 125-slot block rows, upper-triangle scatter + mirror, BC projection)
//   buildDiagonal                   ImplicitSolver.h:605-665   (matrix-free block-Jacobi preconditioner)
//   evaluatePerNodeCNTolerance      ImplicitSolver.h:667-697   (row a14: per-node tolerance of the CN exit test)
//   computeResidual                 ImplicitSolver.h:125-155   (row a14: gravity + forces + inertia, transformResidual, project)
//   updateState / totalEnergy       ImplicitSolver.h:237-275   (row a14: line-search energy)
//   multiply                        ImplicitSolver.h:741-763   (rows a13 / a16: matrix-free and assembled operator)
//   shouldExitByCN                  ImplicitSolver.h:171-215
// run unmodified.  Under them sit, also the reference's own code: the grid (Lib/MPM/MpmGrid.h over Lib/SPGrid/Core, see mpmgrid_ref_shim.cpp), the
// constitutive model ZIRAN::CorotatedIsotropic (Lib/Ziran/Physics/ConstitutiveModel/CorotatedIsotropic.h: updateScratch, firstPiola, psi,
// firstPiolaDerivative, firstPiolaDifferential) and the assembled operator SparseMatrix / MultigridOperator (Projects/multigrid/SparseMatrixFast.h).
// What the stand-in simulation supplies in place of MpmSimulationBase / MpmForceBase / FBasedMpmForceHelper / MassLumpedInertia (they need Scene /
// DataManager / Particles / TBB containers and cannot be compiled here) are their particle loops around those calls, written out below statement
// for statement with the lines they follow:
//   FBasedMpmForceHelper::runLambdaWithDifferential / computePerNodeCNTolerance   Lib/MPM/Force/FBasedMpmForceHelper.h:63-121,123-155
//   FBasedMpmForceHelper::{evolveStrain, updateImplicitState, totalEnergy, computeStressDifferential}   FBasedMpmForceHelper.cpp:66-161
//   MpmForceBase::{evalInterpolantAndGradient, rasterizeForceToTVStack<false>, addScaledForceDifferential, updatePositionBasedState}
//                                                                                  Lib/MPM/Force/MpmForceBase.cpp:100-153,212-306,310-327
//   MassLumpedInertia::{totalEnergy, addScaledForces, addScaledForceDifferential}  Lib/Ziran/Physics/LagrangianForce/Inertia.cpp:14-56
// Built by oracle/Makefile into oracle/_ref/libimplicit_ref.so; tests/golden/make_implicit_golden.py writes tests/golden/implicit_ref.npz from it and
// tests/test_oracle_implicit_ref.py compares the oracle's restatement (oracle_force.inl, oracle_matrix.inl) and the CUDA path with it.
#include "mpmgrid_ref_shim.cpp"

#include <functional>
#include <map>
#include <string>
#include <unordered_map>
#include <Ziran/Math/Linear/DenseExt.h>
#include <Ziran/CS/Util/Logging.h>
#include <Ziran/CS/Util/ErrorContext.h>
#include <Ziran/CS/Util/Timer.h>
#include <Ziran/Physics/ConstitutiveModel/HyperelasticConstitutiveModel.h>
#include <Ziran/Physics/ConstitutiveModel/CorotatedIsotropic.h>

// names ImplicitSolver.h uses from headers that cannot be compiled here: AttributeName / DataManager (ref_shim/Ziran/CS/DataStructure/DataManager.h, shared
// with PlasticityApplier.cpp below) and CollisionNode (CollisionObject.h: same members, nothing else)
#include <Ziran/Physics/PlasticityApplier.cpp> // the reference's return mappings, compiled where they lie (see plasticity_ref_shim.cpp)
namespace ZIRAN {
template <class T_, int dim_>
struct CollisionNode { // Lib/Ziran/Math/Geometry/CollisionObject.h:16-25
    int node_id;
    Matrix<T_, dim_, dim_> P;
    Matrix<T_, dim_, dim_> R, Rinv;
    bool shouldRotate;
};
} // namespace ZIRAN

#include <ImplicitSolver.h>
#include <Ziran/Math/Nonlinear/ExtendedNewtonsMethod.h>
#include <Ziran/Math/Nonlinear/LBFGS.h>

// SquareMatrix::comp (SquareMatrix.h:39-46) falls off its end without a return value for equal colour keys (undefined behaviour that GCC compiles
// into a crash, see mg_ref_shim.cpp): defined as 0 here as there, the rest of the function is the reference's loop
namespace ZIRAN {
template <>
inline int SquareMatrix<double, 3>::comp(const std::array<int, 3>& a, const std::array<int, 3>& b)
{
    for (int i = 0; i < 3; ++i)
        if (a[i] < b[i]) return -1;
        else if (a[i] > b[i]) return 1;
    return 0;
}
} // namespace ZIRAN

namespace {
typedef Eigen::Matrix<T, 9, 9> Hessian9;
typedef Matrix<T, dim, Eigen::Dynamic> TVStack;
typedef Vector<T, Eigen::Dynamic> Vec;
struct MockSim;

struct MockParticles : public DataManager {
    std::vector<TV>* Xp = nullptr;
    int count = 0;
    struct XView {
        MockParticles* p;
        TV& operator[](int i) const { return (*p->Xp)[i]; }
    } X{this};
};

struct MockHelper {
    MockSim& sim;
    explicit MockHelper(MockSim& s) : sim(s) {}
    void runLambdaWithDifferential(const std::vector<int>& particle_order, const std::vector<std::pair<int, int>>& particle_group,
        const std::vector<uint64_t>& block_offset, std::function<void(int, const Hessian9&, const TM&, const T&, const T&, bool)> func, int opt = 0);
    void computePerNodeCNTolerance(const std::vector<int>& particle_order, const std::vector<std::pair<int, int>>& particle_group,
        const std::vector<uint64_t>& block_offset, std::function<void(int, const Hessian9&, const T&, bool)> func);
};

struct MockForce {
    MockSim& sim;
    std::vector<std::unique_ptr<MockHelper>> helpers;
    explicit MockForce(MockSim& s) : sim(s) { helpers.emplace_back(new MockHelper(s)); }
    void restoreStrain();
    void evolveStrainWithDt() {}
    template <class Func>
    void evalInterpolantAndGradient(Func&& f, TVStack& f_eval, std::vector<TM>& grad_f);
    void rasterizeForceToTVStack(const T scale, TVStack& force);
    void updatePositionBasedState();
    T totalEnergy();
    void addScaledForces(const T scale, TVStack& forces) { rasterizeForceToTVStack(scale, forces); }
    void addScaledForceDifferential(const T scale, const TVStack& dv, TVStack& df);
};

struct MockInertia { // MassLumpedInertia, Inertia.cpp:14-56
    MockSim& sim;
    explicit MockInertia(MockSim& s) : sim(s) {}
    void updatePositionBasedState() {}
    T totalEnergy() const;
    void addScaledForces(T scale, TVStack& forces) const;
    void addScaledForceDifferential(T scale, const TVStack& dx, TVStack& df) const;
};

struct MockSim : public RefSim {
    typedef double Scalar;
    static const int dim = 3;
    T dt = 0;
    TV gravity;
    Vec mass_matrix;
    TVStack dv, vn;
    bool quasistatic = false, full_implicit = false, verbose = false, project_pd = true;
    std::vector<CollisionNode<T, 3>> collision_nodes;
    MockParticles particles;
    std::vector<T> vol, mu, lam, Jp;
    bool cn_first = true;
    double cn_dPdFNorm = -1, cn_dPdFNorm_max = -1;
    std::vector<TM> F, Fn;
    TVStack scratch_vp, scratch_fp;
    std::vector<TM> scratch_stress;
    std::unique_ptr<MockForce> force;
    std::vector<MockForce*> forces;
    std::unique_ptr<MockInertia> inertia;
    std::unique_ptr<ImplicitSolverObjective<MockSim>> objective;

    int getSubstep() const { return 0; }
    int getFrame() const { return 0; }
    void moveNodes(const TVStack& dv_in) // MpmSimulationBase.cpp:736-747
    {
        if (dv.data() == dv_in.data())
            return;
        for (int i = 0; i < num_nodes; ++i)
            dv.col(i) = dv_in.col(i);
    }
    void addScaledForces(const T scale, TVStack& f) // SimulationBase: over `forces`
    {
        for (auto& lf : forces) lf->addScaledForces(scale, f);
    }
    void addScaledForceDifferentials(const T scale, const TVStack& x, TVStack& f)
    {
        for (auto& lf : forces) lf->addScaledForceDifferential(scale, x, f);
    }
    template <class M>
    void particlesToMultigrids(M&) {}
    CorotatedIsotropic<T, 3> model(int i) const
    {
        CorotatedIsotropic<T, 3> m;
        m.mu = mu[i];
        m.lambda = lam[i];
        m.project = project_pd;
        return m;
    }
};

// FBasedMpmForceHelper.h:63-121 with opt == 0
void MockHelper::runLambdaWithDifferential(const std::vector<int>& particle_order, const std::vector<std::pair<int, int>>& particle_group,
    const std::vector<uint64_t>& block_offset, std::function<void(int, const Hessian9&, const TM&, const T&, const T&, bool)> func, int opt)
{
    for (uint64_t color = 0; color < (1 << dim); ++color)
        tbb::parallel_for(0, (int)particle_group.size(), [&](int group_idx) {
            if ((block_offset[group_idx] & ((1 << dim) - 1)) != color)
                return;
            for (int idx = particle_group[group_idx].first; idx <= particle_group[group_idx].second; ++idx) {
                int i = particle_order[idx];
                auto& F = sim.F[i];
                auto model = sim.model(i);
                auto& Fn_local = sim.Fn[i];
                CorotatedIsotropicScratch<T, 3> s;
                Hessian9 firstPiolaDerivative;
                model.updateScratch(F, s);
                model.firstPiolaDerivative(s, firstPiolaDerivative);
                func(i, firstPiolaDerivative, Fn_local, (T)-1, (T)-1, false);
            }
        });
}

// FBasedMpmForceHelper.h:123-155
void MockHelper::computePerNodeCNTolerance(const std::vector<int>& particle_order, const std::vector<std::pair<int, int>>& particle_group,
    const std::vector<uint64_t>& block_offset, std::function<void(int, const Hessian9&, const T&, bool)> func)
{
    for (uint64_t color = 0; color < (1 << dim); ++color)
        tbb::parallel_for(0, (int)particle_group.size(), [&](int group_idx) {
            if ((block_offset[group_idx] & ((1 << dim) - 1)) != color)
                return;
            for (int idx = particle_group[group_idx].first; idx <= particle_group[group_idx].second; ++idx) {
                int i = particle_order[idx];
                auto model = sim.model(i);
                CorotatedIsotropicScratch<T, 3> s;
                Hessian9 firstPiolaDerivative;
                model.updateScratch(TM::Identity(), s);
                model.firstPiolaDerivative(s, firstPiolaDerivative);
                func(i, firstPiolaDerivative, (T)-1, false);
            }
        });
}

void MockForce::restoreStrain() // FBasedMpmForceHelper.cpp:36-44
{
    for (int i = 0; i < sim.count; ++i) sim.F[i] = sim.Fn[i];
}

// MpmForceBase.cpp:212-248
template <class Func>
void MockForce::evalInterpolantAndGradient(Func&& f, TVStack& f_eval, std::vector<TM>& grad_f)
{
    auto& grid = sim.grid;
    grid.iterateTouchedGrid([&](IV node, GridState<T, dim>& g) {
        g.new_v = TV::Zero();
    });
    grid.iterateGrid([&](IV node, GridState<T, dim>& g) {
        g.new_v = f((int)g.idx);
    });
    for (uint64_t color = 0; color < (1 << dim); ++color) {
        tbb::parallel_for(0, (int)sim.particle_group.size(), [&](int group_idx) {
            if ((sim.block_offset[group_idx] & ((1 << dim) - 1)) != color)
                return;
            for (int idx = sim.particle_group[group_idx].first; idx <= sim.particle_group[group_idx].second; ++idx) {
                int i = sim.particle_order[idx];
                TM& grad_fp = grad_f[i];
                TV& Xp = sim.X[i];
                BSplineWeights<T, dim> spline(Xp, sim.dx);
                grad_fp = TM::Zero();
                f_eval.col(i) = TV::Zero();
                grid.iterateKernel(spline, sim.particle_base_offset[i], [&](const IV& node, T w, const TV& dw, GridState<T, dim>& g) {
                    grad_fp.noalias() += g.new_v * dw.transpose();
                    f_eval.col(i) += g.new_v * w;
                });
            }
        });
    }
}

// MpmForceBase.cpp:100-153, USE_MLS_MPM == false
void MockForce::rasterizeForceToTVStack(const T scale, TVStack& force)
{
    auto& grid = sim.grid;
    grid.iterateGrid([&](IV node, GridState<T, dim>& g) {
        g.new_v = TV::Zero();
    });
    for (uint64_t color = 0; color < (1 << dim); ++color) {
        tbb::parallel_for(0, (int)sim.particle_group.size(), [&](int group_idx) {
            if ((sim.block_offset[group_idx] & ((1 << dim) - 1)) != color)
                return;
            for (int idx = sim.particle_group[group_idx].first; idx <= sim.particle_group[group_idx].second; ++idx) {
                int i = sim.particle_order[idx];
                TV& Xp = sim.X[i];
                TM4 stress_density = TM4::Zero();
                TM& stress = sim.scratch_stress[i];
                stress_density.template block<dim, dim>(0, 0) = stress;
                TV fp = sim.scratch_fp.col(i);
                stress_density.template block<dim, 1>(0, dim) = -fp;
                BSplineWeights<T, dim> spline(Xp, sim.dx);
                grid.iterateKernel(spline, sim.particle_base_offset[i],
                    [&](const IV& node, T w, const TV& dw, GridState<T, dim>& g) {
                        TV4 weight = TV4::Zero();
                        weight.template block<dim, 1>(0, 0) = dw;
                        weight(3) = w;
                        TV4 delta = (stress_density * weight);
                        TV d3 = delta.template block<dim, 1>(0, 0);
                        g.new_v -= scale * d3;
                    });
            }
        });
    }
    grid.iterateGrid([&](IV node, GridState<T, dim>& g) {
        force.col(g.idx) += g.new_v;
    });
}

// MpmForceBase.cpp:310-327 (computeVAndGradV :88-92, restoreStrain, evolveStrain -> FBasedMpmForceHelper.cpp:100-114, updateParticleImplicitState :184-209
// -> FBasedMpmForceHelper.cpp:66-96)
void MockForce::updatePositionBasedState()
{
    evalInterpolantAndGradient([&](int node_id) -> TV { TV r = sim.vn.col(node_id); r += sim.dv.col(node_id); return r; }, sim.scratch_vp, sim.scratch_gradV);
    restoreStrain();
    for (int p = 0; p < sim.count; ++p) {
        auto& F = sim.F[p];
        F = (TM::Identity() + ((T)sim.dt) * sim.scratch_gradV[p]) * F;
    }
    for (int b = 0; b < sim.count; ++b) {
        sim.scratch_fp.col(b) = TV::Zero();
        sim.scratch_stress[b] = TM::Zero();
    }
    for (int p = 0; p < sim.count; ++p) {
        auto constitutive_model = sim.model(p);
        CorotatedIsotropicScratch<T, 3> scratch;
        const auto& element_measure = sim.vol[p];
        const auto& F = sim.F[p];
        constitutive_model.updateScratch(F, scratch);
        TM vPFnT_local;
        constitutive_model.firstPiola(scratch, vPFnT_local);
        vPFnT_local = element_measure * vPFnT_local * sim.Fn[p].transpose();
        sim.scratch_stress[p] += vPFnT_local;
    }
}

// MpmForceBase.cpp:345-366 with FBasedMpmForceHelper.cpp:116-135 (one range, summed in particle order)
T MockForce::totalEnergy()
{
    double e = 0.0;
    for (int p = 0; p < sim.count; ++p) {
        auto constitutive_model = sim.model(p);
        CorotatedIsotropicScratch<T, 3> scratch;
        constitutive_model.updateScratch(sim.F[p], scratch);
        e += sim.vol[p] * constitutive_model.psi(scratch);
    }
    return e;
}

// MpmForceBase.cpp:262-306 with FBasedMpmForceHelper.cpp:137-160
void MockForce::addScaledForceDifferential(const T scale, const TVStack& dv, TVStack& df)
{
    evalInterpolantAndGradient([&](int node_id) -> TV { TV r = dv.col(node_id); return r; }, sim.scratch_vp, sim.scratch_gradV);
    for (int b = 0; b < sim.count; ++b) {
        sim.scratch_fp.col(b) = TV::Zero();
        sim.scratch_stress[b] = TM::Zero();
    }
    for (int p = 0; p < sim.count; ++p) {
        auto constitutive_model = sim.model(p);
        CorotatedIsotropicScratch<T, 3> scratch;
        constitutive_model.updateScratch(sim.F[p], scratch); // the reference keeps the scratch of the last updateImplicitState: same F
        const auto& element_measure = sim.vol[p];
        TM dP;
        const auto& Fn_local = sim.Fn[p];
        constitutive_model.firstPiolaDifferential(scratch, sim.scratch_gradV[p] * Fn_local, dP);
        sim.scratch_stress[p] += dP * element_measure * Fn_local.transpose();
    }
    rasterizeForceToTVStack(scale, df);
}

T MockInertia::totalEnergy() const
{
    T ke = 0;
    for (int p = 0; p < sim.dv.cols(); ++p) ke += sim.dv.col(p).squaredNorm() * sim.mass_matrix(p);
    return ke / 2;
}
void MockInertia::addScaledForces(T scale, TVStack& forces) const
{
    scale /= sim.dt;
    for (int p = 0; p < forces.cols(); p++)
        forces.col(p) -= scale * sim.mass_matrix(p) * sim.dv.col(p);
}
void MockInertia::addScaledForceDifferential(T scale, const TVStack& dx, TVStack& df) const
{
    scale /= sim.dt * sim.dt;
    for (int p = 0; p < df.cols(); p++)
        df.col(p) -= scale * sim.mass_matrix(p) * dx.col(p);
}

TM load9(const double* p)
{
    TM m;
    for (int q = 0; q < 9; ++q) m(q) = p[q];
    return m;
}
} // namespace

extern "C" {

// ONE simulation and ONE objective per process, like the reference: MultigridBuilder::build keeps a function-static colour-marking lambda that
// captured the id2coord vector of the first build by reference (MultigridPreconditioner.h:581) and SparseMatrix::rebuildPreconditioner a
// function-static operator (SparseMatrixFast.h:49).  create() hands out that one object, re-initialised.
static MockSim* g_sim = nullptr;

void* implicit_ref_create(double dx, double dt, const double* gravity)
{
    if (!g_sim) {
        g_sim = new MockSim();
        g_sim->force.reset(new MockForce(*g_sim));
        g_sim->forces.push_back(g_sim->force.get());
        g_sim->inertia.reset(new MockInertia(*g_sim));
        g_sim->objective.reset(new ImplicitSolverObjective<MockSim>(*g_sim));
    }
    MockSim* s = g_sim;
    s->dx = dx;
    s->D_inverse = 4 / (dx * dx);
    s->dt = dt;
    for (int d = 0; d < 3; ++d) s->gravity(d) = gravity[d];
    s->collision_nodes.clear();
    s->cn_first = true; s->cn_dPdFNorm = -1; s->cn_dPdFNorm_max = -1;
    s->objective->initialize([](TVStack&) {});
    s->objective->setPreconditioner([](const TVStack& x, TVStack& b) { b = x; });
    s->objective->matrix_free = false;
    s->objective->updated = false;
    s->objective->minres.setTolerance(1); // the constructor's values (ImplicitSolver.h:86-88): a --usecn solve overwrites them
    s->objective->cg.setTolerance(1);
    return s;
}
void implicit_ref_destroy(void*) {}
int implicit_ref_begin_step(void* h);
void implicit_ref_get_F(void* h, double* F);

// the first half of advanceOneTimeStep (MultigridSimulation.h:235-281): reinitialize -> sortParticlesAndPolluteGrid, particlesToGrid, and the part of
// startBackwardEuler that does not need the collision objects (buildMassMatrix, dv / vn sized, vn = grid velocity, backupStrain); returns num_nodes
int implicit_ref_begin_step(void* h)
{
    MockSim* s = (MockSim*)h;
    s->collision_nodes.clear();
    mpmgrid_ref_sort(s);
    int nn = mpmgrid_ref_p2g(s);
    s->mass_matrix.resize(nn); s->dv.resize(3, nn); s->vn.resize(3, nn);
    s->dv.setZero();
    s->grid.iterateGrid([&](IV node, GridState<T, dim>& g) { // buildMassMatrix :817-826, vn
        s->mass_matrix(g.idx) = g.m;
        s->vn.col(g.idx) = g.v;
    });
    for (int i = 0; i < s->count; ++i) s->Fn[i] = s->F[i]; // backupStrain
    return nn;
}

// particles (the oracle's buffer layouts: matrices column-major) -> sort -> P2G on the reference grid code; F is the strain at the start of the step
int implicit_ref_setup(void* h, long n, const double* X, const double* V, const double* mass, const double* C, const double* F, const double* vol,
    const double* mu, const double* lam, int project)
{
    MockSim* s = (MockSim*)h;
    mpmgrid_ref_set_particles(s, n, X, V, mass, C);
    s->vol.assign(vol, vol + n); s->mu.assign(mu, mu + n); s->lam.assign(lam, lam + n);
    s->F.resize(n); s->Fn.resize(n); s->scratch_stress.resize(n);
    for (long i = 0; i < n; ++i) s->Fn[i] = s->F[i] = load9(F + 9 * i); // backupStrain
    s->project_pd = project != 0;
    s->scratch_vp.resize(3, (int)n); s->scratch_fp.resize(3, (int)n);
    s->particles.Xp = &s->X; s->particles.count = (int)n; s->particles.measure = &s->vol; s->particles.m = &s->mass;
    s->Jp.assign(n, 1.0);
    return implicit_ref_begin_step(h);
}

// collision nodes (the products of buildInitialDvAndVnForNewton) and the projection MultigridSimulation::initialize hands to the objective
// (Projects/multigrid/MultigridSimulation.h:104-125): mode 1 = BC-projected system with slip nodes in their rotated frame, mode 0 = P projection
void implicit_ref_set_bc(void* h, int mode, int n_bc, const int* node_id, const double* P, const double* R, const double* Rinv, const int* slip)
{
    MockSim* s = (MockSim*)h;
    s->collision_nodes.resize(n_bc);
    for (int b = 0; b < n_bc; ++b) {
        auto& c = s->collision_nodes[b];
        c.node_id = node_id[b];
        c.P = P ? load9(P + 9 * b) : TM::Zero();
        c.R = R ? load9(R + 9 * b) : TM::Zero();
        c.Rinv = Rinv ? load9(Rinv + 9 * b) : TM::Zero();
        c.shouldRotate = slip ? slip[b] != 0 : false;
    }
    HOTSettings::systemBCProject = true;
    HOTSettings::boundaryType = mode;
    if (HOTSettings::systemBCProject && HOTSettings::boundaryType == 1)
        s->objective->initialize(
            [s](TVStack& dv) {
                for (auto iter = s->collision_nodes.begin(); iter != s->collision_nodes.end(); ++iter) {
                    int node_id = iter->node_id;
                    if (iter->shouldRotate)
                        dv(0, node_id) = 0;
                    else
                        dv.col(node_id).setZero();
                }
            });
    else
        s->objective->initialize(
            [s](TVStack& dv) {
                for (auto iter = s->collision_nodes.begin(); iter != s->collision_nodes.end(); ++iter) {
                    int node_id = (*iter).node_id;
                    TV v = dv.col(node_id);
                    dv.col(node_id) = (*iter).P * v; // CollisionNode::project (CollisionObject.h:30-34)
                }
            });
}

void implicit_ref_set_dv(void* h, const double* dv)
{
    MockSim* s = (MockSim*)h;
    for (int i = 0; i < s->num_nodes; ++i)
        for (int d = 0; d < 3; ++d) s->dv(d, i) = dv[3 * i + d];
}

// ImplicitSolverObjective::updateState (moveNodes + updatePositionBasedState + energy); returns Ek
double implicit_ref_update_state(void* h, const double* dv_in, int linesearch)
{
    MockSim* s = (MockSim*)h;
    TVStack dv(3, s->num_nodes);
    for (int i = 0; i < s->num_nodes; ++i)
        for (int d = 0; d < 3; ++d) dv(d, i) = dv_in ? dv_in[3 * i + d] : s->dv(d, i);
    HOTSettings::linesearch = linesearch != 0;
    s->objective->updated = false;
    s->objective->updateState(dv, true);
    return linesearch ? s->objective->Ek : 0.0;
}

void implicit_ref_get_F(void* h, double* F)
{
    MockSim* s = (MockSim*)h;
    for (int i = 0; i < s->count; ++i)
        for (int q = 0; q < 9; ++q) F[9 * i + q] = s->F[i](q);
}

void implicit_ref_compute_residual(void* h, double* r)
{
    MockSim* s = (MockSim*)h;
    TVStack residual(3, s->num_nodes);
    residual.setZero();
    s->objective->updated = false;
    s->objective->computeResidual(residual, true);
    for (int i = 0; i < s->num_nodes; ++i)
        for (int d = 0; d < 3; ++d) r[3 * i + d] = residual(d, i);
}

void implicit_ref_build_matrix(void* h, int bcproject, int* entryCol, double* entryVal)
{
    MockSim* s = (MockSim*)h;
    auto& O = *s->objective;
    HOTSettings::useBaselineMultigrid = false;
    if (bcproject) {
        O.rhs.resize(3, s->num_nodes);
        O.rhs.setZero();
        O.template buildMatrix<true>();
    }
    else
        O.template buildMatrix<false>();
    for (size_t e = 0; e < O.entryCol.size(); ++e) {
        entryCol[e] = O.entryCol[e];
        for (int q = 0; q < 9; ++q) entryVal[9 * e + q] = O.entryVal[e](q);
    }
}

void implicit_ref_build_diagonal(void* h, int Ainv, double* diag_inv)
{
    MockSim* s = (MockSim*)h;
    auto& O = *s->objective;
    std::vector<TM> keep = O.entryVal; // buildDiagonal reuses entryVal (ImplicitSolver.h:609)
    HOTSettings::Ainv = Ainv;
    O.buildDiagonal();
    for (int i = 0; i < s->num_nodes; ++i)
        for (int q = 0; q < 9; ++q) diag_inv[9 * i + q] = O.diagVal[i](q);
    O.entryVal = keep;
}

void implicit_ref_cn_tolerance(void* h, double eps, double dt, double* tol)
{
    MockSim* s = (MockSim*)h;
    auto& O = *s->objective;
    O.evaluatePerNodeCNTolerance(eps, dt);
    for (int i = 0; i < s->num_nodes; ++i) tol[i] = O.nodeCNTol[i];
}

// ImplicitSolverObjective::multiply: matrix-free (through the force differentials) or with the assembled block rows
void implicit_ref_multiply(void* h, int matrix_free, const double* x_in, double* b_out)
{
    MockSim* s = (MockSim*)h;
    auto& O = *s->objective;
    TVStack x(3, s->num_nodes), b(3, s->num_nodes);
    for (int i = 0; i < s->num_nodes; ++i)
        for (int d = 0; d < 3; ++d) x(d, i) = x_in[3 * i + d];
    O.matrix_free = matrix_free != 0;
    O.multiply(x, b);
    for (int i = 0; i < s->num_nodes; ++i)
        for (int d = 0; d < 3; ++d) b_out[3 * i + d] = b(d, i);
}

int implicit_ref_should_exit_by_cn(void* h, const double* r, int useCN, double cneps)
{
    MockSim* s = (MockSim*)h;
    auto& O = *s->objective;
    TVStack residual(3, s->num_nodes);
    for (int i = 0; i < s->num_nodes; ++i)
        for (int d = 0; d < 3; ++d) residual(d, i) = r[3 * i + d];
    HOTSettings::useCN = useCN != 0;
    HOTSettings::cneps = cneps;
    return O.shouldExitByCN(residual) ? 1 : 0;
}

// MultigridSimulation::backwardEulerStep (Projects/multigrid/MultigridSimulation.h:188-233) around the reference's own solver templates
// ZIRAN::ExtendedNewtonsMethod<Objective> (Lib/Ziran/Math/Nonlinear/ExtendedNewtonsMethod.h:39-66: lsolver 2 = PN-PCG / PN-MGPCG through
// ImplicitSolverObjective::computeStep, InexactConjugateGradient and MultigridBuilder / MultigridOperator) and ZIRAN::LBFGS<Objective>
// (LBFGS.h:300-437: lsolver 3 = HOT) on the objective above.  The simulation's dv (already holding buildInitialDvAndVnForNewton's start value,
// implicit_ref_set_dv) is the solution vector, as in :219-221.  out = {iterations (shouldExitByCN calls - 1), converged, final |residual|}
int implicit_ref_backward_euler_step(void* h, int lsolver, int levels, int smoother, int coarse_solver, int Ainv, int linesearch, int usecn, double cneps,
    int max_iterations, int adaptive_h, int matfree, int bcproject, int max_linear_iterations, int times, int levelscale, double topomega, double* dv_out, double* out)
{
    MockSim* s = (MockSim*)h;
    auto& objective = *s->objective;
    HOTSettings::lsolver = lsolver; HOTSettings::levelCnt = levels; HOTSettings::smoother = smoother; HOTSettings::coarseSolver = coarse_solver;
    HOTSettings::Ainv = Ainv; HOTSettings::times = times; HOTSettings::linesearch = linesearch != 0; HOTSettings::useCN = usecn != 0; HOTSettings::cneps = cneps;
    HOTSettings::useAdaptiveHessian = adaptive_h != 0; HOTSettings::debugMode = 0; HOTSettings::useBaselineMultigrid = false;
    HOTSettings::topDownMGS = false; HOTSettings::levelscale = levelscale; HOTSettings::topomega = topomega; HOTSettings::matrixFree = matfree != 0;
    objective.matrix_free = matfree != 0;
    objective.minres.max_iterations = max_linear_iterations; // the scene set-up raises both from the constructor's 20 / 10000 (MultigridInit3D.h:85-87)
    objective.cg.max_iterations = max_linear_iterations;
    HOTSettings::systemBCProject = bcproject != 0; // (--matfree runs without --bcproject: computeStep adds dRhs, which only buildMatrix<true> sizes)
    ExtendedNewtonsMethod<ImplicitSolverObjective<MockSim>> newton(objective, (T)1, max_iterations);
    LBFGS<ImplicitSolverObjective<MockSim>> lbfgs(objective, (T)1, max_iterations);
    // startBackwardEuler :166-186 (mass matrix, dv / vn and the collision nodes were set up by implicit_ref_setup / _set_bc / _set_dv)
    objective.setPreconditioner([s](const TVStack& in, TVStack& out) {
        for (int i = 0; i < s->num_nodes; i++) {
            for (int d = 0; d < dim; d++) {
                out(d, i) = in(d, i) / s->mass_matrix(i);
            }
        }
    });
    for (int i = 0; i < s->count; ++i) s->Fn[i] = s->F[i]; // force->backupStrain()
    objective.reinitialize();
    T maxcntol = -1;
    if (HOTSettings::useCN) {
        // computeCharacteristicNorm :128-164; its function-static first-step cache (`first`, dPdFNorm, dPdFNorm_max) lives in the simulation object
        // here and is reset by implicit_ref_create, i.e. once per "process": later steps keep the first step's norms even when the snow model hardens
        double& dPdFNorm = s->cn_dPdFNorm;
        double& dPdFNorm_max = s->cn_dPdFNorm_max;
        if (s->cn_first) {
            for (int i = 0; i < s->count; ++i) {
                auto model = s->model(i);
                CorotatedIsotropicScratch<T, 3> sc;
                Hessian9 firstPiolaDerivative;
                model.updateScratch(TM::Identity(), sc);
                model.firstPiolaDerivative(sc, firstPiolaDerivative);
                double curdPdFNorm = firstPiolaDerivative.norm();
                if (dPdFNorm < 0 || curdPdFNorm < dPdFNorm)
                    dPdFNorm = curdPdFNorm;
                if (dPdFNorm_max < 0 || curdPdFNorm > dPdFNorm_max)
                    dPdFNorm_max = curdPdFNorm;
            }
        }
        if (dPdFNorm > 0)
            s->cn_first = false;
        objective.evaluatePerNodeCNTolerance(HOTSettings::cneps, s->dt);
        if (dPdFNorm_max != -1)
            maxcntol = HOTSettings::cneps * s->dt * 24 * std::sqrt(s->dv.cols()) * s->dx * s->dx * dPdFNorm_max;
        newton.tolerance = maxcntol;
        lbfgs.tolerance = maxcntol;
        objective.minres.setTolerance(maxcntol);
        objective.cg.setTolerance(maxcntol);
    }
    else {
        lbfgs.tolerance = newton.tolerance = HOTSettings::cneps;
    }
    objective.isNewStep = true;
    objective.curIter = 0;
    if (HOTSettings::linesearch)
        objective.resetLSFlag(s->dv);
    bool converged;
    if (HOTSettings::lsolver != 3)
        converged = newton.solve(s->dv, false);
    else
        converged = lbfgs.solve(s->dv, false, HOTSettings::linesearch);
    out[0] = converged ? objective.curIter - 1 : max_iterations; // (curIter counts the shouldExitByCN calls)
    out[1] = converged ? 1 : 0;
    out[2] = maxcntol;
    for (int i = 0; i < s->num_nodes; ++i)
        for (int d = 0; d < 3; ++d) dv_out[3 * i + d] = s->dv(d, i);
    s->force->restoreStrain();
    return 0;
}

// the second half of advanceOneTimeStep: gridToParticles(dt) = constructNewVelocityFromNewtonResult + the G2P of mpmgrid_ref_shim.cpp, then
// force->evolveStrain(dt) (FBasedMpmForceHelper.cpp:100-114) and applyPlasticity (MpmSimulationBase.cpp:1039-1064) with the reference's return mappings:
// model 0 none, 1 VonMisesFixedCorotated(q[0]), 2 SnowPlasticity(q[0..4]) carrying Jp and hardening mu / lambda per particle
void implicit_ref_end_step(void* h, double dt, int plastic_model, const double* q, int* flags)
{
    MockSim* s = (MockSim*)h;
    std::vector<double> dv(3 * (size_t)s->num_nodes);
    for (int i = 0; i < s->num_nodes; ++i)
        for (int d = 0; d < 3; ++d) dv[3 * i + d] = s->dv(d, i);
    mpmgrid_ref_g2p(s, dv.data(), dt, flags);
    for (int p = 0; p < s->count; ++p) {
        auto& F = s->F[p];
        F = (TM::Identity() + ((T)dt) * s->scratch_gradV[p]) * F;
    }
    for (int p = 0; p < s->count && plastic_model; ++p) {
        CorotatedIsotropic<T, 3> c = s->model(p);
        if (plastic_model == 2) {
            SnowPlasticity<T> pl(q[0], q[1], q[2], q[3], q[4]);
            pl.Jp = s->Jp[p];
            pl.projectStrain(c, s->F[p]);
            s->Jp[p] = pl.Jp;
        }
        else {
            VonMisesFixedCorotated<T, 3> pl(q[0]);
            pl.projectStrain(c, s->F[p]);
        }
        s->mu[p] = c.mu;
        s->lam[p] = c.lambda;
    }
}

void implicit_ref_get_state(void* h, double* X, double* V, double* C, double* F, double* Jp, double* mu, double* lam)
{
    MockSim* s = (MockSim*)h;
    std::vector<double> G(9 * (size_t)s->count);
    mpmgrid_ref_get_particles(s, X, V, C, G.data());
    implicit_ref_get_F(h, F);
    for (int i = 0; i < s->count; ++i) { Jp[i] = s->Jp[i]; mu[i] = s->mu[i]; lam[i] = s->lam[i]; }
}

void implicit_ref_get_id2coord(void* h, int* coord)
{
    MockSim* s = (MockSim*)h;
    s->grid.iterateGrid([&](IV node, GridState<T, dim>& g) {
        for (int d = 0; d < 3; ++d) coord[3 * g.idx + d] = node(d);
    });
}

} // extern "C"
