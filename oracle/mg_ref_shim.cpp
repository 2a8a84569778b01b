// TEST INFRASTRUCTURE ONLY - built into oracle/_ref/libziran_ref.so.
// Runs the reference's OWN Galerkin-multigrid code - ZIRAN::MultigridBuilder::build, SquareMatrix::{buildDiagonal, buildTransposeMatrix,
// buildCoarseMatrix, multiply, estimate2norm}, the 8-colour 4^3-block marking, and MultigridOperator's smoothers and V-cycle
// (Projects/multigrid/{MultigridPreconditioner.h, MPMMultigridMatrix.h, SquareMatrix.h}, compiled where they lie) - on a level-0 system
// handed over as plain arrays (id2coord, entryCol, entryVal, mass: the builder's own interface, MultigridPreconditioner.h:553-554), so
// that the oracle's restatement of rows a16-a20 (oracle_matrix.inl) and the CUDA path can be pinned to the reference's code.
// Stand-ins (oracle/ref_shim): the Eigen subset these headers use (mini_eigen.h + mini_eigen_dyn.h), TBB's parallel_for /
// parallel_reduce executed serially, logging / timer macros as no-ops, inert sparse-solver classes.
#include <array>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <cstdlib>
#include <functional>
#include <memory>
#include <unordered_map>
#include <vector>
#include <Ziran/Math/Linear/DenseExt.h>
#include <Ziran/CS/Util/Logging.h>
#include <Ziran/CS/Util/ErrorContext.h>
#include <Ziran/CS/Util/Timer.h>
#include <tbb/tbb.h>
#include <MultigridPreconditioner.h>

// SquareMatrix::comp (SquareMatrix.h:39-46) falls off its end without a return value when the two colour keys are equal - undefined
// behaviour, and it IS reached (a Gauss-Seidel row compares its own key with itself for the diagonal and the padding entries).  The
// result is only tested with `< 0` / `> 0` on entries whose vector value is still zero, so any returned value gives the same sweep; GCC,
// however, compiles the missing return as unreachable code and the call crashes.  This explicit specialisation for <double, 3> defines
// the case as 0 (like the oracle and DESIGN.md section 2); everything else of the function is the reference's loop restated.
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
using T = double;
constexpr int dim = 3;
using Op = ZIRAN::MultigridOperator<T, int, dim>;
using Builder = ZIRAN::MultigridBuilder<T, int, dim>;
using TM = ZIRAN::Matrix<T, dim, dim>;
using IV = ZIRAN::Vector<int, dim>;
using TVStack = ZIRAN::Matrix<T, dim, Eigen::Dynamic>;
using Vec = ZIRAN::Vector<T, Eigen::Dynamic>;

// like SparseMatrix::rebuildPreconditioner (SparseMatrixFast.h:46-58): ONE operator, and the id2coord vector the builder's static
// colour-marking lambda captures by reference (MultigridPreconditioner.h:581) is the same object on every build
Op g_mgp;
std::vector<IV> g_id2coord;
std::vector<int> g_entryCol;
std::vector<TM> g_entryVal;
Vec g_mass, g_cntol;
TVStack g_dRhs;
std::function<void(TVStack&)> g_project = [](TVStack&) {};
int g_levels = 0;

void to_stack(const double* p, int n, TVStack& v)
{
    v.resize(dim, n);
    std::memcpy(v.data(), p, sizeof(double) * 3 * (size_t)n);
}
} // namespace

extern "C" {
// entryVal: 9 doubles per entry, column-major like Eigen's (and the oracle's) 3x3 blocks
int zr_mg_build(int n, const int* id2coord, const int* entryCol, const double* entryVal, const double* mass, int levels, int smoother,
    int coarse_solver, int Ainv, int times, int levelscale, double topomega, double cneps)
{
    HOTSettings::levelCnt = levels; HOTSettings::smoother = smoother; HOTSettings::coarseSolver = coarse_solver; HOTSettings::Ainv = Ainv;
    HOTSettings::times = times; HOTSettings::levelscale = levelscale; HOTSettings::topomega = topomega; HOTSettings::cneps = cneps;
    HOTSettings::systemBCProject = true; HOTSettings::topDownMGS = false;
    const size_t ne = (size_t)n * 125;
    g_id2coord.resize(n);
    for (int i = 0; i < n; ++i) g_id2coord[i] = IV(id2coord[3 * i], id2coord[3 * i + 1], id2coord[3 * i + 2]);
    g_entryCol.assign(entryCol, entryCol + ne);
    g_entryVal.resize(ne);
    for (size_t e = 0; e < ne; ++e) std::memcpy(g_entryVal[e].data(), entryVal + 9 * e, 9 * sizeof(double));
    g_mass.resize(n); g_cntol.resize(n);
    for (int i = 0; i < n; ++i) { g_mass(i) = mass[i]; g_cntol(i) = 0; }
    g_dRhs.resize(dim, n); g_dRhs.setZero();
    Builder b;
    b.build(g_mgp, g_mass, g_project, g_dRhs, levels, g_id2coord, g_entryCol, g_entryVal, g_cntol);
    g_levels = levels;
    return 0;
}
int zr_mg_level_dofs(int level) { return level < (int)g_mgp.dofs.size() ? g_mgp.dofs[level] : -1; }
// kind 0: system matrix of `level`, 1: prolongation level -> level + 1 (rows: fine nodes), 2: restriction (rows: coarse nodes)
int zr_mg_level_matrix(int level, int kind, int* colsize, int* entryCol, double* entryVal)
{
    const ZIRAN::SquareMatrix<T, dim>* m = kind == 0 ? g_mgp.sysmats[level]->_mat.get() : (kind == 1 ? g_mgp.promats[level]->_mat.get() : g_mgp.promats[level]->_matT.get());
    if (!m) return 1;
    *colsize = m->colsize;
    if (entryCol) std::memcpy(entryCol, m->entryCol.data(), m->entryCol.size() * sizeof(int));
    if (entryVal)
        for (size_t e = 0; e < m->entryVal.size(); ++e) std::memcpy(entryVal + 9 * e, m->entryVal[e].data(), 9 * sizeof(double));
    return 0;
}
long zr_mg_level_entries(int level, int kind)
{
    const ZIRAN::SquareMatrix<T, dim>* m = kind == 0 ? g_mgp.sysmats[level]->_mat.get() : (kind == 1 ? g_mgp.promats[level]->_mat.get() : g_mgp.promats[level]->_matT.get());
    return m ? (long)m->entryCol.size() : -1;
}
int zr_mg_level_diagonal(int level, double* D, double* Dinv)
{
    const auto& m = *g_mgp.sysmats[level]->_mat;
    const auto& inv = HOTSettings::Ainv == 0 ? m.diagonalEntry : m.diagonalBlock;
    for (size_t i = 0; i < m.diagonalVal.size(); ++i) {
        std::memcpy(D + 9 * i, m.diagonalVal[i].data(), 9 * sizeof(double));
        std::memcpy(Dinv + 9 * i, inv[i].data(), 9 * sizeof(double));
    }
    return 0;
}
int zr_mg_color_order(int level, int* out3)
{
    const auto& m = *g_mgp.sysmats[level]->_mat;
    for (size_t i = 0; i < m.colorOrder.size(); ++i)
        for (int k = 0; k < 3; ++k) out3[3 * i + k] = m.colorOrder[i][k];
    return (int)m.colorOrder.size();
}
int zr_mg_two_norm(int level, double* lmax, double* lmin)
{
    *lmax = g_mgp.sysmats[level]->_mat->lMax; *lmin = g_mgp.sysmats[level]->_mat->lMin;
    return 0;
}
// SquareMatrix::estimate2norm (SquareMatrix.h:375-475) as the reference runs it: its start vector is seeded with srand(time(NULL)).  The call is
// bracketed by time() reads (repeated if the second changed in between), so the seed is known afterwards and the +-1 start vector is regenerated with
// the same rand() sequence (SquareMatrix.h:404-417: columns outer, components inner, (rand() & 7) >= 4 ? 1 : -1) and handed back.
int zr_mg_estimate_two_norm(int level, double* start, double* lmax, double* lmin)
{
    auto& m = *g_mgp.sysmats[level]->_mat;
    for (int attempt = 0; attempt < 100; ++attempt) {
        const time_t t0 = time(NULL);
        m.estimate2norm();
        if (time(NULL) != t0) continue;
        srand(t0);
        const int n = g_mgp.dofs[level];
        for (int i = 0; i < n; ++i)
            for (int d = 0; d < dim; ++d) start[3 * i + d] = (rand() & 0x7) >= 4 ? 1.0 : -1.0;
        *lmax = m.lMax; *lmin = m.lMin;
        return 0;
    }
    return 1;
}
// lMax / lMin of a level set from outside (the Chebyshev smoother reads them)
int zr_mg_set_two_norm(int level, double lmax, double lmin)
{
    g_mgp.sysmats[level]->_mat->lMax = lmax; g_mgp.sysmats[level]->_mat->lMin = lmin;
    return 0;
}
// SparseMPMMatrix::multiply of the level's system matrix
int zr_mg_spmv(int level, const double* x, double* b)
{
    const int n = g_mgp.dofs[level];
    TVStack xv, bv;
    to_stack(x, n, xv); bv.resize(dim, n);
    g_mgp.sysmats[level]->multiply(xv, bv);
    std::memcpy(b, bv.data(), sizeof(double) * 3 * (size_t)n);
    return 0;
}
// MultigridOperator::operator() (MultigridPreconditioner.h:362-421)
int zr_mg_vcycle(const double* in, double* out)
{
    const int n = g_mgp.dofs[0];
    TVStack iv, ov;
    to_stack(in, n, iv); ov.resize(dim, n);
    g_mgp(iv, ov);
    std::memcpy(out, ov.data(), sizeof(double) * 3 * (size_t)n);
    return 0;
}
// one smoother call on a level: kind = the -smoother integer; initial_residual (nullable) feeds cg_smooth's stopping test
int zr_mg_smooth(int level, int kind, double* u, double* r, int iterations, double tolerance, const double* initial_residual)
{
    const int n = g_mgp.dofs[level];
    TVStack uv, rv;
    to_stack(u, n, uv); to_stack(r, n, rv);
    Op::level = level;
    if (initial_residual) to_stack(initial_residual, n, Op::initialResiduals[level]);
    auto& A = *g_mgp.sysmats[level];
    auto& du = g_mgp.dus[level];
    auto& dAu = g_mgp.dAus[level];
    switch (kind) {
    case 0: Op::jacobi_smooth(uv, rv, du, dAu, A, iterations, tolerance); break;
    case 1: Op::optimal_jacobi_smooth(uv, rv, du, dAu, A, iterations, tolerance); break;
    case 2: Op::cg_smooth(uv, rv, du, dAu, A, iterations, tolerance); break;
    case 5: Op::gs_smooth(uv, rv, du, dAu, A, iterations, tolerance); break;
    case 6: Op::chebyshev_smooth(uv, rv, du, dAu, A, iterations, tolerance); break;
    default: return 1;
    }
    std::memcpy(u, uv.data(), sizeof(double) * 3 * (size_t)n);
    std::memcpy(r, rv.data(), sizeof(double) * 3 * (size_t)n);
    return 0;
}
}
