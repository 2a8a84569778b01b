// TEST INFRASTRUCTURE ONLY — included by hot_oracle.cpp.  Restates the assembled-matrix / Galerkin-multigrid half of the
// hot path: a15 matrix assembly + BC projection, a16 block-ELL SpMV / diagonal, a17 hierarchy build (prolongation by
// linear weights, R = P^T, A_c = R (A P), 8-colour 4^3 block ordering), a18 symmetric block Gauss-Seidel, a19 Jacobi /
// optimal Jacobi / PCG smoothers, a20 V-cycle.  Data layout and loop structure follow the reference (fixed-width rows
// of 3x3 blocks, row-parallel SpMV, colour-serial / block-parallel / node-serial GS) so that the OpenMP timings of this
// file are the "reference CPU path" V-cycle baseline.

#include <array>
#include <unordered_map>

namespace {

// Projects/multigrid/SquareMatrix.h:11-35
struct SqMat {
    int colsize = 0;
    std::vector<int> entryCol;
    std::vector<double> entryVal; // 9 per entry, column-major 3x3
    std::vector<double> diagonalVal, diagonalEntry, diagonalBlock;
    std::array<std::vector<std::vector<int>>, 8> coloredBlockDofs;
    std::vector<std::array<int, 3>> colorOrder;
    double lMin = 1e-8, lMax = 1e2; // SquareMatrix.h:37, set by estimate2norm
    int rows() const { return colsize ? (int)(entryCol.size() / colsize) : 0; }
};

// SquareMatrix::comp (SquareMatrix.h:39-46); the reference has no return for equal keys (UB) - defined as 0 here
inline int color_comp(const std::array<int, 3>& a, const std::array<int, 3>& b)
{
    for (int i = 0; i < 3; ++i)
        if (a[i] < b[i]) return -1;
        else if (a[i] > b[i]) return 1;
    return 0;
}

inline void m3_mulv_add(const double* A, const double* x, double* y) // y += A x
{
    for (int r = 0; r < 3; ++r) y[r] += A[r] * x[0] + A[r + 3] * x[1] + A[r + 6] * x[2];
}
inline bool m3_inverse(const double* A, double* B)
{
    double c[9];
    cofactor3(A, c); // c = det * A^-T
    double det = A[0] * c[0] + A[3] * c[3] + A[6] * c[6];
    for (int r = 0; r < 3; ++r)
        for (int cc = 0; cc < 3; ++cc) B[r + 3 * cc] = c[cc + 3 * r] / det;
    return det != 0;
}

// SquareMatrix::multiply, SquareMatrix.h:477-487
void sq_multiply(const SqMat& m, const double* x, double* b)
{
    const int n = m.rows(), cs = m.colsize;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
        double sum[3] = {0, 0, 0};
        for (int idx = i * cs; idx < (i + 1) * cs; ++idx) m3_mulv_add(&m.entryVal[9 * (size_t)idx], x + 3 * (size_t)m.entryCol[idx], sum);
        b[3 * (size_t)i] = sum[0]; b[3 * (size_t)i + 1] = sum[1]; b[3 * (size_t)i + 2] = sum[2];
    }
}

// SquareMatrix::buildDiagonal, SquareMatrix.h:301-324
void sq_build_diagonal(SqMat& m, int opt)
{
    const int n = m.rows(), cs = m.colsize;
    m.diagonalVal.assign(9 * (size_t)n, 0.0);
    m.diagonalBlock.assign(9 * (size_t)n, 0.0);
    if (opt == 0) m.diagonalEntry.assign(9 * (size_t)n, 0.0);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
        double* d = &m.diagonalVal[9 * (size_t)i];
        for (int idx = i * cs; idx < (i + 1) * cs; ++idx)
            if (m.entryCol[idx] == i)
                for (int q = 0; q < 9; ++q) d[q] += m.entryVal[9 * (size_t)idx + q];
        if (opt == 0)
            for (int q = 0; q < 3; ++q) m.diagonalEntry[9 * (size_t)i + 4 * q] = 1.0 / d[4 * q];
        m3_inverse(d, &m.diagonalBlock[9 * (size_t)i]);
    }
}

// SquareMatrix::buildCoarseMatrix, SquareMatrix.h:526-571: this = l * r by per-row hash maps
void sq_build_product(SqMat& out, const SqMat& l, const SqMat& r)
{
    const int n = l.rows();
    int colsize = 0;
#pragma omp parallel for reduction(max : colsize)
    for (int i = 0; i < n; ++i) {
        std::unordered_map<int, bool> mp;
        for (int j = i * l.colsize; j < (i + 1) * l.colsize; ++j) {
            int jj = l.entryCol[j];
            for (int k = jj * r.colsize; k < (jj + 1) * r.colsize; ++k) mp[r.entryCol[k]] = true;
        }
        colsize = std::max(colsize, (int)mp.size());
    }
    out.colsize = colsize;
    out.entryCol.assign((size_t)n * colsize, 0);
    out.entryVal.assign(9 * (size_t)n * colsize, 0.0);
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < n; ++i) {
        std::unordered_map<int, std::array<double, 9>> mp;
        for (int j = i * l.colsize; j < (i + 1) * l.colsize; ++j) {
            int jj = l.entryCol[j];
            const double* L = &l.entryVal[9 * (size_t)j];
            for (int k = jj * r.colsize; k < (jj + 1) * r.colsize; ++k) {
                int kk = r.entryCol[k];
                auto it = mp.find(kk);
                if (it == mp.end()) it = mp.emplace(kk, std::array<double, 9>{}).first;
                double t[9];
                mat_mul(L, &r.entryVal[9 * (size_t)k], t);
                for (int q = 0; q < 9; ++q) it->second[q] += t[q];
            }
        }
        int idx = i * colsize;
        for (const auto& p : mp) {
            out.entryCol[idx] = p.first;
            std::copy(p.second.begin(), p.second.end(), &out.entryVal[9 * (size_t)idx]);
            idx++;
        }
        for (; idx < (i + 1) * colsize; ++idx) out.entryCol[idx] = i > 0 ? 0 : 1; // zero padding, :565-569
    }
}

// SquareMatrix::buildTransposeMatrix, SquareMatrix.h:573-607
void sq_build_transpose(SqMat& out, const SqMat& l, int rowcnt)
{
    std::vector<std::unordered_map<int, std::array<double, 9>>> data(rowcnt);
    const int n = l.rows();
    for (int i = 0; i < n; ++i)
        for (int j = i * l.colsize; j < (i + 1) * l.colsize; ++j) {
            int jj = l.entryCol[j];
            auto it = data[jj].find(i);
            if (it == data[jj].end()) it = data[jj].emplace(i, std::array<double, 9>{}).first;
            for (int q = 0; q < 9; ++q) it->second[q] += l.entryVal[9 * (size_t)j + q]; // (blocks are w*I: no block transpose in the reference)
        }
    int colsize = 0;
    for (auto& d : data) colsize = std::max(colsize, (int)d.size());
    out.colsize = colsize;
    out.entryCol.assign((size_t)rowcnt * colsize, 0);
    out.entryVal.assign(9 * (size_t)rowcnt * colsize, 0.0);
    for (int i = 0; i < rowcnt; ++i) {
        int idx = i * colsize;
        for (const auto& p : data[i]) {
            out.entryCol[idx] = p.first;
            std::copy(p.second.begin(), p.second.end(), &out.entryVal[9 * (size_t)idx]);
            idx++;
        }
        for (; idx < (i + 1) * colsize; ++idx) out.entryCol[idx] = i > 0 ? 0 : 1;
    }
}

// markColors, MultigridPreconditioner.h:582-605
void mark_colors(const std::vector<int>& id2coord, SqMat& m)
{
    const int n = (int)id2coord.size() / 3;
    m.colorOrder.assign(n, {0, 0, 0});
    for (auto& b : m.coloredBlockDofs) b.clear();
    std::array<std::unordered_map<unsigned long long, int>, 8> blockIds;
    const unsigned long long seed = 100007;
    for (int i = 0; i < n; ++i) {
        int b[3] = {id2coord[3 * i] >> 2, id2coord[3 * i + 1] >> 2, id2coord[3 * i + 2] >> 2};
        int color = ((b[0] & 1) << 2) | ((b[1] & 1) << 1) | (b[2] & 1);
        unsigned long long key = (unsigned long long)b[0] * seed * seed + (unsigned long long)b[1] * seed + (unsigned long long)b[2];
        auto it = blockIds[color].find(key);
        if (it == blockIds[color].end()) {
            it = blockIds[color].emplace(key, (int)m.coloredBlockDofs[color].size()).first;
            m.coloredBlockDofs[color].emplace_back();
        }
        auto& nodes = m.coloredBlockDofs[color][it->second];
        nodes.push_back(i);
        m.colorOrder[i] = {color, it->second, (int)nodes.size()};
    }
}

struct MatrixState {
    // level-0 assembly (ImplicitSolverObjective members, ImplicitSolver.h:49-60)
    std::vector<int> id2coord, entryCol;
    std::vector<double> entryVal, diagVal;
    bool matrix_built = false;
    // HOTSettings (Configurations.h:18-42)
    int smoother = 5, coarseSolver = 2, Ainv = 1, levelCnt = 3, times = 1, levelscale = 0;
    double topomega = 0.1;
    double cneps = 0.0; // HOTSettings::cneps of the running solve (0 outside a solve): top.tolFunc = cneps^2
    bool bcproject = true;
    // MultigridOperator state (MultigridPreconditioner.h:53-77)
    std::vector<int> dofs;
    std::vector<std::vector<int>> coords; // id2coord per level
    std::vector<SqMat> sysmats, promats, resmats;
    std::vector<std::vector<double>> residuals, initialResiduals, sols, dus, dAus, tmps;
    int level = 0;
    bool mg_built = false;
    double last_timing[10][4];
    int last_cg_iters = 0;
    std::vector<double> dv0; // accepted iterate of the last backward-Euler solve
};

MatrixState& matrix_of(Sim* s)
{
    if (!s->matrix_state) s->matrix_state = new MatrixState;
    return *(MatrixState*)s->matrix_state;
}

inline int linear_offset125(const int* d) { return (d[0] + 2) * 25 + (d[1] + 2) * 5 + d[2] + 2; } // ImplicitSolver.h:465-468

void fill_id2coord(Sim* s, std::vector<int>& out)
{
    out.assign(3 * (size_t)s->num_nodes, 0);
    orc_get_id2coord(s, out.data());
}

inline void vec_axpy(std::vector<double>& y, double a, const std::vector<double>& x)
{
    const long n = (long)y.size();
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) y[i] += a * x[i];
}
inline double vec_dot(const std::vector<double>& a, const std::vector<double>& b)
{ // MultigridOperator::dotProduct: serial Eigen array sum (MultigridPreconditioner.h:155-158)
    double s = 0;
    for (size_t i = 0; i < a.size(); ++i) s += a[i] * b[i];
    return s;
}

// scaler_func: scale_diagonal_entry_inverse / scale_diagonal_block_inverse, MultigridPreconditioner.h:143-154
void mg_scale(const MatrixState& M, const SqMat& A, const std::vector<double>& r, std::vector<double>& mr)
{
    const std::vector<double>& D = M.Ainv == 0 ? A.diagonalEntry : A.diagonalBlock;
    const int n = A.rows();
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
        double y[3] = {0, 0, 0};
        m3_mulv_add(&D[9 * (size_t)i], &r[3 * (size_t)i], y);
        mr[3 * (size_t)i] = y[0]; mr[3 * (size_t)i + 1] = y[1]; mr[3 * (size_t)i + 2] = y[2];
    }
}

using Vd = std::vector<double>;

// A.project: only level 0 without --bcproject carries the BC projection (MultigridPreconditioner.h:695-699)
void mg_project(Sim* s, MatrixState& M, int level, Vd& v)
{
    if (level == 0 && !M.bcproject) bc_project(s, v.data());
}

// jacobi_smooth, MultigridPreconditioner.h:160-173
void jacobi_smooth(Sim* s, MatrixState& M, int level, Vd& u, Vd& r, Vd& du, Vd& dAu, int iterations)
{
    const SqMat& A = M.sysmats[level];
    for (; iterations--;) {
        mg_scale(M, A, r, du);
        for (auto& v : du) v *= M.topomega;
        vec_axpy(u, 1.0, du);
        sq_multiply(A, du.data(), dAu.data());
        mg_project(s, M, level, dAu);
        vec_axpy(r, -1.0, dAu);
    }
}
// optimal_jacobi_smooth, :174-189
void optimal_jacobi_smooth(Sim* s, MatrixState& M, int level, Vd& u, Vd& r, Vd& du, Vd& dAu, int iterations, double tolerance)
{
    const SqMat& A = M.sysmats[level];
    for (; iterations--;) {
        if (std::sqrt(vec_dot(r, r)) < tolerance) break;
        mg_scale(M, A, r, du);
        sq_multiply(A, du.data(), dAu.data());
        mg_project(s, M, level, dAu);
        double omega = vec_dot(du, r) / vec_dot(du, dAu);
        vec_axpy(u, omega, du);
        vec_axpy(r, -omega, dAu);
    }
}
// chebyshev_smooth, MultigridPreconditioner.h:227-264: d, c from the 2-norm estimate of A (lMax, lMin = lMax / 30), applied to D^-1 A
void chebyshev_smooth(Sim* s, MatrixState& M, int level, Vd& u, Vd& r, Vd& du, Vd& dAu, int iterations)
{
    const SqMat& A = M.sysmats[level];
    Vd& p = M.tmps[level];
    const double d = (A.lMax + A.lMin) / 2, c = (A.lMax - A.lMin) / 2;
    int cnt = 1;
    iterations--;
    mg_scale(M, A, r, p);
    double alpha = 1 / d, beta;
    du = p;
    sq_multiply(A, du.data(), dAu.data());
    mg_project(s, M, level, dAu);
    vec_axpy(u, alpha, du);
    vec_axpy(r, -alpha, dAu);
    for (; iterations-- > 0; ++cnt) {
        mg_scale(M, A, r, p);
        beta = 0.5 * c * c * alpha * alpha;
        if (cnt > 1) beta *= 0.5;
        alpha = 1 / (d - beta / alpha);
        for (size_t i = 0; i < du.size(); ++i) du[i] = p[i] + beta * du[i];
        sq_multiply(A, du.data(), dAu.data());
        mg_project(s, M, level, dAu);
        vec_axpy(u, alpha, du);
        vec_axpy(r, -alpha, dAu);
    }
}
// SquareMatrix::estimate2norm (SquareMatrix.h:375-475): power iteration on A A from a +-1 start vector.  The reference seeds the
// start with srand(time(NULL)); any start converges to the same 2-norm within `tol`, here a fixed hash of the entry index.
inline double sign_pattern(size_t t) { return ((uint32_t)(t * 2654435761u) >> 16) & 1u ? 1.0 : -1.0; }
void estimate2norm(SqMat& A, double tol = 1e-6, const double* start = nullptr /* the +-1 start vector (tests hand over the reference's own) */)
{
    const int MaxIters = 512;
    const size_t m = 3 * (size_t)A.rows();
    Vd v(m), x(m);
    for (size_t t = 0; t < m; ++t) v[t] = start ? start[t] : sign_pattern(t);
    sq_multiply(A, v.data(), x.data());
    for (auto& a : x) a = std::fabs(a);
    double e = std::sqrt(vec_dot(x, x));
    if (e == 0) {
        A.lMin = A.lMax = 0;
        return;
    }
    for (auto& a : x) a /= e;
    double e0 = 0;
    for (int iter = 0; iter < MaxIters && std::fabs(e - e0) > tol * e; iter++) {
        e0 = e;
        sq_multiply(A, x.data(), v.data());
        sq_multiply(A, v.data(), x.data());
        const double normx = std::sqrt(vec_dot(x, x));
        e = normx / std::sqrt(vec_dot(v, v));
        for (auto& a : x) a /= normx;
    }
    A.lMax = e;
    A.lMin = A.lMax / 30; // "experience"
}

// cg_smooth, :190-226
void cg_smooth(Sim* s, MatrixState& M, int level, Vd& u, Vd& r, Vd& du, Vd& dAu, int iterations)
{
    const SqMat& A = M.sysmats[level];
    Vd& z = M.tmps[level];
    mg_scale(M, A, M.initialResiduals[level], z);
    double zTrk0 = vec_dot(z, M.initialResiduals[level]);
    mg_scale(M, A, r, z);
    du = z;
    double zTrk = vec_dot(z, r);
    const double cgratio = 0.5;
    double tolerance = zTrk0 * cgratio * cgratio;
    int cnt = 0;
    for (; iterations--;) {
        if (zTrk < tolerance) break;
        sq_multiply(A, du.data(), dAu.data());
        mg_project(s, M, level, dAu);
        double omega = zTrk / vec_dot(dAu, du);
        vec_axpy(u, omega, du);
        vec_axpy(r, -omega, dAu);
        mg_scale(M, A, r, z);
        double zTrkPre = zTrk;
        zTrk = vec_dot(z, r);
        double beta = zTrk / zTrkPre;
        const long n = (long)du.size();
#pragma omp parallel for schedule(static)
        for (long i = 0; i < n; ++i) du[i] = z[i] + beta * du[i];
        ++cnt;
    }
    M.last_cg_iters = cnt;
}
// gs_smooth, :266-318
void gs_smooth(Sim* s, MatrixState& M, int level, Vd& u, Vd& r, Vd& du, Vd& dAu, int iterations)
{
    const SqMat& mat = M.sysmats[level];
    Vd& hdu = M.tmps[level];
    const int cs = mat.colsize, n = mat.rows();
    iterations = (iterations + 1) >> 1;
    for (; iterations--;) {
        std::fill(hdu.begin(), hdu.end(), 0.0);
        for (int c = 0; c < 8; ++c) {
            const auto& blocks = mat.coloredBlockDofs[c];
#pragma omp parallel for schedule(dynamic, 1)
            for (int bid = 0; bid < (int)blocks.size(); ++bid)
                for (int ii = 0; ii < (int)blocks[bid].size(); ++ii) {
                    const int i = blocks[bid][ii];
                    double sigma[3] = {0, 0, 0};
                    for (int idx = i * cs; idx < (i + 1) * cs; ++idx) {
                        const int col = mat.entryCol[idx];
                        if (color_comp(mat.colorOrder[col], mat.colorOrder[i]) < 0) m3_mulv_add(&mat.entryVal[9 * (size_t)idx], &hdu[3 * (size_t)col], sigma);
                    }
                    double rhs[3] = {r[3 * (size_t)i] - sigma[0], r[3 * (size_t)i + 1] - sigma[1], r[3 * (size_t)i + 2] - sigma[2]}, y[3] = {0, 0, 0};
                    m3_mulv_add(&mat.diagonalBlock[9 * (size_t)i], rhs, y);
                    hdu[3 * (size_t)i] = y[0]; hdu[3 * (size_t)i + 1] = y[1]; hdu[3 * (size_t)i + 2] = y[2];
                }
        }
        for (int i = 0; i < n; ++i) { // serial in the reference (:292-293)
            double y[3] = {0, 0, 0};
            m3_mulv_add(&mat.diagonalVal[9 * (size_t)i], &hdu[3 * (size_t)i], y);
            hdu[3 * (size_t)i] = y[0]; hdu[3 * (size_t)i + 1] = y[1]; hdu[3 * (size_t)i + 2] = y[2];
        }
        std::fill(du.begin(), du.end(), 0.0);
        for (int c = 7; c >= 0; --c) {
            const auto& blocks = mat.coloredBlockDofs[c];
#pragma omp parallel for schedule(dynamic, 1)
            for (int bid = 0; bid < (int)blocks.size(); ++bid)
                for (int ii = (int)blocks[bid].size() - 1; ii >= 0; --ii) {
                    const int i = blocks[bid][ii];
                    double sigma[3] = {0, 0, 0};
                    for (int idx = i * cs; idx < (i + 1) * cs; ++idx) {
                        const int col = mat.entryCol[idx];
                        if (color_comp(mat.colorOrder[col], mat.colorOrder[i]) > 0) m3_mulv_add(&mat.entryVal[9 * (size_t)idx], &du[3 * (size_t)col], sigma);
                    }
                    double rhs[3] = {hdu[3 * (size_t)i] - sigma[0], hdu[3 * (size_t)i + 1] - sigma[1], hdu[3 * (size_t)i + 2] - sigma[2]}, y[3] = {0, 0, 0};
                    m3_mulv_add(&mat.diagonalBlock[9 * (size_t)i], rhs, y);
                    du[3 * (size_t)i] = y[0]; du[3 * (size_t)i + 1] = y[1]; du[3 * (size_t)i + 2] = y[2];
                }
        }
        vec_axpy(u, 1.0, du);
        sq_multiply(mat, du.data(), dAu.data());
        mg_project(s, M, level, dAu);
        vec_axpy(r, -1.0, dAu);
    }
}

// selectSmoother, MultigridPreconditioner.h:496-521 (integer codes of -smoother / -coarseSolver)
int run_smoother(Sim* s, MatrixState& M, int kind, int level, Vd& u, Vd& r, Vd& du, Vd& dAu, int iterations, double tolerance)
{
    M.level = level;
    switch (kind) {
    case 0: jacobi_smooth(s, M, level, u, r, du, dAu, iterations); return 0;
    case 1: optimal_jacobi_smooth(s, M, level, u, r, du, dAu, iterations, tolerance); return 0;
    case 2: cg_smooth(s, M, level, u, r, du, dAu, iterations); return 0;
    case 5: gs_smooth(s, M, level, u, r, du, dAu, iterations); return 0;
    case 6: chebyshev_smooth(s, M, level, u, r, du, dAu, iterations); return 0;
    default: return fail(s, "No proper smoother is selected! (supported: 0 Jacobi, 1 optimal Jacobi, 2 PCG, 5 GS, 6 Chebyshev)");
    }
}

// iteration policy, setup_parameters MultigridPreconditioner.h:524-551 (topDownMGS = false)
inline int mg_regular_iters(const MatrixState& M, int level) { return M.times + level * M.levelscale; }
inline int mg_top_iters(const MatrixState& M, int level)
{
    if (M.levelCnt == 1) return mg_regular_iters(M, level);
    if (!(M.coarseSolver == 2 || M.coarseSolver == 6)) return mg_regular_iters(M, level) * 3;
    return 10000;
}

// MultigridOperator::operator(), MultigridPreconditioner.h:362-421
int mg_vcycle(Sim* s, MatrixState& M, const double* in, double* out)
{
    const int L = (int)M.sysmats.size();
    for (int i = 0; i < 10; ++i)
        for (int j = 0; j < 4; ++j) M.last_timing[i][j] = 0;
    auto now = [] { return omp_get_wtime(); };
    M.residuals[0].assign(in, in + 3 * (size_t)M.dofs[0]); // correctResidualProjection adds dRhs == 0 (ImplicitSolver.h:483-486)
    Vd outv(3 * (size_t)M.dofs[0], 0.0);
    if (L > 1) sq_multiply(M.resmats[0], M.residuals[0].data(), M.initialResiduals[1].data());
    else M.initialResiduals[0] = M.residuals[0];
    for (int l = 1; l < L - 1; ++l) sq_multiply(M.resmats[l], M.initialResiduals[l].data(), M.initialResiduals[l + 1].data());
    int level, rc = 0;
    const double top_tol = M.cneps * M.cneps; // top.tolFunc (MultigridPreconditioner.h:529): optimal Jacobi stops early on it
    for (level = 0; level < L - 1; ++level) {
        Vd& sol = level == 0 ? outv : M.sols[level];
        double t0 = now();
        rc = run_smoother(s, M, M.smoother, level, sol, M.residuals[level], M.dus[level], M.dAus[level], mg_regular_iters(M, level), 0.0);
        if (rc) return rc;
        M.last_timing[level][0] += now() - t0;
        t0 = now();
        sq_multiply(M.resmats[level], M.residuals[level].data(), M.residuals[level + 1].data());
        M.last_timing[level][1] += now() - t0;
        std::fill(M.sols[level + 1].begin(), M.sols[level + 1].end(), 0.0);
    }
    {
        double t0 = now();
        rc = run_smoother(s, M, M.coarseSolver, level, level == 0 ? outv : M.sols[level], M.residuals[level], M.dus[level], M.dAus[level],
            mg_top_iters(M, level), top_tol);
        if (rc) return rc;
        M.last_timing[level][0] += now() - t0;
    }
    for (--level; level >= 0; --level) {
        Vd& sol = level == 0 ? outv : M.sols[level];
        double t0 = now();
        sq_multiply(M.promats[level], M.sols[level + 1].data(), M.dus[level].data());
        M.last_timing[level][2] += now() - t0;
        t0 = now();
        vec_axpy(sol, 1.0, M.dus[level]);
        sq_multiply(M.sysmats[level], M.dus[level].data(), M.dAus[level].data());
        vec_axpy(M.residuals[level], -1.0, M.dAus[level]);
        M.last_timing[level][3] += now() - t0;
        t0 = now();
        rc = run_smoother(s, M, M.smoother, level, sol, M.residuals[level], M.dus[level], M.dAus[level], mg_regular_iters(M, level), 0.0);
        if (rc) return rc;
        M.last_timing[level][0] += now() - t0;
    }
    std::copy(outv.begin(), outv.end(), out);
    return 0;
}

} // namespace

extern "C" {

// a15: ImplicitSolverObjective::buildMatrix<projectSystem>, ImplicitSolver.h:470-603
int orc_build_matrix(void* h, int bcproject)
{
    Sim* s = (Sim*)h;
    ForceState& f = force_of(s);
    MatrixState& M = matrix_of(s);
    if ((long)f.scratch.size() != s->N) return fail(s, "orc_build_matrix: call orc_update_state first");
    const int nn = s->num_nodes;
    fill_id2coord(s, M.id2coord);
    M.entryCol.assign((size_t)nn * 125, -1);
    M.entryVal.assign(9 * (size_t)nn * 125, 0.0);
    const int zero3[3] = {0, 0, 0};
    for (int i = 0; i < nn; ++i) { // inertia term
        size_t e = (size_t)i * 125 + linear_offset125(zero3);
        M.entryCol[e] = i;
        for (int q = 0; q < 3; ++q) M.entryVal[9 * e + 4 * q] = s->mass_matrix[i];
    }
    const double force_scale = s->dt * s->dt;
    s->for_colored_groups([&](int grp) {
        for (int sidx = s->groups[grp].first; sidx <= s->groups[grp].second; ++sidx) {
            int i = s->order[sidx];
            // runLambdaWithDifferential (FBasedMpmForceHelper.h:63-121): updateScratch(F) + firstPiolaDerivative
            Scratch sc;
            update_scratch(&s->F[9 * i], s->mu[i], s->lambda[i], f.project, sc);
            double ddF[81];
            first_piola_derivative(sc, ddF);
            const double* Fn = &f.Fn[9 * i];
            const double vol = s->vol[i];
            Spline sp(&s->X[3 * i], s->dx);
            double cw[27][3];
            int cnode[27][3], cidx[27], cnt = 0;
            s->iterate_kernel(sp, s->base_offset[i], [&](const int* node, double, const double* dw, GridState& g) {
                if (g.idx < 0) return;
                for (int v = 0; v < 3; ++v) cw[cnt][v] = Fn[3 * v] * dw[0] + Fn[3 * v + 1] * dw[1] + Fn[3 * v + 2] * dw[2]; // Fn^T dw
                cnode[cnt][0] = node[0]; cnode[cnt][1] = node[1]; cnode[cnt][2] = node[2];
                cidx[cnt++] = (int)g.idx;
            });
            for (int a = 0; a < cnt; ++a)
                for (int b = 0; b < cnt; ++b) {
                    if (cidx[b] < cidx[a]) continue;
                    double delta[9] = {0};
                    for (int q = 0; q < 3; ++q)
                        for (int v = 0; v < 3; ++v) {
                            const double ww = cw[a][v] * cw[b][q];
                            for (int cc = 0; cc < 3; ++cc)
                                for (int rr = 0; rr < 3; ++rr) delta[rr + 3 * cc] += ddF[(3 * v + rr) + 9 * (3 * q + cc)] * ww;
                        }
                    for (int q = 0; q < 9; ++q) delta[q] *= force_scale * vol;
                    int d[3] = {cnode[a][0] - cnode[b][0], cnode[a][1] - cnode[b][1], cnode[a][2] - cnode[b][2]};
                    size_t e = (size_t)cidx[a] * 125 + linear_offset125(d);
                    M.entryCol[e] = cidx[b];
                    for (int q = 0; q < 9; ++q) M.entryVal[9 * e + q] += delta[q];
                    if (cidx[a] != cidx[b]) {
                        int dn[3] = {-d[0], -d[1], -d[2]};
                        size_t et = (size_t)cidx[b] * 125 + linear_offset125(dn);
                        M.entryCol[et] = cidx[a];
                        for (int cc = 0; cc < 3; ++cc)
                            for (int rr = 0; rr < 3; ++rr) M.entryVal[9 * et + rr + 3 * cc] += delta[cc + 3 * rr];
                    }
                }
        }
    });
    M.bcproject = bcproject != 0;
    if (bcproject) { // :554-593
        std::vector<int> bc_of(nn, -1);
        for (size_t b = 0; b < f.bc_node.size(); ++b) bc_of[f.bc_node[b]] = (int)b;
#pragma omp parallel for schedule(static)
        for (int i = 0; i < nn; ++i) {
            const int bi = bc_of[i];
            const bool iCollide = bi >= 0, iSlip = iCollide ? f.bc_slip[bi] != 0 : false;
            for (size_t st = (size_t)i * 125; st < (size_t)(i + 1) * 125; ++st) {
                int j = M.entryCol[st];
                if (j == -1) {
                    M.entryCol[st] = i > 0 ? 0 : 1;
                    continue;
                }
                const int bj = bc_of[j];
                const bool jCollide = bj >= 0;
                if (!iCollide && !jCollide) continue;
                const bool jSlip = jCollide ? f.bc_slip[bj] != 0 : false;
                double* val = &M.entryVal[9 * st];
                if ((iCollide && !iSlip) || (jCollide && !jSlip)) {
                    for (int q = 0; q < 9; ++q) val[q] = 0;
                    if (j == i) val[0] = val[4] = val[8] = 1;
                    continue;
                }
                if (iSlip) mat_mul(&f.bc_R[9 * (size_t)bi], val, val);
                if (jSlip) mat_mul(val, &f.bc_Rinv[9 * (size_t)bj], val);
                if (iSlip) val[0] = val[3] = val[6] = 0;
                if (jSlip) val[0] = val[1] = val[2] = 0;
                if (i == j) val[0] = 1;
            }
        }
    }
    else {
        for (int i = 0; i < nn; ++i)
            for (size_t st = (size_t)i * 125; st < (size_t)(i + 1) * 125; ++st)
                if (M.entryCol[st] == -1) M.entryCol[st] = i > 0 ? 0 : 1;
    }
    M.matrix_built = true;
    M.mg_built = false;
    return 0;
}

// ImplicitSolverObjective::buildDiagonal (matrix-free block-Jacobi preconditioner), ImplicitSolver.h:605-665
int orc_build_diagonal(void* h, int Ainv, double* diag_inv /* 9 per node, nullable */)
{
    Sim* s = (Sim*)h;
    ForceState& f = force_of(s);
    MatrixState& M = matrix_of(s);
    if ((long)f.scratch.size() != s->N) return fail(s, "orc_build_diagonal: call orc_update_state first");
    const int nn = s->num_nodes;
    std::vector<double> D(9 * (size_t)nn, 0.0);
    for (int i = 0; i < nn; ++i)
        for (int q = 0; q < 3; ++q) D[9 * (size_t)i + 4 * q] = s->mass_matrix[i];
    const double force_scale = s->dt * s->dt;
    s->for_colored_groups([&](int grp) {
        for (int sidx = s->groups[grp].first; sidx <= s->groups[grp].second; ++sidx) {
            int i = s->order[sidx];
            Scratch sc;
            update_scratch(&s->F[9 * i], s->mu[i], s->lambda[i], f.project, sc);
            double ddF[81];
            first_piola_derivative(sc, ddF);
            const double* Fn = &f.Fn[9 * i];
            Spline sp(&s->X[3 * i], s->dx);
            s->iterate_kernel(sp, s->base_offset[i], [&](const int*, double, const double* dw, GridState& g) {
                if (g.idx < 0) return;
                double w[3];
                for (int v = 0; v < 3; ++v) w[v] = Fn[3 * v] * dw[0] + Fn[3 * v + 1] * dw[1] + Fn[3 * v + 2] * dw[2];
                double* d = &D[9 * (size_t)g.idx];
                for (int q = 0; q < 3; ++q)
                    for (int v = 0; v < 3; ++v)
                        for (int cc = 0; cc < 3; ++cc)
                            for (int rr = 0; rr < 3; ++rr) d[rr + 3 * cc] += force_scale * s->vol[i] * ddF[(3 * v + rr) + 9 * (3 * q + cc)] * w[v] * w[q];
            });
        }
    });
    M.diagVal.assign(9 * (size_t)nn, 0.0);
    for (int i = 0; i < nn; ++i) {
        if (Ainv == 0)
            for (int q = 0; q < 3; ++q) M.diagVal[9 * (size_t)i + 4 * q] = 1.0 / D[9 * (size_t)i + 4 * q];
        else m3_inverse(&D[9 * (size_t)i], &M.diagVal[9 * (size_t)i]);
    }
    if (diag_inv) std::copy(M.diagVal.begin(), M.diagVal.end(), diag_inv);
    return 0;
}

int orc_get_matrix(void* h, int* entryCol, double* entryVal)
{
    Sim* s = (Sim*)h;
    MatrixState& M = matrix_of(s);
    if (!M.matrix_built) return fail(s, "orc_get_matrix: call orc_build_matrix first");
    if (entryCol) std::copy(M.entryCol.begin(), M.entryCol.end(), entryCol);
    if (entryVal) std::copy(M.entryVal.begin(), M.entryVal.end(), entryVal);
    return 0;
}

// MultigridBuilder::build, MultigridPreconditioner.h:553-703
int orc_build_mg(void* h, int levels, int smoother, int coarseSolver, int Ainv, int times, int levelscale, double topomega)
{
    Sim* s = (Sim*)h;
    MatrixState& M = matrix_of(s);
    if (!M.matrix_built) return fail(s, "orc_build_mg: call orc_build_matrix first");
    if (levels < 1 || levels > 10) return fail(s, "Level depth exceeds 10! Too Deep!");
    if (!M.bcproject && levels > 1) return fail(s, "multigrid needs the BC-projected system (ImplicitSolver.h:339)");
    M.levelCnt = levels; M.smoother = smoother; M.coarseSolver = coarseSolver; M.Ainv = Ainv; M.times = times; M.levelscale = levelscale;
    M.topomega = topomega;
    M.sysmats.assign(levels, SqMat());
    M.promats.assign(levels - 1, SqMat());
    M.resmats.assign(levels - 1, SqMat());
    M.coords.assign(levels, {});
    M.sysmats[0].colsize = 125;
    M.sysmats[0].entryCol = M.entryCol;
    M.sysmats[0].entryVal = M.entryVal;
    sq_build_diagonal(M.sysmats[0], Ainv);
    M.coords[0] = M.id2coord;
    const bool colors = coarseSolver == 5 || smoother == 5;
    if (colors) mark_colors(M.coords[0], M.sysmats[0]);
    if ((coarseSolver == 6 && levels == 1) || (smoother == 6 && levels > 1)) estimate2norm(M.sysmats[0]); // :610-611
    M.dofs.assign(1, (int)M.id2coord.size() / 3);
    const double w1d[2][3] = {{0.0, 1.0, 0.0}, {0.0, 0.5, 0.5}}; // linear_weight_template :445-466
    const unsigned long long seed = 100007;
    for (int level = 0; level < levels - 1; ++level) {
        const std::vector<int>& fine = M.coords[level];
        const int nf = (int)fine.size() / 3;
        std::vector<int> coarse;
        std::unordered_map<unsigned long long, int> coord2id;
        SqMat& P = M.promats[level];
        P.colsize = 8;
        P.entryCol.assign((size_t)nf * 8, 0);
        P.entryVal.assign(9 * (size_t)nf * 8, 0.0);
        for (int i = 0; i < nf; ++i) {
            const int x = fine[3 * i], y = fine[3 * i + 1], z = fine[3 * i + 2];
            const int loc[3] = {x & 1, y & 1, z & 1};
            for (int nx = x / 2; nx <= x / 2 + 1; ++nx)
                for (int ny = y / 2; ny <= y / 2 + 1; ++ny)
                    for (int nz = z / 2; nz <= z / 2 + 1; ++nz) {
                        const int lin = (nx - x / 2) * 4 + (ny - y / 2) * 2 + nz - z / 2;
                        const double weight = w1d[loc[0]][nx - x / 2 + 1] * w1d[loc[1]][ny - y / 2 + 1] * w1d[loc[2]][nz - z / 2 + 1];
                        if (weight == 0) {
                            P.entryCol[(size_t)i * 8 + lin] = P.entryCol[(size_t)i * 8];
                            continue;
                        }
                        unsigned long long key = (unsigned long long)nx * seed * seed + (unsigned long long)ny * seed + (unsigned long long)nz;
                        auto it = coord2id.find(key);
                        if (it == coord2id.end()) {
                            coarse.push_back(nx); coarse.push_back(ny); coarse.push_back(nz);
                            it = coord2id.emplace(key, (int)coarse.size() / 3 - 1).first;
                        }
                        P.entryCol[(size_t)i * 8 + lin] = it->second;
                        for (int q = 0; q < 3; ++q) P.entryVal[9 * ((size_t)i * 8 + lin) + 4 * q] = weight;
                    }
        }
        M.coords[level + 1] = coarse;
        const int nc = (int)coarse.size() / 3;
        sq_build_transpose(M.resmats[level], P, nc);
        SqMat AP;
        sq_build_product(AP, M.sysmats[level], P);
        sq_build_product(M.sysmats[level + 1], M.resmats[level], AP);
        sq_build_diagonal(M.sysmats[level + 1], Ainv);
        if (colors) mark_colors(M.coords[level + 1], M.sysmats[level + 1]);
        if ((coarseSolver == 6 && level + 2 == levels) || (smoother == 6 && level + 2 < levels)) estimate2norm(M.sysmats[level + 1]); // :682-683
        M.dofs.push_back(nc);
    }
    // MultigridOperator::init :83-118
    auto alloc = [&](std::vector<Vd>& v) {
        v.assign(levels, Vd());
        for (int l = 0; l < levels; ++l) v[l].assign(3 * (size_t)M.dofs[l], 0.0);
    };
    alloc(M.residuals); alloc(M.initialResiduals); alloc(M.sols); alloc(M.dus); alloc(M.dAus); alloc(M.tmps);
    M.mg_built = true;
    return 0;
}

int orc_mg_levels(void* h) { return (int)matrix_of((Sim*)h).sysmats.size(); }
int orc_estimate_2norm(void* h, int level, double* lmax_lmin)
{
    MatrixState& M = matrix_of((Sim*)h);
    if (level < 0 || level >= (int)M.sysmats.size()) return -1;
    estimate2norm(M.sysmats[level]);
    lmax_lmin[0] = M.sysmats[level].lMax; lmax_lmin[1] = M.sysmats[level].lMin;
    return 0;
}
// the same from a given start vector (3 x dofs of the level): what tests/test_oracle_mg_ref.py uses to follow the reference's own seeded start
int orc_estimate_2norm_from(void* h, int level, const double* start, double* lmax_lmin)
{
    MatrixState& M = matrix_of((Sim*)h);
    if (level < 0 || level >= (int)M.sysmats.size()) return -1;
    estimate2norm(M.sysmats[level], 1e-6, start);
    lmax_lmin[0] = M.sysmats[level].lMax; lmax_lmin[1] = M.sysmats[level].lMin;
    return 0;
}
int orc_get_level_dofs(void* h, int* dofs)
{
    MatrixState& M = matrix_of((Sim*)h);
    std::copy(M.dofs.begin(), M.dofs.end(), dofs);
    return 0;
}
int orc_get_level_coords(void* h, int level, int* coord)
{
    MatrixState& M = matrix_of((Sim*)h);
    std::copy(M.coords[level].begin(), M.coords[level].end(), coord);
    return 0;
}
// kind 0: system matrix, 1: prolongation, 2: restriction.  First call with col == val == NULL to get colsize.
int orc_get_level_matrix(void* h, int level, int kind, int* colsize, int* col, double* val)
{
    Sim* s = (Sim*)h;
    MatrixState& M = matrix_of(s);
    if (!M.mg_built) return fail(s, "orc_get_level_matrix: call orc_build_mg first");
    const SqMat& m = kind == 0 ? M.sysmats[level] : (kind == 1 ? M.promats[level] : M.resmats[level]);
    if (colsize) *colsize = m.colsize;
    if (col) std::copy(m.entryCol.begin(), m.entryCol.end(), col);
    if (val) std::copy(m.entryVal.begin(), m.entryVal.end(), val);
    return 0;
}
int orc_get_level_diagonal(void* h, int level, double* diagonalVal, double* diagonalInv)
{
    MatrixState& M = matrix_of((Sim*)h);
    const SqMat& m = M.sysmats[level];
    if (diagonalVal) std::copy(m.diagonalVal.begin(), m.diagonalVal.end(), diagonalVal);
    if (diagonalInv) {
        const Vd& D = M.Ainv == 0 ? m.diagonalEntry : m.diagonalBlock;
        std::copy(D.begin(), D.end(), diagonalInv);
    }
    return 0;
}
int orc_get_color_order(void* h, int level, int* order3)
{
    MatrixState& M = matrix_of((Sim*)h);
    const SqMat& m = M.sysmats[level];
    for (size_t i = 0; i < m.colorOrder.size(); ++i)
        for (int d = 0; d < 3; ++d) order3[3 * i + d] = m.colorOrder[i][d];
    return 0;
}

// a16: SquareMatrix::multiply on level `level` (level 0 before build_mg: SparseMatrix::multiply, SparseMatrixFast.h:60-73)
int orc_spmv(void* h, int level, const double* x, double* b)
{
    Sim* s = (Sim*)h;
    MatrixState& M = matrix_of(s);
    if (!M.matrix_built) return fail(s, "orc_spmv: call orc_build_matrix first");
    if (level == 0 && !M.mg_built) {
        SqMat tmp;
        tmp.colsize = 125;
        tmp.entryCol.swap(M.entryCol); tmp.entryVal.swap(M.entryVal);
        sq_multiply(tmp, x, b);
        tmp.entryCol.swap(M.entryCol); tmp.entryVal.swap(M.entryVal);
        return 0;
    }
    if (level < 0 || level >= (int)M.sysmats.size()) return fail(s, "orc_spmv: bad level");
    sq_multiply(M.sysmats[level], x, b);
    return 0;
}
// SparseMPMMatrix::transposeMultiply / multiply, MPMMultigridMatrix.h:63-70
int orc_restrict(void* h, int level, const double* fine, double* coarse)
{
    Sim* s = (Sim*)h;
    MatrixState& M = matrix_of(s);
    if (!M.mg_built || level < 0 || level + 1 >= (int)M.sysmats.size()) return fail(s, "orc_restrict: bad level");
    sq_multiply(M.resmats[level], fine, coarse);
    return 0;
}
int orc_prolong(void* h, int level, const double* coarse, double* fine)
{
    Sim* s = (Sim*)h;
    MatrixState& M = matrix_of(s);
    if (!M.mg_built || level < 0 || level + 1 >= (int)M.sysmats.size()) return fail(s, "orc_prolong: bad level");
    sq_multiply(M.promats[level], coarse, fine);
    return 0;
}

// one smoother call with the reference's signature smoothFunc(u, r, du, dAu, A, iterations, tolerance)
// (MultigridPreconditioner.h:68-73).  For kind 2 (PCG) the stopping test uses initial_residual (initialResiduals[level]).
int orc_smooth(void* h, int level, int kind, double* u, double* r, int iterations, double tolerance, const double* initial_residual)
{
    Sim* s = (Sim*)h;
    MatrixState& M = matrix_of(s);
    if (!M.mg_built || level < 0 || level >= (int)M.sysmats.size()) return fail(s, "orc_smooth: bad level");
    const size_t n = 3 * (size_t)M.dofs[level];
    Vd uu(u, u + n), rr(r, r + n);
    if (initial_residual) M.initialResiduals[level].assign(initial_residual, initial_residual + n);
    int rc = run_smoother(s, M, kind, level, uu, rr, M.dus[level], M.dAus[level], iterations, tolerance);
    if (rc) return rc;
    std::copy(uu.begin(), uu.end(), u);
    std::copy(rr.begin(), rr.end(), r);
    return 0;
}

// a20
int orc_vcycle(void* h, const double* in, double* out)
{
    Sim* s = (Sim*)h;
    MatrixState& M = matrix_of(s);
    if (!M.mg_built) return fail(s, "orc_vcycle: call orc_build_mg first");
    return mg_vcycle(s, M, in, out);
}
// the reference's per-level [smooth, restrict, prolongate, merge] table of the last V-cycle, seconds (:417-419)
int orc_vcycle_timing(void* h, double* t40, int* coarse_cg_iters)
{
    MatrixState& M = matrix_of((Sim*)h);
    for (int i = 0; i < 10; ++i)
        for (int j = 0; j < 4; ++j) t40[4 * i + j] = M.last_timing[i][j];
    if (coarse_cg_iters) *coarse_cg_iters = M.last_cg_iters;
    return 0;
}

} // extern "C"
