// TEST INFRASTRUCTURE ONLY — included by hot_oracle.cpp.  Restates the elastic force model of the hot path:
// a9 evalInterpolantAndGradient, a10 FBasedMpmForceHelper, a11 CorotatedIsotropic, a12 force rasterisation,
// a13 matrix-free Hessian apply, a14 objective (residual / energy / BC projection / CN tolerance).
//
// SVD note: the reference uses an implicit-shift QR SVD (Lib/Ziran/Math/Linear/ImplicitQRSVD.h:354-533) and Eigen's
// SelfAdjointEigenSolver for makePD (EigenDecomposition.h:126-135); both live in / depend on code that cannot be
// built here.  They are restated with one-sided / two-sided Jacobi iterations and the reference's sign
// convention (U, V rotations, sigma0 >= sigma1 >= |sigma2|, sign on sigma2).  Everything downstream (psi, P, dP,
// dPdF) is invariant to the remaining freedom; tests pin these against numpy.linalg (tests/test_oracle_force.py).

namespace {

// ---- 3x3 SVD: F = U diag(sigma) V^T ---------------------------------------------------------------------
void svd3(const double* F, double* U, double* sig, double* V)
{
    double A[9], W[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    std::memcpy(A, F, sizeof A);
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double* ap = A + 3 * p; double* aq = A + 3 * q;
                double alpha = ap[0] * ap[0] + ap[1] * ap[1] + ap[2] * ap[2];
                double beta = aq[0] * aq[0] + aq[1] * aq[1] + aq[2] * aq[2];
                double gamma = ap[0] * aq[0] + ap[1] * aq[1] + ap[2] * aq[2];
                if (gamma == 0 || std::fabs(gamma) <= 1e-17 * std::sqrt(alpha * beta)) continue;
                off = std::max(off, std::fabs(gamma) / std::sqrt(alpha * beta));
                double zeta = (beta - alpha) / (2 * gamma);
                double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1 + zeta * zeta));
                double c = 1 / std::sqrt(1 + t * t), s = c * t;
                for (int r = 0; r < 3; ++r) {
                    double x = ap[r], y = aq[r];
                    ap[r] = c * x - s * y; aq[r] = s * x + c * y;
                    double wx = W[3 * p + r], wy = W[3 * q + r];
                    W[3 * p + r] = c * wx - s * wy; W[3 * q + r] = s * wx + c * wy;
                }
            }
        if (off < 1e-16) break;
    }
    double n[3];
    int ord[3] = {0, 1, 2};
    for (int c = 0; c < 3; ++c) n[c] = std::sqrt(A[3 * c] * A[3 * c] + A[3 * c + 1] * A[3 * c + 1] + A[3 * c + 2] * A[3 * c + 2]);
    std::sort(ord, ord + 3, [&](int a, int b) { return n[a] > n[b]; });
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) V[3 * c + r] = W[3 * ord[c] + r];
    if (det3(V) < 0)
        for (int r = 0; r < 3; ++r) V[6 + r] = -V[6 + r];
    // U columns: normalised F v_c for the two largest, third by cross product (det U = +1), sigma2 signed
    double FV[9];
    mat_mul(F, V, FV);
    for (int c = 0; c < 2; ++c) {
        double l = std::sqrt(FV[3 * c] * FV[3 * c] + FV[3 * c + 1] * FV[3 * c + 1] + FV[3 * c + 2] * FV[3 * c + 2]);
        sig[c] = l;
        for (int r = 0; r < 3; ++r) U[3 * c + r] = l > 0 ? FV[3 * c + r] / l : (r == c ? 1.0 : 0.0);
    }
    // re-orthogonalise column 1 against column 0 (only matters for nearly rank-deficient F)
    double d01 = U[0] * U[3] + U[1] * U[4] + U[2] * U[5];
    for (int r = 0; r < 3; ++r) U[3 + r] -= d01 * U[r];
    double l1 = std::sqrt(U[3] * U[3] + U[4] * U[4] + U[5] * U[5]);
    for (int r = 0; r < 3; ++r) U[3 + r] /= l1;
    U[6] = U[1] * U[5] - U[2] * U[4];
    U[7] = U[2] * U[3] - U[0] * U[5];
    U[8] = U[0] * U[4] - U[1] * U[3];
    sig[2] = U[6] * FV[6] + U[7] * FV[7] + U[8] * FV[8];
}

// symmetric eigen-decomposition (two-sided Jacobi), S = Q diag(l) Q^T, n = 2 or 3, column-major
template <int n>
void sym_eig(const double* S, double* Q, double* l)
{
    double A[n * n];
    std::memcpy(A, S, sizeof A);
    for (int i = 0; i < n * n; ++i) Q[i] = (i % (n + 1) == 0) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 50; ++sweep) {
        double off = 0, diag = 0;
        for (int p = 0; p < n; ++p)
            for (int q = 0; q < n; ++q) (p == q ? diag : off) += A[p + n * q] * A[p + n * q];
        if (off <= 1e-34 * diag || off == 0) break;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                double apq = A[p + n * q];
                if (apq == 0) continue;
                double theta = (A[q + n * q] - A[p + n * p]) / (2 * apq);
                double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(1 + theta * theta));
                double c = 1 / std::sqrt(1 + t * t), s = c * t;
                for (int k = 0; k < n; ++k) { // A <- A J
                    double x = A[k + n * p], y = A[k + n * q];
                    A[k + n * p] = c * x - s * y; A[k + n * q] = s * x + c * y;
                }
                for (int k = 0; k < n; ++k) { // A <- J^T A
                    double x = A[p + n * k], y = A[q + n * k];
                    A[p + n * k] = c * x - s * y; A[q + n * k] = s * x + c * y;
                }
                for (int k = 0; k < n; ++k) {
                    double x = Q[k + n * p], y = Q[k + n * q];
                    Q[k + n * p] = c * x - s * y; Q[k + n * q] = s * x + c * y;
                }
            }
    }
    for (int i = 0; i < n; ++i) l[i] = A[i + n * i];
}

// makePD, Lib/Ziran/Math/Linear/EigenDecomposition.h:126-135: clamp negative eigenvalues to zero, recompose
template <int n>
void make_pd(double* S)
{
    double Q[n * n], l[n];
    sym_eig<n>(S, Q, l);
    for (int i = 0; i < n; ++i)
        if (l[i] < 0.0) l[i] = 0.0;
    for (int c = 0; c < n; ++c)
        for (int r = 0; r < n; ++r) {
            double v = 0;
            for (int k = 0; k < n; ++k) v += Q[r + n * k] * l[k] * Q[c + n * k];
            S[r + n * c] = v;
        }
}

// MathTools.h:163-175
inline double clamp_small_magnitude(double x, double eps)
{
    if (x < -eps) return x;
    if (x < 0) return -eps;
    if (x < eps) return eps;
    return x;
}

// DenseExt.h:240-252 cofactorMatrix (= J F^-T)
inline void cofactor3(const double* F, double* A)
{
    A[0] = F[4] * F[8] - F[7] * F[5]; A[3] = F[7] * F[2] - F[1] * F[8]; A[6] = F[1] * F[5] - F[4] * F[2];
    A[1] = F[6] * F[5] - F[3] * F[8]; A[4] = F[0] * F[8] - F[6] * F[2]; A[7] = F[3] * F[2] - F[0] * F[5];
    A[2] = F[3] * F[7] - F[6] * F[4]; A[5] = F[6] * F[1] - F[0] * F[7]; A[8] = F[0] * F[4] - F[3] * F[1];
}

// CorotatedIsotropic<T,3>::Scratch + SvdBasedIsotropicHelper<T,3>
// 0 = CorotatedIsotropic (the reference's model), 1 = neo-Hookean (extension, see include/hot_b200.h: hot_set_constitutive_model).
// Process-wide in the oracle (test infrastructure: one model per test), set by orc_set_constitutive_model.
static int g_constitutive_model = 0;

struct Scratch {
    int model = 0;
    double J, F[9], U[9], V[9], R[9], JFinvT[9], sigma[3];
    double psi0, psi1, psi2, psi00, psi11, psi22, psi01, psi02, psi12, m01, p01, m02, p02, m12, p12;
    double Aij[9], B01[4], B12[4], B20[4];
};

// CorotatedIsotropic.h:110-144
void update_scratch(const double* F, double mu, double lambda, bool project, Scratch& s)
{
    std::memcpy(s.F, F, sizeof s.F);
    svd3(s.F, s.U, s.sigma, s.V);
    mat_mul_bt(s.U, s.V, s.R);
    cofactor3(s.F, s.JFinvT);
    s.J = s.sigma[0] * s.sigma[1] * s.sigma[2];
    s.model = g_constitutive_model;
    const double eps = 1e-6;
    if (s.model == 1) {
        // neo-Hookean in the SvdBasedIsotropicHelper framework: psi_i = mu s_i + c / s_i with c = lambda log J - mu
        const double c = lambda * std::log(s.J) - mu, i0 = 1 / s.sigma[0], i1 = 1 / s.sigma[1], i2 = 1 / s.sigma[2];
        s.psi0 = mu * s.sigma[0] + c * i0; s.psi1 = mu * s.sigma[1] + c * i1; s.psi2 = mu * s.sigma[2] + c * i2;
        s.psi00 = mu + (lambda - c) * i0 * i0; s.psi11 = mu + (lambda - c) * i1 * i1; s.psi22 = mu + (lambda - c) * i2 * i2;
        s.psi01 = lambda * i0 * i1; s.psi02 = lambda * i0 * i2; s.psi12 = lambda * i1 * i2;
        s.m01 = mu - c * i0 * i1; s.m02 = mu - c * i0 * i2; s.m12 = mu - c * i1 * i2;
    }
    const double _2mu = mu * 2, _lambda = lambda * (s.J - 1);
    const double Sprod[3] = {s.sigma[1] * s.sigma[2], s.sigma[0] * s.sigma[2], s.sigma[0] * s.sigma[1]};
    if (s.model == 0) {
    s.psi0 = _2mu * (s.sigma[0] - 1) + _lambda * Sprod[0];
    s.psi1 = _2mu * (s.sigma[1] - 1) + _lambda * Sprod[1];
    s.psi2 = _2mu * (s.sigma[2] - 1) + _lambda * Sprod[2];
    s.psi00 = _2mu + lambda * Sprod[0] * Sprod[0];
    s.psi11 = _2mu + lambda * Sprod[1] * Sprod[1];
    s.psi22 = _2mu + lambda * Sprod[2] * Sprod[2];
    s.psi01 = _lambda * s.sigma[2] + lambda * Sprod[0] * Sprod[1];
    s.psi02 = _lambda * s.sigma[1] + lambda * Sprod[0] * Sprod[2];
    s.psi12 = _lambda * s.sigma[0] + lambda * Sprod[1] * Sprod[2];
    s.m01 = _2mu - _lambda * s.sigma[2];
    s.m02 = _2mu - _lambda * s.sigma[1];
    s.m12 = _2mu - _lambda * s.sigma[0];
    }
    s.p01 = (s.psi0 + s.psi1) / clamp_small_magnitude(s.sigma[0] + s.sigma[1], eps);
    s.p02 = (s.psi0 + s.psi2) / clamp_small_magnitude(s.sigma[0] + s.sigma[2], eps);
    s.p12 = (s.psi1 + s.psi2) / clamp_small_magnitude(s.sigma[1] + s.sigma[2], eps);
    // buildMatrixBlock, SvdBasedIsotropicHelper.h:223-237
    s.Aij[0] = s.psi00; s.Aij[4] = s.psi11; s.Aij[8] = s.psi22;
    s.Aij[3] = s.Aij[1] = s.psi01; s.Aij[6] = s.Aij[2] = s.psi02; s.Aij[7] = s.Aij[5] = s.psi12;
    s.B01[0] = s.B01[3] = (s.m01 + s.p01) * 0.5; s.B01[1] = s.B01[2] = (s.m01 - s.p01) * 0.5;
    s.B12[0] = s.B12[3] = (s.m12 + s.p12) * 0.5; s.B12[1] = s.B12[2] = (s.m12 - s.p12) * 0.5;
    s.B20[0] = s.B20[3] = (s.m02 + s.p02) * 0.5; s.B20[1] = s.B20[2] = (s.m02 - s.p02) * 0.5;
    if (project) { // projectABBlock :239-247
        make_pd<3>(s.Aij);
        make_pd<2>(s.B01);
        make_pd<2>(s.B12);
        make_pd<2>(s.B20);
    }
}

// CorotatedIsotropic.h:151-155
inline double psi_of(const Scratch& s, double mu, double lambda)
{
    if (s.model == 1) {
        const double lj = std::log(s.J);
        return 0.5 * mu * (s.sigma[0] * s.sigma[0] + s.sigma[1] * s.sigma[1] + s.sigma[2] * s.sigma[2] - 3) - mu * lj + 0.5 * lambda * lj * lj;
    }
    double n2 = 0;
    for (int q = 0; q < 9; ++q) n2 += (s.F[q] - s.R[q]) * (s.F[q] - s.R[q]);
    double Jm1 = s.J - 1;
    return mu * n2 + 0.5 * lambda * Jm1 * Jm1;
}
// :157-160
inline void first_piola(const Scratch& s, double mu, double lambda, double* P)
{
    if (s.model == 1) { // P = U diag(psi_i) V^T
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) P[r + 3 * c] = s.U[r] * s.psi0 * s.V[c] + s.U[r + 3] * s.psi1 * s.V[c + 3] + s.U[r + 6] * s.psi2 * s.V[c + 6];
        return;
    }
    for (int q = 0; q < 9; ++q) P[q] = 2 * mu * (s.F[q] - s.R[q]) + lambda * (s.J - 1) * s.JFinvT[q];
}
// :162-171 with dPdFOfSigmaContractProjected (SvdBasedIsotropicHelper.h:271-282).  After buildMatrixBlock the
// unprojected contraction (:256-268) is the same formula with the unprojected blocks.
inline void first_piola_differential(const Scratch& s, const double* dF, double* dP)
{
    double D[9], K[9], t[9];
    mat_mul_at(s.U, dF, t);
    mat_mul(t, s.V, D);
#define M_(X, r, c) X[(r) + 3 * (c)]
#define B_(X, r, c) X[(r) + 2 * (c)]
    M_(K, 0, 0) = M_(s.Aij, 0, 0) * M_(D, 0, 0) + M_(s.Aij, 0, 1) * M_(D, 1, 1) + M_(s.Aij, 0, 2) * M_(D, 2, 2);
    M_(K, 1, 1) = M_(s.Aij, 1, 0) * M_(D, 0, 0) + M_(s.Aij, 1, 1) * M_(D, 1, 1) + M_(s.Aij, 1, 2) * M_(D, 2, 2);
    M_(K, 2, 2) = M_(s.Aij, 2, 0) * M_(D, 0, 0) + M_(s.Aij, 2, 1) * M_(D, 1, 1) + M_(s.Aij, 2, 2) * M_(D, 2, 2);
    M_(K, 0, 1) = B_(s.B01, 0, 0) * M_(D, 0, 1) + B_(s.B01, 0, 1) * M_(D, 1, 0);
    M_(K, 1, 0) = B_(s.B01, 1, 0) * M_(D, 0, 1) + B_(s.B01, 1, 1) * M_(D, 1, 0);
    M_(K, 0, 2) = B_(s.B20, 0, 0) * M_(D, 0, 2) + B_(s.B20, 0, 1) * M_(D, 2, 0);
    M_(K, 2, 0) = B_(s.B20, 1, 0) * M_(D, 0, 2) + B_(s.B20, 1, 1) * M_(D, 2, 0);
    M_(K, 1, 2) = B_(s.B12, 0, 0) * M_(D, 1, 2) + B_(s.B12, 0, 1) * M_(D, 2, 1);
    M_(K, 2, 1) = B_(s.B12, 1, 0) * M_(D, 1, 2) + B_(s.B12, 1, 1) * M_(D, 2, 1);
    mat_mul(s.U, K, t);
    mat_mul_bt(t, s.V, dP);
}
// dense 9x9 dPdF(ij, rs), ij = i + 3 j: CorotatedIsotropic.h:198-227 (projected branch; with project == false the
// blocks hold the unprojected values, which reproduces the other branch term by term)
void first_piola_derivative(const Scratch& ss, double* dPdF /* 81, column-major 9x9 */)
{
    const double *U = ss.U, *V = ss.V, *A = ss.Aij, *B01 = ss.B01, *B12 = ss.B12, *B20 = ss.B20;
    for (int ij = 0; ij < 9; ++ij) {
        int j = ij / 3, i = ij - j * 3;
        for (int rs = 0; rs <= ij; ++rs) {
            int s = rs / 3, r = rs - s * 3;
            double v = M_(A, 0, 0) * M_(U, i, 0) * M_(V, j, 0) * M_(U, r, 0) * M_(V, s, 0) + M_(A, 0, 1) * M_(U, i, 0) * M_(V, j, 0) * M_(U, r, 1) * M_(V, s, 1)
                + M_(A, 0, 2) * M_(U, i, 0) * M_(V, j, 0) * M_(U, r, 2) * M_(V, s, 2) + M_(A, 0, 1) * M_(U, i, 1) * M_(V, j, 1) * M_(U, r, 0) * M_(V, s, 0)
                + M_(A, 1, 1) * M_(U, i, 1) * M_(V, j, 1) * M_(U, r, 1) * M_(V, s, 1) + M_(A, 1, 2) * M_(U, i, 1) * M_(V, j, 1) * M_(U, r, 2) * M_(V, s, 2)
                + M_(A, 0, 2) * M_(U, i, 2) * M_(V, j, 2) * M_(U, r, 0) * M_(V, s, 0) + M_(A, 1, 2) * M_(U, i, 2) * M_(V, j, 2) * M_(U, r, 1) * M_(V, s, 1)
                + M_(A, 2, 2) * M_(U, i, 2) * M_(V, j, 2) * M_(U, r, 2) * M_(V, s, 2)
                + B_(B01, 0, 0) * M_(U, i, 0) * M_(V, j, 1) * M_(U, r, 0) * M_(V, s, 1) + B_(B01, 0, 1) * M_(U, i, 0) * M_(V, j, 1) * M_(U, r, 1) * M_(V, s, 0)
                + B_(B01, 1, 0) * M_(U, i, 1) * M_(V, j, 0) * M_(U, r, 0) * M_(V, s, 1) + B_(B01, 1, 1) * M_(U, i, 1) * M_(V, j, 0) * M_(U, r, 1) * M_(V, s, 0)
                + B_(B12, 0, 0) * M_(U, i, 1) * M_(V, j, 2) * M_(U, r, 1) * M_(V, s, 2) + B_(B12, 0, 1) * M_(U, i, 1) * M_(V, j, 2) * M_(U, r, 2) * M_(V, s, 1)
                + B_(B12, 1, 0) * M_(U, i, 2) * M_(V, j, 1) * M_(U, r, 1) * M_(V, s, 2) + B_(B12, 1, 1) * M_(U, i, 2) * M_(V, j, 1) * M_(U, r, 2) * M_(V, s, 1)
                + B_(B20, 1, 1) * M_(U, i, 0) * M_(V, j, 2) * M_(U, r, 0) * M_(V, s, 2) + B_(B20, 1, 0) * M_(U, i, 0) * M_(V, j, 2) * M_(U, r, 2) * M_(V, s, 0)
                + B_(B20, 0, 1) * M_(U, i, 2) * M_(V, j, 0) * M_(U, r, 0) * M_(V, s, 2) + B_(B20, 0, 0) * M_(U, i, 2) * M_(V, j, 0) * M_(U, r, 2) * M_(V, s, 0);
            dPdF[ij + 9 * rs] = dPdF[rs + 9 * ij] = v;
        }
    }
}
#undef M_
#undef B_

// per-particle force state of the oracle (original particle order)
struct ForceState {
    bool project = true; // CorotatedIsotropic::project (default true, :60)
    std::vector<double> Fn, stress, vp, gradV;
    std::vector<Scratch> scratch;
    // BC table (a8 output, host-evaluated): CollisionNode{node_id,P,R,Rinv,shouldRotate} (CollisionObject.h:16-45)
    int bc_mode = 0; // HOTSettings::boundaryType && systemBCProject: 0 = project with P, 1 = slip rotation mode
    std::vector<int> bc_node, bc_slip;
    std::vector<double> bc_P, bc_R, bc_Rinv;
    std::vector<double> nodeCNTol;
};

} // namespace

#define FS(s) (*(ForceState*)(s)->force_state)

namespace {

ForceState& force_of(Sim* s)
{
    if (!s->force_state) s->force_state = new ForceState;
    return FS(s);
}

// a9: MpmForceBase::evalInterpolantAndGradient, Lib/MPM/Force/MpmForceBase.cpp:213-248 (field given per DOF)
void eval_interpolant_and_gradient(Sim* s, const double* field, std::vector<double>& f_eval, std::vector<double>& grad_f)
{
    const long n = s->N;
    f_eval.assign(3 * n, 0.0);
    grad_f.assign(9 * n, 0.0);
#pragma omp parallel for
    for (long a = 0; a < (long)s->grid.size(); ++a) {
        GridState& g = s->grid[a];
        g.new_v[0] = g.new_v[1] = g.new_v[2] = 0;
        if (g.idx >= 0)
            for (int d = 0; d < 3; ++d) g.new_v[d] = field[3 * g.idx + d];
    }
    s->for_colored_groups([&](int grp) {
        for (int sidx = s->groups[grp].first; sidx <= s->groups[grp].second; ++sidx) {
            int i = s->order[sidx];
            Spline sp(&s->X[3 * i], s->dx);
            double* G = &grad_f[9 * i];
            double* f = &f_eval[3 * i];
            s->iterate_kernel(sp, s->base_offset[i], [&](const int*, double w, const double* dw, GridState& g) {
                for (int c = 0; c < 3; ++c)
                    for (int r = 0; r < 3; ++r) G[r + 3 * c] += g.new_v[r] * dw[c];
                for (int r = 0; r < 3; ++r) f[r] += g.new_v[r] * w;
            });
        }
    });
}

// a12: MpmForceBase::rasterizeForceToTVStack<false>, MpmForceBase.cpp:100-153 (fp == 0: no meshed forces here)
void rasterize_force(Sim* s, double scale, const std::vector<double>& stress, double* force)
{
#pragma omp parallel for
    for (long a = 0; a < (long)s->grid.size(); ++a) s->grid[a].new_v[0] = s->grid[a].new_v[1] = s->grid[a].new_v[2] = 0;
    s->for_colored_groups([&](int grp) {
        for (int sidx = s->groups[grp].first; sidx <= s->groups[grp].second; ++sidx) {
            int i = s->order[sidx];
            Spline sp(&s->X[3 * i], s->dx);
            const double* S = &stress[9 * i];
            s->iterate_kernel(sp, s->base_offset[i], [&](const int*, double, const double* dw, GridState& g) {
                for (int r = 0; r < 3; ++r) g.new_v[r] -= scale * (S[r] * dw[0] + S[r + 3] * dw[1] + S[r + 6] * dw[2]);
            });
        }
    });
    for (auto& g : s->grid)
        if (g.idx >= 0)
            for (int d = 0; d < 3; ++d) force[3 * g.idx + d] += g.new_v[d];
}

// objective.project (MultigridSimulation.h:104-125) and transformResidual (ImplicitSolver.h:117-125)
void bc_project(Sim* s, double* v)
{
    ForceState& f = force_of(s);
    for (size_t b = 0; b < f.bc_node.size(); ++b) {
        double* x = v + 3 * (size_t)f.bc_node[b];
        if (f.bc_mode == 1) {
            if (f.bc_slip[b]) x[0] = 0;
            else x[0] = x[1] = x[2] = 0;
        }
        else {
            const double* P = &f.bc_P[9 * b];
            double y[3];
            for (int r = 0; r < 3; ++r) y[r] = P[r] * x[0] + P[r + 3] * x[1] + P[r + 6] * x[2];
            x[0] = y[0]; x[1] = y[1]; x[2] = y[2];
        }
    }
}
void bc_rotate(Sim* s, double* v, bool inverse)
{
    ForceState& f = force_of(s);
    if (f.bc_mode != 1) return;
    for (size_t b = 0; b < f.bc_node.size(); ++b)
        if (f.bc_slip[b]) {
            double* x = v + 3 * (size_t)f.bc_node[b];
            const double* R = inverse ? &f.bc_Rinv[9 * b] : &f.bc_R[9 * b];
            double y[3];
            for (int r = 0; r < 3; ++r) y[r] = R[r] * x[0] + R[r + 3] * x[1] + R[r + 6] * x[2];
            x[0] = y[0]; x[1] = y[1]; x[2] = y[2];
        }
}

// ImplicitSolver.h:254-275 + Inertia.cpp:16-30 + MpmForceBase.cpp:349-369
double total_energy(Sim* s)
{
    ForceState& f = force_of(s);
    double e = 0;
#pragma omp parallel for reduction(+ : e)
    for (long i = 0; i < s->N; ++i) e += s->vol[i] * psi_of(f.scratch[i], s->mu[i], s->lambda[i]);
    double ke = 0, ge = 0;
    for (int i = 0; i < s->num_nodes; ++i) {
        const double* d = &s->dv[3 * (size_t)i];
        ke += (d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) * s->mass_matrix[i];
        ge += (s->gravity[0] * d[0] + s->gravity[1] * d[1] + s->gravity[2] * d[2]) * s->mass_matrix[i];
    }
    return e + ke / 2 - s->dt * ge;
}

} // namespace

extern "C" {

int orc_set_project(void* h, int project)
{
    force_of((Sim*)h).project = project != 0;
    return 0;
}

// a8 output crossing the boundary: the BC table + Newton initial guess for the collided nodes
// (MpmSimulationBase.cpp:1139-1184).  dv = gravity*dt on free nodes, dv_bc on BC nodes; vn = v.
int orc_set_bc(void* h, int mode, int n_bc, const int* node_id, const double* P, const double* R, const double* Rinv, const int* slip,
    const double* dv_bc)
{
    Sim* s = (Sim*)h;
    ForceState& f = force_of(s);
    f.bc_mode = mode;
    f.bc_node.assign(node_id, node_id + n_bc);
    f.bc_slip.assign(n_bc, 0);
    if (slip) f.bc_slip.assign(slip, slip + n_bc);
    f.bc_P.assign(9 * (size_t)n_bc, 0.0);
    if (P) f.bc_P.assign(P, P + 9 * (size_t)n_bc);
    f.bc_R.assign(9 * (size_t)n_bc, 0.0); f.bc_Rinv.assign(9 * (size_t)n_bc, 0.0);
    if (R) f.bc_R.assign(R, R + 9 * (size_t)n_bc);
    if (Rinv) f.bc_Rinv.assign(Rinv, Rinv + 9 * (size_t)n_bc);
    for (int i = 0; i < s->num_nodes; ++i)
        for (int d = 0; d < 3; ++d) s->dv[3 * (size_t)i + d] = s->gravity[d] * s->dt;
    for (int b = 0; b < n_bc; ++b)
        for (int d = 0; d < 3; ++d) s->dv[3 * (size_t)node_id[b] + d] = dv_bc ? dv_bc[3 * b + d] : 0.0;
    for (auto& g : s->grid)
        if (g.idx >= 0)
            for (int d = 0; d < 3; ++d) s->vn[3 * g.idx + d] = g.v[d];
    return 0;
}
int orc_get_dv(void* h, double* dv)
{
    Sim* s = (Sim*)h;
    std::copy(s->dv.begin(), s->dv.begin() + 3 * (size_t)s->num_nodes, dv);
    return 0;
}

// FBasedMpmForceHelper::backupStrain / restoreStrain, FBasedMpmForceHelper.cpp:25-44
int orc_backup_strain(void* h)
{
    Sim* s = (Sim*)h;
    force_of(s).Fn = s->F;
    return 0;
}
int orc_restore_strain(void* h)
{
    Sim* s = (Sim*)h;
    s->F = force_of(s).Fn;
    return 0;
}

// ImplicitSolverObjective::updateState (ImplicitSolver.h:237-252): moveNodes, updatePositionBasedState
// (MpmForceBase.cpp:308-328 -> computeVAndGradV, restoreStrain, evolveStrain, updateParticleImplicitState), Ek.
int orc_update_state(void* h, const double* dv, double* energy)
{
    Sim* s = (Sim*)h;
    ForceState& f = force_of(s);
    if (f.Fn.size() != s->F.size()) return fail(s, "orc_update_state: call orc_backup_strain first (startBackwardEuler)");
    const long n = s->N;
    if (dv && dv != s->dv.data()) s->dv.assign(dv, dv + 3 * (size_t)s->num_nodes); // moveNodes (no-op when dv aliases, :738-739)
    std::vector<double> field(3 * (size_t)s->num_nodes);
    for (size_t q = 0; q < field.size(); ++q) field[q] = s->vn[q] + s->dv[q];
    eval_interpolant_and_gradient(s, field.data(), f.vp, f.gradV);
    f.scratch.resize(n);
    f.stress.assign(9 * n, 0.0);
    const double dt = s->dt;
#pragma omp parallel for
    for (long i = 0; i < n; ++i) {
        // restoreStrain + evolveStrain: F = (I + dt gradV) Fn  (FBasedMpmForceHelper.cpp:100-114)
        double A[9];
        for (int q = 0; q < 9; ++q) A[q] = dt * f.gradV[9 * i + q];
        A[0] += 1; A[4] += 1; A[8] += 1;
        mat_mul(A, &f.Fn[9 * i], &s->F[9 * i]);
        // updateImplicitState: vPFnT = vol * P * Fn^T (:72-97)
        update_scratch(&s->F[9 * i], s->mu[i], s->lambda[i], f.project, f.scratch[i]);
        double P[9], t[9];
        first_piola(f.scratch[i], s->mu[i], s->lambda[i], P);
        mat_mul_bt(P, &f.Fn[9 * i], t);
        for (int q = 0; q < 9; ++q) f.stress[9 * i + q] = s->vol[i] * t[q];
    }
    if (energy) *energy = total_energy(s);
    return 0;
}

int orc_get_stress(void* h, double* vPFnT, double* F)
{
    Sim* s = (Sim*)h;
    ForceState& f = force_of(s);
    if (vPFnT) std::copy(f.stress.begin(), f.stress.end(), vPFnT);
    if (F) std::copy(s->F.begin(), s->F.end(), F);
    return 0;
}

// ImplicitSolverObjective::computeResidual, ImplicitSolver.h:128-155
int orc_compute_residual(void* h, double* r)
{
    Sim* s = (Sim*)h;
    ForceState& f = force_of(s);
    const int nn = s->num_nodes;
    for (int i = 0; i < nn; ++i)
        for (int d = 0; d < 3; ++d) r[3 * (size_t)i + d] = s->dt * s->gravity[d] * s->mass_matrix[i];
    rasterize_force(s, s->dt, f.stress, r);
    for (int i = 0; i < nn; ++i) // inertia->addScaledForces(dt): scale/dt = 1  (Inertia.cpp:33-41)
        for (int d = 0; d < 3; ++d) r[3 * (size_t)i + d] -= s->mass_matrix[i] * s->dv[3 * (size_t)i + d];
    bc_rotate(s, r, false);
    bc_project(s, r);
    return 0;
}

int orc_project(void* h, double* v)
{
    bc_project((Sim*)h, v);
    return 0;
}

// a13: ImplicitSolverObjective::multiply with matrix_free (ImplicitSolver.h:741-763):
// b = M x + dt^2 * sum_p [vol dP(grad_x Fn) Fn^T] grad_w   (MpmForceBase.cpp:261-306, FBasedMpmForceHelper.cpp:138-160,
// Inertia.cpp:45-53)
int orc_hessian_apply_mf(void* h, const double* x, double* b)
{
    Sim* s = (Sim*)h;
    ForceState& f = force_of(s);
    const long n = s->N;
    const int nn = s->num_nodes;
    if ((long)f.scratch.size() != n) return fail(s, "orc_hessian_apply_mf: call orc_update_state first");
    for (int i = 0; i < nn; ++i)
        for (int d = 0; d < 3; ++d) b[3 * (size_t)i + d] = s->mass_matrix[i] * x[3 * (size_t)i + d];
    std::vector<double> dvp, gradDv, dstress(9 * n);
    eval_interpolant_and_gradient(s, x, dvp, gradDv);
#pragma omp parallel for
    for (long i = 0; i < n; ++i) {
        double dF[9], dP[9], t[9];
        mat_mul(&gradDv[9 * i], &f.Fn[9 * i], dF);
        first_piola_differential(f.scratch[i], dF, dP);
        mat_mul_bt(dP, &f.Fn[9 * i], t);
        for (int q = 0; q < 9; ++q) dstress[9 * i + q] = s->vol[i] * t[q];
    }
    rasterize_force(s, -(s->dt * s->dt), dstress, b);
    return 0;
}

// a18: evaluatePerNodeCNTolerance (ImplicitSolver.h:667-696) with computePerNodeCNTolerance evaluating dPdF at F = I
// (FBasedMpmForceHelper.h:123-157)
int orc_eval_cn_tolerance(void* h, double eps, double dt, double* tol)
{
    Sim* s = (Sim*)h;
    ForceState& f = force_of(s);
    f.nodeCNTol.assign(s->num_nodes, 0.0);
    const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    s->for_colored_groups([&](int grp) {
        for (int sidx = s->groups[grp].first; sidx <= s->groups[grp].second; ++sidx) {
            int i = s->order[sidx];
            Scratch sc;
            double H[81], nrm = 0;
            update_scratch(I, s->mu[i], s->lambda[i], f.project, sc);
            first_piola_derivative(sc, H);
            for (int q = 0; q < 81; ++q) nrm += H[q] * H[q];
            nrm = std::sqrt(nrm);
            Spline sp(&s->X[3 * i], s->dx);
            s->iterate_kernel(sp, s->base_offset[i], [&](const int*, double w, const double*, GridState& g) {
                if (g.idx < 0) return;
                f.nodeCNTol[g.idx] += w * s->mass[i] * nrm;
            });
        }
    });
    for (int i = 0; i < s->num_nodes; ++i) f.nodeCNTol[i] *= (eps * 24 * s->dx * s->dx * dt) / s->mass_matrix[i];
    if (tol) std::copy(f.nodeCNTol.begin(), f.nodeCNTol.end(), tol);
    return 0;
}

// MpmSimulationBase::applyPlasticity (MpmSimulationBase.cpp:1044-1064): VonMisesFixedCorotated::projectStrain
// (PlasticityApplier.cpp:94-131) or SnowPlasticity::projectStrain (:16-50) on every particle's F
int orc_set_constitutive_model(void* /*h*/, int model)
{
    g_constitutive_model = model;
    return 0;
}
int orc_set_plasticity(void* h, int model, const double* params)
{
    Sim* s = (Sim*)h;
    s->plastic_model = model;
    for (int k = 0; k < 5; ++k) s->plastic_param[k] = (model && params) ? params[k] : 0.0;
    return 0;
}
int orc_apply_plasticity(void* h)
{
    Sim* s = (Sim*)h;
    if (s->plastic_model == 0) return 0;
    const double* q = s->plastic_param;
#pragma omp parallel for
    for (long i = 0; i < s->N; ++i) {
        double* F = &s->F[9 * i];
        double U[9], V[9], sig[3], sn[3];
        svd3(F, U, sig, V);
        if (s->plastic_model == 1) {
            const double mu = s->mu[i], lambda = s->lambda[i];
            for (int d = 0; d < 3; ++d) sig[d] = std::max(1e-4, sig[d]);
            const double J = sig[0] * sig[1] * sig[2];
            double tau[3], tr = 0;
            for (int d = 0; d < 3; ++d) {
                tau[d] = 2 * mu * (sig[d] - 1) * sig[d] + lambda * (J - 1) * J;
                tr += tau[d];
            }
            double sd[3], n2 = 0;
            for (int d = 0; d < 3; ++d) {
                sd[d] = tau[d] - tr / 3;
                n2 += sd[d] * sd[d];
            }
            const double s_norm = std::sqrt(n2), scaled_tauy = std::sqrt(2.0 / 3.0) * q[0];
            if (s_norm - scaled_tauy <= 0) continue;
            const double alpha = scaled_tauy / s_norm;
            for (int d = 0; d < 3; ++d) {
                const double tau_new = alpha * sd[d] + tr / 3;
                const double b2m4ac = mu * mu - 2 * mu * (lambda * (J - 1) * J - tau_new);
                sn[d] = (mu + std::sqrt(b2m4ac)) / (2 * mu);
            }
        }
        else if (s->plastic_model == 3) {
            // Drucker-Prager extension (Klar et al. 2016, Hencky-strain return mapping); not in the reference, see hot_b200.h
            const double mu = s->mu[i], lambda = s->lambda[i], sphi = std::sin(q[0] * 0.017453292519943295),
                         alpha = std::sqrt(2.0 / 3.0) * 2.0 * sphi / (3.0 - sphi), coh = q[1];
            double eps[3], tr = 0;
            for (int d = 0; d < 3; ++d) {
                eps[d] = std::log(std::max(sig[d], 1e-6)) - coh;
                tr += eps[d];
            }
            double dev[3], n2 = 0;
            for (int d = 0; d < 3; ++d) {
                dev[d] = eps[d] - tr / 3;
                n2 += dev[d] * dev[d];
            }
            const double nrm = std::sqrt(n2);
            if (tr >= 0) {
                for (int d = 0; d < 3; ++d) sn[d] = std::exp(coh);
            }
            else if (nrm == 0) continue; // hydrostatic compression: inside the cone
            else {
                const double dgamma = nrm + (3 * lambda + 2 * mu) / (2 * mu) * tr * alpha;
                if (dgamma <= 0) continue;
                for (int d = 0; d < 3; ++d) sn[d] = std::exp(eps[d] - dgamma * dev[d] / nrm + coh);
            }
        }
        else {
            double Fe_det = 1;
            for (int d = 0; d < 3; ++d) {
                sn[d] = std::max(std::min(sig[d], 1 + q[2]), 1 - q[1]);
                Fe_det *= sn[d];
            }
            double Jnew = s->Jp[i] * det3(F) / Fe_det;
            if (!(Jnew <= q[4])) Jnew = q[4];
            if (!(Jnew >= q[3])) Jnew = q[3];
            const double hd = std::exp(q[0] * (s->Jp[i] - Jnew));
            s->mu[i] *= hd;
            s->lambda[i] *= hd;
            s->Jp[i] = Jnew;
        }
        double Fe[9];
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) Fe[r + 3 * c] = U[r] * sn[0] * V[c] + U[r + 3] * sn[1] * V[c + 3] + U[r + 6] * sn[2] * V[c + 6];
        std::copy(Fe, Fe + 9, F);
    }
    return 0;
}
int orc_get_plastic_state(void* h, double* Jp, double* mu, double* lambda)
{
    Sim* s = (Sim*)h;
    if (Jp) std::copy(s->Jp.begin(), s->Jp.end(), Jp);
    if (mu) std::copy(s->mu.begin(), s->mu.end(), mu);
    if (lambda) std::copy(s->lambda.begin(), s->lambda.end(), lambda);
    return 0;
}

// single-particle constitutive evaluation for unit tests against numpy
int orc_constitutive(const double* F, double mu, double lambda, int project, double* psi, double* P, double* dPdF81,
    const double* dF, double* dP, double* U, double* sigma, double* V)
{
    Scratch sc;
    update_scratch(F, mu, lambda, project != 0, sc);
    if (psi) *psi = psi_of(sc, mu, lambda);
    if (P) first_piola(sc, mu, lambda, P);
    if (dPdF81) first_piola_derivative(sc, dPdF81);
    if (dF && dP) first_piola_differential(sc, dF, dP);
    if (U) std::copy(sc.U, sc.U + 9, U);
    if (sigma) std::copy(sc.sigma, sc.sigma + 3, sigma);
    if (V) std::copy(sc.V, sc.V + 9, V);
    return 0;
}

} // extern "C"
