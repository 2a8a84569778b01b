// TEST INFRASTRUCTURE ONLY.  C entry points around the reference's OWN headers, compiled from where they lie under
// /root/reference/Lib (oracle/Makefile, target `ref`) against oracle/ref_shim/ (a minimal Eigen / Tick stand-in: neither library is in
// this image).  Used by tests/golden/make_ziran_golden.py to generate golden vectors and by tests/test_oracle_ziran_ref.py to
// check the oracle's restatement against the reference code directly.  Reference code exercised:
//   Lib/Ziran/Math/MathTools.h:15-25            int_floor
//   Lib/Ziran/Math/Splines/BSplines.h:10-29,55-81   baseNode<2>, computeBSplineWeights (quadratic)
//   Lib/Ziran/Math/Linear/Givens.h, ImplicitQRSVD.h:282-513   3x3 implicit-QR SVD with its sign / ordering conventions
//   Lib/Ziran/Math/Linear/DenseExt.h:240-252    cofactorMatrix
//   Lib/Ziran/Math/Linear/EigenDecomposition.h:126-135   makePD (on the stand-in's Jacobi eigen-solver)
//   Lib/Ziran/Physics/ConstitutiveModel/{SvdBasedIsotropicHelper.h, HyperelasticConstitutiveModel.h, CorotatedIsotropic.h:64-230}
#include <Ziran/Math/MathTools.h>
#include <Ziran/Math/Splines/BSplines.h>
#include <Ziran/Math/Linear/ImplicitQRSVD.h>
#include <Ziran/Physics/ConstitutiveModel/HyperelasticConstitutiveModel.h>
#include <Ziran/Physics/ConstitutiveModel/CorotatedIsotropic.h>

using namespace ZIRAN;
typedef Matrix<double, 3, 3> TM;
typedef Vector<double, 3> TV;

static TM load(const double* p)
{
    TM m;
    for (int q = 0; q < 9; ++q) m(q) = p[q]; // column-major like Eigen's default (Forward.h:10-13)
    return m;
}
static void store(const TM& m, double* p)
{
    for (int q = 0; q < 9; ++q) p[q] = m(q);
}

extern "C" {

int ziran_ref_int_floor(double x) { return MATH_TOOLS::int_floor(x); }

// x in index space (X * one_over_dx): base node and the three weights / weight derivatives of one axis
void ziran_ref_bspline2(long n, const double* x, int* base, double* w, double* dw)
{
    for (long i = 0; i < n; ++i) {
        TV wi, dwi;
        int b;
        computeBSplineWeights(x[i], b, wi, &dwi);
        base[i] = b;
        for (int t = 0; t < 3; ++t) { w[3 * i + t] = wi(t); dw[3 * i + t] = dwi(t); }
    }
}

void ziran_ref_svd3(long n, const double* F, double* U, double* sigma, double* V)
{
    for (long i = 0; i < n; ++i) {
        TM u, v;
        TV s;
        singularValueDecomposition(load(F + 9 * i), u, s, v);
        store(u, U + 9 * i); store(v, V + 9 * i);
        for (int t = 0; t < 3; ++t) sigma[3 * i + t] = s(t);
    }
}

// fixed-corotated model: psi, P, dP = dPdF : dF, and the dense dPdF (81 entries, column-major 9x9, index ij = i + 3 j)
void ziran_ref_corotated(long n, double mu, double lambda, int project, const double* F, const double* dF, double* psi, double* P,
    double* dP, double* dPdF)
{
    CorotatedIsotropic<double, 3> model;
    model.mu = mu;
    model.lambda = lambda;
    model.project = project != 0;
    for (long i = 0; i < n; ++i) {
        CorotatedIsotropicScratch<double, 3> s;
        model.updateScratch(load(F + 9 * i), s);
        if (psi) psi[i] = model.psi(s);
        if (P) {
            TM p;
            model.firstPiola(s, p);
            store(p, P + 9 * i);
        }
        if (dP) {
            TM dp;
            model.firstPiolaDifferential(s, load(dF + 9 * i), dp);
            store(dp, dP + 9 * i);
        }
        if (dPdF) {
            Eigen::Matrix<double, 9, 9> H;
            model.firstPiolaDerivative(s, H);
            for (int q = 0; q < 81; ++q) dPdF[81 * i + q] = H(q);
        }
    }
}

void ziran_ref_lame(double E, double nu, double* mu, double* lambda)
{
    CorotatedIsotropic<double, 3> model(E, nu);
    *mu = model.mu;
    *lambda = model.lambda;
}

double ziran_ref_clamp_small_magnitude(double x, double eps) { return MATH_TOOLS::clamp_small_magnitude(x, eps); }

} // extern "C"
