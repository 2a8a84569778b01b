// TEST INFRASTRUCTURE ONLY — not part of the product path.
//
// The REFERENCE'S OWN plasticity return mappings (row f2): Lib/Ziran/Physics/PlasticityApplier.cpp is compiled where it lies - the whole file,
// with its explicit instantiations - against the Eigen stand-in and inert stand-ins for DataManager / DisjointRanges / Rotation
// (oracle/ref_shim/Ziran/...), and
//   SnowPlasticity<double>::projectStrain            PlasticityApplier.cpp:16-50   (singular values clamped to [1 - theta_c, 1 + theta_s], Jp, hardening)
//   VonMisesFixedCorotated<double,3>::projectStrain  PlasticityApplier.cpp:94-131  (return mapping of the fixed-corotated Kirchhoff stress)
// are called per particle with the reference's CorotatedIsotropic as the constitutive model - which is all PlasticityApplier::applyPlasticity
// (PlasticityApplier.h:38-50) and MpmSimulationBase::applyPlasticity (Lib/MPM/MpmSimulationBase.cpp:1044-1064) do.
// Built by oracle/Makefile into oracle/_ref/libplasticity_ref.so; tests/golden/make_plasticity_golden.py, tests/test_oracle_plasticity_ref.py.
#include <Ziran/CS/Util/Debug.h>
#include <Ziran/Physics/PlasticityApplier.cpp>

using namespace ZIRAN;
typedef Matrix<double, 3, 3> TM;

extern "C" {

// F: n x 9 column-major, updated in place; mu / lambda / Jp per particle, updated in place; q = {psi, theta_c, theta_s, min_Jp, max_Jp}
void zr_plasticity_snow(long n, double* F, double* mu, double* lambda, double* Jp, const double* q)
{
    for (long i = 0; i < n; ++i) {
        SnowPlasticity<double> p(q[0], q[1], q[2], q[3], q[4]);
        p.Jp = Jp[i];
        CorotatedIsotropic<double, 3> c;
        c.mu = mu[i];
        c.lambda = lambda[i];
        TM s;
        for (int k = 0; k < 9; ++k) s(k) = F[9 * i + k];
        p.projectStrain(c, s);
        for (int k = 0; k < 9; ++k) F[9 * i + k] = s(k);
        mu[i] = c.mu;
        lambda[i] = c.lambda;
        Jp[i] = p.Jp;
    }
}

void zr_plasticity_von_mises(long n, double* F, const double* mu, const double* lambda, double yield_stress, int* projected)
{
    for (long i = 0; i < n; ++i) {
        VonMisesFixedCorotated<double, 3> p(yield_stress);
        CorotatedIsotropic<double, 3> c;
        c.mu = mu[i];
        c.lambda = lambda[i];
        TM s;
        for (int k = 0; k < 9; ++k) s(k) = F[9 * i + k];
        projected[i] = p.projectStrain(c, s) ? 1 : 0;
        for (int k = 0; k < 9; ++k) F[9 * i + k] = s(k);
    }
}

} // extern "C"
