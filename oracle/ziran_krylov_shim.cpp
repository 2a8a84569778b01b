// TEST INFRASTRUCTURE ONLY - built into oracle/_ref/libziran_ref.so together with ziran_ref_shim.cpp.
// Runs the reference's OWN Krylov solver classes - ZIRAN::InexactConjugateGradient (Lib/Ziran/Math/Linear/InexactConjugateGradient.h)
// and ZIRAN::Minres (Lib/Ziran/Math/Linear/Minres.h), compiled where they lie - on an operator supplied through C callbacks
// (multiply / project / precondition), so that the oracle's restatement of the iteration (oracle_solver.inl: inexact_pcg,
// minres_solve) can be pinned to the reference's code on the oracle's own MPM systems: same operator, the reference's loop.
// The solvers are templates over the vector type TV; KVec below is the minimal eager vector they compile against.
#include <cmath>
#include <cstddef>
#include <vector>
#include <Ziran/Math/Linear/InexactConjugateGradient.h>
#include <Ziran/Math/Linear/Minres.h>

namespace {
struct KArr {
    std::vector<double> v;
    double sum() const { double s = 0; for (double x : v) s += x; return s; }
};
inline KArr operator*(const KArr& a, const KArr& b)
{
    KArr r; r.v.resize(a.v.size());
    for (size_t i = 0; i < a.v.size(); ++i) r.v[i] = a.v[i] * b.v[i];
    return r;
}
struct KVec {
    using Scalar = double;
    std::vector<double> v;
    size_t size() const { return v.size(); }
    void resizeLike(const KVec& o) { v.resize(o.v.size()); }
    void setZero() { for (double& x : v) x = 0; }
    double squaredNorm() const { double s = 0; for (double x : v) s += x * x; return s; }
    void swap(KVec& o) { v.swap(o.v); }
    KArr array() const { return KArr{v}; }
    KVec& operator+=(const KVec& o) { for (size_t i = 0; i < v.size(); ++i) v[i] += o.v[i]; return *this; }
    KVec& operator-=(const KVec& o) { for (size_t i = 0; i < v.size(); ++i) v[i] -= o.v[i]; return *this; }
    KVec& operator/=(double a) { for (double& x : v) x /= a; return *this; }
};
inline KVec operator-(const KVec& a, const KVec& b) { KVec r = a; r -= b; return r; }
inline KVec operator+(const KVec& a, const KVec& b) { KVec r = a; r += b; return r; }
inline KVec operator*(const KVec& a, double s) { KVec r = a; for (double& x : r.v) x *= s; return r; }
inline KVec operator*(double s, const KVec& a) { return a * s; }
inline KVec operator/(const KVec& a, double s) { KVec r = a; for (double& x : r.v) x /= s; return r; }

typedef void (*zr_apply_fn)(void* user, const double* in, double* out, long n);
typedef void (*zr_project_fn)(void* user, double* v, long n);
struct KOp {
    void* user; zr_apply_fn mul; zr_project_fn proj; zr_apply_fn prec;
    void multiply(const KVec& x, KVec& b) const { b.v.resize(x.v.size()); mul(user, x.v.data(), b.v.data(), (long)x.v.size()); }
    void project(KVec& x) const { if (proj) proj(user, x.v.data(), (long)x.v.size()); }
    void precondition(const KVec& in, KVec& out) const
    {
        out.v.resize(in.v.size());
        if (prec) prec(user, in.v.data(), out.v.data(), (long)in.v.size());
        else out.v = in.v;
    }
};
} // namespace

extern "C" {
// InexactConjugateGradient<double, KOp, KVec>(max_iterations).setTolerance(tolerance); returns solve(A, x, b)
int zr_inexact_cg(void* user, zr_apply_fn mul, zr_project_fn proj, zr_apply_fn prec, long n, double* x, const double* b, int max_iterations,
    double tolerance)
{
    KOp A{user, mul, proj, prec};
    KVec xv, bv; xv.v.assign(x, x + n); bv.v.assign(b, b + n);
    ZIRAN::InexactConjugateGradient<double, KOp, KVec> cg(max_iterations);
    cg.setTolerance(tolerance);
    const int it = cg.solve(A, xv, bv, false);
    for (long i = 0; i < n; ++i) x[i] = xv.v[i];
    return it;
}
// Minres<double, KOp, KVec>(max_iterations) with setTolerance(tolerance), setRelativeTolerance(relative_tolerance)
int zr_minres(void* user, zr_apply_fn mul, zr_project_fn proj, zr_apply_fn prec, long n, double* x, const double* b, int max_iterations,
    double tolerance, double relative_tolerance)
{
    KOp A{user, mul, proj, prec};
    KVec xv, bv; xv.v.assign(x, x + n); bv.v.assign(b, b + n);
    ZIRAN::Minres<double, KOp, KVec> mr(max_iterations);
    mr.setTolerance(tolerance);
    mr.setRelativeTolerance(relative_tolerance);
    const int it = mr.solve(A, xv, bv, false);
    for (long i = 0; i < n; ++i) x[i] = xv.v[i];
    return it;
}
}
