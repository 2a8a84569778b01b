// hot_b200 host side (C++17, header-only): the reference's operator surface for the hot path, re-expressed over the
// C ABI of hot_b200.h.  The reference is in-process C++ (Eigen + TBB), so a drop-in keeps ITS names and call order:
//
//   MpmSimulationBase<T,dim> virtuals  (Lib/MPM/MpmSimulationBase.h:138-220)        -> hot_b200::MpmSimulationB200
//   ImplicitSolverObjective concept    (Projects/multigrid/ImplicitSolver.h:41-47,128-333,741-763) -> ::ImplicitSolverObjectiveB200
//   smoothFunc pointers / -smoother codes (MultigridPreconditioner.h:68-73,496-521)  -> hot_b200::{jacobi,optimal_jacobi,cg,gs}_smooth
//   HOTSettings statics + FLAGS        (Configurations.h:18-42, main.cpp:40-84)      -> hot_b200::HOTSettings, parseFlags
//   AnalyticCollisionObject / CollisionNode (CollisionObject.h:16-45, .cpp:108-149,384-452) -> host-evaluated a8
//
// TVStack here is a std::vector<double> of n x (x,y,z) - the memory layout of Eigen::Matrix<T,3,Dynamic>; 3x3 matrices are
// column-major 9-arrays like Eigen.  With Eigen available, `Eigen::Map<TVStack>(v.data(), 3, n)` views these buffers with
// no copy (INTEGRATION.md shows the Ziran-side glue).  Errors surface as std::runtime_error like ZIRAN_ASSERT
// (Lib/Ziran/CS/Util/Debug.h:19-42).  There is no CPU fallback: construction fails without a usable GPU.
#pragma once
#include "hot_b200.h"
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <istream>
#include <limits>
#include <memory>
#include <ostream>
#include <stdexcept>
#include <string>
#include <vector>

namespace hot_b200 {

using TVStack = std::vector<double>;
using TV = std::array<double, 3>;
using TM = std::array<double, 9>; // column-major

struct HotError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// Projects/multigrid/Configurations.h:18-42 - same names, same defaults
namespace HOTSettings {
inline double cneps = 1e-5;
inline bool useAdaptiveHessian = false, useCN = false, matrixFree = false, project = false, systemBCProject = false, linesearch = false;
inline int boundaryType = 0, lsolver = 0, Ainv = 0, smoother = 0, coarseSolver = 0, levelCnt = 1, times = 1, levelscale = 0, debugMode = 0;
inline double omega = 1, topomega = 0.1;
inline bool useBaselineMultigrid = false, topDownMGS = false; // accepted by the parser like the reference's; not implemented here: optionsFromSettings() refuses them
inline bool revealJacobi = false, revealVcycle = false;       // timing printouts of the reference: accepted, without effect
} // namespace HOTSettings

// the flags of Projects/multigrid/main.cpp:40-84 (every registered flag is accepted, the ones outside the hot path are consumed and dropped); unknown
// flags and flags without their value throw like FLAGS::ParseFlags (Lib/Ziran/CS/Util/CommandLineFlags.h:15-22,185-206).  Compared with the reference's
// own parser in tests/test_flags_ref.py.
inline void parseFlags(int argc, const char* const* argv)
{
    using namespace HOTSettings;
    auto need = [&](int& i) -> const char* {
        if (i + 1 >= argc) throw HotError(std::string("Not enough arguments to ") + argv[i]);
        return argv[++i];
    };
    for (int i = 1; i < argc; ++i) {
        const std::string f = argv[i];
        if (f == "--usecn") useCN = true;
        else if (f == "--adaptiveH") useAdaptiveHessian = true;
        else if (f == "--matfree") matrixFree = true;
        else if (f == "--project") project = true;
        else if (f == "--bcproject") systemBCProject = true;
        else if (f == "--linesearch") linesearch = true;
        else if (f == "--baseline") useBaselineMultigrid = true;
        else if (f == "--topDownMGS") topDownMGS = true;
        else if (f == "--showresidual") revealJacobi = true;
        else if (f == "--showvcycle") revealVcycle = true;
        else if (f == "--3d" || f == "--double" || f == "--help" || f == "--run_diff_test") {}
        else if (f == "-dbg") debugMode = std::stoi(need(i));
        else if (f == "-cneps") cneps = std::stod(need(i));
        else if (f == "-bc") boundaryType = std::stoi(need(i));
        else if (f == "-lsolver") lsolver = std::stoi(need(i));
        else if (f == "-Ainv") Ainv = std::stoi(need(i));
        else if (f == "-smoother") smoother = std::stoi(need(i));
        else if (f == "-coarseSolver") coarseSolver = std::stoi(need(i));
        else if (f == "-mg_level") levelCnt = std::stoi(need(i));
        else if (f == "-mg_times") times = std::stoi(need(i));
        else if (f == "-mg_scale") levelscale = std::stoi(need(i));
        else if (f == "-mg_omega") omega = std::stod(need(i));
        else if (f == "-mg_jomega") topomega = std::stod(need(i));
        else if (f == "-test" || f == "-o" || f == "-t" || f == "-cmd0" || f == "-cmd1" || f == "-script" || f == "-i" || f == "-dtps" || f == "-restart" || f == "-v_mu") need(i);
        else throw HotError("Unknown flag " + f);
    }
}

inline hot_solver_options optionsFromSettings()
{
    if (HOTSettings::useBaselineMultigrid) throw HotError("--baseline: the geometric-multigrid baseline (MultigridSimulation.inl) is not part of this library");
    if (HOTSettings::topDownMGS) throw HotError("--topDownMGS: the top-down multigrid schedule (MultigridPreconditioner.h:534-547) is not part of this library");
    hot_solver_options o;
    hot_default_options(&o);
    o.lsolver = HOTSettings::lsolver; o.matfree = HOTSettings::matrixFree; o.project = HOTSettings::project;
    o.bcproject = HOTSettings::systemBCProject; o.linesearch = HOTSettings::linesearch; o.usecn = HOTSettings::useCN;
    o.adaptive_h = HOTSettings::useAdaptiveHessian; o.mg_level = HOTSettings::levelCnt; o.mg_times = HOTSettings::times;
    o.mg_scale = HOTSettings::levelscale; o.smoother = HOTSettings::smoother; o.coarse_solver = HOTSettings::coarseSolver;
    o.Ainv = HOTSettings::Ainv; o.cneps = HOTSettings::cneps; o.topomega = HOTSettings::topomega;
    return o;
}

// ---- small 3x3 helpers -------------------------------------------------------------------------------------------------
inline TV sub(const TV& a, const TV& b) { return {a[0] - b[0], a[1] - b[1], a[2] - b[2]}; }
inline double dot(const TV& a, const TV& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline double norm(const TV& a) { return std::sqrt(dot(a, a)); }
inline TM identity() { return {1, 0, 0, 0, 1, 0, 0, 0, 1}; }

// ---- a8: collision objects, evaluated on the host (O(N_nodes) analytic tests per step) ------------------------------------
enum COLLISION_OBJECT_TYPE { STICKY = 1, SLIP = 2, SEPARATE = 3, GHOST = 4 }; // CollisionObject.h:51-56

// AnalyticLevelSet (Lib/Ziran/Math/Geometry/AnalyticLevelSet.h): queryInside = signed distance <= 0 + outward normal, in the level
// set's material space; describe() fills the plain-data form the device evaluation takes (hot_collider, hot_b200.h)
struct AnalyticLevelSet {
    virtual ~AnalyticLevelSet() = default;
    virtual bool query(const TV& X, double& phi, TV& n) const = 0; // true when inside (phi <= 0)
    virtual void describe(hot_collider& c) const = 0;
    // material-space bounding box (used by evalMaxSpeed); the reference's base class - and with it HalfSpace - has none (AnalyticLevelSet.cpp:121-125)
    virtual void getBounds(TV&, TV&) const { throw HotError("No bounds available."); }
};
inline TV matVec(const TM& M, const TV& x) { return {M[0] * x[0] + M[3] * x[1] + M[6] * x[2], M[1] * x[0] + M[4] * x[1] + M[7] * x[2], M[2] * x[0] + M[5] * x[1] + M[8] * x[2]}; }
inline TV matTVec(const TM& M, const TV& x) { return {M[0] * x[0] + M[1] * x[1] + M[2] * x[2], M[3] * x[0] + M[4] * x[1] + M[5] * x[2], M[6] * x[0] + M[7] * x[1] + M[8] * x[2]}; }
// Eigen::Quaternion<T>(w, x, y, z).toRotationMatrix(), column-major, the quaternion taken as it is
inline TM quaternionToMatrixRaw(double w, double x, double y, double z)
{
    return {1 - 2 * (y * y + z * z), 2 * (x * y + w * z), 2 * (x * z - w * y), 2 * (x * y - w * z), 1 - 2 * (x * x + z * z), 2 * (y * z + w * x),
        2 * (x * z + w * y), 2 * (y * z - w * x), 1 - 2 * (x * x + y * y)};
}
// Eigen::Quaternion<T>(w, x, y, z).normalized().toRotationMatrix(): what AnalyticBox / CappedCylinder do with their own q (AnalyticLevelSet.h:340, .cpp:492,574)
inline TM quaternionToMatrix(double w, double x, double y, double z)
{
    const double l = std::sqrt(w * w + x * x + y * y + z * z);
    return quaternionToMatrixRaw(w / l, x / l, y / l, z / l);
}
struct HalfSpace : AnalyticLevelSet { // AnalyticLevelSet.cpp:259-304
    TV origin, outward_normal;
    HalfSpace(const TV& o, const TV& n_in) : origin(o)
    {
        const double l = norm(n_in);
        outward_normal = {n_in[0] / l, n_in[1] / l, n_in[2] / l};
    }
    bool query(const TV& X, double& phi, TV& n) const override
    {
        phi = dot(outward_normal, sub(X, origin));
        n = outward_normal;
        return phi <= 0;
    }
    void describe(hot_collider& c) const override
    {
        c.shape = HOT_SHAPE_HALFSPACE;
        for (int d = 0; d < 3; ++d) { c.p[d] = origin[d]; c.p[3 + d] = outward_normal[d]; }
    }
};
struct Sphere : AnalyticLevelSet { // Sphere::queryInside, AnalyticLevelSet.cpp:435-452
    TV center;
    double radius;
    Sphere(const TV& c, double r) : center(c), radius(r) {}
    bool query(const TV& X, double& phi, TV& n) const override
    {
        const TV d = sub(X, center);
        const double l2 = dot(d, d);
        if (!(l2 < radius * radius)) return false;
        const double l = std::sqrt(l2);
        phi = l - radius;
        n = l < 1e-7 ? TV{1, 0, 0} : TV{d[0] / l, d[1] / l, d[2] / l};
        return true;
    }
    void getBounds(TV& lo, TV& hi) const override // AnalyticLevelSet.cpp:464-469
    {
        for (int d = 0; d < 3; ++d) { lo[d] = center[d] - radius; hi[d] = center[d] + radius; }
    }
    void describe(hot_collider& c) const override
    {
        c.shape = HOT_SHAPE_SPHERE;
        for (int d = 0; d < 3; ++d) c.p[d] = center[d];
        c.p[3] = radius;
    }
};
// AnalyticBox (AnalyticLevelSet.h:311-365, .cpp:504-539): primitive box [-half_edges, half_edges], placed by rotation q = <w,x,y,z> and
// translation b; AxisAlignedAnalyticBox(min, max) (.cpp:353-368) is the q = identity case
struct AnalyticBox : AnalyticLevelSet {
    TV half_edges, b;
    TM R;
    bool axis_aligned = false; // built by axisAligned(): stands for the reference's AxisAlignedAnalyticBox, whose bounds are its corners
    TV aa_min{0, 0, 0}, aa_max{0, 0, 0};
    AnalyticBox(const TV& h, const std::array<double, 4>& q, const TV& b_in) : half_edges(h), b(b_in), R(quaternionToMatrix(q[0], q[1], q[2], q[3])) {}
    static AnalyticBox axisAligned(const TV& lo, const TV& hi)
    {
        AnalyticBox box({(hi[0] - lo[0]) / 2, (hi[1] - lo[1]) / 2, (hi[2] - lo[2]) / 2}, {1, 0, 0, 0}, {(lo[0] + hi[0]) / 2, (lo[1] + hi[1]) / 2, (lo[2] + hi[2]) / 2});
        box.axis_aligned = true; box.aa_min = lo; box.aa_max = hi;
        return box;
    }
    void getBounds(TV& lo, TV& hi) const override // AnalyticBox: bounding sphere around b (.cpp:541-548); AxisAlignedAnalyticBox: the box (.cpp:370-374)
    {
        if (axis_aligned) { lo = aa_min; hi = aa_max; return; }
        const double r = norm(half_edges);
        for (int d = 0; d < 3; ++d) { lo[d] = -r + b[d]; hi[d] = r + b[d]; }
    }
    bool query(const TV& X, double& phi, TV& n) const override
    {
        const TV Xp = matTVec(R, sub(X, b));
        const double d[3] = {std::fabs(Xp[0]) - half_edges[0], std::fabs(Xp[1]) - half_edges[1], std::fabs(Xp[2]) - half_edges[2]};
        int a = 0;
        if (d[1] > d[a]) a = 1;
        if (d[2] > d[a]) a = 2;
        const double q0 = std::max(d[0], 0.0), q1 = std::max(d[1], 0.0), q2 = std::max(d[2], 0.0);
        phi = std::min(d[a], 0.0) + std::sqrt(q0 * q0 + q1 * q1 + q2 * q2);
        if (!(phi <= 0)) return false;
        TV Np{0, 0, 0};
        Np[a] = Xp[a] < 0 ? -1.0 : 1.0; // d phi / d X_primitive inside the box: the arg-max axis
        n = matVec(R, Np);
        return true;
    }
    void describe(hot_collider& c) const override
    {
        c.shape = HOT_SHAPE_BOX;
        for (int d = 0; d < 3; ++d) { c.p[d] = half_edges[d]; c.shape_b[d] = b[d]; }
        std::copy(R.begin(), R.end(), c.shape_R);
    }
};
// CappedCylinder (AnalyticLevelSet.h:220-309): primitive along y, centred at the origin, placed by q = <w,x,y,z> and b
struct CappedCylinder : AnalyticLevelSet {
    double radius, height;
    TV b;
    TM R;
    // the reference normalises q = <w,x,y,z> as a 4-vector and hands THAT VECTOR to Eigen::Quaternion (AnalyticLevelSet.h:248-251), whose vector constructor
    // reads coefficients in storage order (x, y, z, w): the rotation actually used is the one of the quaternion <w = q[3], x = q[0], y = q[1], z = q[2]>
    // (for q = <1,0,0,0> a half turn about x, which maps the y-axis cylinder onto itself).  Replicated (tests/test_collider_ref.py).
    CappedCylinder(double r, double h, const std::array<double, 4>& q, const TV& b_in) : radius(r), height(h), b(b_in), R(quaternionToMatrix(q[3], q[0], q[1], q[2])) {}
    bool query(const TV& X, double& phi, TV& n) const override
    {
        const TV Xp = matTVec(R, sub(X, b));
        const double rxz = std::sqrt(Xp[0] * Xp[0] + Xp[2] * Xp[2]);
        const double d0 = rxz - radius, d1 = std::fabs(Xp[1]) - 0.5 * height;
        const double q0 = std::max(d0, 0.0), q1 = std::max(d1, 0.0);
        phi = std::min(std::max(d0, d1), 0.0) + std::sqrt(q0 * q0 + q1 * q1);
        if (!(phi <= 0)) return false;
        TV Np{0, 0, 0};
        if (d0 >= d1) {
            if (rxz > 0) { Np[0] = Xp[0] / rxz; Np[2] = Xp[2] / rxz; }
            else Np[0] = 1.0;
        }
        else Np[1] = Xp[1] < 0 ? -1.0 : 1.0;
        n = matVec(R, Np);
        return true;
    }
    void getBounds(TV& lo, TV& hi) const override // AnalyticLevelSet.h:287-293: bounding sphere around b
    {
        const double r = std::sqrt(radius * radius + (0.5 * height) * (0.5 * height));
        for (int d = 0; d < 3; ++d) { lo[d] = -r + b[d]; hi[d] = r + b[d]; }
    }
    void describe(hot_collider& c) const override
    {
        c.shape = HOT_SHAPE_CAPPED_CYLINDER;
        c.p[0] = radius; c.p[1] = height;
        for (int d = 0; d < 3; ++d) c.shape_b[d] = b[d];
        std::copy(R.begin(), R.end(), c.shape_R);
    }
};

// AnalyticCollisionObject (CollisionObject.h:46-120): level set under x = R s X + b with rates omega, dsdt, dbdt
struct AnalyticCollisionObject {
    std::shared_ptr<AnalyticLevelSet> ls;
    COLLISION_OBJECT_TYPE type;
    double friction = 0;
    TM R{1, 0, 0, 0, 1, 0, 0, 0, 1};
    double s = 1, dsdt = 0;
    TV b{0, 0, 0}, dbdt{0, 0, 0}, omega{0, 0, 0};
    std::function<void(double, AnalyticCollisionObject&)> updateState; // collision_objects[k]->updateState(t + dt), MultigridSimulation.h:292-295
    AnalyticCollisionObject(std::shared_ptr<AnalyticLevelSet> l, COLLISION_OBJECT_TYPE t) : ls(std::move(l)), type(t) {}
    void setRotation(const std::array<double, 4>& q) { R = quaternionToMatrixRaw(q[0], q[1], q[2], q[3]); } // <w, x, y, z>, NOT normalised: Rotation(q), Rotation.h:51-56
    void setAngularVelocity(const TV& w) { omega = w; }
    void setTranslation(const TV& b_in, const TV& dbdt_in) { b = b_in; dbdt = dbdt_in; }

    // evalMaxSpeed, CollisionObject.cpp:201-238: the largest object speed at the corners where the particles' box and the object's bounds meet
    double evalMaxSpeed(const TV& p_min_corner, const TV& p_max_corner) const
    {
        const auto velocity = [&](const TV& x) {
            const TV xb = sub(x, b);
            const double k = dsdt * (1 / s);
            return TV{omega[1] * xb[2] - omega[2] * xb[1] + k * xb[0] + dbdt[0], omega[2] * xb[0] - omega[0] * xb[2] + k * xb[1] + dbdt[1],
                omega[0] * xb[1] - omega[1] * xb[0] + k * xb[2] + dbdt[2]};
        };
        if (dsdt != 0 || norm(omega) != 0) {
            TV lo, hi;
            ls->getBounds(lo, hi);
            const auto overlaps = [](const TV& a_lo, const TV& y, const TV& a_hi) { // (lo < y).any() && (y < hi).any(), as the reference writes it
                return (a_lo[0] < y[0] || a_lo[1] < y[1] || a_lo[2] < y[2]) && (y[0] < a_hi[0] || y[1] < a_hi[1] || y[2] < a_hi[2]);
            };
            double max_speed = 0;
            for (int i = 0; i < 8; ++i) {
                TV x;
                for (int d = 0; d < 3; ++d) x[d] = (i & (1 << d)) ? p_min_corner[d] : p_max_corner[d];
                TV X = matTVec(R, sub(x, b));
                for (int d = 0; d < 3; ++d) X[d] *= 1 / s;
                if (overlaps(lo, X, hi)) max_speed = std::max(max_speed, norm(velocity(x)));
            }
            for (int i = 0; i < 8; ++i) {
                TV X;
                for (int d = 0; d < 3; ++d) X[d] = ((i & (1 << d)) ? lo[d] : hi[d]) * s;
                TV x = matVec(R, X);
                for (int d = 0; d < 3; ++d) x[d] += b[d];
                if (overlaps(p_min_corner, x, p_max_corner)) max_speed = std::max(max_speed, norm(velocity(x)));
            }
            return max_speed;
        }
        return norm(dbdt);
    }

    // detectAndResolveCollision, CollisionObject.cpp:384-452 (material velocity 0)
    bool detectAndResolveCollision(const TV& x, TV& v, TV& n) const
    {
        if (type == GHOST) return false;
        n = {NAN, NAN, NAN};
        const TV xb = sub(x, b);
        const TV Xr = matTVec(R, xb);
        const TV X{Xr[0] / s, Xr[1] / s, Xr[2] / s};
        double phi;
        TV N;
        if (!ls->query(X, phi, N)) return false;
        const double k = dsdt / s;
        const TV vo{omega[1] * xb[2] - omega[2] * xb[1] + k * xb[0] + dbdt[0], omega[2] * xb[0] - omega[0] * xb[2] + k * xb[1] + dbdt[1],
            omega[0] * xb[1] - omega[1] * xb[0] + k * xb[2] + dbdt[2]};
        for (int d = 0; d < 3; ++d) v[d] -= vo[d];
        if (type == STICKY) v = {0, 0, 0};
        else {
            n = matVec(R, N);
            const double dn = dot(v, n);
            if (type == SLIP || dn < 0) {
                for (int d = 0; d < 3; ++d) v[d] -= n[d] * dn;
                if (friction != 0 && dn < 0) {
                    const double l = norm(v);
                    if (-dn * friction < l)
                        for (int d = 0; d < 3; ++d) v[d] += v[d] / l * dn * friction;
                    else v = {0, 0, 0};
                }
            }
        }
        for (int d = 0; d < 3; ++d) v[d] += vo[d];
        return true;
    }
    hot_collider describe() const
    {
        hot_collider c;
        std::memset(&c, 0, sizeof c);
        c.type = (int)type;
        c.friction = friction;
        c.shape_R[0] = c.shape_R[4] = c.shape_R[8] = 1.0;
        ls->describe(c);
        std::copy(R.begin(), R.end(), c.R);
        c.s = s; c.dsdt = dsdt;
        for (int d = 0; d < 3; ++d) { c.b[d] = b[d]; c.dbdt[d] = dbdt[d]; c.omega[d] = omega[d]; }
        return c;
    }
};

// multiObjectCollision, CollisionObject.cpp:108-149
inline bool multiObjectCollision(const std::vector<AnalyticCollisionObject>& objects, const TV& xi, TV& vi, TM& normal_basis, TV& wn)
{
    bool any = false;
    int slip_count = 0;
    normal_basis.fill(0.0);
    for (const auto& o : objects) {
        if (o.type == GHOST) continue;
        TV n;
        const bool collide = o.detectAndResolveCollision(xi, vi, n);
        any = any || collide;
        if (!collide) continue;
        if (o.type == STICKY) {
            wn = {0, 0, 0};
            normal_basis = identity();
            break;
        }
        for (int c = 0; c < slip_count; ++c) { // Gram-Schmidt the normals
            const TV old{normal_basis[3 * c], normal_basis[3 * c + 1], normal_basis[3 * c + 2]};
            const double d = dot(old, n);
            for (int k = 0; k < 3; ++k) n[k] -= d * old[k];
        }
        wn = n;
        const double l = norm(n);
        if (l) {
            for (int k = 0; k < 3; ++k) normal_basis[3 * slip_count + k] = n[k] / l;
            if (++slip_count == 3) break;
        }
    }
    return any;
}

// RotationExtractor<T,3>::rotate (MpmSimulationBase.h:270-281): Quaternion::setFromTwoVectors(a, e_x) as a matrix
inline TM rotateToX(const TV& a_in)
{
    const double l = norm(a_in);
    const TV a{a_in[0] / l, a_in[1] / l, a_in[2] / l};
    const double c = a[0]; // a . e_x
    if (c < -1 + 1e-12) return {-1, 0, 0, 0, -1, 0, 0, 0, 1}; // antiparallel: rotate by pi about z
    const TV v{0, a[2], -a[1]}; // a x e_x
    const double k = 1 / (1 + c);
    // R = I + [v]x + [v]x^2 / (1 + c)
    return {1 + k * (-v[1] * v[1] - v[2] * v[2]), v[2] + k * v[0] * v[1], -v[1] + k * v[0] * v[2],
        -v[2] + k * v[0] * v[1], 1 + k * (-v[0] * v[0] - v[2] * v[2]), v[0] + k * v[1] * v[2],
        v[1] + k * v[0] * v[2], -v[0] + k * v[1] * v[2], 1 + k * (-v[0] * v[0] - v[1] * v[1])};
}

struct CollisionNode { // CollisionObject.h:16-45
    int node_id;
    TM P, R, Rinv;
    bool shouldRotate;
};

// the per-node body of buildInitialDvAndVnForNewton (MpmSimulationBase.cpp:1145-1176): collision test of the node at xi moving with old_v against
// all objects; on a collision fills Z = {node_id, P = I - K K^T, R, R^-1, shouldRotate} and dv_bc = v_after - old_v and returns true
inline bool collisionNodeAt(const std::vector<AnalyticCollisionObject>& objects, const TV& xi, const TV& old_v, int node_id, CollisionNode& Z, TV& dv_bc)
{
    TV vi = old_v, wn{0, 0, 0};
    TM nb;
    if (!multiObjectCollision(objects, xi, vi, nb, wn)) return false;
    const bool isSlip = wn[0] != 0 || wn[1] != 0 || wn[2] != 0;
    Z.node_id = node_id;
    Z.shouldRotate = isSlip;
    Z.R = isSlip ? rotateToX(wn) : identity();
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) Z.Rinv[r + 3 * c] = Z.R[c + 3 * r]; // rotation: inverse = transpose
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            double kk = 0;
            for (int k = 0; k < 3; ++k) kk += nb[r + 3 * k] * nb[c + 3 * k];
            Z.P[r + 3 * c] = (r == c ? 1.0 : 0.0) - kk;
        }
    for (int d = 0; d < 3; ++d) dv_bc[d] = vi[d] - old_v[d];
    return true;
}

// calculateDt (MpmSimulationBase.cpp:789-814) with evalMaxParticleSpeed (:1190-1217) on host arrays: dt = cfl dx / max(particle speed, object speed over
// the particles' box grown by (degree + 2) dx), or max_dt when nothing moves
inline double calculateDtFromArrays(long n, const double* X, const double* V, double dx, double cfl, double max_dt, const std::vector<AnalyticCollisionObject>& objects)
{
    const double Tmin = (double)std::numeric_limits<float>::lowest(); // the reference's initial value of the max reduction (:1196)
    double speed = 0;
    TV hi{Tmin, Tmin, Tmin}, neg_lo{Tmin, Tmin, Tmin};
    for (long i = 0; i < n; ++i) {
        speed = std::max(speed, std::sqrt(V[3 * i] * V[3 * i] + V[3 * i + 1] * V[3 * i + 1] + V[3 * i + 2] * V[3 * i + 2]));
        for (int d = 0; d < 3; ++d) { hi[d] = std::max(hi[d], X[3 * i + d]); neg_lo[d] = std::max(neg_lo[d], -X[3 * i + d]); }
    }
    TV p_min, p_max;
    for (int d = 0; d < 3; ++d) { p_min[d] = -neg_lo[d] - (2 + 2) * dx; p_max[d] = hi[d] + (2 + 2) * dx; } // interpolation_degree 2
    double object_speed = 0;
    for (const auto& o : objects) object_speed = std::max(o.evalMaxSpeed(p_min, p_max), object_speed);
    const double m = std::max(speed, object_speed);
    return m ? cfl * dx / m : max_dt;
}

// the byte stream of the restart files (layout: see MpmSimulationB200::writeState), host arrays in and out; C (the APIC matrices) may be null on writing
inline void writeRestart(std::ostream& out, long n, const double* X, const double* V, const double* mass, const double* vol, const double* F,
    const double* mu, const double* lam, const double* C)
{
    auto w = [&](const void* p, size_t bytes) { out.write(reinterpret_cast<const char*>(p), (std::streamsize)bytes); };
    auto u64 = [&](uint64_t v) { w(&v, 8); };
    auto i32 = [&](int v) { w(&v, 4); };
    auto header = [&](const char* name, uint64_t elem_bytes) {
        const uint64_t len = std::strlen(name);
        u64(len); w(name, len);
        i32(7);                                  // DisjointRanges::lg2_grain_size
        u64(1); u64(8); i32(0); i32((int)n);     // ranges = {[0, n)}
        u64((uint64_t)n); u64(elem_bytes);
    };
    i32((int)n);
    u64(6);
    header("X", 24); w(X, (size_t)n * 24);
    header("V", 24); w(V, (size_t)n * 24);
    header("m", 8); w(mass, (size_t)n * 8);
    header("element measure", 8); w(vol, (size_t)n * 8);
    header("F", 72); w(F, (size_t)n * 72);
    header("CorotatedIsotropic", 24);
    // CorotatedIsotropic<T,3> {bool project; T mu, lambda} (CorotatedIsotropic.h:61-62) is trivially copyable and has no RW<> specialisation, so the
    // reference writes each entry as the 24 raw bytes of the object (BinaryIO.h:41-43,102-108; its write() member is not used): the project flag,
    // 7 bytes of padding, mu, lambda
    const unsigned char head[8] = {(unsigned char)(HOTSettings::project ? 1 : 0), 0, 0, 0, 0, 0, 0, 0};
    for (long i = 0; i < n; ++i) { w(head, 8); w(&mu[i], 8); w(&lam[i], 8); }
    u64(0); u64(12);                             // trimesh_to_write.indices (Vector<int, 3>)
    u64(0); u64(8);                              // segmesh_to_write.indices (Vector<int, 2>)
    if (C) { u64((uint64_t)n); u64(72); w(C, (size_t)n * 72); }
}
struct RestartArrays {
    int n = 0;
    std::vector<double> X, V, m, vol, F, mu, lam, C;
};
inline RestartArrays readRestart(std::istream& in)
{
    RestartArrays A;
    auto r = [&](void* p, size_t bytes) {
        in.read(reinterpret_cast<char*>(p), (std::streamsize)bytes);
        if (!in) throw HotError("readState: truncated restart file");
    };
    auto u64 = [&]() { uint64_t v; r(&v, 8); return v; };
    auto i32 = [&]() { int v; r(&v, 4); return v; };
    const int n = A.n = i32();
    const uint64_t arrays = u64();
    for (uint64_t a = 0; a < arrays; ++a) {
        std::string name(u64(), '\0');
        r(&name[0], name.size());
        (void)i32();
        const uint64_t nr = u64(), rb = u64();
        if (rb != 8) throw HotError("readState: Range size mismatch");
        std::vector<int> ranges(2 * nr);
        r(ranges.data(), nr * 8);
        const uint64_t size = u64(), bytes = u64();
        if ((long)size != n) throw HotError("readState: array " + name + " does not cover all particles");
        auto take = [&](std::vector<double>& dst, uint64_t expect) {
            if (bytes != expect) throw HotError("Read error: The size of the types don't match (" + name + ")");
            dst.resize(size * expect / 8);
            r(dst.data(), size * expect);
        };
        if (name == "X") take(A.X, 24);
        else if (name == "V") take(A.V, 24);
        else if (name == "m") take(A.m, 8);
        else if (name == "element measure") take(A.vol, 8);
        else if (name == "F") take(A.F, 72);
        else if (name == "CorotatedIsotropic") {
            if (bytes != 24) throw HotError("Read error: The size of the types don't match (CorotatedIsotropic)");
            A.mu.resize(size); A.lam.resize(size);
            unsigned char head[8];
            for (uint64_t i = 0; i < size; ++i) { r(head, 8); r(&A.mu[i], 8); r(&A.lam[i], 8); } // (project flag + padding first: the flag stays a run-time setting here)
        }
        else throw HotError("Array " + name + " was not initialized before reading.");
    }
    for (int k = 0; k < 2; ++k) { // mesh index vectors
        const uint64_t size = u64(), bytes = u64();
        in.ignore((std::streamsize)(size * bytes));
    }
    A.C.assign(9 * (size_t)n, 0.0);
    if (in.peek() != std::char_traits<char>::eof()) { // trailing APIC matrices (see writeState)
        const uint64_t size = u64(), bytes = u64();
        if ((long)size == n && bytes == 72) r(A.C.data(), (size_t)n * 72);
    }
    if (A.X.empty() || A.V.empty() || A.m.empty() || A.vol.empty() || A.F.empty() || A.mu.empty()) throw HotError("readState: missing particle arrays");
    return A;
}

// ---- MpmSimulationBase surface -----------------------------------------------------------------------------------------------
class MpmSimulationB200 {
public:
    // public data members the solvers of the reference read directly (MpmSimulationBase.h:69-131)
    double dx, dt = 0, cfl = 0.6, apic_rpic_ratio = 1;
    TV gravity{0, 0, 0};
    int num_nodes = 0;
    std::vector<double> mass_matrix;
    TVStack dv, vn;
    std::vector<AnalyticCollisionObject> collision_objects;
    std::vector<CollisionNode> collision_nodes;
    double time = 0;
    hot_solve_log last_log;

    MpmSimulationB200(double dx_, double apic_rpic_ratio_ = 1.0, double cfl_ = 0.6, int device = -1)
        : dx(dx_), cfl(cfl_), apic_rpic_ratio(apic_rpic_ratio_)
    {
        h = hot_create(dx, apic_rpic_ratio, cfl, device);
        if (!h) throw HotError("hot_create failed: no usable CUDA device (hot_b200 has no CPU fallback)");
        std::memset(&last_log, 0, sizeof last_log);
    }
    ~MpmSimulationB200() { hot_destroy(h); }
    MpmSimulationB200(const MpmSimulationB200&) = delete;
    MpmSimulationB200& operator=(const MpmSimulationB200&) = delete;
    hot_sim* handle() { return h; }

    // particles.X / V / mass, APIC C, F, element measure, CorotatedIsotropic mu / lambda (original order, AoS)
    void setParticles(long n, const double* X, const double* V, const double* mass, const double* C, const double* F, const double* vol,
        const double* mu, const double* lambda)
    {
        check(hot_set_particles(h, n, X, V, mass, C, F, vol, mu, lambda));
        N = n;
        mass_p.assign(mass, mass + n); // constants of the particle set, needed by writeState
        vol_p.assign(vol, vol + n);
    }
    void getParticles(double* X, double* V, double* C, double* F) { check(hot_get_particles(h, X, V, C, F, nullptr)); }
    std::vector<double> mass_p, vol_p;
    long particleCount() const { return N; }

    // calculateDt (MpmSimulationBase.cpp:789-814), the virtual the reference's frame loop calls between steps.  Reads X and V back (O(N) device -> host
    // copy per call; a device-side reduction is the follow-up) and evaluates the rule on the host
    double calculateDt(double max_dt)
    {
        std::vector<double> X(3 * (size_t)N), V(3 * (size_t)N);
        getParticles(X.data(), V.data(), nullptr, nullptr);
        return calculateDtFromArrays(N, X.data(), V.data(), dx, cfl, max_dt, collision_objects);
    }

    void sortParticlesAndPolluteGrid() { check(hot_sort_and_activate(h)); } // MpmSimulationBase.cpp:1066-1137
    void particlesToGrid() { check(hot_p2g(h, &num_nodes)); } // :461-533
    void buildMassMatrix() // :817-826
    {
        mass_matrix.resize(num_nodes);
        if (num_nodes) check(hot_get_mass_matrix(h, mass_matrix.data()));
    }
    // a8 on the device (hot_set_colliders + hot_build_bc): the objects cross the boundary as plain data, nothing comes back but the
    // count; collision_nodes is filled on demand by fetchCollisionNodes().  false: the host evaluation below + hot_set_bc.
    bool device_colliders = true;
    void fetchCollisionNodes()
    {
        int n_bc = 0;
        check(hot_get_bc(h, &n_bc, nullptr, nullptr, nullptr, nullptr, nullptr));
        std::vector<int> node(n_bc), slip(n_bc);
        std::vector<double> P(9 * (size_t)n_bc), R(9 * (size_t)n_bc), Rinv(9 * (size_t)n_bc);
        check(hot_get_bc(h, &n_bc, node.data(), P.data(), R.data(), Rinv.data(), slip.data()));
        collision_nodes.resize(n_bc);
        for (int k = 0; k < n_bc; ++k) {
            CollisionNode& Z = collision_nodes[k];
            Z.node_id = node[k];
            Z.shouldRotate = slip[k] != 0;
            std::copy(P.begin() + 9 * (size_t)k, P.begin() + 9 * (size_t)(k + 1), Z.P.begin());
            std::copy(R.begin() + 9 * (size_t)k, R.begin() + 9 * (size_t)(k + 1), Z.R.begin());
            std::copy(Rinv.begin() + 9 * (size_t)k, Rinv.begin() + 9 * (size_t)(k + 1), Z.Rinv.begin());
        }
    }
    int num_collision_nodes = 0;
    // :1139-1184
    void buildInitialDvAndVnForNewton()
    {
        const int bc_mode = (HOTSettings::systemBCProject && HOTSettings::boundaryType == 1) ? 1 : 0; // MultigridSimulation.h:104-125
        if (device_colliders) {
            std::vector<hot_collider> objs;
            for (const auto& o : collision_objects) objs.push_back(o.describe());
            check(hot_set_dt_gravity(h, dt, gravity.data()));
            check(hot_set_colliders(h, (int)objs.size(), objs.data()));
            check(hot_build_bc(h, bc_mode, &num_collision_nodes));
            collision_nodes.clear();
            dv.resize(3 * (size_t)num_nodes); // (host copies of dv / vn are refreshed by whoever reads them: hot_get_dv)
            return;
        }
        buildInitialDvAndVnForNewtonOnHost();
        num_collision_nodes = (int)collision_nodes.size();
    }
    // the same with the collider tests on the host (O(N_n) analytic tests, the whole grid read back): kept as the comparator
    void buildInitialDvAndVnForNewtonOnHost()
    {
        std::vector<int> coord(3 * (size_t)num_nodes);
        std::vector<long long> idx((size_t)hot_num_pages(h) * 32);
        std::vector<double> gv(3 * idx.size());
        if (num_nodes) check(hot_get_id2coord(h, coord.data()));
        check(hot_get_grid(h, idx.data(), nullptr, gv.data()));
        vn.assign(3 * (size_t)num_nodes, 0.0);
        for (size_t a = 0; a < idx.size(); ++a)
            if (idx[a] >= 0)
                for (int d = 0; d < 3; ++d) vn[3 * idx[a] + d] = gv[3 * a + d];
        collision_nodes.clear();
        std::vector<int> node_id, slip;
        std::vector<double> P, R, Rinv, dv_bc;
        for (int i = 0; i < num_nodes; ++i) {
            const TV xi{coord[3 * i] * dx, coord[3 * i + 1] * dx, coord[3 * i + 2] * dx};
            const TV old_v{vn[3 * i], vn[3 * i + 1], vn[3 * i + 2]};
            CollisionNode Z;
            TV dvi;
            if (!collisionNodeAt(collision_objects, xi, old_v, i, Z, dvi)) continue;
            collision_nodes.push_back(Z);
            node_id.push_back(i);
            slip.push_back(Z.shouldRotate);
            P.insert(P.end(), Z.P.begin(), Z.P.end());
            R.insert(R.end(), Z.R.begin(), Z.R.end());
            Rinv.insert(Rinv.end(), Z.Rinv.begin(), Z.Rinv.end());
            for (int d = 0; d < 3; ++d) dv_bc.push_back(dvi[d]);
        }
        const int mode = (HOTSettings::systemBCProject && HOTSettings::boundaryType == 1) ? 1 : 0; // MultigridSimulation.h:104-125
        check(hot_set_dt_gravity(h, dt, gravity.data()));
        check(hot_set_bc(h, mode, (int)node_id.size(), node_id.data(), P.data(), R.data(), Rinv.data(), slip.data(), dv_bc.data()));
        dv.resize(3 * (size_t)num_nodes);
        if (num_nodes) check(hot_get_dv(h, dv.data()));
    }
    void moveNodes(const TVStack& dv_in) // :735-747 (folded into updateState on the device)
    {
        dv = dv_in;
        check(hot_set_dv(h, dv.data()));
    }
    void addScaledForces(double scale, TVStack& f) { check(hot_add_scaled_forces(h, scale, f.data())); } // :829-833
    void addScaledForceDifferentials(double scale, const TVStack& x, TVStack& f) // :835-840
    {
        check(hot_add_scaled_force_differentials(h, scale, x.data(), f.data()));
    }
    void constructNewVelocityFromNewtonResult() {} // :891-901 - fused into the node-tile staging of hot_g2p
    void gridToParticles(double dt_) // :903-1042 (+ evolveStrain)
    {
        int flags[2] = {0, 0};
        check(hot_g2p(h, dt_, flags));
        faster_than_dx = flags[0] != 0;
        faster_than_half_dx = flags[1] != 0;
    }
    bool faster_than_dx = false, faster_than_half_dx = false;
    // PlasticityApplier (Lib/Ziran/Physics/PlasticityApplier.h): the return mapping hot_g2p runs after evolveStrain (:1039-1064)
    void addVonMisesFixedCorotated(double yield_stress) { check(hot_set_plasticity(h, 1, &yield_stress)); }
    void addSnowPlasticity(double psi = 10, double theta_c = 2e-2, double theta_s = 7.5e-3, double min_Jp = 0.6, double max_Jp = 20)
    {
        const double q[5] = {psi, theta_c, theta_s, min_Jp, max_Jp};
        check(hot_set_plasticity(h, 2, q));
    }
    void applyPlasticity() { check(hot_apply_plasticity(h)); }

    // ---- restart state: MpmSimulationBase::writeState / readState (Lib/MPM/MpmSimulationBase.cpp:755-785) in the reference's
    // binary layout (binary_ver 1), so that a restart_N.dat written here is read by the reference's readState and vice versa:
    //   Scene::writeState (Lib/Ziran/Sim/Scene.h:189-206) = particles.writeData + (no element managers for MPM) + two empty
    //   mesh index vectors; DataManager::writeData (DataManager.h:263-273) = int count, u64 #arrays, then per array its name
    //   (u64 length + bytes, BinaryIO.h:167-172) and DataArray::writeData (DataArray.h:100-105) = int lg2_grain_size (7),
    //   StdVector<Range{int lower, upper}>, StdVector<T> (u64 size, u64 sizeof(T), entries).  Arrays of this path: "X", "V" (TV),
    //   "m", "element measure" (T), "F" (TM, column-major), "CorotatedIsotropic" (the 24 raw bytes of the trivially copyable object: project flag,
    //   padding, mu, lambda).  The reader looks arrays up by NAME (DataManager.h:280-293), the order is free.  Checked against the reference's own
    //   DataManager / BinaryIO code in tests/test_restart_ref.py.
    //   The APIC matrix (scratch_gradV in the reference, written there only for interpolation_degree 1) follows as a trailing
    //   StdVector<TM>; a reader that does not expect it stops before it.
    void writeState(std::ostream& out)
    {
        const long n = N;
        std::vector<double> X(3 * n), V(3 * n), C(9 * n), F(9 * n), mu(n), lam(n);
        getParticles(X.data(), V.data(), C.data(), F.data());
        check(hot_get_plastic_state(h, nullptr, mu.data(), lam.data()));
        writeRestart(out, n, X.data(), V.data(), mass_p.data(), vol_p.data(), F.data(), mu.data(), lam.data(), C.data());
    }
    void readState(std::istream& in)
    {
        RestartArrays a = readRestart(in);
        setParticles(a.n, a.X.data(), a.V.data(), a.m.data(), a.C.data(), a.F.data(), a.vol.data(), a.mu.data(), a.lam.data());
    }

    // One object over the GPUs of a box (include/hot_b200.h "one object over the GPUs of a box"; no counterpart in the single-process
    // reference): one MpmSimulationB200 per process / GPU, each holding its own particles.  Either NCCL inside the library
    // (commUniqueId on rank 0 -> distribute the 128 bytes -> initNccl on every rank) or the caller's collectives (setPartition).
    static std::array<unsigned char, 128> commUniqueId()
    {
        std::array<unsigned char, 128> id;
        if (hot_comm_unique_id(id.data()) != 0) throw HotError("hot_comm_unique_id failed (NCCL not loadable)");
        return id;
    }
    void initNccl(int rank, int world, const std::array<unsigned char, 128>& id) { check(hot_comm_init_nccl(h, rank, world, id.data())); }
    void setPartition(int rank, int world, const hot_transport* transport) { check(hot_set_partition(h, rank, world, transport)); }
    struct Partition { long rank, world, neighbors, shared_pages, exchange_pages, owned_nodes, global_nodes, particles; };
    Partition partition()
    {
        long o[8];
        check(hot_get_partition(h, o));
        return Partition{o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7]};
    }

public:

    // MultigridSimulation::startBackwardEuler (MultigridSimulation.h:167-186)
    void startBackwardEuler()
    {
        buildMassMatrix();
        buildInitialDvAndVnForNewton();
        check(hot_backup_strain(h));
    }
    // MultigridSimulation::backwardEulerStep (:188-233): tolerances, Newton / L-BFGS by HOTSettings::lsolver, restoreStrain
    void backwardEulerStep()
    {
        buildMassMatrix();
        buildInitialDvAndVnForNewton();
        const hot_solver_options o = optionsFromSettings();
        check(hot_backward_euler_step(h, &o, &last_log));
        if (num_nodes) check(hot_get_dv(h, dv.data()));
    }
    // MultigridSimulation::advanceOneTimeStep (:235-297): reinitialize -> P2G -> BE solve -> collider update -> G2P
    void advanceOneTimeStep(double dt_)
    {
        dt = dt_;
        sortParticlesAndPolluteGrid();
        particlesToGrid();
        backwardEulerStep();
        for (auto& o : collision_objects)
            if (o.updateState) o.updateState(time + dt, o);
        gridToParticles(dt);
        time += dt;
    }

    void check(int rc) const
    {
        if (rc != 0) throw HotError(hot_last_error(h));
    }

private:
    hot_sim* h = nullptr;
    long N = 0;
};

// ---- MpmForceHelperBase surface (Lib/MPM/Force/MpmForceHelperBase.h:18-46) of FBasedMpmForceHelper<CorotatedIsotropic<T,3>> ----------
// The helper virtuals the reference's MpmForceBase / ImplicitSolverObjective call.  On the device the helper's per-particle state
// (F, Fn, scratch = SVD, vol P Fn^T) lives in sorted SoA rows of the handle, so the methods are thin: what the reference computes
// in a helper call is computed by the fused kernels behind the C ABI, and a method that has no separate device stage says so.
class FBasedMpmForceHelperB200 {
public:
    using Hessian = std::array<double, 81>; // Eigen::Matrix<T, 9, 9>, column-major, index ij = i + 3 j (CorotatedIsotropic.h:198-227)
    explicit FBasedMpmForceHelperB200(MpmSimulationB200& sim_) : sim(sim_) {}
    void reinitialize() {}                                                                   // FBasedMpmForceHelper.cpp:20-23 (scratch is per step here)
    void backupStrain() { sim.check(hot_backup_strain(sim.handle())); }                     // :25-33
    void restoreStrain() { sim.check(hot_restore_strain(sim.handle())); }                   // :35-44
    bool needGradVn() { return false; }
    // updateImplicitState (:72-97): scratch update + vol P Fn^T of every particle; vPFnT (9 N, original order) is optional.
    // evolveStrain with the trial dv precedes it in ImplicitSolverObjective::updateState; both are one kernel here.
    void updateImplicitState(double* vPFnT = nullptr)
    {
        sim.check(hot_update_state(sim.handle(), nullptr, nullptr));
        if (vPFnT) sim.check(hot_get_stress(sim.handle(), vPFnT, nullptr));
    }
    // evolveStrain (:100-114): F = (I + dt gradV) Fn.  Fused into hot_update_state (trial states) and hot_g2p (end of step): no stage of its own.
    void evolveStrain(double /*dt*/) {}
    double totalEnergy() // :116-136
    {
        double e = 0;
        sim.check(hot_strain_energy(sim.handle(), &e));
        return e;
    }
    // computeStressDifferential (:138-160) + the rasterisation that follows it in MpmForceBase::addScaledForceDifferential
    // (MpmForceBase.cpp:261-306): f += scale * df(x) for a DOF field x
    void computeStressDifferential(double scale, const TVStack& x, TVStack& f)
    {
        sim.check(hot_add_scaled_force_differentials(sim.handle(), scale, x.data(), f.data()));
    }
    // runLambdaWithDifferential (FBasedMpmForceHelper.h:63-120): func(i, dPdF_i, Fn_i, -1, -1, false) for every particle, in the
    // reference's order (8 colour passes over the page groups, sorted order inside a group).  dPdF comes from the device model
    // (hot_corotated_eval); opt 1 stores it, opt 2 reuses the stored one.
    void runLambdaWithDifferential(const std::function<void(int, const Hessian&, const TM&, double, double, bool)>& func, int opt = 0)
    {
        hot_sim* h = sim.handle();
        const long n = sim.particleCount();
        std::vector<double> F(9 * (size_t)n), Fn(9 * (size_t)n);
        sim.check(hot_get_particles(h, nullptr, nullptr, nullptr, F.data(), nullptr));
        sim.check(hot_get_strain_backup(h, Fn.data()));
        if (opt < 2) evaluate(F, stored);
        walk([&](int i) {
            Hessian H;
            TM fn;
            std::copy(stored.begin() + 81 * (size_t)i, stored.begin() + 81 * (size_t)(i + 1), H.begin());
            std::copy(Fn.begin() + 9 * (size_t)i, Fn.begin() + 9 * (size_t)(i + 1), fn.begin());
            func(i, H, fn, -1.0, -1.0, false);
        });
        if (opt == 0) stored.clear();
    }
    // computePerNodeCNTolerance (:123-157): func(i, dPdF(F = I), -1, false)
    void computePerNodeCNTolerance(const std::function<void(int, const Hessian&, double, bool)>& func)
    {
        const long n = sim.particleCount();
        std::vector<double> F(9 * (size_t)n, 0.0), H;
        for (long i = 0; i < n; ++i) F[9 * i] = F[9 * i + 4] = F[9 * i + 8] = 1.0;
        evaluate(F, H);
        walk([&](int i) {
            Hessian Hi;
            std::copy(H.begin() + 81 * (size_t)i, H.begin() + 81 * (size_t)(i + 1), Hi.begin());
            func(i, Hi, -1.0, false);
        });
    }

private:
    MpmSimulationB200& sim;
    std::vector<double> stored; // the function-static dPdF cache of the reference (FBasedMpmForceHelper.h:71)
    // dPdF of every particle: one device evaluation per distinct (mu, lambda)
    void evaluate(const std::vector<double>& F, std::vector<double>& H)
    {
        hot_sim* h = sim.handle();
        const long n = sim.particleCount();
        std::vector<double> mu(n), lam(n);
        sim.check(hot_get_plastic_state(h, nullptr, mu.data(), lam.data()));
        H.assign(81 * (size_t)n, 0.0);
        std::vector<char> done(n, 0);
        std::vector<double> Fb, Hb;
        std::vector<long> ids;
        for (long a = 0; a < n; ++a) {
            if (done[a]) continue;
            ids.clear(); Fb.clear();
            for (long i = a; i < n; ++i)
                if (!done[i] && mu[i] == mu[a] && lam[i] == lam[a]) {
                    done[i] = 1;
                    ids.push_back(i);
                    Fb.insert(Fb.end(), F.begin() + 9 * (size_t)i, F.begin() + 9 * (size_t)(i + 1));
                }
            Hb.resize(81 * ids.size());
            sim.check(hot_corotated_eval(h, (long)ids.size(), Fb.data(), mu[a], lam[a], HOTSettings::project ? 1 : 0, nullptr, nullptr, nullptr, nullptr,
                Hb.data(), nullptr, nullptr, nullptr));
            for (size_t k = 0; k < ids.size(); ++k) std::copy(Hb.begin() + 81 * k, Hb.begin() + 81 * (k + 1), H.begin() + 81 * (size_t)ids[k]);
        }
    }
    // the reference's particle walk: colour passes over the page groups (MpmSimulationBase.h:251-264)
    template <class Fn_>
    void walk(Fn_ f)
    {
        hot_sim* h = sim.handle();
        const long n = sim.particleCount(), G = hot_num_groups(h);
        std::vector<int> order(n), first(G), last(G);
        std::vector<unsigned long long> block(G);
        sim.check(hot_get_sort(h, nullptr, order.data(), nullptr));
        sim.check(hot_get_groups(h, first.data(), last.data(), block.data()));
        for (unsigned long long colour = 0; colour < 8; ++colour)
            for (long g = 0; g < G; ++g) {
                if ((block[g] & 7ull) != colour) continue;
                for (int idx = first[g]; idx <= last[g]; ++idx) f(order[idx]);
            }
    }
};

// ---- Krylov / Newton objective concept ---------------------------------------------------------------------------------------
// ImplicitSolverObjective::shouldExitByCN (ImplicitSolver.h:171-215) on host arrays: without --usecn the l2 norm against cneps, with it the residual
// scaled per node by the CN tolerance against the node count
inline bool shouldExitByCN(const TVStack& residual, const std::vector<double>& nodeCNTol, int num_nodes)
{
    if (!HOTSettings::useCN) {
        double s = 0;
        for (double x : residual) s += x * x;
        return std::sqrt(s) < HOTSettings::cneps;
    }
    double scaled = 0;
    for (int i = 0; i < num_nodes; ++i)
        scaled += (residual[3 * i] * residual[3 * i] + residual[3 * i + 1] * residual[3 * i + 1] + residual[3 * i + 2] * residual[3 * i + 2]) / (nodeCNTol[i] * nodeCNTol[i]);
    if (num_nodes == 0) return true;
    return scaled < num_nodes;
}
// transformResidual (R, :117-125) / recoverSolution (R^-1, :103-115): slip collision nodes live in their rotated frame when the system is BC-projected
inline void rotateCollisionNodes(const std::vector<CollisionNode>& nodes, TVStack& v, bool inverse)
{
    if (!(HOTSettings::systemBCProject && HOTSettings::boundaryType == 1)) return;
    for (const CollisionNode& c : nodes)
        if (c.shouldRotate) {
            const TV x{v[3 * c.node_id], v[3 * c.node_id + 1], v[3 * c.node_id + 2]};
            const TV y = matVec(inverse ? c.Rinv : c.R, x);
            for (int d = 0; d < 3; ++d) v[3 * c.node_id + d] = y[d];
        }
}

class ImplicitSolverObjectiveB200 {
public:
    using Scalar = double;
    using NewtonVector = TVStack;
    MpmSimulationB200& simulation;
    bool matrix_free = false;
    std::function<void(TVStack&)> project;
    std::function<void(const TVStack&, TVStack&)> precondition;
    double Ek = 0;

    explicit ImplicitSolverObjectiveB200(MpmSimulationB200& sim) : simulation(sim)
    {
        project = [this](TVStack& v) { simulation.check(hot_project(simulation.handle(), v.data())); };
        precondition = [this](const TVStack& in, TVStack& out) { // rebuildPreconditioner installs the V-cycle here (SparseMatrixFast.h:46-58)
            out.resize(in.size());
            simulation.check(hot_vcycle(simulation.handle(), in.data(), out.data()));
        };
    }
    void updateState(const TVStack& dv, bool = false) // ImplicitSolver.h:237-252
    {
        simulation.check(hot_update_state(simulation.handle(), dv.data(), HOTSettings::linesearch ? &Ek : nullptr));
    }
    void computeResidual(TVStack& r, bool = false) // :128-155
    {
        r.resize(3 * (size_t)simulation.num_nodes);
        simulation.check(hot_compute_residual(simulation.handle(), r.data()));
    }
    double computeNorm(const TVStack& r) const { return std::sqrt(innerProduct(r, r)); } // :158-171
    double innerProduct(const TVStack& a, const TVStack& b) const // :213-234
    {
        double s = 0;
        for (size_t i = 0; i < a.size(); ++i) s += a[i] * b[i];
        return s;
    }
    void multiply(const TVStack& x, TVStack& b) const // :741-763
    {
        b.resize(x.size());
        if (matrix_free) simulation.check(hot_hessian_apply_mf(simulation.handle(), x.data(), b.data()));
        else simulation.check(hot_spmv(simulation.handle(), 0, x.data(), b.data()));
    }
    void buildMatrix(bool projectSystem) { simulation.check(hot_build_matrix(simulation.handle(), projectSystem)); } // :470-603
    void HinvApproxInit() // :335-353
    {
        buildMatrix(true);
        simulation.check(hot_build_mg(simulation.handle(), HOTSettings::levelCnt, HOTSettings::smoother, HOTSettings::coarseSolver,
            HOTSettings::Ainv, HOTSettings::times, HOTSettings::levelscale, HOTSettings::topomega));
    }
    // the rest of the duck-typed objective the reference's Newton / L-BFGS templates call (ImplicitSolver.h); host logic over the C ABI.  The library
    // runs the same sequence internally (hot_backward_euler_step), which is the tested path; these members exist so that an outer loop kept in the
    // reference can drive the device operators one call at a time.
    std::vector<double> nodeCNTol;
    TVStack dv0;
    double cg_tolerance = 1;      // cg.setTolerance(1) of the constructor (:86-88); backwardEulerStep sets maxcntol with --usecn (MultigridSimulation.h:207)
    int cg_max_iterations = 10000;
    void evaluatePerNodeCNTolerance(double eps, double dt) // :667-696
    {
        nodeCNTol.resize((size_t)simulation.num_nodes);
        simulation.check(hot_eval_cn_tolerance(simulation.handle(), eps, dt, nodeCNTol.data()));
    }
    bool shouldExitByCN(const TVStack& residual) { return hot_b200::shouldExitByCN(residual, nodeCNTol, simulation.num_nodes); } // :171-215
    void transformResidual(TVStack& r) { rotateCollisionNodes(simulation.collision_nodes, r, false); }                         // :117-125
    void recoverSolution(TVStack& ddv) { rotateCollisionNodes(simulation.collision_nodes, ddv, true); }                        // :103-115
    void resetLSFlag(const TVStack& dv) { dv0 = dv; }                                                                           // :277-282
    double lineSearch(TVStack& ddv, TVStack& residual, double alpha) // :313-333 (capped at 60 halvings like the library: a NaN energy would loop forever)
    {
        TVStack dvnew(ddv.size());
        recoverSolution(ddv);
        const bool ls = HOTSettings::linesearch;
        HOTSettings::linesearch = true; // updateState must return the energy
        const double Ek0 = Ek;
        int halvings = 0;
        do {
            for (size_t i = 0; i < ddv.size(); ++i) dvnew[i] = dv0[i] + ddv[i] * alpha;
            updateState(dvnew, true);
            alpha *= 0.5;
        } while (Ek > Ek0 && ++halvings < 60);
        HOTSettings::linesearch = ls;
        alpha *= 2;
        for (double& x : ddv) x *= alpha;
        transformResidual(ddv);
        computeResidual(residual, true);
        dv0 = dvnew;
        return alpha;
    }
    void computeStep(TVStack& ddv, const TVStack& residual, double /* linear_solve_relative_tolerance: the inexact CG derives its own */) // :355-432, -lsolver 2
    {
        ddv.assign(residual.size(), 0.0);
        int preconditioner = 1, iters = 0;
        if (!matrix_free) {
            buildMatrix(HOTSettings::systemBCProject);
            simulation.check(hot_build_mg(simulation.handle(), HOTSettings::levelCnt, HOTSettings::smoother, HOTSettings::coarseSolver, HOTSettings::Ainv,
                HOTSettings::times, HOTSettings::levelscale, HOTSettings::topomega));
            preconditioner = (HOTSettings::levelCnt == 1 && HOTSettings::times == 1) ? 1 : 2; // "force diagonal entry preconditioner", :381-396
        }
        else
            simulation.check(hot_build_diagonal(simulation.handle(), HOTSettings::Ainv, nullptr));
        simulation.check(hot_pcg(simulation.handle(), residual.data(), ddv.data(), cg_tolerance, cg_max_iterations, matrix_free ? 1 : 0, preconditioner, &iters));
        if (HOTSettings::linesearch) {
            TVStack r = residual;
            lineSearch(ddv, r, 1.0);
        }
    }
};

// ---- -smoother / -coarseSolver function-pointer surface (MultigridPreconditioner.h:68-73) -----------------------------------
struct MPMSpMatB200 { // stands for MPMSpMat& A: one level of the device hierarchy
    MpmSimulationB200* sim;
    int level;
};
using SmoothFunc = void (*)(TVStack& u, TVStack& r, TVStack& du, TVStack& dAu, MPMSpMatB200& A, int iterations, double tolerance);
namespace detail {
inline void smooth(int kind, TVStack& u, TVStack& r, MPMSpMatB200& A, int iterations, double tolerance)
{
    A.sim->check(hot_smooth(A.sim->handle(), A.level, kind, u.data(), r.data(), iterations, tolerance, nullptr));
}
} // namespace detail
inline void jacobi_smooth(TVStack& u, TVStack& r, TVStack&, TVStack&, MPMSpMatB200& A, int it, double tol) { detail::smooth(0, u, r, A, it, tol); }
inline void optimal_jacobi_smooth(TVStack& u, TVStack& r, TVStack&, TVStack&, MPMSpMatB200& A, int it, double tol) { detail::smooth(1, u, r, A, it, tol); }
inline void cg_smooth(TVStack& u, TVStack& r, TVStack&, TVStack&, MPMSpMatB200& A, int it, double tol) { detail::smooth(2, u, r, A, it, tol); }
inline void gs_smooth(TVStack& u, TVStack& r, TVStack&, TVStack&, MPMSpMatB200& A, int it, double tol) { detail::smooth(5, u, r, A, it, tol); }
// selectSmoother, MultigridPreconditioner.h:496-521
inline SmoothFunc selectSmoother(int opt)
{
    switch (opt) {
    case 0: return jacobi_smooth;
    case 1: return optimal_jacobi_smooth;
    case 2: return cg_smooth;
    case 5: return gs_smooth;
    default: throw HotError("No proper smoother is selected!");
    }
}

} // namespace hot_b200
