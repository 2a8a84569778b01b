// TEST INFRASTRUCTURE ONLY — not part of the product path.
//
// Row a8 (+ f3): the REFERENCE'S OWN collision objects.  Lib/Ziran/Math/Geometry/AnalyticLevelSet.cpp (HalfSpace, Sphere, AxisAlignedAnalyticBox,
// AnalyticBox, CappedCylinder: signed distance + normal, the latter through Lib/Ziran/Math/Nonlinear/AutoDiff.h) and
// Lib/Ziran/Math/Geometry/CollisionObject.cpp (AnalyticCollisionObject::detectAndResolveCollision :384-452 with the object transform x = R s X + b and its
// rates, STICKY / SLIP / SEPARATE / GHOST, friction; multiObjectCollision :108-149 with its Gram-Schmidt of the slip normals) and Rotation.h are compiled
// where they lie - whole files, with their explicit instantiations, ZIRAN_WITH_VDB undefined (mesh export only) - against the Eigen stand-in
// (oracle/ref_shim/mini_eigen_geom.h restates Eigen's Quaternion formulas: third-party arithmetic).
// Written out here, because Lib/MPM/MpmSimulationBase.{h,cpp} cannot be compiled: the per-node body of buildInitialDvAndVnForNewton
// (MpmSimulationBase.cpp:1139-1184: CollisionNode {P = I - K K^T, R, R^-1, shouldRotate}, Newton initial guess) and RotationExtractor<T,3>::rotate
// (MpmSimulationBase.h:270-281).
// Built by oracle/Makefile into oracle/_ref/libcollider_ref.so; tests/golden/make_collider_golden.py, tests/test_collider_ref.py.
#include <functional>
#include <memory>
#include <Ziran/CS/Util/Debug.h>
#include <Ziran/Math/Geometry/AnalyticLevelSet.cpp>
#include <Ziran/Math/Geometry/CollisionObject.cpp>

using namespace ZIRAN;
namespace {
typedef double T;
constexpr int dim = 3;
typedef Vector<T, 3> TV;
typedef Vector<T, 4> TV4;
typedef Matrix<T, 3, 3> TM;
typedef AnalyticCollisionObject<T, 3> Object;
constexpr int OBJ_DOUBLES = 33;

TM rotate(const TV& a) // RotationExtractor<T, 3>::rotate, MpmSimulationBase.h:270-281
{
    TV unitNorm = TV::Zero();
    unitNorm(0) = 1;
    TM R = static_cast<TM>(Eigen::Quaternion<T>().setFromTwoVectors(a, unitNorm));
    return R;
}

// one object = 33 doubles: type, shape, friction, p[8], shape quaternion <w,x,y,z>, shape_b[3], object quaternion <w,x,y,z>, s, b[3], omega[3], dsdt, dbdt[3]
// shape 0 HalfSpace(origin p[0..2], outward normal p[3..5]); 1 Sphere(center p[0..2], radius p[3]); 2 AnalyticBox(half edges p[0..2], shape q, shape b);
// 3 CappedCylinder(radius p[0], height p[1], shape q, shape b); 4 AxisAlignedAnalyticBox(min p[0..2], max p[3..5])
void build(int n_obj, const double* o, StdVector<std::unique_ptr<Object>>& objects)
{
    for (int k = 0; k < n_obj; ++k, o += OBJ_DOUBLES) {
        const int type = (int)o[0], shape = (int)o[1];
        const double* p = o + 3;
        TV4 sq(o[11], o[12], o[13], o[14]);
        TV sb(o[15], o[16], o[17]);
        std::unique_ptr<AnalyticLevelSet<T, 3>> ls;
        if (shape == 0) ls.reset(new HalfSpace<T, 3>(TV(p[0], p[1], p[2]), TV(p[3], p[4], p[5])));
        else if (shape == 1) ls.reset(new Sphere<T, 3>(TV(p[0], p[1], p[2]), p[3]));
        else if (shape == 2) ls.reset(new AnalyticBox<T, 3>(TV(p[0], p[1], p[2]), sq, sb));
        else if (shape == 3) ls.reset(new CappedCylinder<T, 3>(p[0], p[1], sq, sb));
        else ls.reset(new AxisAlignedAnalyticBox<T, 3>(TV(p[0], p[1], p[2]), TV(p[3], p[4], p[5])));
        objects.emplace_back(new Object(std::move(ls), (Object::COLLISION_OBJECT_TYPE)type));
        Object& obj = *objects.back();
        obj.setFriction(o[2]);
        obj.setRotation(TV4(o[18], o[19], o[20], o[21]));
        obj.setScaling(o[22], o[29]);
        obj.setTranslation(TV(o[23], o[24], o[25]), TV(o[30], o[31], o[32]));
        obj.setAngularVelocity(TV(o[26], o[27], o[28]));
    }
}
} // namespace

extern "C" {

// per point: the body of the grid loop of buildInitialDvAndVnForNewton (MpmSimulationBase.cpp:1145-1182) with node position xi and grid velocity v.
// collide[i]; dv[i] = vi - old_v (collision) or gravity dt; P, R, Rinv column-major; slip[i] = shouldRotate
void zr_colliders_eval(int n_obj, const double* objs, long n, const double* xi_in, const double* v_in, const double* gravity, double dt, int* collide,
    double* dv, double* P, double* R_out, double* Rinv, int* slip)
{
    StdVector<std::unique_ptr<Object>> collision_objects;
    build(n_obj, objs, collision_objects);
    const TV g(gravity[0], gravity[1], gravity[2]);
    for (long i = 0; i < n; ++i) {
        TV old_v(v_in[3 * i], v_in[3 * i + 1], v_in[3 * i + 2]);
        TV vi = old_v;
        TV wn;
        const TV xi(xi_in[3 * i], xi_in[3 * i + 1], xi_in[3 * i + 2]);
        TM normal_basis, R;
        bool any_collision = Object::multiObjectCollision(collision_objects, xi, vi, normal_basis, wn);
        collide[i] = any_collision ? 1 : 0;
        slip[i] = 0;
        TV d;
        TM Pm = TM::Zero(), Ri = TM::Zero();
        R = TM::Zero();
        if (any_collision) {
            bool isSlip = wn != TV::Zero();
            if (isSlip) {
                R = rotate(wn);
            }
            else
                R = TM::Identity();
            Pm = TM::Identity() - normal_basis * normal_basis.transpose();
            Ri = R.inverse();
            slip[i] = isSlip ? 1 : 0;
            d = vi - old_v;
        }
        else
            d = g * dt;
        for (int q = 0; q < 3; ++q) dv[3 * i + q] = d(q);
        for (int q = 0; q < 9; ++q) { P[9 * i + q] = Pm(q); R_out[9 * i + q] = R(q); Rinv[9 * i + q] = Ri(q); }
    }
}

// AnalyticCollisionObject::evalMaxSpeed (CollisionObject.cpp:201-238) of every object for the particle box [p_min, p_max] (what calculateDt,
// MpmSimulationBase.cpp:802-804, asks for); NaN where the reference throws (a rotating / scaling object over a level set without bounds: HalfSpace)
void zr_colliders_max_speed(int n_obj, const double* objs, const double* p_min, const double* p_max, double* out)
{
    StdVector<std::unique_ptr<Object>> collision_objects;
    build(n_obj, objs, collision_objects);
    const TV lo(p_min[0], p_min[1], p_min[2]), hi(p_max[0], p_max[1], p_max[2]);
    for (int k = 0; k < n_obj; ++k) {
        try {
            out[k] = collision_objects[k]->evalMaxSpeed(lo, hi);
        }
        catch (...) {
            out[k] = NAN;
        }
    }
}

} // extern "C"
