// The host collision-object mirror of include/hot_b200_host.hpp (HalfSpace / Sphere / AnalyticBox / CappedCylinder, AnalyticCollisionObject,
// multiObjectCollision, collisionNodeAt) evaluated on a list of points, for the comparison with the reference's own collision-object code
// (oracle/collider_ref_shim.cpp, tests/test_collider_ref.py).  No device call is made.   colliders_ref <in.bin> <out.bin>
// in.bin: int64 n_obj, int64 n, double dt, double gravity[3], n_obj x 33 doubles (layout of oracle/collider_ref_shim.cpp), xi[3n], v[3n]
// out.bin: per point 1 + 3 + 9 + 9 + 9 + 1 doubles: collide, dv, P, R, Rinv, slip
#include "hot_b200_host.hpp"
#include <cstdio>
#include <cstdlib>
#include <string>

using namespace hot_b200;

int main(int argc, char** argv)
{
    if (argc < 3) return 2;
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 3;
    long long n_obj = 0, n = 0;
    double dt = 0, g[3];
    if (std::fread(&n_obj, 8, 1, f) != 1 || std::fread(&n, 8, 1, f) != 1 || std::fread(&dt, 8, 1, f) != 1 || std::fread(g, 8, 3, f) != 3) return 4;
    auto rd = [&](size_t k) { std::vector<double> v(k); if (std::fread(v.data(), 8, k, f) != k) std::exit(5); return v; };
    const std::vector<double> objs = rd(33 * (size_t)n_obj), xi = rd(3 * (size_t)n), v = rd(3 * (size_t)n);
    std::fclose(f);
    std::vector<AnalyticCollisionObject> objects;
    for (long long k = 0; k < n_obj; ++k) {
        const double* o = &objs[33 * (size_t)k];
        const int type = (int)o[0], shape = (int)o[1];
        const double* p = o + 3;
        const std::array<double, 4> sq{o[11], o[12], o[13], o[14]};
        const TV sb{o[15], o[16], o[17]};
        std::shared_ptr<AnalyticLevelSet> ls;
        if (shape == 0) ls = std::make_shared<HalfSpace>(TV{p[0], p[1], p[2]}, TV{p[3], p[4], p[5]});
        else if (shape == 1) ls = std::make_shared<Sphere>(TV{p[0], p[1], p[2]}, p[3]);
        else if (shape == 2) ls = std::make_shared<AnalyticBox>(TV{p[0], p[1], p[2]}, sq, sb);
        else if (shape == 3) ls = std::make_shared<CappedCylinder>(p[0], p[1], sq, sb);
        else ls = std::make_shared<AnalyticBox>(AnalyticBox::axisAligned({p[0], p[1], p[2]}, {p[3], p[4], p[5]}));
        objects.emplace_back(ls, (COLLISION_OBJECT_TYPE)type);
        AnalyticCollisionObject& obj = objects.back();
        obj.friction = o[2];
        obj.setRotation({o[18], o[19], o[20], o[21]});
        obj.s = o[22]; obj.dsdt = o[29];
        obj.setTranslation({o[23], o[24], o[25]}, {o[30], o[31], o[32]});
        obj.setAngularVelocity({o[26], o[27], o[28]});
    }
    FILE* out = std::fopen(argv[2], "wb");
    if (!out) return 6;
    if (argc > 3 && std::string(argv[3]) == "speed") { // evalMaxSpeed of every object for the box [xi[0], xi[1]]; NaN where the mirror throws
        const TV lo{xi[0], xi[1], xi[2]}, hi{xi[3], xi[4], xi[5]};
        for (const auto& o : objects) {
            double sp;
            try { sp = o.evalMaxSpeed(lo, hi); }
            catch (const HotError&) { sp = std::nan(""); }
            std::fwrite(&sp, 8, 1, out);
        }
        // calculateDt of a particle set = the two box corners moving with v[0], v[1], among these objects (cfl 0.6, dx 1/32, max_dt 1e-2): skipping the
        // objects the reference would throw on
        std::vector<AnalyticCollisionObject> ok;
        for (const auto& o : objects) { try { (void)o.evalMaxSpeed(lo, hi); ok.push_back(o); } catch (const HotError&) {} }
        const double dtc = calculateDtFromArrays(2, xi.data(), v.data(), 1.0 / 32, 0.6, 1e-2, ok);
        std::fwrite(&dtc, 8, 1, out);
        std::fclose(out);
        return 0;
    }
    for (long long i = 0; i < n; ++i) {
        const TV x{xi[3 * i], xi[3 * i + 1], xi[3 * i + 2]}, old_v{v[3 * i], v[3 * i + 1], v[3 * i + 2]};
        CollisionNode Z;
        Z.P.fill(0.0); Z.R.fill(0.0); Z.Rinv.fill(0.0); Z.shouldRotate = false;
        TV dv{g[0] * dt, g[1] * dt, g[2] * dt}; // Newton initial guess of a free node
        const bool hit = collisionNodeAt(objects, x, old_v, (int)i, Z, dv);
        double rec[32];
        rec[0] = hit ? 1.0 : 0.0;
        for (int d = 0; d < 3; ++d) rec[1 + d] = dv[d];
        for (int q = 0; q < 9; ++q) { rec[4 + q] = Z.P[q]; rec[13 + q] = Z.R[q]; rec[22 + q] = Z.Rinv[q]; }
        rec[31] = Z.shouldRotate ? 1.0 : 0.0;
        std::fwrite(rec, 8, 32, out);
    }
    std::fclose(out);
    return 0;
}
