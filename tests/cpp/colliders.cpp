// Device evaluation of the collision objects (hot_set_colliders + hot_build_bc, a8 / SURVEY 8f rank 3) against the host evaluation of
// include/hot_b200_host.hpp (CollisionObject.cpp:108-149,384-452 restated) on the same grid: same collision nodes, P, R, R^-1, slip
// flags and Newton initial guess.   colliders <in.bin> <dt>     (in.bin as for host_step.cpp)
#include "hot_b200_host.hpp"
#include <algorithm>
#include <cstdio>
#include <cstdlib>

using namespace hot_b200;

int main(int argc, char** argv)
{
    if (argc < 3) return 2;
    try {
        FILE* f = std::fopen(argv[1], "rb");
        if (!f) throw HotError("cannot open input");
        long long n; double dx;
        if (std::fread(&n, 8, 1, f) != 1 || std::fread(&dx, 8, 1, f) != 1) throw HotError("bad header");
        auto rd = [&](size_t k) { std::vector<double> v(k); if (std::fread(v.data(), 8, k, f) != k) throw HotError("short read"); return v; };
        auto X = rd(3 * n), V = rd(3 * n), m = rd(n), C = rd(9 * n), F = rd(9 * n), vol = rd(n), mu = rd(n), lam = rd(n);
        std::fclose(f);
        MpmSimulationB200 sim(dx);
        sim.dt = std::atof(argv[2]);
        sim.gravity = {0, -9.8, 0};
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (long long i = 0; i < n; ++i)
            for (int d = 0; d < 3; ++d) { lo[d] = std::min(lo[d], X[3 * i + d]); hi[d] = std::max(hi[d], X[3 * i + d]); }
        const TV c{(lo[0] + hi[0]) / 2, (lo[1] + hi[1]) / 2, (lo[2] + hi[2]) / 2};
        const double L = hi[1] - lo[1];
        // one object of every shape and type around the block: tilted SLIP ground, moving STICKY sphere, rotating STICKY capped cylinder
        // at the top (the twisting bar's clamp, MultigridInit3D.h:628-661), SEPARATE box with friction poking into a side, a GHOST
        sim.collision_objects.emplace_back(std::make_shared<HalfSpace>(TV{0, lo[1] + 0.12 * L, 0}, TV{0.15, 1.0, -0.1}), SLIP);
        sim.collision_objects.back().friction = 0.3;
        sim.collision_objects.emplace_back(std::make_shared<Sphere>(TV{0, 0, 0}, 0.3 * L), STICKY);
        sim.collision_objects.back().setTranslation({lo[0], c[1], c[2]}, {0.2, -0.1, 0.05});
        sim.collision_objects.emplace_back(std::make_shared<CappedCylinder>(0.6 * L, 0.25 * L, std::array<double, 4>{1, 0, 0, 0}, TV{0, 0, 0}), STICKY);
        sim.collision_objects.back().setRotation({std::cos(0.35), 0, std::sin(0.35), 0});
        sim.collision_objects.back().setAngularVelocity({0, 2 * M_PI, 0});
        sim.collision_objects.back().setTranslation({c[0], hi[1], c[2]}, {0, 0.3, 0});
        sim.collision_objects.emplace_back(std::make_shared<AnalyticBox>(TV{0.2 * L, 0.15 * L, 0.5 * L}, std::array<double, 4>{0.9, 0.1, 0.3, -0.2}, TV{hi[0], c[1], c[2]}), SEPARATE);
        sim.collision_objects.back().friction = 0.5;
        sim.collision_objects.emplace_back(std::make_shared<AnalyticBox>(AnalyticBox::axisAligned({lo[0], lo[1], hi[2] - 0.1 * L}, {hi[0], lo[1] + 0.3 * L, hi[2] + L})), SLIP);
        sim.collision_objects.emplace_back(std::make_shared<Sphere>(c, 10 * L), GHOST);
        sim.setParticles(n, X.data(), V.data(), m.data(), C.data(), F.data(), vol.data(), mu.data(), lam.data());
        sim.sortParticlesAndPolluteGrid();
        sim.particlesToGrid();
        sim.buildMassMatrix();

        sim.device_colliders = false;
        sim.buildInitialDvAndVnForNewton();
        std::vector<CollisionNode> host_nodes = sim.collision_nodes;
        TVStack host_dv = sim.dv;
        std::sort(host_nodes.begin(), host_nodes.end(), [](const CollisionNode& a, const CollisionNode& b) { return a.node_id < b.node_id; });

        sim.device_colliders = true;
        sim.buildInitialDvAndVnForNewton();
        sim.fetchCollisionNodes();
        TVStack dev_dv(3 * (size_t)sim.num_nodes);
        sim.check(hot_get_dv(sim.handle(), dev_dv.data()));

        std::printf("nodes %d\nhost_bc %zu\ndevice_bc %d\n", sim.num_nodes, host_nodes.size(), sim.num_collision_nodes);
        if (host_nodes.size() != sim.collision_nodes.size()) { std::printf("match 0\n"); return 0; }
        double eP = 0, eR = 0, eI = 0, edv = 0;
        int bad_id = 0, bad_slip = 0, n_slip = 0;
        for (size_t k = 0; k < host_nodes.size(); ++k) {
            const CollisionNode &a = host_nodes[k], &b = sim.collision_nodes[k];
            bad_id += a.node_id != b.node_id;
            bad_slip += a.shouldRotate != b.shouldRotate;
            n_slip += a.shouldRotate;
            for (int q = 0; q < 9; ++q) {
                eP = std::max(eP, std::fabs(a.P[q] - b.P[q]));
                eR = std::max(eR, std::fabs(a.R[q] - b.R[q]));
                eI = std::max(eI, std::fabs(a.Rinv[q] - b.Rinv[q]));
            }
        }
        double sdv = 0;
        for (size_t q = 0; q < host_dv.size(); ++q) { edv = std::max(edv, std::fabs(host_dv[q] - dev_dv[q])); sdv = std::max(sdv, std::fabs(host_dv[q])); }
        std::printf("bad_id %d\nbad_slip %d\nslip_nodes %d\nerr_P %.3e\nerr_R %.3e\nerr_Rinv %.3e\nerr_dv %.3e\nscale_dv %.3e\nmatch 1\n", bad_id, bad_slip, n_slip, eP, eR, eI, edv, sdv);
    }
    catch (const std::exception& e) {
        std::fprintf(stderr, "colliders: %s\n", e.what());
        return 1;
    }
    return 0;
}
