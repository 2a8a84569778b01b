// Drives the C++ host mirror (include/hot_b200_host.hpp) the way Projects/multigrid/main.cpp drives MultigridSimulation:
// flags -> HOTSettings, scene (particles + collision objects), advanceOneTimeStep x n.  Used by tests/test_gpu_host_cpp.py.
//   host_step <in.bin> <out.bin> <n_steps> <dt> <ground_y> <ground_type 1|2> [reference flags ...]
// in.bin : i64 n, f64 dx, then X[3n] V[3n] m[n] C[9n] F[9n] vol[n] mu[n] lam[n]
// out.bin: i64 n, X V C F, then per step {i32 iterations, i32 converged, i32 n_nodes, i32 n_bc, f64 last residual norm}
#include <fstream>
#include "hot_b200_host.hpp"
#include <cstdio>
#include <cstdlib>

using namespace hot_b200;

int main(int argc, char** argv)
{
    if (argc < 7) { std::fprintf(stderr, "usage\n"); return 2; }
    try {
        FILE* f = std::fopen(argv[1], "rb");
        if (!f) throw HotError("cannot open input");
        long long n; double dx;
        if (std::fread(&n, 8, 1, f) != 1 || std::fread(&dx, 8, 1, f) != 1) throw HotError("bad header");
        auto rd = [&](size_t k) { std::vector<double> v(k); if (std::fread(v.data(), 8, k, f) != k) throw HotError("short read"); return v; };
        auto X = rd(3 * n), V = rd(3 * n), m = rd(n), C = rd(9 * n), F = rd(9 * n), vol = rd(n), mu = rd(n), lam = rd(n);
        std::fclose(f);
        const int steps = std::atoi(argv[3]);
        const double dt = std::atof(argv[4]), ground = std::atof(argv[5]);
        const int gtype = std::atoi(argv[6]);
        std::vector<const char*> flags{argv[0]};
        for (int i = 7; i < argc; ++i) flags.push_back(argv[i]);
        parseFlags((int)flags.size(), flags.data());

        MpmSimulationB200 sim(dx);
        sim.device_colliders = std::getenv("HOT_HOST_COLLIDERS") == nullptr; // a8 on the device (default) or on the host
        sim.gravity = {0, -9.8, 0};
        sim.collision_objects.emplace_back(std::make_shared<HalfSpace>(TV{0, ground, 0}, TV{0, 1, 0}), (COLLISION_OBJECT_TYPE)gtype);
        sim.setParticles(n, X.data(), V.data(), m.data(), C.data(), F.data(), vol.data(), mu.data(), lam.data());
        FILE* o = std::fopen(argv[2], "wb");
        if (!o) throw HotError("cannot open output");
        std::vector<double> rec;
        for (int s = 0; s < steps; ++s) {
            sim.advanceOneTimeStep(dt);
            const hot_solve_log& L = sim.last_log;
            rec.push_back(L.iterations); rec.push_back(L.converged); rec.push_back(sim.num_nodes); rec.push_back((double)sim.num_collision_nodes);
            rec.push_back(L.n_log ? L.residual_norm[L.n_log - 1] : 0.0);
        }
        sim.getParticles(X.data(), V.data(), C.data(), F.data());
        std::fwrite(&n, 8, 1, o);
        std::fwrite(X.data(), 8, X.size(), o); std::fwrite(V.data(), 8, V.size(), o);
        std::fwrite(C.data(), 8, C.size(), o); std::fwrite(F.data(), 8, F.size(), o);
        std::fwrite(rec.data(), 8, rec.size(), o);
        std::fclose(o);
        // the -smoother function-pointer surface stays callable after the step (hierarchy of the last solve)
        if (HOTSettings::lsolver == 3 || !HOTSettings::matrixFree) {
            MPMSpMatB200 A{&sim, 0};
            TVStack u(3 * (size_t)sim.num_nodes, 0.0), r(3 * (size_t)sim.num_nodes, 1.0), du, dAu;
            selectSmoother(HOTSettings::smoother)(u, r, du, dAu, A, 1, 0.0);
        }
        if (const char* rf = std::getenv("HOT_RESTART_FILE")) { // writeState -> readState into a second simulation: identical particle state
            {
                std::ofstream os(rf, std::ios::binary);
                sim.writeState(os);
            }
            MpmSimulationB200 sim2(dx);
            std::ifstream is(rf, std::ios::binary);
            sim2.readState(is);
            std::vector<double> X2(3 * n), V2(3 * n), C2(9 * n), F2(9 * n);
            sim2.getParticles(X2.data(), V2.data(), C2.data(), F2.data());
            if (sim2.particleCount() != n || X2 != X || V2 != V || C2 != C || F2 != F || sim2.mass_p != m || sim2.vol_p != vol)
                throw HotError("restart round trip changed the particle state");
            std::printf("restart ok\n");
        }
        std::printf("ok %lld particles, %d nodes, %d bc nodes, %d iterations\n", n, sim.num_nodes, sim.num_collision_nodes, sim.last_log.iterations);
    }
    catch (const std::exception& e) {
        std::fprintf(stderr, "host_step: %s\n", e.what());
        return 1;
    }
    return 0;
}
