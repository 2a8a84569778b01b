// Exercises FBasedMpmForceHelperB200 (include/hot_b200_host.hpp), the mirror of MpmForceHelperBase (Lib/MPM/Force/MpmForceHelperBase.h:18-46):
//   force_helper <in.bin> <dt>      in.bin as for host_step.cpp
// prints: strain energy, number of particles visited by runLambdaWithDifferential, sum of dPdF(0,0), sum of Fn(0,0), the same two
// sums for the stored / reused Hessians (opt 1 / 2), the sum of dPdF(F = I)(0,0) from computePerNodeCNTolerance, and whether the visit
// order was the reference's (colour passes over the page groups).
#include "hot_b200_host.hpp"
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
        const char* flags[] = {argv[0], "--project"};
        parseFlags(2, flags);
        MpmSimulationB200 sim(dx);
        sim.dt = std::atof(argv[2]);
        sim.gravity = {0, -9.8, 0};
        sim.setParticles(n, X.data(), V.data(), m.data(), C.data(), F.data(), vol.data(), mu.data(), lam.data());
        sim.sortParticlesAndPolluteGrid();
        sim.particlesToGrid();
        sim.buildMassMatrix();
        sim.buildInitialDvAndVnForNewton();
        FBasedMpmForceHelperB200 helper(sim);
        helper.backupStrain();
        helper.updateImplicitState();
        std::printf("energy %.17g\n", helper.totalEnergy());
        long visited = 0;
        double sh = 0, sf = 0;
        std::vector<int> seq;
        helper.runLambdaWithDifferential([&](int i, const FBasedMpmForceHelperB200::Hessian& H, const TM& Fn, double, double, bool) {
            ++visited; sh += H[0]; sf += Fn[0]; seq.push_back(i);
        }, 1);
        std::printf("visited %ld\nsum_dPdF00 %.17g\nsum_Fn00 %.17g\n", visited, sh, sf);
        double sh2 = 0;
        helper.runLambdaWithDifferential([&](int, const FBasedMpmForceHelperB200::Hessian& H, const TM&, double, double, bool) { sh2 += H[0]; }, 2);
        std::printf("sum_dPdF00_reused %.17g\n", sh2);
        double si = 0;
        helper.computePerNodeCNTolerance([&](int, const FBasedMpmForceHelperB200::Hessian& H, double, bool) { si += H[0]; });
        std::printf("sum_dPdF00_identity %.17g\n", si);
        std::printf("first_visited %d\nlast_visited %d\n", seq.front(), seq.back());
        helper.restoreStrain();
    }
    catch (const std::exception& e) {
        std::fprintf(stderr, "force_helper: %s\n", e.what());
        return 1;
    }
    return 0;
}
