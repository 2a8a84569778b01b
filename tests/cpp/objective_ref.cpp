// hot_b200::shouldExitByCN (include/hot_b200_host.hpp, the exit test ImplicitSolverObjectiveB200 uses) on host arrays, for the comparison with the
// reference's own ImplicitSolverObjective::shouldExitByCN (tests/test_oracle_implicit_ref.py).  No device call is made.
//   objective_ref <in.bin>      in.bin: int64 n, int64 useCN, double cneps, double scale, residual[3n], nodeCNTol[n]   -> prints 0 / 1 for scale * residual
#include "hot_b200_host.hpp"
#include <cstdio>

int main(int argc, char** argv)
{
    if (argc < 2) return 2;
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 3;
    long long n = 0, useCN = 0;
    double cneps = 0, scale = 1;
    if (std::fread(&n, 8, 1, f) != 1 || std::fread(&useCN, 8, 1, f) != 1 || std::fread(&cneps, 8, 1, f) != 1 || std::fread(&scale, 8, 1, f) != 1) return 4;
    hot_b200::TVStack r(3 * (size_t)n);
    std::vector<double> tol((size_t)n);
    if (std::fread(r.data(), 8, r.size(), f) != r.size() || std::fread(tol.data(), 8, tol.size(), f) != tol.size()) return 5;
    std::fclose(f);
    for (double& x : r) x *= scale;
    hot_b200::HOTSettings::useCN = useCN != 0;
    hot_b200::HOTSettings::cneps = cneps;
    std::printf("%d\n", hot_b200::shouldExitByCN(r, tol, (int)n) ? 1 : 0);
    return 0;
}
