// hot_b200::parseFlags (include/hot_b200_host.hpp) on the command line given to this program: prints the 20 settings in the order of
// oracle/flags_ref_shim.cpp, or "error <message>" with exit status 1.  No device call is made.  (tests/test_flags_ref.py)
#include "hot_b200_host.hpp"
#include <cstdio>

int main(int argc, char** argv)
{
    using namespace hot_b200;
    try {
        parseFlags(argc, argv);
    }
    catch (const std::exception& e) {
        std::printf("error %s\n", e.what());
        return 1;
    }
    using namespace HOTSettings;
    std::printf("%.17g %d %d %d %d %d %d %d %d %d %d %d %d %d %d %d %.17g %.17g %d %d\n", cneps, (int)useAdaptiveHessian, (int)useCN, (int)matrixFree, (int)project,
        (int)systemBCProject, (int)linesearch, boundaryType, lsolver, Ainv, smoother, coarseSolver, levelCnt, times, levelscale, debugMode, omega, topomega,
        (int)useBaselineMultigrid, (int)topDownMGS);
    return 0;
}
