// TEST INFRASTRUCTURE ONLY — not part of the product path.
//
// Row (b), the flag surface: the REFERENCE'S OWN command-line parser (Lib/Ziran/CS/Util/CommandLineFlags.h, Strings.h) and settings object
// (Projects/multigrid/Configurations.h), compiled where they lie, with the flags of Projects/multigrid/main.cpp:40-84 registered the way main() registers
// them (main.cpp itself holds main() and the scene set-up and is not compiled: the FLAGS::Register lines are written out below, same names, same targets).
// zr_flags_parse resets the settings to the defaults they had at load time, runs FLAGS::ParseFlags on a command line and returns the settings - or the
// message of the exception the reference throws.  Built by oracle/Makefile into oracle/_ref/libflags_ref.so; tests/test_flags_ref.py compares
// hot_b200::parseFlags (include/hot_b200_host.hpp) with it on the command lines of Projects/multigrid/tog.sh and on malformed ones.
#include <cstdio>
#include <cstring>
#include <functional>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
#include <Ziran/CS/Util/CommandLineFlags.h>
#include <Configurations.h>

using namespace ZIRAN;

namespace {
bool displayHelp = false, three_d = false, run_diff_test = false, use_double = false;
int test_number = -1, restart = 0, num_threads = -1;
double diff_test_perturbation_scale = 1;
std::string script_file_name;
std::vector<std::string> inline_strings;

// Projects/multigrid/main.cpp:40-84
FLAGS::Register helpflag("--help", "Print help (this message) and exit", displayHelp);
FLAGS::Register scriptflag("-script", "Lua script to read for initial data", script_file_name);
FLAGS::Register iflag("-i", "Append string to script", inline_strings);
FLAGS::Register test_number_flag("-test", "Test number (non-lua test)", test_number);
FLAGS::Register three_d_flag("--3d", "Dimension is 3(non-lua test)", three_d);
FLAGS::Register run_diff_test_flag("--run_diff_test", "Run diff test (non-lua test)", run_diff_test);
FLAGS::Register diff_test_perturbation_scale_flag("-dtps", "diff_test_perturbation_scale (non-lua test)", diff_test_perturbation_scale);
FLAGS::Register double_flag("--double", "Dimension (non-lua test)", use_double);
FLAGS::Register restart_flag("-restart", "Restart frame (non-lua test)", restart);
FLAGS::Register v_mu_flag("-v_mu", "v_mu", CaseSettings::v_mu);
FLAGS::Register output_folder_flag("-o", "output_folder overwrite", CaseSettings::output_folder);
FLAGS::Register dbg_flag("-dbg", "debug mode (0: close 1: open)", HOTSettings::debugMode);
FLAGS::Register cneps_flag("-cneps", "epsilon for characteristic norm", HOTSettings::cneps);
FLAGS::Register usel_baselinemg_flag("--baseline", "if use baseline multigrid", HOTSettings::useBaselineMultigrid);
FLAGS::Register usecn_flag("--usecn", "if use characteristic norm", HOTSettings::useCN);
FLAGS::Register use_adaptiveH_flag("--adaptiveH", "if use adaptive hessian", HOTSettings::useAdaptiveHessian);
FLAGS::Register matrix_flag("--matfree", "if matrix-free", HOTSettings::matrixFree);
FLAGS::Register proj_flag("--project", "if project matrix", HOTSettings::project);
FLAGS::Register bc_proj_flag("--bcproject", "if project boundary on matrix", HOTSettings::systemBCProject);
FLAGS::Register linesearch_flag("--linesearch", "if using linesearch", HOTSettings::linesearch);
FLAGS::Register boundary_flag("-bc", "boundary condition type", HOTSettings::boundaryType);
FLAGS::Register lsolver_flag("-lsolver", "linear solver type", HOTSettings::lsolver);
FLAGS::Register Ainv_flag("-Ainv", "A~ inverse", HOTSettings::Ainv);
FLAGS::Register smoother_flag("-smoother", "smoother type", HOTSettings::smoother);
FLAGS::Register coarseSolver_flag("-coarseSolver", "coarse solver type", HOTSettings::coarseSolver);
FLAGS::Register mglevel_flag("-mg_level", "multigrid level", HOTSettings::levelCnt);
FLAGS::Register mgtimes_flag("-mg_times", "smoother times", HOTSettings::times);
FLAGS::Register levelscale_flag("-mg_scale", "multigrid smoother time scale", HOTSettings::levelscale);
FLAGS::Register omega_flag("-mg_omega", "Gauss Seidel omega", HOTSettings::omega);
FLAGS::Register jomega_flag("-mg_jomega", "Jacobi omega", HOTSettings::topomega);
FLAGS::Register reveal_residual_flag("--showresidual", "multigrid residual time", HOTSettings::revealJacobi);
FLAGS::Register reveal_vcycle_flag("--showvcycle", "multigrid vcycle time", HOTSettings::revealVcycle);
FLAGS::Register topDownMGS_flag("--topDownMGS", "use top down MG solver", HOTSettings::topDownMGS);
FLAGS::Register cmd0_flag("-cmd0", "cmd0", CmdArgument::cmd0);
FLAGS::Register cmd1_flag("-cmd1", "cmd1", CmdArgument::cmd1);
FLAGS::Register thread_flag("-t", "Set number of threads", num_threads);

constexpr int NS = 20;
void snapshot(double* s)
{
    s[0] = HOTSettings::cneps; s[1] = HOTSettings::useAdaptiveHessian; s[2] = HOTSettings::useCN; s[3] = HOTSettings::matrixFree; s[4] = HOTSettings::project;
    s[5] = HOTSettings::systemBCProject; s[6] = HOTSettings::linesearch; s[7] = HOTSettings::boundaryType; s[8] = HOTSettings::lsolver; s[9] = HOTSettings::Ainv;
    s[10] = HOTSettings::smoother; s[11] = HOTSettings::coarseSolver; s[12] = HOTSettings::levelCnt; s[13] = HOTSettings::times; s[14] = HOTSettings::levelscale;
    s[15] = HOTSettings::debugMode; s[16] = HOTSettings::omega; s[17] = HOTSettings::topomega; s[18] = HOTSettings::useBaselineMultigrid; s[19] = HOTSettings::topDownMGS;
}
void restore(const double* s)
{
    HOTSettings::cneps = s[0]; HOTSettings::useAdaptiveHessian = s[1] != 0; HOTSettings::useCN = s[2] != 0; HOTSettings::matrixFree = s[3] != 0;
    HOTSettings::project = s[4] != 0; HOTSettings::systemBCProject = s[5] != 0; HOTSettings::linesearch = s[6] != 0; HOTSettings::boundaryType = (int)s[7];
    HOTSettings::lsolver = (int)s[8]; HOTSettings::Ainv = (int)s[9]; HOTSettings::smoother = (int)s[10]; HOTSettings::coarseSolver = (int)s[11];
    HOTSettings::levelCnt = (int)s[12]; HOTSettings::times = (int)s[13]; HOTSettings::levelscale = (int)s[14]; HOTSettings::debugMode = (int)s[15];
    HOTSettings::omega = s[16]; HOTSettings::topomega = s[17]; HOTSettings::useBaselineMultigrid = s[18] != 0; HOTSettings::topDownMGS = s[19] != 0;
}
struct Defaults {
    double s[NS];
    Defaults() { snapshot(s); }
} defaults; // (Configurations.h's initialisers, read once when the library is loaded)
} // namespace

extern "C" {
// argv[0] is the program name.  out: 20 settings in the order of snapshot().  Returns 0, or 1 with the reference's exception message in err
int zr_flags_parse(int argc, char** argv, double* out, char* err, int err_cap)
{
    restore(defaults.s);
    int rc = 0;
    try {
        FLAGS::ParseFlags(argc, argv);
    }
    catch (std::exception& e) {
        std::snprintf(err, err_cap, "%s", e.what());
        rc = 1;
    }
    snapshot(out);
    return rc;
}
} // extern "C"
