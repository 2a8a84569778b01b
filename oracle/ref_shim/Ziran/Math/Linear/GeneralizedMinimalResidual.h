// TEST INFRASTRUCTURE ONLY - stands in for Lib/Ziran/Math/Linear/GeneralizedMinimalResidual.h.  Projects/multigrid/MultigridPreconditioner.h includes it without using
// anything of it; Projects/multigrid/ImplicitSolver.h holds a GMRES member that only its full_implicit branch calls, which none of the pinned code paths
// takes (implicit_ref_shim.cpp) - so the class below has the members the constructor touches and a solve() that refuses to run.
#pragma once
#include <stdexcept>
namespace ZIRAN {
template <class T, class TM, class TV>
class GeneralizedMinimalResidual {
public:
    GeneralizedMinimalResidual(const int) {}
    void setTolerance(T) {}
    void setRelativeTolerance(T) {}
    int solve(const TM&, TV&, const TV&, const bool = false) { throw std::runtime_error("GMRES stand-in: not part of the pinned code paths"); }
};
} // namespace ZIRAN
