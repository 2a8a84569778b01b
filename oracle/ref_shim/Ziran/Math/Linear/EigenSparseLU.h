// TEST INFRASTRUCTURE ONLY - stands in for Lib/Ziran/Math/Linear/EigenSparseLU.h, which Projects/multigrid/MultigridPreconditioner.h includes without using anything of it
// in the pinned code paths (MultigridBuilder::build, the smoothers, MultigridOperator::operator()).
#pragma once
