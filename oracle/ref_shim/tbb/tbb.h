// TEST INFRASTRUCTURE ONLY - stands in for <tbb/tbb.h> (TBB is not in this image).  The reference's solver headers include it; the pinned
// code paths (InexactConjugateGradient::solve, Minres::solve, LBFGS::solve) use nothing from it, a debugging helper of LBFGS.h names
// tbb::parallel_for, which runs serially here.
#pragma once
namespace tbb {
template <class Index, class F>
inline void parallel_for(Index first, Index last, const F& f) { for (Index i = first; i < last; ++i) f(i); }
} // namespace tbb
