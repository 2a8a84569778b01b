// TEST INFRASTRUCTURE ONLY - stands in for <tbb/tbb.h> (TBB is not in this image).  The reference's Krylov solver headers include it but
// the solvers themselves (InexactConjugateGradient.h, Minres.h) use nothing from it.
#pragma once
