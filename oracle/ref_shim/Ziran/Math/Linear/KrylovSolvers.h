// TEST INFRASTRUCTURE ONLY - stands in for Lib/Ziran/Math/Linear/KrylovSolvers.h, whose helper classes (Jacobi preconditioner,
// Dirichlet projection, Eigen sparse-matrix adapter) need Eigen's sparse module.  The solvers under test
// (InexactConjugateGradient.h, Minres.h, LinearSolver.h) only need what it includes: the Matrix / Vector aliases, GivensRotation and
// the logging / timer macros.
#pragma once
#include <Ziran/Math/Linear/DenseExt.h>
#include <Ziran/Math/Linear/Givens.h>
#include <Ziran/CS/Util/Timer.h>
#include <Ziran/CS/Util/Logging.h>
#include <Eigen/Core>
