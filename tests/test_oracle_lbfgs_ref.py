"""Pins the oracle's restatement of row a22 (the L-BFGS loop of the HOT solver) to the REFERENCE'S OWN code: tests/golden/lbfgs_ref.npz holds
implicit substeps whose loop is ZIRAN::LBFGS<Objective>::solve (Lib/Ziran/Math/Nonlinear/LBFGS.h:300-437), compiled where it lies and driven
on the oracle's objective (oracle/lbfgs_ref_shim.cpp, tests/golden/make_lbfgs_golden.py).  The oracle's lbfgs_solve - and the CUDA solver
through the C ABI - must take the same number of iterations, V-cycles and matrix builds, follow the same residual history and end at the
same velocity increment: ring-buffer history (8 deep, the long cases wrap it), two-loop recursion, rebuild schedule
(--adaptiveH: every 16 iterations), iteration cap, line search on / off."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("make_lbfgs_golden", os.path.join(ROOT, "tests", "golden", "make_lbfgs_golden.py"))
gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gen)
G = np.load(os.path.join(ROOT, "tests", "golden", "lbfgs_ref.npz"))
IDS = [c[0] for c in gen.CASES]


def _check(name, log, dv, rtol_hist, rtol_dv):
    assert log["iterations"] == int(G[name + "_iterations"]), f"{name}: {log['iterations']} iterations, the reference's loop takes {int(G[name + '_iterations'])}"
    assert bool(log["converged"]) == bool(G[name + "_converged"])
    assert log["total_linear_iterations"] == int(G[name + "_vcycles"])
    assert log["matrix_builds"] == int(G[name + "_matrix_builds"])
    h, hr = np.asarray(log["residual_norm"]), G[name + "_residual_norm"]
    assert len(h) == len(hr)
    assert np.abs(h - hr).max() <= rtol_hist * hr[0] + 1e-3 * np.abs(hr).min()   # (the last entries sit at the convergence threshold)
    assert np.abs(dv - G[name + "_dv"]).max() <= rtol_dv * np.abs(G[name + "_dv"]).max()


@pytest.mark.parametrize("case", gen.CASES, ids=IDS)
def test_oracle_lbfgs_against_reference_code(oracle, case):
    name, sc_args, opts = case
    s = gen.scene(oracle.OracleSim, **sc_args)
    log = s.backwardEulerStep(**opts)
    # same objective, same loop: rounding differs only in the dot products (array order on the reference side, fixed tree in the oracle)
    _check(name, log, s.get_dv(), 1e-9, 1e-7)
    s.close()


@pytest.mark.skipif(not os.path.exists(gen.REF_LIB), reason="oracle/_ref/libhot_oracle_lbfgsref.so not built (needs /root/reference)")
def test_reference_loop_reproduces_the_golden_vectors():
    ref = gen.reference_binding()
    name, sc_args, opts = gen.CASES[2]
    s = gen.scene(ref.OracleSim, **sc_args)
    log = s.backwardEulerStepReferenceLBFGS(**opts)
    _check(name, log, s.get_dv(), 1e-13, 1e-12)
    s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("case", gen.CASES, ids=IDS)
def test_cuda_lbfgs_against_reference_code(hot, case):
    """hot_backward_euler_step (solver.cu + the whole device path under it) against substeps driven by the reference's LBFGS loop"""
    name, sc_args, opts = case
    s = gen.scene(hot.MpmSimulationB200, **sc_args)
    log = s.backwardEulerStep(**opts)
    _check(name, log, s.get_dv(), 1e-5, 1e-5)
    s.close()
