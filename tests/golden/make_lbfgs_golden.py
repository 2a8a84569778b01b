"""Generates tests/golden/lbfgs_ref.npz: whole implicit substeps whose L-BFGS loop is the REFERENCE'S OWN ZIRAN::LBFGS<Objective>::solve
(Lib/Ziran/Math/Nonlinear/LBFGS.h:300-437, compiled where it lies into oracle/_ref/libhot_oracle_lbfgsref.so, oracle/lbfgs_ref_shim.cpp)
driven on the oracle's objective (updateState / computeResidual / shouldExitByCN / HinvApproxInit / precondition / project /
lineSearch / recoverSolution / transformResidual = the oracle's functions of the same rows).  tests/test_oracle_lbfgs_ref.py compares
the oracle's restatement of the loop (oracle_solver.inl: lbfgs_solve, row a22) and the CUDA solver (hot_backward_euler_step) with
these iteration counts, residual histories and velocity increments.
Run in the build container (needs /root/reference for `make -C oracle ref`):  python tests/golden/make_lbfgs_golden.py"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libhot_oracle_lbfgsref.so")

HOT = dict(lsolver=3, mg_level=3, smoother=5, coarse_solver=2, project=1, linesearch=1, bcproject=1, usecn=1)   # tog.sh:38
# (name, scene arguments, solver options)
CASES = [
    ("hot_default", dict(cells=(6, 6, 6), E=4e5, dt=5e-3, seed=0), dict(HOT)),
    ("hot_no_linesearch", dict(cells=(6, 6, 6), E=4e5, dt=5e-3, seed=0), dict(HOT, linesearch=0)),
    ("hot_stiff_long", dict(cells=(7, 9, 6), E=2e6, dt=8e-3, seed=1), dict(HOT, cneps=1e-9)),          # more iterations than the 8-deep history
    ("hot_adaptive_hessian", dict(cells=(7, 9, 6), E=2e6, dt=8e-3, seed=1), dict(HOT, cneps=1e-9, adaptive_h=1)),
    ("hot_no_cn_iteration_cap", dict(cells=(6, 5, 6), E=3e5, dt=5e-3, seed=2), dict(HOT, usecn=0, cneps=1e-14, max_lbfgs_iterations=11)),
]


def reference_binding():
    """the oracle binding bound to the library that also carries the reference's LBFGS loop"""
    os.environ["HOT_ORACLE_LIB"] = REF_LIB
    try:
        spec = importlib.util.spec_from_file_location("oracle_binding_lbfgsref", os.path.join(ROOT, "tests", "oracle_binding.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        del os.environ["HOT_ORACLE_LIB"]
    return mod


def scene(make_sim, cells, E, dt, seed):
    """a seeded block with a perturbed deformation gradient resting on a sticky floor, ready for backwardEulerStep"""
    from hot_b200 import scenes
    sc = scenes.block(cells, 1.0 / 32, ppc=6, seed=seed, E=E)
    rng = np.random.default_rng(seed)
    sc["F"] = sc["F"] + 0.08 * (rng.random(sc["F"].shape) - 0.5)
    s = make_sim(sc["dx"])
    s.set_particles(sc["X"], sc["V"], sc["mass"], sc["C"], sc["F"], sc["vol"], sc["mu"], sc["lam"])
    s.set_dt_gravity(dt, (0.0, -9.8, 0.0))
    s.sortParticlesAndPolluteGrid(); s.particlesToGrid()
    coord = s.get_id2coord()
    bc = np.nonzero(coord[:, 1] <= coord[:, 1].min() + 1)[0].astype(np.int32)
    s.set_bc(bc, P=np.zeros((len(bc), 9)), dv_bc=np.zeros((len(bc), 3)))
    return s


def main():
    ref = reference_binding()
    out = {}
    for name, sc_args, opts in CASES:
        s = scene(ref.OracleSim, **sc_args)
        log = s.backwardEulerStepReferenceLBFGS(**opts)
        dv = s.get_dv()
        out[name + "_iterations"] = np.int64(log["iterations"])
        out[name + "_converged"] = np.int64(log["converged"])
        out[name + "_vcycles"] = np.int64(log["total_linear_iterations"])
        out[name + "_matrix_builds"] = np.int64(log["matrix_builds"])
        out[name + "_residual_norm"] = np.asarray(log["residual_norm"], dtype=np.float64)
        out[name + "_dv"] = dv
        print(name, "iterations", log["iterations"], "converged", log["converged"], "V-cycles", log["total_linear_iterations"],
              "matrix builds", log["matrix_builds"], "residuals", np.asarray(log["residual_norm"])[:3], "...", np.asarray(log["residual_norm"])[-1])
        s.close()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "lbfgs_ref.npz"), **out)


if __name__ == "__main__":
    main()
