"""Pins rows a9-a15 (v / grad v gather + F update, fixed-corotated energy and stress, force scatter, matrix-free Hessian apply, residual, CN
tolerance, line-search energy, buildMatrix with and without BC projection, buildDiagonal) to the REFERENCE'S OWN implicit-solver objective:
tests/golden/implicit_ref.npz was produced by ZIRAN::ImplicitSolverObjective<Simulation> (Projects/multigrid/ImplicitSolver.h), compiled where
it lies and instantiated on a stand-in simulation over the reference's MpmGrid / SPGrid grid code and its CorotatedIsotropic model
(oracle/implicit_ref_shim.cpp -> oracle/_ref/libimplicit_ref.so; tests/golden/make_implicit_golden.py).  The oracle's restatement and the CUDA
path through the C ABI must reproduce them; tolerances (relative to the largest magnitude of each field) are in TOL below - the sums run in
another order than the reference's serial loops."""
import importlib.util
import os

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("make_implicit_golden", os.path.join(ROOT, "tests", "golden", "make_implicit_golden.py"))
gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gen)
G = np.load(os.path.join(ROOT, "tests", "golden", "implicit_ref.npz"))
TOL = dict(energy=1e-11, F=1e-13, residual=1e-11, matrix=1e-11, Ax=1e-10, diag=1e-10, cntol=1e-12)


def _close(a, b, tol, what):
    a = np.asarray(a); b = np.asarray(b)
    assert a.shape == b.shape, what
    err = np.abs(a - b).max()
    assert err <= tol * max(np.abs(b).max(), 1e-300), (what, err, np.abs(b).max())


def _csr(col, val, n):
    blocks = val.reshape(-1, 3, 3).transpose(0, 2, 1)            # column-major 3 x 3 blocks
    A = sp.bsr_matrix((blocks, col.reshape(-1), np.arange(0, n * 125 + 1, 125)), shape=(3 * n, 3 * n)).tocsr()
    A.sum_duplicates()
    return A


def _check(make_sim, name, exact_slots):
    g = lambda k: G[f"{name}/{k}"]
    dt, mode, project = float(g("dt")), int(g("mode")), bool(g("project"))
    s = make_sim(float(g("dx")))
    s.set_particles(*[g("in_" + k) for k in ("X", "V", "mass", "C", "F", "vol", "mu", "lam")])
    s.set_dt_gravity(dt, gen.GRAVITY)
    s.set_project(project)
    s.sortParticlesAndPolluteGrid()
    n = s.particlesToGrid()
    assert n == int(g("num_nodes"))
    node, slip = g("bc_node"), g("bc_slip")
    s.backupStrain()
    if mode == 1:
        s.set_bc(node, R=g("bc_R"), Rinv=g("bc_Rinv"), slip=slip, dv_bc=np.zeros((len(node), 3)), mode=1)
    else:
        s.set_bc(node, P=g("bc_P"), dv_bc=np.zeros((len(node), 3)), mode=0)
    # a9-a11 + a14: moveNodes, grad v gather, F = (I + dt grad v) Fn, energy (elastic + kinetic - gravity work)
    _close(s.updateState(g("dv")), g("energy"), TOL["energy"], "energy")
    _close(s.get_stress()[1], g("F"), TOL["F"], "F")
    # a12 + a14: dt g m + dt f - m dv, rotated / projected at the collision nodes
    _close(s.computeResidual(), g("residual"), TOL["residual"], "residual")
    # a13: matrix-free Hessian apply
    _close(s.multiply(g("x_mf")), g("Ax_mf"), TOL["Ax"], "Ax_mf")
    # a15: block rows without / with BC projection
    for bc in (0, 1):
        s.buildMatrix(bool(bc))
        col, val = s.get_matrix()
        if exact_slots:
            assert np.array_equal(col, g(f"col{bc}")), f"col{bc}"        # slot layout incl. the "empty slot" column convention (:597-602)
        _close(val.sum(1), g(f"valsum{bc}"), TOL["matrix"], f"valsum{bc}")
        if name in gen.FULL_MATRIX:
            A, B = _csr(col, val, n), _csr(g(f"col{bc}"), g(f"val{bc}"), n)
            assert abs(A - B).max() <= TOL["matrix"] * abs(B).max(), f"val{bc}"
        _close(s.spmv(0, g(f"x{bc}")), g(f"Ax{bc}"), TOL["Ax"], f"Ax{bc}")
    for a in (0, 1):
        _close(s.buildDiagonal(a), g(f"diag{a}"), TOL["diag"], f"diag{a}")
    _close(s.evaluatePerNodeCNTolerance(gen.CN_EPS, dt), g("cntol"), TOL["cntol"], "cntol")


@pytest.mark.parametrize("name", list(gen.CASES))
def test_oracle_objective_against_reference_code(oracle, name):
    _check(oracle.OracleSim, name, exact_slots=True)


def test_golden_exit_tests_bracket_the_threshold():
    """shouldExitByCN (ImplicitSolver.h:171-215) on a residual scaled just below / above the threshold, with and without the CN scaling"""
    for name in gen.CASES:
        assert list(G[f"{name}/exit"]) == [1, 0, 1, 0]
        r, tol, n = G[f"{name}/residual"], G[f"{name}/cntol"], int(G[f"{name}/num_nodes"])
        f = float(G[f"{name}/exit_scale"])
        scaled = ((f * r) ** 2).sum(1) / tol ** 2                  # the rule the oracle's and the library's solvers apply (oracle_solver.inl:33-50)
        assert abs(scaled.sum() / n - 1.0) < 1e-9


def test_host_exit_test_against_reference_code(tmp_path):
    """hot_b200::shouldExitByCN (the exit test of ImplicitSolverObjectiveB200, include/hot_b200_host.hpp; driver tests/cpp/objective_ref.cpp) takes the
    decisions of the reference's ImplicitSolverObjective::shouldExitByCN on residuals scaled just below / above the threshold, with and without --usecn"""
    import struct
    import subprocess
    exe = str(tmp_path / "objective_ref")
    lib = os.path.join(ROOT, "hot_b200", "lib")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "objective_ref.cpp"),
                           "-o", exe, "-L", lib, "-lhot_b200", f"-Wl,-rpath,{lib}", "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"])
    for name in gen.CASES:
        r, tol, f = G[f"{name}/residual"], G[f"{name}/cntol"], float(G[f"{name}/exit_scale"])
        l2 = float(np.sqrt((r ** 2).sum()))
        calls = [(1, 0.0, 0.99 * f), (1, 0.0, 1.01 * f), (0, 1.01 * l2, 1.0), (0, 0.99 * l2, 1.0)]      # (useCN, cneps, scale): make_implicit_golden.py
        got = []
        for k, (usecn, cneps, scale) in enumerate(calls):
            inp = str(tmp_path / f"{name}{k}.bin")
            with open(inp, "wb") as fh:
                fh.write(struct.pack("<qqdd", len(tol), usecn, cneps, scale))
                fh.write(np.ascontiguousarray(r, dtype=np.float64).tobytes()); fh.write(np.ascontiguousarray(tol, dtype=np.float64).tobytes())
            got.append(int(subprocess.check_output([exe, inp], text=True).strip()))
        assert got == [int(x) for x in G[f"{name}/exit"]], (name, got)


def _check_step(make_sim, name, rtol_dv):
    """a whole implicit solve (row a24 + the Newton / L-BFGS outer loops a22): same number of nonlinear iterations, same tolerance, same dv"""
    sc_args, opts = gen.STEPS[name]
    s, _, _ = gen.step_scene(make_sim, **sc_args)
    log = s.backwardEulerStep(**opts)
    g = lambda k: G[f"step/{name}/{k}"]
    assert log["iterations"] == int(g("iterations")), (name, log["iterations"], int(g("iterations")))
    assert bool(log["converged"]) == bool(g("converged"))
    if opts.get("usecn", 1):
        assert abs(log["tolerance"] - float(g("tolerance"))) <= 1e-12 * float(g("tolerance"))
    dv = s.get_dv()
    assert np.abs(dv - g("dv")).max() <= rtol_dv * np.abs(g("dv")).max(), (name, np.abs(dv - g("dv")).max() / np.abs(g("dv")).max())
    s.close()


@pytest.mark.parametrize("name", list(gen.STEPS))
def test_oracle_whole_solve_against_reference_code(oracle, name):
    """the oracle's backwardEulerStep against MultigridSimulation::backwardEulerStep's sequence run in the reference's ExtendedNewtonsMethod / LBFGS /
    ImplicitSolverObjective / InexactConjugateGradient / MultigridBuilder / MultigridOperator code"""
    _check_step(oracle.OracleSim, name, 1e-9)


def test_whole_solve_golden_agrees_with_the_lbfgs_golden():
    """the same substeps with the reference's L-BFGS loop on the ORACLE'S objective (lbfgs_ref.npz) and on the REFERENCE'S objective (here)"""
    L = np.load(os.path.join(ROOT, "tests", "golden", "lbfgs_ref.npz"))
    for name in ("hot_default", "hot_no_linesearch", "hot_stiff_long"):
        assert int(L[name + "_iterations"]) == int(G[f"step/{name}/iterations"])
        assert np.abs(L[name + "_dv"] - G[f"step/{name}/dv"]).max() <= 1e-9 * np.abs(L[name + "_dv"]).max()


def _check_run(make_sim, name, rtol):
    """three whole time steps (sort, P2G, implicit solve, G2P, evolveStrain, return mapping) with the state carried through the re-sorts: the same
    nonlinear iterations in every step, the same particle state at the end, the same plastic history after every step"""
    sc_args, opts, plastic = gen.RUNS[name]
    states, its = gen.run_sim(make_sim, sc_args, opts, plastic)
    assert its == [int(x) for x in G[f"run/{name}/iterations"]], (name, its)
    for k, st in enumerate(states):
        for key, v in st.items():
            gk = f"run/{name}/{k}/{key}"
            if gk in G.files:
                ref = G[gk]
                assert np.abs(v - ref).max() <= rtol * max(np.abs(ref).max(), 1e-300), (name, k, key, np.abs(v - ref).max() / np.abs(ref).max())


@pytest.mark.parametrize("name", list(gen.RUNS))
def test_oracle_time_steps_against_reference_code(oracle, name):
    """MultigridSimulation::advanceOneTimeStep's sequence (MultigridSimulation.h:235-297) in the reference's grid / objective / solver / multigrid /
    plasticity code against the oracle's, including the first-step cache of the characteristic norm while the snow model hardens"""
    _check_run(oracle.OracleSim, name, 1e-8)


@pytest.mark.skipif(not os.path.exists(gen.REF_LIB), reason="oracle/_ref/libimplicit_ref.so not built (needs /root/reference)")
def test_reference_objective_reproduces_the_golden_vectors():
    name = "slip"
    g = lambda k: G[f"{name}/{k}"]
    inp = {k: g("in_" + k) for k in ("X", "V", "mass", "C", "F", "vol", "mu", "lam")}
    ref = gen.Reference(float(g("dx")), float(g("dt")))
    n = ref.setup(inp, bool(g("project")))
    ref.set_bc(int(g("mode")), g("bc_node"), g("bc_P"), g("bc_R"), g("bc_Rinv"), g("bc_slip"))
    assert ref.updateState(g("dv")) == float(g("energy"))
    assert np.array_equal(ref.computeResidual(), g("residual"))
    col, val = ref.buildMatrix(1)
    assert np.array_equal(col, g("col1")) and np.array_equal(val, g("val1"))
    assert np.array_equal(ref.evaluatePerNodeCNTolerance(gen.CN_EPS, float(g("dt"))), g("cntol"))


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(gen.CASES))
def test_cuda_objective_against_reference_code(hot, name):
    _check(hot.MpmSimulationB200, name, exact_slots=False)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["hot_default", "hot_no_linesearch", "hot_stiff_long"])
def test_cuda_whole_solve_against_reference_code(hot, name):
    """hot_backward_euler_step against the reference's own solve (the substeps of test_oracle_lbfgs_ref.py, same bars)"""
    _check_step(hot.MpmSimulationB200, name, 1e-5)
