"""Pins the oracle's outer solvers (a21 inexact PCG, a22 L-BFGS, Newton, line search, a24 backward-Euler glue):
PCG against a textbook implementation on the exported matrix, and the nonlinear solves against each other - PN-PCG(mf),
PN-PCG, PN-MGPCG and HOT (L-BFGS + 3-level MG) must reach the same minimiser of the incremental potential."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from hot_b200 import scenes
from test_oracle_matrix import ell_to_csr


def _scene(oracle, cells=(6, 7, 6), E=2e4, dt=4e-3, seed=2):
    sc = scenes.block(cells, 0.04, ppc=6, seed=seed, E=E)
    o = oracle.OracleSim(sc["dx"])
    o.set_particles(sc["X"], sc["V"], sc["mass"], sc["C"], sc["F"], sc["vol"], sc["mu"], sc["lam"])
    o.set_dt_gravity(dt, (0, -9.8, 0))
    o.sortParticlesAndPolluteGrid()
    o.particlesToGrid()
    coord = o.get_id2coord()
    bc = np.nonzero(coord[:, 1] <= coord[:, 1].min() + 1)[0].astype(np.int32)
    v = o.get_grid()[2]; idx = o.get_grid()[0]
    vn = np.zeros((o.num_nodes, 3)); vn[idx[idx >= 0]] = v[idx >= 0]
    o.set_bc(bc, P=np.zeros((len(bc), 9)), dv_bc=-vn[bc])            # sticky floor: v + dv = 0
    return sc, o, bc


def test_inexact_pcg_vs_textbook(oracle):
    sc, o, bc = _scene(oracle)
    o.backupStrain(); o.updateState()
    o.buildMatrix(True); o.buildMultigrid(levels=3)
    n = o.num_nodes
    A = ell_to_csr(*o.level_matrix(0, 0), n)
    b = o.project(np.random.default_rng(0).random((n, 3)) - 0.5)
    x, it = o.pcg(b, tolerance=1e-30, max_iterations=7, preconditioner=1)   # forcing = 0.5 -> stops at ||r||_M-1 halved
    _, Di = o.level_diagonal(0)
    Minv = sp.block_diag([sp.csr_matrix(m) for m in Di.reshape(n, 3, 3).transpose(0, 2, 1)], format="csr")
    xr = np.zeros(3 * n); r = b.reshape(-1).copy(); z = Minv @ r; p = z.copy(); zr = z @ r; k = 0
    target = min(0.5, np.sqrt(max(np.sqrt(zr), 1e-30))) * np.sqrt(zr)
    while np.sqrt(zr) >= target and k < 7:
        q = A @ p; a = zr / (q @ p); xr += a * p; r -= a * q; z = Minv @ r; zn = z @ r; p = z + (zn / zr) * p; zr = zn; k += 1
    assert it == k and it >= 1
    assert np.abs(x.reshape(-1) - xr).max() < 1e-11 * np.abs(xr).max()
    # V-cycle preconditioning converges in fewer iterations than block Jacobi
    x1, it1 = o.pcg(b, tolerance=1e-30, max_iterations=200, preconditioner=1)
    x2, it2 = o.pcg(b, tolerance=1e-30, max_iterations=200, preconditioner=2)
    assert it2 <= it1


CONFIGS = {
    "pn_pcg_mf": dict(lsolver=2, matfree=1, bcproject=0, mg_level=1),
    "pn_pcg": dict(lsolver=2, matfree=0, bcproject=0, mg_level=1),
    "pn_mgpcg": dict(lsolver=2, matfree=0, bcproject=1, mg_level=3),
    "pn_minres_mf": dict(lsolver=1, matfree=1, bcproject=0, mg_level=1),            # -lsolver 1: Newton + MINRES (Minres.h)
    "pn_mgminres": dict(lsolver=1, matfree=0, bcproject=1, mg_level=3),
    "hot": dict(lsolver=3, bcproject=1, mg_level=3),
    "lbfgs_h": dict(lsolver=3, bcproject=0, mg_level=1, mg_times=10000, smoother=2, coarse_solver=2),
}


@pytest.fixture(scope="module")
def solves(oracle):
    out = {}
    for name, kw in CONFIGS.items():
        sc, o, bc = _scene(oracle)
        log = o.backwardEulerStep(max_newton_iterations=30, cneps=1e-9, **kw)
        out[name] = (o.get_dv0(), log, bc, o.get_dv())
    return out


@pytest.mark.parametrize("name", list(CONFIGS))
def test_backward_euler_converges(solves, name):
    dv, log, bc, _ = solves[name]
    assert log["converged"], log
    assert log["scaled_norm"][-1] < 1.0                               # shouldExitByCN: sum |r|^2/tol^2 < n
    assert log["residual_norm"][-1] < 1e-3 * log["residual_norm"][0]
    e = log["energy"]
    assert all(e[i + 1] <= e[i] + 1e-12 * abs(e[i]) for i in range(1, len(e) - 1))   # line search: monotone energy
    assert log["matrix_builds"] == (0 if name in ("pn_pcg_mf", "pn_minres_mf") else (1 if name in ("hot", "lbfgs_h") else log["iterations"]))


def test_all_solvers_agree(solves):
    """the accepted iterates (ImplicitSolverObjective::dv0) of all five solver configurations coincide"""
    ref = solves["pn_pcg_mf"][0]
    scale = np.abs(ref).max()
    for name in CONFIGS:
        assert np.abs(solves[name][0] - ref).max() < 2e-4 * scale, name


def test_linesearch_leaves_one_extra_step_in_dv(oracle, solves):
    """with --linesearch the reference exits with dv = dv0 + last accepted step (x aliases simulation.dv
    and `x += step` runs after lineSearch already moved the nodes); without it dv == dv0."""
    dv0, log, bc, dv = solves["hot"]
    assert np.abs(dv - dv0).max() > 0
    sc, o, _ = _scene(oracle)
    o.backwardEulerStep(lsolver=2, matfree=1, bcproject=0, mg_level=1, linesearch=0, max_newton_iterations=30, cneps=1e-9)
    assert np.array_equal(o.get_dv(), o.get_dv0())
    assert np.abs(o.get_dv0() - dv0).max() < 2e-4 * np.abs(dv0).max()


def test_option_errors(oracle):
    sc, o, bc = _scene(oracle, cells=(4, 4, 4))
    with pytest.raises(RuntimeError):
        o.backwardEulerStep(lsolver=3, matfree=1)                     # README:13-15
    with pytest.raises(RuntimeError):
        o.backwardEulerStep(lsolver=0)
