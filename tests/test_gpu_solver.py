"""Parity of the CUDA solvers (a21 PCG, a22 L-BFGS, Newton, line search, backward-Euler glue) against the CPU oracle.

Iteration counts, matrix rebuild counts and line-search probe counts must be IDENTICAL (same algorithm, same stopping
tests); per-iteration residual norms within 1e-5 relative (BASELINE.json: "per-substep residuals within 1e-5 relative of
the CPU reference"); energies 1e-9 relative; resulting dv / dv0 1e-6 of the field magnitude."""
import numpy as np
import pytest

from hot_b200 import scenes

pytestmark = pytest.mark.gpu


def _scene(hot, oracle, cells=(6, 7, 6), E=2e4, dt=4e-3, seed=2):
    sc = scenes.block(cells, 0.04, ppc=6, seed=seed, E=E)
    g = hot.MpmSimulationB200(sc["dx"]); o = oracle.OracleSim(sc["dx"])
    for s in (g, o):
        s.set_particles(sc["X"], sc["V"], sc["mass"], sc["C"], sc["F"], sc["vol"], sc["mu"], sc["lam"])
        s.set_dt_gravity(dt, (0, -9.8, 0))
        s.sortParticlesAndPolluteGrid()
        s.particlesToGrid()
        coord = s.get_id2coord()
        bc = np.nonzero(coord[:, 1] <= coord[:, 1].min() + 1)[0].astype(np.int32)
        idx, _, v = s.get_grid()
        vn = np.zeros((s.num_nodes, 3)); vn[idx[idx >= 0]] = v[idx >= 0]
        s.set_bc(bc, P=np.zeros((len(bc), 9)), dv_bc=-vn[bc])
    return sc, g, o, bc


def test_pcg_parity(hot, oracle):
    sc, g, o, bc = _scene(hot, oracle)
    for s in (g, o):
        s.backupStrain(); s.updateState()
        s.buildMatrix(True); s.buildMultigrid(levels=3)
    n = g.num_nodes
    b = o.project(np.random.default_rng(0).random((n, 3)) - 0.5)
    for scale in (1.0, 1e-6):       # small right-hand sides tighten the forcing term sqrt(||r||) -> more iterations
        for matfree, pre in [(False, 1), (False, 2), (True, 1), (False, 0)]:
            xg, ig = g.pcg(scale * b, tolerance=1e-30, max_iterations=40, matfree=matfree, preconditioner=pre)
            xo, io = o.pcg(scale * b, tolerance=1e-30, max_iterations=40, matfree=matfree, preconditioner=pre)
            assert ig == io and ig >= 1, (matfree, pre, ig, io)
            assert np.abs(xg - xo).max() < 1e-8 * np.abs(xo).max()


CONFIGS = {
    "pn_pcg_mf": dict(lsolver=2, matfree=1, bcproject=0, mg_level=1),
    "pn_pcg": dict(lsolver=2, matfree=0, bcproject=0, mg_level=1),
    "pn_mgpcg": dict(lsolver=2, matfree=0, bcproject=1, mg_level=3),
    "pn_minres_mf": dict(lsolver=1, matfree=1, bcproject=0, mg_level=1),            # -lsolver 1: Newton + MINRES (Minres.h)
    "pn_mgminres": dict(lsolver=1, matfree=0, bcproject=1, mg_level=3),
    "hot": dict(lsolver=3, bcproject=1, mg_level=3),
    "hot_nolinesearch": dict(lsolver=3, bcproject=1, mg_level=3, linesearch=0, usecn=0, cneps=1e-6),
    "lbfgs_h": dict(lsolver=3, bcproject=0, mg_level=1, mg_times=10000, smoother=2, coarse_solver=2),
}


@pytest.mark.parametrize("name", list(CONFIGS))
def test_backward_euler_step_parity(hot, oracle, name):
    sc, g, o, bc = _scene(hot, oracle)
    kw = dict(max_newton_iterations=30, max_lbfgs_iterations=200, **CONFIGS[name])
    lg, lo = g.backwardEulerStep(**kw), o.backwardEulerStep(**kw)
    assert lo["converged"] and lg["converged"]
    for k in ("iterations", "matrix_builds", "total_linear_iterations", "total_linesearch_probes", "linear_iterations"):
        assert lg[k] == lo[k], (k, lg[k], lo[k])
    assert abs(lg["tolerance"] - lo["tolerance"]) <= 1e-12 * lo["tolerance"]
    rg, ro = np.array(lg["residual_norm"]), np.array(lo["residual_norm"])
    assert np.abs(rg - ro).max() <= 1e-5 * ro.max()                  # per-iteration residual parity
    assert np.all(np.abs(rg - ro) <= 1e-5 * ro + 1e-9 * ro[0])
    if kw.get("linesearch", 1):
        eg, eo = np.array(lg["energy"]), np.array(lo["energy"])
        assert np.abs(eg - eo).max() <= 1e-9 * np.abs(eo).max()
    scale = np.abs(o.get_dv()).max()
    assert np.abs(g.get_dv() - o.get_dv()).max() < 1e-6 * scale
    assert np.abs(g.get_dv0() - o.get_dv0()).max() < 1e-6 * scale
    # the step leaves the strain restored and the grid ready for G2P: finish the time step on both
    g.gridToParticles(4e-3); o.gridToParticles(4e-3)
    pg, po = g.get_particles(), o.get_particles()
    for k in ("X", "V", "F"):
        assert np.abs(pg[k] - po[k]).max() < 1e-6 * np.abs(po[k]).max()


def test_option_errors(hot, oracle):
    sc, g, o, bc = _scene(hot, oracle, cells=(4, 4, 4))
    with pytest.raises(hot.HotError):
        g.backwardEulerStep(lsolver=3, matfree=1)
    with pytest.raises(hot.HotError):
        g.backwardEulerStep(lsolver=0)
