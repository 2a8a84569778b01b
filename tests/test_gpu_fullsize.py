"""Parity against the CPU oracle at the sizes BASELINE.json names (not only on the small scenes of the other test files):
full-size C2 (951 k particles) and C4 (8 M particles) for the sort, P2G, G2P, residual and matrix-free multiply; the V-cycle and
a PN-MGPCG solve on the 232 k-particle slab of the C2 bar; scaled-down instances of the C3 / C5 stand-in scenes.

Tolerances as in the small-scene tests: integer outputs (keys, order, groups, pages in first-Set order, DOF ids) bit-exact;
masses rtol 1e-13; fields 1e-11 of the field magnitude (different summation order); solver residual norms 1e-5
relative with identical iteration counts (BASELINE.json)."""
import numpy as np
import pytest

from hot_b200 import scenes

pytestmark = pytest.mark.gpu


def _close(a, b, tol=1e-11):
    np.testing.assert_allclose(a, b, rtol=0, atol=tol * max(np.abs(b).max(), 1e-300))


SCENES = {
    "c2_full": lambda: scenes.config_c2(),
    "c4_full": lambda: scenes.config_c4(),
    "c3_quarter": lambda: scenes.config_c3(scale=0.25),
    "c5_quarter": lambda: scenes.config_c5(scale=0.25),
}


def _floor_caps(coord, cells=2):
    y = coord[:, 1]
    return np.nonzero((y <= y.min() + cells) | (y >= y.max() - cells))[0].astype(np.int32)


@pytest.mark.parametrize("name", list(SCENES))
def test_transfers_and_operators_at_config_size(hot, oracle, name):
    sc = SCENES[name]()
    dt = 1e-4 if name.startswith("c5") else 1e-3
    g = hot.MpmSimulationB200(sc["dx"]); o = oracle.OracleSim(sc["dx"])
    for s in (g, o):
        s.set_particles(sc["X"], sc["V"], sc["mass"], sc["C"], sc["F"], sc["vol"], sc["mu"], sc["lam"])
        s.set_dt_gravity(dt, (0, -9.8, 0))
        s.sortParticlesAndPolluteGrid()
    # a5 / a2: keys, order, base offsets, groups, page list in first-Set order
    for a, b in zip(g.get_sort(), o.get_sort()):
        assert (a == b).all()
    for a, b in zip(g.get_groups(), o.get_groups()):
        assert (a == b).all()
    assert (g.get_pages() == o.get_pages()).all()
    # a6 / a7
    n = g.particlesToGrid()
    assert n == o.particlesToGrid()
    gi, gm, gv = g.get_grid(); oi, om, ov = o.get_grid()
    assert (gi == oi).all()
    np.testing.assert_allclose(gm, om, rtol=1e-13, atol=0)
    _close(gv, ov)
    # a9-a14 on the same grid: updateState, residual, matrix-free multiply
    coord = o.get_id2coord()
    bc = _floor_caps(coord)
    rng = np.random.default_rng(1)
    for s in (g, o):
        s.backupStrain()
        s.set_bc(bc, P=np.zeros((len(bc), 9)), dv_bc=np.zeros((len(bc), 3)))
    vmax = np.abs(ov).max()
    dv = o.get_dv() + 0.02 * vmax * (rng.random((n, 3)) - 0.5)
    eg, eo = g.updateState(dv), o.updateState(dv)
    assert abs(eg - eo) <= 1e-11 * abs(eo)
    _close(g.computeResidual(), o.computeResidual())
    x = rng.random((n, 3)) - 0.5
    _close(g.multiply(x), o.multiply(x))
    for s in (g, o):
        s.restoreStrain()
    # a23: G2P + evolveStrain with the trial dv
    fg, fo = g.gridToParticles(dt), o.gridToParticles(dt)
    assert fg == fo
    pg, po = g.get_particles(), o.get_particles()
    for k, tol in (("X", 1e-14), ("V", 1e-12), ("C", 1e-11), ("gradV", 1e-11), ("F", 1e-12)):
        _close(pg[k], po[k], tol)
    o.close()


def _slab(hot, oracle, dt=1.0 / 480):
    """the 232 k-particle slab of the C2 bar that bench.py uses as the bounded CPU sample"""
    sc = scenes.block((22, 40, 22), 0.12 / 22, ppc=12, origin_cells=(16, 16, 16), rho=2000.0, E=1e5, nu=0.3, seed=0)
    g = hot.MpmSimulationB200(sc["dx"]); o = oracle.OracleSim(sc["dx"])
    for s in (g, o):
        s.set_particles(sc["X"], sc["V"], sc["mass"], sc["C"], sc["F"], sc["vol"], sc["mu"], sc["lam"])
        s.set_dt_gravity(dt, (0.0, 0.0, 0.0))
        s.sortParticlesAndPolluteGrid(); s.particlesToGrid()
        bc = _floor_caps(s.get_id2coord(), cells=8)
        s.set_bc(bc, P=np.zeros((len(bc), 9)), dv_bc=np.zeros((len(bc), 3)))
    return sc, g, o


def test_vcycle_on_c2_slab(hot, oracle):
    sc, g, o = _slab(hot, oracle)
    for s in (g, o):
        s.backupStrain(); s.updateState()
        s.buildMatrix(True)
        s.buildMultigrid(levels=3, smoother=5, coarseSolver=2, Ainv=1, times=1)
    assert g.level_dofs() == o.level_dofs()
    r = o.computeResidual()
    _close(g.computeResidual(), r)
    zg, zo = g.vcycle(r), o.vcycle(r)
    _close(zg, zo, 1e-9)
    assert g.vcycle_timing()[1] == o.vcycle_timing()[1]          # coarse-level PCG iterations
    o.close()


def test_pn_mgpcg_substep_on_c2_slab(hot, oracle):
    sc, g, o = _slab(hot, oracle)
    kw = dict(lsolver=2, matfree=0, bcproject=1, mg_level=3, smoother=5, coarse_solver=2, project=1, linesearch=1, usecn=1)
    lg, lo = g.backwardEulerStep(**kw), o.backwardEulerStep(**kw)
    assert lg["converged"] and lo["converged"]
    for k in ("iterations", "matrix_builds", "total_linear_iterations", "total_linesearch_probes"):
        assert lg[k] == lo[k], (k, lg[k], lo[k])
    rg, ro = np.array(lg["residual_norm"]), np.array(lo["residual_norm"])
    assert np.all(np.abs(rg - ro) <= 1e-5 * ro + 1e-9 * ro[0])
    assert np.abs(g.get_dv() - o.get_dv()).max() < 1e-6 * np.abs(o.get_dv()).max()
    o.close()
