"""Parity of the CUDA assembled-matrix / Galerkin-multigrid path (a15-a20) against the CPU oracle through the C ABI.

Bit-exact: coarse node sets and ids (first-touch order), node coordinates per level, the Gauss-Seidel sweep order
(colour, block, position).  fp64 tolerance: matrix entries 1e-11 of the matrix magnitude (different summation order in
the assembly), operator applications and smoother outputs 1e-10 of the field magnitude.
Also the reference's own checks on the GPU path: assembled == matrix-free (ImplicitSolver.h:698-739), symmetry
(SquareMatrix.h:84-109)."""
import numpy as np
import pytest
import scipy.sparse as sp

from hot_b200 import scenes

pytestmark = pytest.mark.gpu


def ell_to_csr(col, val, ncols):
    n, cs = col.shape
    if val.ndim == 2:                                              # scalar weights -> w * I3
        val = val[:, :, None] * np.eye(3).reshape(1, 1, 9)
    blocks = val.reshape(n * cs, 3, 3).transpose(0, 2, 1)
    A = sp.bsr_matrix((blocks, col.reshape(-1), np.arange(0, n * cs + 1, cs)), shape=(3 * n, 3 * ncols)).tocsr()
    A.sum_duplicates()
    return A


def _close(a, b, tol=1e-10):
    assert np.abs(a - b).max() <= tol * max(np.abs(b).max(), 1e-300)


def _pair(hot, oracle, cells=(9, 10, 9), seed=4, bc=True, slip=False, dx=0.04):
    sc = scenes.block(cells, dx, ppc=6, seed=seed, E=1e4)
    g = hot.MpmSimulationB200(sc["dx"]); o = oracle.OracleSim(sc["dx"])
    rng = np.random.default_rng(seed)
    dvp = None
    for s in (g, o):
        s.set_particles(sc["X"], sc["V"], sc["mass"], sc["C"], sc["F"], sc["vol"], sc["mu"], sc["lam"])
        s.set_dt_gravity(2e-3, (0, -9.8, 0))
        s.sortParticlesAndPolluteGrid(); s.particlesToGrid(); s.backupStrain()
        coord = s.get_id2coord()
        bcn = np.nonzero(coord[:, 1] <= coord[:, 1].min() + 1)[0].astype(np.int32) if bc else np.zeros(0, dtype=np.int32)
        if slip:
            th = 0.3
            R = np.array([[np.cos(th), np.sin(th), 0], [-np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
            kw = dict(R=np.tile(R.T.reshape(1, 9), (len(bcn), 1)), Rinv=np.tile(R.reshape(1, 9), (len(bcn), 1)),
                      slip=(np.arange(len(bcn)) % 2).astype(np.int32), mode=1)
        else:
            kw = dict(P=np.zeros((len(bcn), 9)))
        s.set_bc(bcn, dv_bc=np.zeros((len(bcn), 3)), **kw)
        if dvp is None:
            dvp = s.get_dv() + 0.2 * (rng.random((s.num_nodes, 3)) - 0.5)
        s.updateState(dvp)
    return g, o, bcn


@pytest.mark.parametrize("bcproject,slip", [(False, False), (True, False), (True, True)])
def test_assembled_matrix_parity(hot, oracle, bcproject, slip):
    g, o, bc = _pair(hot, oracle, cells=(5, 6, 5), slip=slip)
    g.buildMatrix(bcproject); o.buildMatrix(bcproject)
    n = g.num_nodes
    Ag = ell_to_csr(*g.get_matrix(), n); Ao = ell_to_csr(*o.get_matrix(), n)
    assert abs(Ag - Ao).max() < 1e-11 * abs(Ao).max()
    assert abs(Ag - Ag.T).max() < 1e-10 * abs(Ag).max()
    x = np.random.default_rng(0).random((n, 3)) - 0.5
    _close(g.spmv(0, x), o.spmv(0, x))
    if not bcproject:
        _close(g.spmv(0, x), g.multiply(x))                          # matrixSanityCheck on the GPU path
    _close(g.buildDiagonal(1), o.buildDiagonal(1), 1e-10)
    _close(g.buildDiagonal(0), o.buildDiagonal(0), 1e-10)


@pytest.fixture(scope="module")
def mg(hot, oracle):
    g, o, bc = _pair(hot, oracle)
    for s in (g, o):
        s.buildMatrix(True)
        s.buildMultigrid(levels=3, smoother=5, coarseSolver=2, Ainv=1)
    return g, o


def test_hierarchy_parity(mg):
    g, o = mg
    dofs = o.level_dofs()
    assert g.level_dofs() == dofs
    for l in range(3):
        assert (g.level_coords(l) == o.level_coords(l)).all()          # coarse ids in first-touch order: bit-exact
        assert (g.color_order(l) == o.color_order(l)).all()            # GS sweep order: bit-exact
        Ag = ell_to_csr(*g.level_matrix(l, 0), dofs[l]); Ao = ell_to_csr(*o.level_matrix(l, 0), dofs[l])
        assert abs(Ag - Ao).max() < 1e-11 * abs(Ao).max()
        Dg, Dig = g.level_diagonal(l); Do, Dio = o.level_diagonal(l)
        _close(Dg, Do, 1e-11); _close(Dig, Dio, 1e-9)
        x = np.random.default_rng(l).random((dofs[l], 3)) - 0.5
        _close(g.spmv(l, x), o.spmv(l, x))
    for l in range(2):
        Pg = ell_to_csr(*g.level_matrix(l, 1), dofs[l + 1]); Po = ell_to_csr(*o.level_matrix(l, 1), dofs[l + 1])
        assert abs(Pg - Po).max() == 0
        Rg = ell_to_csr(*g.level_matrix(l, 2), dofs[l]); Ro = ell_to_csr(*o.level_matrix(l, 2), dofs[l])
        assert abs(Rg - Ro).max() == 0
        x = np.random.default_rng(l).random((dofs[l], 3)) - 0.5
        _close(g.restrict(l, x), o.restrict(l, x), 1e-13)
        y = np.random.default_rng(l).random((dofs[l + 1], 3)) - 0.5
        _close(g.prolong(l, y), o.prolong(l, y), 1e-13)


@pytest.mark.parametrize("level", [0, 1, 2])
@pytest.mark.parametrize("kind,iters", [(5, 1), (5, 3), (0, 2), (1, 2), (2, 10000)])
def test_smoother_parity(mg, level, kind, iters):
    g, o = mg
    n = o.level_dofs()[level]
    rng = np.random.default_rng(10 * level + kind)
    r0 = rng.random((n, 3)) - 0.5
    u0 = 0.01 * (rng.random((n, 3)) - 0.5)
    ug, rg = g.smooth(level, kind, u0, r0, iters, initial_residual=4.0 * r0 if kind == 2 else None)
    uo, ro = o.smooth(level, kind, u0, r0, iters, initial_residual=4.0 * r0 if kind == 2 else None)
    _close(ug, uo); _close(rg, ro)


@pytest.mark.parametrize("level", [0, 1, 2])
def test_chebyshev_smoother_and_2norm_estimate_parity(mg, level):
    """-smoother 6 (chebyshev_smooth, MultigridPreconditioner.h:227-264) with the 2-norm estimate of SquareMatrix::estimate2norm
    (SquareMatrix.h:375-475): same fixed start vector on both sides, so the power iterations agree step by step"""
    g, o = mg
    (gmax, gmin), (omax, omin) = g.estimate2norm(level), o.estimate2norm(level)
    assert abs(gmax - omax) <= 1e-9 * omax and abs(gmin - omax / 30) <= 1e-9 * omax
    n = o.level_dofs()[level]
    rng = np.random.default_rng(60 + level)
    r0 = rng.random((n, 3)) - 0.5
    u0 = 0.01 * (rng.random((n, 3)) - 0.5)
    for iters in (1, 4):
        ug, rg = g.smooth(level, 6, u0, r0, iters)
        uo, ro = o.smooth(level, 6, u0, r0, iters)
        _close(ug, uo); _close(rg, ro)


@pytest.mark.parametrize("smoother,coarse", [(5, 2), (5, 5), (0, 2), (1, 0), (6, 2)])
def test_vcycle_parity(hot, oracle, smoother, coarse):
    g, o, _ = _pair(hot, oracle)
    for s in (g, o):
        s.buildMatrix(True)
        s.buildMultigrid(levels=3, smoother=smoother, coarseSolver=coarse, Ainv=1, times=2 if smoother != 5 else 1)
    n = g.num_nodes
    b = np.random.default_rng(1).random((n, 3)) - 0.5
    zg, zo = g.vcycle(b), o.vcycle(b)
    _close(zg, zo, 1e-9)
    t, _ = g.vcycle_timing()
    assert t[:3, 0].min() > 0
    assert g.vcycle_bench(2) > 0


def test_single_level_and_ainv0(hot, oracle):
    g, o, _ = _pair(hot, oracle, cells=(5, 6, 5))
    for s in (g, o):
        s.buildMatrix(True)
        s.buildMultigrid(levels=1, smoother=5, coarseSolver=2, Ainv=0)
    b = np.random.default_rng(2).random((g.num_nodes, 3)) - 0.5
    _close(g.vcycle(b), o.vcycle(b), 1e-9)


def test_mg_argument_errors(hot, oracle):
    g, o, _ = _pair(hot, oracle, cells=(4, 4, 4))
    with pytest.raises(hot.HotError):
        g.buildMultigrid()                 # no matrix yet
    g.buildMatrix(False)
    with pytest.raises(hot.HotError):
        g.buildMultigrid(levels=3)         # multigrid needs --bcproject (ImplicitSolver.h:339)
    g.buildMatrix(True)
    with pytest.raises(hot.HotError):
        g.buildMultigrid(smoother=7)       # incomplete Cholesky (Eigen IncompleteCholesky) is not provided
    with pytest.raises(hot.HotError):
        g.buildMultigrid(levels=11)
