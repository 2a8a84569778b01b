"""Pins the oracle's restatement of rows a16-a20 (block-row SpMV, diagonals, Galerkin hierarchy, 8-colour block ordering, Gauss-Seidel /
Jacobi / optimal-Jacobi / PCG smoothers, V-cycle) to the REFERENCE'S OWN code: tests/golden/mg_ref.npz was produced by
ZIRAN::MultigridBuilder::build, SquareMatrix and MultigridOperator (Projects/multigrid/{MultigridPreconditioner.h, MPMMultigridMatrix.h,
SquareMatrix.h}) compiled where they lie (oracle/mg_ref_shim.cpp -> oracle/_ref/libziran_ref.so) and run on the level-0 system the oracle
assembles (tests/golden/make_mg_golden.py).  The oracle - and the CUDA path through the C ABI - must build the same levels (node counts,
colour keys bit-exact; prolongation, diagonals and the action of every level's matrix to rounding) and return the same smoother and
V-cycle results."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("make_mg_golden", os.path.join(ROOT, "tests", "golden", "make_mg_golden.py"))
gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gen)
G = np.load(os.path.join(ROOT, "tests", "golden", "mg_ref.npz"))
L = gen.LEVELS


def _close(a, b, tol):
    assert a.shape == b.shape
    assert np.abs(a - b).max() <= tol * max(np.abs(b).max(), 1e-300)


def _levels(s, tol):
    """hierarchy of the (GS, PCG) configuration against the reference's"""
    s.buildMultigrid(levels=L, smoother=5, coarseSolver=2, Ainv=1, times=1)
    dofs = s.level_dofs()
    assert list(dofs) == [int(x) for x in G["s5c2t1_dofs"]]
    for l in range(L):
        x, u0 = gen.vectors(dofs[l], 10 + l)
        _close(s.spmv(l, x), G[f"spmv{l}"], tol)                       # a16 / a17: the level's matrix through its action
        D, Di = s.level_diagonal(l)
        _close(D, G[f"diag{l}"], tol); _close(Di, G[f"dinv{l}"], 10 * tol)
        assert np.array_equal(s.color_order(l), G[f"color{l}"])          # (colour, first-seen block, position in block): bit-exact
    for l in range(L - 1):
        pc, pv = s.level_matrix(l, 1)
        w = pv[:, :, 0] if pv.ndim == 3 else pv     # (the oracle keeps weight * I blocks like the reference, the CUDA side scalar weights)
        assert np.array_equal(pc[w != 0], G[f"pcol{l}"][G[f"pw{l}"] != 0])   # parents of every fine node, in the reference's slot order
        _close(w, G[f"pw{l}"], 1e-15)
    return dofs


def _smoothers(s, dofs, tol):
    for l in range(L):
        x, u0 = gen.vectors(dofs[l], 10 + l)
        for kind, iters in gen.SMOOTHERS.items():
            u, r = s.smooth(l, kind, u0, x, iters, initial_residual=4.0 * x if kind == 2 else None)
            _close(u, G[f"smooth{kind}_l{l}_u"], tol); _close(r, G[f"smooth{kind}_l{l}_r"], tol)


def _vcycles(s, tol):
    for smoother, coarse, times in gen.CONFIGS:
        s.buildMultigrid(levels=L, smoother=smoother, coarseSolver=coarse, Ainv=1, times=times)
        b, _ = gen.vectors(s.level_dofs()[0], 1)
        _close(s.vcycle(b), G[f"s{smoother}c{coarse}t{times}_vcycle"], tol)


def test_oracle_multigrid_against_reference_code(oracle):
    o = gen.scene(oracle.OracleSim)
    dofs = _levels(o, 1e-12)
    _smoothers(o, dofs, 1e-10)
    _vcycles(o, 1e-9)


@pytest.mark.skipif(not os.path.exists(gen.REF_LIB), reason="oracle/_ref/libziran_ref.so not built (needs /root/reference)")
def test_reference_multigrid_reproduces_the_golden_vectors(oracle):
    o = gen.scene(oracle.OracleSim)
    ref = gen.Reference()
    ref.build(o, 5, 2, 1)
    b, _ = gen.vectors(ref.dofs[0], 1)
    _close(ref.vcycle(b), G["s5c2t1_vcycle"], 1e-13)
    x, u0 = gen.vectors(ref.dofs[1], 11)
    _close(ref.smooth(1, 5, u0, x, 2)[0], G["smooth5_l1_u"], 1e-13)


def test_oracle_chebyshev_against_reference_code(oracle):
    """f4, -smoother 6: SquareMatrix::estimate2norm (SquareMatrix.h:375-475) and chebyshev_smooth (MultigridPreconditioner.h:227-264) of the reference.
    The reference seeds the start vector of the power iteration from the clock; cheb_ref.npz keeps the vector it drew (recovered after the call), and the
    oracle's iteration started from the same vector must land on the same lMax / lMin; with them the Chebyshev sweeps and the V-cycle must agree."""
    Cb = np.load(os.path.join(ROOT, "tests", "golden", "cheb_ref.npz"))
    o = gen.scene(oracle.OracleSim)
    o.buildMultigrid(levels=L, smoother=6, coarseSolver=2, Ainv=1, times=1)
    dofs = o.level_dofs()
    assert list(dofs) == [int(x) for x in Cb["dofs"]]
    for l in range(L):
        lmax, lmin = o.estimate2norm_from(l, Cb[f"start{l}"].astype(np.float64))
        assert abs(lmax - float(Cb[f"lmax{l}"])) <= 1e-12 * float(Cb[f"lmax{l}"]) and abs(lmin - float(Cb[f"lmin{l}"])) <= 1e-12 * float(Cb[f"lmin{l}"])
        lmax_own, _ = o.estimate2norm(l)                       # the oracle's own fixed start vector: the same norm to the iteration's tolerance
        assert abs(lmax_own - lmax) <= 1e-4 * lmax
        o.estimate2norm_from(l, Cb[f"start{l}"].astype(np.float64))
    for l in range(L):
        x, u0 = gen.vectors(dofs[l], 10 + l)
        u, r = o.smooth(l, 6, u0, x, gen.CHEB_ITERS)
        _close(u, Cb[f"cheb_l{l}_u"], 1e-10); _close(r, Cb[f"cheb_l{l}_r"], 1e-10)
    b, _ = gen.vectors(dofs[0], 1)
    _close(o.vcycle(b), Cb["vcycle"], 1e-9)


@pytest.mark.gpu
def test_cuda_multigrid_against_reference_code(hot):
    g = gen.scene(hot.MpmSimulationB200)
    dofs = _levels(g, 1e-11)
    _smoothers(g, dofs, 1e-9)
    _vcycles(g, 1e-8)
