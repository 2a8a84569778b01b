"""Pins rows a3 / a5 / a6 / a7 / a23 (node record, particle sort + page activation, APIC P2G, DOF numbering + normalisation, G2P) to the
REFERENCE'S OWN grid code: tests/golden/mpmgrid_ref.npz was produced by GridState / BSplineWeights / MpmGrid::{iterateKernel, getNumNodes,
iterateGrid} (Lib/MPM/MpmGrid.h) over the reference's SPGrid page map and B-spline header, compiled where they lie
(oracle/mpmgrid_ref_shim.cpp -> oracle/_ref/libmpmgrid_ref.so; tests/golden/make_mpmgrid_golden.py).  The oracle's restatement
(oracle/hot_oracle.cpp) and the CUDA path through the C ABI must reproduce: sort keys, order, page groups, page list in first-Set order,
DOF ids and id2coord bit-exactly; node masses to 1e-13, node velocities and particle results to 1e-11 / 1e-12 of the field magnitude (the
sums run in another order than the reference's serial loop); the CFL flags exactly."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("make_mpmgrid_golden", os.path.join(ROOT, "tests", "golden", "make_mpmgrid_golden.py"))
gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gen)
G = np.load(os.path.join(ROOT, "tests", "golden", "mpmgrid_ref.npz"))
EXACT = ("sorter", "order", "base", "first", "last", "blk", "pages", "num_nodes", "idx", "id2coord", "flags")
# relative to the largest magnitude of the field
CLOSE = {"m": 1e-13, "v": 1e-11, "pX": 1e-14, "pV": 1e-12, "pC": 1e-11, "pgradV": 1e-11}


def _inputs(name):
    inp = {k: G[f"{name}/in_{k}"] for k in ("X", "V", "mass", "C", "dv")}
    n = len(inp["mass"])
    inp["F"] = np.tile(np.eye(3).reshape(1, 9), (n, 1))
    inp["vol"] = np.ones(n); inp["mu"] = np.ones(n); inp["lam"] = np.ones(n)
    return inp


def _check(make_sim, name, scale=1.0):
    sim = make_sim(float(G[f"{name}/dx"]), apic_rpic_ratio=float(G[f"{name}/ratio"]))
    out = gen.run(sim, _inputs(name))
    for k in EXACT:
        a, b = np.asarray(out[k]), G[f"{name}/{k}"]
        assert a.shape == b.shape and (a == b).all(), (name, k)
    for k, tol in CLOSE.items():
        a, b = out[k], G[f"{name}/{k}"]
        assert a.shape == b.shape, (name, k)
        assert np.abs(a - b).max() <= scale * tol * max(np.abs(b).max(), 1e-300), (name, k, np.abs(a - b).max())


def test_gridstate_layout_is_the_128_byte_record():
    """MpmGrid.h:14-34 as the reference's compiler lays it out: v, m, new_v, idx at 0 / 24 / 32 / 56 of 128 bytes, 32 nodes per 4 KB page"""
    assert list(G["layout"]) == [128, 0, 24, 32, 56, 32]


@pytest.mark.parametrize("name", list(gen.CASES))
def test_oracle_transfers_against_reference_grid_code(oracle, name):
    _check(oracle.OracleSim, name)


@pytest.mark.skipif(not os.path.exists(gen.REF_LIB), reason="oracle/_ref/libmpmgrid_ref.so not built (needs /root/reference)")
def test_reference_grid_code_reproduces_the_golden_vectors():
    assert list(gen.Reference(0.1).layout()) == list(G["layout"])
    for name in ("tiny", "page_corner"):
        out = gen.run(gen.Reference(float(G[f"{name}/dx"]), apic_rpic_ratio=float(G[f"{name}/ratio"])), _inputs(name))
        for k in EXACT + tuple(CLOSE):
            assert np.array_equal(np.asarray(out[k]), G[f"{name}/{k}"]), (name, k)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(gen.CASES))
def test_cuda_transfers_against_reference_grid_code(hot, name):
    _check(hot.MpmSimulationB200, name)
