"""Row (e): one object partitioned over 2 and 3 ranks must reproduce the single-GPU result.  The ranks run as separate
processes on the same GPU and sum their exchange buffers through gloo, so this runs on a one-GPU box; bench.py --gpus N uses
the same library path with NCCL over NVLink.  Bit-exact: DOF numbering.  fp64 tolerance 1e-11 (fields) - the summation order
of interface nodes changes with the partition; identical Newton / PCG / line-search counts."""
import numpy as np
import pytest

from test_dist_cpu import launch
import dist_worker

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def single(hot):
    sc = dist_worker.scene()
    sim = hot.MpmSimulationB200(sc["dx"])
    n, bc = dist_worker.setup(sim, sc)
    ref = {"n_nodes": n}
    ref["grid_idx"], ref["grid_m"], ref["grid_v"] = sim.get_grid()
    sim.backupStrain()
    rng = np.random.default_rng(7)
    dv = sim.get_dv() + 0.2 * (rng.random((n, 3)) - 0.5)
    ref["energy"] = sim.updateState(dv)
    ref["residual"] = sim.computeResidual()
    x = rng.random((n, 3)) - 0.5
    ref["multiply"] = sim.multiply(x)
    ref["cn_tol"] = sim.evaluatePerNodeCNTolerance(1e-7, 4e-3)
    ref["diag"] = sim.buildDiagonal(1)
    sim.restoreStrain()
    dist_worker.setup(sim, sc)
    log = sim.backwardEulerStep(lsolver=2, matfree=1, bcproject=0, mg_level=1, max_newton_iterations=30, cneps=1e-8)
    ref["log_iters"] = np.array([log["iterations"], log["total_linear_iterations"], log["total_linesearch_probes"], int(log["converged"])])
    ref["log_res"] = np.array(log["residual_norm"])
    ref["dv0"] = sim.get_dv0()
    sim.gridToParticles(4e-3)
    ref["P"] = sim.get_particles()
    return ref


def _close(a, b, tol=1e-11):
    assert np.abs(a - b).max() <= tol * max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_object_matches_single_gpu(single, tmp_path, world):
    res = launch("gpu", world, tmp_path, timeout=600)
    n = single["n_nodes"]
    parts = [r["part"] for r in res]
    # contiguous, covering partitions of groups, particles and DOF ids; a non-empty interface
    assert parts[0][0] == 0 and parts[0][2] == 0 and parts[0][4] == 0
    for a, b in zip(parts, parts[1:]):
        assert a[1] == b[0] and a[3] == b[2] and a[5] == b[4]
    assert parts[-1][5] == n and all(p[6] == parts[0][6] > 0 for p in parts)
    own_all = np.concatenate([r["own"] for r in res])
    assert len(np.unique(own_all)) == len(single["P"]["X"]) == len(own_all)
    for r, part in zip(res, parts):
        assert r["n_nodes"] == n and (r["grid_idx"] == single["grid_idx"]).all()       # replicated numbering: bit-exact
        d0, d1 = int(part[4]), int(part[5])
        sel = (single["grid_idx"] >= d0) & (single["grid_idx"] < d1)                    # grid values: complete on the owned nodes
        _close(r["grid_m"][sel], single["grid_m"][sel], 1e-13); _close(r["grid_v"][sel], single["grid_v"][sel])
        assert abs(r["energy"] - single["energy"]) <= 1e-12 * abs(single["energy"])
        for k in ("residual", "multiply", "cn_tol", "diag"):                         # valid on the owned nodes (and ghosts)
            _close(r[k][d0:d1], single[k][d0:d1])
        assert (r["log_iters"] == single["log_iters"]).all()
        assert np.abs(r["log_res"] - single["log_res"]).max() <= 1e-5 * single["log_res"].max()
        _close(r["dv0"][d0:d1], single["dv0"][d0:d1], 1e-7)
        for k in ("X", "V", "F", "C"):
            _close(r["P_" + k], single["P"][k][r["own"]], 1e-7)
