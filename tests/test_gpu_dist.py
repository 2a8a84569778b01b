"""Row (e): one object partitioned over 2 and 3 ranks (every rank holds ITS slab of the particles, pages at the seams are shared)
must reproduce the single-GPU result on every node it holds, through two whole time steps in which the particles move.  The
ranks run as separate processes on the same GPU and serve the library's collectives through gloo, so this runs on a one-GPU box;
bench.py --gpus N uses the same library path with NCCL inside the library.  Nodes are matched by coordinate (ids are local).
fp64 tolerance 1e-11 (fields; the summation order on shared nodes changes with the partition); identical Newton / PCG /
line-search counts."""
import numpy as np
import pytest

from test_dist_cpu import launch
import dist_worker

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def single(hot):
    sc = dist_worker.scene()
    sim = hot.MpmSimulationB200(sc["dx"])
    ymin = int(np.floor(sc["X"][:, 1].min() / sc["dx"] - 0.5))
    return dist_worker.run_object(sim, sc, np.arange(len(sc["mass"])), ymin)


def _key(coord):
    c = coord.astype(np.int64)
    return (c[:, 0] * 4096 + c[:, 1]) * 4096 + c[:, 2]


def _close(a, b, tol=1e-11):
    assert np.abs(a - b).max() <= tol * max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_object_matches_single_gpu(single, tmp_path, world):
    res = launch("gpu", world, tmp_path, timeout=900)
    order = np.argsort(_key(single["coord"]))
    skey = _key(single["coord"])[order]
    order2 = np.argsort(_key(single["coord_solve"]))
    skey2 = _key(single["coord_solve"])[order2]
    n_particles = len(single["P_X"])
    sel_all = np.concatenate([r["sel"] for r in res])
    assert len(np.unique(sel_all)) == n_particles == len(sel_all)
    assert sum(int(r["part"][5]) for r in res) == len(skey)                       # owned nodes partition the object's nodes
    seen = np.zeros(len(skey), dtype=int)
    for r in res:
        part = r["part"]
        assert part[1] == world and part[2] >= 1 and part[3] > 0 and part[6] == len(skey) and part[7] == len(r["sel"])
        pos = np.searchsorted(skey, _key(r["coord"]))
        assert (skey[pos] == _key(r["coord"])).all()                               # every local node exists in the single-GPU object
        at = order[pos]
        seen[pos] += 1
        _close(r["grid_m"], single["grid_m"][at], 1e-13); _close(r["grid_v"], single["grid_v"][at])
        assert abs(r["energy"] - single["energy"]) <= 1e-12 * abs(single["energy"])
        for k in ("residual", "multiply", "cn_tol", "diag"):
            _close(r[k], single[k][at])
        assert (r["log_iters"] == single["log_iters"]).all()
        for k in ("log_res0", "log_res1"):
            assert np.abs(r[k] - single[k]).max() <= 1e-5 * single[k].max()
        pos2 = np.searchsorted(skey2, _key(r["coord_solve"]))
        _close(r["dv0"], single["dv0"][order2[pos2]], 1e-7)
        for k in ("X", "V", "F", "C"):                                             # after TWO steps
            _close(r["P_" + k], single["P_" + k][r["sel"]], 1e-7)
    assert (seen >= 1).all() and (seen > 1).any()                                  # all nodes covered, some shared


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="one rank per physical GPU: needs >= 2 GPUs (gpurun --gpus 2)")
def test_peer_memory_transport_equals_nccl_transport(single, tmp_path):
    """the shared-page exchange through peer memory (P2P stores into the neighbours' arenas + flags) and through grouped
    ncclSend / ncclRecv add the same partial sums in the same order; the scatters themselves add page-group sums with RED atomics,
    so two runs agree to rounding, not bitwise: 1e-13 between the transports, the usual tolerances against the single-GPU object"""
    world = min(_n_gpus(), 4)
    (tmp_path / "peer").mkdir(); (tmp_path / "nccl").mkdir()
    peer = launch("nccl", world, tmp_path / "peer", timeout=900)
    nccl = launch("nccl", world, tmp_path / "nccl", timeout=900, env={"HOT_XCHG": "nccl"})
    order = np.argsort(_key(single["coord"])); skey = _key(single["coord"])[order]
    for a, b in zip(peer, nccl):
        assert "peer memory" in str(a["transport"]) and "ncclSend" in str(b["transport"])
        for k in ("grid_m", "grid_v", "residual", "multiply", "cn_tol", "diag"):
            _close(a[k], b[k], 1e-13)
        for k in ("dv0", "P_X", "P_V", "P_F", "P_C"):
            _close(a[k], b[k], 1e-9)
        assert (a["log_iters"] == b["log_iters"]).all()
        assert (a["log_iters"] == single["log_iters"]).all()
        at = order[np.searchsorted(skey, _key(a["coord"]))]
        _close(a["grid_m"], single["grid_m"][at], 1e-13); _close(a["grid_v"], single["grid_v"][at])
        for k in ("residual", "multiply", "cn_tol", "diag"):
            _close(a[k], single[k][at])
        for k in ("X", "V", "F", "C"):
            _close(a["P_" + k], single["P_" + k][a["sel"]], 1e-7)
