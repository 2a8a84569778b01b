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


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


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


@pytest.fixture(scope="module")
def single_mg(hot):
    sc = dist_worker.scene()
    sim = hot.MpmSimulationB200(sc["dx"])
    ymin = int(np.floor(sc["X"][:, 1].min() / sc["dx"] - 0.5))
    return dist_worker.run_mg(sim, sc, np.arange(len(sc["mass"])), ymin)


def _match(skey_order, coord):
    skey, order = skey_order
    pos = np.searchsorted(skey, _key(coord))
    assert (skey[pos] == _key(coord)).all()
    return order[pos], pos


@pytest.mark.parametrize("world", [2, 3, -2])
def test_partitioned_multigrid_path_matches_single_gpu(single_mg, tmp_path, world):
    """Assembled matrix, Galerkin hierarchy and V-cycle of a partitioned object (ghost ring on): level 0 distributed (rows summed over
    the sharers, take-over exchange after every SpMV / Gauss-Seidel colour phase), levels >= 1 replicated (all-reduced Galerkin
    product and restriction).  Order-independent operators (SpMV on every level, restriction, prolongation, a Jacobi V-cycle with the
    PCG coarse solve) must agree with the single-GPU object to rounding (1e-10); the Gauss-Seidel V-cycle sweeps blocks in the LOCAL
    node order, so it is compared as a solver: same contraction, HOT / PN-MGPCG converge to the same state with the same or nearly
    the same iteration counts."""
    if world < 0:                                                            # one rank per physical GPU, NCCL + peer-memory transport
        if _n_gpus() < -world:
            pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
        res = launch("nccl", -world, tmp_path, timeout=900, env={"HOT_TEST_MG": "1"})
        assert all("peer memory" in str(r["transport"]) for r in res)
    else:
        res = launch("gpu_mg", world, tmp_path, timeout=900)
    S = single_mg
    def keyed(c):
        o = np.argsort(_key(c)); return _key(c)[o], o
    k0, k1, k2 = keyed(S["coord"]), keyed(S["coord1"]), keyed(S["coord2"])
    gs_single = np.linalg.norm(S["vcycle_gs_residual"]) / np.linalg.norm(S["rhs"])
    assert gs_single < 0.6
    seen = np.zeros(len(k0[0]), dtype=int)
    num = np.zeros(len(k0[0])); den = np.zeros(len(k0[0]))
    for r in res:
        at0, pos0 = _match(k0, r["coord"])
        seen[pos0] += 1
        assert int(r["global_nodes"]) == len(k0[0])
        assert (r["dofs"][1:] == S["dofs"][1:]).all()                       # the replicated coarse levels are the single-GPU ones
        at1, _ = _match(k1, r["coord1"]); at2, _ = _match(k2, r["coord2"])
        assert len(at1) == len(k1[0]) and len(at2) == len(k2[0])
        _close(r["spmv0"], S["spmv0"][at0], 1e-10)
        _close(r["spmv1"], S["spmv1"][at1], 1e-10)
        _close(r["spmv2"], S["spmv2"][at2], 1e-10)
        _close(r["restrict0"], S["restrict0"][at1], 1e-10)
        _close(r["prolong0"], S["prolong0"][at0], 1e-10)
        _close(r["rhs"], S["rhs"][at0], 1e-10)
        _close(r["vcycle_jacobi"], S["vcycle_jacobi"][at0], 1e-9)
        num[pos0] = (r["vcycle_gs_residual"] ** 2).sum(1); den[pos0] = (r["rhs"] ** 2).sum(1)
        assert r["hot_log"][1] == 1 and abs(int(r["hot_log"][0]) - int(S["hot_log"][0])) <= 2
        assert r["pn_log"][2] == 1 and abs(int(r["pn_log"][0]) - int(S["pn_log"][0])) <= 1
        assert abs(r["hot_res"][0] - S["hot_res"][0]) <= 1e-9 * S["hot_res"][0]
        _close(r["hot_dv0"], S["hot_dv0"][at0], 2e-4)                        # both stop at the same tolerance, on different sweeps
        for k in ("X", "V", "F"):
            _close(r["P_" + k], S["P_" + k][r["sel"]], 1e-4)
    assert (seen >= 1).all() and (seen > 1).any()
    gs_part = np.sqrt(num.sum() / den.sum())
    assert gs_part < 0.6 and abs(gs_part - gs_single) < 0.1                  # the same smoother up to the in-block order
