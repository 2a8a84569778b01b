"""worker of tests/test_gpu_dist.py and tests/test_dist_cpu.py (one process per rank, rendezvous on 127.0.0.1)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def scene():
    from hot_b200 import scenes
    sc = scenes.block((6, 14, 6), 0.04, ppc=6, seed=2, E=2e4)
    return sc


def setup(sim, sc, dt=4e-3):
    sim.set_particles(sc["X"], sc["V"], sc["mass"], sc["C"], sc["F"], sc["vol"], sc["mu"], sc["lam"])
    sim.set_dt_gravity(dt, (0, -9.8, 0))
    sim.sortParticlesAndPolluteGrid()
    n = sim.particlesToGrid()
    coord = sim.get_id2coord()
    bc = np.nonzero(coord[:, 1] <= coord[:, 1].min() + 1)[0].astype(np.int32)
    idx, _, v = sim.get_grid()
    vn = np.zeros((n, 3)); vn[idx[idx >= 0]] = v[idx >= 0]
    sim.set_bc(bc, P=np.zeros((len(bc), 9)), dv_bc=-vn[bc])
    return n, bc


def gpu_worker(rank, world, port, out_dir):
    """partitioned run on ONE physical GPU: every rank opens its own handle on cuda:0 and the all-reduce callback stages the
    exchange buffer through gloo (host) - the library neither knows nor cares which transport sums the buffer"""
    import torch
    import torch.distributed as dist
    import hot_b200
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    dev = torch.device("cuda", 0)
    sim = hot_b200.MpmSimulationB200(scene()["dx"], device=0)

    def alloc(n):
        t = torch.empty(n, dtype=torch.float64, device=dev)
        return t, t.data_ptr()

    def allreduce(buf, op, count):
        torch.cuda.synchronize()
        h = buf[:count].cpu()
        dist.all_reduce(h, op=dist.ReduceOp.MAX if op == 1 else dist.ReduceOp.SUM)
        buf[:count].copy_(h)
        torch.cuda.synchronize()

    sim.set_partition(rank, world, allreduce, alloc)
    sc = scene()
    n, bc = setup(sim, sc)
    part = sim.get_partition()
    res = {"part": np.array([part[k] for k in ("group0", "group1", "particle0", "particle1", "dof0", "dof1", "n_interface")]),
           "n_nodes": n}
    idx, m, v = sim.get_grid()
    res["grid_idx"] = idx; res["grid_m"] = m; res["grid_v"] = v
    sim.backupStrain()
    rng = np.random.default_rng(7)
    dv = sim.get_dv() + 0.2 * (rng.random((n, 3)) - 0.5)
    res["energy"] = sim.updateState(dv)
    res["residual"] = sim.computeResidual()
    x = rng.random((n, 3)) - 0.5
    res["multiply"] = sim.multiply(x)
    res["cn_tol"] = sim.evaluatePerNodeCNTolerance(1e-7, 4e-3)
    res["diag"] = sim.buildDiagonal(1)
    sim.restoreStrain()
    n2, bc2 = setup(sim, sc)                     # fresh step for the solve
    log = sim.backwardEulerStep(lsolver=2, matfree=1, bcproject=0, mg_level=1, max_newton_iterations=30, cneps=1e-8)
    res["log_iters"] = np.array([log["iterations"], log["total_linear_iterations"], log["total_linesearch_probes"], int(log["converged"])])
    res["log_res"] = np.array(log["residual_norm"])
    res["dv0"] = sim.get_dv0()
    part = sim.get_partition()
    order = sim.get_sort()[1]                     # before G2P: moving the particles invalidates the sort
    own = order[part["particle0"]:part["particle1"]]
    sim.gridToParticles(4e-3)
    p = sim.get_particles()
    res["own"] = own
    for k in ("X", "V", "F", "C"):
        res["P_" + k] = p[k][own]
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **res)
    dist.destroy_process_group()


def cpu_worker(rank, world, port, out_dir):
    """host-side protocol of the interface exchange under gloo (no GPU): contributions of the ranks that touch a node are
    summed, every rank ends with the same interface values; the group cut is contiguous, covering and balanced"""
    import torch
    import torch.distributed as dist
    from hot_b200.dist import split_groups
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(0)                      # same stream on every rank
    sizes = rng.integers(1, 400, size=500)
    group_first = np.concatenate([[0], np.cumsum(sizes)]).tolist()
    cut = split_groups(group_first, group_first[-1], world)
    # a rank's partial scatter: its groups add into the nodes [g, g + 3) (stencil overlap across the cut = interface)
    n_nodes = len(sizes) + 2
    full = np.zeros(n_nodes)
    partial = np.zeros(n_nodes)
    for g, s in enumerate(sizes):
        full[g:g + 3] += s
        if cut[rank] <= g < cut[rank + 1]:
            partial[g:g + 3] += s
    touched = [set(range(cut[r], cut[r + 1] + 2)) if cut[r + 1] > cut[r] else set() for r in range(world)]
    iface = sorted(i for i in range(n_nodes) if sum(i in t for t in touched) >= 2)
    buf = torch.from_numpy(partial[iface].copy())
    dist.all_reduce(buf)
    partial[iface] = buf.numpy()
    mine = sorted(touched[rank])
    ok = np.allclose(partial[mine], full[mine])
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), ok=ok, cut=np.array(cut), n_iface=len(iface))
    dist.destroy_process_group()


if __name__ == "__main__":
    kind, rank, world, port, out_dir = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    (gpu_worker if kind == "gpu" else cpu_worker)(rank, world, port, out_dir)
