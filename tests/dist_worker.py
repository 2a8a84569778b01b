"""worker of tests/test_gpu_dist.py and tests/test_dist_cpu.py (one process per rank, rendezvous on 127.0.0.1)"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

DT = 4e-3


def scene():
    from hot_b200 import scenes
    return scenes.block((6, 14, 6), 0.04, ppc=6, seed=2, E=2e4)


def node_field(coord, seed):
    """a smooth field of the node COORDINATES (node ids differ between ranks and from a single-GPU run)"""
    c = coord.astype(np.float64)
    return np.stack([np.sin(0.37 * c[:, 0] + 0.11 * c[:, 1] + seed), np.cos(0.23 * c[:, 1] - 0.19 * c[:, 2] + seed), np.sin(0.29 * c[:, 2] + 0.31 * c[:, 0] - seed)], 1)


def setup(sim, sc, sel, ymin):
    sim.set_particles(sc["X"][sel], sc["V"][sel], sc["mass"][sel], sc["C"][sel], sc["F"][sel], sc["vol"][sel], sc["mu"][sel], sc["lam"][sel])
    sim.set_dt_gravity(DT, (0, -9.8, 0))
    return begin_step(sim, ymin)


def begin_step(sim, ymin):
    sim.sortParticlesAndPolluteGrid()
    n = sim.particlesToGrid()
    coord = sim.get_id2coord()
    bc = np.nonzero(coord[:, 1] <= ymin + 1)[0].astype(np.int32)
    idx, _, v = sim.get_grid()
    vn = np.zeros((n, 3)); vn[idx[idx >= 0]] = v[idx >= 0]
    sim.set_bc(bc, P=np.zeros((len(bc), 9)), dv_bc=-vn[bc])
    return n, coord


def run_object(sim, sc, sel, ymin):
    """what both the partitioned ranks and the single-GPU reference do; everything is returned per LOCAL node with its coordinates"""
    res = {}
    n, coord = setup(sim, sc, sel, ymin)
    res["coord"] = coord
    part = sim.get_partition()                                   # of the first step's sort
    res["part"] = np.array([part[k] for k in ("rank", "world", "neighbors", "shared_pages", "exchange_pages", "owned_nodes", "global_nodes", "particles")])
    idx, m, v = sim.get_grid()
    act = idx >= 0
    gm = np.zeros(n); gv = np.zeros((n, 3)); gm[idx[act]] = m[act]; gv[idx[act]] = v[act]
    res["grid_m"] = gm; res["grid_v"] = gv
    sim.backupStrain()
    dv = sim.get_dv() + 0.2 * node_field(coord, 0.3)
    res["energy"] = sim.updateState(dv)
    res["residual"] = sim.computeResidual()
    res["multiply"] = sim.multiply(node_field(coord, 1.7))
    res["cn_tol"] = sim.evaluatePerNodeCNTolerance(1e-7, DT)
    res["diag"] = sim.buildDiagonal(1)
    sim.restoreStrain()
    # two whole time steps with the matrix-free PN-PCG solver: the particles move between them
    logs = []
    for step in range(2):
        if step:
            n, coord = begin_step(sim, ymin)
        log = sim.backwardEulerStep(lsolver=2, matfree=1, bcproject=0, mg_level=1, max_newton_iterations=30, cneps=1e-8)
        logs.append(log)
        if step == 0:
            res["dv0"] = sim.get_dv0(); res["coord_solve"] = coord
        sim.gridToParticles(DT)
    res["log_iters"] = np.array([[l["iterations"], l["total_linear_iterations"], l["total_linesearch_probes"], int(l["converged"])] for l in logs])
    res["log_res0"] = np.array(logs[0]["residual_norm"]); res["log_res1"] = np.array(logs[1]["residual_norm"])
    p = sim.get_particles()
    for k in ("X", "V", "F", "C"):
        res["P_" + k] = p[k]
    return res


HOT_FLAGS = dict(lsolver=3, mg_level=3, smoother=5, coarse_solver=2, project=1, linesearch=1, bcproject=1, usecn=1, cneps=1e-7)  # tog.sh:38


def run_mg(sim, sc, sel, ymin):
    """assembled matrix / Galerkin hierarchy / V-cycle / HOT solve; everything per LOCAL node (level 0) or per coarse node with its
    coordinates, so that a partitioned run can be matched against the single-GPU one"""
    res = {}
    n, coord = setup(sim, sc, sel, ymin)
    res["coord"] = coord
    res["global_nodes"] = np.array(sim.get_partition()["global_nodes"])
    sim.backupStrain()
    sim.updateState(sim.get_dv() + 0.2 * node_field(coord, 0.3))
    sim.buildMatrix(True)
    x = node_field(coord, 1.7)
    # order-independent smoother first: Jacobi V-cycle (smoother 0) with the PCG coarse solve must agree to rounding
    sim.buildMultigrid(levels=3, smoother=0, coarseSolver=2, Ainv=1, times=2)
    res["dofs"] = np.array(sim.level_dofs())
    res["spmv0"] = sim.spmv(0, x)
    c1 = sim.level_coords(1); c2 = sim.level_coords(2)
    res["coord1"] = c1; res["coord2"] = c2
    res["spmv1"] = sim.spmv(1, node_field(c1, 0.9))
    res["spmv2"] = sim.spmv(2, node_field(c2, 0.4))
    res["restrict0"] = sim.restrict(0, x)
    res["prolong0"] = sim.prolong(0, node_field(c1, 0.5))
    r = sim.spmv(0, node_field(coord, 2.1))          # a right-hand side in the range of the (BC-projected) operator
    res["rhs"] = r
    res["vcycle_jacobi"] = sim.vcycle(r)
    # the HOT smoother: coloured block Gauss-Seidel (in-block order follows the LOCAL numbering: compared as a solver, not entry-wise)
    sim.buildMultigrid(levels=3, smoother=5, coarseSolver=2, Ainv=1, times=1)
    z = sim.vcycle(r)
    res["vcycle_gs"] = z
    res["vcycle_gs_residual"] = r - sim.spmv(0, z)
    sim.restoreStrain()
    log = sim.backwardEulerStep(**HOT_FLAGS)
    res["hot_log"] = np.array([log["iterations"], int(log["converged"])])
    res["hot_res"] = np.array(log["residual_norm"])
    res["hot_dv0"] = sim.get_dv0()
    sim.gridToParticles(DT)
    # second step: PN-MGPCG (Newton + PCG on the assembled matrix, V-cycle preconditioner, tog.sh:49)
    n, coord = begin_step(sim, ymin)
    log = sim.backwardEulerStep(lsolver=2, matfree=0, mg_level=3, smoother=5, coarse_solver=2, project=1, linesearch=1, bcproject=1, usecn=1, cneps=1e-7)
    res["pn_log"] = np.array([log["iterations"], log["total_linear_iterations"], int(log["converged"])])
    sim.gridToParticles(DT)
    p = sim.get_particles()
    for k in ("X", "V", "F"):
        res["P_" + k] = p[k]
    return res


def gpu_mg_worker(rank, world, port, out_dir):
    """like gpu_worker, with the ghost ring on: assembled matrix, replicated coarse levels, take-over exchanges"""
    import torch.distributed as dist
    import hot_b200
    from hot_b200.dist import host_partition, split_slabs
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    sc = scene()
    sel = split_slabs(sc["X"], world)[rank]
    sim = hot_b200.MpmSimulationB200(sc["dx"], device=0)
    host_partition(sim)
    sim.set_ghost_ring(True)
    ymin = int(np.floor(sc["X"][:, 1].min() / sc["dx"] - 0.5))
    res = run_mg(sim, sc, sel, ymin)
    res["sel"] = sel
    part = sim.get_partition()
    res["part"] = np.array([part[k] for k in ("rank", "world", "neighbors", "shared_pages", "exchange_pages", "owned_nodes", "global_nodes", "particles")])
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **res)
    dist.destroy_process_group()


def gpu_worker(rank, world, port, out_dir):
    """partitioned run on ONE physical GPU: every rank opens its own handle on cuda:0 with ITS slab of the particles; the library's
    collectives are served by gloo on host copies (hot_b200.dist.host_partition) - NCCL cannot run two ranks on one device"""
    import torch.distributed as dist
    import hot_b200
    from hot_b200.dist import host_partition, split_slabs
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    sc = scene()
    sel = split_slabs(sc["X"], world)[rank]
    sim = hot_b200.MpmSimulationB200(sc["dx"], device=0)
    host_partition(sim)
    ymin = int(np.floor(sc["X"][:, 1].min() / sc["dx"] - 0.5))
    res = run_object(sim, sc, sel, ymin)
    res["sel"] = sel
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **res)
    dist.destroy_process_group()


def nccl_worker(rank, world, port, out_dir):
    """one rank per physical GPU, NCCL communicator inside the library (hot_comm_init_nccl); HOT_XCHG selects the transport of the
    shared-page exchange (peer memory by default, `nccl` = grouped ncclSend / ncclRecv)"""
    import torch
    import torch.distributed as dist
    import hot_b200
    from hot_b200.dist import nccl_partition, split_slabs
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    sc = scene()
    sel = split_slabs(sc["X"], world)[rank]
    sim = hot_b200.MpmSimulationB200(sc["dx"], device=rank)
    nccl_partition(sim, torch.device("cuda", rank))
    ymin = int(np.floor(sc["X"][:, 1].min() / sc["dx"] - 0.5))
    if os.environ.get("HOT_TEST_MG") == "1":                       # the assembled-matrix / multigrid path (ghost ring on)
        sim.set_ghost_ring(True)
        res = run_mg(sim, sc, sel, ymin)
    else:
        res = run_object(sim, sc, sel, ymin)
    res["sel"] = sel
    res["transport"] = np.array(sim.get_transport())
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **res)
    dist.destroy_process_group()


def cpu_worker(rank, world, port, out_dir):
    """the N > 1 host logic without a GPU: the library's shared-page tables (hot_share_tables, the code dist_after_sort runs) on page
    sets that overlap between ranks, then pack -> neighbour exchange -> unpack in ascending rank order emulated in numpy over gloo
    with the same collectives the GPU test uses.  Every sharer must end with bit-identical totals = the sum over the sharers in
    ascending rank order."""
    import torch
    import torch.distributed as dist
    from hot_b200._lib import load_library
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lib = load_library()
    rng = np.random.default_rng(100 + rank)
    # page ids of rank r: an own block + random pages of a common pool (shared by 2, 3, ... ranks)
    own = np.arange(1000 * rank, 1000 * rank + 300)
    pool = np.arange(50000, 50400)
    pids = np.unique(np.concatenate([own, rng.choice(pool, size=150, replace=False)])).astype(np.uint32)
    slot_sorted = rng.permutation(len(pids)).astype(np.int32)          # local slot of the i-th smallest page id
    value = rng.random((len(pids), 4))                                    # partial sums per local slot
    # all-gather of counts and padded lists (what dist_after_sort does on the device)
    counts = [torch.zeros(1, dtype=torch.int32) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([len(pids)], dtype=torch.int32))
    counts = np.array([int(c) for c in counts], dtype=np.int32)
    maxp = int(counts.max())
    mine = np.full(maxp, 0xffffffff, dtype=np.uint32); mine[:len(pids)] = pids
    parts = [torch.empty(maxp, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(parts, torch.from_numpy(mine.astype(np.int64)))
    all_pids = np.stack([p.numpy() for p in parts]).astype(np.uint32)
    nm = len(pids)
    n_nbr, n_x, n_sh = C.c_int(), C.c_int(), C.c_int()
    nbr_rank = np.zeros(world, np.int32); nbr_off = np.zeros(world, np.int64); nbr_cnt = np.zeros(world, np.int64)
    x_slot = np.zeros(max(1, (world - 1) * nm), np.int32); sh_slot = np.zeros(nm, np.int32); sh_ptr = np.zeros(nm + 1, np.int32)
    sh_entry = np.zeros(world * nm, np.int32); sh_owned = np.zeros(nm, np.int32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.hot_share_tables(rank, world, maxp, vp(counts), vp(np.ascontiguousarray(all_pids)), vp(slot_sorted), C.byref(n_nbr), vp(nbr_rank), vp(nbr_off),
                              vp(nbr_cnt), C.byref(n_x), vp(x_slot), C.byref(n_sh), vp(sh_slot), vp(sh_ptr), vp(sh_entry), vp(sh_owned))
    assert rc == 0
    n_nbr, n_x, n_sh = n_nbr.value, n_x.value, n_sh.value
    # pack -> exchange -> unpack (ascending rank order, -1 = own partial)
    send = value[x_slot[:n_x]]
    recv = np.zeros_like(send)
    reqs = []
    for j in range(n_nbr):
        a, b = int(nbr_off[j]), int(nbr_off[j] + nbr_cnt[j])
        reqs.append(dist.isend(torch.from_numpy(send[a:b].copy()), dst=int(nbr_rank[j])))
        rbuf = torch.empty((b - a, 4), dtype=torch.float64)
        reqs.append((dist.irecv(rbuf, src=int(nbr_rank[j])), a, b, rbuf))
    for q in reqs:
        if isinstance(q, tuple):
            q[0].wait(); recv[q[1]:q[2]] = q[3].numpy()
        else:
            q.wait()
    total = value.copy()
    for p in range(n_sh):
        acc = None
        for e in sh_entry[sh_ptr[p]:sh_ptr[p + 1]]:
            x = value[sh_slot[p]] if e < 0 else recv[e]
            acc = x.copy() if acc is None else acc + x
        total[sh_slot[p]] = acc
    # expected: gather everyone's (page id, partial) and add in ascending rank order
    allv = [torch.empty((maxp, 4), dtype=torch.float64) for _ in range(world)]
    by_id = np.zeros((maxp, 4)); by_id[:nm] = value[slot_sorted]           # by_id[i] = partial of the i-th smallest page id
    dist.all_gather(allv, torch.from_numpy(by_id))
    ok = True
    shared_ids = set()
    for i, pid in enumerate(pids):
        acc = None; sharers = 0
        for r in range(world):
            k = np.searchsorted(all_pids[r, :counts[r]], pid)
            if k < counts[r] and all_pids[r, k] == pid:
                x = allv[r][k].numpy(); acc = x.copy() if acc is None else acc + x; sharers += 1
        ok = ok and np.array_equal(total[slot_sorted[i]], acc)             # bit-exact
        if sharers > 1:
            shared_ids.add(int(pid))
            lowest = min(r for r in range(world) if pid in all_pids[r, :counts[r]])
            k = list(sh_slot[:n_sh]).index(slot_sorted[i])
            ok = ok and int(sh_owned[k]) == int(lowest == rank)
    ok = ok and n_sh == len(shared_ids) and all(nbr_rank[j] < nbr_rank[j + 1] for j in range(n_nbr - 1))
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), ok=ok, n_shared=n_sh, n_nbr=n_nbr, n_x=n_x)
    dist.destroy_process_group()


if __name__ == "__main__":
    kind, rank, world, port, out_dir = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    {"gpu": gpu_worker, "gpu_mg": gpu_mg_worker, "nccl": nccl_worker, "cpu": cpu_worker}[kind](rank, world, port, out_dir)
