"""torch.distributed plumbing for an object partitioned over several GPUs (one process per GPU).

`nccl_partition`: NCCL lives INSIDE the library (hot_comm_init_nccl): rank 0 draws the ncclUniqueId, torch.distributed only carries
its 128 bytes to the other ranks.  `host_partition`: the library's three collectives served by any torch.distributed backend on
host copies (gloo in the tests, which run several ranks on one GPU - something NCCL refuses to do)."""
import numpy as np
import torch
import torch.distributed as dist


def comm_unique_id(lib=None):
    import ctypes as C
    from ._lib import load_library
    lib = lib or load_library()
    buf = (C.c_ubyte * 128)()
    if lib.hot_comm_unique_id(buf) != 0:
        raise RuntimeError("hot_comm_unique_id failed (NCCL not loadable)")
    return bytes(buf)


def nccl_partition(sim, device=None, group=None):
    """hook `sim` to an NCCL communicator of its own over the ranks of `group`; returns (rank, world)"""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1:
        return rank, world
    on_gpu = dist.get_backend(group) == "nccl"
    t = torch.zeros(128, dtype=torch.uint8, device=device if on_gpu else "cpu")
    if rank == 0:
        t.copy_(torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(t, src=0, group=group)
    sim.init_nccl(rank, world, bytes(t.cpu().numpy().tobytes()))
    return rank, world


def host_partition(sim, group=None):
    """the library's collectives through torch.distributed on host arrays; returns (rank, world)"""
    rank, world = dist.get_rank(group), dist.get_world_size(group)

    def all_reduce(a, op):
        t = torch.from_numpy(a.copy())
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == 1 else dist.ReduceOp.SUM, group=group)
        return t.numpy()

    def all_gather(a):
        parts = [torch.empty(len(a), dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(parts, torch.from_numpy(a.copy()), group=group)
        return torch.cat(parts).numpy()

    def neighbor_exchange(peers, send):
        recv = [torch.empty(len(s), dtype=torch.float64) for s in send]
        reqs = []
        for p, s_, r_ in zip(peers, send, recv):
            reqs.append(dist.isend(torch.from_numpy(s_.copy()), dst=p, group=group))
            reqs.append(dist.irecv(r_, src=p, group=group))
        for q in reqs:
            q.wait()
        return [r_.numpy() for r_ in recv]

    sim.set_partition(rank, world, all_reduce, all_gather, neighbor_exchange)
    return rank, world


def split_slabs(X, world, axis=1):
    """cut an object into `world` slabs of equal particle count along `axis` (the partition bench.py and the tests hand to the
    ranks); returns the particle indices of every rank.  Host logic, testable on CPU."""
    order = np.argsort(X[:, axis], kind="stable")
    cuts = [(len(order) * r) // world for r in range(world + 1)]
    return [np.sort(order[cuts[r]:cuts[r + 1]]) for r in range(world)]
