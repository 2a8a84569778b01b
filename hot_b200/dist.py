"""torch.distributed plumbing for a partitioned object: the exchange buffer is a torch tensor, the all-reduce is
torch.distributed.all_reduce (NCCL over NVLink on the GPU box) enqueued on the stream the library works on."""
import torch
import torch.distributed as dist


def torch_partition(sim, device, group=None):
    """hook `sim` (MpmSimulationB200) to the default process group; returns (rank, world)"""
    rank, world = dist.get_rank(group), dist.get_world_size(group)

    def alloc(n):
        t = torch.empty(n, dtype=torch.float64, device=device)
        return t, t.data_ptr()

    def allreduce(buf, op, count):
        dist.all_reduce(buf[:count], op=dist.ReduceOp.MAX if op == 1 else dist.ReduceOp.SUM, group=group)

    sim.set_partition(rank, world, allreduce, alloc)
    return rank, world


def split_groups(group_first, n_particles, world):
    """the balanced contiguous cut of the page groups that dist.cu::dist_after_sort computes (host logic, testable on CPU):
    group_first = first sorted particle of every group + [n_particles]; returns world + 1 group boundaries"""
    import bisect
    cut = [0] * (world + 1)
    cut[world] = len(group_first) - 1
    for r in range(1, world):
        target = int(float(n_particles) * r / world)
        c = bisect.bisect_left(group_first, target)
        cut[r] = max(min(c, len(group_first) - 1), cut[r - 1])
    return cut
