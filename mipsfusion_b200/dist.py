"""Host-side sharding helpers for the paths that partition naturally across the GPUs of one box
(SURVEY.md 8e): pose candidates (RandomOptimizer), submaps / grid slabs (joint query) and ray batches
(data-parallel mapping).  One process per GPU, torch.distributed (NCCL on the GPUs; the same code runs
under gloo in the CPU tests).  Nothing here touches the kernels."""
import torch


def world(group=None):
    """(world_size, rank) of `group`, (1, 0) when torch.distributed is not initialised / group is None."""
    import torch.distributed as dist
    if group is None or not dist.is_available() or not dist.is_initialized():
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def shard_range(total, world_size, rank):
    """Contiguous, balanced-by-ceiling shard [begin, begin+count) of range(total): every rank gets
    ceil(total / world) items except the tail ranks (possibly 0).  -> (begin, count, per)"""
    per = (total + world_size - 1) // world_size
    begin = min(rank * per, total)
    return begin, min(per, total - begin), per


def round_robin(total, world_size, rank):
    """Items m with m % world == rank (submap m lives on GPU m mod G)."""
    return [m for m in range(total) if m % world_size == rank]


def allgather_rows(local, total, group):
    """local: (per, ...) rows of this rank's shard (padded to `per`); -> (total, ...) rows of all shards in order."""
    import torch.distributed as dist
    ws, _ = world(group)
    if ws == 1:
        return local[:total]
    out = torch.empty((ws * local.shape[0],) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out[:total]


def allreduce_sum_(t, group):
    import torch.distributed as dist
    if world(group)[0] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def allreduce_max_(t, group):
    import torch.distributed as dist
    if world(group)[0] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return t


def average_gradients_(tensors, group):
    """Sum-all-reduce then divide by the world size (data-parallel mapping: every rank holds the loss of its own
    equally sized ray batch, normalised by its local batch size and weighted with the GLOBAL mask counts)."""
    ws, _ = world(group)
    if ws == 1:
        return tensors
    for t in tensors:
        allreduce_sum_(t, group)
        t.mul_(1.0 / ws)
    return tensors
