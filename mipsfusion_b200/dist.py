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


def allreduce_sum_async_(t, group):
    """Starts the all-reduce on the collective library's own stream (ordered after the work already queued on the current
    stream); the caller queues independent kernels and then calls ``.wait()`` on the returned handle (None for a world of 1)."""
    import torch.distributed as dist
    if world(group)[0] > 1:
        return dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=True)
    return None


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


def DeviceTypeCUDA():
    from torch._C._autograd import DeviceType
    return DeviceType.CUDA


class PeerArena:
    """One float32 arena per rank in symmetric (peer-mapped) memory: every rank can load / store every other rank's
    copy over NVLink with plain pointers.  Used by the data-parallel mapper to fuse the gradient reduce-scatter, Adam
    and the parameter all-gather into one kernel (mf_adam_step_sharded).  ``regions``: dict name -> number of floats
    (each region is padded to a multiple of 4 floats, 16-byte aligned)."""

    def __init__(self, regions, device, group, use_multicast=True):
        import torch.distributed._symmetric_memory as symm_mem
        self.offsets, self.sizes = {}, {}
        off = 0
        for name, n in regions.items():
            n4 = (int(n) + 3) // 4 * 4
            self.offsets[name], self.sizes[name] = off, n4
            off += n4
        self.group = group
        self.world, self.rank = world(group)
        self.buf = symm_mem.empty(off, dtype=torch.float32, device=device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, group.group_name if hasattr(group, "group_name") else group)
        self.peer_bases = [int(x) for x in self.handle.buffer_ptrs]
        # NVSwitch multicast mapping of the arena (NVLS), 0 when the fabric / driver does not offer it
        self.multicast_base = 0
        if use_multicast:
            try:
                if self.handle.has_multicast_support(DeviceTypeCUDA(), device.index if device.index is not None else torch.cuda.current_device()):
                    self.multicast_base = int(self.handle.multicast_ptr or 0)
            except Exception:
                try:
                    self.multicast_base = int(self.handle.multicast_ptr or 0)
                except Exception:
                    self.multicast_base = 0
        if len(self.peer_bases) != self.world:
            raise RuntimeError("symmetric memory rendezvous returned %d peers for a world of %d" % (len(self.peer_bases), self.world))

    def view(self, name, n=None):
        o = self.offsets[name]
        return self.buf[o:o + (self.sizes[name] if n is None else int(n))]

    def barrier(self):
        """Cross-GPU barrier on the current stream (device side, no host synchronisation)."""
        self.handle.barrier(channel=0)
