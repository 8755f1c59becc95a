"""Data-parallel plumbing for the sampling path: patches / volumes are independent work items (no cross-item
state, GroupNorm is per sample), so each rank (one process per GPU) takes a contiguous block of items and the only
exchange is the final gather of decoded slabs (NCCL over NVLink on GPUs; gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """contiguous block [lo, hi) of `n_items` for `rank`; the first n_items % world ranks get one extra item"""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def patch_grid(full, size, stride):
    """sliding-window start offsets along one axis (reference inference/sampler.py:388-395)"""
    return sorted(set(list(range(0, full - size + 1, stride)) + [max(0, full - size)]))


def volume_work_items(n_volumes, H, W, patch=192, stride=96):
    """flat (volume, h0, w0) list; contiguous sharding keeps a volume's patches on one rank when counts divide"""
    hs, ws = patch_grid(H, patch, stride), patch_grid(W, patch, stride)
    return [(v, h0, w0) for v in range(n_volumes) for h0 in hs for w0 in ws]


def gather_slabs(local, counts=None, group=None):
    """all-gather decoded slabs (n_r, C, D, H, W) along dim 0 -> (sum n_r, C, D, H, W) on every rank.
    counts: per-rank item counts when they differ (ragged shards are padded to the max for the collective)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    if counts is None:
        out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    nmax = max(counts)
    pad = torch.zeros((nmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * nmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * nmax: r * nmax + counts[r]] for r in range(world)], dim=0)


class SlabGatherer:
    """The final gather of decoded slabs, taken off the critical path: every call issues one asynchronous
    all_gather_into_tensor (NCCL's own stream, ordered after the producer by an event) into one of `depth` rotating
    output buffers and returns immediately, so the next batch's encode / sampling overlaps the transfer and the ranks
    are not re-synchronised once per batch (a per-batch blocking collective couples every step to the slowest,
    power-capped GPU).  `finish()` waits for everything outstanding; buffers are reused, so consume a result before
    `depth` further gathers are issued."""

    def __init__(self, depth=2, group=None):
        self.depth, self.group, self.bufs, self.pending, self.i = depth, group, {}, [], 0

    def gather(self, local):
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return local
        world = dist.get_world_size(self.group)
        key = (tuple(local.shape), local.dtype, local.device)
        if key not in self.bufs:
            self.bufs[key] = [torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype,
                                          device=local.device) for _ in range(self.depth)]
        out = self.bufs[key][self.i % self.depth]
        self.i += 1
        while len(self.pending) >= self.depth:  # the buffer about to be overwritten must have been produced
            self.pending.pop(0)[0].wait()
        work = dist.all_gather_into_tensor(out, local.contiguous(), group=self.group, async_op=True)
        self.pending.append((work, local))  # keep the source alive until the collective has read it
        return out

    def finish(self):
        for work, _ in self.pending:
            work.wait()
        self.pending = []
