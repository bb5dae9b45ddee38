"""Batch sharding across the GPUs of one box: instances are independent, so each rank owns a
contiguous block and the only collective is the gather of the per-instance results.  Pure host logic (torch
tensors on whatever device the process group uses), exercised on CPU with the gloo backend."""
import numpy as np


def shard_range(total, rank, world):
    """Contiguous block [lo, hi) of `total` instances owned by `rank`; blocks differ by at most one instance."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(total, world):
    return [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]


def gather_results(u0, status, iters, total, dist=None):
    """All-gather the applied controls [b][8] (float64) and status / SQP iteration counts [b] (int32) of every rank
    into arrays of the whole batch (ragged blocks are padded to the largest block for the collective)."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return u0, status, iters
    world = dist.get_world_size()
    sizes = shard_sizes(total, world)
    m = max(sizes)
    dev = u0.device
    pu = torch.zeros((m, 8), dtype=torch.float64, device=dev); pu[: u0.shape[0]] = u0
    ps = torch.zeros((m, 2), dtype=torch.int32, device=dev); ps[: status.shape[0], 0] = status; ps[: iters.shape[0], 1] = iters
    gu = torch.empty((world * m, 8), dtype=torch.float64, device=dev)
    gs = torch.empty((world * m, 2), dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(gu, pu)
    dist.all_gather_into_tensor(gs, ps)
    keep = torch.cat([torch.arange(r * m, r * m + sizes[r], device=dev) for r in range(world)])
    return gu[keep], gs[keep, 0], gs[keep, 1]


class ResultGatherer:
    """gather_results with every buffer allocated once (the per-cycle path of bench.py): equal blocks of b instances per rank."""

    def __init__(self, b, world, device):
        import torch
        self.b, self.world = b, world
        self.ps = torch.zeros((b, 2), dtype=torch.int32, device=device)
        self.gu = torch.empty((world * b, 8), dtype=torch.float64, device=device)
        self.gs = torch.empty((world * b, 2), dtype=torch.int32, device=device)

    def __call__(self, u0, status, iters, dist):
        self.ps[:, 0] = status; self.ps[:, 1] = iters
        dist.all_gather_into_tensor(self.gu, u0)
        dist.all_gather_into_tensor(self.gs, self.ps)
        return self.gu, self.gs[:, 0], self.gs[:, 1]


def reduce_counters(solved, failed, max_iters, dist=None):
    """Sum / max of the per-rank statistics (solved, failed, largest SQP iteration count)."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return int(solved), int(failed), int(max_iters)
    t = torch.tensor([solved, failed], dtype=torch.int64)
    mx = torch.tensor([max_iters], dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    return int(t[0]), int(t[1]), int(mx[0])
