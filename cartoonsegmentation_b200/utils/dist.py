"""Multi-GPU plumbing of the hot path: images are independent, so ranks share nothing but a final counter gather.

  shard_indices   image i -> rank i mod world (the reference itself is a batch-1 loop over images, animeinsseg/__init__.py:485)
  gather_counters ONE all_gather of a small per-rank struct (frames, device ms, ...) -- the only collective of a run
  max_over_ranks  the timed region is reported as the maximum over ranks
"""
import torch
import torch.distributed as dist


def shard_indices(n, rank, world):
    return list(range(rank, n, world))


def gather_counters(counters: torch.Tensor):
    """counters: 1-D float64 tensor on the rank's device -> [world, len] tensor on the same device."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return counters[None]
    out = [torch.empty_like(counters) for _ in range(dist.get_world_size())]
    dist.all_gather(out, counters)
    return torch.stack(out)


def max_over_ranks(value: float, device):
    t = torch.tensor([value], device=device, dtype=torch.float64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
