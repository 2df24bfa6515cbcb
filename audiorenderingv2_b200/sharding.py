"""Ray-range and source sharding across the GPUs of one box.

The seeded ray set is global: direction = f(seed, global ray id), so any partition of
[0, N) over ranks reproduces the single-GPU result after the IR histograms are summed
(NCCL all-reduce).  Convolution shards by source; a single source is never split.
"""
from __future__ import annotations


def ray_range(rank: int, world: int, n_rays: int):
    """Contiguous slice [begin, begin+count) of the N-ray set traced by `rank`."""
    if not (0 <= rank < world) or n_rays < 0:
        raise ValueError("bad rank/world/n_rays")
    base, rem = divmod(n_rays, world)
    begin = rank * base + min(rank, rem)
    return begin, base + (1 if rank < rem else 0)


def sources_of(rank: int, world: int, n_sources: int):
    """Source s lives on GPU s mod world."""
    return [s for s in range(n_sources) if s % world == rank]


def all_reduce_hist(hist_tensor, world: int):
    """Sum the per-rank fp64 IR histograms in place (torch.distributed, NCCL on GPUs,
    gloo in the CPU tests)."""
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(hist_tensor)
    return hist_tensor
