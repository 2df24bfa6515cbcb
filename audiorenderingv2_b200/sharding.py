"""Ray and source sharding across the GPUs of one box.

The seeded ray set is global: direction = f(seed, global ray id), so any partition of
[0, N) over ranks reproduces the single-GPU result after the IR histograms are summed
(NCCL all-reduce).  libarv2's default partition is by direction tiles (direction_tile_rank),
contiguous id ranges (ray_range) on request.  Convolution shards by source; a single source is never split.
"""
from __future__ import annotations


def ray_range(rank: int, world: int, n_rays: int):
    """Contiguous slice [begin, begin+count) of the N-ray set traced by `rank`."""
    if not (0 <= rank < world) or n_rays < 0:
        raise ValueError("bad rank/world/n_rays")
    base, rem = divmod(n_rays, world)
    begin = rank * base + min(rank, rem)
    return begin, base + (1 if rank < rem else 0)


def direction_tile_bits(n_total: int) -> int:
    """Tiles of the octahedral map for a set of n_total rays: 2^14, fewer when a tile would hold under 256 rays
    (the rule of libarv2's ensure_tiles)."""
    bits = 14
    while bits > 6 and (n_total >> bits) < 256:
        bits -= 1
    return bits


def direction_tile_rank(directions, world: int, tile_bits: int):
    """Host-side statement of the default multi-GPU shard rule (csrc/trace.cu: direction_select_kernel): the rank that
    traces a ray with emission direction d.  d is mapped to the octahedral square (u, v) in [-1, 1]^2, quantised to
    16 bits per axis, Morton-interleaved; the top tile_bits bits name the tile, tile mod world the rank.  float32
    throughout, like the kernel (which may contract a multiply-add: rays on a tile edge can differ)."""
    import numpy as np
    d = np.asarray(directions, np.float32).reshape(-1, 3)
    inv = np.float32(1.0) / (np.abs(d[:, 0]) + np.abs(d[:, 1]) + np.abs(d[:, 2]) + np.float32(1e-30))
    u, v = d[:, 0] * inv, d[:, 1] * inv
    neg = d[:, 2] < 0
    uu = (np.float32(1.0) - np.abs(v)) * np.copysign(np.float32(1.0), u)
    vv = (np.float32(1.0) - np.abs(u)) * np.copysign(np.float32(1.0), v)
    u, v = np.where(neg, uu, u).astype(np.float32), np.where(neg, vv, v).astype(np.float32)
    q = lambda t: np.clip((t * np.float32(0.5) + np.float32(0.5)) * np.float32(65536.0), 0.0, 65535.0).astype(np.uint32)

    def spread(x):
        x = (x | (x << 8)) & np.uint32(0x00FF00FF); x = (x | (x << 4)) & np.uint32(0x0F0F0F0F)
        x = (x | (x << 2)) & np.uint32(0x33333333); x = (x | (x << 1)) & np.uint32(0x55555555)
        return x
    key = spread(q(u)) | (spread(q(v)) << np.uint32(1))
    return ((key >> np.uint32(32 - tile_bits)) % np.uint32(world)).astype(np.int64)


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
