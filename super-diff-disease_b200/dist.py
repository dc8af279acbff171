"""Batch-sharded sampling across the GPUs of one box.

Every quantity on the path is per-sample (conv, GroupNorm, kappa, log q), so the batch shards with
NO data-path collective: rank r of G samples global indices [lo, hi) with both models' weights
resident, and one all_gather of the finished samples closes the call.  Noise is keyed by GLOBAL
sample index, so the gathered result is bit-identical for any G.
"""
import torch
import torch.distributed as dist


def shard_range(global_batch, rank, world_size):
    """Contiguous, balanced [lo, hi) of the global batch owned by ``rank``."""
    base, rem = divmod(global_batch, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sharded_sample(local_fn, global_batch, sample_shape, device, group=None, gather=True):
    """Run ``local_fn(lo, hi) -> tensor [hi-lo, *sample_shape]`` on this rank's shard and gather.

    local_fn is typically ``lambda lo, hi: superposed_sample(models, ddpm, (hi-lo,1,H,W), device,
    seed=s, sample_offset=lo)``.  Returns the full [global_batch, *sample_shape] tensor on every rank
    (one all_gather; uneven shards are padded to the largest shard for the collective).
    """
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    lo, hi = shard_range(global_batch, rank, world)
    local = local_fn(lo, hi)
    assert tuple(local.shape) == (hi - lo,) + tuple(sample_shape), (local.shape, lo, hi)
    if world == 1 or not gather:
        return local
    sizes = [shard_range(global_batch, r, world) for r in range(world)]
    mx = max(h - l for l, h in sizes)
    pad = torch.zeros((mx,) + tuple(sample_shape), dtype=local.dtype, device=device)
    pad[: hi - lo] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][: h - l] for r, (l, h) in enumerate(sizes)], dim=0)
