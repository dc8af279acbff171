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


def _gather_batch_dim(local, dim, sizes, rank, device, group):
    """all_gather of tensors that differ only in the length of ``dim`` (the batch axis); padded to the largest shard."""
    world = len(sizes)
    mx = max(h - l for l, h in sizes)
    moved = local.movedim(dim, 0).contiguous()
    pad = torch.zeros((mx,) + tuple(moved.shape[1:]), dtype=local.dtype, device=device)
    pad[: moved.shape[0]] = moved
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][: h - l] for r, (l, h) in enumerate(sizes)], dim=0).movedim(0, dim)


def sharded_sample(local_fn, global_batch, sample_shape, device, group=None, gather=True, trajectories=False):
    """Run ``local_fn(lo, hi)`` on this rank's shard ``[lo, hi)`` of the global batch and gather.

    local_fn is typically ``lambda lo, hi: superposed_sample(models, ddpm, (hi-lo,1,H,W), device, seed=s,
    sample_offset=lo)`` and returns x ``[hi-lo, *sample_shape]``.  With ``trajectories=True`` it returns
    ``(x, kappas [T, hi-lo, M], logq [T+1, hi-lo, M])`` (``return_trajectory=True``) and the kappa / log q trajectories are
    gathered along their batch axis too.  Returns the full-batch tensor(s) on every rank: one all_gather
    per tensor; uneven shards are padded to the largest shard for the collective.
    """
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    lo, hi = shard_range(global_batch, rank, world)
    out = local_fn(lo, hi)
    if trajectories:
        local, kap, lq = out
        assert kap.shape[1] == hi - lo and lq.shape[1] == hi - lo, (kap.shape, lq.shape, lo, hi)
    else:
        local = out
    assert tuple(local.shape) == (hi - lo,) + tuple(sample_shape), (local.shape, lo, hi)
    if world == 1 or not gather:
        return out
    sizes = [shard_range(global_batch, r, world) for r in range(world)]
    x = _gather_batch_dim(local, 0, sizes, rank, device, group)
    if not trajectories:
        return x
    return x, _gather_batch_dim(kap, 1, sizes, rank, device, group), _gather_batch_dim(lq, 1, sizes, rank, device, group)
