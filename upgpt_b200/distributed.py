"""Multi-GPU layer of the hot path: batch sharding + the single all-gather of decoded frames.

Every image is an independent denoising chain, so ranks never talk during sampling; weights are replicated.  The only
collective is one all-gather of the decoded uint8 frames (NCCL over NVLink/NVSwitch on GPUs, gloo in CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous [lo, hi) slice of n samples owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(t, rank=None, world=None):
    """Slice of a batch-major tensor (or list of tensors / dict of them) owned by this rank."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    if isinstance(t, dict):
        return {k: shard_batch(v, rank, world) for k, v in t.items()}
    if isinstance(t, (list, tuple)):
        return [shard_batch(v, rank, world) for v in t]
    lo, hi = shard_range(t.shape[0], rank, world)
    return t[lo:hi]


def gather_frames(frames, n_total=None, sizes=None):
    """All-gathers per-rank frame tensors (B_r, ...) into (sum B_r, ...) on every rank, preserving global sample order.
    Uneven shards are padded to the largest shard for the collective and trimmed afterwards. sizes: the per-rank shard sizes when the
    caller knows them (e.g. shard_range): skips the size exchange, leaving ONE collective and no host synchronisation."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return frames
    world = dist.get_world_size()
    if sizes is None:
        n_local = torch.tensor([frames.shape[0]], device=frames.device, dtype=torch.int64)
        sizes = [torch.zeros_like(n_local) for _ in range(world)]
        dist.all_gather(sizes, n_local)
    sizes = [int(s) for s in sizes]
    mx = max(sizes)
    if frames.shape[0] < mx:
        pad = frames.new_zeros((mx - frames.shape[0],) + tuple(frames.shape[1:]))
        frames = torch.cat([frames, pad], 0)
    out = frames.new_empty((world * mx,) + tuple(frames.shape[1:]))
    dist.all_gather_into_tensor(out, frames.contiguous())
    if all(s == mx for s in sizes):
        return out
    return torch.cat([out[r * mx:r * mx + sizes[r]] for r in range(world)], 0)


def per_sample_noise(shape_per_sample, global_indices, seed, steps=None, device=None):
    """Gaussian noise whose value for a sample depends on (seed, the sample's GLOBAL index) only -- not on the rank that draws it or on
    how many ranks share the batch ('per-sample generator offsets so results are independent of rank count').
    -> (len(global_indices), *shape_per_sample), or (steps, len(global_indices), *shape_per_sample) for per-step sampler noise."""
    outs = []
    for gi in global_indices:
        g = torch.Generator().manual_seed((int(seed) * 1000003 + int(gi)) & 0x7FFFFFFFFFFFFFFF)
        full = ((steps,) if steps is not None else ()) + tuple(shape_per_sample)
        outs.append(torch.randn(full, generator=g))
    t = torch.stack(outs, 1 if steps is not None else 0)
    return t if device is None else t.to(device)
