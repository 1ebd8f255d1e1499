"""Seed sharding across the GPUs of one box.

Every seed is an independent unit of work, so the path shards with no data-path collective: rank r owns the
contiguous seed range [r*S/G, (r+1)*S/G), the input cloud and the weights are replicated, and ONE all-gather of
the displaced points (padded to the largest shard, trimmed afterwards) runs at the end.  torch.distributed is
plumbing (NCCL on the GPU box, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_bounds(S, world):
    """Contiguous, balanced ranges: the first S % world ranks get one extra seed."""
    base, extra = divmod(int(S), int(world))
    bounds, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        bounds.append((lo, hi))
        lo = hi
    return bounds


def shard_range(S, rank, world):
    return shard_bounds(S, world)[rank]


def all_gather_rows(local, S, group=None):
    """local: [S_r, C] tensor of this rank's rows (in shard order) -> [S, C] on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    bounds = shard_bounds(S, world)
    pad = max(hi - lo for lo, hi in bounds)
    buf = torch.zeros(pad, *local.shape[1:], dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    out = torch.empty(world * pad, *local.shape[1:], dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf, group=group) if local.is_cuda else \
        dist.all_gather(list(out.view(world, pad, *local.shape[1:]).unbind(0)), buf, group=group)
    out = out.view(world, pad, *local.shape[1:])
    return torch.cat([out[r, : hi - lo] for r, (lo, hi) in enumerate(bounds)], dim=0)


def upsample_sharded(generator, d_cloud, d_seeds, group=None):
    """Run Generator3D6.displace_device on this rank's seed range and all-gather the result.
    d_cloud [N,3] f64 and d_seeds [S,3] f64 are replicated device tensors."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    S = d_seeds.shape[0]
    lo, hi = shard_range(S, rank, world)
    local = generator.displace_device(d_cloud, d_seeds[lo:hi].contiguous())
    return all_gather_rows(local, S, group)
