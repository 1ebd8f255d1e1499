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
    even = all(hi - lo == pad for lo, hi in bounds)
    if even:
        buf = local.contiguous()                      # equal shards: gather straight into the result, no padding copy
    else:
        buf = torch.zeros(pad, *local.shape[1:], dtype=local.dtype, device=local.device)
        buf[: local.shape[0]] = local
    out = torch.empty(world * pad, *local.shape[1:], dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf, group=group) if local.is_cuda else \
        dist.all_gather(list(out.view(world, pad, *local.shape[1:]).unbind(0)), buf, group=group)
    if even:
        return out
    out = out.view(world, pad, *local.shape[1:])
    return torch.cat([out[r, : hi - lo] for r, (lo, hi) in enumerate(bounds)], dim=0)


def shard_batch_offsets(seed_off, lo, hi):
    """Prefix table of a batched problem restricted to the flat seed range [lo, hi): cloud b keeps the seeds
    [max(seed_off[b], lo), min(seed_off[b+1], hi)) -- possibly none -- and the clouds stay replicated."""
    import numpy as np
    so = np.clip(np.asarray(seed_off, dtype=np.int64), lo, hi) - lo
    return so


def upsample_sharded(generator, d_cloud, d_seeds, group=None, batch=None):
    """Run Generator3D6.displace_device on this rank's seed range and all-gather the result.
    d_cloud [N,3] f64 and d_seeds [S,3] f64 are replicated device tensors; batch = (cloud_off, seed_off) for a batch of
    independent clouds (the flattened (cloud, seed) list is what gets sharded)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    S = d_seeds.shape[0]
    lo, hi = shard_range(S, rank, world)
    if batch is not None:
        batch = (batch[0], shard_batch_offsets(batch[1], lo, hi))
    local = generator.displace_device(d_cloud, d_seeds[lo:hi].contiguous(), batch=batch)
    return all_gather_rows(local, S, group)


def upsample_sharded_host(generator, cloud, seeds, group=None, batch=None):
    """Host arrays in, host array out, on every rank: pinned H2D of the (replicated) cloud and of THIS rank's seed range,
    the device pipeline, the all-gather of the displaced points and the D2H copy of the gathered [S,3] result.
    This is the end-to-end call bench.py times (`e2e`)."""
    import numpy as np
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    S = seeds.shape[0]
    lo, hi = shard_range(S, rank, world)
    dev = generator.device
    with torch.cuda.device(dev):
        h_cloud = generator._pinned("cloud", cloud.shape, torch.float64)
        h_cloud.copy_(torch.from_numpy(np.ascontiguousarray(cloud)))
        h_seeds = generator._pinned("seeds", (hi - lo, 3), torch.float64)
        h_seeds.copy_(torch.from_numpy(np.ascontiguousarray(seeds[lo:hi])))
        d_cloud = h_cloud.to(dev, non_blocking=True)
        d_seeds = h_seeds.to(dev, non_blocking=True)
        b = None if batch is None else (batch[0], shard_batch_offsets(batch[1], lo, hi))
        local = generator.displace_device(d_cloud, d_seeds, batch=b)
        full = all_gather_rows(local, S, group)
        h_out = generator._pinned("out_full", (S, 3), torch.float64)
        h_out.copy_(full, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        from . import _native as N
        N.check_device("upsample_sharded_host")
    return h_out.numpy()


def outlier_keep_sharded(d_points, threshold=1.5, k=30, group=None, shard=None):
    """generation.py:176-183 over seeds sharded across ranks.  d_points [S,3] f64: the gathered displaced points (every rank
    holds all of them).  Each rank queries the k self-neighbours of ITS rows, the row means are all-gathered and summed in
    the single-GPU order, each rank masks its rows and the masks are all-gathered: returns the full boolean keep mask [S]
    (cuda), bit-identical to Generator3D6._outlier_filter's for any world size.  `shard` = (rank, world) overrides the
    process group (single-process emulation of the ranks in tests)."""
    import ctypes
    from . import _native as N
    L = N.lib()
    S = d_points.shape[0]
    if S < k:
        raise ValueError("k must be less than or equal to the number of training points")
    dev = d_points.device
    if shard is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        ranks = [rank]
    else:
        world, ranks = shard[1], list(range(shard[1])) if shard[0] is None else [shard[0]]
    means = {}
    with torch.cuda.device(dev):
        st = N.stream_ptr(dev)
        kws = torch.empty(L.sapcu_knn_workspace_bytes(S), dtype=torch.uint8, device=dev)
        for r in ranks:
            lo, hi = shard_range(S, r, world)
            q = d_points[lo:hi].contiguous()
            idx = torch.empty(hi - lo, k, dtype=torch.int32, device=dev)
            N.check(L.sapcu_knn(N.ptr(d_points), S, N.ptr(q), hi - lo, k, N.ptr(idx), N.ptr(kws), kws.numel(), st), "sapcu_knn(outlier shard)")
            m = torch.empty(hi - lo, dtype=torch.float64, device=dev)
            N.check(L.sapcu_knn_mean_dist(N.ptr(d_points), S, N.ptr(q), hi - lo, N.ptr(idx), k, N.ptr(m), st), "sapcu_knn_mean_dist")
            means[r] = m
        if shard is None:
            mean_all = all_gather_rows(means[ranks[0]].view(-1, 1), S, group).view(-1).contiguous()
        else:
            mean_all = torch.cat([means[r] for r in range(world)]).contiguous()      # emulated ranks: all shards computed here
        ws = torch.empty(256, dtype=torch.uint8, device=dev)
        keeps = {}
        for r in ranks:
            lo, hi = shard_range(S, r, world)
            keep = torch.empty(hi - lo, dtype=torch.uint8, device=dev)
            N.check(L.sapcu_outlier_mask_from_means(N.ptr(mean_all), S, lo, hi - lo, float(threshold), N.ptr(keep), N.ptr(ws), ws.numel(), st),
                    "sapcu_outlier_mask_from_means")
            keeps[r] = keep
        if shard is None:
            keep_all = all_gather_rows(keeps[ranks[0]].view(-1, 1), S, group).view(-1)
        else:
            keep_all = torch.cat([keeps[r] for r in range(world)])
    return keep_all.bool()
