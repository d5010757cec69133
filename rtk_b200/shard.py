"""Ray sharding across the GPUs of one box: the scene is replicated, rays are split
into contiguous ranges, there is no reduction, and the only exchange is the gather of the hit
records on rank 0.  One process per GPU, torch.distributed for the plumbing (NCCL on GPUs; the
CPU test tier drives the same code over gloo)."""
import numpy as np


def ray_range(rank, world, n):
    """Contiguous range [lo, hi) of rank `rank` out of `world` over n rays; ranges tile [0, n)."""
    return (n * rank) // world, (n * (rank + 1)) // world


def gather_records(local, n_total, dist, rank, world, device):
    """Gather per-rank record arrays (numpy structured, contiguous ranges in rank order) on rank 0.
    Ranges may differ in length by one, so records are padded to the longest range."""
    import torch
    itemsize = local.dtype.itemsize
    longest = max(ray_range(r, world, n_total)[1] - ray_range(r, world, n_total)[0] for r in range(world))
    buf = np.zeros((longest, itemsize), dtype=np.uint8)
    buf[:len(local)] = local.view(np.uint8).reshape(len(local), itemsize)
    t = torch.from_numpy(buf).to(device)
    out = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, out, dst=0)
    if rank != 0:
        return None
    parts = []
    for r in range(world):
        lo, hi = ray_range(r, world, n_total)
        parts.append(out[r][:hi - lo].cpu().numpy().reshape(-1).view(local.dtype))
    return np.concatenate(parts)


def trace_sharded(scene, rays, dist, rank, world, device="cpu"):
    """Every rank traces its own range of `rays` against its replica of `scene` (api.Scene) and
    rank 0 receives all expanded hits and masks, in ray order.  Returns (hits, mask) on rank 0,
    (None, None) elsewhere."""
    lo, hi = ray_range(rank, world, len(rays))
    hits, mask, _ = scene.trace_rays(np.ascontiguousarray(rays[lo:hi]))
    all_hits = gather_records(hits, len(rays), dist, rank, world, device)
    all_mask = gather_records(mask.view(np.dtype([("m", "u1")])), len(rays), dist, rank, world, device)
    if rank != 0:
        return None, None
    return all_hits, all_mask.view(np.uint8)
