"""Ray sharding and the one collective of the path.

Rays are independent: rank r of W traces the contiguous range
[r*n/W, (r+1)*n/W) of the bundle with the replicated step table; nothing is
exchanged during the trace.  The only cross-ray operation downstream of
`seqtrace` is the spot statistic of RayBundleAnalysis
(analysis/ray_analysis.py:44-86): every rank reduces its shard to 8 partial sums
on the device (pyr_spot_sums) and ONE all-reduce of those 64 bytes (NCCL over
NVLink/NVSwitch on the GPU box, gloo in the CPU tests) gives the global
centroid / RMS.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous, balanced partition: returns (lo, hi) of `rank`."""
    lo = (n * rank) // world
    hi = (n * (rank + 1)) // world
    return lo, hi


def shard_bundle(x0, k0, e0, rank=None, world=None):
    """Slice (3, n) arrays / tensors to this rank's rays (views, no copy)."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    (lo, hi) = shard_range(x0.shape[1], rank, world)
    return (x0[:, lo:hi], k0[:, lo:hi], None if e0 is None else e0[:, lo:hi],
            (lo, hi))


def allreduce_spot_sums(sums):
    """In-place sum of the 8 partial sums over all ranks (no-op for 1 rank)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    return sums


def global_spot(x_last, flags_last=None, shift=None):
    """Centroid and RMS spot radius of the surviving rays of ALL ranks.

    x_last: (3, n_local) CUDA tensor (rows may be strided), flags_last: (n_local)
    uint8 PYR_RAY_* flags or None, shift: common reference point of the sums
    (e.g. engine.last_surface_origin).  Returns (centroid[3], rms, count)."""
    from . import engine
    sums = engine.spot_sums(x_last, flags_last, shift=shift)
    allreduce_spot_sums(sums)
    host = sums.cpu()
    (c, rms) = engine.spot_from_sums(host, shift)
    return c, rms, float(host[3])


def gather_spot_points(x_last, flags_last=None, dst=0):
    """Spot-diagram points (x, y of surviving rays) of all ranks on `dst`
    (16 B/ray over NVLink); returns a (2, N) tensor on dst, None elsewhere."""
    xy = x_last[:2]
    if flags_last is not None:
        xy = xy[:, (flags_last & 2) != 0]
    xy = xy.contiguous()
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return xy
    world = dist.get_world_size()
    count = torch.tensor([xy.shape[1]], dtype=torch.int64, device=xy.device)
    counts = [torch.zeros_like(count) for _ in range(world)]
    dist.all_gather(counts, count)
    width = int(max(int(c.item()) for c in counts))
    padded = torch.zeros((2, width), dtype=xy.dtype, device=xy.device)
    padded[:, :xy.shape[1]] = xy
    bufs = [torch.empty_like(padded) for _ in range(world)] \
        if dist.get_rank() == dst else None
    dist.gather(padded, bufs, dst=dst)
    if dist.get_rank() != dst:
        return None
    return torch.cat([b[:, :int(c.item())] for (b, c) in zip(bufs, counts)], dim=1)
