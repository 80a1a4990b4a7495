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


class SpotPoints(object):
    """Spot-diagram points of all ranks on the destination rank: `xy` (world, 2, width)
    and `counts` (world,) int64, both on the device; columns [0, counts[r]) of block r are
    rank r's points.  Nothing here synchronises with the host; `points()` does (it has to
    know the total to allocate)."""

    def __init__(self, xy, counts):
        (self.xy, self.counts) = (xy, counts)

    def points(self):
        c = [int(v) for v in self.counts.tolist()]
        return torch.cat([self.xy[r, :, :c[r]] for r in range(len(c))], dim=1)


def gather_fixed_width(xy, count, dst=0, out=None):
    """Gather every rank's (2, width) point buffer and its point count on `dst` (one
    gather of fixed-width buffers + one of the counts: no host round trip, no ragged
    sizes on the wire).  `width` must be the same on all ranks.  Returns SpotPoints on
    dst, None elsewhere; a single process returns its own buffer."""
    count = count.reshape(1)
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return SpotPoints(xy[None], count)
    world = dist.get_world_size()
    me = dist.get_rank()
    if me == dst:
        if out is None:
            out = SpotPoints(torch.empty((world,) + tuple(xy.shape), dtype=xy.dtype, device=xy.device),
                             torch.empty((world,), dtype=count.dtype, device=count.device))
        dist.gather(xy, list(out.xy.unbind(0)), dst=dst)
        dist.gather(count, list(out.counts.split(1)), dst=dst)
        return out
    dist.gather(xy, None, dst=dst)
    dist.gather(count, None, dst=dst)
    return None


def gather_spot_points(x_last, flags_last=None, dst=0, width=None, frame=None, out=None,
                       local=None):
    """Spot diagram of all ranks on `dst`: each rank compacts (x, y) of its surviving
    rays on the device (pyr_spot_points: block-aggregated atomic cursor, count stays on
    the device) into a buffer of `width` columns (default: this rank's ray count -- pass
    the largest shard size when shards differ), then ONE fixed-width gather over
    NVLink (16 B/ray) plus the counts.  `frame`: PyrFrame of the last surface (the
    reference's get_spot reports local coordinates), None = global.  `local` / `out`:
    reusable buffers (engine.spot_points output / SpotPoints) for benchmark loops."""
    from . import engine
    n = x_last.shape[1]
    width = n if width is None else int(width)
    (xy, count) = engine.spot_points(x_last, flags_last, frame=frame, width=width, out=local)
    return gather_fixed_width(xy, count, dst=dst, out=out)
