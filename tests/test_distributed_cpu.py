"""CPU (gloo, world_size 2): the multi-rank host logic -- ray sharding, the
8-double all-reduce of the spot sums and the reference's centroid / RMS
normalisations.  The per-rank partial sums are formed with torch on the CPU
here (on the GPU box they come from pyr_spot_sums)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, xfile, out):
    sys.path.insert(0, ROOT)
    from pyrate_b200 import distributed as pd
    from pyrate_b200 import engine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x = torch.from_numpy(np.load(xfile))
    (xs, _, _, (lo, hi)) = pd.shard_bundle(x, x, None)
    sums = torch.zeros(8, dtype=torch.float64)
    sums[0:3] = xs.sum(1)
    sums[3] = xs.shape[1]
    sums[4:7] = (xs * xs).sum(1)
    pd.allreduce_spot_sums(sums)
    (c, rms) = engine.spot_from_sums(sums)
    np.save(out % rank, np.array(list(c) + [rms, sums[3].item(), lo, hi]))
    # the optional spot-diagram gather (BASELINE config 5: "NCCL spot gather"): surviving
    # rays of every rank, ragged widths, on rank 0
    # (the compaction itself is the native kernel pyr_spot_points on the GPU box; here the
    # fixed-width buffer + count are made with torch and the COLLECTIVE logic is tested)
    flags = torch.full((xs.shape[1],), 3, dtype=torch.uint8)
    flags[rank::3] = 1                                    # HIT but not ALIVE: dropped
    keep = (flags & 2) != 0
    width = (x.shape[1] + world - 1) // world             # largest shard: same on all ranks
    xy = torch.zeros((2, width), dtype=torch.float64)
    xy[:, :int(keep.sum())] = xs[:2][:, keep]
    pts = pd.gather_fixed_width(xy, keep.sum().to(torch.int64), dst=0)
    if rank == 0:
        assert pts.xy.shape == (world, 2, width) and pts.counts.tolist() == \
            [int(((torch.arange(hi_ - lo_) % 3) != r).sum())
             for (r, (lo_, hi_)) in enumerate(pd.shard_range(x.shape[1], q, world) for q in range(world))]
        np.save(out % 99, pts.points().numpy())
    else:
        assert pts is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sharded_spot_matches_reference_statistic(tmp_path):
    g = np.load(os.path.join(ROOT, "tests", "golden", "spot.npz"))
    xfile = str(tmp_path / "x.npy")
    np.save(xfile, g["x"])
    out = str(tmp_path / "r%d.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, xfile, out), nprocs=2, join=True)
    n = g["x"].shape[1]
    for r in range(2):
        v = np.load(out % r)
        assert np.allclose(v[:3], g["centroid"], rtol=1e-12, atol=1e-14)
        assert np.isclose(v[3], float(g["rms"]), rtol=1e-9)
        assert v[4] == n
    (lo0, hi0) = np.load(out % 0)[5:7]
    (lo1, hi1) = np.load(out % 1)[5:7]
    assert (lo0, hi0, hi1) == (0, n // 2, n) and lo1 == hi0
    want = []
    for (r, (lo, hi)) in enumerate(((int(lo0), int(hi0)), (int(lo1), int(hi1)))):
        keep = np.ones(hi - lo, dtype=bool)
        keep[r::3] = False
        want.append(g["x"][:2, lo:hi][:, keep])
    assert np.array_equal(np.load(out % 99), np.concatenate(want, axis=1))


def test_shard_range_partitions():
    from pyrate_b200 import distributed as pd
    for n in (0, 1, 7, 9997351):
        for w in (1, 2, 3, 8):
            ranges = [pd.shard_range(n, r, w) for r in range(w)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            for (a, b) in zip(ranges[:-1], ranges[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for (lo, hi) in ranges]
            assert max(sizes) - min(sizes) <= 1


def test_hexapolar_shards_concatenate_to_the_full_raster():
    from pyrate_b200 import configs
    from pyrate_b200 import distributed as pd
    for rings in (1, 3, 40):
        (px, py) = configs.hexapolar(rings)
        n = px.size
        parts = [configs.hexapolar_range(rings, *pd.shard_range(n, r, 8)) for r in range(8)]
        assert np.allclose(np.concatenate([p[0] for p in parts]), px, atol=1e-15)
        assert np.allclose(np.concatenate([p[1] for p in parts]), py, atol=1e-15)
