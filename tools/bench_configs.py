#!/usr/bin/env python
"""The other BASELINE.json configurations in their stated multi-GPU form (bench.py keeps
the headline C2 line the driver reads):

  C3 demo_asphere      1e7 rays, 1 GPU
  C4 anisotropic       1e6 rays in total, ray-sharded over the ranks (o/e split -> 4e6)
  C5 GRIN              1e8 rays in total, ray-sharded + NCCL gather of the spot points

  python tools/bench_configs.py --config c5_grin --rays-total 100000000 [--steps K]
  torchrun --nproc-per-node N tools/bench_configs.py --config ...

STRONG scaling: the total ray count is fixed, rank r traces the contiguous shard
[r n/W, (r+1) n/W) (pyrate_b200.distributed.shard_range), generated shard-wise on the
host.  One JSON line on rank 0; time = max over ranks of CUDA-event time per step.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

DEFAULT_TOTAL = {"c1_doublet": 1000519, "c2_doublegauss": 9997351, "c3_asphere": 9997351,
                 "c4_anisotropic": 1000519, "c5_grin": 99999907}
# algorithmic bytes per (output) ray-entry, SURVEY 8(d)
REC_BYTES = {"c4_anisotropic": 121.0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c5_grin")
    ap.add_argument("--rays-total", type=int, default=0)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--gather", action="store_true", help="also gather the spot points on rank 0")
    args = ap.parse_args()

    import numpy as np
    import torch
    import torch.distributed as dist
    import pyrate_b200 as pb
    from pyrate_b200 import configs, engine, lowering
    from pyrate_b200 import distributed as pd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    spec = configs.CONFIGS[args.config]
    total = args.rays_total or DEFAULT_TOTAL.get(args.config, 1000519)
    rings = configs.rings_for(total)
    total = configs.hexapolar_count(rings)
    (lo, hi) = pd.shard_range(total, rank, world)
    b = spec["bundle"]
    (x0h, k0h, e0h) = configs.collimated_shard(rings, b["radius"], b["z0"], lo, hi)
    (s, seq) = configs.build_system(spec, pb.api())
    lowered = lowering.lower(s, seq, configs.DLINE)
    (x0, k0, e0) = engine.device_bundle(x0h, k0h, e0h, dev)
    origin = engine.last_surface_origin(lowered)
    pool = engine.RecordPool()
    spot = torch.zeros(8, dtype=torch.float64, device=dev)

    def step(events=None):
        rec = engine.trace(lowered, x0, k0, e0, configs.DLINE, device=dev, pool=pool,
                           events=events)
        spot.zero_()
        last = rec.hit[-1]
        engine.spot_sums(last, rec.flags[-1], out=spot, shift=origin)
        pd.allreduce_spot_sums(spot)
        pts = pd.gather_spot_points(last, rec.flags[-1]) if args.gather else None
        return rec, pts

    for _ in range(max(args.warmup, 3)):
        (rec, pts) = step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    ev = []
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        (rec, pts) = step(ev)
    t1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([t0.elapsed_time(t1) / args.steps], dtype=torch.float64, device=dev)
    launches = len(ev) // args.steps
    kms = torch.tensor([sum(a.elapsed_time(b_) for (a, b_) in ev) / args.steps],
                       dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kms, op=dist.ReduceOp.MAX)
    out_entries = sum(rec.n_out)                       # output ray-entries of this rank
    tot_entries = torch.tensor([float(out_entries)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot_entries)
    (c, rms) = engine.spot_from_sums(spot.cpu(), origin)
    if rank == 0:
        line = {"config": args.config, "n_gpus": world, "rays_total": total,
                "entries": len(lowered), "scaling": "strong",
                "ms_per_step": float(ms.item()), "trace_kernels_ms": float(kms.item()),
                "launches_per_step": launches,
                "ray_entries_per_s": float(tot_entries.item()) / (float(ms.item()) * 1e-3),
                "rays_per_s": total / (float(ms.item()) * 1e-3),
                "spot_rms": rms, "spot_count": float(spot[3].item()),
                "gathered_points": None if pts is None else int(pts.shape[1])}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
