"""Concurrent host<->device copy ceiling of the node: every rank copies pinned host
buffers to / from its GPU with bare cudaMemcpyAsync (torch copy_, non_blocking) at the
same time; reports per-rank and aggregate GB/s for H2D alone, D2H alone and both
directions together.  Run under torchrun with one rank per GPU:
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_ceiling.py
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import bind_to_gpu_numa  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
bind = bind_to_gpu_numa(local) if "--no-bind" not in sys.argv else {"bound": False, "why": "--no-bind"}
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
MB = 512
n = MB * (1 << 20) // 8
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
h_in.fill_(1.0)
d_in = torch.empty(n, dtype=torch.float64, device=dev)
d_out = torch.ones(n, dtype=torch.float64, device=dev)
s1 = torch.cuda.Stream(dev)
s2 = torch.cuda.Stream(dev)


def run(h2d, d2h, iters=6):
    def once():
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    once()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t = time.perf_counter()
    for _ in range(iters):
        once()
    torch.cuda.synchronize(dev)
    dt = (time.perf_counter() - t) / iters
    gb = (int(h2d) + int(d2h)) * n * 8 / 1e9
    mine = torch.tensor([gb / dt], dtype=torch.float64, device=dev)
    tmax = torch.tensor([dt], dtype=torch.float64, device=dev)
    rates = [torch.zeros_like(mine) for _ in range(world)]
    if world > 1:
        dist.all_gather(rates, mine)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    else:
        rates = [mine]
    return {"per_rank_gbs": [round(float(r.item()), 1) for r in rates],
            "aggregate_gbs": round(world * gb / float(tmax.item()), 1)}


res = {"world": world, "buffer_mb": MB, "h2d": run(True, False), "d2h": run(False, True),
       "both": run(True, True)}
binds = [None] * world
if world > 1:
    dist.all_gather_object(binds, bind)
else:
    binds = [bind]
res["binding"] = binds
try:
    res["affinity_rank0"] = len(os.sched_getaffinity(0))
    res["numa_nodes"] = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
except Exception:
    pass
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
