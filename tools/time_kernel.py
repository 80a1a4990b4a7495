"""Kernel-only timing of one BASELINE workload (CUDA events, record buffers
preallocated).  Usage: python tools/time_kernel.py [config] [rays] [iters] [record_e]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import pyrate_b200 as pb  # noqa: E402
from pyrate_b200 import _native, configs, engine, lowering  # noqa: E402

if os.environ.get("PYR_TOOLS_LIB"):
    _native.use_tools_library(os.environ["PYR_TOOLS_LIB"])     # a variant build (make tools TOOLS_OUT=...)
elif os.environ.get("PYR_LEAN_VARIANT") or os.environ.get("PYR_DEBUG_RECORD_LAST"):
    _native.use_tools_library()      # the A/B knobs exist in the `make tools` build only

name = sys.argv[1] if len(sys.argv) > 1 else "c2_doublegauss"
rays = int(sys.argv[2]) if len(sys.argv) > 2 else 0
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
record_e = bool(int(sys.argv[4])) if len(sys.argv) > 4 else False
spec = configs.CONFIGS[name]
rings = configs.rings_for(rays) if rays else spec["bundle"]["rings"]
(x0, k0, e0) = configs.config_bundle(spec, rings)
(s, seq) = configs.build_system(spec, pb.api())
lowered = lowering.lower(s, seq, configs.DLINE)
dev = torch.device("cuda", 0)
(x0, k0, e0) = engine.device_bundle(x0, k0, e0, dev)
pool = engine.RecordPool()
for _ in range(3):
    engine.trace(lowered, x0, k0, e0, configs.DLINE, device=dev, pool=pool, record_e=record_e)
torch.cuda.synchronize()
ms = []
for _ in range(iters):
    ev = []
    rec = engine.trace(lowered, x0, k0, e0, configs.DLINE, device=dev, pool=pool,
                       record_e=record_e, events=ev)
    torch.cuda.synchronize()
    ms.append(sum(a.elapsed_time(b) for (a, b) in ev))
ms.sort()
n = x0.shape[1]
nsteps = len(lowered)
alive = int(((rec.flags[-1] & 2) != 0).sum())
print("%s variant=%s rays=%d entries=%d alive_end=%d  median %.4f ms  min %.4f ms  -> %.3e ray-entries/s (median)" %
      (name, os.environ.get("PYR_LEAN_VARIANT", "-"), n, nsteps, alive, ms[len(ms) // 2], ms[0],
       n * nsteps / (ms[len(ms) // 2] * 1e-3)))
