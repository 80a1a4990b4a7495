"""Short single-GPU target for ncu: a few launches of the trace kernel on a
BASELINE workload (default: the C2 double-Gauss bundle).  Usage:
   ncu ... python tools/profile_target.py [config] [rays] [launches]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import pyrate_b200 as pb  # noqa: E402
from pyrate_b200 import configs, engine, lowering  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2_doublegauss"
rays = int(sys.argv[2]) if len(sys.argv) > 2 else 0
launches = int(sys.argv[3]) if len(sys.argv) > 3 else 5
generated = len(sys.argv) > 4 and sys.argv[4] == "gen"      # bundle generated in the kernel prologue
spec = configs.CONFIGS[name]
rings = configs.rings_for(rays) if rays else spec["bundle"]["rings"]
(x0, k0, e0) = configs.config_bundle(spec, rings)
(s, seq) = configs.build_system(spec, pb.api())
lowered = lowering.lower(s, seq, configs.DLINE)
dev = torch.device("cuda", 0)
(x0, k0, e0) = engine.device_bundle(x0, k0, e0, dev)
gen = None
if generated:
    from pyrate_b200 import bundlegen
    gen = bundlegen.config_generator(spec, rings)
for _ in range(launches):
    if gen is not None:
        rec = engine.trace(lowered, None, None, None, configs.DLINE, device=dev, gen=gen)
    else:
        rec = engine.trace(lowered, x0, k0, e0, configs.DLINE, device=dev)
torch.cuda.synchronize()
print(name, x0.shape[1], "rays", launches, "launches ok")
