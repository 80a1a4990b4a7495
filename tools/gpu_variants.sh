#!/bin/bash
for v in 0 41 43 3; do PYR_LEAN_VARIANT=$v timeout 300 python tools/time_kernel.py c2_doublegauss 0 20; done
PYR_LEAN_VARIANT=41 timeout 300 python - <<'PY'
# parity of the TMA-store variant against the default one (bit for bit)
import os, sys
sys.path.insert(0, os.getcwd())
import torch
import pyrate_b200 as pb
from pyrate_b200 import configs, engine, lowering
for name, rings in (("c2_doublegauss", 300), ("x1_tilted", 101), ("x3_vignette", 77)):
    spec = configs.CONFIGS[name]
    (x0, k0, e0) = configs.config_bundle(spec, rings)
    (s, seq) = configs.build_system(spec, pb.api())
    low = lowering.lower(s, seq, configs.DLINE)
    rec = engine.trace(low, x0, k0, e0, configs.DLINE)
    torch.cuda.synchronize()
    import subprocess, pickle
    ref = {k: [t.cpu() for t in getattr(rec, k)] for k in ("hit", "k", "flags")}
    torch.save(ref, "/tmp/tma_%s.pt" % name)
print("saved")
PY
PYR_LEAN_VARIANT=0 timeout 300 python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import torch
import pyrate_b200 as pb
from pyrate_b200 import configs, engine, lowering
for name, rings in (("c2_doublegauss", 300), ("x1_tilted", 101), ("x3_vignette", 77)):
    spec = configs.CONFIGS[name]
    (x0, k0, e0) = configs.config_bundle(spec, rings)
    (s, seq) = configs.build_system(spec, pb.api())
    low = lowering.lower(s, seq, configs.DLINE)
    rec = engine.trace(low, x0, k0, e0, configs.DLINE)
    ref = torch.load("/tmp/tma_%s.pt" % name)
    ok = True
    for k in ("hit", "k", "flags"):
        for (a, b) in zip(getattr(rec, k), ref[k]):
            a = a.cpu()
            if a.dtype.is_floating_point:
                ok = ok and torch.equal(torch.nan_to_num(a), torch.nan_to_num(b))
            else:
                ok = ok and torch.equal(a, b)
    print(name, "TMA-store variant == default:", ok)
PY
