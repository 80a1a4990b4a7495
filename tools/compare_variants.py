"""Bitwise comparison of a kernel variant (PYR_LEAN_VARIANT) against saved records.
   python tools/compare_variants.py save|check"""
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

mode = sys.argv[1]
ok_all = True
for (name, rings) in (("c2_doublegauss", 300), ("x1_tilted", 101), ("x3_vignette", 77),
                      ("c1_doublet", 8), ("c3_asphere", 200)):
    spec = configs.CONFIGS[name]
    (x0, k0, e0) = configs.config_bundle(spec, rings)
    (s, seq) = configs.build_system(spec, pb.api())
    low = lowering.lower(s, seq, configs.DLINE)
    rec = engine.trace(low, x0, k0, e0, configs.DLINE)
    torch.cuda.synchronize()
    cur = {k: [t.cpu() for t in getattr(rec, k)] for k in ("hit", "k", "flags")}
    fn = "/tmp/variant_ref_%s.pt" % name
    if mode == "save":
        torch.save(cur, fn)
        continue
    ref = torch.load(fn)
    ok = True
    for k in ("hit", "k", "flags"):
        for (a, b) in zip(cur[k], ref[k]):
            if a.dtype.is_floating_point:
                ok = ok and torch.equal(torch.nan_to_num(a), torch.nan_to_num(b))
            else:
                ok = ok and torch.equal(a, b)
    print(name, "variant", os.environ.get("PYR_TOOLS_LIB") or os.environ.get("PYR_LEAN_VARIANT"), "== default:", ok)
    ok_all = ok_all and ok
print("saved" if mode == "save" else ("ALL EQUAL" if ok_all else "MISMATCH"))
