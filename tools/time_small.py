"""Host-side latency of the public API on small bundles (the optimiser-loop use case):
python tools/time_small.py [config] [rings]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import pyrate_b200 as pb  # noqa: E402
from pyrate_b200 import configs, engine, lowering  # noqa: E402
from pyrate_b200.raytracer.analysis.ray_analysis import RayBundleAnalysis  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2_doublegauss"
rings = int(sys.argv[2]) if len(sys.argv) > 2 else 18
spec = configs.CONFIGS[name]
(x0, k0, e0) = configs.config_bundle(spec, rings)
(s, seq) = configs.build_system(spec, pb.api())
bundle = pb.RayBundle(x0, k0, e0, wave=configs.DLINE)
dev_bundle = bundle.to("cuda")


def timeit(fn, n=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e3


low = lowering.lower(s, seq, configs.DLINE)
pool = engine.RecordPool()
(xd, kd, ed) = engine.device_bundle(x0, k0, e0)
print("%s, %d rays, %d entries" % (name, x0.shape[1], len(low)))
print("lowering.lower            %.3f ms" % timeit(lambda: lowering.lower(s, seq, configs.DLINE)))
print("engine.trace (device in, pooled records) %.3f ms" %
      timeit(lambda: engine.trace(low, xd, kd, ed, configs.DLINE, pool=pool)))
print("engine.trace (device in)  %.3f ms" % timeit(lambda: engine.trace(low, xd, kd, ed, configs.DLINE)))
print("seqtrace (device bundle)  %.3f ms" % timeit(lambda: s.seqtrace(dev_bundle, seq)))
print("seqtrace (host bundle)    %.3f ms" % timeit(lambda: s.seqtrace(bundle, seq)))


def merit():
    p = s.seqtrace(dev_bundle, seq)[0]
    return RayBundleAnalysis(p.raybundles[-1]).get_rms_spot_size_centroid()


print("seqtrace + RMS spot (merit function) %.3f ms" % timeit(merit))


from pyrate_b200.merit import MeritTrace  # noqa: E402
from pyrate_b200 import bundlegen  # noqa: E402

mt = MeritTrace(s, seq, dev_bundle)
print("MeritTrace (resident arrays): refresh + trace + spot sums + read-back  %.3f ms" % timeit(mt))
print("MeritTrace (resident arrays): refresh only                          %.3f ms" % timeit(mt.refresh))
print("MeritTrace (resident arrays): trace + spot sums + read-back          %.3f ms" %
      timeit(lambda: mt(refresh=False)))
mg = MeritTrace(s, seq, pb.RayBundle(generator=bundlegen.config_generator(spec, rings), wave=configs.DLINE))
print("MeritTrace (generated bundle): refresh + trace + spot sums + read-back %.3f ms" % timeit(mg))
print("rms %.12g vs general API %.12g" % (mt(), merit()))
