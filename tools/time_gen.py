"""Fused generate-and-trace against the trace of resident arrays (kernel time, CUDA events)
and the host entry with a generator (wall time per call).
Usage: python tools/time_gen.py [config] [rays] [iters]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import pyrate_b200 as pb  # noqa: E402
from pyrate_b200 import _native as _nat_for_variant  # noqa: E402
if os.environ.get('PYR_TOOLS_LIB'):
    _nat_for_variant.use_tools_library(os.environ['PYR_TOOLS_LIB'])
from pyrate_b200 import bundlegen, configs, engine, lowering  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2_doublegauss"
rays = int(sys.argv[2]) if len(sys.argv) > 2 else 0
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
spec = configs.CONFIGS[name]
rings = configs.rings_for(rays) if rays else spec["bundle"]["rings"]
gen = bundlegen.config_generator(spec, rings)
(s, seq) = configs.build_system(spec, pb.api())
lowered = lowering.lower(s, seq, configs.DLINE)
dev = torch.device("cuda", 0)
pool = engine.RecordPool()


def timed(**kw):
    for _ in range(3):
        engine.trace(lowered, kw.get("x0"), kw.get("k0"), kw.get("e0"), configs.DLINE, device=dev,
                     pool=pool, gen=kw.get("gen"))
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        ev = []
        engine.trace(lowered, kw.get("x0"), kw.get("k0"), kw.get("e0"), configs.DLINE, device=dev,
                     pool=pool, gen=kw.get("gen"), events=ev)
        torch.cuda.synchronize()
        ms.append(sum(a.elapsed_time(b) for (a, b) in ev))
    ms.sort()
    return ms[len(ms) // 2], ms[0]


n = gen.n
(g_med, g_min) = timed(gen=gen)
assert not gen.materialised
(x0, k0, e0) = bundlegen.config_generator(spec, rings).materialise(dev)
(m_med, m_min) = timed(x0=x0, k0=k0, e0=e0)
ns = len(lowered)
print("%s %d rays x %d entries: generated %.4f ms (min %.4f) = %.0f GB/s of records | from memory "
      "%.4f ms (min %.4f)" % (name, n, ns, g_med, g_min, n * ns * 49.0 / g_med / 1e6, m_med, m_min))
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
buf = torch.empty((9, engine._round_up(n, 16)), dtype=torch.float64, device=dev)
g2 = bundlegen.config_generator(spec, rings)
for _ in range(3):
    g2._cache = None
    g2.materialise(dev)
torch.cuda.synchronize()
ev[0].record()
for _ in range(iters):
    g2._cache = None
    g2.materialise(dev)
ev[1].record()
torch.cuda.synchronize()
print("pyr_generate_bundle alone (incl. zero-fill of the buffer): %.4f ms" % (ev[0].elapsed_time(ev[1]) / iters))
if all(ls.st.before.kind == 0 and ls.st.after.kind == 0 for ls in lowered) or name == "c5_grin":
    for chunk in (1 << 19, 1 << 20, 1 << 21):
        ht = engine.HostTracer(lowered, n, chunk_rays=chunk, device=dev)
        for _ in range(3):
            ht(gen=gen)
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(iters):
            ht(gen=gen)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t) / iters
        print("host entry, generator, chunk %8d: %.3f ms per call (D2H %d MB -> %.1f GB/s)" %
              (chunk, 1e3 * dt, ht.d2h_bytes >> 20, ht.d2h_bytes / dt / 1e9))
