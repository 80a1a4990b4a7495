"""End-to-end (host buffers) timing of pyr_trace_host against the PCIe ceiling:
chunk-size sweep on C2 plus plain pinned copies of the same byte counts.
python tools/e2e_sweep.py [rings]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import pyrate_b200 as pb  # noqa: E402
from pyrate_b200 import configs, engine, lowering  # noqa: E402

rings = int(sys.argv[1]) if len(sys.argv) > 1 else 1825
spec = configs.CONFIGS["c2_doublegauss"]
(x0, k0, e0) = configs.config_bundle(spec, rings)
n = x0.shape[1]
(s, seq) = configs.build_system(spec, pb.api())
low = lowering.lower(s, seq, configs.DLINE)
(hx, hk, he) = (torch.from_numpy(a).contiguous().pin_memory() for a in (x0, k0, e0))


def wall(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t) * 1e3)
    return sorted(ts)[len(ts) // 2]


# PCIe ceiling: the same bytes as plain pinned copies
dev_in = torch.empty((9, n), dtype=torch.float64, device="cuda")
dev_out = torch.empty((6, n), dtype=torch.float64, device="cuda")
host_out = torch.empty((6, n), dtype=torch.float64).pin_memory()
host_in = torch.empty((9, n), dtype=torch.float64).pin_memory()
(s1, s2) = (torch.cuda.Stream(), torch.cuda.Stream())


def h2d(rows):
    dev_in[:rows].copy_(host_in[:rows], non_blocking=True)


def both(rows):
    with torch.cuda.stream(s1):
        dev_in[:rows].copy_(host_in[:rows], non_blocking=True)
    with torch.cuda.stream(s2):
        host_out.copy_(dev_out, non_blocking=True)


for rows in (9, 6):
    t = wall(lambda: h2d(rows))
    print("H2D %d rows (%.0f MB) alone: %.2f ms = %.1f GB/s" % (rows, rows * n * 8e-6, t, rows * n * 8e-6 / t))
t = wall(lambda: host_out.copy_(dev_out, non_blocking=True))
print("D2H 6 rows (%.0f MB) alone: %.2f ms = %.1f GB/s" % (6 * n * 8e-6, t, 6 * n * 8e-6 / t))
for rows in (9, 6):
    t = wall(lambda: both(rows))
    print("H2D %d rows + D2H 6 rows concurrently: %.2f ms" % (rows, t))

for with_e in (True, False):
    for chunk in (1 << 18, 1 << 19, 1 << 20, 1 << 21):
        ht = engine.HostTracer(low, n, chunk_rays=chunk)
        t = wall(lambda: ht(hx, hk, he if with_e else None))
        print("pyr_trace_host E0 %s chunk %8d: %.2f ms  (%.3g ray-surfaces/s)" %
              ("uploaded" if with_e else "default ", chunk, t, n * 10 / t * 1e3))
        del ht
