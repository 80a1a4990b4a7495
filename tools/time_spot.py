import sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyrate_b200 import engine
n = 9997351
ld = (n + 15) // 16 * 16
x = torch.randn((3, ld), dtype=torch.float64, device="cuda")
f = torch.full((ld,), 3, dtype=torch.uint8, device="cuda")
out = torch.zeros(8, dtype=torch.float64, device="cuda")
big = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for i in range(12):
    big.zero_()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); engine.spot_sums(x[:, :n], f[:n], out=out, shift=[0, 0, 0]); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ts.sort(); print("spot_sums 1e7 rays: median %.4f ms min %.4f ms -> %.0f GB/s" % (ts[6], ts[0], 25.0 * n / ts[6] * 1e-6))
