"""Practical HBM ceilings on this GPU for the access mixes of the trace kernels:
pure write (fill), copy (read+write), and a 78-row write stream like C2's record
pattern (13 entries x 6 rows)."""
import torch

dev = torch.device("cuda", 0)
n = 9997360
rows = 78


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


buf = torch.empty((rows, n), dtype=torch.float64, device=dev)
src = torch.empty((rows, n), dtype=torch.float64, device=dev).normal_()
gb = buf.numel() * 8 / 1e9
t = timeit(lambda: buf.fill_(1.5))
print("pure write  fill_ %.2f GB: %.3f ms -> %.0f GB/s" % (gb, t, gb / t * 1e3))
t = timeit(lambda: buf.copy_(src))
print("copy (r+w)  %.2f GB moved: %.3f ms -> %.0f GB/s" % (2 * gb, t, 2 * gb / t * 1e3))
t = timeit(lambda: torch.cuda.memset if False else buf.zero_())
print("pure write  zero_ %.2f GB: %.3f ms -> %.0f GB/s" % (gb, t, gb / t * 1e3))
small = torch.empty((9, n), dtype=torch.float64, device=dev).normal_()
t = timeit(lambda: small.sum())
print("pure read   sum %.2f GB: %.3f ms -> %.0f GB/s" % (small.numel() * 8 / 1e9, t,
                                                          small.numel() * 8 / 1e9 / t * 1e3))
