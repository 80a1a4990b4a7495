#!/usr/bin/env python
"""Headline benchmark: ray-surface intersections/s of OpticalSystem.seqtrace on
the 10-surface Rudolph double-Gauss (13 sequence entries), 9 997 351-ray
hexapolar bundle per GPU (BASELINE.json configs[1]), FP64.

  python bench.py --gpus N --steps K --warmup W            # this engine
  python bench.py --impl reference --gpus N ...            # CPU arm (oracle port)

One "step" = one full pass of the bundle through the element sequence (one
persistent kernel launch per GPU + the spot-sum kernel; for N > 1 also the one
NCCL all-reduce of the 8 spot sums).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIG = "c2_doublegauss"
S_SEQ = 13            # sequence entries (bytes move for every one of them)
S_COUNTED = 10        # refracting surfaces with a real shape (headline count)
BYTES_PER_RAY_ENTRY = 49.0 + 72.0 / S_SEQ      # 54.54 B
METRIC = "ray-surface intersections/sec"
CPU_BLOCK = 1000


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler(object):
    """SM clock and throttle reasons sampled DURING the timed region: NVML in a
    background thread (sub-millisecond period; nvidia-smi as a fallback)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
            0x4: "sw_power_cap"}

    def __init__(self, index=0):
        self.index = index
        self.sm = []
        self.max_mhz = None
        self.reasons = set()
        self.stop = False
        self.thread = None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        nv = self.nvml
        self.sm.append(int(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
        try:
            mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        except Exception:
            mask = 0
        for (bit, name) in self.BITS.items():
            if mask & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(
            ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
             "--format=csv,noheader,nounits"], capture_output=True, text=True,
            timeout=5).stdout.strip().splitlines()
        if not out:
            return
        c = [v.strip() for v in out[0].split(",")]
        if c[0].isdigit():
            self.sm.append(int(c[0]))
        if len(c) > 1 and c[1].isdigit():
            self.max_mhz = int(c[1])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for i in range(4):
            if len(c) >= 6 and c[2 + i].lower().startswith("active"):
                self.reasons.add(names[i])

    def _run(self):
        while not self.stop:
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            time.sleep(0.0005 if self.nvml is not None else 0.05)

    def __enter__(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.thread.join(timeout=6)

    def summary(self):
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(sm),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ---------------------------------------------------------------------------
# CPU arm: the oracle port (NumPy restatement of the reference's seqtrace)
# ---------------------------------------------------------------------------
_SHARED = {}        # bundle arrays inherited by forked workers (copy-on-write, no pickling)


def _cpu_worker(args):
    (lo, hi) = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyrate_np as onp
    from pyrate_b200 import configs
    try:                                    # one BLAS/LAPACK thread per worker
        import threadpoolctl
        threadpoolctl.threadpool_limits(1)
    except Exception:
        pass
    (x0, k0, e0) = (_SHARED["x0"], _SHARED["k0"], _SHARED["e0"])
    system = onp.system_from_spec(configs.CONFIGS[CONFIG])
    t = time.perf_counter()
    # cache-blocked: 1000-ray pieces keep every NumPy temporary below the
    # malloc mmap threshold (no page-fault storm when all cores run) -- the
    # fastest way we found to run the reference algorithm on the host
    for a in range(lo, hi, CPU_BLOCK):
        b = min(a + CPU_BLOCK, hi)
        onp.seqtrace(system, x0[:, a:b], k0[:, a:b], e0[:, a:b], wave=configs.DLINE)
    return time.perf_counter() - t


def cpu_bundle(nrays):
    from pyrate_b200 import configs
    spec = configs.CONFIGS[CONFIG]
    (x0, k0, e0) = configs.config_bundle(spec, configs.rings_for(nrays))
    _SHARED.update(x0=x0, k0=k0, e0=e0)
    return x0.shape[1]


def cpu_pass(n, procs, pool=None):
    """One pass of the shared bundle through the oracle port on `procs` processes
    (ray-sharded, one single-threaded NumPy process per core)."""
    # many small jobs: dynamic load balance across the cores
    njobs = procs if procs == 1 else min(max(procs * 4, 1), max(n // CPU_BLOCK, 1))
    bounds = [(i * n) // njobs for i in range(njobs + 1)]
    jobs = [(bounds[i], bounds[i + 1]) for i in range(njobs)]
    t = time.perf_counter()
    if procs == 1 or pool is None:
        for j in jobs:
            _cpu_worker(j)
    else:
        pool.map(_cpu_worker, jobs, chunksize=1)
    return n, time.perf_counter() - t


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # one process per core: BLAS/OpenMP pools must not spawn (and spin) a thread
    # per core inside every worker -- set before NumPy is first imported
    for var in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS",
                "NUMEXPR_NUM_THREADS"):
        os.environ[var] = "1"
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    nrays = int(args.cpu_rays) if args.cpu_rays else min(50000 * cores, 4000000)
    n = cpu_bundle(nrays)                       # BEFORE the fork: workers inherit the arrays
    pool = mp.get_context("fork").Pool(cores) if cores > 1 else None
    try:
        for _ in range(max(args.warmup, 1)):
            cpu_pass(min(n, max(2000 * cores, 2000)), cores, pool)
        total_t = 0.0
        for _ in range(args.steps):
            (n, dt) = cpu_pass(n, cores, pool)
            total_t += dt
    finally:
        if pool is not None:
            pool.close()
            pool.join()
    ms = 1e3 * total_t / args.steps
    val = n * S_COUNTED / (ms * 1e-3)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "ray-surfaces/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 1),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "double-Gauss (Rudolph 1897) 10 refracting conic "
                       "surfaces / 13 sequence entries, hexapolar bundle, single "
                       "wavelength; CPU sample of %d rays per step" % n,
                       "rays_per_step": n, "s_counted": S_COUNTED, "s_seq": S_SEQ,
                       "value_all_entries": n * S_SEQ / (ms * 1e-3)},
            "cpu_baseline": {"value": val, "unit": "ray-surfaces/s", "cores": cores,
                             "kind": "port",
                             "sample": "%d rays x %d steps through oracle/pyrate_np.py (NumPy "
                                       "restatement of the reference seqtrace incl. its "
                                       "per-refraction 3x3 SVD for E), ray-sharded over %d "
                                       "single-threaded processes; the Python reference itself "
                                       "cannot travel to the GPU box" % (n, args.steps, cores)},
            "e2e": {"value": val, "unit": "ray-surfaces/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import pyrate_b200 as pb
    from pyrate_b200 import configs, engine, lowering

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    spec = configs.CONFIGS[CONFIG]
    rings = configs.rings_for(args.rays) if args.rays else spec["bundle"]["rings"]
    # weak scaling: every rank traces its own full-size bundle (a different field
    # angle per rank so the shards are not copies of each other)
    ang = 0.25 * rank * np.pi / 180.0
    (x0h, k0h, e0h) = configs.config_bundle(spec, rings, (0.0, np.sin(ang), np.cos(ang)),
                                            (1.0, 0.0, 0.0))
    n = x0h.shape[1]
    (s, seq) = configs.build_system(spec, pb.api())
    lowered = lowering.lower(s, seq, configs.DLINE)
    # resident inputs in the engine's row-aligned layout (what any host upload
    # through the public API produces)
    (x0, k0, e0) = engine.device_bundle(x0h, k0h, e0h, dev)
    spot = torch.zeros(8, dtype=torch.float64, device=dev)
    origin = engine.last_surface_origin(lowered)
    pool = engine.RecordPool()      # record buffers allocated once, reused per step

    def step():
        rec = engine.trace(lowered, x0, k0, e0, configs.DLINE, device=dev, pool=pool)
        spot.zero_()
        engine.spot_sums(rec.hit[-1], rec.flags[-1], out=spot, shift=origin)
        if world > 1:
            dist.all_reduce(spot)
        return rec

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        rec = step()
    sync()
    # kernel-only timing of the trace launch (CUDA events on the launch stream)
    kern_ms = []
    ev = []          # (start, end) CUDA events recorded around each native trace launch
    with ClockSampler(local) as clocks:
        sync()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for i in range(args.steps):
            rec = engine.trace(lowered, x0, k0, e0, configs.DLINE, device=dev, pool=pool,
                               events=ev)
            spot.zero_()
            engine.spot_sums(rec.hit[-1], rec.flags[-1], out=spot, shift=origin)
            if world > 1:
                dist.all_reduce(spot)
        t1.record()
        sync()
        total_ms = t0.elapsed_time(t1)
        kern_ms = [a.elapsed_time(b) for (a, b) in ev]
    tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms = float(tmax.item())
    ms_step = total_ms / args.steps
    value = world * n * S_COUNTED / (ms_step * 1e-3)
    (centroid, rms) = engine.spot_from_sums(spot.cpu(), origin)

    # ---- end to end through the C ABI with host buffers (rank-local) ----
    e2e = None
    if not args.no_e2e:
        ht = engine.HostTracer(lowered, n, chunk_rays=args.chunk, device=dev)
        (xp, kp, ep) = (torch.from_numpy(x0h).pin_memory(), torch.from_numpy(k0h).pin_memory(),
                        torch.from_numpy(e0h).pin_memory())
        for _ in range(3):
            ht(xp, kp, ep)
        sync()
        t = time.perf_counter()
        for _ in range(args.steps):
            ht(xp, kp, ep)
        torch.cuda.synchronize(dev)
        dt = torch.tensor([(time.perf_counter() - t) / args.steps], dtype=torch.float64,
                          device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * n * S_COUNTED / float(dt.item()), "unit": "ray-surfaces/s",
               "h2d_bytes_per_step": ht.h2d_bytes, "d2h_bytes_per_step": ht.d2h_bytes,
               "ms_per_step": 1e3 * float(dt.item()),
               "what": "pyr_trace_host: pinned host x0,k0,E0 -> H2D -> trace -> D2H of the "
                       "image-plane record (x, k, flags) + 8 spot sums, 4-slot pipeline, "
                       "ramped chunks"}
        (c2, rms2) = engine.spot_from_sums(ht.spot8, origin)
        e2e["spot_rms"] = rms2
        e2e["spot_count"] = float(ht.spot8[3])
        if world == 1:
            # informational: the same call with E0=None, the reference's default field
            # (0,1,0) (ray.py:71-73; perpendicular to this on-axis bundle's k like the
            # uploaded (1,0,0), so the trace is the same), which is not uploaded: 48 B/ray
            ht(xp, kp, None)
            sync()
            t = time.perf_counter()
            for _ in range(args.steps):
                ht(xp, kp, None)
            torch.cuda.synchronize(dev)
            dt2 = (time.perf_counter() - t) / args.steps
            e2e["default_e0"] = {"value": n * S_COUNTED / dt2, "ms_per_step": 1e3 * dt2,
                                 "h2d_bytes_per_step": 48 * n,
                                 "spot_rms": engine.spot_from_sums(ht.spot8, origin)[1]}

    if rank == 0:
        (peak, peak_kind) = measured_peaks()
        kms = sorted(kern_ms)[len(kern_ms) // 2]
        algo_bytes = n * S_SEQ * BYTES_PER_RAY_ENTRY
        achieved = algo_bytes / (kms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("trace_real_kernel_bytes_per_launch")
            except Exception:
                traffic = None
        cpu = None
        if world == 1 and not args.no_cpu:
            (cn, cdt) = cpu_pass(cpu_bundle(int(args.cpu_rays) if args.cpu_rays else 200000), 1)
            # (single process; `--impl reference` times the all-cores variant)
            cpu = {"value": cn * S_COUNTED / cdt, "unit": "ray-surfaces/s", "cores": 1,
                   "kind": "port",
                   "sample": "%d rays of the same workload, one pass, oracle/pyrate_np.py "
                             "single process (NumPy restatement incl. 3x3 SVD for E)" % cn}
        line = {"metric": METRIC, "value": value, "unit": "ray-surfaces/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": "double-Gauss (Rudolph 1897) 10 refracting conic "
                           "surfaces / 13 sequence entries, %d-ray hexapolar bundle per GPU, "
                           "single wavelength (BASELINE configs[1])" % n,
                           "rays_per_gpu": n, "s_counted": S_COUNTED, "s_seq": S_SEQ,
                           "value_all_entries": world * n * S_SEQ / (ms_step * 1e-3),
                           "l2": "inputs 0.72 GB + per-step records 6.4 GB per pass >> 126 MB L2 "
                                 "(no flush needed)",
                           "parallelism": "rays sharded over %d GPU(s); one 8-double NCCL "
                                          "all-reduce of the spot sums per step" % world,
                           "spot_rms": rms, "spot_count": float(spot[3].item())},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": traffic,
                             "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)",
                             "kernel": "trace_real_kernel<2,false,0,2,3> (lean, TMA in/out)",
                             "kernel_ms": kms,
                             "algorithmic_bytes_per_launch": algo_bytes},
                "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": 2 * args.steps,
                "clocks": clocks.summary()}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rays", type=int, default=0, help="rays per GPU (default: config)")
    ap.add_argument("--chunk", type=int, default=1 << 20, help="e2e chunk size in rays")
    ap.add_argument("--cpu-rays", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
