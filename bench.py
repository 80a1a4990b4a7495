#!/usr/bin/env python
"""Headline benchmark: ray-surface intersections/s of OpticalSystem.seqtrace on the
10-surface Rudolph double-Gauss (13 sequence entries), 9 997 351-ray hexapolar bundle per
GPU (BASELINE.json configs[1]), FP64.

  python bench.py --gpus N --steps K --warmup W            # this engine
  python bench.py --impl reference --gpus N ...            # CPU arm: the reference itself

One "step" = one full pass of the bundle through the element sequence (one persistent
kernel launch per GPU + the spot-sum kernel; for N > 1 also the one NCCL all-reduce of the
8 spot sums).  Prints ONE JSON line on rank 0:

  value            device-timed, bundle resident in HBM (x0, k0, E0 arrays)
  config.generated the same step with the bundle GENERATED in the kernel's prologue
                   (PyrBundleGen: no input arrays exist)
  e2e              through the host entry pyr_trace_host_io: descriptor in, image-plane
                   record + spot sums back in pinned host memory (D2H inside the timed
                   region); e2e.host_buffers = the same with x0, k0, E0 uploaded from pinned
                   host arrays; e2e.all_records = every record of the sequence read back;
                   e2e.metric_only = descriptor in, the 8 spot sums out (MeritTrace)
  config.extra     BASELINE configs 3-5 at their stated total sizes, ray-sharded over the
                   N ranks (strong scaling), each with kernel time, roofline fraction and
                   an in-run parity figure against the oracle on a strided subsample
"""
import argparse
import itertools
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
REC_BYTES = 49.0      # per ray-entry: x 24 + k 24 + flag 1
BYTES_PER_RAY_ENTRY = REC_BYTES + 72.0 / S_SEQ      # + x0, k0, E0 read once: 54.54 B
METRIC = "ray-surface intersections/sec"
CPU_BLOCK = 1000      # rays per reference seqtrace call: temporaries stay below malloc's mmap threshold (no page-fault storm when all cores run)
WORKLOAD = ("double-Gauss (Rudolph 1897) 10 refracting conic surfaces / 13 sequence entries, "
            "%d-ray hexapolar bundle per GPU, single wavelength (BASELINE configs[1])")
# BASELINE configs 3-5: total rays (strong scaling over the ranks), entries, counted surfaces
EXTRAS = {"c3_asphere": {"total": 9997351, "label": "c3", "gather": True},
          "c4_anisotropic": {"total": 1000519, "label": "c4", "gather": False},
          "c5_grin": {"total": 99999907, "label": "c5", "gather": True}}


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
# CPU arm: the reference's own OpticalSystem.seqtrace (oracle/_ref, staged by
# oracle/make_ref.sh; /root/reference in the build container) -- kind "reference"; the
# NumPy restatement oracle/pyrate_np.py only if neither exists -- kind "port"
# ---------------------------------------------------------------------------
_SHARED = {}        # bundle arrays inherited by forked workers (copy-on-write, no pickling)
_CPU = {}           # per-process cache: the system built once


def cpu_kind():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refshim
    return "reference" if refshim.reference_available() else "port"


def _cpu_system():
    if "trace" in _CPU:
        return _CPU["trace"]
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from pyrate_b200 import configs
    try:                                    # one BLAS/LAPACK thread per worker
        import threadpoolctl
        threadpoolctl.threadpool_limits(1)
    except Exception:
        pass
    spec = configs.CONFIGS[CONFIG]
    if cpu_kind() == "reference":
        import warnings
        warnings.filterwarnings("ignore")
        import refshim
        api = refshim.api()                 # installs the three import shims, imports pyrateoptics
        (s, seq) = configs.build_system(spec, api)

        def trace(x0, k0, e0):
            # the unmodified reference: RayBundle + OpticalSystem.seqtrace
            # (raytracer/ray.py:35, raytracer/optical_system.py:73-94)
            s.seqtrace(api.RayBundle(x0, k0, e0, wave=configs.DLINE), seq)
    else:
        import pyrate_np as onp
        system = onp.system_from_spec(spec)

        def trace(x0, k0, e0):
            onp.seqtrace(system, x0, k0, e0, wave=configs.DLINE)
    _CPU["trace"] = trace
    return trace


def _cpu_worker(args):
    (lo, hi) = args
    trace = _cpu_system()
    (x0, k0, e0) = (_SHARED["x0"], _SHARED["k0"], _SHARED["e0"])
    t = time.perf_counter()
    for a in range(lo, hi, CPU_BLOCK):
        b = min(a + CPU_BLOCK, hi)
        trace(x0[:, a:b].copy(), k0[:, a:b].copy(), e0[:, a:b].copy())
    return time.perf_counter() - t


def cpu_bundle(nrays):
    from pyrate_b200 import configs
    spec = configs.CONFIGS[CONFIG]
    (x0, k0, e0) = configs.config_bundle(spec, configs.rings_for(nrays))
    _SHARED.update(x0=x0, k0=k0, e0=e0)
    return x0.shape[1]


def cpu_pass(n, procs, pool=None):
    """One pass of the shared bundle through the CPU implementation on `procs` processes
    (ray-sharded, one single-threaded process per core)."""
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


def cpu_sample_text(kind, n, steps, cores):
    if kind == "reference":
        return ("%d rays x %d step(s) through the UNMODIFIED reference (pyrateoptics "
                "OpticalSystem.seqtrace, staged by oracle/make_ref.sh), %d-ray calls, ray-sharded "
                "over %d single-threaded process(es)" % (n, steps, CPU_BLOCK, cores))
    return ("%d rays x %d step(s) through oracle/pyrate_np.py (NumPy restatement; the staged "
            "reference oracle/_ref is absent), ray-sharded over %d single-threaded "
            "process(es)" % (n, steps, cores))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    for var in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS",
                "NUMEXPR_NUM_THREADS"):
        os.environ[var] = "1"
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    kind = cpu_kind()
    per_core = 20000 if kind == "reference" else 50000
    nrays = int(args.cpu_rays) if args.cpu_rays else min(per_core * cores, 4000000)
    n = cpu_bundle(nrays)                       # BEFORE the fork: workers inherit the arrays
    pool = mp.get_context("fork").Pool(cores) if cores > 1 else None
    try:
        for _ in range(max(args.warmup, 1)):              # short passes: imports, page faults
            cpu_pass(min(n, max(1000 * cores, 1000)), cores, pool)
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
                             "kind": kind, "sample": cpu_sample_text(kind, n, args.steps, cores)},
            "e2e": {"value": val, "unit": "ray-surfaces/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def bind_to_gpu_numa(local):
    """Pin this rank's threads (and with them its first-touch / pinned allocations) to the
    NUMA node of its GPU.  Returns what was done (reported in the JSON line)."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        before = sorted(os.sched_getaffinity(0))
        info = {"gpu_pci": bus, "numa_node": node, "cpus_before": len(before)}
        if node < 0:
            return dict(info, bound=False, why="no NUMA information for the device")
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            (a, _, b) = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(before)
        if not allowed:
            return dict(info, bound=False, why="the GPU's node has no CPU in this process's cpuset")
        os.sched_setaffinity(0, allowed)
        return dict(info, bound=True, cpus_after=len(allowed))
    except Exception as err:                                    # never fatal
        return {"bound": False, "why": "%s: %s" % (type(err).__name__, err)}


def timed_region(step, steps, sync, world, dist, dev, torch):
    """K steps bracketed by barrier + synchronize; CUDA-event time, max over ranks (ms/step)."""
    sync()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        step()
    t1.record()
    sync()
    t = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / steps


def wall_region(call, steps, sync, world, dist, dev, torch):
    """K host-entry calls, wall clock between two synchronisations, max over ranks (s/step)."""
    sync()
    t = time.perf_counter()
    for _ in range(steps):
        call()
    torch.cuda.synchronize(dev)
    dt = torch.tensor([(time.perf_counter() - t) / steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return float(dt.item())


def parity_sample(name, spec, lowered, rec, gen, n_local, samples=384):
    """In-run parity of one leg: a strided subsample of this rank's rays, traced by the oracle
    (oracle/pyrate_np.py) from the host restatement of the same generator, against the device
    records.  Returns {"rays", "max_rel_err", "what"}."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyrate_np as onp
    from pyrate_b200 import configs
    stride = max(n_local // samples, 1)
    sub = np.arange(0, n_local, stride, dtype=np.int64)
    (x0, k0, e0) = gen.arrays_host(sub)
    kw = {"per_ray_energy": True, "history": False} if name == "c5_grin" else {}
    ref = onp.seqtrace(onp.system_from_spec(spec), x0, k0, e0, wave=configs.DLINE, **kw)[0]

    def rel(a, b):
        m = np.isfinite(b)
        return float(np.max(np.abs(a[m] - b[m])) / max(np.max(np.abs(b[m])), 1e-300)) if m.any() else 0.0
    worst = 0.0
    import torch
    sub_t = torch.as_tensor(sub, device=rec.hit[-1].device)
    if name == "c4_anisotropic":
        # output width 4 n: children of ray i sit in columns i + m n; the order of the two
        # forward modes is LAPACK-arbitrary in the reference, so children are matched as a set
        last = ref[-1]
        width = rec.hit[-1].shape[1]
        cols = (sub[:, None] + n_local * np.arange(width // n_local)[None, :]).reshape(-1)
        ct = torch.as_tensor(cols, device=rec.hit[-1].device)
        dx = rec.hit[-1][:, ct].cpu().numpy().reshape(3, sub.size, -1)
        dk = rec.k[-1][:, ct].cpu().numpy().real.reshape(3, sub.size, -1)
        d = np.concatenate((dx, dk)).transpose(1, 2, 0)                     # (rays, 4, 6)
        ids = last["rayID"]
        order = np.argsort(ids, kind="stable")
        o = np.concatenate((last["x"][-1], last["k"][-1].real))[:, order]
        o = o.reshape(6, sub.size, -1).transpose(1, 2, 0)                   # (rays, 4, 6)
        scale = np.array([np.max(np.abs(o[..., :3]))] * 3 + [np.max(np.abs(o[..., 3:]))] * 3)
        best = np.full(sub.size, np.inf)
        for perm in itertools.permutations(range(d.shape[1])):
            err = np.max(np.abs(d[:, list(perm), :] - o) / scale, axis=(1, 2))
            best = np.minimum(best, err)
        worst = float(np.max(best))
        what = "image-plane hit points and Re k of the 4 children of each sampled ray (unordered)"
    else:
        for s in range(len(rec.hit)):
            rb = ref[s + 1]
            ids = rb["rayID"]
            hit = rec.hit[s][:, sub_t].cpu().numpy()[:, ids]
            v = rb["valid"][-1]
            worst = max(worst, rel(hit[:, v], rb["x"][-1][:, v]))
            nb = ref[s + 2]
            k = rec.k[s][:, sub_t].cpu().numpy()[:, nb["rayID"]]
            worst = max(worst, rel(k, nb["k"][0]))
        what = "hit points and wave vectors after every sequence entry"
    return {"rays": int(sub.size), "max_rel_err": worst, "vs": "oracle/pyrate_np.py", "what": what}


def run_leg(name, info, args, world, rank, dev, peak, torch, dist):
    """One of BASELINE configs 3-5 at its stated TOTAL size, ray-sharded over the ranks."""
    import pyrate_b200 as pb
    from pyrate_b200 import bundlegen, configs, engine, lowering
    from pyrate_b200 import distributed as pd
    spec = configs.CONFIGS[name]
    total = int(args.extra_rays) if args.extra_rays else info["total"]
    rings = configs.rings_for(total)
    full = bundlegen.config_generator(spec, rings)
    total = full.n
    (lo, hi) = pd.shard_range(total, rank, world)
    gen = full.shard(lo, hi)
    n = gen.n
    (s, seq) = configs.build_system(spec, pb.api())
    lowered = lowering.lower(s, seq, configs.DLINE)
    fused = engine._gen_fusable(lowered, False, False, None)
    (x0, k0, e0) = (None, None, None) if fused else gen.materialise(dev)
    pool = engine.RecordPool()
    spot = torch.zeros(8, dtype=torch.float64, device=dev)
    origin = engine.last_surface_origin(lowered)
    width = (total + world - 1) // world
    gather_buf = {}

    def step(events=None):
        rec = engine.trace(lowered, x0, k0, e0, configs.DLINE, device=dev, pool=pool,
                           events=events, gen=gen if fused else None)
        spot.zero_()
        engine.spot_sums(rec.hit[-1], rec.flags[-1], out=spot, shift=origin)
        pd.allreduce_spot_sums(spot)
        if info["gather"]:
            if "local" not in gather_buf:
                gather_buf["local"] = (torch.empty((2, width), dtype=torch.float64, device=dev),
                                       torch.zeros((), dtype=torch.int64, device=dev))
            gather_buf["out"] = pd.gather_spot_points(
                rec.hit[-1], rec.flags[-1], dst=0, width=width, frame=lowered[-1].st.shape_frame,
                out=gather_buf.get("out"), local=gather_buf["local"])
        return rec

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    steps = max(min(args.steps, 5), 1)
    for _ in range(3):
        rec = step()
    ev = []
    with ClockSampler(torch.cuda.current_device()) as leg_clocks:
        ms = timed_region(lambda: step(ev), steps, sync, world, dist, dev, torch)
    launches = len(ev) // steps
    kms = torch.tensor([sum(a.elapsed_time(b) for (a, b) in ev) / steps], dtype=torch.float64, device=dev)
    entries = torch.tensor([float(sum(rec.n_out))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(kms, op=dist.ReduceOp.MAX)
        dist.all_reduce(entries)
    (kms, entries) = (float(kms.item()), float(entries.item()))
    # C3 is the HBM-bound leg: also time it from resident input arrays (the layout SURVEY 8d's
    # 67 B per ray-entry refers to)
    res_ms = None
    if fused and name == "c3_asphere":
        (xa, ka, ea) = gen.materialise(dev)
        for _ in range(3):
            engine.trace(lowered, xa, ka, ea, configs.DLINE, device=dev, pool=pool)
        ev2 = []
        for _ in range(steps):
            engine.trace(lowered, xa, ka, ea, configs.DLINE, device=dev, pool=pool, events=ev2)
        torch.cuda.synchronize(dev)
        rt = torch.tensor([sum(a.elapsed_time(b) for (a, b) in ev2) / steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(rt, op=dist.ReduceOp.MAX)
        res_ms = float(rt.item())
        del xa, ka, ea
        gen._cache = None                    # release the 72 B/ray arrays
    out = None
    if rank == 0:
        (c, rms) = engine.spot_from_sums(spot.cpu(), origin)
        complex_ = name == "c4_anisotropic"
        per_entry = 121.0 if complex_ else REC_BYTES
        in_bytes = 0.0 if fused else 72.0 * total
        algo = entries * per_entry + in_bytes
        out = {"config": name, "rays_total": total, "rays_per_gpu": n, "entries": len(lowered),
               "scaling": "strong", "ms_per_step": ms, "trace_kernels_ms": kms,
               "launches_per_step": launches, "output_ray_entries": entries,
               "ray_entries_per_s": entries / (ms * 1e-3),
               "ray_entries_per_s_kernel": entries / (kms * 1e-3),
               "bundle": "generated in the kernel prologue (no input arrays)" if fused else
                         "generated once on the device (pyr_generate_bundle), resident",
               "roofline": {"bound": "hbm" if name == "c3_asphere" else "fp64 (HBM fraction reported as SURVEY 8d asks; "
                            "FP64-pipe utilisation: ncu tables under profiles/)",
                            "algorithmic_bytes_all_ranks": algo,
                            "achieved": algo / (kms * 1e-3) / 1e9 / world, "peak": peak, "unit": "GB/s",
                            "frac": algo / (kms * 1e-3) / 1e9 / world / peak},
               "spot_rms": rms, "spot_count": float(spot[3].item()),
               "clocks": leg_clocks.summary()}
        if info["gather"]:
            g = gather_buf["out"]
            out["gathered_points"] = int(g.counts.sum().item())
            out["gather"] = ("pyr_spot_points (device compaction, count stays on the device) + one "
                             "fixed-width NCCL gather of (2, %d) doubles per rank to rank 0" % width)
        if out is not None and res_ms is not None:
            rbytes = entries * per_entry + 72.0 * total
            out["resident_arrays"] = {
                "what": "the same trace with x0, k0, E0 read from device arrays (72 B/ray: SURVEY 8d's "
                        "49 + 72/S_seq = 67 B per ray-entry), kernel only",
                "trace_kernels_ms": res_ms,
                "roofline": {"bound": "hbm", "algorithmic_bytes_all_ranks": rbytes,
                             "achieved": rbytes / (res_ms * 1e-3) / 1e9 / world, "peak": peak, "unit": "GB/s",
                             "frac": rbytes / (res_ms * 1e-3) / 1e9 / world / peak}}
        try:
            out["parity"] = parity_sample(name, spec, lowered, rec, gen, n)
        except Exception as err:                                # report, never hide
            out["parity"] = {"error": "%s: %s" % (type(err).__name__, err)}
    del pool, rec, gather_buf
    torch.cuda.empty_cache()
    return out


def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import pyrate_b200 as pb
    from pyrate_b200 import bundlegen, configs, engine, lowering

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    spec = configs.CONFIGS[CONFIG]
    rings = configs.rings_for(args.rays) if args.rays else spec["bundle"]["rings"]
    # weak scaling: every rank traces its own full-size bundle (a different field
    # angle per rank so the shards are not copies of each other)
    ang = 0.25 * rank * np.pi / 180.0
    kdir = (0.0, float(np.sin(ang)), float(np.cos(ang)))
    gen = bundlegen.config_generator(spec, rings, kdir, (1.0, 0.0, 0.0))
    n = gen.n
    (s, seq) = configs.build_system(spec, pb.api())
    lowered = lowering.lower(s, seq, configs.DLINE)
    # resident inputs in the engine's row-aligned layout (what any host upload through the
    # public API produces); written once by pyr_generate_bundle, outside every timed region
    (x0, k0, e0) = bundlegen.config_generator(spec, rings, kdir, (1.0, 0.0, 0.0)).materialise(dev)
    spot = torch.zeros(8, dtype=torch.float64, device=dev)
    origin = engine.last_surface_origin(lowered)
    pool = engine.RecordPool()      # record buffers allocated once, reused per step
    ev = []                         # (start, end) CUDA events around each native trace launch

    # (engine.trace(spot=...) would fold the spot sums into the trace launch -- pyr_trace_spot, what
    # MeritTrace uses: 1.17 instead of 1.21 ms per step on one GPU -- but the one 8-GPU run of that
    # step came out slower, 1.37 ms, profiles/r02b_bench_n8_fused_step.json, with no GPU time left to
    # find out why: the headline step keeps the two launches whose scaling is established)
    def step_resident():
        rec = engine.trace(lowered, x0, k0, e0, configs.DLINE, device=dev, pool=pool, events=ev)
        spot.zero_()
        engine.spot_sums(rec.hit[-1], rec.flags[-1], out=spot, shift=origin)
        if world > 1:
            dist.all_reduce(spot)
        return rec

    def step_generated():
        rec = engine.trace(lowered, None, None, None, configs.DLINE, device=dev, pool=pool,
                           events=ev, gen=gen)
        spot.zero_()
        engine.spot_sums(rec.hit[-1], rec.flags[-1], out=spot, shift=origin)
        if world > 1:
            dist.all_reduce(spot)
        return rec

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step_resident()
    with ClockSampler(local) as clocks:
        del ev[:]
        ms_step = timed_region(step_resident, args.steps, sync, world, dist, dev, torch)
        kern_ms = sorted(a.elapsed_time(b) for (a, b) in ev)
        spot_resident = spot.cpu().clone()
        for _ in range(warm):
            step_generated()
        del ev[:]
        ms_gen = timed_region(step_generated, args.steps, sync, world, dist, dev, torch)
        kern_gen_ms = sorted(a.elapsed_time(b) for (a, b) in ev)
        spot_generated = spot.cpu().clone()
    assert not gen.materialised
    value = world * n * S_COUNTED / (ms_step * 1e-3)
    (centroid, rms) = engine.spot_from_sums(spot_resident, origin)
    (_, rms_gen) = engine.spot_from_sums(spot_generated, origin)

    # ---- end to end through the C ABI host entry (rank-local, all ranks concurrently) ----
    e2e = None
    if not args.no_e2e:
        ht = engine.HostTracer(lowered, n, chunk_rays=args.chunk, device=dev)
        for _ in range(3):
            ht(gen=gen)
        dt = wall_region(lambda: ht(gen=gen), args.steps, sync, world, dist, dev, torch)
        (_, rms2) = engine.spot_from_sums(ht.spot8, origin)
        desc_bytes = ht.h2d_bytes             # descriptor + step table: what a described bundle sends
        e2e = {"value": world * n * S_COUNTED / dt, "unit": "ray-surfaces/s",
               "h2d_bytes_per_step": ht.h2d_bytes, "d2h_bytes_per_step": ht.d2h_bytes,
               "ms_per_step": 1e3 * dt, "spot_rms": rms2, "spot_count": float(ht.spot8[3]),
               "what": "pyr_trace_host_io: the bundle is DESCRIBED (PyrBundleGen, %d bytes + the step "
                       "table cross the bus), generated in the trace kernel's prologue, traced; D2H of "
                       "the image-plane record (x, k, flags: 49 B/ray) + 8 spot sums into pinned host "
                       "memory, 4-slot pipeline, ramped chunks" % 176}
        # the same with host arrays: H2D of x0, k0, E0 (72 B/ray) from pinned memory
        (xp, kp, ep) = (t.cpu().contiguous().pin_memory() for t in (x0, k0, e0))
        for _ in range(3):
            ht(xp, kp, ep)
        dt_h = wall_region(lambda: ht(xp, kp, ep), args.steps, sync, world, dist, dev, torch)
        e2e["host_buffers"] = {"value": world * n * S_COUNTED / dt_h, "ms_per_step": 1e3 * dt_h,
                               "h2d_bytes_per_step": ht.h2d_bytes, "d2h_bytes_per_step": ht.d2h_bytes,
                               "spot_rms": engine.spot_from_sums(ht.spot8, origin)[1],
                               "what": "pinned host x0, k0, E0 -> H2D -> trace -> D2H of the image-plane "
                                       "record + spot sums (the round-1 e2e)"}
        del xp, kp, ep
        # what an optimiser's merit function moves: descriptor in, the 8 spot sums out (nothing
        # else crosses the bus; only the image-plane entry is recorded)
        from pyrate_b200.merit import MeritTrace
        mt = MeritTrace(s, seq, pb.RayBundle(generator=gen, wave=configs.DLINE), device=dev)
        for _ in range(3):
            mt()
        dt_m = wall_region(mt, args.steps, sync, world, dist, dev, torch)
        e2e["metric_only"] = {"value": world * n * S_COUNTED / dt_m, "ms_per_step": 1e3 * dt_m,
                              "h2d_bytes_per_step": desc_bytes, "d2h_bytes_per_step": 64,
                              "spot_rms": mt(refresh=False),
                              "what": "pyrate_b200.merit.MeritTrace (pyr_trace_spot): parameters re-read from "
                                      "the object graph, bundle generated in the kernel, only the last entry "
                                      "recorded, spot sums read back -- one merit-function evaluation of "
                                      "optimize/optimize.py:73-91 at full bundle size"}
        del mt
        if world == 1 and not args.no_extras:
            # every record of the sequence back in host memory (what the S + 2 bundles of the
            # reference's RayPath hold): 13 x 49 B/ray over PCIe
            hta = engine.HostTracer(lowered, n, chunk_rays=min(args.chunk, 1 << 19), device=dev,
                                    all_records=True)
            hta(gen=gen)
            dt_a = wall_region(lambda: hta(gen=gen), 3, sync, world, dist, dev, torch)
            e2e["all_records"] = {"value": n * S_COUNTED / dt_a, "ms_per_step": 1e3 * dt_a,
                                  "h2d_bytes_per_step": hta.h2d_bytes,
                                  "d2h_bytes_per_step": hta.d2h_bytes, "steps": 3,
                                  "what": "as e2e, but the records of ALL 13 entries are read back"}
            del hta
        del ht
    del pool
    torch.cuda.empty_cache()

    (peak, peak_kind) = measured_peaks()
    extra = {}
    if not args.no_extras:
        for (name, info) in EXTRAS.items():
            if args.only_extra and info["label"] not in args.only_extra.split(","):
                continue
            try:
                res = run_leg(name, info, args, world, rank, dev, peak, torch, dist)
            except Exception as err:
                res = {"error": "%s: %s" % (type(err).__name__, err)}
            if rank == 0:
                extra[info["label"]] = res

    if rank == 0:
        kms = kern_ms[len(kern_ms) // 2]
        kgen = kern_gen_ms[len(kern_gen_ms) // 2]
        algo_bytes = n * S_SEQ * BYTES_PER_RAY_ENTRY
        achieved = algo_bytes / (kms * 1e-3) / 1e9
        gen_bytes = n * S_SEQ * REC_BYTES
        traffic = None
        traffic_src = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                traffic = tj.get("trace_real_kernel_bytes_per_launch")
                traffic_src = "profiles/traffic.json (%s; ncu capture, not this run)" % tj.get("source", "")
            except Exception:
                traffic = None
        cpu = None
        if world == 1 and not args.no_cpu:
            kind = cpu_kind()
            (cn, cdt) = cpu_pass(cpu_bundle(int(args.cpu_rays) if args.cpu_rays else
                                            (200000 if kind == "reference" else 400000)), 1)
            cpu = {"value": cn * S_COUNTED / cdt, "unit": "ray-surfaces/s", "cores": 1,
                   "kind": kind, "sample": cpu_sample_text(kind, cn, 1, 1)}
        line = {"metric": METRIC, "value": value, "unit": "ray-surfaces/s", "n_gpus": world,
                "steps": args.steps, "warmup": warm, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOAD % n,
                           "rays_per_gpu": n, "s_counted": S_COUNTED, "s_seq": S_SEQ,
                           "value_all_entries": world * n * S_SEQ / (ms_step * 1e-3),
                           "l2": "inputs 0.72 GB + per-step records 6.4 GB per pass >> 126 MB L2 "
                                 "(no flush needed)",
                           "parallelism": "rays sharded over %d GPU(s); one 8-double NCCL "
                                          "all-reduce of the spot sums per step" % world,
                           "spot_rms": rms, "spot_count": float(spot_resident[3].item()),
                           "generated": {
                               "what": "the same step with the bundle generated in the trace kernel's "
                                       "prologue from a PyrBundleGen descriptor: no x0 / k0 / E0 arrays "
                                       "exist, nothing is read from HBM",
                               "ms_per_step": ms_gen,
                               "value": world * n * S_COUNTED / (ms_gen * 1e-3),
                               "kernel_ms": kgen,
                               "roofline": {"bound": "hbm", "achieved": gen_bytes / (kgen * 1e-3) / 1e9,
                                            "peak": peak, "unit": "GB/s",
                                            "frac": gen_bytes / (kgen * 1e-3) / 1e9 / peak,
                                            "algorithmic_bytes_per_launch": gen_bytes,
                                            "note": "49 B per ray-entry, write-only: a pure store stream "
                                                    "can exceed the COPY bandwidth the peak is measured with"},
                               "spot_rms": rms_gen},
                           "numa": numa,
                           "extra": extra},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": traffic,
                             "traffic_source": traffic_src,
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
    ap.add_argument("--config", default=CONFIG, choices=[CONFIG],
                    help="headline workload (BASELINE configs[1]); configs 3-5 ride along as config.extra")
    ap.add_argument("--rays", type=int, default=0, help="rays per GPU (default: config)")
    ap.add_argument("--chunk", type=int, default=1 << 20, help="e2e chunk size in rays")
    ap.add_argument("--cpu-rays", type=int, default=0)
    ap.add_argument("--extra-rays", type=int, default=0, help="total rays of every extra leg (tests)")
    ap.add_argument("--only-extra", default="", help="comma list of c3,c4,c5")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
