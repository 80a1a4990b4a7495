"""Host driver of the native trace: allocates the per-step records as torch CUDA
tensors, fills the output pointers of the PyrStep table, launches pyr_trace()
(one launch per non-splitting stretch of the sequence) and wraps the records as
RayBundle / RayPath views with the reference's structure
(optical_system.py:73-94: path = [b0, b0, b1, ..., bS]).

PyTorch is plumbing here (device memory, streams); all ray arithmetic is in
libpyrate_b200.so.  No CPU fallback: without a CUDA device every entry point
raises.
"""
import ctypes as C

import numpy as np

import torch

from . import _native as nat
from . import lowering
from .raytracer.ray import RayBundle, RayPath, as_tensor

LD_ALIGN = 16           # doubles: rows start on 128-byte boundaries
MAX_STEPS_PER_LAUNCH = 40       # kMaxSteps
MAX_AUX_PER_LAUNCH = 10         # kMaxAux
MAX_SPLITS_PER_LAUNCH = 6       # kMaxSplits (csrc/pyr_aniso.cu)


def _needs_aux(st):
    """Mirror of pack() in csrc/pyr_trace.cu: how many DAux records this entry needs."""
    same_frame = st.aperture_kind == nat.AP_BASE or \
        bytes(st.shape_frame) == bytes(st.aperture_frame)
    if st.shape_kind == nat.SHAPE_COMBINATION:
        return int(st.n_terms)
    return int(st.shape_kind != nat.SHAPE_CONIC or not same_frame or
               st.mode != nat.STEP_FULL or st.before.kind != nat.MEDIUM_ISO_CONST or
               st.after.kind != nat.MEDIUM_ISO_CONST)


def bind_grid(st, device):
    """Grid-sag spline arrays of a lowered step (lowering leaves them in st._grid):
    upload once per device and point the step at them."""
    g = getattr(st, "_grid", None)
    if g is None:
        return
    cache = getattr(st, "_grid_dev", None)
    if cache is None or cache[0] != device:
        cache = (device, [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(device)
                          for a in g])
        st._grid_dev = cache
    (tx, ty, c) = cache[1]
    (st.grid_tx, st.grid_ty, st.grid_c) = (tx.data_ptr(), ty.data_ptr(), c.data_ptr())
    (st.grid_nx, st.grid_ny) = (tx.numel(), ty.numel())


class DeviceRequired(RuntimeError):
    pass


def require_cuda():
    if not torch.cuda.is_available():
        raise DeviceRequired(
            "pyrate_b200: a CUDA device is required (the trace runs only as "
            "sm_100a kernels; there is no CPU fallback)")
    lib = nat.load()
    return lib


def compute_device():
    """The CUDA device ray data lives on (raises without one: no CPU fallback)."""
    require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def _round_up(v, m):
    return (v + m - 1) // m * m


class TraceRecord(object):
    """Per-step device records of one seqtrace call.

    hit[s]   (3, n_in[s])   global hit point at sequence entry s
    k[s]     (3, n_out[s])  wave vector after the deflection (real or complex)
    e[s]     (3, n_out[s])  E field after the deflection, or None
    flags[s] (n_in[s])      PYR_RAY_HIT | PYR_RAY_ALIVE bits
    n_out[s] = 2 n_in[s] on a birefringent (splitting) step.
    """

    def __init__(self):
        self._x0 = self._k0 = self._e0 = None
        self.gen = None         # bundlegen.BundleGen of a generated trace (x0 / k0 / e0 lazy)
        self.n0 = 0
        self.device = None
        self.hit = []
        self.k = []
        self.e = []
        self.flags = []
        self.n_in = []
        self.n_out = []
        self.split = []
        self.lowered = None
        self.wave = None
        self.grin_hist = {}     # step index -> {"x": (M,3,n), "k": (M,3,n), "valid": (M,n), "count": (n)}

    def _initial(self, i):
        v = (self._x0, self._k0, self._e0)[i]
        if v is None and self.gen is not None:
            # the kernel generated the rays in registers; the caller wants to look at the
            # initial bundle: one pyr_generate_bundle launch, on first access only
            v = self.gen.materialise(self.device)[i]
        return v

    x0 = property(lambda self: self._initial(0), lambda self, v: setattr(self, "_x0", v))
    k0 = property(lambda self: self._initial(1), lambda self, v: setattr(self, "_k0", v))
    e0 = property(lambda self: self._initial(2), lambda self, v: setattr(self, "_e0", v))


def _empty(shape, dtype, device, pool):
    if pool is not None:
        return pool.get(shape, dtype, device)
    return torch.empty(shape, dtype=dtype, device=device)


def _alloc_rows(rows, n, device, complex_=False, pool=None):
    ld = _round_up(max(n, 1), LD_ALIGN)
    shape = (rows, 3, ld, 2) if complex_ else (rows, 3, ld)
    return _empty(shape, torch.float64, device, pool), ld


def _launch(lib, steps, lo, hi, x, k, e, alive, n, n_x, ld_in, flags, stream,
            events=None, wave_end=None, gen=None, device=None, spot=None):
    """spot = (out8 tensor, shift ctypes array or None): the launch also leaves the spot sums of its
    last entry in out8 (pyr_trace_spot, asynchronous form: fused into the trace kernel where the
    launch is a single conic-only kernel, a second kernel otherwise)."""
    if n == 0:
        if spot is not None:
            spot[0].zero_()
        return                      # empty bundle: nothing to trace
    arr = (nat.PyrStep * (hi - lo))()
    for i in range(lo, hi):
        if getattr(steps[i], "_grid", None) is not None:
            bind_grid(steps[i], x.device if x is not None else device)
        C.memmove(C.addressof(arr[i - lo]), C.addressof(steps[i]),
                  C.sizeof(nat.PyrStep))
    rin = nat.PyrRaysIn()
    if gen is not None:
        rin.gen = C.pointer(gen)    # PyrBundleGen: the kernel generates the rays itself
    else:
        rin.x = x.data_ptr()
        rin.k = k.data_ptr()
        rin.e = e.data_ptr() if e is not None else None
        rin.alive = alive.data_ptr() if alive is not None else None
    rin.ld = ld_in
    rin.n_x = n_x
    if wave_end is not None and len(wave_end) > 1:
        rin.n_waves = len(wave_end)
        for (w, end) in enumerate(wave_end):
            rin.wave_end[w] = int(end)
    if events is not None:
        # CUDA events recorded back to back with the launch on the launch stream:
        # the pair brackets the kernel itself, not the host-side packing
        ext = torch.cuda.ExternalStream(stream.value) if stream.value else \
            torch.cuda.current_stream()
        (a, b) = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        a.record(ext)
        if spot is not None:
            nat.check(lib.pyr_trace_spot(arr, hi - lo, C.byref(rin), n, flags, spot[1], spot[0].data_ptr(),
                                         None, stream))
        else:
            nat.check(lib.pyr_trace(arr, hi - lo, C.byref(rin), n, flags, stream))
        b.record(ext)
        events.append((a, b))
        return
    if spot is not None:
        nat.check(lib.pyr_trace_spot(arr, hi - lo, C.byref(rin), n, flags, spot[1], spot[0].data_ptr(),
                                     None, stream))
        return
    nat.check(lib.pyr_trace(arr, hi - lo, C.byref(rin), n, flags, stream))


def _padded(t, complex_=False, pad=True):
    """(3, n) tensor -> buffer the kernels can read, and its leading dimension.

    pad=False hands a contiguous tensor through unchanged (ld = n; the kernels
    fall back to 64-bit loads when rows are not 16-byte aligned) -- used for
    the caller's input bundle, which is never copied needlessly."""
    n = t.shape[1]
    if complex_:
        tc = t.to(torch.complex128)
        if not pad and tc.is_contiguous():
            return torch.view_as_real(tc), max(n, 1)
        ld = _round_up(max(n, 1), LD_ALIGN)
        buf = torch.zeros((3, ld, 2), dtype=torch.float64, device=t.device)
        buf[:, :n, :] = torch.view_as_real(tc)
        return buf, ld
    if t.is_complex():
        raise ValueError("complex data in a real-valued stretch")
    if not pad and t.is_contiguous():
        return t, max(n, 1)
    ld = _round_up(max(n, 1), LD_ALIGN)
    if n == ld and t.is_contiguous():
        return t, ld
    buf = torch.zeros((3, ld), dtype=torch.float64, device=t.device)
    buf[:, :n] = t
    return buf, ld


def device_bundle(x0, k0, e0, device=None, complex_=False, like_ld=None):
    """Bundle rows as the kernels want them: float64 (complex128 if `complex_`) CUDA
    tensors of shape (3, n) with unit column stride and ONE common leading dimension.

    Host data (NumPy / CPU tensors) is uploaded straight into row-padded buffers
    (ld a multiple of 16 doubles, rows 128-byte aligned: the layout that enables the
    TMA-staged input path) -- no extra pass.  Device tensors that already have a
    common row stride are used in place (the kernels fall back to plain loads
    when the rows are not 16-byte aligned); anything else is copied once."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None \
        else torch.device(device)
    ts = [None if t is None else as_tensor(t) for t in (x0, k0, e0)]
    n = ts[0].shape[1]

    def usable(t, want_complex):
        return (t.device == device and t.dim() == 2 and t.shape[0] == 3 and
                (t.stride(1) == 1 or n <= 1) and t.is_complex() == want_complex)

    wants = [False, complex_, complex_]
    strides = {t.stride(0) for (t, w) in zip(ts, wants) if t is not None and usable(t, w)}
    all_ok = all(t is None or usable(t, w) for (t, w) in zip(ts, wants))
    if all_ok and len(strides) == 1 and (like_ld is None or like_ld in strides) and n > 1:
        return tuple(ts)
    ld = like_ld if like_ld is not None else _round_up(max(n, 1), LD_ALIGN)
    out = []
    for (t, w) in zip(ts, wants):
        if t is None:
            out.append(None)
            continue
        if usable(t, w) and t.stride(0) == ld:
            out.append(t)
            continue
        buf = torch.empty((3, ld), dtype=torch.complex128 if w else torch.float64,
                          device=device)
        buf[:, n:] = 0
        src = t.to(torch.complex128) if w else t
        buf[:, :n].copy_(src, non_blocking=True)
        out.append(buf[:, :n])
    return tuple(out)


class RecordPool(object):
    """Opt-in reuse of the per-step record buffers across calls with identical
    shapes (optimisation loops, benchmarks): a TraceRecord returned with a pool
    is overwritten by the next trace that uses the same pool."""

    def __init__(self):
        self.bufs = {}
        self.cursor = 0

    def begin(self):
        self.cursor = 0

    def get(self, shape, dtype, device):
        key = (self.cursor, tuple(shape), dtype, str(device))
        self.cursor += 1
        t = self.bufs.get(key)
        if t is None:
            t = torch.empty(shape, dtype=dtype, device=device)
            self.bufs[key] = t
        return t


def _grin_lockstep(lib, st, x, k, e, alive, n, ld, device, stream, want_history):
    """pyr_grin_lockstep on the state handed over to a GRIN segment: returns (x, k, alive,
    hist) with fresh (3, ld) / (ld) buffers; hist = the reference's per-iteration rows."""
    xo = torch.empty((3, ld), dtype=torch.float64, device=device)
    ko = torch.empty((3, ld), dtype=torch.float64, device=device)
    ao = torch.zeros((ld,), dtype=torch.uint8, device=device)
    scratch = torch.empty((lib.pyr_grin_lockstep_scratch(ld),), dtype=torch.uint8, device=device)
    iters = torch.zeros((1,), dtype=torch.int32, device=device)

    def run(hx, hk, hv, rows):
        nat.check(lib.pyr_grin_lockstep(
            C.byref(st), x.data_ptr(), k.data_ptr(), e.data_ptr() if e is not None else None,
            alive.data_ptr() if alive is not None else None, ld, n, xo.data_ptr(), ko.data_ptr(),
            ao.data_ptr(), scratch.data_ptr(), iters.data_ptr(),
            hx.data_ptr() if hx is not None else None, hk.data_ptr() if hk is not None else None,
            hv.data_ptr() if hv is not None else None, rows, stream))
    run(None, None, None, 0)
    hist = None
    if want_history:
        m = int(iters.item())
        hist = {"x": torch.zeros((m, 3, ld), dtype=torch.float64, device=device),
                "k": torch.zeros((m, 3, ld), dtype=torch.float64, device=device),
                "valid": torch.zeros((m, ld), dtype=torch.uint8, device=device),
                "count": torch.full((n,), m, dtype=torch.int32, device=device),
                "rows": m, "n": n}
        if m > 0:
            run(hist["x"], hist["k"], hist["valid"], m)
    return xo, ko, ao, hist


_USER_GRIN = {}     # (id(material), device index) -> grin_jit.UserGrin (compiled + verified)


def _is_user_grin(medium):
    return medium.kind == nat.MEDIUM_ISO_GRIN and medium.grin_profile == nat.GRIN_USER


def _user_grin(material, medium, device):
    """The NVRTC-compiled kernels of a user-defined GRIN material (compiled once per source,
    checked against the material's Python functions once per material and device)."""
    from . import grin_jit
    key = (id(material), device.index)
    ug = _USER_GRIN.get(key)
    if ug is None or ug[0] is not material:
        u = grin_jit.UserGrin(material.annotations["device_source"], device)
        u.verify(material, medium.frame)
        ug = _USER_GRIN[key] = (material, u)
    return ug[1]


def _compose_to_shape(mat_frame, shape_frame):
    """Frame mapping material-local -> shape-local coordinates: x_s = Rs^T (Rm x_m + om - os)."""
    rm = np.asarray(list(mat_frame.r)).reshape(3, 3)
    rs = np.asarray(list(shape_frame.r)).reshape(3, 3)
    (om, os_) = (np.asarray(list(mat_frame.o)), np.asarray(list(shape_frame.o)))
    return ((rs.T @ rm).reshape(-1), rs.T @ (om - os_))


def _gen_fusable(lowered, record_e, grin_history, wave_end):
    """Can the trace kernel generate the rays itself (PyrRaysIn.gen)?  Real-valued,
    non-splitting sequences without E recording, grid-sag / combination shapes,
    wavelength batches or integrator history; otherwise the bundle is written out once
    (pyr_generate_bundle) and traced from the arrays."""
    if record_e or grin_history or wave_end is not None:
        return False
    for (i, ls) in enumerate(lowered):
        st = ls.st
        if nat.MEDIUM_ANISO in (st.before.kind, st.after.kind) or st.split:
            return False
        if _is_user_grin(st.before) or _is_user_grin(st.after):
            return False
        if st.shape_kind in (nat.SHAPE_GRIDSAG, nat.SHAPE_COMBINATION):
            return False
        if i > 0 and st.dir_mode == nat.DIR_POYNTING:
            return False
    return True


def trace(lowered, x0, k0, e0, wave, record_e=False, device=None, stream=None,
          pool=None, events=None, grin_history=False, _hist_rows=None, wave_end=None,
          gen=None, grin_lockstep=False, spot=None):
    """Run the lowered sequence on the device.  Returns a TraceRecord.

    spot: (out8, shift) -- also leave the spot sums of the last entry about `shift` (3 floats or
    None) in the float64 CUDA tensor out8 (overwritten; what `spot_sums` of the last record gives):
    one C call with the last launch (pyr_trace_spot), fused into the trace kernel where it can be.

    events: optional list; a (start, end) pair of CUDA timing events is appended
    per native launch (kernel-only timing for benchmarks).
    wave_end: wavelength batch (`lowering.lower_batch` table): cumulative ray counts of
    the bundles that lie back to back in x0 / k0 / e0; segment w uses the w-th media
    indices of every entry.  One launch for all wavelengths.
    grin_history: also record every integrator step of GRIN segments (the rows the
    reference appends, material_grin.py:198-205).  Memory grows with steps x rays:
    meant for small bundles.  Runs the trace twice (step counts first).
    grin_lockstep: GRIN segments run the reference's exact lock-step loop with its
    bundle-summed energy test (pyr_grin_lockstep, material_grin.py:139, :164-176) instead
    of integrating every ray independently -- for small bundles."""
    lib = require_cuda()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None \
        else torch.device(device)
    if gen is not None and (gen.materialised or
                            not _gen_fusable(lowered, record_e, grin_history or grin_lockstep,
                                             wave_end)):
        (x0, k0, e0) = gen.materialise(device)
        gen = None
    nsteps = len(lowered)
    rec = TraceRecord()
    rec.lowered = lowered
    rec.wave = wave
    rec.device = device
    if gen is not None:
        # `gen`: a bundlegen.BundleGen -- the rays are generated in the prologue of the
        # first launch; x0 / k0 / e0 of the record materialise on first access
        complex_in = False
        n0 = gen.n
        rec.gen = gen
        gen_desc = gen.descriptor(device)
    else:
        complex_in = bool(as_tensor(k0).is_complex() or
                          (e0 is not None and as_tensor(e0).is_complex()))
        (x0, k0, e0) = device_bundle(x0, k0, e0, device, complex_in)
        if x0.is_complex():
            raise ValueError("ray positions must be real")
        n0 = x0.shape[1]
        (rec.x0, rec.k0, rec.e0) = (x0, k0, e0)
        gen_desc = None
    rec.n0 = n0
    steps = [ls.st for ls in lowered]
    stream_ptr = C.c_void_p(torch.cuda.current_stream(device).cuda_stream
                            if stream is None else stream)

    # stretches: [lo, hi) with no ray doubling except possibly at hi-1; the
    # run turns complex at the first anisotropic medium and stays complex
    # (the reference's dtype behaviour)
    first_aniso = None
    for (i, ls) in enumerate(lowered):
        if ls.st.after.kind == nat.MEDIUM_ANISO or ls.st.before.kind == nat.MEDIUM_ANISO:
            first_aniso = i
            break
    # ONE launch per stretch: the complex-valued kernel walks the mode tree of every ray
    # itself (birefringent interfaces double the rays inside the launch), so the only cut
    # is where the run turns complex
    cuts = [0]
    if first_aniso is not None and first_aniso > 0:
        cuts.append(first_aniso)
    # one launch carries at most MAX_STEPS_PER_LAUNCH entries, MAX_AUX_PER_LAUNCH entries
    # with an auxiliary record (kernel parameter block, csrc/pyr_device.cuh) and
    # MAX_SPLITS_PER_LAUNCH doubling steps; longer sequences continue from the last
    # recorded state in a further launch
    if grin_lockstep:
        # the lock-step integrator is a kernel of its own in front of every GRIN segment
        cuts += [i for (i, ls) in enumerate(lowered)
                 if i > 0 and ls.st.before.kind == nat.MEDIUM_ISO_GRIN]
    user_grin = any(_is_user_grin(ls.st.before) or _is_user_grin(ls.st.after) for ls in lowered)
    if user_grin:
        # user-written index functions (NVRTC): the segment is integrated by its own kernel in
        # front of the exit step; the entrance step runs alone, in two phases, because the
        # refraction into the medium needs n at every ray's hit point
        if grin_history or grin_lockstep:
            raise lowering.LoweringError("grin_history / grin_lockstep: catalogue GRIN profiles only")
        for (i, ls) in enumerate(lowered):
            if _is_user_grin(ls.st.before) and i > 0:
                cuts.append(i)
            if _is_user_grin(ls.st.after):
                cuts += [i, i + 1] if i > 0 else [i + 1]
        cuts = [c for c in cuts if c < nsteps]
    cuts = sorted(set(cuts)) + [nsteps]
    bounded = [0]
    for hi in cuts[1:]:
        lo = bounded[-1]
        (count, aux, splits) = (0, 0, 0)
        for i in range(lo, hi):
            a = _needs_aux(lowered[i].st)
            sp = int(bool(lowered[i].st.split))
            if count + 1 > MAX_STEPS_PER_LAUNCH or aux + a > MAX_AUX_PER_LAUNCH or \
                    splits + sp > MAX_SPLITS_PER_LAUNCH:
                bounded.append(i)
                (count, aux, splits) = (0, 0, 0)
            count += 1
            aux += a
            splits += sp
        bounded.append(hi)
    cuts = bounded

    if pool is not None:
        pool.begin()
    with torch.cuda.device(device):
        # device_bundle() guarantees unit column stride and ONE leading dimension
        if gen is not None:
            (cur_x, ld_x, x0, k0) = (None, max(n0, 1), None, None)
        else:
            (cur_x, ld_x) = (x0, max(x0.stride(0), 1))
        if complex_in:
            (cur_k, ld_k) = (torch.view_as_real(k0), ld_x)
        else:
            (cur_k, ld_k) = (k0, ld_x)
        cur_e = None
        if gen is not None:
            pass
        elif e0 is not None:
            cur_e = torch.view_as_real(e0) if complex_in else e0
        elif complex_in or first_aniso is not None:
            tmp = torch.zeros((3, n0), dtype=torch.float64, device=device)
            tmp[1] = 1.0                      # ray.py:71-73 default
            (_, _, cur_e) = device_bundle(x0, x0, tmp, device, complex_in, like_ld=ld_x)
            e0 = cur_e
            cur_e = torch.view_as_real(cur_e) if complex_in else cur_e
        cur_alive = None
        n = n0
        n_x = n0
        is_complex = complex_in
        spot_done = False
        for ci in range(len(cuts) - 1):
            (lo, hi) = (cuts[ci], cuts[ci + 1])
            seg_complex = is_complex or (first_aniso is not None and lo >= first_aniso)
            if seg_complex and not is_complex:
                # promote the state handed over from the real stretch
                (cur_k, ld_k) = _padded(cur_k[:, :n], True)
                if cur_e is None:
                    raise RuntimeError("internal: E must be recorded ahead of an "
                                       "anisotropic medium")
                (cur_e, _) = _padded(cur_e[:, :n], True)
                is_complex = True
                if ld_k != ld_x:
                    nx = torch.zeros((3, ld_k), dtype=torch.float64, device=device)
                    nx[:, :n_x] = cur_x[:, :n_x]
                    (cur_x, ld_x) = (nx, ld_k)
            launch_steps = steps
            if grin_lockstep and not seg_complex and steps[lo].before.kind == nat.MEDIUM_ISO_GRIN:
                # reference-exact GRIN segment: integrate in lock-step, then let the trace
                # kernel start from the frozen state in front of the surface
                ld_s = max(cur_x.stride(0), 1)
                (gx, gk, ga, hist) = _grin_lockstep(
                    lib, steps[lo], cur_x, cur_k, cur_e if steps[lo].dir_mode == nat.DIR_POYNTING else None,
                    cur_alive, n, ld_s, device, stream_ptr, grin_history)
                (cur_x, cur_k, cur_alive, ld_x, ld_k, cur_e) = (gx, gk, ga, ld_s, ld_s, None)
                if hist is not None:
                    rec.grin_hist[lo] = hist
                patched = nat.PyrStep.from_buffer_copy(steps[lo])
                patched.before = _probe_medium()          # homogeneous: nothing left to integrate
                patched.dir_mode = nat.DIR_K
                patched.k_norm_hint = 0.0
                patched._grid = getattr(steps[lo], "_grid", None)
                launch_steps = list(steps)
                launch_steps[lo] = patched
            if user_grin and _is_user_grin(steps[lo].before):
                if seg_complex or steps[lo].shape_kind != nat.SHAPE_CONIC:
                    raise lowering.LoweringError("a user-defined GRIN medium must end at a conic surface "
                                                 "(real-valued sequences)")
                ug = _user_grin(lowered[lo].before_obj, steps[lo].before, device)
                ld_s = max(cur_x.stride(0), 1)
                (gx, gk, ga) = ug.propagate(steps[lo].before,
                                            _compose_to_shape(steps[lo].before.frame, steps[lo].shape_frame),
                                            steps[lo].curv, steps[lo].cc, cur_x, cur_k, cur_alive, n, ld_s)
                (cur_x, cur_k, cur_alive, ld_x, ld_k, cur_e) = (gx, gk, ga, ld_s, ld_s, None)
                patched = nat.PyrStep.from_buffer_copy(launch_steps[lo])
                patched.before = _probe_medium()
                patched.dir_mode = nat.DIR_K
                patched.k_norm_hint = 0.0
                patched._grid = None
                launch_steps = list(launch_steps)
                launch_steps[lo] = patched
            want_e = record_e or seg_complex or \
                (first_aniso is not None and hi == first_aniso)
            rows = hi - lo
            any_split = any(bool(steps[i].split) for i in range(lo, hi))
            # per-step record buffers.  Real stretches (never doubling): one allocation per
            # array, `rows` records of the bundle's width.  Complex stretches: the width
            # doubles behind every birefringent interface (hstack order: mode a of column c
            # stays in c, mode b goes to w + c, material_anisotropic.py:89-100), so every
            # step gets buffers of its own width.
            bufs = []            # per step: (x (3, ld), flags (ld), k, e, w_in, w_out, ld, ld2)
            if not any_split:
                (xbuf, ld) = _alloc_rows(rows, n, device, pool=pool)
                fbuf = _empty((rows, ld), torch.uint8, device, pool)
                (kbuf, _) = _alloc_rows(rows, n, device, seg_complex, pool=pool)
                ebuf = None
                if want_e:
                    (ebuf, _) = _alloc_rows(rows, n, device, seg_complex, pool=pool)
                for r in range(rows):
                    bufs.append((xbuf[r], fbuf[r], kbuf[r], ebuf[r] if ebuf is not None else None,
                                 n, n, ld, ld))
            else:
                w = n
                for i in range(lo, hi):
                    w_out = 2 * w if steps[i].split else w
                    (ld, ld2) = (_round_up(max(w, 1), LD_ALIGN), _round_up(max(w_out, 1), LD_ALIGN))
                    bufs.append((_empty((3, ld), torch.float64, device, pool),
                                 _empty((ld,), torch.uint8, device, pool),
                                 _empty((3, ld2, 2), torch.float64, device, pool),
                                 _empty((3, ld2, 2), torch.float64, device, pool),
                                 w, w_out, ld, ld2))
                    w = w_out
            for i in range(lo, hi):
                st = launch_steps[i]
                (bx, bf, bk, be, w_in, w_out, ld, ld2) = bufs[i - lo]
                (st.grin_hist_x, st.grin_hist_k, st.grin_hist_valid, st.grin_hist_count) = \
                    (None, None, None, None)
                st.grin_hist_rows = 0
                if grin_history and st.before.kind == nat.MEDIUM_ISO_GRIN and not seg_complex:
                    cnt = torch.zeros((ld,), dtype=torch.int32, device=device)
                    st.grin_hist_count = cnt.data_ptr()
                    h = {"count": cnt[:n]}
                    if _hist_rows is not None and _hist_rows.get(i, 0) > 0:
                        m = int(_hist_rows[i])
                        h["x"] = torch.zeros((m, 3, ld), dtype=torch.float64, device=device)
                        h["k"] = torch.zeros((m, 3, ld), dtype=torch.float64, device=device)
                        h["valid"] = torch.zeros((m, ld), dtype=torch.uint8, device=device)
                        (st.grin_hist_x, st.grin_hist_k, st.grin_hist_valid) = \
                            (h["x"].data_ptr(), h["k"].data_ptr(), h["valid"].data_ptr())
                        st.grin_hist_rows = m
                        h["rows"] = m
                        h["n"] = n
                    rec.grin_hist[i] = h
                st.out_x = bx.data_ptr()
                st.out_flags = bf.data_ptr()
                st.out_k = bk.data_ptr()
                st.out_e = be.data_ptr() if be is not None else None
                st.ld_out = ld
                st.ld_out2 = ld2        # k / e of a doubling step are 2 w wide
            flags = 0
            if seg_complex:
                flags |= nat.F_COMPLEX | nat.F_RECORD_E
            elif want_e:
                flags |= nat.F_RECORD_E
            if lo == 0 and cur_e is None and not seg_complex:
                cur_e_arg = None            # engine substitutes (0, 1, 0)
            else:
                cur_e_arg = cur_e
            if user_grin and _is_user_grin(steps[lo].after):
                # entrance into a user-defined GRIN medium (this stretch is the one step):
                # (A) propagate + intersect + aperture, (B) n at the hit points from the user's
                # kernel, (C) refraction with the per-ray index -- all into the same record
                assert hi == lo + 1 and not seg_complex
                (bx, bf, bk, be, w_in, w_out, ld, ld2) = bufs[0]
                phase_a = nat.PyrStep.from_buffer_copy(launch_steps[lo])
                phase_a.mode = nat.STEP_PROPAGATE_ONLY
                phase_a._grid = getattr(launch_steps[lo], "_grid", None)
                _launch(lib, [phase_a], 0, 1, cur_x, cur_k, cur_e_arg, cur_alive, n, n_x,
                        ld_k, flags, stream_ptr, events, device=device)
                ug = _user_grin(lowered[lo].after_obj, steps[lo].after, device)
                n_hit = ug.index_at(steps[lo].after.frame, bx, n, ld)
                phase_c = nat.PyrStep.from_buffer_copy(launch_steps[lo])
                phase_c.mode = nat.STEP_DEFLECT_ONLY
                phase_c.dir_mode = nat.DIR_K
                phase_c.k_norm_hint = 0.0
                phase_c.after_n_rays = n_hit.data_ptr()
                phase_c._grid = getattr(launch_steps[lo], "_grid", None)
                kin = bk.clone()                       # phase A left k unchanged in the record
                _launch(lib, [phase_c], 0, 1, bx, kin, None, bf, n, n, ld, flags & ~nat.F_RECORD_E,
                        stream_ptr, events, device=device)
            else:
                spot_arg = None
                if spot is not None and hi == nsteps and not seg_complex and not launch_steps[hi - 1].split:
                    sh = None if spot[1] is None else (C.c_double * 3)(*[float(v) for v in spot[1]])
                    spot_arg = (spot[0], sh)
                    spot_done = True
                _launch(lib, launch_steps, lo, hi, cur_x, cur_k, cur_e_arg, cur_alive, n, n_x,
                        ld_k, flags, stream_ptr, events, wave_end=wave_end,
                        gen=gen_desc if ci == 0 else None, device=device, spot=spot_arg)
            for i in range(lo, hi):
                (bx, bf, bk, be, w_in, w_out, ld, ld2) = bufs[i - lo]
                rec.hit.append(bx[:, :w_in])
                rec.flags.append(bf[:w_in])
                rec.n_in.append(w_in)
                rec.n_out.append(w_out)
                rec.split.append(w_out != w_in)
                if seg_complex:
                    rec.k.append(torch.view_as_complex(bk)[:, :w_out])
                    rec.e.append(torch.view_as_complex(be)[:, :w_out])
                else:
                    rec.k.append(bk[:, :w_out])
                    rec.e.append(be[:, :w_out] if be is not None else None)
            # state handed to the next stretch
            if hi < nsteps:
                (bx, bf, bk, be, w_in, w_out, ld, ld2) = bufs[-1]
                (cur_x, ld_x, cur_alive) = (bx, ld, bf)
                (cur_k, cur_e, ld_k) = (bk, be, ld2)
                (n_x, n) = (w_in, w_out)
                if ld_k != ld_x:
                    # x / alive are read with column i % n_x from rows of leading
                    # dimension ld_k: re-pad to the common ld
                    nx = torch.zeros((3, ld_k), dtype=torch.float64, device=device)
                    nx[:, :n_x] = cur_x[:, :n_x]
                    (cur_x, ld_x) = (nx, ld_k)
    if spot is not None and not spot_done:
        # the last launch could not carry the sums (split step, GRIN lock-step / user GRIN phases, ...)
        spot[0].zero_()
        if rec.hit:
            spot_sums(rec.hit[-1], rec.flags[-1], out=spot[0], shift=spot[1])
    if grin_history and _hist_rows is None and any("x" not in h for h in rec.grin_hist.values()):
        # first pass gave the step counts; second pass records the rows
        rows = {i: int(h["count"].max().item()) if h["count"].numel() else 0
                for (i, h) in rec.grin_hist.items()}
        return trace(lowered, rec.x0, rec.k0, rec.e0, wave, record_e=record_e, device=device,
                     stream=stream, pool=pool, events=events, grin_history=True,
                     _hist_rows=rows, grin_lockstep=grin_lockstep)
    return rec


# ---------------------------------------------------------------------------
# TraceRecord -> RayPath views
# ---------------------------------------------------------------------------
def _hit_mask(flags):
    return (flags & nat.RAY_HIT) != 0


def _alive_mask(flags):
    return (flags & nat.RAY_ALIVE) != 0


def _lazy_efield(k):
    """A unit E perpendicular to k (any such vector satisfies the reference's
    contract: its own choice is an arbitrary SVD null vector)."""
    kr = k.real if k.is_complex() else k
    a = torch.zeros_like(kr)
    idx = kr.abs().argmin(dim=-2, keepdim=True)
    a.scatter_(-2, idx, 1.0)
    t = a - (a * kr).sum(-2, keepdim=True) / (kr * kr).sum(-2, keepdim=True) * kr
    t = t / torch.sqrt((t * t).sum(-2, keepdim=True))
    return t.to(k.dtype)


class _Lazy(object):
    """A value computed on first use (device work is deferred until a bundle field
    is actually read: building the RayPath of a trace launches nothing)."""
    __slots__ = ("fn", "val", "done")

    def __init__(self, fn):
        (self.fn, self.val, self.done) = (fn, None, False)

    def __call__(self):
        if not self.done:
            (self.val, self.done) = (self.fn(), True)
            self.fn = None
        return self.val


def _const(v):
    lz = _Lazy(None)
    (lz.val, lz.done) = (v, True)
    return lz


class _BundleBuilder(object):
    """Materialises one RayBundle of a traced path on first field access.

    start_x / mask / ids: lazies of the state right after the deflection that opens
    the bundle (full width), of the rays the reference keeps
    (material_isotropic.py:194-199) and of their ids; hit / hit_flags: record of the
    next intersect (None for the last bundle, which has a single row).
    """

    def __init__(self, start_x, start_k, start_e, mask, ids, hit, hit_flags, hist=None):
        (self.start_x, self.start_k, self.start_e) = (start_x, start_k, start_e)
        (self.mask, self.ids, self.hit, self.hit_flags) = (mask, ids, hit, hit_flags)
        self.hist = hist
        self.cache = None

    def build(self):
        if self.cache is not None:
            return self.cache
        mask = self.mask()
        if mask is None or bool(mask.all()):
            def take(t):
                return t
        else:
            def take(t):
                return t[..., mask]
        x0 = take(self.start_x())
        start_k = self.start_k() if isinstance(self.start_k, _Lazy) else self.start_k
        start_e = self.start_e() if isinstance(self.start_e, _Lazy) else self.start_e
        k0 = take(start_k)
        e_full = start_e if start_e is not None else _lazy_efield(start_k)
        e0 = take(e_full)
        ids = take(self.ids())
        ones = torch.ones(x0.shape[-1], dtype=torch.bool, device=x0.device)
        if self.hit is None:
            out = {"x": x0.unsqueeze(0), "k": k0.unsqueeze(0),
                   "Efield": e0.unsqueeze(0), "valid": ones.unsqueeze(0),
                   "rayID": ids}
        elif self.hist is not None and "x" in self.hist:
            # GRIN segment with integrator history: rows = start, one per integrator
            # step (lock-step: a finished ray repeats its last row), the intersection
            h = self.hist
            (m, n) = (h["rows"], h["n"])
            last = (h["count"].to(torch.int64) - 1).clamp(min=0)
            row = torch.minimum(torch.arange(m, device=x0.device)[:, None], last[None, :])
            hx = take(torch.gather(h["x"][:, :, :n], 0, row[:, None, :].expand(m, 3, n)))
            hk = take(torch.gather(h["k"][:, :, :n], 0, row[:, None, :].expand(m, 3, n)))
            hv = take(torch.gather(h["valid"][:, :n], 0, row) != 0)
            hv = torch.cummin(hv.to(torch.uint8), dim=0).values != 0
            x = torch.cat((x0.unsqueeze(0), hx, take(self.hit).unsqueeze(0)))
            k = torch.cat((k0.unsqueeze(0), hk, hk[-1:]))
            e = _lazy_efield(k) if start_e is None else \
                torch.cat((e0.unsqueeze(0), _lazy_efield(k[1:])))
            valid = torch.cat((ones.unsqueeze(0), hv, take(_hit_mask(self.hit_flags)).unsqueeze(0)))
            out = {"x": x, "k": k, "Efield": e, "valid": valid, "rayID": ids}
        else:
            x = torch.stack((x0, take(self.hit)))
            out = {"x": x, "k": k0.unsqueeze(0).expand(2, -1, -1),
                   "Efield": e0.unsqueeze(0).expand(2, -1, -1),
                   "valid": torch.stack((ones, take(_hit_mask(self.hit_flags)))),
                   "rayID": ids}
        self.cache = out
        return out

    def field(self, name):
        return lambda: self.build()[name]


def paths_from_record(rec, splitup=False):
    """list[RayPath] with the reference's structure: path = [b0, b0, b1..bS]
    (optical_system.py:74-91); with splitup=True every birefringent interface
    forks the path (2^m paths of n0 rays), otherwise it doubles the rays
    (one path, hstack order, material_anisotropic.py:87-113)."""
    nsteps = len(rec.hit)
    generated = rec.gen is not None and rec._x0 is None
    n0 = rec.n0 if generated else rec.x0.shape[1]
    dev = rec.device if generated else rec.x0.device
    nsplit = sum(1 for s in rec.split if s)
    npaths = (2 ** nsplit) if splitup else 1

    def default_e0():
        e = torch.zeros_like(rec.x0)
        e[1] = 1.0
        return e
    e0 = None if generated else rec.e0
    paths = []
    for p in range(npaths):
        def cols(t, width):
            # after q splits the device bundle has 2^q blocks of n0 columns; path p
            # lives in block p mod 2^q.  Without path forking the records already
            # have exactly the bundle's width: no view op at all.
            if not splitup:
                return t
            b = p % (width // n0)
            return t[..., b * n0:(b + 1) * n0]

        if generated:
            # lazies: the first bundle's fields call pyr_generate_bundle when first read
            sx = _Lazy(lambda: rec.x0)
            (sk, se) = (_Lazy(lambda: rec.k0), _Lazy(lambda: rec.e0))
        else:
            sx = _const(rec.x0)
            (sk, se) = (rec.k0, e0 if e0 is not None else default_e0())
        mask = _const(None)
        ids = _Lazy(lambda: torch.arange(n0, device=dev))
        bundles = []
        for s in range(nsteps):
            hit = cols(rec.hit[s], rec.n_in[s])
            fl = cols(rec.flags[s], rec.n_in[s])
            bb = _BundleBuilder(sx, sk, se, mask, ids, hit, fl,
                                hist=rec.grin_hist.get(s) if not splitup else None)
            splitted = bool(s >= 1 and rec.split[s - 1] and not splitup)
            bundles.append(RayBundle(_lazy={f: bb.field(f) for f in RayBundle._FIELDS},
                                     wave=rec.wave, splitted=splitted))
            sk = cols(rec.k[s], rec.n_out[s])
            se = cols(rec.e[s], rec.n_out[s]) if rec.e[s] is not None else None
            # survivors: the device ALIVE bit is cumulative; an anisotropic
            # deflection keeps every ray it was handed (material_anisotropic.py
            # :87-100 has no validity filter), which the kernel mirrors
            if rec.split[s] and not splitup:
                sx = _Lazy(lambda hit=hit: torch.cat((hit, hit), dim=1))
                mask = _Lazy(lambda fl=fl: torch.cat((_alive_mask(fl), _alive_mask(fl))))
                ids = _Lazy(lambda ids=ids: torch.cat((ids(), ids())))
            else:
                sx = _const(hit)
                mask = _Lazy(lambda fl=fl: _alive_mask(fl))
        bb = _BundleBuilder(sx, sk, se, mask, ids, None, None)
        splitted = bool(nsteps >= 1 and rec.split[nsteps - 1] and not splitup)
        bundles.append(RayBundle(_lazy={f: bb.field(f) for f in RayBundle._FIELDS},
                                 wave=rec.wave, splitted=splitted))
        # every element's seqtrace starts its path with the bundle it was handed
        # (optical_element.py:331) and the system appends that path to its own
        # (optical_system.py:83-91): the hand-over bundle appears twice
        path = RayPath(bundles[0])
        for s in range(nsteps):
            if s == 0 or rec.lowered[s].elem_index != rec.lowered[s - 1].elem_index:
                path.appendRayBundle(path.raybundles[-1])
            path.appendRayBundle(bundles[s + 1])
        paths.append(path)
    return paths


# ---------------------------------------------------------------------------
# entry used by OpticalSystem.seqtrace
# ---------------------------------------------------------------------------
def seqtrace(system, initialbundle, elementsequence, splitup=False,
             record_e=False, grin_history=False, grin_lockstep=False):
    lowered = lowering.lower(system, elementsequence, initialbundle.wave,
                             splitup=splitup)
    gen = getattr(initialbundle, "generator", None)
    if gen is not None:
        # a generated bundle (OpticalSystemAnalysis.aim): the kernel expands it in registers
        rec = trace(lowered, None, None, None, initialbundle.wave, record_e=record_e,
                    grin_history=grin_history, gen=gen, grin_lockstep=grin_lockstep)
        paths = paths_from_record(rec, splitup=splitup)
        for p in paths:
            p.record = rec
        return paths
    if initialbundle.x.shape[0] != 1:
        x0 = initialbundle.x[-1]
        k0 = initialbundle.k[-1]
        e0 = initialbundle.Efield[-1]
    else:
        (x0, k0, e0) = (initialbundle.x[0], initialbundle.k[0],
                        initialbundle.Efield[0])
    rec = trace(lowered, x0, k0, e0, initialbundle.wave, record_e=record_e,
                grin_history=grin_history, grin_lockstep=grin_lockstep)
    paths = paths_from_record(rec, splitup=splitup)
    for p in paths:
        p.record = rec
    return paths


def _column_view(rec, lo, hi, lowered, wave):
    """TraceRecord of the rays [lo, hi) of a batch record (zero-copy column slices)."""
    sub = TraceRecord()
    sub.lowered = lowered
    sub.wave = wave
    (sub.x0, sub.k0) = (rec.x0[:, lo:hi], rec.k0[:, lo:hi])
    sub.e0 = rec.e0[:, lo:hi] if rec.e0 is not None else None
    for s in range(len(rec.hit)):
        sub.hit.append(rec.hit[s][:, lo:hi])
        sub.k.append(rec.k[s][:, lo:hi])
        sub.e.append(rec.e[s][:, lo:hi] if rec.e[s] is not None else None)
        sub.flags.append(rec.flags[s][lo:hi])
        sub.n_in.append(hi - lo)
        sub.n_out.append(hi - lo)
        sub.split.append(False)
    return sub


def seqtrace_batch(system, bundles, elementsequence, record_e=False):
    """Trace several bundles of DIFFERENT wavelength through the same sequence in one
    native launch per group of PYR_MAX_WAVES wavelengths (the F / d / C bundles of
    demos/demo_doublegauss.py:189-213, which the reference traces one after the other).
    Media indices are evaluated per wavelength at lowering time; the kernel picks them
    per ray from the segment the ray lies in.  Returns one list[RayPath] per bundle,
    exactly what `seqtrace` returns for it.  Sequences with crystals or GRIN media, and
    record_e, fall back to one launch per bundle."""
    bundles = list(bundles)
    out = [None] * len(bundles)
    if record_e or len(bundles) == 1:
        return [seqtrace(system, b, elementsequence, record_e=record_e) for b in bundles]
    for g0 in range(0, len(bundles), nat.MAX_WAVES):
        group = bundles[g0:g0 + nat.MAX_WAVES]
        waves = [b.wave for b in group]
        try:
            (per_wave, batch) = lowering.lower_batch(system, elementsequence, waves)
        except lowering.LoweringError:
            for (i, b) in enumerate(group):
                out[g0 + i] = seqtrace(system, b, elementsequence)
            continue
        require_cuda()
        dev = torch.device("cuda", torch.cuda.current_device())
        rows = []
        for b in group:
            rows.append([as_tensor(t[-1], dev) for t in (b.x, b.k, b.Efield)])
            if any(t.is_complex() for t in rows[-1]):
                raise ValueError("wavelength batches are real-valued")
        (x0, k0, e0) = (torch.cat([r[c] for r in rows], dim=1) for c in range(3))
        ends = np.cumsum([r[0].shape[1] for r in rows]).tolist()
        try:
            rec = trace(batch, x0, k0, e0, waves[0], wave_end=ends)
        except nat.NativeError as err:
            if getattr(err, "code", None) != nat.E_UNSUPPORTED:
                raise
            # a feature the batch instantiations do not carry: one launch per bundle
            for (i, b) in enumerate(group):
                out[g0 + i] = seqtrace(system, b, elementsequence)
            continue
        lo = 0
        for (i, (b, hi)) in enumerate(zip(group, ends)):
            sub = _column_view(rec, lo, hi, per_wave[i], b.wave)
            paths = paths_from_record(sub)
            for p in paths:
                p.record = sub
            out[g0 + i] = paths
            lo = hi
    return out


# ---------------------------------------------------------------------------
# stand-alone plugin calls (Material.propagate / refract, Surface.intersect)
# ---------------------------------------------------------------------------
def _single_step(st, bundle, mode, record_e=True, ignore_validity=False):
    """One stand-alone plugin call: launch `st` in `mode` on the bundle's last row.
    Returns (x, k, e, flags, n): k / e are 2n wide when the step splits (mode a in
    column i, mode b in n + i) and complex whenever the fields or a medium are."""
    lib = require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    x = as_tensor(bundle.x[-1], dev).contiguous()
    k = as_tensor(bundle.k[-1], dev).contiguous()
    e = as_tensor(bundle.Efield[-1], dev).contiguous()
    complex_ = bool(k.is_complex() or e.is_complex() or
                    st.after.kind == nat.MEDIUM_ANISO or st.before.kind == nat.MEDIUM_ANISO)
    if x.is_complex():
        x = x.real.contiguous()
    n = x.shape[1]
    st.mode = mode
    split = bool(complex_ and mode != nat.STEP_PROPAGATE_ONLY and
                 st.after.kind == nat.MEDIUM_ANISO)
    st.split = 1 if split else 0
    (xb, ld) = _padded(x)
    (kb, _) = _padded(k, complex_)
    (eb, _) = _padded(e, complex_)
    ld2 = _round_up(2 * max(n, 1), LD_ALIGN) if split else ld
    tail = (2,) if complex_ else ()
    ox = torch.empty((3, ld), dtype=torch.float64, device=dev)
    ok = torch.empty((3, ld2) + tail, dtype=torch.float64, device=dev)
    oe = torch.empty((3, ld2) + tail, dtype=torch.float64, device=dev)
    of = torch.empty((ld,), dtype=torch.uint8, device=dev)
    (st.out_x, st.out_k, st.out_e, st.out_flags) = (ox.data_ptr(), ok.data_ptr(),
                                                    oe.data_ptr(), of.data_ptr())
    st.ld_out = ld
    st.ld_out2 = ld2
    alive_p = torch.zeros((ld,), dtype=torch.uint8, device=dev)
    if ignore_validity:
        alive_p[:n] = nat.RAY_ALIVE
    else:
        alive_p[:n] = bundle.valid[-1].to(dev).to(torch.uint8) * nat.RAY_ALIVE
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    flags = nat.F_RECORD_E | (nat.F_COMPLEX if complex_ else 0)
    with torch.cuda.device(dev):
        _launch(lib, [st], 0, 1, xb, kb, eb, alive_p, n, n, ld, flags, stream)
    width = 2 * n if split else n
    if complex_:
        (ok, oe) = (torch.view_as_complex(ok), torch.view_as_complex(oe))
    return (ox[:, :n], ok[:, :width], oe[:, :width], of[:n], n)


def _probe_medium():
    m = nat.PyrMedium()
    m.kind = nat.MEDIUM_ISO_CONST
    m.n = 1.0
    for i in range(3):
        m.frame.r[i * 4] = 1.0
    return m


def surface_intersect(surface, bundle, remove_rays_outside_aperture=True,
                      medium=None):
    st = nat.PyrStep()
    lowering.lower_surface(surface, st)
    if not remove_rays_outside_aperture:
        st.aperture_kind = nat.AP_BASE
    st.before = _probe_medium() if medium is None else \
        lowering.lower_medium(medium, bundle.wave)
    st.after = _probe_medium()
    st.dir_mode = nat.DIR_POYNTING
    (ox, ok, oe, of, _) = _single_step(st, bundle, nat.STEP_PROPAGATE_ONLY)
    valid = _hit_mask(of) | ~bundle.valid[-1].to(of.device)
    # RayBundle.append ANDs with the previous row itself (ray.py:100)
    bundle.append(ox, ok if medium is not None else bundle.k[-1].to(ox.device),
                  bundle.Efield[-1].to(ox.device), valid)


def shape_intersect(shape, bundle):
    class _S(object):
        pass
    from .raytracer.aperture import BaseAperture
    s = _S()
    s.shape = shape
    s.aperture = BaseAperture.p(shape.lc)
    surface_intersect(s, bundle, remove_rays_outside_aperture=False)


def material_propagate(material, bundle, next_surface):
    surface_intersect(next_surface, bundle, True, medium=material)


def material_deflect(material, bundle, surface, mirror=False, splitup=False):
    """Material.refract / reflect as a stand-alone call.  Isotropic media return one
    bundle of the surviving rays (material_isotropic.py:194-199); anisotropic media
    keep every ray they are handed and return both forward modes -- one bundle of 2N
    rays in hstack order, or two bundles with `splitup` (material_anisotropic.py
    :87-113, :133-155)."""
    st = nat.PyrStep()
    lowering.lower_surface(surface, st)
    st.aperture_kind = nat.AP_BASE
    st.before = _probe_medium()
    st.after = lowering.lower_medium(material, bundle.wave)
    st.interaction = nat.REFLECT if mirror else nat.REFRACT
    aniso = st.after.kind == nat.MEDIUM_ANISO
    st.dir_mode = nat.DIR_POYNTING if aniso else nat.DIR_K
    (ox, ok, oe, of, n) = _single_step(st, bundle, nat.STEP_DEFLECT_ONLY,
                                       ignore_validity=aniso)
    ids = bundle.rayID.to(ox.device)
    if not aniso:
        alive = _alive_mask(of)
        return (RayBundle(ox[:, alive], ok[:, alive], oe[:, alive], ids[alive],
                          wave=bundle.wave),)
    if splitup:
        return (RayBundle(ox, ok[:, :n], oe[:, :n], ids, wave=bundle.wave),
                RayBundle(ox, ok[:, n:], oe[:, n:], ids, wave=bundle.wave))
    return (RayBundle(torch.cat((ox, ox), dim=1), ok, oe, torch.cat((ids, ids)),
                      wave=bundle.wave, splitted=True),)


# ---------------------------------------------------------------------------
# end-to-end host entry (pyr_trace_host): host bundle in, final record out
# ---------------------------------------------------------------------------
class HostTracer(object):
    """Reusable binding of pyr_trace_host_io for one lowered (real, non-splitting)
    sequence: owns the device workspace and pinned result buffers.

    all_records=True also returns the hit points / wave vectors / flags of EVERY
    sequence entry -- what the S + 2 bundles of OpticalSystem.seqtrace carry
    (optical_system.py:73-94) -- in pinned (n_steps, 3, n) / (n_steps, n) arrays."""

    def __init__(self, lowered, n_rays, chunk_rays=1 << 20, device=None, all_records=False):
        self.lib = require_cuda()
        self.device = torch.device("cuda", torch.cuda.current_device()) \
            if device is None else torch.device(device)
        self.lowered = lowered
        for ls in lowered:
            bind_grid(ls.st, self.device)
        self.steps = lowering.step_array(lowered)
        self.n = int(n_rays)
        self.chunk = int(min(chunk_rays, max(self.n, 1)))
        self.all_records = bool(all_records)
        ns = len(lowered)
        self.crystal = any(nat.MEDIUM_ANISO in (ls.st.before.kind, ls.st.after.kind) for ls in lowered)
        (self.mult_x, self.mult_k) = (1, 1)
        if self.crystal:
            # birefringent media: the last record is that of the doubled bundle, k / E complex
            if self.all_records:
                raise ValueError("all_records: real-valued sequences only")
            (mx, mk) = (C.c_int64(1), C.c_int64(1))
            nbytes = self.lib.pyr_trace_host_crystal_workspace(self.steps, ns, self.chunk,
                                                               C.byref(mx), C.byref(mk))
            if nbytes <= 0:
                raise nat.NativeError("pyr_trace_host_crystal_workspace: sequence not supported")
            (self.mult_x, self.mult_k) = (int(mx.value), int(mk.value))
        else:
            nbytes = self.lib.pyr_trace_host_io_workspace(ns, self.chunk, int(self.all_records))
        self.workspace = torch.empty((nbytes + 256,), dtype=torch.uint8,
                                     device=self.device)
        off = (-self.workspace.data_ptr()) % 256
        self.ws_ptr = self.workspace.data_ptr() + off
        self.ws_bytes = nbytes
        kdt = torch.complex128 if self.crystal else torch.float64
        self.x_last = torch.empty((3, self.n * self.mult_x), dtype=torch.float64).pin_memory()
        self.k_last = torch.empty((3, self.n * self.mult_k), dtype=kdt).pin_memory()
        self.e_last = torch.empty((3, self.n * self.mult_k), dtype=kdt).pin_memory() if self.crystal else None
        self.flags_last = torch.empty((self.n * self.mult_x,), dtype=torch.uint8).pin_memory()
        self.spot8 = torch.zeros((8,), dtype=torch.float64).pin_memory()
        self.x_all = self.k_all = self.flags_all = None
        if self.all_records:
            self.x_all = torch.empty((ns, 3, self.n), dtype=torch.float64).pin_memory()
            self.k_all = torch.empty((ns, 3, self.n), dtype=torch.float64).pin_memory()
            self.flags_all = torch.empty((ns, self.n), dtype=torch.uint8).pin_memory()
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def __call__(self, x0=None, k0=None, e0=None, gen=None):
        """x0, k0, e0: contiguous (3, n) float64 CPU tensors (pinned for overlap);
        e0=None is the reference's default field (0, 1, 0) (ray.py:71-73) and is not
        uploaded.  gen: a bundlegen.BundleGen instead of the arrays -- nothing but the
        descriptor crosses the bus on the way in.  Returns (x_last, k_last, flags_last,
        spot8) host tensors (and fills x_all / k_all / flags_all with all_records)."""
        io = nat.PyrHostIO()
        ns = len(self.lowered)
        if gen is not None:
            assert gen.n == self.n
            desc = gen.descriptor(self.device)
            io.gen = C.pointer(desc)
            self.h2d_bytes = C.sizeof(nat.PyrBundleGen) + ns * C.sizeof(nat.PyrStep) + \
                (0 if gen.raster.rows is None else 0)        # row table: uploaded once, cached
        else:
            for t in (x0, k0, e0):
                assert t is None or (t.device.type == "cpu" and t.dtype == torch.float64 and
                                     t.is_contiguous() and t.shape == (3, self.n))
            io.x0 = x0.data_ptr()
            io.k0 = k0.data_ptr()
            io.e0 = None if e0 is None else e0.data_ptr()
            self.h2d_bytes = (72 if e0 is not None else 48) * self.n + ns * C.sizeof(nat.PyrStep)
        io.x_last = self.x_last.data_ptr()
        io.k_last = self.k_last.data_ptr()
        io.flags_last = self.flags_last.data_ptr()
        io.spot8 = self.spot8.data_ptr()
        self.d2h_bytes = 49 * self.n + 64
        if self.crystal:
            io.e_last = self.e_last.data_ptr()
            self.d2h_bytes = (25 * self.mult_x + 96 * self.mult_k) * self.n + 64
        if self.all_records:
            io.x_all = self.x_all.data_ptr()
            io.k_all = self.k_all.data_ptr()
            io.flags_all = self.flags_all.data_ptr()
            self.d2h_bytes += 49 * self.n * ns
        with torch.cuda.device(self.device):
            nat.check(self.lib.pyr_trace_host_io(self.steps, ns, C.byref(io), self.n,
                                                 self.ws_ptr, self.ws_bytes, self.chunk))
        return (self.x_last, self.k_last, self.flags_last, self.spot8)


def spot_sums(x, flags=None, mask=nat.RAY_ALIVE, out=None, shift=None):
    """Device partial sums of RayBundleAnalysis (ray_analysis.py:44-86) about
    the reference point `shift` (3 floats, default origin): out[0:3] =
    sum (x - shift), out[3] = count, out[4:7] = sum (x - shift)^2 (float64 CUDA
    tensor, accumulated).  Sums of several ranks (same shift) can be all-reduced
    before `spot_from_sums`."""
    lib = require_cuda()
    dev = x.device
    if out is None:
        out = torch.zeros((8,), dtype=torch.float64, device=dev)
    assert x.dim() == 2 and x.shape[0] == 3 and x.stride(1) == 1
    ld = x.stride(0)
    n = x.shape[1]
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    sh = None if shift is None else (C.c_double * 3)(*[float(v) for v in shift])
    with torch.cuda.device(dev):
        nat.check(lib.pyr_spot_sums(x.data_ptr(), ld,
                                    flags.data_ptr() if flags is not None else None,
                                    mask, n, sh, out.data_ptr(), stream))
    return out


def spot_points(x, flags=None, mask=nat.RAY_ALIVE, frame=None, width=None, out=None):
    """(x, y) of the rays with `flags & mask` (flags None = all), compacted on the device:
    returns (xy (2, width) float64, count () int64), both CUDA tensors -- the count is NOT
    read back.  frame: nat.PyrFrame of the surface whose local coordinates are wanted
    (OpticalSystemAnalysis.get_spot, reference :283-303), None = global.  Point order is
    unspecified.  out: (xy, count) from a previous call to reuse."""
    lib = require_cuda()
    dev = x.device
    assert x.dim() == 2 and x.shape[0] == 3 and x.stride(1) == 1
    n = x.shape[1]
    width = n if width is None else int(width)
    if out is None:
        out = (torch.empty((2, max(width, 1)), dtype=torch.float64, device=dev),
               torch.zeros((), dtype=torch.int64, device=dev))
    (xy, count) = out
    assert xy.shape[1] >= min(width, max(n, 1)) and xy.is_contiguous()
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    with torch.cuda.device(dev):
        nat.check(lib.pyr_spot_points(x.data_ptr(), x.stride(0),
                                      flags.data_ptr() if flags is not None else None,
                                      mask, n, C.byref(frame) if frame is not None else None,
                                      xy.data_ptr(), xy.shape[1], count.data_ptr(), stream))
    return xy, count


def spot_from_sums(s, shift=None):
    """(centroid[3], rms about the centroid) with the reference's
    normalisations: centroid = sum/(N + 1e-17), rms = sqrt(sum|x-c|^2 /
    (N - 1 + 1e-17)); `shift` = the reference point the sums were taken about."""
    s = [float(v) for v in s]
    n = s[3]
    c = [s[i] / (n + 1e-17) for i in range(3)]
    ss = sum(s[4 + i] - 2.0 * c[i] * s[i] + n * c[i] * c[i] for i in range(3))
    if shift is not None:
        c = [c[i] + float(shift[i]) for i in range(3)]
    return c, (max(ss, 0.0) / (n - 1 + 1e-17)) ** 0.5


def last_surface_origin(lowered):
    """Global vertex of the last sequence entry: the natural reference point
    of the spot sums."""
    return [float(v) for v in lowered[-1].st.shape_frame.o]
