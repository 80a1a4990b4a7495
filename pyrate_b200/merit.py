"""Optimiser-loop binding of the trace: one merit-function evaluation in ~0.1 ms.

The reference's optimisers (optimize/optimize.py:73-91) call a merit function once per
function evaluation; the usual one (demos/demo_doublegauss.py:144-147) is
`seqtrace` + `RayBundleAnalysis.get_rms_spot_size`, 74 ms there for a 1 027-ray
double-Gauss.  Through the general API this engine pays mostly Python: lowering the object
graph, allocating records, building lazy RayPath views, a torch round trip for the spot.
`MeritTrace` binds ONE (system, sequence, bundle) and keeps everything that does not change
between evaluations:

  * the step table is lowered once; `refresh()` re-reads every optimisable quantity
    (curvatures, conic constants, polynomial coefficients, frames, indices) straight into
    the ctypes records -- frames only when LocalCoordinates.update() replaced them;
  * only the LAST entry is recorded (the kernel skips every other store);
  * trace + spot sums + 64-byte read-back are one C call (`pyr_trace_spot`).

Real-valued, non-splitting sequences that fit one launch; anything else raises (use
`OpticalSystem.seqtrace`).
"""
import ctypes as C

import numpy as np
import torch

from . import _native as nat
from . import engine, lowering


def _frame_view(frame):
    """float64[12] NumPy view of a PyrFrame living inside a ctypes array."""
    return np.ctypeslib.as_array((C.c_double * 12).from_address(C.addressof(frame)))


class _FrameSlot(object):
    __slots__ = ("lc", "views", "basis", "origin")

    def __init__(self, lc):
        (self.lc, self.views, self.basis, self.origin) = (lc, [], None, None)

    def refresh(self):
        lc = self.lc
        if lc.localbasis is self.basis and lc.globalcoordinates is self.origin:
            return False                      # LocalCoordinates.update() replaces both arrays
        (self.basis, self.origin) = (lc.localbasis, lc.globalcoordinates)
        flat = np.concatenate((np.asarray(self.basis, dtype=np.float64).reshape(-1),
                               np.asarray(self.origin, dtype=np.float64).reshape(-1)))
        for v in self.views:
            v[:] = flat
        return True


class LiveStepTable(object):
    """The lowered step table of one (system, sequence, wavelength), kept alive: `refresh()`
    re-reads every optimisable quantity into the SAME ctypes records.  Host-only logic."""

    def __init__(self, system, elementsequence, wave):
        self.system = system
        self.sequence = elementsequence
        self.wave = wave
        low = lowering.lower(system, elementsequence, wave)
        self.lowered = low
        self.arr = lowering.step_array(low)                   # THE step table, patched in place
        self.n_steps = len(low)
        self._frames = {}
        self._plan = []
        for (i, ls) in enumerate(low):
            st = self.arr[i]
            surface = system.elements[ls.elemkey].surfaces[ls.surfkey]
            conic = st.shape_kind in (nat.SHAPE_CONIC, nat.SHAPE_CYLINDER)
            self._plan.append((st, surface, conic, ls.before_obj, ls.after_obj))
            for (lc, frame) in ((surface.shape.lc, st.shape_frame), (surface.aperture.lc, st.aperture_frame),
                                (ls.before_obj.lc, st.before.frame), (ls.after_obj.lc, st.after.frame)):
                slot = self._frames.get(id(lc))
                if slot is None:
                    slot = self._frames[id(lc)] = _FrameSlot(lc)
                slot.views.append(_frame_view(frame))
        self.refresh()

    def refresh(self):
        """Re-read every optimisable quantity into the step table (in place)."""
        for slot in self._frames.values():
            slot.refresh()
        index = {}
        wave = self.wave
        for (st, surface, conic, before, after) in self._plan:
            shape = surface.shape
            if conic:
                st.curv = shape.curvature()
                st.cc = shape.conic()
            else:
                (ox, ok, oe, of, ld) = (st.out_x, st.out_k, st.out_e, st.out_flags, st.ld_out)
                (gx, gy, gc) = (st.grid_tx, st.grid_ty, st.grid_c)
                lowering.lower_surface(surface, st)        # coefficients, Newton settings
                (st.out_x, st.out_k, st.out_e, st.out_flags, st.ld_out) = (ox, ok, oe, of, ld)
                (st.grid_tx, st.grid_ty, st.grid_c) = (gx, gy, gc)
            for (mat, med, hint) in ((before, st.before, True), (after, st.after, False)):
                if med.kind != nat.MEDIUM_ISO_CONST:
                    continue
                n = index.get(id(mat))
                if n is None:
                    n = index[id(mat)] = float(mat.get_optical_index(None, wave))
                med.n = n
                if hint and st.k_norm_hint > 0.0:
                    st.k_norm_hint = n


class MeritTrace(LiveStepTable):

    def __init__(self, system, elementsequence, bundle, device=None):
        self.lib = engine.require_cuda()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None \
            else torch.device(device)
        LiveStepTable.__init__(self, system, elementsequence, bundle.wave)
        low = self.lowered
        if len(low) > engine.MAX_STEPS_PER_LAUNCH or \
                sum(engine._needs_aux(ls.st) for ls in low) > engine.MAX_AUX_PER_LAUNCH:
            raise lowering.LoweringError("MeritTrace: the sequence does not fit one launch")
        for ls in low:
            st = ls.st
            if nat.MEDIUM_ANISO in (st.before.kind, st.after.kind) or st.split:
                raise lowering.LoweringError("MeritTrace: real-valued, non-splitting sequences only")
        # ---- the bundle: a generator (expanded in the kernel) or resident arrays ----
        self.gen = getattr(bundle, "generator", None)
        self.rin = nat.PyrRaysIn()
        if self.gen is not None and engine._gen_fusable(low, False, False, None):
            self.n = self.gen.n
            self._desc = self.gen.descriptor(self.device)
            self.rin.gen = C.pointer(self._desc)
        else:
            (x, k, e) = (bundle.x[-1], bundle.k[-1], bundle.Efield[-1])
            (self._x, self._k, self._e) = engine.device_bundle(x, k, e, self.device)
            self.n = self._x.shape[1]
            (self.rin.x, self.rin.k, self.rin.e) = (self._x.data_ptr(), self._k.data_ptr(),
                                                    self._e.data_ptr())
            self.rin.ld = max(self._x.stride(0), 1)
            self.rin.n_x = self.n
        # ---- records: the last entry only ----
        ld = engine._round_up(max(self.n, 1), engine.LD_ALIGN)
        self.x_last = torch.empty((3, ld), dtype=torch.float64, device=self.device)
        self.flags_last = torch.empty((ld,), dtype=torch.uint8, device=self.device)
        for i in range(self.n_steps):
            st = self.arr[i]
            (st.out_x, st.out_k, st.out_e, st.out_flags) = (None, None, None, None)
            st.ld_out = ld
            if getattr(low[i].st, "_grid", None) is not None:
                engine.bind_grid(low[i].st, self.device)
                (st.grid_tx, st.grid_ty, st.grid_c) = (low[i].st.grid_tx, low[i].st.grid_ty,
                                                       low[i].st.grid_c)
        last = self.arr[self.n_steps - 1]
        (last.out_x, last.out_flags) = (self.x_last.data_ptr(), self.flags_last.data_ptr())
        self.spot_dev = torch.zeros((8,), dtype=torch.float64, device=self.device)
        self.spot_host = torch.zeros((8,), dtype=torch.float64).pin_memory()
        self._spot_np = self.spot_host.numpy()
        self._shift = (C.c_double * 3)()
        self._stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def refresh(self):
        LiveStepTable.refresh(self)
        if hasattr(self, "_shift"):
            o = self.arr[self.n_steps - 1].shape_frame.o
            (self._shift[0], self._shift[1], self._shift[2]) = (o[0], o[1], o[2])

    def spot_sums(self, refresh=True):
        """Trace and return the 8 spot sums of the image-plane record (NumPy view of pinned
        host memory: sum (x - o), count, sum (x - o)^2 about the last surface's vertex o)."""
        if refresh:
            self.refresh()
        nat.check(self.lib.pyr_trace_spot(self.arr, self.n_steps, C.byref(self.rin), self.n, 0,
                                          self._shift, self.spot_dev.data_ptr(),
                                          self.spot_host.data_ptr(), self._stream))
        return self._spot_np

    def __call__(self, refresh=True):
        """RMS spot radius about the centroid (RayBundleAnalysis.get_rms_spot_size_centroid,
        analysis/ray_analysis.py:44-86) of the current state of the system."""
        s = self.spot_sums(refresh)
        return engine.spot_from_sums(s, self._shift)[1]

    def centroid_and_rms(self, refresh=True):
        s = self.spot_sums(refresh)
        (c, rms) = engine.spot_from_sums(s, self._shift)
        return c, rms, float(s[3])
