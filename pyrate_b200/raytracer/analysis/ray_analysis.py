"""Bundle statistics right after the trace (host API mirror of reference
raytracer/analysis/ray_analysis.py:33-170): the step optimiser merit functions
consume (demos/demo_doublegauss.py:144-147).

Centroid and RMS spot size -- the only cross-ray reduction of the path and the
only collective of a multi-GPU run -- use the native pyr_spot_sums kernel when
the bundle lives on a CUDA device (sums taken about the first ray, so they do
not cancel catastrophically); everything else is a handful of torch
expressions on the per-surface records, wherever they live.
"""
import math

import torch

from ..globalconstants import numerical_tolerance
from ..ray import as_tensor


class RayBundleAnalysis(object):

    def __init__(self, raybundle, name=""):
        self.raybundle = raybundle
        self.name = name

    # ---- spot statistics (ray_analysis.py:44-86) ----
    def _sums(self, shift):
        x = self.raybundle.x[-1]
        if x.is_cuda and x.shape[1] > 0 and x.stride(1) == 1:
            from ... import engine
            return engine.spot_sums(x, None, shift=shift).cpu()
        d = x - torch.as_tensor(shift, dtype=x.dtype, device=x.device)[:, None]
        s = torch.zeros(8, dtype=torch.float64)
        s[0:3] = d.sum(1).cpu()
        s[3] = x.shape[1]
        s[4:7] = (d * d).sum(1).cpu()
        return s

    def _anchor(self):
        x = self.raybundle.x[-1]
        if x.shape[1] == 0:
            return [0.0, 0.0, 0.0]
        return [float(v) for v in x[:, 0].cpu()]

    def get_centroid_position(self):
        shift = self._anchor()
        s = self._sums(shift)
        n = float(s[3])
        return torch.tensor([float(s[i]) / (n + numerical_tolerance) + shift[i] *
                             (n / (n + numerical_tolerance)) for i in range(3)],
                            dtype=torch.float64)

    def get_rms_spot_size(self, reference_pos):
        ref = [float(v) for v in as_tensor(reference_pos).reshape(-1)]
        s = self._sums(ref)
        n = float(s[3])
        return math.sqrt(float(s[4] + s[5] + s[6]) / (n - 1 + numerical_tolerance))

    def get_rms_spot_size_centroid(self):
        return self.get_rms_spot_size(self.get_centroid_position())

    # ---- directions (ray_analysis.py:88-138) ----
    def get_centroid_direction(self):
        d = self.raybundle.returnKtoD()[-1]
        com = d.sum(1)
        return com / torch.sqrt((com * com).sum())

    def get_rms_angluar_size(self, ref_direction):
        d = self.raybundle.returnKtoD()[-1]
        ref = as_tensor(ref_direction, d.device).reshape(3, 1).to(d.dtype)
        cross = torch.linalg.cross(d, ref.expand_as(d), dim=0)
        return math.asin(math.sqrt(float((cross * cross).sum()) / d.shape[1]))

    def get_rms_angluar_size_centroid(self):
        return self.get_rms_angluar_size(self.get_centroid_direction())

    # ---- path integrals (ray_analysis.py:140-170) ----
    def get_arc_length(self, first=0, last=None):
        last_no = 0 if last is None else last
        x = self.raybundle.x
        stop = x.shape[0] - 1 + last_no if last_no <= 0 else -1 + last_no
        seg = x[first + 1:last] - x[first:stop]
        return torch.sqrt((seg * seg).sum(1)).sum(0)

    def get_phase_difference(self, first=0, last=None):
        last_no = 0 if last is None else last
        x = self.raybundle.x
        k = self.raybundle.k
        k = k.real if k.is_complex() else k
        stop = x.shape[0] - 1 + last_no if last_no <= 0 else -1 + last_no
        dph = x[first + 1:last] * k[first + 1:last] - x[first:stop] * k[first:stop]
        return dph.sum(1).sum(0)


class RayPathAnalysis(object):
    """Path integrals over all bundles of a RayPath (reference :170-213).  As in the
    reference every bundle of the path must have the same width (no ray dropped on
    the way), and the hand-over bundle that appears twice in a system-level path
    (optical_system.py:74-91) is counted twice."""

    def __init__(self, raypath, name=""):
        self.raypath = raypath
        self.name = name

    def _sum(self, method, first, last):
        total = None
        for raybundle in self.raypath.raybundles[first:last]:
            part = getattr(RayBundleAnalysis(raybundle), method)()
            total = part.clone() if total is None else total + part
        if total is None:
            n = self.raypath.raybundles[0].x.shape[-1]
            total = torch.zeros(n, dtype=torch.float64)
        return total

    def get_arc_length(self, first=0, last=None):
        return self._sum("get_arc_length", first, last)

    def get_phase_difference(self, first=0, last=None):
        return self._sum("get_phase_difference", first, last)

    def get_relative_phase_difference(self, first=0, last=None, referenceray=None,
                                      wavelength=None):
        """Phase differences relative to a chief ray, optionally in waves."""
        res = self.get_phase_difference(first=first, last=last)
        if referenceray is not None:
            res = res - res[referenceray]
        if wavelength is not None:
            res = res / wavelength
        return res
