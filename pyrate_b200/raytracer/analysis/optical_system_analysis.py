"""Bundle generation and trace wrappers (host API mirror of reference
raytracer/analysis/optical_system_analysis.py:43-320): the callers right above
`seqtrace`.  `aim()` describes the bundle by a generator (pyrate_b200.bundlegen:
raster, radius, start, direction, index -- 176 bytes) that the trace kernel expands in
registers, so a bundle of 1e7 rays never crosses the PCIe bus and is never read from
HBM; `collimated_bundle` / `divergent_bundle` return the reference's explicit
(origin, k, E) arrays, built in O(N) NumPy -- the reference runs a 3x3 generalised
eigen-solve PER RAY for this even in vacuum (material/material.py:476-497).  Every
trace goes through the native engine.
"""
import numpy as np

from ...sampling2d.raster import RectGrid
from ..globalconstants import degree, standard_wavelength
from ..ray import RayBundle
from .ray_analysis import RayBundleAnalysis


def _perpendicular_unit(direction):
    """Some unit E with E.d = 0 per column (the reference's choice comes out of an
    eigen-solver and is arbitrary in the plane)."""
    d = np.asarray(direction, dtype=float)
    axis = np.zeros_like(d)
    axis[np.argmin(np.abs(d), axis=0), np.arange(d.shape[1])] = 1.0
    e = axis - np.sum(axis * d, axis=0) / np.sum(d * d, axis=0) * d
    return e / np.sqrt(np.sum(e * e, axis=0))


class OpticalSystemAnalysis(object):

    def __init__(self, os, seq, name=""):
        self.opticalsystem = os
        self.name = name
        self.sequence = seq
        self.field_raster = RectGrid()
        self.pupil_raster = RectGrid()
        self.initial_bundles = None

    # ---- sequence + per-element analyses (reference :63-81) ----
    def set_sequence(self, seq):
        from .optical_element_analysis import OpticalElementAnalysis
        self._sequence = seq
        self.opticalelementanalysis_dict = {
            elem: OpticalElementAnalysis(self.opticalsystem.elements[elem], elemseq,
                                         name="oea_" + elem)
            for (elem, elemseq) in seq}

    def get_sequence(self):
        return self._sequence

    sequence = property(fget=get_sequence, fset=set_sequence)

    # ---- bundle generation (reference :83-165) ----
    def _background_index(self, wave):
        mat = self.opticalsystem.material_background
        names = {c.__name__ for c in type(mat).__mro__}
        if "IsotropicMaterial" not in names or "IsotropicGrinMaterial" in names:
            raise NotImplementedError("bundle generation needs a homogeneous isotropic "
                                      "background medium")
        return float(mat.get_optical_index(None, wave))

    def _bundle(self, origin, unit, wave):
        n = self._background_index(wave)
        return (origin, n * unit, _perpendicular_unit(unit))

    def collimated_bundle(self, nrays, properties_dict=None, wave=standard_wavelength):
        p = properties_dict or {}
        raster = p.get("raster", RectGrid())
        radius = p.get("radius", 1.0)
        (ay, ax) = (p.get("angley", 0.0), p.get("anglex", 0.0))
        (px, py) = raster.getGrid(nrays)
        origin = np.vstack((radius * px + p.get("startx", 0.),
                            radius * py + p.get("starty", 0.),
                            p.get("startz", 0.) * np.ones_like(px)))
        unit = np.empty_like(origin)
        unit[0] = np.sin(ay) * np.cos(ax)
        unit[1] = np.sin(ax)
        unit[2] = np.cos(ay) * np.cos(ax)
        return self._bundle(origin, unit, wave)

    def divergent_bundle(self, nrays, properties_dict=None, wave=standard_wavelength):
        p = properties_dict or {}
        raster = p.get("raster", RectGrid())
        radius = p.get("radius", 45.0 * degree)
        (ay, ax) = (p.get("angley", 0.0), p.get("anglex", 0.0))
        (gx, gy) = raster.getGrid(nrays)
        origin = np.vstack((p.get("startx", 0.) * np.ones_like(gx),
                            p.get("starty", 0.) * np.ones_like(gx),
                            p.get("startz", 0.) * np.ones_like(gx)))
        unit = np.empty_like(origin)
        unit[0] = np.sin(ay + radius * gx) * np.cos(ax + radius * gy)
        unit[1] = np.sin(ax + radius * gy)
        unit[2] = np.cos(ay + radius * gx) * np.cos(ax + radius * gy)
        return self._bundle(origin, unit, wave)

    def bundle_generator(self, nrays, properties_dict=None, bundletype="collimated",
                         wave=standard_wavelength):
        """The bundle of collimated_bundle / divergent_bundle as a device generator
        (pyrate_b200.bundlegen.BundleGen), or None for rasters without one."""
        from ... import bundlegen
        from ..._native import BUNDLE_COLLIMATED, BUNDLE_DIVERGENT
        p = properties_dict or {}
        spec = bundlegen.spec_for_raster(p.get("raster", RectGrid()), nrays)
        if spec is None:
            return None
        (ay, ax) = (p.get("angley", 0.0), p.get("anglex", 0.0))
        start = (p.get("startx", 0.), p.get("starty", 0.), p.get("startz", 0.))
        n = self._background_index(wave)
        if bundletype == "collimated":
            unit = (float(np.sin(ay) * np.cos(ax)), float(np.sin(ax)),
                    float(np.cos(ay) * np.cos(ax)))
            return bundlegen.BundleGen(spec, BUNDLE_COLLIMATED, p.get("radius", 1.0), start,
                                       unit, None, n)
        return bundlegen.BundleGen(spec, BUNDLE_DIVERGENT, p.get("radius", 45.0 * degree),
                                   start, (ay, ax, 0.0), None, n)

    def aim(self, numrays, rays_dict, bundletype="collimated", wave=standard_wavelength):
        import torch
        gen = self.bundle_generator(numrays, rays_dict, bundletype, wave) \
            if torch.cuda.is_available() else None
        if gen is not None:
            # described, not built: the trace kernel generates the rays in registers
            self.initial_bundles = [RayBundle(generator=gen, wave=wave)]
            return
        make = {"collimated": self.collimated_bundle,
                "divergent": self.divergent_bundle}[bundletype]
        (org, kvec, evec) = make(numrays, rays_dict, wave=wave)
        self.initial_bundles = [RayBundle(x0=org, k0=kvec, Efield0=evec, wave=wave)]

    # ---- traces (reference :184-260) ----
    def trace(self, **kwargs):
        return [self.opticalsystem.seqtrace(ib, self.sequence, **kwargs)
                for ib in self.initial_bundles]

    def trace_3d_global(self, x0, k0, wave=standard_wavelength, **kwargs):
        n = np.shape(x0)[1]
        e0 = np.zeros((3, n))
        e0[1] = 1.0                                   # canonical_ey, reference :205-207
        self.initial_bundles = [RayBundle(x0, k0, e0, wave=wave)]
        return [[[(rb.x[0], rb.k[0]) for rb in rp.raybundles] for rp in fp]
                for fp in self.trace(**kwargs)]

    def _flat_sequence(self):
        return [(elem, surf) for (elem, elemseq) in self.sequence for (surf, _) in elemseq]

    def trace_3d_local(self, **kwargs):
        res = self.trace_3d_global(**kwargs)
        flat = self._flat_sequence()
        out = []
        for fp in res:
            fp_out = []
            for rp in fp:
                items = []
                for ((elem, surf), (x, k)) in zip(flat, rp):
                    lc = self.opticalsystem.elements[elem].surfaces[surf].rootcoordinatesystem
                    items.append((lc.returnGlobalToLocalPoints(x),
                                  lc.returnGlobalToLocalDirections(k)))
                fp_out.append(items)
            out.append(fp_out)
        return out

    def trace_2d_local(self, **kwargs):
        return [[[(x[:2], k[:2]) for (x, k) in rp] for rp in fp]
                for fp in self.trace_3d_local(**kwargs)]

    def get_spot(self, raypath):
        """(x, y of the last bundle in the last surface's frame, RMS spot radius
        about the centroid) -- reference :283-303."""
        (last_elem, last_seq) = self.sequence[-1]
        (last_surf, _) = last_seq[-1]
        lc = self.opticalsystem.elements[last_elem].surfaces[last_surf].rootcoordinatesystem
        last = raypath.raybundles[-1]
        xs = lc.returnGlobalToLocalPoints(last.x[-1])
        return (xs[0:2, :], RayBundleAnalysis(last).get_rms_spot_size_centroid())


def raytrace(s, seq, numrays, rays_dict, bundletype="collimated", traceoptions=None,
             wave=standard_wavelength):
    """Convenience function of the reference (pyrateoptics/__init__.py:457-465)."""
    osa = OpticalSystemAnalysis(s, seq)
    osa.aim(numrays, rays_dict, bundletype=bundletype, wave=wave)
    return osa.trace(**(traceoptions or {}))
