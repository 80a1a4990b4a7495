"""Ray state carriers (host API mirror of reference raytracer/ray.py:34-260).

`RayBundle` keeps the reference's attribute names and index layout -- `x`, `k`,
`Efield`: (P, 3, N) with P the history axis, `valid`: (P, N) bool, `rayID`: (N,),
`wave`, `splitted` -- but holds PyTorch tensors (CUDA after a trace; the
constructor also accepts NumPy arrays).  Bundles produced by the native trace
are *views* of the per-step device records: rows are sliced (zero-copy) when
every ray survived and gathered lazily, on first attribute access, when rays
were dropped (the reference compacts at every refract,
material_isotropic.py:194-199).
"""
import numpy as np
import torch

from .globalconstants import standard_wavelength


def as_tensor(a, device=None):
    if isinstance(a, torch.Tensor):
        t = a
    else:
        t = torch.from_numpy(np.ascontiguousarray(a))
    if not (t.is_complex() or t.dtype == torch.float64):
        t = t.to(torch.float64)
    if device is not None and t.device != torch.device(device):
        t = t.to(device)
    return t


class RayBundle(object):
    _FIELDS = ("x", "k", "Efield", "valid", "rayID")

    def __init__(self, x0=None, k0=None, Efield0=None, rayID=None,
                 wave=standard_wavelength, splitted=False, _lazy=None, generator=None):
        self.wave = wave
        self.splitted = splitted
        self._store = {}
        self.generator = generator
        if _lazy is not None:
            self._store.update(_lazy)
            return
        if generator is not None:
            # a bundle described by a generator (pyrate_b200.bundlegen.BundleGen): traced
            # without ever existing in memory; the fields materialise on the device (one
            # pyr_generate_bundle launch) when somebody reads them
            def field(i):
                return lambda: generator.materialise()[i].unsqueeze(0)

            def dev():
                return generator.materialise()[0].device
            self._store = {"x": field(0), "k": field(1), "Efield": field(2),
                           "valid": lambda: torch.ones((1, generator.n), dtype=torch.bool,
                                                       device=dev()),
                           "rayID": lambda: torch.arange(generator.n, device=dev())}
            return
        x0 = as_tensor(x0)
        k0 = as_tensor(k0, x0.device)
        n = x0.shape[1]
        if Efield0 is None or len(Efield0) == 0:
            e0 = torch.zeros((3, n), dtype=torch.float64, device=x0.device)
            e0[1] = 1.0                         # ray.py:71-73
        else:
            e0 = as_tensor(Efield0, x0.device)
        if rayID is None or len(rayID) == 0:
            rid = torch.arange(n, device=x0.device)
        else:
            rid = rayID if isinstance(rayID, torch.Tensor) else \
                torch.as_tensor(np.asarray(rayID), device=x0.device)
        self._store = {"x": x0.reshape(1, 3, n), "k": k0.reshape(1, 3, n),
                       "Efield": e0.reshape(1, 3, n),
                       "valid": torch.ones((1, n), dtype=torch.bool,
                                           device=x0.device),
                       "rayID": rid}

    # ---- lazily materialised fields ----
    def _get(self, name):
        v = self._store[name]
        if callable(v):
            v = v()
            self._store[name] = v
        return v

    x = property(lambda self: self._get("x"),
                 lambda self, v: self._store.__setitem__("x", v))
    k = property(lambda self: self._get("k"),
                 lambda self, v: self._store.__setitem__("k", v))
    Efield = property(lambda self: self._get("Efield"),
                      lambda self, v: self._store.__setitem__("Efield", v))
    valid = property(lambda self: self._get("valid"),
                     lambda self, v: self._store.__setitem__("valid", v))
    rayID = property(lambda self: self._get("rayID"),
                     lambda self, v: self._store.__setitem__("rayID", v))

    @property
    def device(self):
        return self.x.device

    def to(self, device):
        out = RayBundle(_lazy={f: self._get(f).to(device) for f in self._FIELDS},
                        wave=self.wave, splitted=self.splitted)
        return out

    def numpy(self):
        """Dict of NumPy copies (x, k, Efield, valid, rayID)."""
        return {f: self._get(f).detach().cpu().numpy() for f in self._FIELDS}

    def append(self, xnew, knew, Enew, Validnew):
        """ray.py:83-105: validity is cumulative."""
        dev = self.x.device
        xnew = as_tensor(xnew, dev)
        knew = as_tensor(knew, dev)
        Enew = as_tensor(Enew, dev)
        vnew = torch.as_tensor(Validnew, device=dev).to(torch.bool)
        (k_old, e_old) = (self.k, self.Efield)
        if knew.is_complex() and not k_old.is_complex():
            k_old = k_old.to(knew.dtype)
        if Enew.is_complex() and not e_old.is_complex():
            e_old = e_old.to(Enew.dtype)
        self.x = torch.cat((self.x, xnew[None]))
        self.k = torch.cat((k_old, knew[None].to(k_old.dtype)))
        self.Efield = torch.cat((e_old, Enew[None].to(e_old.dtype)))
        self.valid = torch.cat((self.valid, (self.valid[-1] & vnew)[None]))

    def clone(self):
        return RayBundle(_lazy={f: self._get(f).clone() for f in self._FIELDS},
                         wave=self.wave, splitted=self.splitted)

    def returnKtoD(self):
        """Poynting direction for every history row (ray.py:136-152)."""
        (k, e) = (self.k, self.Efield)
        if e.is_complex() or k.is_complex():
            e = e.to(torch.complex128)
            k = k.to(torch.complex128)
            abs_e2 = (e.conj() * e).sum(1, keepdim=True)
            ek = (e * k).sum(1, keepdim=True)
            s = (abs_e2 * k - ek * e.conj()).real
        else:
            abs_e2 = (e * e).sum(1, keepdim=True)
            ek = (e * k).sum(1, keepdim=True)
            s = abs_e2 * k - ek * e
        return s / torch.sqrt((s * s).sum(1, keepdim=True))

    def returnLocalComponents(self, lc, num):
        return (lc.returnGlobalToLocalPoints(self.x[num]),
                lc.returnGlobalToLocalDirections(self.k[num]),
                lc.returnGlobalToLocalDirections(self.Efield[num]))

    def returnLocalD(self, lc, num):
        return lc.returnGlobalToLocalDirections(self.returnKtoD()[num])

    def appendLocalComponents(self, lc, xloc, kloc, Eloc, valid):
        self.append(lc.returnLocalToGlobalPoints(xloc),
                    lc.returnLocalToGlobalDirections(kloc),
                    lc.returnLocalToGlobalDirections(Eloc), valid)

    def getLocalSurfaceNormal(self, surface, material, xglob):
        xl = surface.shape.lc.returnGlobalToLocalPoints(xglob)
        nl = surface.shape.getNormal(xl[0], xl[1])
        return material.lc.returnOtherToActualDirections(nl, surface.shape.lc)


class RayPath(object):

    def __init__(self, initialraybundle=None):
        self.raybundles = [] if initialraybundle is None else [initialraybundle]

    def appendRayBundle(self, raybundle):
        self.raybundles.append(raybundle)

    def appendRayPath(self, raypath):
        self.raybundles += raypath.raybundles

    def containsSplitted(self):
        return any(r.splitted for r in self.raybundles)


def returnDtoK(direction):
    return direction
