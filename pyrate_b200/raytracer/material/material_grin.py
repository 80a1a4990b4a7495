"""Isotropic GRIN medium (host API mirror of reference
raytracer/material/material_grin.py:37-220).

The reference evaluates arbitrary user Python (`nfunc`, `dndx`, ... taken from a
source string, core/functionobject.py:99-119) inside its integrator loop.  The
device integrator instead evaluates a closed catalogue of index profiles
(include/pyrate_b200.h, PyrGrinProfile).  The profile is declared in
`annotations["device_profile"]`:

    {"kind": "gaussian_xy", "params": [n0, g, a, b],          # n0 + g exp(-a x^2 - b y^2)
     "boundary": {"kind": "cylinder", "params": [r]}}
    {"kind": "poly_rz", "params": [n0, nr2, nr4, nr6, nz1, nz2, nz3], ...}

and, when Python source is given as in the reference, the lowering pass samples
`nfunc` / `dnd*` / `bnd` on random points and raises if they disagree with the
declared profile -- a mismatch fails loudly instead of tracing another medium.
"""
import numpy as np

from ...core import FloatOptimizableVariable
from ..globalconstants import standard_wavelength
from .material_isotropic import IsotropicMaterial

PROFILE_KINDS = {"gaussian_xy": 0, "poly_rz": 1}
BOUNDARY_KINDS = {"none": 0, "cylinder": 1, "box": 2, "sphere": 3}


def profile_functions(profile):
    """NumPy closures (n, dndx, dndy, dndz, bnd) of a device profile dict."""
    kind = profile["kind"]
    p = list(profile["params"])
    if kind == "gaussian_xy":
        (n0, g, a, b) = p[:4]

        def nfun(x):
            return n0 + g * np.exp(-a * x[0] ** 2 - b * x[1] ** 2)

        def dndx(x):
            return -2. * a * x[0] * g * np.exp(-a * x[0] ** 2 - b * x[1] ** 2)

        def dndy(x):
            return -2. * b * x[1] * g * np.exp(-a * x[0] ** 2 - b * x[1] ** 2)

        def dndz(x):
            return np.zeros_like(x[0])
    elif kind == "poly_rz":
        p = p + [0.0] * (7 - len(p))

        def nfun(x):
            r2 = x[0] ** 2 + x[1] ** 2
            return (p[0] + p[1] * r2 + p[2] * r2 ** 2 + p[3] * r2 ** 3 +
                    p[4] * x[2] + p[5] * x[2] ** 2 + p[6] * x[2] ** 3)

        def _dr(x):
            r2 = x[0] ** 2 + x[1] ** 2
            return 2. * (p[1] + 2. * p[2] * r2 + 3. * p[3] * r2 ** 2)

        def dndx(x):
            return x[0] * _dr(x)

        def dndy(x):
            return x[1] * _dr(x)

        def dndz(x):
            return p[4] + 2. * p[5] * x[2] + 3. * p[6] * x[2] ** 2
    else:
        raise ValueError("unknown GRIN device profile %r" % (kind,))
    bspec = profile.get("boundary", {"kind": "none", "params": []})
    (bk, bp) = (bspec["kind"], list(bspec.get("params", [])))
    if bk == "none":
        def bnd(x):
            return np.ones_like(x[0], dtype=bool)
    elif bk == "cylinder":
        def bnd(x):
            return x[0] ** 2 + x[1] ** 2 < bp[0] ** 2
    elif bk == "box":
        def bnd(x):
            return (np.abs(x[0]) < bp[0]) & (np.abs(x[1]) < bp[1])
    elif bk == "sphere":
        def bnd(x):
            return x[0] ** 2 + x[1] ** 2 + x[2] ** 2 < bp[0] ** 2
    else:
        raise ValueError("unknown GRIN boundary %r" % (bk,))
    return (nfun, dndx, dndy, dndz, bnd)


class IsotropicGrinMaterial(IsotropicMaterial):

    @classmethod
    def p(cls, lc, mysource=None, nfun_name="nfunc", dndx_name="dndx",
          dndy_name="dndy", dndz_name="dndz", bnd_name="bnd",
          parameterlist=None, name="", comment="", device_profile=None, device_source=None):
        params = {}
        for (pname, value) in (parameterlist or []):
            params[pname] = FloatOptimizableVariable(value, name=pname)
        ann = {"comment": comment, "ds": 0.1, "energyviolation": 1e-2,
               "f_name": nfun_name, "dfdx_name": dndx_name,
               "dfdy_name": dndy_name, "dfdz_name": dndz_name,
               "bnd_name": bnd_name, "source": mysource}
        if device_profile is not None:
            ann["device_profile"] = device_profile
        if device_source is not None:
            # {"n": ..., "dndx": ..., "dndy": ..., "dndz": ..., "inside": ..., "params": [...]}:
            # CUDA C++ expressions in x, y, z (material frame) and p[] -- the device
            # counterpart of the Python source above, compiled by pyrate_b200/grin_jit.py
            ann["device_source"] = device_source
        return cls(ann, {"lc": lc, "params": params}, name=name)

    def initialize_from_annotations(self):
        self._user_functions = None

    def user_functions(self):
        """(nfunc, dndx, dndy, dndz, bnd) from the Python source, or None.

        The source is only ever executed to VERIFY the declared device profile
        (never inside the trace)."""
        src = self.annotations.get("source")
        if not src:
            return None
        if self._user_functions is None:
            env = {}
            exec(src, env)                      # same trust model as the reference
            kw = {k: v() for (k, v) in self.params.items()}
            names = [self.annotations[k] for k in
                     ("f_name", "dfdx_name", "dfdy_name", "dfdz_name")]
            funcs = [(lambda x, f=env[nm]: f(x, **kw)) for nm in names]
            funcs.append(env[self.annotations["bnd_name"]])
            self._user_functions = tuple(funcs)
        return self._user_functions

    def device_profile(self):
        prof = self.annotations.get("device_profile")
        if prof is None and self.annotations.get("device_source") is not None:
            return None
        if prof is None:
            raise NotImplementedError(
                "IsotropicGrinMaterial %r has no annotations['device_profile']: the "
                "device integrator only evaluates the catalogue profiles "
                "(gaussian_xy, poly_rz); arbitrary Python index functions are not "
                "executed on the GPU" % (self.name,))
        return prof

    def get_optical_index(self, x, wave=standard_wavelength):
        prof = self.device_profile()
        if prof is None:                          # user profile: the Python source itself
            return self.user_functions()[0](np.asarray(x))
        return profile_functions(prof)[0](np.asarray(x))

    def in_boundary(self, pos):
        prof = self.device_profile()
        if prof is None:
            return self.user_functions()[4](np.asarray(pos))
        return profile_functions(prof)[4](np.asarray(pos))
