"""TEST INFRASTRUCTURE ONLY -- loader for the unmodified reference (mess42/pyrate).

Makes `/root/reference` importable under NumPy 2 / without matplotlib by
installing three monkey-patches *before* `import pyrateoptics` (SURVEY.md
Appendix C).  Nothing under /root/reference is modified.  In the build container
it loads /root/reference; on the GPU box (no /root/reference) it loads the copy
of the unmodified package that `oracle/make_ref.sh` staged under `oracle/_ref/`
(git-ignored).  Used by `oracle/gen_golden.py` to produce the committed fixtures
in `tests/golden/`, by the `not gpu` tests that cross-check the NumPy restatement
against the live reference when it is present, and by `bench.py`'s CPU arm.
"""
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REFERENCE_ROOT = os.environ.get("PYRATE_REFERENCE_ROOT", "/root/reference")
if not os.path.isdir(os.path.join(REFERENCE_ROOT, "pyrateoptics")) and \
        os.path.isdir(os.path.join(_STAGED, "pyrateoptics")):
    # the GPU box has no /root/reference: the unmodified package staged by
    # oracle/make_ref.sh (bench.py's CPU arm, kind "reference")
    REFERENCE_ROOT = _STAGED


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pyrateoptics"))


def install():
    """Install shims and put the reference on sys.path. Idempotent."""
    import numpy as np
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    # reference localcoordinates.py:82-83,275 and helpers_math.py:82 use np.lib.eye
    if not hasattr(np.lib, "eye"):
        np.lib.eye = np.eye
    if not hasattr(np, "float"):      # reference tests only
        np.float = float
    # pyrateoptics/__init__.py:37-39 imports matplotlib at top level (draw only)
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.widgets",
                 "matplotlib.testing", "matplotlib.testing.decorators"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].__version__ = "3.0.0"
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib.widgets"].Slider = object
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import logging
    logging.disable(logging.CRITICAL)


def api():
    """Namespace of the reference classes the config builder needs."""
    install()
    from pyrateoptics.raytracer.optical_system import OpticalSystem
    from pyrateoptics.raytracer.optical_element import OpticalElement
    from pyrateoptics.raytracer.localcoordinates import LocalCoordinates
    from pyrateoptics.raytracer.surface import Surface
    from pyrateoptics.raytracer.surface_shape import (Conic, Cylinder, Asphere, Biconic,
                                                      XYPolynomials, ZernikeFringe,
                                                      ZernikeANSI, GridSag,
                                                      LinearCombination)
    from pyrateoptics.raytracer.aperture import (BaseAperture,
                                                 CircularAperture,
                                                 RectangularAperture)
    from pyrateoptics.raytracer.material.material_isotropic import (
        ConstantIndexGlass, ModelGlass)
    from pyrateoptics.raytracer.material.material_isotropic_tir import (
        ConstantIndexGlassTIR)
    from pyrateoptics.raytracer.material.material_anisotropic import (
        AnisotropicMaterial)
    from pyrateoptics.raytracer.material.material_grin import (
        IsotropicGrinMaterial)
    from pyrateoptics.raytracer.ray import RayBundle, RayPath
    return types.SimpleNamespace(**locals())
