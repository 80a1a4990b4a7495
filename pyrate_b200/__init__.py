"""pyrate_b200 -- Blackwell-native sequential optical trace engine behind the
pyrateoptics API (OpticalSystem.seqtrace and the classes it touches).

Host classes mirror the reference's names and signatures; all ray arithmetic
runs in hand-written sm_100a CUDA (pyrate_b200/csrc) reached through the C ABI
of include/pyrate_b200.h.  There is no CPU fallback.
"""
import types

from .raytracer.aperture import (BaseAperture, CircularAperture,  # noqa: F401
                                 RectangularAperture)
from .raytracer.globalconstants import (degree, numerical_tolerance,  # noqa: F401
                                        standard_wavelength)
from .raytracer.localcoordinates import LocalCoordinates  # noqa: F401
from .raytracer.material.material_anisotropic import AnisotropicMaterial  # noqa: F401
from .raytracer.material.material_glasscat import CatalogMaterial  # noqa: F401
from .raytracer.material.material_grin import IsotropicGrinMaterial  # noqa: F401
from .raytracer.material.material_isotropic import (ConstantIndexGlass,  # noqa: F401
                                                    ModelGlass)
from .raytracer.material.material_isotropic_tir import (ConstantIndexGlassTIR,  # noqa: F401
                                                        IsotropicMaterialTIR)
from .raytracer.optical_element import OpticalElement  # noqa: F401
from .raytracer.optical_system import OpticalSystem  # noqa: F401
from .raytracer.ray import RayBundle, RayPath  # noqa: F401
from .raytracer.surface import Surface  # noqa: F401
from .raytracer.surface_shape import (Asphere, Biconic, Conic, Cylinder, GridSag,  # noqa: F401
                                      LinearCombination, XYPolynomials,
                                      ZernikeANSI, ZernikeFringe,
                                      accessible_shapes)

__version__ = "0.1.0"


def api():
    """Class namespace for `pyrate_b200.configs.build_system(spec, api)`."""
    return types.SimpleNamespace(
        OpticalSystem=OpticalSystem, OpticalElement=OpticalElement,
        LocalCoordinates=LocalCoordinates, Surface=Surface, Conic=Conic, Cylinder=Cylinder,
        Asphere=Asphere, Biconic=Biconic, XYPolynomials=XYPolynomials,
        ZernikeFringe=ZernikeFringe, ZernikeANSI=ZernikeANSI,
        GridSag=GridSag, LinearCombination=LinearCombination,
        BaseAperture=BaseAperture,
        CircularAperture=CircularAperture, RectangularAperture=RectangularAperture,
        ConstantIndexGlass=ConstantIndexGlass, ModelGlass=ModelGlass,
        ConstantIndexGlassTIR=ConstantIndexGlassTIR,
        AnisotropicMaterial=AnisotropicMaterial,
        IsotropicGrinMaterial=IsotropicGrinMaterial, RayBundle=RayBundle,
        RayPath=RayPath)


# ---------------------------------------------------------------------------
# convenience builders (mirror of reference pyrateoptics/__init__.py:83-258)
# ---------------------------------------------------------------------------
def build_rotationally_symmetric_optical_system(builduplist, **kwargs):
    """builduplist rows: (r, cc, thickness, mat, name, optdict); `r` is the radius
    of curvature (abs(r) <= 1e-17 -> plane), `thickness` the distance from the
    PREVIOUS surface (decz), `mat` None or something convertible to float
    (-> ConstantIndexGlass).  Returns (OpticalSystem, sequence)."""
    rows = []
    for (r, cc, thickness, mat, name, optdict) in builduplist:
        curv = 1. / r if abs(r) > numerical_tolerance else 0.
        rows.append(({"shape": "Conic", "curv": curv, "cc": cc},
                     {"decz": thickness}, mat, name, optdict))
    return build_simple_optical_system(rows, **kwargs)


def build_simple_optical_element(lc0, builduplist, material_db_path="", name=""):
    """builduplist rows: (surfdict, coordbreakdict, mat, name, optdict) with
    surfdict = {"shape": "Conic" | "Asphere" | "XYPolynomials", <shape kwargs>,
    "aperture": None | BaseAperture | {"type": ..., ...}}."""
    elem = OpticalElement.p(lc0, name=name)
    refname = lc0.name
    lastmat = None
    seq = []
    for (surfdict, coordbreakdict, mat, surf_name, optdict) in builduplist:
        surfdict = dict(surfdict)
        lc = elem.addLocalCoordinateSystem(
            LocalCoordinates.p(name=surf_name + "_lc", **coordbreakdict),
            refname=refname)
        shapetype = "shape_" + surfdict.pop("shape", "Conic")
        aperture = surfdict.pop("aperture", None)
        if shapetype not in accessible_shapes:
            raise NotImplementedError(
                "%s is outside the sequential-trace engine (supported: %s)" %
                (shapetype, ", ".join(sorted(accessible_shapes))))
        surf = Surface.p(lc, name=surf_name + "_surf", aperture=aperture,
                         shape=accessible_shapes[shapetype].p(
                             lc, name=name + "_shape", **surfdict))
        if mat is not None:
            try:
                n = float(mat)
            except (TypeError, ValueError):
                raise NotImplementedError(
                    "material %r: only constant-index glasses (floats) are built "
                    "here; the refractiveindex.info catalogue reader is outside "
                    "the seqtrace path (add the material object with "
                    "OpticalElement.addMaterial instead)" % (mat,))
            mat = "constantindexglass_" + str(mat)
            elem.addMaterial(mat, ConstantIndexGlass.p(lc, n=n))
        elem.addSurface(surf_name, surf, (lastmat, mat))
        lastmat = mat
        refname = lc.name
        seq.append((surf_name, optdict))
    return (elem, (name, seq))


def build_simple_optical_system(builduplist, material_db_path="", name=""):
    s = OpticalSystem.p(name=name)
    lc0 = s.addLocalCoordinateSystem(LocalCoordinates.p(name="object", decz=0.0),
                                     refname=s.rootcoordinatesystem.name)
    (elem, elem_seq) = build_simple_optical_element(
        lc0, builduplist, material_db_path=material_db_path, name="stdelem")
    s.addElement("stdelem", elem)
    s.material_background.set_name("background")
    return (s, [elem_seq])

from .raytracer.analysis.optical_system_analysis import (OpticalSystemAnalysis,  # noqa: E402,F401
                                                         raytrace)
from .raytracer.analysis.ray_analysis import RayBundleAnalysis  # noqa: E402,F401
from .raytracer.analysis.optical_element_analysis import OpticalElementAnalysis  # noqa: E402,F401
from .raytracer.aim import Aimy  # noqa: E402,F401
from .raytracer.helpers import (build_pilotbundle, build_pilotbundle_complex,  # noqa: E402,F401
                                choose_nearest)
