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
from .raytracer.material.material_grin import IsotropicGrinMaterial  # noqa: F401
from .raytracer.material.material_isotropic import (ConstantIndexGlass,  # noqa: F401
                                                    ModelGlass)
from .raytracer.optical_element import OpticalElement  # noqa: F401
from .raytracer.optical_system import OpticalSystem  # noqa: F401
from .raytracer.ray import RayBundle, RayPath  # noqa: F401
from .raytracer.surface import Surface  # noqa: F401
from .raytracer.surface_shape import (Asphere, Conic, XYPolynomials,  # noqa: F401
                                      accessible_shapes)

__version__ = "0.1.0"


def api():
    """Class namespace for `pyrate_b200.configs.build_system(spec, api)`."""
    return types.SimpleNamespace(
        OpticalSystem=OpticalSystem, OpticalElement=OpticalElement,
        LocalCoordinates=LocalCoordinates, Surface=Surface, Conic=Conic,
        Asphere=Asphere, XYPolynomials=XYPolynomials, BaseAperture=BaseAperture,
        CircularAperture=CircularAperture, RectangularAperture=RectangularAperture,
        ConstantIndexGlass=ConstantIndexGlass, ModelGlass=ModelGlass,
        AnisotropicMaterial=AnisotropicMaterial,
        IsotropicGrinMaterial=IsotropicGrinMaterial, RayBundle=RayBundle,
        RayPath=RayPath)
