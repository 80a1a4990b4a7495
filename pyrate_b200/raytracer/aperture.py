"""Apertures (host API mirror of reference raytracer/aperture.py:33-154).

Same classes, `.p(...)` signatures, annotations ("minradius", "maxradius",
"width", "height", "typicaldimension") and factory.  The mask itself is fused
into the native per-surface step; `are_points_in_aperture` is kept for callers
that test points on the host (NumPy or torch).
"""
import math

from ..core import ClassWithOptimizableVariables


class BaseAperture(ClassWithOptimizableVariables):

    @classmethod
    def p(cls, lc, name="", *_):
        return cls({"typicaldimension": 1e16}, {"lc": lc}, name=name)

    def setKind(self):
        self.kind = "aperture"

    def get_typical_dimension(self):
        return self.annotations["typicaldimension"]

    def get_boolean_function(self):
        return lambda x, y: (x == x) | True

    def are_points_in_aperture(self, x_intersection, y_intersection):
        return self.get_boolean_function()(x_intersection, y_intersection)


class CircularAperture(BaseAperture):

    @classmethod
    def p(cls, lc, maxradius=1.0, minradius=0.0, name="", *_):
        return cls({"maxradius": maxradius, "minradius": minradius,
                    "typicaldimension": maxradius}, {"lc": lc}, name=name)

    def setKind(self):
        self.kind = "aperture_Circular"

    def get_boolean_function(self):
        (rmin, rmax) = (self.annotations["minradius"], self.annotations["maxradius"])
        return lambda x, y: ((x * x + y * y >= rmin ** 2) &
                             (x * x + y * y <= rmax ** 2))


class RectangularAperture(BaseAperture):

    @classmethod
    def p(cls, lc, width=1.0, height=1.0, name="", *_):
        return cls({"width": width, "height": height,
                    "typicaldimension": math.sqrt(width ** 2 + height ** 2)},
                   {"lc": lc}, name=name)

    def setKind(self):
        self.kind = "aperture_Rectangle"

    def get_boolean_function(self):
        (w, h) = (self.annotations["width"], self.annotations["height"])
        return lambda x, y: ((x >= -w * 0.5) & (x <= w * 0.5) &
                             (y >= -h * 0.5) & (y <= h * 0.5))


ACCESSIBLE_APERTURES = {None: BaseAperture,
                        "CircularAperture": CircularAperture,
                        "RectangularAperture": RectangularAperture}


def create_aperture(localcoordinates, ap_dict):
    ap_dict = dict(ap_dict)
    ap_type = ap_dict.pop("type", None)
    return ACCESSIBLE_APERTURES[ap_type].p(localcoordinates, **ap_dict)
