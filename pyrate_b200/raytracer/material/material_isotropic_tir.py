"""Isotropic media with the angle-form Snell refraction (host API mirror of
reference raytracer/material/material_isotropic_tir.py: IsotropicMaterialTIR
:36-119, ConstantIndexGlassTIR :122-137).

The reference's `refract` there computes sin(theta') = n1/n2 sin(theta), marks
sin(theta') > 1 as total internal reflection (ray invalid, dropped) and builds
the outgoing direction from the normal and the tangential part; that is the same
map k -> k_par + sqrt(n2^2 - k_par.k_par) n as IsotropicMaterial.refract
(material_isotropic.py:163-199) with the same validity rule, written with angles.
The native step therefore serves both classes (lowering keys on the
`IsotropicMaterial` base); parity of the two forms is pinned by the reference
fixture tests/golden/seqtrace_x13_tirglass.npz.

Deliberate normalisations: the reference takes k in global and the surface normal in
material coordinates (:52 vs :66), which is only right for material frames parallel to
the global frame -- here both live in one frame; the reference's TIR refract does not AND its validity
with the incoming bundle's (`valid = (1-TIR)`, :84), so vignetted or missed
rays come back to life there; here a dead ray stays dead, as in every other
material."""
from ...core import FloatOptimizableVariable
from .material_isotropic import IsotropicMaterial


class IsotropicMaterialTIR(IsotropicMaterial):

    def setKind(self):
        self.kind = "isotropicmaterialTIR"


class ConstantIndexGlassTIR(IsotropicMaterialTIR):

    @classmethod
    def p(cls, lc, n=1.0, name="", comment=""):
        return cls({"comment": comment},
                   {"lc": lc, "n": FloatOptimizableVariable(n, name="refractive index")},
                   name=name)

    def setKind(self):
        self.kind = "constantindexglass"

    def get_optical_index(self, x, wave):
        return self.n.evaluate()
