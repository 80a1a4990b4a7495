"""Material base classes (host API mirror of reference
raytracer/material/material.py:36-96).  `propagate` mutates the bundle,
`refract` / `reflect` return fresh bundles -- the reference's ownership rules
-- but every ray operation is a launch of the native engine."""
from ...core import ClassWithOptimizableVariables
from ..globalconstants import standard_wavelength


class Material(ClassWithOptimizableVariables):

    @classmethod
    def p(cls, lc, name="", comment=""):
        return cls({"comment": comment}, {"lc": lc}, name=name)

    def setKind(self):
        self.kind = "material"

    def refract(self, raybundle, actualSurface, splitup=False):
        from ... import engine
        return engine.material_deflect(self, raybundle, actualSurface,
                                       mirror=False, splitup=splitup)

    def reflect(self, raybundle, actualSurface, splitup=False):
        from ... import engine
        return engine.material_deflect(self, raybundle, actualSurface,
                                       mirror=True, splitup=splitup)

    def propagate(self, raybundle, nextSurface):
        from ... import engine
        engine.material_propagate(self, raybundle, nextSurface)


class MaxwellMaterial(Material):

    def setKind(self):
        self.kind = "maxwellmaterial"

    def get_epsilon_tensor(self, x, wave=standard_wavelength):
        raise NotImplementedError()
