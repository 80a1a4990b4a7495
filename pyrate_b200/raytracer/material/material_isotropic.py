"""Isotropic media (host API mirror of reference
raytracer/material/material_isotropic.py: IsotropicMaterial :38-247,
ConstantIndexGlass :250-265, ModelGlass :268-360).  Only the scalar index is
evaluated on the host (once per surface per trace); Snell's law runs in the
native per-surface step."""
import numpy as np

from ...core import FloatOptimizableVariable
from ..globalconstants import standard_wavelength
from .material import MaxwellMaterial


class IsotropicMaterial(MaxwellMaterial):

    def setKind(self):
        self.kind = "isotropicmaterial"

    def get_optical_index(self, xpos, wave):
        raise NotImplementedError()

    def get_isotropic_epsilon(self, xpos, wave=standard_wavelength):
        return self.get_optical_index(xpos, wave=wave) ** 2

    def get_epsilon_tensor(self, xpos, wave=standard_wavelength):
        n = np.shape(xpos)[1]
        eps = np.zeros((3, 3, n))
        for i in range(3):
            eps[i, i, :] = 1.
        return eps * self.get_isotropic_epsilon(xpos, wave=wave)


class ConstantIndexGlass(IsotropicMaterial):

    @classmethod
    def p(cls, lc, n=1.0, name="", comment=""):
        return cls({"comment": comment},
                   {"lc": lc, "n": FloatOptimizableVariable(n, name="refractive index")},
                   name=name)

    def setKind(self):
        self.kind = "constantindexglass"

    def get_optical_index(self, x, wave):
        return self.n.evaluate()


class ModelGlass(IsotropicMaterial):
    """Conrady model n = n0 + A / wave + B / wave**3.5 (reference :299-309)."""

    @classmethod
    def p(cls, lc, n0_A_B=(1.49749699179, 0.0100998734374 * 1e-3,
                           0.000328623343942 * (1e-3) ** 3.5),
          name="", comment=""):
        (n0, a, b) = n0_A_B
        return cls({"comment": comment},
                   {"lc": lc,
                    "n0": FloatOptimizableVariable(n0, name="Conrady n0"),
                    "A": FloatOptimizableVariable(a, name="Conrady A"),
                    "B": FloatOptimizableVariable(b, name="Conrady B")},
                   name=name)

    def setKind(self):
        self.kind = "modelglass"

    def get_optical_index(self, x, wave):
        return self.n0() + self.A() / wave + self.B() / (wave ** 3.5)

    def calcCoefficientsFrom_nd_vd_PgF(self, nd=1.51680, vd=64.17, PgF=0.5349):
        # same (swapped A/B) assignment as the reference :322-325
        nf_minus_nc = (nd - 1) / vd
        b = (0.454670392956 * nf_minus_nc * (PgF - 0.445154791693)) * (1e-3) ** 3.5
        a = (1.87513751845 * nf_minus_nc - b * 15.2203074842) * 1e-3
        n0 = nd - 1.70194862906e3 * a - 6.43150432188 * (1e3 ** 3.5) * b
        self.n0.setvalue(n0)
        self.A.setvalue(b)
        self.B.setvalue(a)

    def calcCoefficientsFrom_nd_vd(self, nd=1.51680, vd=64.17):
        self.calcCoefficientsFrom_nd_vd_PgF(nd, vd, 0.6438 - 0.001682 * vd)
