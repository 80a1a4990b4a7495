"""Homogeneous anisotropic medium (host API mirror of reference
raytracer/material/material_anisotropic.py:34-155).  The eigenpolarisation
solve and the o/e ray split run in the native complex kernel."""
import numpy as np

from ..globalconstants import standard_wavelength
from .material import MaxwellMaterial


class AnisotropicMaterial(MaxwellMaterial):

    @classmethod
    def p(cls, lc, epstensor, name="", comment=""):
        return cls({"comment": comment,
                    "epstensor": np.asarray(epstensor).tolist()},
                   {"lc": lc}, name=name)

    def initialize_from_annotations(self):
        self.epstensor = np.array(self.annotations["epstensor"])

    def get_epsilon_tensor(self, x, wave=standard_wavelength):
        return np.repeat(self.epstensor[:, :, np.newaxis], np.shape(x)[1], axis=2)
