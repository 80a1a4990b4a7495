"""Element-level analysis (host API mirror of reference
raytracer/analysis/optical_element_analysis.py:33-66)."""
import numpy as np


class OpticalElementAnalysis(object):

    def __init__(self, oe, elemseq, name=""):
        self.opticalelement = oe
        self.elementsequence = elemseq
        self.name = name

    def calc_xyuv(self, parthitlist, pilotbundle, fullsequence, background_medium):
        """Product of the XYUV matrices along `parthitlist` (real 4x4 convention of
        the reference, :46-66); the pilot bundle is traced by the native engine."""
        (_, matrices) = self.opticalelement.calculateXYUV(
            pilotbundle, fullsequence, background_medium, pilotbundle_generation="real")
        tmp = np.eye(4)
        for hit in parthitlist:
            tmp = np.dot(matrices[hit], tmp)
        return tmp
