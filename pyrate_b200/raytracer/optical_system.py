"""Optical system (host API mirror of reference raytracer/optical_system.py
:39-94, :217-227): dict of elements + background material, and `seqtrace`,
THE path this package accelerates.

    paths = s.seqtrace(initialbundle, elementsequence, splitup=False)

keeps the reference's contract -- the caller's bundle is not mutated, the result
is `list[RayPath]`, path = [b0, b0, b1, ..., bS] with (P, 3, N) bundle rows --
but the rays live in CUDA tensors and the whole element sequence runs as one
persistent sm_100a kernel.
"""
from .localcoordinates import LocalCoordinates, LocalCoordinatesTreeBase
from .material.material_isotropic import ConstantIndexGlass


class OpticalSystem(LocalCoordinatesTreeBase):

    @classmethod
    def p(cls, rootlc=None, matbackground=None, name=""):
        if rootlc is None:
            rootlc = LocalCoordinates.p(name="global")
        if matbackground is None:
            matbackground = ConstantIndexGlass.p(rootlc, 1.0, name="background")
        return cls({}, {"rootcoordinatesystem": rootlc,
                        "material_background": matbackground,
                        "elements": {}}, name=name)

    def setKind(self):
        self.kind = "opticalsystem"

    def addElement(self, key, element):
        if not self.checkForRootConnection(element.rootcoordinatesystem):
            raise Exception("OpticalElement root should be connected to root "
                            "of OpticalSystem")
        self.elements[key] = element

    def removeElement(self, key):
        if key in self.elements:
            self.elements.pop(key)

    def seqtrace(self, initialbundle, elementsequence, splitup=False,
                 record_efield=False, grin_history=False):
        """Sequential trace on the GPU.

        :param initialbundle: RayBundle (never written)
        :param elementsequence: [(elemkey, [(surfkey, {"is_mirror": bool,
                                "is_stop": bool}), ...]), ...]
        :param splitup: fork the path at birefringent interfaces instead of
                        doubling the rays (material_anisotropic.py:87-113)
        :param record_efield: also transport and record E through isotropic
                        media (otherwise `Efield` of isotropic bundles is
                        produced on demand as some unit vector perpendicular to
                        k -- the reference's own choice is SVD-arbitrary)
        :param grin_history: record every integrator step of GRIN segments in the
                        bundle rows like the reference does (small bundles only:
                        steps x rays x 49 B)
        :return: list[RayPath]
        """
        from .. import engine
        return engine.seqtrace(self, initialbundle, elementsequence,
                               splitup=splitup, record_e=record_efield,
                               grin_history=grin_history)
