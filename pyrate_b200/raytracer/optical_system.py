"""Optical system (host API mirror of reference raytracer/optical_system.py
:39-94, :217-227): dict of elements + background material, and `seqtrace`,
THE path this package accelerates.

    paths = s.seqtrace(initialbundle, elementsequence, splitup=False)

keeps the reference's contract -- the caller's bundle is not mutated, the result
is `list[RayPath]`, path = [b0, b0, b1, ..., bS] with (P, 3, N) bundle rows --
but the rays live in CUDA tensors and the whole element sequence runs as one
persistent sm_100a kernel.
"""
import numpy as np

from .localcoordinates import LocalCoordinates, LocalCoordinatesTreeBase
from .ray import RayPath
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

    def seqtrace_batch(self, bundles, elementsequence, record_efield=False):
        """`[self.seqtrace(b, elementsequence) for b in bundles]` for bundles of different
        wavelength (e.g. the F, d, C bundles of demos/demo_doublegauss.py:189-213) in ONE
        native launch: dispersion is resolved per wavelength at lowering time and the
        kernel selects the media indices per ray.  Returns one list[RayPath] per bundle."""
        from .. import engine
        return engine.seqtrace_batch(self, bundles, elementsequence, record_e=record_efield)

    def para_seqtrace(self, pilotbundle, initialbundle, elementsequence,
                      pilotraypathsequence=None, use6x6=True,
                      pilotbundle_generation="complex"):
        """Linearised trace about a pilot ray, element by element (reference :105-128).
        Returns (pilot RayPath, RayPath of the linearised bundles)."""
        rpath = RayPath(initialbundle)
        pilotpath = RayPath(pilotbundle)
        if pilotraypathsequence is None:
            pilotraypathsequence = tuple(0 for _ in elementsequence)
        for ((elem, subseq), prp_nr) in zip(elementsequence, pilotraypathsequence):
            (append_pilotpath, append_rpath) = self.elements[elem].para_seqtrace(
                pilotpath.raybundles[-1], rpath.raybundles[-1], subseq,
                self.material_background, pilotraypath_nr=prp_nr,
                pilotbundle_generation=pilotbundle_generation)
            rpath.appendRayPath(append_rpath)
            pilotpath.appendRayPath(append_pilotpath)
        return (pilotpath, rpath)

    def sequence_to_hitlist(self, elementsequence):
        return [(elem, self.elements[elem].sequence_to_hitlist(seq))
                for (elem, seq) in elementsequence]

    def extractXYUV(self, pilotbundle, elementsequence, pilotraypathsequence=None,
                    pilotbundle_generation="complex"):
        """(object -> stop, stop -> image) transfer matrices: products of the
        per-surface-pair XYUV matrices on either side of the one surface flagged
        `is_stop` (reference :134-214).  None (with a warning) unless exactly one stop
        is flagged."""
        pilotpath = RayPath(pilotbundle)
        if pilotraypathsequence is None:
            pilotraypathsequence = tuple(0 for _ in elementsequence)
        stops_found = sum(1 for (_, subseq) in elementsequence
                          for (_, options_dict) in subseq
                          if options_dict.get("is_stop", False))
        if stops_found != 1:
            self.warning("%d stops found. need exactly 1!" % (stops_found,))
            return None
        size = 6 if pilotbundle_generation.lower() == "complex" else 4
        lst_matrix_pairs = []
        for ((elem, subseq), prp_nr) in zip(elementsequence, pilotraypathsequence):
            (hitlist, optionshitlist_dict) = self.elements[elem].sequence_to_hitlist(subseq)
            (append_pilotpath, elem_matrices) = self.elements[elem].calculateXYUV(
                pilotpath.raybundles[-1], subseq, self.material_background,
                pilotraypath_nr=prp_nr, pilotbundle_generation=pilotbundle_generation)
            pilotpath.appendRayPath(append_pilotpath)
            (m1, m2) = (np.eye(size), np.eye(size))
            found_stop = False
            for h in hitlist:
                (d1, d2) = optionshitlist_dict[h]
                if d1.get("is_stop", False) and not d2.get("is_stop", False):
                    found_stop = True
                if not found_stop:
                    m1 = np.dot(elem_matrices[h], m1)
                else:
                    m2 = np.dot(elem_matrices[h], m2)
            lst_matrix_pairs.append((m1, m2, found_stop))
        (m_obj_stop, m_stop_img) = (np.eye(size), np.eye(size))
        obj_stop_branch = True
        for (m1, m2, found_stop) in lst_matrix_pairs:
            if obj_stop_branch:
                m_obj_stop = np.dot(m1, m_obj_stop)
                if found_stop:
                    m_stop_img = np.dot(m2, m_stop_img)
                    obj_stop_branch = False
            else:
                m_stop_img = np.dot(m1, m_stop_img)
        return (m_obj_stop, m_stop_img)

    def seqtrace(self, initialbundle, elementsequence, splitup=False,
                 record_efield=False, grin_history=False, grin_lockstep=False):
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
        :param grin_lockstep: GRIN segments follow the reference's loop literally --
                        all rays step until every ray is final, energy test summed
                        over the bundle (material_grin.py:139, :164-176) -- instead
                        of the per-ray normalisation; for small bundles
        :return: list[RayPath]
        """
        from .. import engine
        return engine.seqtrace(self, initialbundle, elementsequence,
                               splitup=splitup, record_e=record_efield,
                               grin_history=grin_history, grin_lockstep=grin_lockstep)
