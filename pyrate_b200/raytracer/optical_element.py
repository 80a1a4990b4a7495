"""Optical element = surfaces + materials + surface->(material-, material+) map
(host API mirror of reference raytracer/optical_element.py:36-130, :324-379).

`seqtrace` keeps the reference signature; the per-surface loop itself is lowered
into one persistent native launch (pyrate_b200/lowering.py, engine.py).
"""
from .localcoordinates import LocalCoordinatesTreeBase


class OpticalElement(LocalCoordinatesTreeBase):

    @classmethod
    def p(cls, lc, name=""):
        return cls({"surf_mat_connection": {}},
                   {"surfaces": {}, "materials": {}, "rootcoordinatesystem": lc},
                   name=name)

    def setKind(self):
        self.kind = "opticalelement"

    def addSurface(self, key, surface_object, materialkeys):
        (minus_key, plus_key) = materialkeys
        if not self.checkForRootConnection(surface_object.rootcoordinatesystem):
            raise Exception("surface coordinate system should be connected "
                            "to OpticalElement root coordinate system")
        self.surfaces[key] = surface_object
        self.annotations["surf_mat_connection"][key] = (minus_key, plus_key)

    def changeMaterialsForSurface(self, key, materialkeys):
        if key in self.annotations["surf_mat_connection"]:
            self.annotations["surf_mat_connection"][key] = tuple(materialkeys)

    def getSurfaces(self):
        return self.surfaces

    def getConnection(self, key):
        return self.annotations["surf_mat_connection"][key]

    def addMaterial(self, key, material_object, comment=""):
        if not self.checkForRootConnection(material_object.lc):
            raise Exception("material coordinate system should be connected "
                            "to OpticalElement root coordinate system")
        if key not in self.materials:
            self.materials[key] = material_object
            self.materials[key].comment = comment
        else:
            self.warning("Material key " + str(key) +
                         " already taken. Material will not be added.")

    def findoutWhichMaterial(self, mat1, mat2, current_mat):
        """Toggle by object identity (reference :109-126)."""
        return mat2 if (mat1 is current_mat) else mat1

    def sequence_to_hitlist(self, seq):
        counts = {}
        keys = []
        for s in seq:
            counts[s] = counts.get(s, 0) + 1
            keys.append((s, counts[s]))
        return list(zip(keys[:-1], keys[1:]))

    def seqtrace(self, raybundle, sequence, background_medium, splitup=False):
        """Trace `raybundle` through `sequence` of this element only; returns
        list[RayPath] whose first bundle is `raybundle`'s traced copy."""
        from .. import engine

        class _OneElementSystem(object):
            pass
        sysview = _OneElementSystem()
        sysview.material_background = background_medium
        sysview.elements = {"_self": self}
        paths = engine.seqtrace(sysview, raybundle, [("_self", sequence)],
                                splitup=splitup)
        for p in paths:           # element-level paths start with ONE copy
            p.raybundles = p.raybundles[1:]
        return paths
