"""Optical element = surfaces + materials + surface->(material-, material+) map
(host API mirror of reference raytracer/optical_element.py:36-130, :324-379).

`seqtrace` keeps the reference signature; the per-surface loop itself is lowered
into one persistent native launch (pyrate_b200/lowering.py, engine.py).
`calculateXYUV` / `para_seqtrace` (reference :165-322, :381-469) are the paraxial
callers of that path: they trace a small pilot bundle natively and fit / apply linear
transfer matrices (xyuv.py).
"""
from . import xyuv
from .localcoordinates import LocalCoordinatesTreeBase
from .ray import RayBundle, RayPath, as_tensor


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
        """(hitlist, optionshitlistdict): surface pairs with a hit counter, which tells
        multiple passes between the same two surfaces apart (reference :128-151)."""
        return xyuv.sequence_to_hitlist(seq)

    def hitlist_to_sequence(self, hitlist_pair):
        return xyuv.hitlist_to_sequence(hitlist_pair)

    def calculateXYUV(self, pilotinitbundle, sequence, background_medium,
                      pilotraypath_nr=0, pilotbundle_generation="complex"):
        """Trace the pilot bundle through `sequence` (native engine) and fit the linear
        transfer matrix of every surface pair in the local x, y, kx, ky (+ Im kx, Im ky)
        coordinates of the two surfaces (reference :165-322).  Returns
        (pilotraypath, {(s1, s2, hit): matrix, (s2, s1, hit): inverse fit})."""
        (hitlist, _) = self.sequence_to_hitlist(sequence)
        pilotraypaths = self.seqtrace(pilotinitbundle, sequence, background_medium,
                                      splitup=True)
        self.info("found %d pilotraypaths" % (len(pilotraypaths),))
        self.info("selected no %d via pilotraypath_nr parameter" % (pilotraypath_nr,))
        pilotraypath = pilotraypaths[pilotraypath_nr]
        bundles = pilotraypath.raybundles
        (px, pk) = ([], [])
        width = bundles[0].x.shape[-1]
        for b in bundles[:len(hitlist) + 1]:
            if b.x.shape[-1] != width or not bool(b.valid[-1].all()):
                raise Exception("pilot bundle lost rays on its way: XYUV matrices need "
                                "every pilot ray at every surface (smaller pilot "
                                "deltas, or a different pilot ray)")
            px.append(b.x[-1])
            pk.append(b.k[-1])
        matrices = xyuv.transfer_matrices(self.surfaces, hitlist, px, pk,
                                          pilotbundle_generation)
        return (pilotraypath, matrices)

    def para_seqtrace(self, pilotbundle, raybundle, sequence, background_medium,
                      pilotraypath_nr=0, pilotbundle_generation="complex"):
        """Linearised trace of `raybundle` about the pilot ray (reference :381-469):
        per surface pair the deviation from the pilot ray is mapped by the XYUV
        matrix; no intersection and no aperture test takes place.  Returns
        (pilotraypath, RayPath of the linearised bundles)."""
        from .. import engine
        import torch
        dev = engine.compute_device()
        rpath = RayPath(raybundle)
        (pilotraypath, matrices) = self.calculateXYUV(
            pilotbundle, sequence, background_medium, pilotraypath_nr=pilotraypath_nr,
            pilotbundle_generation=pilotbundle_generation)
        (hitlist, _) = self.sequence_to_hitlist(sequence)
        for (ps, pe, surfhit) in zip(pilotraypath.raybundles[:-1],
                                     pilotraypath.raybundles[1:], hitlist):
            (surf_start_key, surf_end_key, _) = surfhit
            last = rpath.raybundles[-1]
            x0_glob = as_tensor(last.x[-1], dev)
            k0_glob = as_tensor(last.k[-1], dev)
            newbundle = RayBundle(x0_glob, k0_glob, None, last.rayID.to(dev),
                                  wave=last.wave)
            (x1, k1) = xyuv.para_step(
                self.surfaces[surf_start_key].rootcoordinatesystem,
                self.surfaces[surf_end_key].rootcoordinatesystem, matrices[surfhit],
                x0_glob, k0_glob, ps.x[-1][:, 0].cpu().numpy(), ps.k[-1][:, 0].cpu().numpy(),
                pe.x[-1][:, 0].cpu().numpy(), pe.k[-1][:, 0].cpu().numpy(),
                pilotbundle_generation)
            newbundle.append(x1, k1, newbundle.Efield[0],
                             torch.ones(x1.shape[1], dtype=torch.bool, device=dev))
            rpath.appendRayBundle(newbundle)
        return (pilotraypath, rpath)

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
