"""Surface = shape + aperture + root frame (host API mirror of reference
raytracer/surface.py:35-135).  `intersect` runs on the device."""
from .aperture import BaseAperture, create_aperture
from .localcoordinates import LocalCoordinatesTreeBase
from .surface_shape import Conic


class Surface(LocalCoordinatesTreeBase):

    @classmethod
    def p(cls, rootlc, shape=None, aperture=None, name=""):
        if shape is None:
            shape = Conic.p(rootlc)
        inst = cls({}, {"rootcoordinatesystem": rootlc}, name=name)
        aperture_ = BaseAperture.p(rootlc)
        if isinstance(aperture, BaseAperture):
            aperture_ = aperture
        elif isinstance(aperture, dict):
            aperture_ = create_aperture(rootlc, aperture)
        inst.setShape(shape)
        inst.setAperture(aperture_)
        return inst

    def setKind(self):
        self.kind = "surface"

    def setAperture(self, apert):
        if not self.checkForRootConnection(apert.lc):
            raise Exception("Aperture coordinate system should be connected "
                            "to surface coordinate system")
        self._aperture = apert

    def getAperture(self):
        return self._aperture

    aperture = property(getAperture, setAperture)

    def setShape(self, shape):
        if not self.checkForRootConnection(shape.lc):
            raise Exception("Shape coordinate system should be connected "
                            "to surface coordinate system")
        self._shape = shape

    def getShape(self):
        return self._shape

    shape = property(getShape, setShape)

    def intersect(self, raybundle, remove_rays_outside_aperture=True):
        """Shape intersect + aperture mask, appended to `raybundle` (device)."""
        from .. import engine
        engine.surface_intersect(self, raybundle, remove_rays_outside_aperture)

    def getCentralCurvature(self, ray=None):
        """Vertex curvature of the shape (reference :253-257)."""
        return self.shape.getCentralCurvature()
