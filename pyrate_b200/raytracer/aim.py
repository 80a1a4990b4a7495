"""Paraxial ray aiming (host API mirror of reference raytracer/aim.py: Aimy :46-320).

`Aimy` traces a pilot bundle from the object surface (native engine, through
`OpticalSystem.extractXYUV`), keeps the object->stop and stop->image transfer
matrices and turns a field specification (angle or object height) into an initial
`RayBundle` whose rays fill the stop to first order.  Everything here is arithmetic on
4x4 / 6x6 matrices and <= num_pupil_points rays; the traces it triggers run on the
device.
"""
import numpy as np

from ..core import ClassWithOptimizableVariables
from ..sampling2d.raster import RectGrid
from .globalconstants import degree, standard_wavelength
from .helpers import (build_pilotbundle, build_pilotbundle_complex, rodrigues,
                      _perpendicular)
from .ray import RayBundle, returnDtoK


def _np(a):
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    return np.asarray(a)


class FieldManager(object):
    """Placeholder of the reference (:39-43)."""


class Aimy(ClassWithOptimizableVariables):

    def __init__(self, s, seq, wave=standard_wavelength, num_pupil_points=100,
                 stopsize=10, pilotbundle_solution=-1, pilotbundle_generation="complex",
                 pilotbundle_delta_angle=1 * degree, pilotbundle_delta_size=0.1,
                 pilotbundle_sampling_points=3, name=""):
        super(Aimy, self).__init__(name=name)
        self.field_raster = RectGrid()
        self.pupil_raster = RectGrid()
        self.stopsize = stopsize
        self.num_pupil_points = num_pupil_points
        self.wave = wave
        self.pilotbundle_solution = pilotbundle_solution
        self.pilotbundle_generation = pilotbundle_generation
        self.pilotbundle_delta_angle = pilotbundle_delta_angle
        self.pilotbundle_delta_size = pilotbundle_delta_size
        self.pilotbundle_sampling_points = pilotbundle_sampling_points
        self.update(s, seq)

    def setKind(self):
        self.kind = "aimy"

    def extract_abcd(self, xyuv):
        """2x2 blocks of a transfer matrix; only the real k rows / columns (:84-96)."""
        return (xyuv[0:2, 0:2], xyuv[0:2, 2:4], xyuv[2:4, 0:2], xyuv[2:4, 2:4])

    def update(self, system, seq):
        """Pilot bundle from the first surface of the sequence, then the object->stop
        and stop->image matrices (:98-153)."""
        obj_dx = self.pilotbundle_delta_size
        obj_dphi = self.pilotbundle_delta_angle
        (first_element_name, first_element_seq) = seq[0]
        (objsurfname, _) = first_element_seq[0]
        self.objectsurface = system.elements[first_element_name].surfaces[objsurfname]
        self.start_material = system.material_background
        build = {"real": build_pilotbundle,
                 "complex": build_pilotbundle_complex}[self.pilotbundle_generation.lower()]
        pilotbundles = build(self.objectsurface, self.start_material, (obj_dx, obj_dx),
                             (obj_dphi, obj_dphi),
                             num_sampling_points=self.pilotbundle_sampling_points)
        self.pilotbundle = pilotbundles[self.pilotbundle_solution]
        (self.m_obj_stop, self.m_stop_img) = system.extractXYUV(
            self.pilotbundle, seq, pilotbundle_generation=self.pilotbundle_generation)

    def _stop_raster(self):
        (xraster, yraster) = self.pupil_raster.getGrid(self.num_pupil_points)
        return np.vstack((xraster, yraster)) * self.stopsize

    def aim_core_angle_known(self, theta2d):
        """Start positions on the object surface for a field given as two angles
        (:155-190)."""
        (thetax, thetay) = theta2d
        rmfinal = np.dot(rodrigues(thetay, [1, 0, 0]), rodrigues(thetax, [0, 1, 0]))
        lc = self.objectsurface.rootcoordinatesystem
        dpilot_global = _np(self.pilotbundle.returnKtoD()[0, :, 0])
        kpilot_global = _np(self.pilotbundle.k[0, :, 0])
        dpilot_object = lc.returnGlobalToLocalDirections(dpilot_global.reshape(3, 1))
        kpilot_object = lc.returnGlobalToLocalDirections(kpilot_global.reshape(3, 1))
        dr_stop = self._stop_raster()
        kpilot_object = np.repeat(kpilot_object, dr_stop.shape[1], axis=1)
        dvec = np.dot(rmfinal, dpilot_object)
        dk_obj = (returnDtoK(dvec) - kpilot_object)[0:2, :]
        (a_obj_stop, b_obj_stop, _, _) = self.extract_abcd(self.m_obj_stop)
        dr_obj = np.dot(np.linalg.inv(a_obj_stop), dr_stop - np.dot(b_obj_stop, dk_obj))
        return (dr_obj, dk_obj)

    def aim_core_k_known(self, dk_obj):
        (a_obj_stop, b_obj_stop, _, _) = self.extract_abcd(self.m_obj_stop)
        dr_stop = self._stop_raster()
        dk_obj2 = np.repeat(np.asarray(dk_obj)[:, np.newaxis], dr_stop.shape[1], axis=1)
        dr_obj = np.dot(np.linalg.inv(a_obj_stop), dr_stop - np.dot(b_obj_stop, dk_obj2))
        return (dr_obj, dk_obj2)

    def aim_core_r_known(self, delta_xy):
        """Start directions for a field given as an object height (:213-239)."""
        (a_obj_stop, b_obj_stop, _, _) = self.extract_abcd(self.m_obj_stop)
        dr_stop = self._stop_raster()
        dr_obj = np.repeat(np.asarray(delta_xy)[:, np.newaxis], dr_stop.shape[1], axis=1)
        dk_obj = np.dot(np.linalg.inv(b_obj_stop), dr_stop - np.dot(a_obj_stop, dr_obj))
        return (dr_obj, dk_obj)

    def aim(self, delta_xy, fieldtype="angle"):
        """Initial RayBundle for a field point (:241-318).  As in the reference the
        result is linearised (k is the pilot k plus a tangential increment and
        violates |k| = n at second order), and E is some unit vector with E.k = 0 (the
        reference takes an arbitrary SVD null vector there, :301-312)."""
        if fieldtype == "angle":
            (dr_obj, dk_obj) = self.aim_core_angle_known(delta_xy)
        elif fieldtype == "objectheight":
            (dr_obj, dk_obj) = self.aim_core_r_known(delta_xy)
        else:
            raise NotImplementedError()
        num_points = dr_obj.shape[1]
        dr_obj3d = np.vstack((dr_obj, np.zeros(num_points)))
        dk_obj3d = np.vstack((dk_obj, np.zeros(num_points)))
        lc = self.objectsurface.rootcoordinatesystem
        basis = np.asarray(lc.localbasis)
        # the reference adds global-frame increments to object-frame pilot values
        # (:262-275, transposed basis); kept as is -- identical for the usual object
        # surface without tilt
        xp_objsurf = lc.returnGlobalToLocalPoints(
            _np(self.pilotbundle.x[0, :, 0]).reshape(3, 1))
        xparabasal = np.repeat(xp_objsurf, num_points, axis=1) + \
            np.dot(basis.T, dr_obj3d)
        kp_objsurf = lc.returnGlobalToLocalDirections(
            _np(self.pilotbundle.k[0, :, 0]).reshape(3, 1))
        kparabasal = np.repeat(kp_objsurf, num_points, axis=1) + \
            np.dot(basis.T, dk_obj3d)
        efield = _perpendicular(kparabasal)
        if not (np.iscomplexobj(kparabasal) and np.any(kparabasal.imag)):
            (kparabasal, efield) = (np.real(kparabasal), np.real(efield))
        return RayBundle(np.real(xparabasal), kparabasal, efield, wave=self.wave)
