"""Tree of decentered / tilted frames (host side).

API mirror of reference raytracer/localcoordinates.py: `LocalCoordinates.p`
:44-105, tilt order `calculateMatrixFromTilt` :171-180, composition `update`
:264-307, point / direction / tensor maps :383-435, `Actual<->Other` :354-381.
Frames are tiny (3x3 + 3) and are rebuilt on the host before every trace; the
device only sees the finished matrices in the step table (lowering.py).  The
transform helpers accept NumPy arrays or torch tensors (any device).
"""
import math

import numpy as np

from ..core import ClassWithOptimizableVariables, FloatOptimizableVariable

try:
    import torch
except Exception:          # pragma: no cover
    torch = None


def _axis_rotations(tiltx, tilty, tiltz):
    (cx, sx) = (math.cos(tiltx), math.sin(tiltx))
    (cy, sy) = (math.cos(tilty), math.sin(tilty))
    (cz, sz) = (math.cos(tiltz), math.sin(tiltz))
    rx = np.array([[1., 0., 0.], [0., cx, -sx], [0., sx, cx]])
    ry = np.array([[cy, 0., sy], [0., 1., 0.], [-sy, 0., cy]])
    rz = np.array([[cz, -sz, 0.], [sz, cz, 0.], [0., 0., 1.]])
    return (rx, ry, rz)


def _is_torch(a):
    return torch is not None and isinstance(a, torch.Tensor)


def _mat_for(mat, like):
    if _is_torch(like):
        m = torch.as_tensor(mat, dtype=torch.float64, device=like.device)
        return m.to(like.dtype) if like.is_complex() else m
    return mat


class LocalCoordinates(ClassWithOptimizableVariables):

    @classmethod
    def p(cls, name="", **kwargs):
        structure = {}
        for key in ("decx", "decy", "decz", "tiltx", "tilty", "tiltz"):
            structure[key] = FloatOptimizableVariable(kwargs.get(key, 0.0),
                                                      name=key)
        structure["parent"] = None
        annotations = {"tiltThenDecenter": kwargs.get("tiltThenDecenter", 0)}
        lc = cls(annotations, structure, name)
        lc.update()
        return lc

    def setKind(self):
        self.kind = "localcoordinates"

    def initialize_from_annotations(self):
        self._children = []
        self.globalcoordinates = np.zeros(3)
        self.localdecenter = np.zeros(3)
        self.localrotation = np.eye(3)
        self.localbasis = np.eye(3)

    @property
    def children(self):
        return self._children

    @property
    def tiltThenDecenter(self):
        return self.annotations["tiltThenDecenter"]

    # ---- tree ----
    def addChild(self, childlc):
        childlc.parent = self
        childlc.update()
        self._children.append(childlc)
        return childlc

    def addChildToReference(self, refname, childlc):
        if self.name == refname:
            self.addChild(childlc)
        else:
            for ch in self._children:
                ch.addChildToReference(refname, childlc)
        return childlc

    def returnConnectedNames(self):
        names = [self.name]
        for ch in self._children:
            names += ch.returnConnectedNames()
        return names

    def returnConnectedChildren(self):
        lst = [self]
        for ch in self._children:
            lst += ch.returnConnectedChildren()
        return lst

    # ---- frame maths ----
    def calculateMatrixFromTilt(self, tiltx, tilty, tiltz, tiltThenDecenter=0):
        (rx, ry, rz) = _axis_rotations(tiltx, tilty, tiltz)
        if tiltThenDecenter == 0:
            return rz @ (ry @ rx)
        return rx @ (ry @ rz)

    def calculate(self):
        self.localdecenter = np.array([self.decx(), self.decy(), self.decz()])
        self.localrotation = self.calculateMatrixFromTilt(
            self.tiltx(), self.tilty(), self.tiltz(),
            self.annotations["tiltThenDecenter"])

    def update(self):
        self.calculate()
        if self.parent is not None:
            (pcoord, pbasis) = (self.parent.globalcoordinates,
                                self.parent.localbasis)
        else:
            (pcoord, pbasis) = (np.zeros(3), np.eye(3))
        self.localbasis = pbasis @ self.localrotation
        if self.annotations["tiltThenDecenter"] == 0:
            self.globalcoordinates = pcoord + pbasis @ self.localdecenter
        else:
            self.globalcoordinates = pcoord + self.localbasis @ self.localdecenter
        for ch in self._children:
            ch.update()

    # ---- maps (NumPy or torch, (3, N)) ----
    def returnLocalToGlobalPoints(self, localpts):
        b = _mat_for(self.localbasis, localpts)
        o = _mat_for(self.globalcoordinates, localpts)
        return b @ localpts + o[:, None]

    def returnLocalToGlobalDirections(self, localdirs):
        return _mat_for(self.localbasis, localdirs) @ localdirs

    def returnGlobalToLocalPoints(self, globalpts):
        b = _mat_for(self.localbasis, globalpts)
        o = _mat_for(self.globalcoordinates, globalpts)
        return b.T @ (globalpts - o[:, None])

    def returnGlobalToLocalDirections(self, globaldirs):
        return _mat_for(self.localbasis, globaldirs).T @ globaldirs

    def returnGlobalToLocalTensors(self, globaltensor):
        b = _mat_for(self.localbasis, globaltensor)
        ein = torch.einsum if _is_torch(globaltensor) else np.einsum
        return ein("ji,jkn,kl->iln", b, globaltensor, b)

    def returnLocalToGlobalTensors(self, localtensor):
        b = _mat_for(self.localbasis, localtensor)
        ein = torch.einsum if _is_torch(localtensor) else np.einsum
        return ein("ij,jkn,lk->iln", b, localtensor, b)

    def returnActualToOtherPoints(self, localpts, lcother):
        return lcother.returnGlobalToLocalPoints(
            self.returnLocalToGlobalPoints(localpts))

    def returnOtherToActualPoints(self, otherpts, lcother):
        return self.returnGlobalToLocalPoints(
            lcother.returnLocalToGlobalPoints(otherpts))

    def returnActualToOtherDirections(self, localdirs, lcother):
        return lcother.returnGlobalToLocalDirections(
            self.returnLocalToGlobalDirections(localdirs))

    def returnOtherToActualDirections(self, otherdirs, lcother):
        return self.returnGlobalToLocalDirections(
            lcother.returnLocalToGlobalDirections(otherdirs))

    def returnActualToOtherTensors(self, localtensors, lcother):
        return lcother.returnGlobalToLocalTensors(
            self.returnLocalToGlobalTensors(localtensors))

    def returnOtherToActualTensors(self, othertensors, lcother):
        return self.returnGlobalToLocalTensors(
            lcother.returnLocalToGlobalTensors(othertensors))

    def pprint(self, n=0):
        s = n * "    " + self.name + " (" + str(self.globalcoordinates) + ")\n"
        for ch in self._children:
            s += ch.pprint(n + 1)
        return s


class LocalCoordinatesTreeBase(ClassWithOptimizableVariables):
    """Connection checks shared by OpticalSystem / OpticalElement / Surface
    (reference raytracer/localcoordinatestreebase.py:31-88)."""

    @classmethod
    def p(cls, rootcoordinatesystem, name=""):
        return cls({}, {"rootcoordinatesystem": rootcoordinatesystem},
                   name=name)

    def checkForRootConnection(self, lc):
        return any(lc is c for c in
                   self.rootcoordinatesystem.returnConnectedChildren())

    def addLocalCoordinateSystem(self, lc, refname):
        allnames = self.rootcoordinatesystem.returnConnectedNames()
        if lc.name in allnames:
            lc.name = lc.name + "_" + str(len(allnames))
        if refname not in allnames:
            refname = self.rootcoordinatesystem.name
        self.rootcoordinatesystem.addChildToReference(refname, lc)
        self.rootcoordinatesystem.update()
        return lc
