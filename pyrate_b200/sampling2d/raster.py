"""Pupil / field rasters (host API mirror of reference sampling2d/raster.py:36-166)
plus the hexapolar raster BASELINE.json's bundles use.  A raster returns normalised
coordinates on the unit disk; bundle generation scales them (host-side, O(N) NumPy,
once per bundle -- the rays themselves are traced on the device)."""
import math

import numpy as np


class RectGrid(object):
    """Square lattice clipped to the unit disk (reference :36-60)."""

    def getGrid(self, nray):
        per_dim = int(round(math.sqrt(nray * 4.0 / math.pi)))
        dx = 1. / per_dim
        x1d = np.linspace(-1 + .25 * dx, 1 - .25 * dx, per_dim)
        (xp, yp) = np.meshgrid(x1d, x1d)
        (xp, yp) = (xp.reshape(-1), yp.reshape(-1))
        keep = xp ** 2 + yp ** 2 <= 1
        return (xp[keep], yp[keep])


class HexGrid(RectGrid):
    """Hexagonal Bravais lattice (two interleaved rectangular ones), :62-91."""

    def getGrid(self, nray):
        nx = int(round(math.sqrt(2 * math.sqrt(3) * nray / math.pi) + 1))
        x1d = np.linspace(-1, 1, nx)
        y1d = x1d * math.sqrt(3)
        (dx, dy) = (x1d[1] - x1d[0], y1d[1] - y1d[0])
        (xa, ya) = np.meshgrid(x1d, y1d)
        (xa, ya) = (xa.reshape(-1), ya.reshape(-1))
        (xb, yb) = (xa + 0.5 * dx, ya + 0.5 * dy)
        ka = xa ** 2 + ya ** 2 <= 1
        kb = xb ** 2 + yb ** 2 <= 1
        return (np.hstack((xa[ka], xb[kb])), np.hstack((ya[ka], yb[kb])))


class RandomGrid(RectGrid):
    def getGrid(self, nray):
        m = int(round(nray * 4.0 / math.pi))
        xp = 2. * np.random.random(m) - 1.
        yp = 2. * np.random.random(m) - 1.
        keep = xp ** 2 + yp ** 2 <= 1.
        return (xp[keep], yp[keep])


class MeridionalFan(RectGrid):
    def getGrid(self, nray, phi=0.):
        lin = np.linspace(-1, 1, nray)
        alpha = phi / 180. * math.pi
        return (lin * -math.sin(alpha), lin * math.cos(alpha))


class SagitalFan(RectGrid):
    def getGrid(self, nray, phi=0.):
        return MeridionalFan().getGrid(nray, phi - 90.)


class ChiefAndComa(RectGrid):
    def getGrid(self, nray, phi=0.):
        a = phi / 180. * math.pi
        (s, c) = (math.sin(a), math.cos(a))
        return (np.array([0, 0, -s, s, c, -c], dtype=float),
                np.array([0, 0, c, -c, s, -s], dtype=float))


class Single(RectGrid):
    def getGrid(self, nray, xpup=0.0, ypup=0.0):
        return (np.array([xpup]), np.array([ypup]))


class CircularGrid(RectGrid):
    """Polar grid with the same number of points on every ring (:150-166)."""

    def getGrid(self, nray, requidistant=True):
        m = int(round(math.sqrt(nray)))
        r = np.linspace(0, 1, num=m)
        if not requidistant:
            r = np.sqrt(r)
        phi = np.linspace(0, 2. * math.pi, num=m, endpoint=False)
        (rr, pp) = np.meshgrid(r, phi)
        return ((rr * np.cos(pp)).flatten(), (rr * np.sin(pp)).flatten())


class HexapolarGrid(RectGrid):
    """Ring j carries 6 j points (not in the reference): the smallest
    ring count with at least `nray` points is used."""

    def getGrid(self, nray):
        from ..configs import hexapolar, rings_for
        return hexapolar(rings_for(nray))
