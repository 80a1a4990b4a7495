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


def poisson_disk_2d(width, height, radius, tries=30, rng=None):
    """Poisson-disk ("blue noise") samples of [0, width) x [0, height) with pairwise
    distance >= radius: Bridson's grid-accelerated dart throwing (the reference ships a
    slower grid-based variant of the same hard-core process, sampling2d/pds.py).
    Returns an (N, 2) array."""
    rng = np.random.default_rng() if rng is None else rng
    cell = radius / math.sqrt(2.0)
    (gw, gh) = (int(math.ceil(width / cell)), int(math.ceil(height / cell)))
    grid = -np.ones((gw, gh), dtype=np.int64)
    pts = [np.array([rng.uniform(0, width), rng.uniform(0, height)])]
    grid[int(pts[0][0] / cell), int(pts[0][1] / cell)] = 0
    active = [0]
    while active:
        pick = int(rng.integers(len(active)))
        base = pts[active[pick]]
        for _ in range(tries):
            (rad, ang) = (radius * math.sqrt(rng.uniform(1.0, 4.0)), rng.uniform(0, 2 * math.pi))
            cand = base + rad * np.array([math.cos(ang), math.sin(ang)])
            if not (0 <= cand[0] < width and 0 <= cand[1] < height):
                continue
            (ci, cj) = (int(cand[0] / cell), int(cand[1] / cell))
            near = grid[max(ci - 2, 0):ci + 3, max(cj - 2, 0):cj + 3].ravel()
            near = near[near >= 0]
            if all(np.sum((pts[q] - cand) ** 2) >= radius * radius for q in near):
                grid[ci, cj] = len(pts)
                active.append(len(pts))
                pts.append(cand)
                break
        else:
            active.pop(pick)
    return np.array(pts)


class PoissonDiskSampling(RectGrid):
    """Random pupil points with a minimum mutual distance (reference raster.py:106-125):
    about `nray` points in the unit disk."""

    def getGrid(self, nray, rng=None):
        per_dim = max(1, int(round(math.sqrt(nray * 4.0 / math.pi))))
        sample = poisson_disk_2d(2.0, 2.0, 1.0 / per_dim, rng=rng)
        (xs, ys) = (sample[:, 0] - 1.0, sample[:, 1] - 1.0)
        keep = xs ** 2 + ys ** 2 <= 1
        return (xs[keep], ys[keep])


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
