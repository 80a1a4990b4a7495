"""Bundle generators: the description of a ray bundle the device expands in registers.

Host mirror of what sits right above `seqtrace` in the reference --
`OpticalSystemAnalysis.collimated_bundle` / `divergent_bundle`
(raytracer/analysis/optical_system_analysis.py:83-165) over the rasters of
sampling2d/raster.py:36-166 -- reduced to what a bundle in a homogeneous isotropic
background really is: (raster, radius, start point, direction, index).  A
`BundleGen` is <= 176 bytes on the wire (`PyrBundleGen`, include/pyrate_b200.h) plus,
for the two clipped lattices, a per-row table of a few KB; ray i of a call is raster
point `first + i`, computed by the trace kernel's prologue (csrc/pyr_gen.cuh) or
written out by `pyr_generate_bundle`.  Nothing here traces or generates rays on the
host: `points_host()` exists for the CPU tests that pin the index arithmetic on the
NumPy rasters.
"""
import ctypes as C
import math

import numpy as np

from . import _native as nat


def raster_rows(xa, ya, radius2=1.0):
    """Row table of a lattice clipped to the unit disk.

    Lattice point (ix, iy) = (xa[ix], ya[iy]) is kept iff xa[ix]**2 + ya[iy]**2 <= 1,
    evaluated in floating point exactly like the NumPy rasters do (square, square, add,
    compare).  xa is ascending, so the kept ix of a row form one interval; returns
    (prefix[nrows + 1], first[nrows]): number of kept points ahead of row iy, and the
    first kept ix of the row.  O(n log n) instead of the O(n^2) mask."""
    xa = np.asarray(xa, dtype=np.float64)
    ya = np.asarray(ya, dtype=np.float64)
    n = xa.size
    xx = xa ** 2
    yy = ya ** 2

    def pred(ix, rows):
        ok = (ix >= 0) & (ix < n)
        out = np.zeros(ix.shape, dtype=bool)
        out[ok] = xx[ix[ok]] + yy[rows[ok]] <= radius2
        return out
    rows = np.arange(ya.size)
    with np.errstate(invalid="ignore"):
        lim = np.sqrt(np.maximum(radius2 - yy, 0.0))
    a = np.searchsorted(xa, -lim, side="left").astype(np.int64)
    b = np.searchsorted(xa, lim, side="right").astype(np.int64) - 1
    centre = int(np.argmin(np.abs(xa)))
    empty = b < a
    a[empty] = centre
    b[empty] = centre - 1
    # exact fix-up at the rim (rounding of the squares): grow, then shrink
    for _ in range(8):
        grow_a = pred(a - 1, rows)
        grow_b = pred(b + 1, rows)
        nonempty = b >= a
        shrink_a = nonempty & ~pred(a, rows)
        shrink_b = nonempty & ~pred(b, rows)
        if not (grow_a.any() or grow_b.any() or shrink_a.any() or shrink_b.any()):
            break
        a = a - grow_a + (shrink_a & ~grow_a)
        b = b + grow_b - (shrink_b & ~grow_b)
    count = np.maximum(b - a + 1, 0)
    prefix = np.concatenate(([0], np.cumsum(count))).astype(np.int64)
    return prefix, a.astype(np.int64)


def _linspace_params(start, stop, num):
    """(start, step, stop) exactly as numpy.linspace computes its points:
    y = arange(num) * step + start, y[-1] = stop."""
    start = np.float64(start)
    stop = np.float64(stop)
    div = num - 1
    step = (stop - start) / div if div > 0 else np.float64(0.0)
    return float(start), float(step), float(stop)


class RasterSpec(object):
    """Device description of one raster instance: kind, param, total and the tables."""

    def __init__(self, kind, param, total, lin=(0.0, 0.0, 0.0), aux=(0.0, 0.0), rows=None,
                 flags=0):
        (self.kind, self.param, self.total) = (int(kind), int(param), int(total))
        (self.lin, self.aux, self.rows, self.flags) = (tuple(lin), tuple(aux), rows, int(flags))


def hexapolar_spec(rings):
    rings = int(rings)
    return RasterSpec(nat.RASTER_HEXAPOLAR, rings, 1 + 3 * rings * (rings + 1))


def rect_spec(nray):
    """RectGrid.getGrid(nray), sampling2d/raster.py:40-60."""
    per_dim = int(round(math.sqrt(nray * 4.0 / math.pi)))
    dx = 1. / per_dim
    lin = _linspace_params(-1 + .25 * dx, 1 - .25 * dx, per_dim)
    x1d = np.linspace(-1 + .25 * dx, 1 - .25 * dx, per_dim)
    (prefix, first) = raster_rows(x1d, x1d)
    return RasterSpec(nat.RASTER_RECT, per_dim, int(prefix[-1]), lin,
                      rows=np.concatenate((prefix, first)))


def hex_spec(nray):
    """HexGrid.getGrid(nray), sampling2d/raster.py:63-91: lattice 1, then lattice 2."""
    nx = int(round(math.sqrt(2 * math.sqrt(3) * nray / math.pi) + 1))
    x1d = np.linspace(-1, 1, nx)
    y1d = x1d * math.sqrt(3)
    (dx, dy) = (x1d[1] - x1d[0], y1d[1] - y1d[0])
    (pa, fa) = raster_rows(x1d, y1d)
    (pb, fb) = raster_rows(x1d + 0.5 * dx, y1d + 0.5 * dy)
    prefix = np.concatenate((pa, pa[-1] + pb[1:]))
    first = np.concatenate((fa, fb))
    return RasterSpec(nat.RASTER_HEX, nx, int(prefix[-1]), _linspace_params(-1, 1, nx),
                      aux=(float(0.5 * dx), float(0.5 * dy)),
                      rows=np.concatenate((prefix, first)))


def circular_spec(nray, requidistant=True):
    """CircularGrid.getGrid(nray, requidistant), sampling2d/raster.py:151-166."""
    m = int(round(math.sqrt(nray)))
    (_, rstep, _) = _linspace_params(0, 1, m)
    pstep = float((np.float64(2. * math.pi) - 0.0) / m)      # endpoint=False: delta / num
    return RasterSpec(nat.RASTER_CIRCULAR, m, m * m, lin=(0.0, rstep, 1.0), aux=(pstep, 0.0),
                      flags=0 if requidistant else nat.GEN_SQRT_R)


def spec_for_raster(raster, nray):
    """RasterSpec of a raster object of pyrate_b200.sampling2d.raster, or None when the
    device has no generator for it (random rasters, fans, single points: tiny or
    irregular -- those bundles are built on the host and uploaded)."""
    from .sampling2d import raster as R
    t = type(raster)
    if t is R.RectGrid:
        return rect_spec(nray)
    if t is R.HexGrid:
        return hex_spec(nray)
    if t is R.CircularGrid:
        return circular_spec(nray)
    if t is R.HexapolarGrid:
        from .configs import rings_for
        return hexapolar_spec(rings_for(nray))
    return None


class BundleGen(object):
    """A generated bundle (or a shard [first, first + n) of one)."""

    def __init__(self, raster, bundle=nat.BUNDLE_COLLIMATED, radius=1.0, start=(0., 0., 0.),
                 direction=(0., 0., 1.), efield=None, n_index=1.0, first=0, count=None):
        self.raster = raster
        self.bundle = int(bundle)
        self.radius = float(radius)
        self.start = tuple(float(v) for v in start)
        self.direction = tuple(float(v) for v in direction)
        self.efield = None if efield is None else tuple(float(v) for v in efield)
        self.n_index = float(n_index)
        self.first = int(first)
        self.n = int(raster.total - self.first if count is None else count)
        assert 0 <= self.first and self.first + self.n <= raster.total
        self._rows_dev = {}
        self._cache = None          # materialised (x, k, e) device tensors

    def shard(self, lo, hi):
        """Rays [lo, hi) of this bundle (ray sharding across ranks / chunks)."""
        g = BundleGen(self.raster, self.bundle, self.radius, self.start, self.direction,
                      self.efield, self.n_index, self.first + int(lo), int(hi) - int(lo))
        g._rows_dev = self._rows_dev
        return g

    def descriptor(self, device):
        """PyrBundleGen for `device` (uploads the row table once per device)."""
        import torch
        g = nat.PyrBundleGen()
        r = self.raster
        (g.raster, g.bundle, g.param, g.first, g.total) = (r.kind, self.bundle, r.param,
                                                           self.first, r.total)
        g.flags = r.flags | (nat.GEN_E_PERP if self.efield is None else 0)
        (g.lin_start, g.lin_step, g.lin_stop) = r.lin
        g.aux[:] = r.aux
        g.radius = self.radius
        g.start[:] = self.start
        g.dir[:] = self.direction
        g.e[:] = self.efield if self.efield is not None else (0.0, 0.0, 0.0)
        g.n_index = self.n_index
        if r.rows is not None:
            key = str(device)
            t = self._rows_dev.get(key)
            if t is None:
                t = torch.from_numpy(np.ascontiguousarray(r.rows, dtype=np.int64)).to(device)
                self._rows_dev[key] = t
            g.rows = t.data_ptr()
        return g

    def materialise(self, device=None):
        """(x0, k0, E0) as (3, n) CUDA tensors in the engine's row-padded layout: one
        pyr_generate_bundle launch, cached."""
        import torch
        from . import engine
        lib = engine.require_cuda()
        device = torch.device("cuda", torch.cuda.current_device()) if device is None \
            else torch.device(device)
        if self._cache is not None and self._cache[0].device == device:
            return self._cache
        ld = engine._round_up(max(self.n, 1), engine.LD_ALIGN)
        buf = torch.zeros((3, 3, ld), dtype=torch.float64, device=device)
        desc = self.descriptor(device)
        stream = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        with torch.cuda.device(device):
            nat.check(lib.pyr_generate_bundle(C.byref(desc), self.n, buf[0].data_ptr(),
                                              buf[1].data_ptr(), buf[2].data_ptr(), ld, stream))
        self._cache = (buf[0, :, :self.n], buf[1, :, :self.n], buf[2, :, :self.n])
        return self._cache

    @property
    def materialised(self):
        return self._cache is not None

    # ---- CPU restatement of the index arithmetic (tests only) ----
    def points_host(self, index=None):
        """Normalised raster coordinates of this shard (or of its rays `index`), by the
        same index arithmetic the device uses (csrc/pyr_gen.cuh: gen_point)."""
        r = self.raster
        if index is None:
            idx = np.arange(self.first, self.first + self.n, dtype=np.int64)
        else:
            idx = self.first + np.asarray(index, dtype=np.int64)
        (ls, lstep, lstop) = r.lin

        def lin(i):
            v = i.astype(np.float64) * lstep + ls
            return np.where((i == r.param - 1) & (r.param > 1), lstop, v)
        if r.kind == nat.RASTER_HEXAPOLAR:
            j = np.floor((3.0 + np.sqrt(9.0 + 12.0 * np.maximum(idx - 1, 0))) / 6.0).astype(np.int64)
            j = np.maximum(j, 1)
            j = np.where(idx < 1 + 3 * (j - 1) * j, j - 1, j)
            j = np.where(idx >= 1 + 3 * j * (j + 1), j + 1, j)
            i = idx - (1 + 3 * (j - 1) * j)
            jj = np.maximum(j, 1)
            ang = 2.0 * math.pi * i / (6.0 * jj)
            rad = jj / float(max(r.param, 1))
            return (np.where(idx <= 0, 0.0, rad * np.cos(ang)),
                    np.where(idx <= 0, 0.0, rad * np.sin(ang)))
        if r.kind in (nat.RASTER_RECT, nat.RASTER_HEX):
            nrows = r.param if r.kind == nat.RASTER_RECT else 2 * r.param
            prefix = r.rows[:nrows + 1]
            first = r.rows[nrows + 1:]
            row = np.searchsorted(prefix, idx, side="right") - 1
            ix = first[row] + (idx - prefix[row])
            if r.kind == nat.RASTER_RECT:
                return lin(ix), lin(row)
            second = row >= r.param
            px = lin(ix)
            py = lin(np.where(second, row - r.param, row)) * math.sqrt(3)
            return (np.where(second, px + r.aux[0], px), np.where(second, py + r.aux[1], py))
        m = max(r.param, 1)
        (iphi, ir) = (idx // m, idx % m)
        rr = np.where((ir == m - 1) & (m > 1), 1.0, ir.astype(np.float64) * lstep)
        if r.flags & nat.GEN_SQRT_R:
            rr = np.sqrt(rr)
        phi = iphi.astype(np.float64) * r.aux[0]
        return rr * np.cos(phi), rr * np.sin(phi)

    def arrays_host(self, index=None):
        """(x0, k0, E0) NumPy arrays by the CPU restatement (tests / parity samples only)."""
        (px, py) = self.points_host(index)
        n = px.size
        x = np.empty((3, n))
        d = np.empty((3, n))
        if self.bundle == nat.BUNDLE_COLLIMATED:
            x[0] = self.radius * px + self.start[0]
            x[1] = self.radius * py + self.start[1]
            x[2] = self.start[2]
            d[:] = np.asarray(self.direction)[:, None]
        else:
            x[:] = np.asarray(self.start)[:, None]
            (ay, ax) = (self.direction[0] + self.radius * px, self.direction[1] + self.radius * py)
            d[0] = np.sin(ay) * np.cos(ax)
            d[1] = np.sin(ax)
            d[2] = np.cos(ay) * np.cos(ax)
        k = self.n_index * d
        if self.efield is None:
            axis = np.zeros_like(d)
            axis[np.argmin(np.abs(d), axis=0), np.arange(n)] = 1.0
            e = axis - np.sum(axis * d, axis=0) / np.sum(d * d, axis=0) * d
            e = e / np.sqrt(np.sum(e * e, axis=0))
        else:
            e = np.repeat(np.asarray(self.efield, dtype=float)[:, None], n, axis=1)
        return x, k, e


def hexapolar_collimated(rings, radius, z0, kdir=(0.0, 0.0, 1.0), efield=(0.0, 1.0, 0.0),
                         n_index=1.0):
    """The generator of configs.collimated_bundle(rings, radius, z0, kdir, efield)."""
    return BundleGen(hexapolar_spec(rings), nat.BUNDLE_COLLIMATED, radius, (0.0, 0.0, z0),
                     kdir, efield, n_index)


def config_generator(spec, rings=None, kdir=(0.0, 0.0, 1.0), efield=(0.0, 1.0, 0.0)):
    """Generator of configs.config_bundle(spec, rings, kdir, efield)."""
    b = spec["bundle"]
    return hexapolar_collimated(b["rings"] if rings is None else rings, b["radius"], b["z0"],
                                kdir, efield)
