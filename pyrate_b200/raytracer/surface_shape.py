"""Surface shapes (host API mirror of reference raytracer/surface_shape.py).

`Conic` :158-325, `Asphere` :520-606, `XYPolynomials` :780-858 with the same
`.p(...)` signatures, parameter containers (`curvature` / `conic` variables,
`params` dict with keys "curv", "cc", "A2", ..., "CX{m}Y{n}", "normradius") and
annotations ("tol", "iterations", "numcoefficients").  `getSag` / `getGrad` /
`getNormal` / `getHessian` evaluate on NumPy arrays or torch tensors (plotting and
analysis helpers).  `intersect(raybundle)` is the device path: one
PYR_STEP_PROPAGATE_ONLY launch of the native engine (no host arithmetic on rays).
"""
import numpy as np

from ..core import ClassWithOptimizableVariables, FloatOptimizableVariable

try:
    import torch
except Exception:          # pragma: no cover
    torch = None


def _lib(a):
    return torch if (torch is not None and isinstance(a, torch.Tensor)) else np


class Shape(ClassWithOptimizableVariables):

    def setKind(self):
        self.kind = "shape"

    def getSag(self, x, y):
        raise NotImplementedError()

    def getGrad(self, x, y):
        raise NotImplementedError()

    def getHessian(self, x, y):
        raise NotImplementedError()

    def getCentralCurvature(self):
        raise NotImplementedError()

    def getNormal(self, x, y):
        xp = _lib(x)
        g = self.getGrad(x, y)
        return g / xp.sqrt((g * g).sum(0))

    def intersect(self, raybundle):
        """Appends the intersection row to `raybundle` (device)."""
        from .. import engine
        engine.shape_intersect(self, raybundle)


class Conic(Shape):

    @classmethod
    def p(cls, lc, curv=0.0, cc=0.0, name=""):
        return cls({}, {"curvature": FloatOptimizableVariable(curv, name="curvature"),
                        "conic": FloatOptimizableVariable(cc, name="conic constant"),
                        "lc": lc}, name)

    def setKind(self):
        self.kind = "shape_Conic"

    def getCentralCurvature(self):
        return self.curvature()

    def conic_function(self, rsquared):
        xp = _lib(rsquared)
        (curv, cc) = (self.curvature(), self.conic())
        s = 1 - (1 + cc) * curv ** 2 * rsquared
        bad = s <= 0
        nan = float("nan")
        r2 = xp.where(bad, xp.full_like(rsquared, nan), rsquared)
        s = xp.where(bad, xp.zeros_like(s), s)
        return curv * r2 / (1 + xp.sqrt(s))

    def getSag(self, x, y):
        return self.conic_function(x * x + y * y)

    def getGrad(self, x, y):
        xp = _lib(x)
        (curv, cc) = (self.curvature(), self.conic())
        z = self.getSag(x, y)
        return xp.stack((-curv * x, -curv * y, 1. - curv * z * (1 + cc)))

    def getHessian(self, x, y):
        xp = _lib(x)
        (curv, cc) = (self.curvature(), self.conic())
        h = xp.zeros((3, 3) + tuple(x.shape), dtype=x.dtype)
        h[0, 0] = curv
        h[1, 1] = curv
        h[2, 2] = curv * (1 + cc)
        return h


class Cylinder(Conic):
    """Conic section in y, extruded along x (reference surface_shape.py:328-388).  Sag and
    `.p` as in the reference; gradient / Hessian of the actual surface (the reference
    inherits Conic's rotationally symmetric ones, inconsistent with its own sag)."""

    def setKind(self):
        self.kind = "shape_Cylinder"

    def getSag(self, x, y):
        return self.conic_function(y * y + 0 * x)

    def getGrad(self, x, y):
        xp = _lib(x)
        (curv, cc) = (self.curvature(), self.conic())
        z = self.getSag(x, y)
        return xp.stack((0 * x, -curv * y, 1. - curv * z * (1 + cc)))

    def getHessian(self, x, y):
        h = super(Cylinder, self).getHessian(x, y)
        h[0, 0] = 0
        return h


class FreeShape(Shape):

    @staticmethod
    def createAnnotationsAndStructure(lc, paramlist=(), tol=1e-6, iterations=10):
        params = {}
        for (name, value) in paramlist:
            params[name] = FloatOptimizableVariable(value, name=name)
        return ({"tol": tol, "iterations": iterations},
                {"lc": lc, "params": params})

    def getGrad(self, x, y):
        return self.gradF(x, y, self.getSag(x, y))

    def getHessian(self, x, y):
        return self.hessF(x, y, self.getSag(x, y))


class ExplicitShape(FreeShape):
    """z = F(x, y); the reference solves the ray equation with fsolve
    (:448-465); here the native engine runs a per-ray Newton iteration."""

    def getSag(self, x, y):
        return self.F(x, y)


class Asphere(ExplicitShape):

    @classmethod
    def p(cls, lc, curv=0, cc=0, coefficients=None, name=""):
        coefficients = [] if coefficients is None else list(coefficients)
        plist = [("curv", curv), ("cc", cc)] + \
                [("A" + str(2 * i + 2), v) for (i, v) in enumerate(coefficients)]
        (ann, struct) = FreeShape.createAnnotationsAndStructure(lc, plist)
        ann["numcoefficients"] = len(coefficients)
        return cls(ann, struct, name)

    def setKind(self):
        self.kind = "shape_Asphere"

    def getAsphereParameters(self):
        return (self.params["curv"](), self.params["cc"](),
                [self.params["A" + str(2 * i + 2)]()
                 for i in range(self.annotations["numcoefficients"])])

    def getCentralCurvature(self):
        return self.params["curv"]()

    def sqrtfun(self, r2):
        (curv, cc, _) = self.getAsphereParameters()
        return _lib(r2).sqrt(1 - curv ** 2 * (1 + cc) * r2)

    def F(self, x, y):
        (curv, cc, acoeffs) = self.getAsphereParameters()
        r2 = x * x + y * y
        res = curv * r2 / (1 + self.sqrtfun(r2))
        for (n, an) in enumerate(acoeffs):
            res = res + an * r2 ** (n + 1)
        return res

    def gradF(self, x, y, z):
        xp = _lib(x)
        (curv, cc, acoeffs) = self.getAsphereParameters()
        r2 = x * x + y * y
        radial = curv / self.sqrtfun(r2)
        for (n, an) in enumerate(acoeffs):
            radial = radial + 2. * (n + 1) * an * r2 ** n
        return xp.stack((-x * radial, -y * radial, xp.ones_like(x)))

    def hessF(self, x, y, z):
        xp = _lib(x)
        (curv, cc, acoeffs) = self.getAsphereParameters()
        r2 = x * x + y * y
        sq = self.sqrtfun(r2)
        main1 = -curv / (2. * sq)
        main2 = -curv ** 3 * (1 + cc) / (4. * sq)
        for (n, an) in enumerate(acoeffs):
            main1 = main1 - an * (n + 1) * r2 ** n
            if n >= 1:
                main2 = main2 - an * (n + 1) * n * r2 ** (n - 1)
        h = xp.zeros((3, 3) + tuple(x.shape), dtype=x.dtype)
        h[0, 0] = 2 * (2 * main2 * x * x + main1)
        h[1, 1] = 2 * (2 * main2 * y * y + main1)
        h[0, 1] = h[1, 0] = 4 * main2 * x * y
        return h


class Biconic(ExplicitShape):
    """Polynomial biconic (reference surface_shape.py:609-706)."""

    @classmethod
    def p(cls, lc, curvx=0, ccx=0, curvy=0, ccy=0, coefficients=None, name=""):
        coefficients = [] if coefficients is None else list(coefficients)
        plist = [("curvx", curvx), ("curvy", curvy), ("ccx", ccx), ("ccy", ccy)] + \
                [("A" + str(2 * i + 2), a) for (i, (a, b)) in enumerate(coefficients)] + \
                [("B" + str(2 * i + 2), b) for (i, (a, b)) in enumerate(coefficients)]
        (ann, struct) = FreeShape.createAnnotationsAndStructure(lc, plist)
        ann["numcoefficients"] = len(coefficients)
        return cls(ann, struct, name)

    def setKind(self):
        self.kind = "shape_Biconic"

    def getBiconicParameters(self):
        return (self.params["curvx"](), self.params["curvy"](), self.params["ccx"](),
                self.params["ccy"](),
                [(self.params["A" + str(2 * i + 2)](), self.params["B" + str(2 * i + 2)]())
                 for i in range(self.annotations["numcoefficients"])])

    def getCentralCurvature(self):
        return 0.5 * (self.params["curvx"]() + self.params["curvy"]())

    def sqrtfun(self, x, y):
        (cx, cy, kx, ky, _) = self.getBiconicParameters()
        return _lib(x).sqrt(1 - cx ** 2 * (1 + kx) * x * x - cy ** 2 * (1 + ky) * y * y)

    def F(self, x, y):
        (cx, cy, kx, ky, coeffs) = self.getBiconicParameters()
        (r2, ast2) = (x * x + y * y, x * x - y * y)
        res = (cx * x * x + cy * y * y) / (1 + self.sqrtfun(x, y))
        for (n, (an, bn)) in enumerate(coeffs):
            res = res + an * (r2 - bn * ast2) ** (n + 1)
        return res

    def gradF(self, x, y, z):
        xp = _lib(x)
        (cx, cy, kx, ky, coeffs) = self.getBiconicParameters()
        (r2, ast2) = (x * x + y * y, x * x - y * y)
        sq = self.sqrtfun(x, y)
        base = cx * x * x + cy * y * y
        den = (sq + 1) ** 2 * sq
        gx = -cx * x * (cx * (kx + 1) * base + 2 * (sq + 1) * sq) / den
        gy = -cy * y * (cy * (ky + 1) * base + 2 * (sq + 1) * sq) / den
        for (n, (an, bn)) in enumerate(coeffs):
            u = r2 - bn * ast2
            gx = gx + 2 * an * (n + 1) * x * (bn - 1) * u ** n
            gy = gy - 2 * an * (n + 1) * y * (bn + 1) * u ** n
        return xp.stack((gx, gy, xp.ones_like(x)))


class XYPolynomials(ExplicitShape):

    @classmethod
    def p(cls, lc, normradius=1.0, coefficients=None, name=""):
        coefficients = [] if coefficients is None else list(coefficients)
        plist = [("normradius", normradius)] + \
                [("CX" + str(xp_) + "Y" + str(yp_), c)
                 for (xp_, yp_, c) in coefficients]
        (ann, struct) = FreeShape.createAnnotationsAndStructure(lc, plist)
        return cls(ann, struct, name)

    def setKind(self):
        self.kind = "shape_XYPolynomials"

    def getXYParameters(self):
        terms = []
        for (key, var) in self.params.items():
            if key[0] == "C":
                (xpow, ypow) = key[2:].split("Y")
                terms.append([int(xpow), int(ypow), var()])
        return (self.params["normradius"](), terms)

    def getCentralCurvature(self):
        (nr, terms) = self.getXYParameters()
        c = 0.0
        for (xpow, ypow, coeff) in terms:
            if (xpow, ypow) in ((2, 0), (0, 2)):
                c += coeff / nr ** 2
        return c

    def F(self, x, y):
        (nr, terms) = self.getXYParameters()
        res = _lib(x).zeros_like(x)
        for (xpow, ypow, c) in terms:
            res = res + x ** xpow * y ** ypow * (c / nr ** (xpow + ypow))
        return res

    def gradF(self, x, y, z):
        xp = _lib(x)
        (nr, terms) = self.getXYParameters()
        gx = xp.zeros_like(x)
        gy = xp.zeros_like(x)
        for (xpow, ypow, c) in terms:
            c = c / nr ** (xpow + ypow)
            if xpow >= 1:
                gx = gx - xpow * x ** (xpow - 1) * y ** ypow * c
            if ypow >= 1:
                gy = gy - ypow * x ** xpow * y ** (ypow - 1) * c
        return xp.stack((gx, gy, xp.ones_like(x)))

    def hessF(self, x, y, z):
        xp = _lib(x)
        (nr, terms) = self.getXYParameters()
        h = xp.zeros((3, 3) + tuple(x.shape), dtype=x.dtype)
        for (xpow, ypow, c) in terms:
            c = c / nr ** (xpow + ypow)
            if xpow >= 2:
                h[0, 0] -= xpow * (xpow - 1) * x ** (xpow - 2) * y ** ypow * c
            if xpow >= 1 and ypow >= 1:
                h[0, 1] -= xpow * ypow * x ** (xpow - 1) * y ** (ypow - 1) * c
            if ypow >= 2:
                h[1, 1] -= ypow * (ypow - 1) * x ** xpow * y ** (ypow - 2) * c
        h[1, 0] = h[0, 1]
        return h


def _poly_mul(a, b):
    out = {}
    for ((ax, ay), av) in a.items():
        for ((bx, by), bv) in b.items():
            key = (ax + bx, ay + by)
            out[key] = out.get(key, 0.0) + av * bv
    return out


def zernike_monomials(n, m):
    """Unnormalised Zernike polynomial Z_n^m = R_n^|m|(rho) cos / sin(|m| phi) (the
    reference's convention, surface_shape.py:1058-1083: m < 0 -> sin) as a dict
    {(px, py): coefficient} of monomials x^px y^py in the normalised coordinates."""
    import math
    om = abs(m)
    # angular part: Re / Im of (x + i y)^om
    ang = {}
    for j in range(om + 1):
        c = math.comb(om, j)
        # i^j: j even -> real (-1)^(j/2), j odd -> imaginary (-1)^((j-1)/2)
        if m >= 0 and j % 2 == 0:
            ang[(om - j, j)] = ang.get((om - j, j), 0.0) + c * (-1) ** (j // 2)
        if m < 0 and j % 2 == 1:
            ang[(om - j, j)] = ang.get((om - j, j), 0.0) + c * (-1) ** ((j - 1) // 2)
    if om == 0 and m >= 0:
        ang = {(0, 0): 1.0}
    r2 = {(2, 0): 1.0, (0, 2): 1.0}
    total = {}
    for k in range((n - om) // 2 + 1):
        coef = ((-1) ** k * math.factorial(n - k) /
                (math.factorial(k) * math.factorial((n + om) // 2 - k) *
                 math.factorial((n - om) // 2 - k)))
        term = {(0, 0): coef}
        for _ in range((n - om) // 2 - k):          # rho^(n - 2k) = rho^om (rho^2)^((n-om)/2-k)
            term = _poly_mul(term, r2)
        for (key, v) in _poly_mul(term, ang).items():
            total[key] = total.get(key, 0.0) + v
    return {k: v for (k, v) in total.items() if v != 0.0}


class Zernike(ExplicitShape):
    """Zernike series z = sum_j Z_j zern_j(x / R, y / R) (reference surface_shape.py
    :927-1113).  The series is expanded into monomials once (`getXYTerms`), which is also
    what the device evaluates (PYR_SHAPE_XYPOLY) -- unlike the reference's polar
    gradient (:1085-1095) the result is finite on the axis."""

    @classmethod
    def p(cls, lc, normradius=1., coefficients=None, name=""):
        coefficients = [] if coefficients is None else list(coefficients)
        plist = [("normradius", normradius)] + \
                [("Z" + str(i + 1), v) for (i, v) in enumerate(coefficients)]
        (ann, struct) = FreeShape.createAnnotationsAndStructure(lc, plist)
        ann["numcoefficients"] = len(coefficients)
        return cls(ann, struct, name)

    def setKind(self):
        self.kind = "shape_Zernike"

    def getZernikeParameters(self):
        return (self.params["normradius"](),
                [self.params["Z" + str(i + 1)]()
                 for i in range(self.annotations["numcoefficients"])])

    @staticmethod
    def jtonm(j):
        raise NotImplementedError()

    @staticmethod
    def nmtoj(n_m_pair):
        raise NotImplementedError()

    def getXYTerms(self):
        """(normradius, [(xpow, ypow, coefficient), ...]) of the whole series."""
        (nr, zc) = self.getZernikeParameters()
        total = {}
        for (j, val) in enumerate(zc, 1):
            if val == 0.0:
                continue
            (n, m) = self.jtonm(j)
            for (key, v) in zernike_monomials(n, m).items():
                total[key] = total.get(key, 0.0) + val * v
        return (nr, [(px, py, c) for ((px, py), c) in sorted(total.items()) if c != 0.0])

    def getCentralCurvature(self):
        (nr, terms) = self.getXYTerms()
        return sum(c / nr ** 2 for (px, py, c) in terms if (px, py) in ((2, 0), (0, 2)))

    def F(self, x, y):
        (nr, terms) = self.getXYTerms()
        res = _lib(x).zeros_like(x)
        for (px, py, c) in terms:
            res = res + x ** px * y ** py * (c / nr ** (px + py))
        return res

    def gradF(self, x, y, z):
        xp = _lib(x)
        (nr, terms) = self.getXYTerms()
        gx = xp.zeros_like(x)
        gy = xp.zeros_like(x)
        for (px, py, c) in terms:
            c = c / nr ** (px + py)
            if px >= 1:
                gx = gx - px * x ** (px - 1) * y ** py * c
            if py >= 1:
                gy = gy - py * x ** px * y ** (py - 1) * c
        return xp.stack((gx, gy, xp.ones_like(x)))


class ZernikeFringe(Zernike):

    def setKind(self):
        self.kind = "shape_ZernikeFringe"

    @staticmethod
    def jtonm(j):
        import math
        nsq = math.ceil(math.sqrt(j)) ** 2
        m_plus_n = int(2 * math.sqrt(nsq) - 2)
        m = int(math.ceil((nsq - j) / 2))
        n = m_plus_n - m
        return (n, int((-1) ** ((nsq - j) % 2)) * m)

    @staticmethod
    def nmtoj(n_m_pair):
        (n, m) = n_m_pair
        sign = (m > 0) - (m < 0)
        return int(((n + abs(m)) / 2 + 1) ** 2 - 2 * abs(m) + (1 - sign) / 2)


class ZernikeANSI(Zernike):

    def setKind(self):
        self.kind = "shape_ZernikeANSI"

    @staticmethod
    def jtonm(j):
        import math
        j -= 1
        n = math.floor((-1. + math.sqrt(1. + 8. * j)) * 0.5)
        m = n - 2 * j + n * (n + 1)
        return (n, -m)

    @staticmethod
    def nmtoj(n_m_pair):
        (n, m) = n_m_pair
        return int(((n + 2) * n + m) / 2) + 1


class GridSag(ExplicitShape):
    """Sag given on a rectangular grid, interpolated by the bicubic spline scipy's
    RectBivariateSpline fits through it (reference :861-924).  The device evaluates
    the same spline from its FITPACK knots / coefficients (csrc/pyr_shapes.cuh);
    the host-side getSag / getGrad below are for plotting and analysis (NumPy)."""

    @classmethod
    def p(cls, lc, xlin_ylin_zgrid, tol=1e-4, iterations=10, name=""):
        (xlinspace, ylinspace, zgrid) = xlin_ylin_zgrid
        (ann, struct) = FreeShape.createAnnotationsAndStructure(lc, paramlist=())
        ann["xlinspace"] = np.asarray(xlinspace).tolist()
        ann["ylinspace"] = np.asarray(ylinspace).tolist()
        ann["zgrid"] = np.asarray(zgrid).tolist()
        ann["tol"] = tol
        ann["iterations"] = iterations
        return cls(ann, struct, name)

    def setKind(self):
        self.kind = "shape_GridSag"

    def initialize_from_annotations(self):
        from scipy.interpolate import RectBivariateSpline
        self.interpolant = RectBivariateSpline(np.array(self.annotations["xlinspace"]),
                                               np.array(self.annotations["ylinspace"]),
                                               np.array(self.annotations["zgrid"]))

    @staticmethod
    def _host(a):
        return a.detach().cpu().numpy() if (torch is not None and isinstance(a, torch.Tensor)) \
            else np.asarray(a)

    def F(self, x, y):
        return self.interpolant.ev(self._host(x), self._host(y))

    def gradF(self, x, y, z):
        (x, y) = (self._host(x), self._host(y))
        return np.stack((-self.interpolant.ev(x, y, dx=1), -self.interpolant.ev(x, y, dy=1),
                         np.ones_like(x)))

    def hessF(self, x, y, z):
        (x, y) = (self._host(x), self._host(y))
        h = np.zeros((3, 3) + x.shape)
        h[0, 0] = -self.interpolant.ev(x, y, dx=2)
        h[0, 1] = h[1, 0] = -self.interpolant.ev(x, y, dx=1, dy=1)
        h[1, 1] = -self.interpolant.ev(x, y, dy=2)
        return h

    def getCentralCurvature(self):
        h = self.hessF(np.zeros(1), np.zeros(1), None)
        return float(-0.5 * (h[0, 0, 0] + h[1, 1, 0]))


class LinearCombination(ExplicitShape):
    """z = sum_i c_i F_i over sub-shapes (reference :709-777; the Zemax importer builds
    asphere + decentred Zernike this way, io/zmx.py:755-775).  On the device the
    sub-shape frames may differ from this shape's frame by a translation."""

    @classmethod
    def p(cls, lc, list_of_coefficients_and_shapes=None, name=""):
        pairs = [] if list_of_coefficients_and_shapes is None else \
            list(list_of_coefficients_and_shapes)
        (ann, struct) = FreeShape.createAnnotationsAndStructure(lc)
        ann["list_shape_coefficients"] = [c for (c, _) in pairs]
        struct["list_shapes"] = [s for (_, s) in pairs]
        return cls(ann, struct, name)

    def setKind(self):
        self.kind = "shape_LinearCombination"

    def _offset(self, shape):
        b = np.asarray(self.lc.localbasis, dtype=float)
        return b.T @ (np.asarray(shape.lc.globalcoordinates, dtype=float) -
                      np.asarray(self.lc.globalcoordinates, dtype=float))

    def F(self, x, y):
        z = 0.0
        for (c, shape) in zip(self.annotations["list_shape_coefficients"], self.list_shapes):
            o = self._offset(shape)
            z = z + c * (shape.getSag(x - o[0], y - o[1]) + o[2])
        return z

    def gradF(self, x, y, z):
        """Gradient of z - F.  Conic terms enter with the true sag gradient (the
        reference adds their implicit-function gradient and renormalises the z
        component by the sum of coefficients, :738-752)."""
        xp = _lib(x)
        (gx, gy) = (0.0, 0.0)
        for (c, shape) in zip(self.annotations["list_shape_coefficients"], self.list_shapes):
            o = self._offset(shape)
            g = shape.getGrad(x - o[0], y - o[1])
            gx = gx + c * g[0] / g[2]
            gy = gy + c * g[1] / g[2]
        return xp.stack((gx, gy, xp.ones_like(x)))

    def getCentralCurvature(self):
        return sum(c * s.getCentralCurvature() for (c, s) in
                   zip(self.annotations["list_shape_coefficients"], self.list_shapes))


accessible_shapes = {"shape_Conic": Conic, "shape_Cylinder": Cylinder, "shape_Asphere": Asphere,
                     "shape_Biconic": Biconic, "shape_XYPolynomials": XYPolynomials,
                     "shape_ZernikeFringe": ZernikeFringe, "shape_ZernikeANSI": ZernikeANSI,
                     "shape_GridSag": GridSag,
                     "shape_LinearCombination": LinearCombination}
