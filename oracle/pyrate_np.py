"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference's sequential
trace (`OpticalSystem.seqtrace`, mess42/pyrate pyrateoptics 0.4.0).

Not part of the product: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.  The product path
(pyrate_b200/) never does and has no CPU fallback.

Parity status: PINNED.  Every function below is checked against outputs of the
unmodified reference (generated in the build container by oracle/gen_golden.py,
committed under tests/golden/) by tests/test_oracle_golden.py.  The reference's
own tests do not pin seqtrace output, so the fixtures are
the pin.  Two reference behaviours are LAPACK/MINPACK-arbitrary and therefore
compared as invariants only: the isotropic E-field picked by an SVD with a
2-fold degenerate null space (material_isotropic.py:106-128) and the iterate
path of fsolve (surface_shape.py:457; the converged root is compared).

All citations are file:line relative to /root/reference/pyrateoptics/raytracer.

The restatement is vectorised over rays, works on plain dicts / arrays and
keeps the reference's data semantics:
  * a *bundle* is {"x","k","E": (P,3,N), "valid": (P,N) bool, "rayID": (N,)}
    (ray.py:34-105); `append` is a cumulative-AND of validity (ray.py:100)
  * a *path* is a list of bundles; path[0] is path[1] (optical_system.py:74,
    optical_element.py:331)
"""
import math

import numpy as np

try:                                   # anisotropic restatement only
    import scipy.linalg as sla
except Exception:                      # pragma: no cover
    sla = None


# ---------------------------------------------------------------------------
# frames: localcoordinates.py:171-180 (tilt order), :264-293 (composition),
#         helpers_math.py:69-83 (Rodrigues), :383-413 (point/direction maps)
# ---------------------------------------------------------------------------
def rodrigues(angle, axis):
    (a0, a1, a2) = axis
    km = np.array([[0., -a2, a1], [a2, 0., -a0], [-a1, a0, 0.]])
    return np.eye(3) + math.sin(angle) * km + (1. - math.cos(angle)) * (km @ km)


def tilt_matrix(tiltx, tilty, tiltz, tilt_then_decenter=0):
    rx = rodrigues(tiltx, (1, 0, 0))
    ry = rodrigues(tilty, (0, 1, 0))
    rz = rodrigues(tiltz, (0, 0, 1))
    if tilt_then_decenter == 0:
        return rz @ (ry @ rx)
    return rx @ (ry @ rz)


def child_frame(parent, decx=0., decy=0., decz=0., tiltx=0., tilty=0.,
                tiltz=0., tiltThenDecenter=0):
    """parent / result: (basis 3x3, origin 3)."""
    (pb, po) = parent
    rot = tilt_matrix(tiltx, tilty, tiltz, tiltThenDecenter)
    basis = pb @ rot
    dec = np.array([decx, decy, decz], dtype=float)
    if tiltThenDecenter == 0:
        origin = po + pb @ dec
    else:
        origin = po + basis @ dec
    return (basis, origin)


ROOT_FRAME = (np.eye(3), np.zeros(3))


def g2l_pts(frame, pts):
    (b, o) = frame
    return b.T @ (pts.T - o).T


def l2g_pts(frame, pts):
    (b, o) = frame
    return ((b @ pts).T + o).T


def g2l_dir(frame, d):
    return frame[0].T @ d


def l2g_dir(frame, d):
    return frame[0] @ d


# ---------------------------------------------------------------------------
# shapes: surface_shape.py
# ---------------------------------------------------------------------------
def conic_sag(curv, cc, x, y):
    """Conic.getSag / conic_function :198-219 (NaN where undefined)."""
    r2 = x * x + y * y
    sqrtterm = 1. - (1. + cc) * curv ** 2 * r2
    r2 = np.where(sqrtterm <= 0., np.nan, r2)
    sqrtterm = np.where(sqrtterm <= 0., 0., sqrtterm)
    return curv * r2 / (1. + np.sqrt(sqrtterm))


def conic_grad(curv, cc, x, y):
    """Conic.getGrad :221-237."""
    z = conic_sag(curv, cc, x, y)
    return np.vstack((-curv * x, -curv * y, 1. - curv * z * (1. + cc)))


def asphere_sag(curv, cc, acoeffs, x, y):
    """Asphere.F :529-537."""
    r2 = x * x + y * y
    with np.errstate(invalid="ignore"):
        res = curv * r2 / (1. + np.sqrt(1. - curv ** 2 * (1. + cc) * r2))
    for (n, an) in enumerate(acoeffs):
        res = res + an * r2 ** (n + 1)
    return res


def asphere_grad(curv, cc, acoeffs, x, y):
    """Asphere.gradF :539-555."""
    r2 = x * x + y * y
    with np.errstate(invalid="ignore"):
        sq = np.sqrt(1. - curv ** 2 * (1. + cc) * r2)
    gx = -curv * x / sq
    gy = -curv * y / sq
    for (n, an) in enumerate(acoeffs):
        gx = gx - 2. * x * (n + 1) * an * r2 ** n
        gy = gy - 2. * y * (n + 1) * an * r2 ** n
    return np.vstack((gx, gy, np.ones_like(x)))


def xypoly_sag(normradius, coeffs, x, y):
    """XYPolynomials.F :785-793; coeffs = [(xpow, ypow, c), ...]."""
    res = np.zeros_like(x)
    for (xp, yp, c) in coeffs:
        res = res + x ** int(xp) * y ** int(yp) * c / normradius ** (int(xp) + int(yp))
    return res


def xypoly_grad(normradius, coeffs, x, y):
    """XYPolynomials.gradF :795-807."""
    gx = np.zeros_like(x)
    gy = np.zeros_like(x)
    for (xp, yp, c) in coeffs:
        (xp, yp) = (int(xp), int(yp))
        norm = 1. / normradius ** (xp + yp)
        xpm1 = x ** (xp - 1) if xp >= 1 else np.zeros_like(x)
        ypm1 = y ** (yp - 1) if yp >= 1 else np.zeros_like(x)
        gx = gx - xp * xpm1 * y ** yp * c * norm
        gy = gy - yp * x ** xp * ypm1 * c * norm
    return np.vstack((gx, gy, np.ones_like(x)))


def biconic_sag(cx, cy, ccx, ccy, coeffs, x, y):
    """Biconic.F surface_shape.py:618-629; coeffs = [(a_n, b_n), ...]."""
    r2 = x * x + y * y
    ast2 = x * x - y * y
    with np.errstate(invalid="ignore"):
        sq = np.sqrt(1 - cx ** 2 * (1 + ccx) * x ** 2 - cy ** 2 * (1 + ccy) * y ** 2)
    res = (cx * x ** 2 + cy * y ** 2) / (1 + sq)
    for (n, (an, bn)) in enumerate(coeffs):
        res = res + an * (r2 - bn * ast2) ** (n + 1)
    return res


def biconic_grad(cx, cy, ccx, ccy, coeffs, x, y):
    """Biconic.gradF surface_shape.py:631-649."""
    r2 = x * x + y * y
    ast2 = x * x - y * y
    with np.errstate(invalid="ignore", divide="ignore"):
        sq = np.sqrt(1 - cx ** 2 * (1 + ccx) * x ** 2 - cy ** 2 * (1 + ccy) * y ** 2)
        gx = -cx * x * (cx * (ccx + 1) * (cx * x ** 2 + cy * y ** 2) + 2 * (sq + 1) * sq) / ((sq + 1) ** 2 * sq)
        gy = -cy * y * (cy * (ccy + 1) * (cx * x ** 2 + cy * y ** 2) + 2 * (sq + 1) * sq) / ((sq + 1) ** 2 * sq)
    for (n, (an, bn)) in enumerate(coeffs):
        gx = gx + 2 * an * (n + 1) * x * (bn - 1) * (-bn * ast2 + r2) ** n
        gy = gy - 2 * an * (n + 1) * y * (bn + 1) * (-bn * ast2 + r2) ** n
    return np.vstack((gx, gy, np.ones_like(x)))


def _zernike_nm(kind, j):
    """Single index -> (n, m): ZernikeFringe.jtonm :1118-1124, ZernikeANSI.jtonm :1139-1143."""
    if kind == "ZernikeFringe":
        nsq = math.ceil(math.sqrt(j)) ** 2
        m = int(math.ceil((nsq - j) / 2))
        n = int(2 * math.sqrt(nsq) - 2) - m
        return (n, int((-1) ** ((nsq - j) % 2)) * m)
    j -= 1
    n = math.floor((-1. + math.sqrt(1. + 8. * j)) * 0.5)
    return (n, -(n - 2 * j + n * (n + 1)))


def _zernike_term(n, m, x, y):
    """Z_n^m = R_n^|m|(rho) cos / sin(|m| phi) (zernike_norm :1058-1083) and its x, y
    derivatives, through w = x + i y:  rho^|m| cos(|m| phi) = Re w^|m|, sin -> Im.

    NOTE: the reference's own gradient (gradzernike_norm :1085-1095) divides the angular
    term by rho once instead of twice and is therefore inconsistent with its sag for
    every term with m != 0 (and NaN on the axis); the restatement keeps the correct
    derivative, and parity with the reference is pinned on the sag and, for traces, on
    rotationally symmetric series (m = 0), where the reference gradient is right."""
    om = abs(m)
    w = x + 1j * y
    r2 = x * x + y * y
    wm = w ** om if om > 0 else np.ones_like(w)
    dwm = om * w ** (om - 1) if om > 0 else np.zeros_like(w)
    pick = (lambda z: z.imag) if m < 0 else (lambda z: z.real)
    ang = pick(wm)
    (ang_x, ang_y) = (pick(dwm), pick(1j * dwm))
    rad = np.zeros_like(x)
    drad = np.zeros_like(x)                  # d rad / d (r2)
    for k in range((n - om) // 2 + 1):
        c = ((-1) ** k * math.factorial(n - k) /
             (math.factorial(k) * math.factorial((n + om) // 2 - k) *
              math.factorial((n - om) // 2 - k)))
        q = (n - om) // 2 - k
        rad = rad + c * r2 ** q
        if q >= 1:
            drad = drad + c * q * r2 ** (q - 1)
    return (rad * ang, 2 * x * drad * ang + rad * ang_x, 2 * y * drad * ang + rad * ang_y)


def zernike_sag_grad(shape, x, y):
    nr = shape["normradius"]
    (xp, yp) = (x / nr, y / nr)
    f = np.zeros_like(x)
    fx = np.zeros_like(x)
    fy = np.zeros_like(x)
    for (j, val) in enumerate(shape["coefficients"], 1):
        (n, m) = _zernike_nm(shape["kind"], j)
        (z, zx, zy) = _zernike_term(n, m, xp, yp)
        f = f + val * z
        fx = fx + val * zx / nr
        fy = fy + val * zy / nr
    return (f, fx, fy)


# ---------------------------------------------------------------------------
# GridSag (surface_shape.py:861-924): scipy.interpolate.RectBivariateSpline is a
# third-party dependency of the reference (SciPy / FITPACK `regrid` for the fit,
# `bispev` / `parder` for the evaluation; SciPy 1.18.1 here, the reference pins none).
# The fit is taken from SciPy (knots tx, ty and B-spline coefficients c); the
# evaluation is restated from FITPACK's published algorithm (fpbisp / fpbspl: clamp
# the argument to the grid, locate the knot interval, de Boor-Cox recursion for the 4
# non-zero cubic basis functions, tensor-product sum) and pinned against SciPy's own
# ev() in tests/test_oracle_golden.py.
# ---------------------------------------------------------------------------
def gridsag_fit(xlin, ylin, zgrid):
    from scipy.interpolate import RectBivariateSpline
    sp = RectBivariateSpline(np.asarray(xlin, dtype=float), np.asarray(ylin, dtype=float),
                             np.asarray(zgrid, dtype=float))
    (tx, ty) = sp.get_knots()
    return (np.array(tx), np.array(ty), np.array(sp.get_coeffs()))


def _bspline_basis(t, x):
    """Cubic basis values h (4, N) and derivatives dh (4, N) at x (N,) plus the interval
    index l (N,), knots t with 4-fold end knots (fpbspl)."""
    n = t.size
    x = np.clip(x, t[3], t[n - 4])
    l = np.clip(np.searchsorted(t, x, side="right") - 1, 3, n - 5)
    h = np.zeros((4, x.size))
    h[0] = 1.0
    quad = None
    for j in range(1, 4):
        hh = h[:j].copy()
        h[0] = 0.0
        for i in range(1, j + 1):
            (li, lj) = (l + i, l + i - j)
            f = hh[i - 1] / (t[li] - t[lj])
            h[i - 1] = h[i - 1] + f * (t[li] - x)
            h[i] = f * (x - t[lj])
        if j == 2:
            quad = h[:3].copy()
    dh = np.zeros_like(h)
    for i in range(4):
        a = quad[i - 1] / (t[l + i] - t[l + i - 3]) if i >= 1 else 0.0
        b = quad[i] / (t[l + i + 1] - t[l + i - 2]) if i <= 2 else 0.0
        dh[i] = 3.0 * (a - b)
    return (h, dh, l)


def gridsag_eval(spline, x, y):
    """(F, dF/dx, dF/dy) of the bicubic spline (tx, ty, c) at points x, y (N,)."""
    (tx, ty, c) = spline
    shape = np.shape(x)
    (x, y) = (np.ravel(np.asarray(x, dtype=float)), np.ravel(np.asarray(y, dtype=float)))
    (hx, dhx, lx) = _bspline_basis(tx, x)
    (hy, dhy, ly) = _bspline_basis(ty, y)
    cm = c.reshape((tx.size - 4, ty.size - 4))
    f = np.zeros_like(x)
    fx = np.zeros_like(x)
    fy = np.zeros_like(x)
    for i in range(4):
        for j in range(4):
            cij = cm[lx - 3 + i, ly - 3 + j]
            f = f + cij * hx[i] * hy[j]
            fx = fx + cij * dhx[i] * hy[j]
            fy = fy + cij * hx[i] * dhy[j]
    return (f.reshape(shape), fx.reshape(shape), fy.reshape(shape))


def _gridsag_spline(shape):
    if "_spline" not in shape:
        shape["_spline"] = gridsag_fit(shape["xlinspace"], shape["ylinspace"], shape["zgrid"])
    return shape["_spline"]


# ---------------------------------------------------------------------------
# LinearCombination (surface_shape.py:709-777)
# ---------------------------------------------------------------------------
def lincomb_sag(shape, x, y):
    """LinearCombination.F :714-731."""
    xlocal = np.vstack((x, y, np.zeros_like(x)))
    z = np.zeros_like(x)
    for (coef, sub) in shape["terms"]:
        xs = g2l_pts(sub["frame"], l2g_pts(shape["frame"], xlocal))
        xs[2] = shape_sag(sub, xs[0], xs[1])
        z = z + coef * g2l_pts(shape["frame"], l2g_pts(sub["frame"], xs))[2]
    return z


def lincomb_grad(shape, x, y):
    """LinearCombination.gradF :733-754 (z component divided by the coefficient sum)."""
    xlocal = np.vstack((x, y, np.zeros_like(x)))
    g = np.zeros_like(xlocal)
    total = 0.
    for (coef, sub) in shape["terms"]:
        xs = g2l_pts(sub["frame"], l2g_pts(shape["frame"], xlocal))
        gs = shape_grad(sub, xs[0], xs[1])
        g = g + coef * g2l_dir(shape["frame"], l2g_dir(sub["frame"], gs))
        total += coef
    g[2] = g[2] / total
    return g


def shape_sag(shape, x, y):
    kind = shape["kind"]
    if kind == "GridSag":
        return gridsag_eval(_gridsag_spline(shape), x, y)[0]
    if kind == "LinearCombination":
        return lincomb_sag(shape, x, y)
    if kind.startswith("Zernike"):
        return zernike_sag_grad(shape, x, y)[0]
    if kind == "Biconic":
        return biconic_sag(shape["curvx"], shape["curvy"], shape["ccx"], shape["ccy"],
                           shape["coefficients"], x, y)
    if kind == "Conic":
        return conic_sag(shape["curv"], shape["cc"], x, y)
    if kind == "Cylinder":                                 # Cylinder.getSag :360-367
        return conic_sag(shape["curv"], shape["cc"], np.zeros_like(x), y)
    if kind == "Asphere":
        return asphere_sag(shape["curv"], shape["cc"], shape["coefficients"], x, y)
    if kind == "XYPolynomials":
        return xypoly_sag(shape["normradius"], shape["coefficients"], x, y)
    raise NotImplementedError(kind)


def shape_grad(shape, x, y):
    kind = shape["kind"]
    if kind == "GridSag":                                  # GridSag.gradF :871-880
        (_, fx, fy) = gridsag_eval(_gridsag_spline(shape), x, y)
        return np.vstack((-fx, -fy, np.ones_like(x)))
    if kind == "LinearCombination":
        return lincomb_grad(shape, x, y)
    if kind.startswith("Zernike"):
        (_, fx, fy) = zernike_sag_grad(shape, x, y)
        return np.vstack((-fx, -fy, np.ones_like(x)))
    if kind == "Biconic":
        return biconic_grad(shape["curvx"], shape["curvy"], shape["ccx"], shape["ccy"],
                            shape["coefficients"], x, y)
    if kind == "Conic":
        return conic_grad(shape["curv"], shape["cc"], x, y)
    if kind == "Cylinder":
        # gradient of the cylinder's own implicit function c (y^2 + (1+cc) z^2) - 2 z (the
        # reference inherits Conic.getGrad, whose x component contradicts Cylinder.getSag)
        g = conic_grad(shape["curv"], shape["cc"], np.zeros_like(x), y)
        g[0] = 0.0
        return g
    if kind == "Asphere":
        return asphere_grad(shape["curv"], shape["cc"], shape["coefficients"], x, y)
    if kind == "XYPolynomials":
        return xypoly_grad(shape["normradius"], shape["coefficients"], x, y)
    raise NotImplementedError(kind)


def shape_normal(shape, x, y):
    """Shape.getNormal :100-112."""
    with np.errstate(invalid="ignore", divide="ignore"):
        g = shape_grad(shape, x, y)
        return g / np.sqrt(np.sum(g ** 2, axis=0))


# ---------------------------------------------------------------------------
# bundle helpers: ray.py
# ---------------------------------------------------------------------------
def new_bundle(x0, k0, e0, ray_id=None, splitted=False):
    """RayBundle.__init__ ray.py:35-75 (default E = (0,1,0))."""
    n = x0.shape[1]
    if e0 is None or len(e0) == 0:
        e0 = np.zeros_like(x0)
        e0[1] = 1.
    return {"x": x0.reshape((1, 3, n)), "k": k0.reshape((1, 3, n)),
            "E": e0.reshape((1, 3, n)), "valid": np.ones((1, n), dtype=bool),
            "rayID": np.arange(n) if ray_id is None else ray_id,
            "splitted": splitted}


def bundle_append(b, x, k, e, valid):
    """RayBundle.append ray.py:83-105."""
    b["x"] = np.vstack((b["x"], x[None]))
    b["k"] = np.vstack((b["k"], k[None]))
    b["E"] = np.vstack((b["E"], e[None]))
    b["valid"] = np.vstack((b["valid"], (b["valid"][-1] * valid)[None]))


def k_to_d(k, e):
    """RayBundle.returnKtoD ray.py:136-152 for one history row."""
    abs_e2 = np.sum(np.conj(e) * e, axis=0)
    ek = np.sum(e * k, axis=0)
    s = np.real(abs_e2 * k - ek * np.conj(e))
    with np.errstate(invalid="ignore", divide="ignore"):
        return s / np.sqrt(np.sum(s ** 2, axis=0))


# ---------------------------------------------------------------------------
# intersect: surface_shape.py:289-325 (conic), :448-465 (explicit shapes),
#            surface.py:116-135 and aperture.py:71-140 (aperture mask)
# ---------------------------------------------------------------------------
NEWTON_MAXIT = 60


def _explicit_t(shape, r0, d):
    """Root of r0_z + t d_z - F(r0_xy + t d_xy) (surface_shape.py:453-458).

    The reference hands the N equations to MINPACK hybrd as one system,
    start t = 0, xtol = 1e-6; the converged root has |residual| ~ 1e-15
    (SURVEY D2).  Restated as a per-ray Newton from the same start, run to
    machine precision.
    """
    t = np.zeros_like(r0[0])
    for _ in range(NEWTON_MAXIT):
        x = r0[0] + t * d[0]
        y = r0[1] + t * d[1]
        with np.errstate(invalid="ignore", divide="ignore"):
            res = r0[2] + t * d[2] - shape_sag(shape, x, y)
            g = shape_grad(shape, x, y)
            dres = g[0] * d[0] + g[1] * d[1] + g[2] * d[2]
            step = res / dres
        step = np.where(np.isfinite(step), step, 0.)
        t = t - step
        if np.all(np.abs(step) <= 1e-15 * (1. + np.abs(t))):
            break
    return t


def shape_intersect(shape, bundle):
    """Shape.intersect: appends one row to `bundle`."""
    frame = shape["frame"]
    r0 = g2l_pts(frame, bundle["x"][-1])
    d = g2l_dir(frame, k_to_d(bundle["k"][-1], bundle["E"][-1]))
    if shape["kind"] == "Conic":
        (curv, cc) = (shape["curv"], shape["cc"])
        f = d[2] - curv * (d[0] * r0[0] + d[1] * r0[1] + d[2] * r0[2] * (1 + cc))
        g = curv * (r0[0] ** 2 + r0[1] ** 2 + r0[2] ** 2 * (1 + cc)) - 2 * r0[2]
        h = -curv - cc * curv * d[2] ** 2
        square = f ** 2 + h * g
        with np.errstate(invalid="ignore", divide="ignore"):
            t = g / (f + np.sqrt(square))
            hit = r0 + d * t
            valid = square >= 0
    elif shape["kind"] == "Cylinder":
        # Cylinder.intersect :369-388 is dead code in the reference (it reads the
        # non-existent raybundle.rayDir) and its H is the rotationally symmetric conic's.
        # Corrected: the ray meets c (y^2 + (1+cc) z^2) - 2 z = 0 where
        #   -H t^2 - 2 F t + G = 0  with  H = -c (d_y^2 + (1+cc) d_z^2);
        # validity like Conic (square >= 0).  PARITY UNPINNED by construction: pinned on the
        # limiting cases instead (tests: x-independence, equality with Conic for d_x = 0 rays
        # in the plane x = 0, hit points satisfy the surface equation).
        (curv, cc) = (shape["curv"], shape["cc"])
        f = d[2] - curv * (d[1] * r0[1] + d[2] * r0[2] * (1 + cc))
        g = curv * (r0[1] ** 2 + r0[2] ** 2 * (1 + cc)) - 2 * r0[2]
        h = -curv * (d[1] ** 2 + (1 + cc) * d[2] ** 2)
        square = f ** 2 + h * g
        with np.errstate(invalid="ignore", divide="ignore"):
            t = g / (f + np.sqrt(square))
            hit = r0 + d * t
            valid = square >= 0
    else:
        t = _explicit_t(shape, r0, d)
        hit = r0 + d * t
        valid = np.ones_like(r0[0], dtype=bool)
    bundle_append(bundle, l2g_pts(frame, hit), bundle["k"][-1], bundle["E"][-1],
                  valid)


def aperture_mask(ap, x, y):
    kind = ap["kind"]
    with np.errstate(invalid="ignore"):
        if kind == "Base":
            return np.ones_like(x, dtype=bool)
        if kind == "Circular":
            r2 = x ** 2 + y ** 2
            return (r2 >= ap["minradius"] ** 2) * (r2 <= ap["maxradius"] ** 2)
        if kind == "Rectangular":
            (w, h) = (ap["width"], ap["height"])
            return (x >= -w * 0.5) * (x <= w * 0.5) * (y >= -h * 0.5) * (y <= h * 0.5)
    raise NotImplementedError(kind)


def surface_intersect(step, bundle):
    """Surface.intersect surface.py:116-135."""
    shape_intersect(step["shape"], bundle)
    ap = step["aperture"]
    loc = g2l_pts(ap["frame"], bundle["x"][-1])
    bundle["valid"][-1] = bundle["valid"][-1] * aperture_mask(ap, loc[0], loc[1])


def local_surface_normal(step, mat, xglob):
    """RayBundle.getLocalSurfaceNormal ray.py:156-161."""
    sh = step["shape"]
    xl = g2l_pts(sh["frame"], xglob)
    nl = shape_normal(sh, xl[0], xl[1])
    return g2l_dir(mat["frame"], l2g_dir(sh["frame"], nl))


# ---------------------------------------------------------------------------
# isotropic media: material/material_isotropic.py
# ---------------------------------------------------------------------------
def optical_index(mat, xlocal, wave):
    kind = mat["kind"]
    if kind in ("ConstantIndexGlass", "ConstantIndexGlassTIR"):   # :264, _tir.py:136
        return mat["n"]
    if kind == "ModelGlass":                  # :299-309 (Conrady)
        (n0, a, b) = mat["n0_A_B"]
        return n0 + a / wave + b / (wave ** 3.5)
    if kind == "IsotropicGrinMaterial":       # material_grin.py:97-98
        return mat["nfunc"](xlocal)
    raise NotImplementedError(kind)


def efield_svd(k, eps_scalar):
    """IsotropicMaterial.calc_e_field :72-128 (vectorised selection)."""
    n = k.shape[1]
    if n == 0:
        return np.zeros_like(k)
    kk = np.sum(k * k, axis=0)
    m = (-np.eye(3)[None] * kk[:, None, None]
         + np.einsum("in,jn->nij", k, k)
         + np.eye(3)[None] * (np.ones(n) * eps_scalar)[:, None, None])
    (u, sv, _) = np.linalg.svd(m)
    idx = np.argsort(np.abs(sv), axis=1)[:, 0]
    return u[np.arange(n), :, idx].T


def isotropic_deflect(mat, bundle, step, wave, mirror):
    """IsotropicMaterial.refract :163-199 / reflect :201-236."""
    fr = mat["frame"]
    xg = bundle["x"][-1]
    k1 = g2l_dir(fr, bundle["k"][-1])
    nrm = local_surface_normal(step, mat, xg)
    xl = g2l_pts(fr, xg)
    valid_normals = np.all(np.isfinite(nrm), axis=0)       # helpers_math.py:32-37
    kin = k1 - np.sum(k1 * nrm, axis=0) * nrm
    nidx = optical_index(mat, xl, wave)
    square = nidx ** 2 - np.sum(kin * kin, axis=0)          # :152-155, k dimensionless
    with np.errstate(invalid="ignore"):
        valid_refr = square > 0                             # :157 (TIR -> invalid)
        xi = np.sqrt(square)
    valid = bundle["valid"][-1] * valid_refr * valid_normals
    k2 = (-kin if mirror else kin) + xi * nrm               # :185 / :224
    k2v = k2[:, valid]
    eps = nidx ** 2 if np.isscalar(nidx) else (nidx ** 2)[valid]
    e2 = efield_svd(k2v, eps)
    return (new_bundle(xg[:, valid], l2g_dir(fr, k2v), l2g_dir(fr, e2),
                       bundle["rayID"][valid]),)


def isotropic_tir_deflect(mat, bundle, step, wave):
    """IsotropicMaterialTIR.refract material_isotropic_tir.py:46-119 (angle form of
    Snell's law), expression by expression.  Normalisations of reference
    defects: k is rotated into the material frame (see below), validity is ANDed with the incoming bundle's and with finite normals
    (the reference takes `1 - TIR` alone, :84, which revives vignetted rays), and the
    E field is computed from the compacted k (the reference passes mismatched widths
    to calc_e_field, :116, and raises as soon as one ray is totally reflected)."""
    fr = mat["frame"]
    xg = bundle["x"][-1]
    # :52 takes k in GLOBAL coordinates while the normal (:66) is in the material frame;
    # identical for material frames parallel to the global one (the pinned fixture),
    # normalised to one frame here for the others
    kin = g2l_dir(fr, bundle["k"][0])                       # :52 (row 0, same as -1 in
    normk = np.sqrt(np.sum(kin ** 2, axis=0))               # a homogeneous medium)
    index_before = normk
    index_after = np.real(optical_index(mat, np.zeros((3, 1)), wave)) * np.ones(normk.shape)
    dir_in = -(kin / normk)                                 # :63
    normal = -local_surface_normal(step, mat, xg)           # :66
    normal = normal / np.sqrt(np.sum(normal ** 2, axis=0))  # :69-71
    costheta = np.sum(dir_in * normal, axis=0)              # :75
    with np.errstate(invalid="ignore"):
        sintheta = np.sqrt(1 - costheta ** 2)
        sinthetadash = index_before / index_after * sintheta    # :79
        tir = sinthetadash > 1.0                                # :82
        costhetadash = np.sqrt(1 - sinthetadash ** 2)           # :87
    valid = (~tir) & bundle["valid"][-1] & np.all(np.isfinite(normal), axis=0)
    a = dir_in - costheta * normal                          # :91
    aout = np.zeros(a.shape)
    aout[:, valid] = -a[:, valid] * (index_before / index_after)[valid]   # :96-97
    dir_out = -normal * costhetadash + aout                 # :99
    dir_out = dir_out / np.sqrt(np.sum(dir_out ** 2, axis=0))             # :100-102
    k2 = dir_out * index_after                              # :105
    k2v = l2g_dir(fr, k2[:, valid])                         # :112
    e2 = efield_svd(k2v, mat["n"] ** 2)
    return (new_bundle(xg[:, valid], k2v, e2, bundle["rayID"][valid]),)


# ---------------------------------------------------------------------------
# anisotropic media: material/material.py:98-153, :214-223, :353-454 and
#                    material/material_anisotropic.py:70-155
# ---------------------------------------------------------------------------
def aniso_modes_sorted(eps, nrm, kpa):
    """sortKnormEField(x, n, kpa, n): 4 modes sorted by ascending S.n.

    Returns (k4, e4) of shape (4,3,N) complex.  Per ray: 6x6 generalised EVP
    (:353-403), drop non-finite eigenvalues, keep the 4 smallest |xi|
    (:441-449), E = lower 3 components (:452), k_i = kpa + xi_i n (:105-106),
    S_i = Re(|E|^2 k - (k.E) conj(E)) (:214-223), argsort of S.n (:148-151).
    """
    n = nrm.shape[1]
    k4 = np.zeros((4, 3, n), dtype=complex)
    e4 = np.zeros((4, 3, n), dtype=complex)
    eye = np.eye(3, dtype=complex)
    zero = np.zeros((3, 3), dtype=complex)
    for j in range(n):
        nv = nrm[:, j]
        kv = kpa[:, j]
        mm = -eye + np.outer(nv, nv)
        cm = np.outer(kv, nv) + np.outer(nv, kv)
        km = np.array(eps, dtype=complex) - np.dot(kv, kv) * eye + np.outer(kv, kv)
        a6 = np.vstack((np.hstack((cm, km)), np.hstack((-eye, zero))))
        b6 = -np.vstack((np.hstack((mm, zero)), np.hstack((zero, eye))))
        (w, vr) = sla.eig(a6, b=b6)
        fin = np.isfinite(w)
        w = w[fin]
        vr = vr[:, fin]
        if len(w) > 4:
            order = np.abs(w).argsort()
            w = w[order][:4]
            vr = vr[:, order][:, :4]
        ev = vr.T[:, 3:]
        kk = kv[None, :] + w[:, None] * nv[None, :]
        s = np.real(np.sum(np.conj(ev) * ev, axis=1)[:, None] * kk
                    - np.sum(kk * ev, axis=1)[:, None] * np.conj(ev))
        order = np.argsort(s @ nv)
        k4[:, :, j] = kk[order]
        e4[:, :, j] = ev[order]
    return (k4, e4)


def uniaxial_decomposition(eps, tol=1e-12):
    """(eps_o, eps_e, axis) of a real symmetric tensor with a twofold eigenvalue
    (eps = eps_o 1 + (eps_e - eps_o) a a^T), or None for isotropic / biaxial tensors."""
    eps = np.asarray(eps)
    if np.iscomplexobj(eps):
        if np.max(np.abs(eps.imag)) > tol:
            return None
        eps = eps.real
    if np.max(np.abs(eps - eps.T)) > tol * np.max(np.abs(eps)):
        return None
    (w, v) = np.linalg.eigh(eps)
    scale = np.max(np.abs(w))
    if abs(w[0] - w[1]) <= tol * scale and abs(w[2] - w[1]) > tol * scale:
        return (0.5 * (w[0] + w[1]), w[2], v[:, 2])
    if abs(w[2] - w[1]) <= tol * scale and abs(w[0] - w[1]) > tol * scale:
        return (0.5 * (w[1] + w[2]), w[0], v[:, 0])
    return None


def uniaxial_xi_roots(eps_o, eps_e, axis, nrm, kpa):
    """Closed form of the four roots xi of the dispersion relation of a uniaxial crystal
    for k = kpa + xi n (what calcXiEigenvectorsNorm, material.py:407-454, gets out of a
    6x6 generalised eigen-solve and the device out of the Fresnel quartic): the quartic
    factorises into the ordinary sphere  k.k = eps_o  and the extraordinary ellipsoid
    eps_o k.k + (eps_e - eps_o) (k.a)^2 = eps_o eps_e.  Returns (4, N) complex:
    ordinary +, ordinary -, extraordinary +, extraordinary -.  Groundwork for a
    root-finder-free crystal step; pinned on the reference's
    eigenvalues by tests/test_oracle_golden.py."""
    nrm = np.asarray(nrm, dtype=float)
    kpa = np.asarray(kpa, dtype=complex)
    a = np.asarray(axis, dtype=float).reshape(3, 1)
    delta = eps_e - eps_o
    kk = np.sum(kpa * kpa, axis=0)
    (ka, na) = (np.sum(kpa * a, axis=0), np.sum(nrm * a, axis=0))
    kn = np.sum(kpa * nrm, axis=0)                       # 0 for an in-plane kpa; kept general
    xo = np.sqrt(kn * kn - (kk - eps_o) + 0j)
    qa = eps_o + delta * na * na
    qb = eps_o * kn + delta * ka * na                    # half the linear coefficient
    qc = eps_o * kk + delta * ka * ka - eps_o * eps_e
    disc = np.sqrt(qb * qb - qa * qc + 0j)
    return np.stack((-kn + xo, -kn - xo, (-qb + disc) / qa, (-qb - disc) / qa))


def anisotropic_deflect(mat, bundle, step, wave, mirror, splitup):
    fr = mat["frame"]
    xg = bundle["x"][-1]
    k1 = g2l_dir(fr, bundle["k"][-1])
    nrm = local_surface_normal(step, mat, xg)
    kin = k1 - np.sum(k1 * nrm, axis=0) * nrm
    (k4, e4) = aniso_modes_sorted(mat["eps"], nrm, kin)
    if mirror:                                  # material_anisotropic.py:133-134
        (ka, kb, ea, eb) = (-k4[0], -k4[1], -e4[0], -e4[1])
    else:                                       # :89-91 / :102-106
        (ka, kb, ea, eb) = (k4[2], k4[3], e4[2], e4[3])
    if not splitup:
        ids = np.hstack((bundle["rayID"], bundle["rayID"]))
        return (new_bundle(np.hstack((xg, xg)), l2g_dir(fr, np.hstack((ka, kb))),
                           l2g_dir(fr, np.hstack((ea, eb))), ids, splitted=True),)
    return (new_bundle(xg, l2g_dir(fr, ka), l2g_dir(fr, ea), bundle["rayID"]),
            new_bundle(xg, l2g_dir(fr, kb), l2g_dir(fr, eb), bundle["rayID"]))


# ---------------------------------------------------------------------------
# GRIN: material/material_grin.py:106-220
# ---------------------------------------------------------------------------
_CBRT2 = 2.0 ** (1. / 3.)
GRIN_C = [1.0 / (2.0 * (2.0 - _CBRT2)), (1.0 - _CBRT2) / (2.0 * (2.0 - _CBRT2)),
          (1.0 - _CBRT2) / (2.0 * (2.0 - _CBRT2)), 1.0 / (2.0 * (2.0 - _CBRT2))]
GRIN_D = [1.0 / (2.0 - _CBRT2), (-_CBRT2) / (2.0 - _CBRT2),
          1.0 / (2.0 - _CBRT2), 0.0]


def grin_integrate(mat, bundle, step, per_ray_energy=False, history=True):
    """symplecticintegrator :106-213 (lock-step over the bundle).

    per_ray_energy=False reproduces the reference's bundle-summed energy test
    (:164-176); True is the per-ray normalisation the device path uses
    (SURVEY Appendix B-10).  history=False appends only the final row.
    """
    fr = mat["frame"]
    tau = mat["ds"]
    pos = g2l_pts(fr, bundle["x"][-1])
    vel = optical_index(mat, pos, None) * g2l_dir(
        fr, k_to_d(bundle["k"][-1], bundle["E"][-1]))
    n = pos.shape[1]
    valid = np.ones(n, dtype=bool)
    final = np.zeros(n, dtype=bool)
    upd_pos = pos.copy()
    upd_vel = vel.copy()
    shape = step["shape"]
    steps = 0
    while not np.all(final):
        steps += 1
        for (c, d) in zip(GRIN_C, GRIN_D):
            pos = pos + tau * c * 2.0 * vel
            optin = mat["nfunc"](pos)
            vel = vel + tau * d * 2.0 * optin * np.array(
                [mat["dndx"](pos), mat["dndy"](pos), mat["dndz"](pos)])
        if per_ray_energy:
            bad = np.abs(np.sum(vel ** 2, axis=0) - optin ** 2) > mat["energyviolation"]
            valid[bad] = False
        else:
            if abs(np.sum(vel ** 2) - np.sum(optin ** 2)) > mat["energyviolation"]:
                valid[:] = False
        xs = g2l_pts(shape["frame"], l2g_pts(fr, pos))
        with np.errstate(invalid="ignore"):
            final = xs[2] - shape_sag(shape, xs[0], xs[1]) > 0
        valid[~mat["bnd"](pos)] = False
        final = final | ~valid
        upd_pos[:, ~final] = pos[:, ~final]
        upd_vel[:, ~final] = vel[:, ~final]
        if history:
            newk = upd_vel / mat["nfunc"](upd_pos)
            e = efield_svd(newk, mat["nfunc"](pos) ** 2)
            bundle_append(bundle, l2g_pts(fr, upd_pos), l2g_dir(fr, newk),
                          l2g_dir(fr, e), valid)
    if not history:
        newk = upd_vel / mat["nfunc"](upd_pos)
        e = efield_svd(newk, mat["nfunc"](pos) ** 2)
        bundle_append(bundle, l2g_pts(fr, upd_pos), l2g_dir(fr, newk),
                      l2g_dir(fr, e), valid)
    return steps


# ---------------------------------------------------------------------------
# the path: optical_element.py:324-379, optical_system.py:73-94
# ---------------------------------------------------------------------------
def material_propagate(mat, bundle, step, **grin_kw):
    if mat["kind"] == "IsotropicGrinMaterial":
        grin_integrate(mat, bundle, step, **grin_kw)     # material_grin.py:215-220
    surface_intersect(step, bundle)                      # material_isotropic.py:238-247


def material_deflect(mat, bundle, step, wave, mirror, splitup):
    if mat["kind"] == "AnisotropicMaterial":
        return anisotropic_deflect(mat, bundle, step, wave, mirror, splitup)
    if mat["kind"] == "ConstantIndexGlassTIR" and not mirror:
        return isotropic_tir_deflect(mat, bundle, step, wave)
    return isotropic_deflect(mat, bundle, step, wave, mirror)


def _copy_bundle(b):
    return {k: (v.copy() if isinstance(v, np.ndarray) else v) for (k, v) in b.items()}


def seqtrace(system, x0, k0, e0, wave=0.5876e-3, splitup=False, **grin_kw):
    """Returns list of paths; a path is a list of bundles.

    `system` = {"background": material, "steps": [step, ...]} with
    step = {"shape", "aperture", "mat_minus", "mat_plus", "is_mirror"};
    materials are compared by identity (optical_element.py:109-126).
    """
    first = new_bundle(np.array(x0, copy=True), np.array(k0, copy=True),
                       None if e0 is None else np.array(e0, copy=True))
    background = system["background"]
    current = background
    paths = [[first]]
    prev_elem = None
    for step in system["steps"]:
        elem = step.get("elem", 0)
        if elem != prev_elem:
            # a new element: its seqtrace starts in the background medium
            # (optical_element.py:328) with a path holding the bundle it was handed
            # (:331), which the system appends to its own path (optical_system.py:91)
            current = background
            for p in paths:
                p.append(p[-1])
            prev_elem = elem
        mirror = bool(step.get("is_mirror", False))
        mn = step["mat_minus"] if step["mat_minus"] is not None else background
        pn = step["mat_plus"] if step["mat_plus"] is not None else background
        for p in paths:
            material_propagate(current, p[-1], step, **grin_kw)
        if not mirror:
            current = pn if (mn is current) else mn
        new_paths = []
        for p in paths:
            res = material_deflect(current, p[-1], step, wave, mirror, splitup)
            for rb in res[1:]:
                memo = {}                  # deepcopy keeps "same object twice"
                q = [memo.setdefault(id(b), _copy_bundle(b)) for b in p]
                q.append(rb)
                new_paths.append(q)
            p.append(res[0])
        paths = paths + new_paths
    # (the very first hand-over is the copied input bundle: path[0] is path[1])
    return paths


# ---------------------------------------------------------------------------
# spot statistics: analysis/ray_analysis.py:44-86
# ---------------------------------------------------------------------------
def centroid(x):
    return np.sum(x, axis=1) / (x.shape[1] + 1e-17)


def rms_spot(x, ref):
    delta = x - np.asarray(ref).reshape((3, 1))
    return math.sqrt(np.sum(delta ** 2) / (x.shape[1] - 1 + 1e-17))


# ---------------------------------------------------------------------------
# building a `system` dict from a pyrate_b200.configs spec (own frame math,
# independent of the product's lowering)
# ---------------------------------------------------------------------------
def _exec_grin_source(source, names):
    env = {}
    exec(source, env)          # test infrastructure; trusted config text
    return [env[n] for n in names]


def _shape_from_spec(skind, skw, frame):
    if skind == "GridSag":
        from pyrate_b200.configs import grid_arrays      # spec data -> arrays only
        (xl, yl, zg) = grid_arrays(skw["grid"])
        return {"kind": skind, "frame": frame, "xlinspace": xl, "ylinspace": yl, "zgrid": zg}
    if skind == "LinearCombination":
        terms = []
        for (coef, tkind, tkw, dec) in skw["terms"]:
            sub_frame = child_frame(frame, **dec) if dec else frame
            terms.append((coef, _shape_from_spec(tkind, tkw, sub_frame)))
        return {"kind": skind, "frame": frame, "terms": terms}
    shape = dict(skw)
    shape["kind"] = skind
    shape["frame"] = frame
    if skind in ("Asphere", "Biconic", "ZernikeFringe", "ZernikeANSI"):
        shape.setdefault("coefficients", [])
    return shape


def system_from_spec(spec):
    frame = child_frame(ROOT_FRAME, decz=0.0)           # "object_lc0"
    background = {"kind": "ConstantIndexGlass", "n": 1.0, "frame": ROOT_FRAME}
    mats = {}
    steps = []
    last = None
    for surf in spec["surfaces"]:
        frame = child_frame(frame, **surf["lc"])
        (skind, skw) = surf["shape"]
        shape = _shape_from_spec(skind, skw, frame)
        if surf["aperture"] is None:
            ap = {"kind": "Base"}
        else:
            (akind, akw) = surf["aperture"]
            ap = dict(akw)
            ap["kind"] = akind.replace("Aperture", "")
            if ap["kind"] == "Circular":
                ap.setdefault("minradius", 0.0)
        ap["frame"] = frame
        key = surf["mat"]
        if key is not None and key not in mats:
            (mkind, mkw) = spec["materials"][key]
            m = {"kind": mkind, "frame": frame}
            if mkind in ("ConstantIndexGlass", "ConstantIndexGlassTIR"):
                m["n"] = mkw["n"]
            elif mkind == "ModelGlass":
                m["n0_A_B"] = tuple(mkw["n0_A_B"])
            elif mkind == "AnisotropicMaterial":
                m["eps"] = np.array(mkw["epstensor"])
            elif mkind == "IsotropicGrinMaterial":
                (m["nfunc"], m["dndx"], m["dndy"], m["dndz"], m["bnd"]) = \
                    _exec_grin_source(mkw["source"], mkw["names"])
                m["ds"] = mkw["ds"]
                m["energyviolation"] = mkw["energyviolation"]
            else:
                raise NotImplementedError(mkind)
            mats[key] = m
        split = spec.get("split_after")
        steps.append({"name": surf["name"], "shape": shape, "aperture": ap,
                      "mat_minus": mats.get(last), "mat_plus": mats.get(key),
                      "is_mirror": bool(surf["opt"].get("is_mirror", False)),
                      "elem": 0 if (split is None or len(steps) < split) else 1})
        last = key
    return {"background": background, "steps": steps}
