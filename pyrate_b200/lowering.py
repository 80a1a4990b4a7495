"""Lowering: (OpticalSystem, elementsequence) -> flat table of PyrStep records.

Walks the object graph exactly as the reference's per-surface loop does
(raytracer/optical_element.py:324-375): material look-up with background
substitution (:344-346), toggle by object identity (:109-126, :353-356),
mirrors keep the current material (:357-358), every element starts in the
background medium (:328).  Every parameter is re-read here on each trace
(`v()` / `.evaluate()`), because optimisers mutate variables between traces
(optimize/optimize.py:73-91).

Objects are inspected by duck typing (class names along the MRO, attribute
names of the reference), so the *reference's own* object graph lowers too --
that is how identical systems are fed to both engines in the parity tests.
"""
import ctypes as C
import threading

import numpy as np

from . import _native as nat
from .raytracer.material.material_grin import (BOUNDARY_KINDS, PROFILE_KINDS,
                                               profile_functions)


class LoweringError(Exception):
    pass


_MRO_CACHE = {}


def _mro_names(obj):
    t = type(obj)
    names = _MRO_CACHE.get(t)
    if names is None:
        names = _MRO_CACHE[t] = frozenset(c.__name__ for c in t.__mro__)
    return names


class _Caches(threading.local):
    """Per-thread caches that live for ONE lower() call (two threads may lower at the
    same time, e.g. one per GPU): frames by id(lc), media by (id(material), wave)."""
    frames = None
    media = None


_CACHES = _Caches()


def _frame(lc):
    cache = _CACHES.frames
    if cache is not None:
        hit = cache.get(id(lc))
        if hit is not None:
            return hit
    f = _frame_uncached(lc)
    if cache is not None:
        cache[id(lc)] = f
    return f


def _frame_uncached(lc):
    f = nat.PyrFrame()
    f.r[:] = np.asarray(lc.localbasis, dtype=np.float64).reshape(-1).tolist()
    f.o[:] = np.asarray(lc.globalcoordinates, dtype=np.float64).reshape(-1).tolist()
    return f


def _value(v):
    return float(v() if callable(v) else v)


# ---------------------------------------------------------------------------
# media
# ---------------------------------------------------------------------------
def _verify_grin_profile(mat, profile, samples=64):
    """Compare user Python index functions with the declared device profile."""
    user = None
    if hasattr(mat, "user_functions"):
        user = mat.user_functions()
    elif hasattr(mat, "nfunc"):             # reference object
        kw = getattr(mat, "params", {})
        user = tuple((lambda x, f=f: f(x, **kw)) for f in
                     (mat.nfunc, mat.dndx, mat.dndy, mat.dndz)) + \
            (mat.boundaryfunction,)
    if user is None:
        return
    ours = profile_functions(profile)
    rng = np.random.default_rng(12345)
    pts = rng.uniform(-3.0, 3.0, (3, samples))
    for (i, (fu, fo)) in enumerate(zip(user[:4], ours[:4])):
        (a, b) = (np.asarray(fu(pts), dtype=float), np.asarray(fo(pts), dtype=float))
        if not np.allclose(a, b, rtol=1e-12, atol=1e-13):
            raise LoweringError(
                "GRIN material %r: annotations['device_profile'] disagrees with the "
                "Python source (function #%d, max |diff| = %.3e); refusing to trace a "
                "different medium" % (getattr(mat, "name", "?"), i,
                                      float(np.max(np.abs(a - b)))))
    far = rng.uniform(-30.0, 30.0, (3, samples))
    if not np.array_equal(np.asarray(user[4](far), dtype=bool),
                          np.asarray(ours[4](far), dtype=bool)):
        raise LoweringError("GRIN material %r: boundary function disagrees with the "
                            "declared device boundary" % (getattr(mat, "name", "?"),))


def lower_medium(mat, wave):
    cache = _CACHES.media
    key = (id(mat), wave)
    if cache is not None:
        hit = cache.get(key)
        if hit is not None:
            return hit
    m = _lower_medium_uncached(mat, wave)
    if cache is not None:
        cache[key] = m
    return m


def _lower_medium_uncached(mat, wave):
    m = nat.PyrMedium()
    names = _mro_names(mat)
    m.frame = _frame(mat.lc)
    if "AnisotropicMaterial" in names:
        m.kind = nat.MEDIUM_ANISO
        eps = np.asarray(mat.epstensor, dtype=complex)
        if eps.shape != (3, 3):
            raise LoweringError("epsilon tensor must be 3x3")
        flat = eps.reshape(-1)
        for i in range(9):
            m.eps[2 * i] = flat[i].real
            m.eps[2 * i + 1] = flat[i].imag
        m.n = 0.0
    elif "IsotropicGrinMaterial" in names:
        m.kind = nat.MEDIUM_ISO_GRIN
        prof = mat.annotations.get("device_profile")
        if prof is None and mat.annotations.get("device_source") is not None:
            # user-written index function (CUDA source, compiled at run time by NVRTC:
            # pyrate_b200/grin_jit.py); the engine runs such segments through its own kernels
            m.grin_profile = nat.GRIN_USER
            m.grin_boundary = nat.BND_NONE
            m.grin_ds = float(mat.annotations["ds"])
            m.grin_energy_tol = float(mat.annotations["energyviolation"])
            m.grin_max_steps = int(mat.annotations.get("max_steps", 0))
            m.n = 0.0
            return m
        if prof is None:
            raise LoweringError(
                "GRIN material %r needs annotations['device_profile'] (catalogue: %s) or "
                "annotations['device_source'] (CUDA expressions, compiled at run time)"
                % (getattr(mat, "name", "?"), ", ".join(sorted(PROFILE_KINDS))))
        _verify_grin_profile(mat, prof)
        m.grin_profile = PROFILE_KINDS[prof["kind"]]
        for (i, v) in enumerate(prof["params"]):
            m.grin_p[i] = float(v)
        bspec = prof.get("boundary", {"kind": "none", "params": []})
        m.grin_boundary = BOUNDARY_KINDS[bspec["kind"]]
        for (i, v) in enumerate(bspec.get("params", [])):
            m.grin_b[i] = float(v)
        m.grin_ds = float(mat.annotations["ds"])
        m.grin_energy_tol = float(mat.annotations["energyviolation"])
        m.grin_max_steps = int(mat.annotations.get("max_steps", 0))
        m.n = 0.0
    elif "IsotropicMaterial" in names:
        m.kind = nat.MEDIUM_ISO_CONST
        m.n = float(mat.get_optical_index(None, wave))
    else:
        raise LoweringError("unsupported material class %s" % type(mat).__name__)
    return m


# ---------------------------------------------------------------------------
# surfaces
# ---------------------------------------------------------------------------
def _xy_terms(shape):
    terms = []
    for (key, var) in shape.params.items():
        if key[0] == "C":
            (xp, yp) = key[2:].split("Y")
            terms.append((int(xp), int(yp), _value(var)))
    return terms


def _zernike_terms(shape):
    """Zernike series (reference or mirror object) -> monomial terms; the index
    convention comes from the object's own jtonm (Fringe / ANSI)."""
    from .raytracer.surface_shape import zernike_monomials
    total = {}
    for i in range(int(shape.annotations["numcoefficients"])):
        val = _value(shape.params["Z" + str(i + 1)])
        if val == 0.0:
            continue
        (n, m) = type(shape).jtonm(i + 1)
        for (key, v) in zernike_monomials(int(n), int(m)).items():
            total[key] = total.get(key, 0.0) + val * v
    return [(px, py, c) for ((px, py), c) in sorted(total.items()) if c != 0.0]


def _grid_spline(shape):
    """GridSag -> FITPACK representation (tx, ty, c) of the bicubic spline.

    The reference keeps a scipy RectBivariateSpline in `interpolant`
    (surface_shape.py:911-921); its knots and coefficients are read directly, so
    the device evaluates the very same spline.  Objects without one (annotations
    only) are fitted here the same way."""
    interp = getattr(shape, "interpolant", None)
    if interp is None:
        from scipy.interpolate import RectBivariateSpline
        ann = shape.annotations
        interp = RectBivariateSpline(np.array(ann["xlinspace"]), np.array(ann["ylinspace"]),
                                     np.array(ann["zgrid"]))
    (tx, ty) = interp.get_knots()
    c = interp.get_coeffs()
    if tuple(interp.degrees) != (3, 3):
        raise LoweringError("grid sag: bicubic splines only")
    (tx, ty, c) = (np.ascontiguousarray(a, dtype=np.float64) for a in (tx, ty, c))
    if c.size != (tx.size - 4) * (ty.size - 4) or tx.size < 8 or ty.size < 8:
        raise LoweringError("grid sag: unexpected spline layout")
    return (tx, ty, c)


def _simple_shape(shape):
    """One non-composite explicit shape -> dict(kind, curv, cc, curv2, cc2,
    normradius, n_coeff, coeff, xpow, ypow, grid); coeff/xpow/ypow are lists in
    the device layout of `kind`."""
    names = _mro_names(shape)
    d = {"curv": 0.0, "cc": 0.0, "curv2": 0.0, "cc2": 0.0, "normradius": 1.0,
         "n_coeff": 0, "coeff": [], "xpow": [], "ypow": [], "grid": None}
    if "Cylinder" in names:
        # surface_shape.py:328-388: a Conic subclass whose own intersect is dead code in
        # the reference (AttributeError at :380); traced with the corrected quadratic
        d["kind"] = nat.SHAPE_CYLINDER
        d["curv"] = _value(shape.curvature)
        d["cc"] = _value(shape.conic)
    elif "Conic" in names:
        d["kind"] = nat.SHAPE_CONIC
        d["curv"] = _value(shape.curvature)
        d["cc"] = _value(shape.conic)
    elif "Asphere" in names:
        d["kind"] = nat.SHAPE_ASPHERE
        d["curv"] = _value(shape.params["curv"])
        d["cc"] = _value(shape.params["cc"])
        ncoef = int(shape.annotations["numcoefficients"])
        d["n_coeff"] = ncoef
        d["coeff"] = [_value(shape.params["A" + str(2 * i + 2)]) for i in range(ncoef)]
    elif "Biconic" in names:
        d["kind"] = nat.SHAPE_BICONIC
        (d["curv"], d["cc"]) = (_value(shape.params["curvx"]), _value(shape.params["ccx"]))
        (d["curv2"], d["cc2"]) = (_value(shape.params["curvy"]), _value(shape.params["ccy"]))
        ncoef = int(shape.annotations["numcoefficients"])
        if ncoef > 16:
            raise LoweringError("too many biconic coefficient pairs")
        d["n_coeff"] = ncoef
        if ncoef:
            co = [0.0] * (16 + ncoef)
            for i in range(ncoef):
                co[i] = _value(shape.params["A" + str(2 * i + 2)])
                co[16 + i] = _value(shape.params["B" + str(2 * i + 2)])
            d["coeff"] = co
    elif "XYPolynomials" in names or "Zernike" in names:
        d["kind"] = nat.SHAPE_XYPOLY
        terms = _zernike_terms(shape) if "Zernike" in names else _xy_terms(shape)
        d["normradius"] = _value(shape.params["normradius"])
        d["n_coeff"] = len(terms)
        for (xp, yp, c) in terms:
            if not (0 <= xp < 64 and 0 <= yp < 64):
                raise LoweringError("XY exponent out of range")
            d["xpow"].append(xp)
            d["ypow"].append(yp)
            d["coeff"].append(c)
    elif "GridSag" in names:
        d["kind"] = nat.SHAPE_GRIDSAG
        d["grid"] = _grid_spline(shape)
    else:
        raise LoweringError("unsupported shape class %s" % type(shape).__name__)
    if len(d["coeff"]) > nat.MAX_COEFF:
        raise LoweringError("too many shape coefficients (%d > %d)" %
                            (len(d["coeff"]), nat.MAX_COEFF))
    d["xpow"] += [0] * (len(d["coeff"]) - len(d["xpow"]))
    d["ypow"] += [0] * (len(d["coeff"]) - len(d["ypow"]))
    return d


def _relative_offset(sub_lc, lc):
    """Origin of `sub_lc` in the frame `lc`; the two frames may differ by a
    translation only."""
    (b, bs) = (np.asarray(lc.localbasis, dtype=float), np.asarray(sub_lc.localbasis, dtype=float))
    if np.max(np.abs(b.T @ bs - np.eye(3))) > 1e-12:
        raise LoweringError("LinearCombination: sub-shape frames rotated against the "
                            "combination's frame are not supported")
    return b.T @ (np.asarray(sub_lc.globalcoordinates, dtype=float) -
                  np.asarray(lc.globalcoordinates, dtype=float))


def lower_surface(surface, st):
    """Fill shape / aperture fields of PyrStep `st`.  Grid-sag spline arrays (host
    NumPy, uploaded by the engine) are left in `st._grid`."""
    shape = surface.shape
    names = _mro_names(shape)
    st.shape_frame = _frame(shape.lc)
    st._grid = None
    if "LinearCombination" in names:
        # surface_shape.py:709-777.  A Conic term is lowered as an asphere without
        # coefficients: sag and TRUE sag gradient (the reference mixes the implicit
        # conic gradient into the sum there, :738-752 "TODO: is this correct?")
        coeffs = list(shape.annotations["list_shape_coefficients"])
        subs = list(shape.list_shapes)
        if not (1 <= len(subs) <= nat.MAX_TERMS) or len(coeffs) != len(subs):
            raise LoweringError("LinearCombination: 1..%d terms supported" % nat.MAX_TERMS)
        st.shape_kind = nat.SHAPE_COMBINATION
        st.n_terms = len(subs)
        off = 0
        for (t, (w, sub)) in enumerate(zip(coeffs, subs)):
            if "LinearCombination" in _mro_names(sub):
                raise LoweringError("nested LinearCombination")
            d = _simple_shape(sub)
            tm = st.terms[t]
            tm.kind = nat.SHAPE_ASPHERE if d["kind"] == nat.SHAPE_CONIC else d["kind"]
            tm.weight = float(w)
            (tm.dx, tm.dy, tm.dz) = [float(v) for v in _relative_offset(sub.lc, shape.lc)]
            (tm.curv, tm.cc, tm.curv2, tm.cc2) = (d["curv"], d["cc"], d["curv2"], d["cc2"])
            tm.normradius = d["normradius"]
            tm.n_coeff = d["n_coeff"]
            tm.coeff_off = off
            tm.coeff_len = len(d["coeff"])
            if off + tm.coeff_len > nat.MAX_COEFF:
                raise LoweringError("LinearCombination: too many coefficients in total")
            for (i, c) in enumerate(d["coeff"]):
                (st.coeff[off + i], st.xpow[off + i], st.ypow[off + i]) = \
                    (c, d["xpow"][i], d["ypow"][i])
            off += tm.coeff_len
            if d["grid"] is not None:
                if st._grid is not None:
                    raise LoweringError("LinearCombination: one GridSag term at most")
                st._grid = d["grid"]
    else:
        d = _simple_shape(shape)
        st.shape_kind = d["kind"]
        (st.curv, st.cc, st.curv2, st.cc2) = (d["curv"], d["cc"], d["curv2"], d["cc2"])
        st.normradius = d["normradius"]
        st.n_coeff = d["n_coeff"]
        for (i, c) in enumerate(d["coeff"]):
            (st.coeff[i], st.xpow[i], st.ypow[i]) = (c, d["xpow"][i], d["ypow"][i])
        st._grid = d["grid"]
    if st.shape_kind not in (nat.SHAPE_CONIC, nat.SHAPE_CYLINDER):
        ann = getattr(shape, "annotations", {})
        # The reference's fsolve stops at xtol = annotations["tol"] (1e-6) but
        # lands at ~1e-15 residual; Newton is run to machine precision and
        # `iterations` is only used as a (generous) cap.
        st.newton_tol = 1e-14
        st.newton_maxit = max(30, 3 * int(ann.get("iterations", 10)))
    ap = surface.aperture
    apn = _mro_names(ap)
    st.aperture_frame = _frame(ap.lc)
    if "CircularAperture" in apn:
        st.aperture_kind = nat.AP_CIRCULAR
        st.aperture_p[0] = float(ap.annotations["minradius"])
        st.aperture_p[1] = float(ap.annotations["maxradius"])
    elif "RectangularAperture" in apn:
        st.aperture_kind = nat.AP_RECTANGULAR
        st.aperture_p[0] = float(ap.annotations["width"])
        st.aperture_p[1] = float(ap.annotations["height"])
    elif "BaseAperture" in apn:
        st.aperture_kind = nat.AP_BASE
    else:
        raise LoweringError("unsupported aperture class %s" % type(ap).__name__)


class LoweredStep(object):
    """One PyrStep plus the bookkeeping the host needs."""
    __slots__ = ("st", "elemkey", "elem_index", "surfkey", "before_obj", "after_obj",
                 "is_aniso_deflect", "is_stop")

    def __init__(self):
        self.st = nat.PyrStep()


def lower(system, elementsequence, wave, splitup=False):
    """Returns list[LoweredStep] for `elementsequence`
    = [(elemkey, [(surfkey, {"is_mirror": .., "is_stop": ..}), ...]), ...]."""
    (_CACHES.frames, _CACHES.media) = ({}, {})
    try:
        return _lower(system, elementsequence, wave)
    finally:
        (_CACHES.frames, _CACHES.media) = (None, None)


def _lower(system, elementsequence, wave):
    background = system.material_background
    out = []
    last_deflector = None          # medium object that set |k| last
    for (elem_index, (elemkey, subseq)) in enumerate(elementsequence):
        if elemkey not in system.elements:
            raise LoweringError("unknown element %r" % (elemkey,))
        elem = system.elements[elemkey]
        current = background                                    # :328
        conn = elem.annotations["surf_mat_connection"]
        for (surfkey, surfoptions) in subseq:
            mirror = bool(surfoptions.get("is_mirror", False))
            surface = elem.surfaces[surfkey]
            (mnkey, pnkey) = conn[surfkey]
            mnmat = elem.materials.get(mnkey, background)       # :345-346
            pnmat = elem.materials.get(pnkey, background)
            before = current
            if not mirror:                                       # :353-358
                current = pnmat if (mnmat is current) else mnmat
            after = current
            ls = LoweredStep()
            st = ls.st
            lower_surface(surface, st)
            st.interaction = nat.REFLECT if mirror else nat.REFRACT
            st.before = lower_medium(before, wave)
            st.after = lower_medium(after, wave)
            first = len(out) == 0
            if first or st.before.kind == nat.MEDIUM_ANISO:
                st.dir_mode = nat.DIR_POYNTING
            else:
                st.dir_mode = nat.DIR_K
            st.k_norm_hint = 0.0
            if (not first and last_deflector is not None and
                    last_deflector is before and
                    st.before.kind == nat.MEDIUM_ISO_CONST):
                st.k_norm_hint = st.before.n
            ls.is_aniso_deflect = st.after.kind == nat.MEDIUM_ANISO
            st.split = 1 if (ls.is_aniso_deflect) else 0
            st.mode = nat.STEP_FULL
            ls.elemkey = elemkey
            ls.elem_index = elem_index
            ls.surfkey = surfkey
            ls.before_obj = before
            ls.after_obj = after
            ls.is_stop = bool(surfoptions.get("is_stop", False))
            out.append(ls)
            last_deflector = after
    return out


def lower_batch(system, elementsequence, waves):
    """Lower one sequence for a batch of wavelengths (<= PYR_MAX_WAVES): returns
    (per-wavelength lowered lists, the batch table).  The batch table is the first
    wavelength's with `before_n_w` / `after_n_w` filled in; everything but the media
    indices must agree between the wavelengths (it does: dispersion is the only
    wavelength dependence of a sequence), and all media must be homogeneous isotropic."""
    waves = [float(w) for w in waves]
    if not (1 <= len(waves) <= nat.MAX_WAVES):
        raise LoweringError("a wavelength batch holds 1..%d wavelengths" % nat.MAX_WAVES)
    per_wave = [lower(system, elementsequence, w) for w in waves]
    batch = lower(system, elementsequence, waves[0])
    for (i, ls) in enumerate(batch):
        st = ls.st
        if st.before.kind != nat.MEDIUM_ISO_CONST or st.after.kind != nat.MEDIUM_ISO_CONST:
            raise LoweringError("wavelength batches need homogeneous isotropic media "
                                "(entry %r)" % (ls.surfkey,))
        if st.shape_kind in (nat.SHAPE_GRIDSAG, nat.SHAPE_COMBINATION):
            raise LoweringError("wavelength batches do not carry grid-sag / combination "
                                "shapes (entry %r): one launch per bundle" % (ls.surfkey,))
        if i > 0 and st.dir_mode == nat.DIR_POYNTING:
            raise LoweringError("wavelength batches need k-directed segments")
        for (w, low) in enumerate(per_wave):
            other = low[i].st
            if (other.shape_kind, other.aperture_kind, other.interaction, other.dir_mode) != \
                    (st.shape_kind, st.aperture_kind, st.interaction, st.dir_mode):
                raise LoweringError("sequence differs between wavelengths")
            st.before_n_w[w] = other.before.n
            st.after_n_w[w] = other.after.n
    return (per_wave, batch)


def step_array(lowered):
    arr = (nat.PyrStep * len(lowered))()
    for (i, ls) in enumerate(lowered):
        C.memmove(C.addressof(arr[i]), C.addressof(ls.st), C.sizeof(nat.PyrStep))
    return arr
