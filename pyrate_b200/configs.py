"""Benchmark / parity configurations of BASELINE.json as plain data.

Every configuration is a *chain* system: each surface frame is the child of the
previous one (as `build_simple_optical_element` does it, reference
pyrateoptics/__init__.py:124-211).  `build_system(spec, api)` instantiates a
spec through any class namespace that follows the pyrateoptics constructor API
(`X.p(...)`): the reference's own classes (oracle/refshim.py, build container
only) or this package's (`pyrate_b200.api()`), which is itself a drop-in check
of the host API mirror.

Sources:
  C1 demos/demo_doublet.py:48-96
  C2 demos/data/double_gauss_rudolph_1897_v2.spd:7-40 (Rudolph 1897, d-line indices)
  C3 demos/demo_asphere.py:47-57 with A4=1e-7, A6=-1e-10
  C4 geometry of demos/demo_anisotropic_doublet.py:55-116, uniaxial crystals
  C5 demos/demo_grin.py:52-146
"""
import math

import numpy as np

DLINE = 0.5876e-3  # mm, reference globalconstants.py:38


def hexapolar(rings):
    """Hexapolar pupil sampling on the unit disk (defined by us).

    Ring 0 is the centre point, ring j = 1..rings carries 6j points at radius
    j/rings and angle 2 pi i/(6j).  N = 1 + 3 R (R + 1).
    """
    rings = int(rings)
    if rings <= 0:
        return np.zeros(1), np.zeros(1)
    j = np.repeat(np.arange(1, rings + 1), 6 * np.arange(1, rings + 1))
    start = 3 * (j - 1) * j          # number of ring points before ring j
    i = np.arange(j.size) - start
    ang = 2.0 * math.pi * i / (6.0 * j)
    rad = j / float(rings)
    px = np.concatenate(([0.0], rad * np.cos(ang)))
    py = np.concatenate(([0.0], rad * np.sin(ang)))
    return px, py


def hexapolar_range(rings, lo, hi):
    """Points lo..hi-1 of hexapolar(rings) without building the whole raster
    (ray shards of a rank in a multi-GPU run)."""
    i = np.arange(lo, hi, dtype=np.int64)
    j = np.floor((3.0 + np.sqrt(np.maximum(9.0 + 12.0 * (i - 1), 0.0))) / 6.0).astype(np.int64)
    j = np.maximum(j, 0)
    # correct rounding at ring boundaries: ring j holds indices [1+3(j-1)j, 1+3j(j+1))
    j = np.where(i < 1 + 3 * (j - 1) * j, j - 1, j)
    j = np.where(i >= 1 + 3 * j * (j + 1), j + 1, j)
    j = np.where(i == 0, 0, j)
    idx = i - (1 + 3 * (j - 1) * j)
    jj = np.maximum(j, 1)
    ang = 2.0 * math.pi * idx / (6.0 * jj)
    rad = j / float(max(rings, 1))
    return np.where(i == 0, 0.0, rad * np.cos(ang)), np.where(i == 0, 0.0, rad * np.sin(ang))


def collimated_shard(rings, radius, z0, lo, hi, kdir=(0.0, 0.0, 1.0),
                     efield=(0.0, 1.0, 0.0)):
    """Rays lo..hi-1 of collimated_bundle(rings, ...)."""
    (px, py) = hexapolar_range(rings, lo, hi)
    n = px.size
    x0 = np.empty((3, n))
    x0[0] = radius * px
    x0[1] = radius * py
    x0[2] = z0
    k0 = np.ascontiguousarray(np.repeat(np.asarray(kdir, dtype=float)[:, None], n, axis=1))
    e0 = np.ascontiguousarray(np.repeat(np.asarray(efield, dtype=float)[:, None], n, axis=1))
    return x0, k0, e0


def hexapolar_count(rings):
    return 1 + 3 * rings * (rings + 1)


def rings_for(nrays):
    """Smallest ring count with at least `nrays` points."""
    r = int(math.ceil((-3.0 + math.sqrt(9.0 + 12.0 * (nrays - 1))) / 6.0))
    while hexapolar_count(r) < nrays:
        r += 1
    return max(r, 0)


def collimated_bundle(rings, radius, z0, kdir=(0.0, 0.0, 1.0),
                      efield=(0.0, 1.0, 0.0)):
    """x0, k0, E0 as (3, N) float64 arrays; |k0| = 1 (background index)."""
    px, py = hexapolar(rings)
    n = px.size
    x0 = np.empty((3, n))
    x0[0] = radius * px
    x0[1] = radius * py
    x0[2] = z0
    k0 = np.repeat(np.asarray(kdir, dtype=float)[:, None], n, axis=1)
    e0 = np.repeat(np.asarray(efield, dtype=float)[:, None], n, axis=1)
    return x0, np.ascontiguousarray(k0), np.ascontiguousarray(e0)


def _conic(name, decz, curv=0.0, cc=0.0, mat=None, aperture=None, opt=None,
           **lckw):
    lc = {"decz": decz}
    lc.update(lckw)
    return {"name": name, "lc": lc,
            "shape": ("Conic", {"curv": curv, "cc": cc}),
            "aperture": aperture, "mat": mat, "opt": opt or {}}


def _circ(r):
    return ("CircularAperture", {"maxradius": r})


C1_DOUBLET = {
    "name": "c1_doublet",
    "surfaces": [
        _conic("stop", 0.0, opt={"is_stop": True}),
        _conic("front", -1.048, curv=1. / 62.8, mat="bk7", aperture=_circ(12.7)),
        _conic("cement", 4.0, curv=-1. / 45.7, mat="sf5", aperture=_circ(12.7)),
        _conic("rear", 2.5, curv=-1. / 128.2, mat=None, aperture=_circ(12.7)),
        _conic("image", 97.2),
    ],
    "materials": {"bk7": ("ConstantIndexGlass", {"n": 1.5168}),
                  "sf5": ("ConstantIndexGlass", {"n": 1.6727})},
    "bundle": {"rings": 18, "radius": 11.43, "z0": -5.0},
    "s_counted": 3,
}

_N1, _N2, _N3 = 1.52345716953278, 1.54813778400421, 1.60341715812683
_DG = [  # (name, radius, decz_before, material after)
    ("obj", 0.0, 0.0, None),
    ("s1", 43.5219015164416, 20.0, "g1"),
    ("s2", -22.9137468057709, 6.63155126149407, "g2"),
    ("s3", -54.9830148717816, 5.21926274183504, None),
    ("s4", -39.5941875735904, 5.86288269028215, "g3"),
    ("s5", 208.80831176475, 4.91037116991389, None),
    ("stop", 0.0, 3.65022739559149, None),
    ("s6", -208.80831176475, 3.65022739559149, "g3"),
    ("s7", 39.5941875735904, 4.91037116991389, None),
    ("s8", 54.9830148717816, 5.86288269028215, "g2"),
    ("s9", 22.9137468057709, 5.21926274183504, "g1"),
    ("s10", -43.5219015164416, 6.63155126149407, None),
    ("image", 0.0, 100.0, None),
]

C2_DOUBLEGAUSS = {
    "name": "c2_doublegauss",
    "surfaces": [
        _conic(nm, dz, curv=(1. / r if abs(r) > 1e-17 else 0.0), mat=mt,
               opt=({"is_stop": True} if nm == "stop" else {}))
        for (nm, r, dz, mt) in _DG],
    "materials": {"g1": ("ConstantIndexGlass", {"n": _N1}),
                  "g2": ("ConstantIndexGlass", {"n": _N2}),
                  "g3": ("ConstantIndexGlass", {"n": _N3})},
    "bundle": {"rings": 1825, "radius": 5.0, "z0": 0.0},
    "s_counted": 10,
}

C3_ASPHERE = {
    "name": "c3_asphere",
    "surfaces": [
        _conic("stop", 0.0, opt={"is_stop": True}),
        _conic("front", 5.0, mat="glass"),
        {"name": "back", "lc": {"decz": 20.0},
         "shape": ("Asphere", {"curv": -1. / 50.0, "cc": -1.0,
                               "coefficients": [0.0, 1e-7, -1e-10]}),
         "aperture": None, "mat": None, "opt": {}},
        _conic("image", 100.0),
    ],
    "materials": {"glass": ("ConstantIndexGlass", {"n": 1.5168})},
    "bundle": {"rings": 1825, "radius": 11.43, "z0": -5.0},
    "s_counted": 2,
}

_EPS1 = np.diag([1.658 ** 2, 1.486 ** 2, 1.658 ** 2]).tolist()  # calcite-like
_EPS2 = np.diag([1.544 ** 2, 1.553 ** 2, 1.544 ** 2]).tolist()  # quartz-like

C4_ANISOTROPIC = {
    "name": "c4_anisotropic",
    "surfaces": [
        _conic("stop", 0.0, opt={"is_stop": True}),
        _conic("front", -1.048, curv=1. / 62.8, mat="crystal1", aperture=_circ(12.7)),
        _conic("cement", 4.0, curv=-1. / 45.7, mat="crystal2", aperture=_circ(12.7)),
        _conic("rear", 2.5, curv=-1. / 128.2, mat=None, aperture=_circ(12.7)),
        _conic("image", 97.2),
    ],
    "materials": {"crystal1": ("AnisotropicMaterial", {"epstensor": _EPS1}),
                  "crystal2": ("AnisotropicMaterial", {"epstensor": _EPS2})},
    "bundle": {"rings": 577, "radius": 11.43, "z0": -5.0},
    "s_counted": 3,
}

GRIN_SOURCE = r"""
import numpy as np

grin_strength = 0.5


def nfunc(x, **kw):
    return grin_strength*np.exp(-x[0]**2 - 4.*x[1]**2)+1.0


def dndx(x, **kw):
    return -2.*x[0]*grin_strength*np.exp(-x[0]**2 - 4.*x[1]**2)


def dndy(x, **kw):
    return -2.*4.*x[1]*grin_strength*np.exp(-x[0]**2 - 4.*x[1]**2)


def dndz(x, **kw):
    return np.zeros_like(x[0])


def bnd(x):
    return x[0]**2 + x[1]**2 < 10.**2
"""

C5_GRIN = {
    "name": "c5_grin",
    "surfaces": [
        _conic("object", 0.0, opt={"is_stop": True}),
        _conic("surf1", 10.0, curv=1. / 24.0, mat="grin", aperture=_circ(5.0),
               tiltx=5. * math.pi / 180.0),
        _conic("surf2", 20.0, curv=-1. / 24.0, mat=None, aperture=_circ(5.0),
               tiltx=10. * math.pi / 180.0),
        _conic("image", 10.0),
    ],
    "materials": {"grin": ("IsotropicGrinMaterial", {
        "source": GRIN_SOURCE,
        "names": ("nfunc", "dndx", "dndy", "dndz", "bnd"),
        "parameterlist": [("n0", 0.5)],
        "ds": 0.05, "energyviolation": 0.01,
        # closed-catalogue device profile: n = n0 + g exp(-a x^2 - b y^2),
        # boundary x^2 + y^2 < r^2  (checked against the Python source at
        # lowering time, pyrate_b200/lowering.py)
        "device_profile": {"kind": "gaussian_xy",
                           "params": [1.0, 0.5, 1.0, 4.0],
                           "boundary": {"kind": "cylinder", "params": [10.0]}},
    })},
    "bundle": {"rings": 5773, "radius": 2.5, "z0": -5.0},
    "s_counted": 2,
}

CONFIGS = {c["name"]: c for c in
           (C1_DOUBLET, C2_DOUBLEGAUSS, C3_ASPHERE, C4_ANISOTROPIC, C5_GRIN)}


def build_system(spec, api):
    """Instantiate `spec` with the classes in `api`; returns (system, seq).

    spec["split_after"] = i puts surfaces[:i] into element "stdelem" and the rest
    into "elem2" (the ray must be in the background medium at the cut)."""
    s = api.OpticalSystem.p(name=spec["name"])
    lc0 = s.addLocalCoordinateSystem(
        api.LocalCoordinates.p(name="object_lc0", decz=0.0),
        refname=s.rootcoordinatesystem.name)
    split = spec.get("split_after")
    elem = api.OpticalElement.p(lc0, name="stdelem")
    elems = [("stdelem", elem, [])]
    refname = lc0.name
    lastmat = None
    made = set()
    for (isurf, surf) in enumerate(spec["surfaces"]):
        if split is not None and isurf == split:
            if lastmat is not None:
                raise ValueError("split_after must cut in the background medium")
            elem = api.OpticalElement.p(lc0, name="elem2")
            elems.append(("elem2", elem, []))
            made = set()
        lc = elem.addLocalCoordinateSystem(
            api.LocalCoordinates.p(name=surf["name"] + "_lc", **surf["lc"]),
            refname=refname)
        (shapekind, shapekw) = surf["shape"]
        shape = make_shape(api, elem, lc, shapekind, shapekw, surf["name"])
        aperture = None
        if surf["aperture"] is not None:
            (apkind, apkw) = surf["aperture"]
            aperture = getattr(api, apkind).p(lc, **apkw)
        surface = api.Surface.p(lc, shape=shape, aperture=aperture,
                                name=surf["name"] + "_surf")
        mat = surf["mat"]
        if mat is not None and mat not in made:
            elem.addMaterial(mat, _make_material(api, lc, spec["materials"][mat],
                                                 mat))
            made.add(mat)
        elem.addSurface(surf["name"], surface, (lastmat, mat))
        lastmat = mat
        refname = lc.name
        elems[-1][2].append((surf["name"], dict(surf["opt"])))
    for (key, el, _) in elems:
        s.addElement(key, el)
    return s, [(key, seq) for (key, _, seq) in elems]


def grid_arrays(grid):
    """Sag grid of a GridSag spec: {"x": (lo, hi, n), "y": (lo, hi, n), "poly":
    [(c, px, py), ...]} -> (x, y, Z) with Z[i, j] = sum c x_i^px y_j^py."""
    x = np.linspace(*grid["x"])
    y = np.linspace(*grid["y"])
    (xx, yy) = np.meshgrid(x, y, indexing="ij")
    z = np.zeros_like(xx)
    for (c, px, py) in grid["poly"]:
        z = z + c * xx ** px * yy ** py
    return (x, y, z)


def make_shape(api, elem, lc, kind, kw, name):
    """Shape object of a spec entry.  GridSag: kw = {"grid": ...} (grid_arrays);
    LinearCombination: kw = {"terms": [(coefficient, kind, kw, decenter), ...]} with
    decenter = None (the combination's frame) or {"decx": .., "decy": ..} (a child
    frame, like the Zemax importer's decentred Zernike term, io/zmx.py:755-775)."""
    if kind == "GridSag":
        return api.GridSag.p(lc, grid_arrays(kw["grid"]))
    if kind == "LinearCombination":
        pairs = []
        for (i, (coef, skind, skw, dec)) in enumerate(kw["terms"]):
            sub_lc = lc
            if dec:
                sub_lc = elem.addLocalCoordinateSystem(
                    api.LocalCoordinates.p(name="%s_term%d_lc" % (name, i), **dec),
                    refname=lc.name)
            pairs.append((coef, make_shape(api, elem, sub_lc, skind, skw,
                                           "%s_term%d" % (name, i))))
        return api.LinearCombination.p(lc, list_of_coefficients_and_shapes=pairs)
    return getattr(api, kind).p(lc, **kw)


def _make_material(api, lc, matspec, name):
    (kind, kw) = matspec
    if kind == "ConstantIndexGlass":
        return api.ConstantIndexGlass.p(lc, n=kw["n"], name=name)
    if kind == "ConstantIndexGlassTIR":
        return api.ConstantIndexGlassTIR.p(lc, n=kw["n"], name=name)
    if kind == "ModelGlass":
        return api.ModelGlass.p(lc, n0_A_B=tuple(kw["n0_A_B"]), name=name)
    if kind == "AnisotropicMaterial":
        return api.AnisotropicMaterial.p(lc, np.array(kw["epstensor"]),
                                         name=name)
    if kind == "IsotropicGrinMaterial":
        m = api.IsotropicGrinMaterial.p(lc, kw["source"], *kw["names"],
                                        parameterlist=list(kw["parameterlist"]),
                                        name=name)
        m.annotations["ds"] = kw["ds"]
        m.annotations["energyviolation"] = kw["energyviolation"]
        # ignored by the reference; read by pyrate_b200's lowering
        if "device_profile" in kw:
            m.annotations["device_profile"] = kw["device_profile"]
        if "device_source" in kw:
            m.annotations["device_source"] = kw["device_source"]
        return m
    raise ValueError("unknown material kind %r" % (kind,))


def config_bundle(spec, rings=None, kdir=(0.0, 0.0, 1.0),
                  efield=(0.0, 1.0, 0.0)):
    b = spec["bundle"]
    return collimated_bundle(b["rings"] if rings is None else rings,
                             b["radius"], b["z0"], kdir, efield)


# ---------------------------------------------------------------------------
# Extra parity-only systems (not BASELINE configs): general frames, mirror,
# rectangular aperture, dispersion model, XY polynomial, vignetting / TIR.
# ---------------------------------------------------------------------------
X1_TILTED = {
    "name": "x1_tilted",
    "surfaces": [
        _conic("stop", 0.0, opt={"is_stop": True}),
        _conic("front", 3.0, curv=1. / 40.0, cc=-0.5, mat="mg",
               aperture=_circ(9.0), decx=0.3, decy=-0.2,
               tiltx=2.0 * math.pi / 180.0, tilty=-1.5 * math.pi / 180.0),
        _conic("back", 5.0, curv=-1. / 55.0, cc=0.3, mat=None,
               aperture=("RectangularAperture", {"width": 14.0, "height": 12.0}),
               tiltz=10.0 * math.pi / 180.0, tiltx=-1.0 * math.pi / 180.0,
               tiltThenDecenter=1, decy=0.25),
        _conic("mirror", 30.0, curv=-1. / 200.0, mat=None,
               tiltx=12.0 * math.pi / 180.0, opt={"is_mirror": True}),
        _conic("image", -25.0, tiltx=12.0 * math.pi / 180.0),
    ],
    "materials": {"mg": ("ModelGlass", {"n0_A_B": (1.49749699179,
                                                   0.0100998734374 * 1e-3,
                                                   0.000328623343942 * (1e-3) ** 3.5)})},
    "bundle": {"rings": 8, "radius": 6.0, "z0": -4.0},
    "s_counted": 3,
}

X2_XYPOLY = {
    "name": "x2_xypoly",
    "surfaces": [
        _conic("stop", 0.0, opt={"is_stop": True}),
        _conic("front", 4.0, curv=1. / 80.0, mat="glass"),
        {"name": "back", "lc": {"decz": 6.0, "tilty": 1.0 * math.pi / 180.0},
         "shape": ("XYPolynomials", {"normradius": 10.0, "coefficients": [
             (2, 0, -0.9), (0, 2, -1.1), (1, 1, 0.05), (3, 0, 0.02),
             (1, 2, -0.03), (4, 0, 0.004), (2, 2, 0.006), (0, 4, -0.005)]}),
         "aperture": None, "mat": None, "opt": {}},
        _conic("image", 60.0),
    ],
    "materials": {"glass": ("ConstantIndexGlass", {"n": 1.62})},
    "bundle": {"rings": 6, "radius": 7.0, "z0": -3.0},
    "s_counted": 2,
}

X3_VIGNETTE = {   # misses, aperture clipping and total internal reflection
    "name": "x3_vignette",
    "surfaces": [
        _conic("stop", 0.0, opt={"is_stop": True}),
        _conic("front", 2.0, curv=1. / 9.0, mat="dense", aperture=_circ(8.5)),
        _conic("back", 6.0, curv=-1. / 7.5, mat=None, aperture=_circ(7.0)),
        _conic("image", 12.0),
    ],
    "materials": {"dense": ("ConstantIndexGlass", {"n": 1.9})},
    "bundle": {"rings": 10, "radius": 9.5, "z0": -2.0},
    "s_counted": 2,
}

def _rotated_biaxial():
    # real symmetric positive-definite tensor with principal values
    # (2.2, 2.6, 3.1) in a frame rotated about a skew axis
    (a, b, c) = (0.4, -0.3, 0.25)
    rx = np.array([[1, 0, 0], [0, math.cos(a), -math.sin(a)], [0, math.sin(a), math.cos(a)]])
    ry = np.array([[math.cos(b), 0, math.sin(b)], [0, 1, 0], [-math.sin(b), 0, math.cos(b)]])
    rz = np.array([[math.cos(c), -math.sin(c), 0], [math.sin(c), math.cos(c), 0], [0, 0, 1]])
    q = rz @ ry @ rx
    return (q @ np.diag([2.2, 2.6, 3.1]) @ q.T).tolist()


X4_BIAXIAL = {   # biaxial crystal lens, tilted material frame, into air
    "name": "x4_biaxial",
    "surfaces": [
        _conic("stop", 0.0, opt={"is_stop": True}),
        _conic("front", 3.0, curv=1. / 50.0, mat="biax", aperture=_circ(8.0),
               tiltx=3.0 * math.pi / 180.0, tilty=-2.0 * math.pi / 180.0),
        _conic("back", 5.0, curv=-1. / 60.0, cc=-0.4, mat=None,
               tiltx=-3.0 * math.pi / 180.0),
        _conic("image", 40.0),
    ],
    "materials": {"biax": ("AnisotropicMaterial", {"epstensor": _rotated_biaxial()})},
    "bundle": {"rings": 3, "radius": 5.0, "z0": -2.0},
    "s_counted": 2,
}

X5_DEGENERATE = {   # the demo's own choice (demos/demo_anisotropic_doublet.py:92-93):
    # "crystals" with isotropic tensors -> every mode pair is degenerate
    "name": "x5_degenerate",
    "surfaces": C4_ANISOTROPIC["surfaces"],
    "materials": {"crystal1": ("AnisotropicMaterial",
                               {"epstensor": (1.5168 ** 2 * np.eye(3)).tolist()}),
                  "crystal2": ("AnisotropicMaterial",
                               {"epstensor": (1.6727 ** 2 * np.eye(3)).tolist()})},
    "bundle": {"rings": 3, "radius": 11.43, "z0": -5.0},
    "s_counted": 3,
}

X6_BICONIC = {
    "name": "x6_biconic",
    "surfaces": [
        _conic("stop", 0.0, opt={"is_stop": True}),
        {"name": "front", "lc": {"decz": 3.0, "tiltz": 20.0 * math.pi / 180.0},
         "shape": ("Biconic", {"curvx": 1. / 35.0, "ccx": -0.6, "curvy": 1. / 60.0,
                               "ccy": 0.4, "coefficients": [(1e-5, 0.3), (-2e-8, -0.2)]}),
         "aperture": _circ(9.0), "mat": "glass", "opt": {}},
        _conic("back", 5.0, curv=-1. / 70.0, mat=None),
        _conic("image", 50.0),
    ],
    "materials": {"glass": ("ConstantIndexGlass", {"n": 1.58})},
    "bundle": {"rings": 6, "radius": 7.0, "z0": -2.0},
    "s_counted": 2,
}

X7_TWO_ELEMENTS = dict(C2_DOUBLEGAUSS, name="x7_two_elements", split_after=6,
                       bundle={"rings": 5, "radius": 5.0, "z0": 0.0})
# the double-Gauss cut into two OpticalElements in the air gap after s5: the element
# sequence has two entries, each element starts again in the background medium
# (optical_element.py:328) and the hand-over bundle appears twice in the path

X8_CRYSTAL_MIRROR = {   # reflection inside a birefringent medium
    # (material_anisotropic.py:115-155: modes [0], [1] of the Poynting sort, negated)
    "name": "x8_crystal_mirror",
    "surfaces": [
        _conic("stop", 0.0, opt={"is_stop": True}),
        _conic("front", 2.0, curv=1. / 45.0, mat="crystal", aperture=_circ(9.0)),
        _conic("mirror", 5.0, curv=-1. / 80.0, mat="crystal", opt={"is_mirror": True},
               tiltx=4.0 * math.pi / 180.0),
        _conic("exit", 6.0, curv=0.0, mat=None),
        _conic("image", 20.0),
    ],
    "materials": {"crystal": ("AnisotropicMaterial", {"epstensor": _EPS1})},
    "bundle": {"rings": 2, "radius": 5.0, "z0": -2.0},
    "s_counted": 3,
}

_ZF_SYM = [0.0] * 16       # rotationally symmetric Fringe series: Z4, Z9, Z16 (m = 0)
(_ZF_SYM[3], _ZF_SYM[8], _ZF_SYM[15]) = (-0.35, 0.02, -0.004)
X9_ZERNIKE = {
    "name": "x9_zernike",
    "surfaces": [
        _conic("stop", 0.0, opt={"is_stop": True}),
        _conic("front", 4.0, curv=1. / 70.0, mat="glass"),
        {"name": "back", "lc": {"decz": 6.0, "decx": 0.4, "tilty": 1.5 * math.pi / 180.0},
         "shape": ("ZernikeFringe", {"normradius": 10.0, "coefficients": _ZF_SYM}),
         "aperture": None, "mat": None, "opt": {}},
        _conic("image", 55.0),
    ],
    "materials": {"glass": ("ConstantIndexGlass", {"n": 1.62})},
    "bundle": {"rings": 6, "radius": 7.0, "z0": -3.0},
    "s_counted": 2,
}
_ZF_GEN = [0.0, 0.01, -0.02, -0.3, 0.03, -0.02, 0.015, -0.01, 0.02, 0.004, -0.006, 0.003,
           -0.002, 0.004, -0.003, -0.004]
X10_ZERNIKE_GENERAL = dict(X9_ZERNIKE, name="x10_zernike_general", surfaces=[
    X9_ZERNIKE["surfaces"][0], X9_ZERNIKE["surfaces"][1],
    dict(X9_ZERNIKE["surfaces"][2],
         shape=("ZernikeFringe", {"normradius": 10.0, "coefficients": _ZF_GEN})),
    X9_ZERNIKE["surfaces"][3]])
# x10 is traced against the oracle only (the reference's Zernike gradient is inconsistent
# with its sag for m != 0 terms, see oracle/pyrate_np.py:_zernike_term)

X11_GRIDSAG = {     # measured-surface data: sag on a 41 x 33 grid (GridSag, :861-924)
    "name": "x11_gridsag",
    "surfaces": [
        _conic("stop", 0.0, opt={"is_stop": True}),
        _conic("front", 4.0, curv=1. / 60.0, mat="glass"),
        {"name": "back", "lc": {"decz": 6.0, "decy": -0.3, "tiltx": 1.0 * math.pi / 180.0},
         "shape": ("GridSag", {"grid": {"x": (-9.0, 9.0, 41), "y": (-8.0, 8.0, 33),
                                        "poly": [(-0.011, 2, 0), (-0.009, 0, 2), (4e-4, 1, 1),
                                                 (2e-5, 3, 0), (-1.5e-5, 1, 2), (3e-6, 4, 0),
                                                 (2e-6, 2, 2), (-1e-6, 0, 4)]}}),
         "aperture": None, "mat": None, "opt": {}},
        _conic("image", 50.0),
    ],
    "materials": {"glass": ("ConstantIndexGlass", {"n": 1.55})},
    "bundle": {"rings": 6, "radius": 7.0, "z0": -3.0},
    "s_counted": 2,
}

_ZF_SYM2 = [0.0] * 9
(_ZF_SYM2[3], _ZF_SYM2[8]) = (0.05, -0.008)
X12_COMBINATION = {   # asphere + decentred XY polynomial + Zernike (m = 0): the shape the
    # Zemax importer builds for "Zernike fringe sag" surfaces (io/zmx.py:755-775)
    "name": "x12_combination",
    "surfaces": [
        _conic("stop", 0.0, opt={"is_stop": True}),
        {"name": "front", "lc": {"decz": 4.0},
         "shape": ("LinearCombination", {"terms": [
             (1.0, "Asphere", {"curv": 1. / 45.0, "cc": -0.8, "coefficients": [2e-6, -1e-9]},
              None),
             (0.5, "XYPolynomials", {"normradius": 10.0,
                                     "coefficients": [(2, 0, 0.02), (1, 1, -0.01),
                                                      (0, 3, 0.004)]},
              {"decx": 0.5, "decy": -0.25}),
             (2.0, "ZernikeFringe", {"normradius": 10.0, "coefficients": _ZF_SYM2},
              {"decx": -0.2, "decy": 0.1})]}),
         "aperture": _circ(9.0), "mat": "glass", "opt": {}},
        _conic("back", 5.0, curv=-1. / 80.0, mat=None),
        _conic("image", 45.0),
    ],
    "materials": {"glass": ("ConstantIndexGlass", {"n": 1.6})},
    "bundle": {"rings": 6, "radius": 7.0, "z0": -2.0},
    "s_counted": 2,
}

X13_TIRGLASS = {   # IsotropicMaterialTIR (angle-form Snell, material_isotropic_tir.py:46-118)
    # entered from a denser glass at a steep cemented surface: rim rays are totally
    # internally reflected (-> invalid, dropped); no ray misses a surface or an aperture
    "name": "x13_tirglass",
    "surfaces": [
        _conic("stop", 0.0, opt={"is_stop": True}),
        _conic("front", 2.0, curv=1. / 12.0, mat="dense", decx=-0.1),
        _conic("cement", 5.0, curv=-1. / 7.5, mat="light", decy=0.2),
        _conic("back", 4.0, curv=1. / 90.0, cc=-1.5, mat=None, tiltx=1.0 * math.pi / 180.0),
        _conic("image", 15.0),
    ],
    "materials": {"dense": ("ConstantIndexGlass", {"n": 1.9}),
                  "light": ("ConstantIndexGlassTIR", {"n": 1.31})},
    # The material frame (the cement surface's) is parallel to the global frame: the
    # reference's TIR refract takes k in GLOBAL and the normal in MATERIAL coordinates
    # (:52 vs :66) and is only right for such frames; the device uses one frame.
    # radius 4: no TIR (the reference's TIR refract raises ValueError as soon as a ray
    # is actually reflected, :116 mixes compacted and uncompacted widths); the TIR
    # branch is traced against the oracle with a wider bundle (tests/test_gpu_parity.py)
    "bundle": {"rings": 6, "radius": 4.0, "z0": -2.0},
    "s_counted": 3,
}

def _conrady(nd, spread):
    """Conrady coefficients (n0, A, B) with index nd at the d line and a dispersion
    scale `spread` (ModelGlass: n = n0 + A / wave + B / wave^3.5, wave in mm)."""
    (a, b) = (0.0100998734374e-3 * spread, 0.000328623343942 * (1e-3) ** 3.5 * spread)
    return (nd - a / DLINE - b / DLINE ** 3.5, a, b)


X14_DISPERSIVE = dict(C2_DOUBLEGAUSS, name="x14_dispersive", materials={
    # the double-Gauss with dispersive (Conrady) glasses: wavelength batches F / d / C
    "g1": ("ModelGlass", {"n0_A_B": _conrady(_N1, 1.0)}),
    "g2": ("ModelGlass", {"n0_A_B": _conrady(_N2, 1.6)}),
    "g3": ("ModelGlass", {"n0_A_B": _conrady(_N3, 2.2)})},
    bundle={"rings": 8, "radius": 5.0, "z0": 0.0})
X15_DISPERSIVE_ASPHERE = dict(C3_ASPHERE, name="x15_dispersive_asphere", materials={
    "glass": ("ModelGlass", {"n0_A_B": _conrady(1.5168, 1.3)})},
    bundle={"rings": 8, "radius": 11.43, "z0": -5.0})

X16_CYLINDER = {   # cylinder lenses (conic sections in y extruded along x), tilted about z so that
    # the extrusion axis is not a coordinate axis of the bundle, one parabolic and one
    # elliptic section; the rear cylinder sits in a crossed orientation
    "name": "x16_cylinder",
    "surfaces": [
        _conic("stop", 0.0, opt={"is_stop": True}),
        {"name": "front", "lc": {"decz": 3.0, "tiltz": 20.0 * math.pi / 180.0},
         "shape": ("Cylinder", {"curv": 1. / 30.0, "cc": -1.0}),
         "aperture": _circ(9.0), "mat": "glass", "opt": {}},
        {"name": "back", "lc": {"decz": 5.0, "tiltz": 70.0 * math.pi / 180.0,
                                "tiltx": 1.0 * math.pi / 180.0},
         "shape": ("Cylinder", {"curv": -1. / 45.0, "cc": 0.4}),
         "aperture": None, "mat": None, "opt": {}},
        _conic("image", 40.0),
    ],
    "materials": {"glass": ("ConstantIndexGlass", {"n": 1.5168})},
    "bundle": {"rings": 8, "radius": 7.0, "z0": -3.0},
    "s_counted": 2,
}

# A GRIN rod with an index profile OUTSIDE the device catalogue (hyperbolic-secant "selfoc"
# profile with an axial taper): Python source for the reference, CUDA expressions for the
# device (compiled at run time, pyrate_b200/grin_jit.py).
SECH_SOURCE = r"""
import numpy as np


def nfunc(x, **kw):
    return 1.55 / np.cosh(0.12 * np.sqrt(x[0]**2 + x[1]**2)) * (1.0 + 0.002 * x[2])


def dndx(x, **kw):
    r = np.sqrt(x[0]**2 + x[1]**2) + 1e-300
    return -1.55 * 0.12 * np.tanh(0.12 * r) / np.cosh(0.12 * r) * (1.0 + 0.002 * x[2]) * x[0] / r


def dndy(x, **kw):
    r = np.sqrt(x[0]**2 + x[1]**2) + 1e-300
    return -1.55 * 0.12 * np.tanh(0.12 * r) / np.cosh(0.12 * r) * (1.0 + 0.002 * x[2]) * x[1] / r


def dndz(x, **kw):
    return 1.55 / np.cosh(0.12 * np.sqrt(x[0]**2 + x[1]**2)) * 0.002


def bnd(x):
    return x[0]**2 + x[1]**2 < 6.0**2
"""

X17_USER_GRIN = {
    "name": "x17_user_grin",
    "surfaces": [
        _conic("object", 0.0, opt={"is_stop": True}),
        _conic("front", 8.0, curv=1. / 40.0, mat="rod", aperture=_circ(4.5)),
        _conic("back", 15.0, curv=-1. / 60.0, mat=None, tiltx=3. * math.pi / 180.0),
        _conic("image", 12.0),
    ],
    "materials": {"rod": ("IsotropicGrinMaterial", {
        "source": SECH_SOURCE,
        "names": ("nfunc", "dndx", "dndy", "dndz", "bnd"),
        "parameterlist": [],
        "ds": 0.05, "energyviolation": 0.01,
        "device_source": {
            "n": "p[0] / cosh(p[1] * sqrt(x * x + y * y)) * (1.0 + p[2] * z)",
            "dndx": "-p[0] * p[1] * tanh(p[1] * (sqrt(x * x + y * y) + 1e-300)) / cosh(p[1] * (sqrt(x * x + y * y) + 1e-300)) * (1.0 + p[2] * z) * x / (sqrt(x * x + y * y) + 1e-300)",
            "dndy": "-p[0] * p[1] * tanh(p[1] * (sqrt(x * x + y * y) + 1e-300)) / cosh(p[1] * (sqrt(x * x + y * y) + 1e-300)) * (1.0 + p[2] * z) * y / (sqrt(x * x + y * y) + 1e-300)",
            "dndz": "p[0] / cosh(p[1] * sqrt(x * x + y * y)) * p[2]",
            "inside": "x * x + y * y < 36.0",
            "params": [1.55, 0.12, 0.002]},
    })},
    "bundle": {"rings": 6, "radius": 3.5, "z0": -4.0},
    "s_counted": 2,
}

CONFIGS.update({c["name"]: c for c in (X17_USER_GRIN, X16_CYLINDER, X13_TIRGLASS, X14_DISPERSIVE, X15_DISPERSIVE_ASPHERE, X1_TILTED, X2_XYPOLY, X3_VIGNETTE, X4_BIAXIAL,
                                       X5_DEGENERATE, X6_BICONIC, X7_TWO_ELEMENTS,
                                       X8_CRYSTAL_MIRROR, X9_ZERNIKE, X10_ZERNIKE_GENERAL,
                                       X11_GRIDSAG, X12_COMBINATION)})
