"""Shared helpers of the parity tests."""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# tolerance classes of BASELINE.json's north_star
TOL_CLOSED_FORM = 1e-10      # isotropic conics: closed-form intersect + Snell
TOL_ITERATED = 1e-6          # asphere / XY polynomial Newton, GRIN integration


def golden_traces():
    return sorted(os.path.basename(f)[len("seqtrace_"):-4]
                  for f in glob.glob(os.path.join(GOLDEN, "seqtrace_*.npz")))


def load_golden(tag):
    return np.load(os.path.join(GOLDEN, "seqtrace_%s.npz" % tag))


def config_of(tag):
    from pyrate_b200 import configs
    names = sorted(configs.CONFIGS, key=len, reverse=True)
    for nm in names:
        if tag.startswith(nm):
            return nm
    raise KeyError(tag)


def tolerance_of(name):
    return TOL_ITERATED if name in ("c3_asphere", "c5_grin", "x2_xypoly", "x6_biconic",
                                    "x9_zernike", "x10_zernike_general", "x11_gridsag",
                                    "x12_combination", "x15_dispersive_asphere") \
        else TOL_CLOSED_FORM


def relerr(a, b):
    """max |a-b| / max |b| over the finite entries of b."""
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    m = np.isfinite(b)
    if not m.any():
        return 0.0
    return float(np.max(np.abs(a[m] - b[m])) / max(1e-300, np.max(np.abs(b[m]))))


def golden_paths(g):
    """Golden npz -> list of paths -> list of bundle dicts (subset of rows for
    long GRIN histories: rows 0, P-2, P-1)."""
    paths = []
    for ip in range(int(g["npaths"])):
        bundles = []
        for ib in range(int(g["p%d_nbundles" % ip])):
            pre = "p%d_b%d_" % (ip, ib)
            b = {"rows": int(g[pre + "rows"]), "x": g[pre + "x"], "k": g[pre + "k"],
                 "valid": g[pre + "valid"], "rayID": g[pre + "rayID"]}
            if pre + "E" in g:
                b["E"] = g[pre + "E"]
            bundles.append(b)
        paths.append(bundles)
    return paths


def compare_bundle(got, ref, tol, what, first_last_only=False, check_k_last=True):
    """got/ref: dicts with x, k (P,3,N), valid (P,N), rayID (N).  `ref` rows may
    be the [0, P-2, P-1] subset of a long history (then got has 2 rows)."""
    assert np.array_equal(np.asarray(got["rayID"]), np.asarray(ref["rayID"])), \
        "%s: rayID differs" % what
    gx, rx = np.asarray(got["x"]), np.asarray(ref["x"])
    gk, rk = np.asarray(got["k"]), np.asarray(ref["k"])
    gv, rv = np.asarray(got["valid"]), np.asarray(ref["valid"])
    if gx.shape[0] != rx.shape[0] or first_last_only:
        sel_g = [0, gx.shape[0] - 1]
        sel_r = [0, rx.shape[0] - 1]
        (gx, gk, gv) = (gx[sel_g], gk[sel_g], gv[sel_g])
        (rx, rk, rv) = (rx[sel_r], rk[sel_r], rv[sel_r])
        if not check_k_last:
            (gk, rk) = (gk[:1], rk[:1])
    assert np.array_equal(gv, rv), "%s: valid rows differ (%d vs %d set)" % (
        what, int(gv.sum()), int(rv.sum()))
    v = rv[-1].astype(bool)
    ex = relerr(gx[:, :, v], rx[:, :, v])
    ek = relerr(gk[:, :, v], rk[:, :, v])
    assert ex <= tol, "%s: x rel err %.3e > %.1e" % (what, ex, tol)
    assert ek <= tol, "%s: k rel err %.3e > %.1e" % (what, ek, tol)
    return max(ex, ek)


def compare_birefringent_bundle(got, ref, tol, what, check_e=True):
    """Bundles downstream of a birefringent interface.  Children of one parent
    ray share its rayID; the order of the two forward modes depends on LAPACK's
    eigenvector normalisation in the reference (material.py:148 sorts S.n of
    un-normalised fields), so children are matched as an unordered set per
    rayID; E is compared up to a complex scalar."""
    gid, rid = np.asarray(got["rayID"]), np.asarray(ref["rayID"])
    assert np.array_equal(np.sort(gid), np.sort(rid)), "%s: rayID multiset differs" % what
    gx, rx = np.asarray(got["x"]), np.asarray(ref["x"])
    gk, rk = np.asarray(got["k"]), np.asarray(ref["k"])
    assert gx.shape == rx.shape and gk.shape == rk.shape, (what, gx.shape, rx.shape)
    gv, rv = np.asarray(got["valid"]), np.asarray(ref["valid"])
    ge = np.asarray(got["E"]) if "E" in got else None
    re_ = np.asarray(ref["E"]) if "E" in ref else None
    scale_x = max(1e-300, np.nanmax(np.abs(rx)))
    scale_k = max(1e-300, np.nanmax(np.abs(rk)))
    worst = 0.0
    for ray in np.unique(rid):
        gc = np.where(gid == ray)[0]
        rc = list(np.where(rid == ray)[0])
        assert len(gc) == len(rc)
        for g in gc:
            feats = [np.max(np.abs(gk[0][:, g] - rk[0][:, r])) / scale_k +
                     np.max(np.abs(gx[-1][:, g] - rx[-1][:, r])) / scale_x for r in rc]
            j = int(np.argmin(feats))
            r = rc.pop(j)
            ex = np.max(np.abs(gx[:, :, g] - rx[:, :, r])) / scale_x
            ek = np.max(np.abs(gk[:, :, g] - rk[:, :, r])) / scale_k
            assert ex <= tol and ek <= tol, "%s ray %d: x %.3e k %.3e" % (what, ray, ex, ek)
            assert np.array_equal(gv[:, g], rv[:, r]), "%s ray %d: valid differs" % (what, ray)
            worst = max(worst, ex, ek)
            if check_e and ge is not None and re_ is not None:
                (a, b) = (ge[0][:, g], re_[0][:, r])
                col = np.abs(np.sum(np.conj(a) * b)) / np.sqrt(
                    np.sum(np.abs(a) ** 2) * np.sum(np.abs(b) ** 2))
                assert col > 1 - 1e-7, "%s ray %d: E not colinear (%.3e)" % (what, ray, 1 - col)
    return worst


def random_spec(seed, explicit=False):
    """A random but traceable chain: decentered / tilted conics (both tilt orders),
    random glasses, apertures of all three kinds, an occasional mirror; with `explicit`
    also a mild even asphere or XY polynomial in place of some conics."""
    from pyrate_b200 import configs
    rng = np.random.default_rng(seed)
    surfaces = [configs._conic("stop", 0.0, opt={"is_stop": True})]
    mats = {}
    inside = False
    nsurf = int(rng.integers(3, 7))
    for i in range(nsurf):
        lc = {"decx": float(rng.uniform(-0.3, 0.3)), "decy": float(rng.uniform(-0.3, 0.3)),
              "tiltx": float(rng.uniform(-0.05, 0.05)), "tilty": float(rng.uniform(-0.05, 0.05)),
              "tiltz": float(rng.uniform(-1.0, 1.0)), "tiltThenDecenter": int(rng.integers(0, 2))}
        mat = None
        if not inside or rng.random() < 0.4:
            mat = "g%d" % i
            mats[mat] = ("ConstantIndexGlass", {"n": float(rng.uniform(1.3, 1.9))})
        inside = mat is not None
        kind = int(rng.integers(0, 3))
        ap = None if kind == 0 else (configs._circ(float(rng.uniform(6.0, 9.0))) if kind == 1 else
                                     ("RectangularAperture", {"width": float(rng.uniform(9, 14)),
                                                              "height": float(rng.uniform(9, 14))}))
        surf = configs._conic("s%d" % i, float(rng.uniform(2.0, 6.0)),
                              curv=float(rng.uniform(-0.03, 0.03)),
                              cc=float(rng.choice([0.0, -1.0, float(rng.uniform(-2, 2))])),
                              mat=mat, aperture=ap, **lc)
        if explicit and rng.random() < 0.4:
            if rng.random() < 0.5:
                surf["shape"] = ("Asphere", {"curv": float(rng.uniform(-0.03, 0.03)),
                                             "cc": float(rng.uniform(-1.5, 0.5)),
                                             "coefficients": [float(rng.uniform(-1e-4, 1e-4)),
                                                              float(rng.uniform(-1e-6, 1e-6)),
                                                              float(rng.uniform(-1e-9, 1e-9))]})
            else:
                surf["shape"] = ("XYPolynomials", {"normradius": 10.0, "coefficients": [
                    (2, 0, float(rng.uniform(-0.8, 0.8))), (0, 2, float(rng.uniform(-0.8, 0.8))),
                    (1, 1, float(rng.uniform(-0.05, 0.05))), (3, 0, float(rng.uniform(-0.02, 0.02))),
                    (2, 2, float(rng.uniform(-0.01, 0.01)))]})
        surfaces.append(surf)
    if inside:
        surfaces.append(configs._conic("exit", 3.0, curv=float(rng.uniform(-0.01, 0.01)), mat=None))
    if rng.random() < 0.5:
        surfaces.append(configs._conic("mirror", 10.0, curv=float(rng.uniform(-0.005, 0.005)),
                                       opt={"is_mirror": True}, tiltx=float(rng.uniform(-0.1, 0.1))))
    surfaces.append(configs._conic("image", 15.0))
    return {"name": "random%d" % seed, "surfaces": surfaces, "materials": mats,
            "bundle": {"rings": 9, "radius": float(rng.uniform(4.0, 7.0)), "z0": -3.0}}
