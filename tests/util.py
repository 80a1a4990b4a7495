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
    return TOL_ITERATED if name in ("c3_asphere", "c5_grin", "x2_xypoly") \
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
