"""CPU: the NumPy restatement (oracle/pyrate_np.py) against the fixtures dumped
from the unmodified reference (oracle/gen_golden.py) -- this is what pins the
oracle.  Bit-level agreement is expected wherever the arithmetic is the same
NumPy expression; the Newton-vs-fsolve and eig-ordering cases get tolerances."""
import os

import numpy as np
import pytest

import pyrate_np as onp
from pyrate_b200 import configs

import util


@pytest.mark.parametrize("tag", util.golden_traces())
def test_seqtrace_matches_reference(tag):
    g = util.load_golden(tag)
    name = util.config_of(tag)
    system = onp.system_from_spec(configs.CONFIGS[name])
    paths = onp.seqtrace(system, g["x0"], g["k0"], g["E0"], wave=configs.DLINE,
                         splitup=bool(g["splitup"]))
    ref_paths = util.golden_paths(g)
    assert len(paths) == len(ref_paths)
    tol = 1e-11 if util.tolerance_of(name) == util.TOL_ITERATED else 1e-13
    for (ip, (path, rpath)) in enumerate(zip(paths, ref_paths)):
        assert len(path) == len(rpath)
        assert path[0] is path[1]
        for (ib, (b, rb)) in enumerate(zip(path, rpath)):
            assert b["x"].shape[0] == rb["rows"]
            got = {"x": b["x"], "k": b["k"], "valid": b["valid"], "rayID": b["rayID"]}
            if rb["rows"] > 3:
                sel = [0, rb["rows"] - 2, rb["rows"] - 1]
                got = {"x": b["x"][sel], "k": b["k"][sel], "valid": b["valid"][sel],
                       "rayID": b["rayID"]}
            util.compare_bundle(got, rb, tol, "%s p%d b%d" % (tag, ip, ib))


def test_frames_match_reference():
    g = np.load(util.GOLDEN + "/frames.npz")
    for i in range(g["params"].shape[0]):
        frame = onp.ROOT_FRAME
        for lvl in range(3):
            (dx, dy, dz, tx, ty, tz, ttd) = g["params"][i, lvl]
            frame = onp.child_frame(frame, decx=dx, decy=dy, decz=dz, tiltx=tx,
                                    tilty=ty, tiltz=tz, tiltThenDecenter=int(ttd))
            assert np.allclose(frame[0], g["basis"][i, lvl], rtol=0, atol=1e-14)
            assert np.allclose(frame[1], g["origin"][i, lvl], rtol=0, atol=1e-13)
    pts = g["pts"]
    assert np.allclose(onp.l2g_pts(frame, pts), g["last_l2g_pts"], atol=1e-13)
    assert np.allclose(onp.g2l_pts(frame, pts), g["last_g2l_pts"], atol=1e-13)
    assert np.allclose(onp.l2g_dir(frame, pts), g["last_l2g_dir"], atol=1e-13)
    assert np.allclose(onp.g2l_dir(frame, pts), g["last_g2l_dir"], atol=1e-13)


def test_shapes_match_reference():
    g = np.load(util.GOLDEN + "/shapes.npz")
    (x, y) = (g["x"], g["y"])
    for (i, (curv, cc)) in enumerate(g["conic_params"]):
        sh = {"kind": "Conic", "curv": curv, "cc": cc}
        assert np.allclose(onp.shape_sag(sh, x, y), g["conic%d_sag" % i],
                           rtol=1e-14, atol=0, equal_nan=True)
        assert np.allclose(onp.shape_grad(sh, x, y), g["conic%d_grad" % i],
                           rtol=1e-14, atol=1e-16, equal_nan=True)
        assert np.allclose(onp.shape_normal(sh, x, y), g["conic%d_normal" % i],
                           rtol=1e-14, atol=1e-16, equal_nan=True)
    ap = g["asph_params"]
    sh = {"kind": "Asphere", "curv": ap[0], "cc": ap[1], "coefficients": list(ap[2:])}
    assert np.allclose(onp.shape_sag(sh, x, y), g["asph_sag"], rtol=1e-14)
    assert np.allclose(onp.shape_grad(sh, x, y), g["asph_grad"], rtol=1e-13, atol=1e-16)
    for (nm, kind) in (("zf", "ZernikeFringe"), ("za", "ZernikeANSI")):
        sh = {"kind": kind, "normradius": 5.0, "coefficients": list(g[nm + "_coeffs"])}
        assert np.allclose(onp.shape_sag(sh, x, y), g[nm + "_sag"], rtol=1e-12, atol=1e-15)
    bp = g["bic_params"]
    sh = {"kind": "Biconic", "curvx": bp[0], "ccx": bp[1], "curvy": bp[2], "ccy": bp[3],
          "coefficients": [(bp[4], bp[5]), (bp[6], bp[7])]}
    assert np.allclose(onp.shape_sag(sh, x, y), g["bic_sag"], rtol=1e-14, atol=1e-18)
    assert np.allclose(onp.shape_grad(sh, x, y), g["bic_grad"], rtol=1e-13, atol=1e-16)
    sh = {"kind": "XYPolynomials", "normradius": float(g["xy_normradius"]),
          "coefficients": [tuple(c) for c in g["xy_coeffs"]]}
    assert np.allclose(onp.shape_sag(sh, x, y), g["xy_sag"], rtol=1e-13, atol=1e-16)
    assert np.allclose(onp.shape_grad(sh, x, y), g["xy_grad"], rtol=1e-13, atol=1e-16)
    assert np.allclose(onp.shape_normal(sh, x, y), g["xy_normal"], rtol=1e-13, atol=1e-16)


def test_gridsag_and_combination_match_reference():
    """GridSag (FITPACK evaluation restated in NumPy) and LinearCombination sag /
    gradient against the unmodified reference, and the restated spline evaluation
    against SciPy's own ev() including points outside the grid (clamped)."""
    from pyrate_b200 import configs
    g = np.load(os.path.join(util.GOLDEN, "shapes2.npz"))
    (x, y) = (g["x"], g["y"])
    grid = configs.X11_GRIDSAG["surfaces"][2]["shape"][1]["grid"]
    (xl, yl, zg) = configs.grid_arrays(grid)
    sh = {"kind": "GridSag", "xlinspace": xl, "ylinspace": yl, "zgrid": zg,
          "frame": onp.ROOT_FRAME}
    assert util.relerr(onp.shape_sag(sh, x, y), g["grid_sag"]) < 1e-13
    assert util.relerr(onp.shape_grad(sh, x, y), g["grid_grad"]) < 1e-12
    from scipy.interpolate import RectBivariateSpline
    sp = RectBivariateSpline(xl, yl, zg)
    rng = np.random.default_rng(5)
    (px, py) = (rng.uniform(-12, 12, 500), rng.uniform(-11, 11, 500))
    (f, fx, fy) = onp.gridsag_eval(onp.gridsag_fit(xl, yl, zg), px, py)
    assert np.max(np.abs(f - sp.ev(px, py))) < 1e-13
    assert np.max(np.abs(fx - sp.ev(px, py, dx=1))) < 1e-12
    assert np.max(np.abs(fy - sp.ev(px, py, dy=1))) < 1e-12
    sub = onp.child_frame(onp.ROOT_FRAME, decx=0.5, decy=-0.25)
    comb = {"kind": "LinearCombination", "frame": onp.ROOT_FRAME, "terms": [
        (1.0, {"kind": "Asphere", "curv": 1. / 45.0, "cc": -0.8,
               "coefficients": [2e-6, -1e-9], "frame": onp.ROOT_FRAME}),
        (0.5, {"kind": "XYPolynomials", "normradius": 10.0,
               "coefficients": [(2, 0, 0.02), (1, 1, -0.01), (0, 3, 0.004)],
               "frame": sub})]}
    assert util.relerr(onp.shape_sag(comb, g["xs"], g["ys"]), g["comb_sag"]) < 1e-13
    assert util.relerr(onp.shape_grad(comb, g["xs"], g["ys"]), g["comb_grad"]) < 1e-13


def test_aniso_modes_match_reference():
    g = np.load(util.GOLDEN + "/aniso_modes.npz")
    for nm in ("uniaxial_y", "biaxial_rot"):
        (k4, e4) = onp.aniso_modes_sorted(g[nm + "_eps"], g[nm + "_n"], g[nm + "_kpa"])
        assert np.allclose(k4, g[nm + "_k4"], rtol=1e-10, atol=1e-12)
        # eigenvectors are defined up to a complex scalar: compare directions
        ref = g[nm + "_e4"]
        num = np.abs(np.sum(np.conj(e4) * ref, axis=1))
        den = np.sqrt(np.sum(np.abs(e4) ** 2, axis=1) * np.sum(np.abs(ref) ** 2, axis=1))
        assert np.all(num / den > 1 - 1e-9)


def test_spot_matches_reference():
    g = np.load(util.GOLDEN + "/spot.npz")
    c = onp.centroid(g["x"])
    assert np.allclose(c, g["centroid"], rtol=1e-15, atol=1e-18)
    assert np.isclose(onp.rms_spot(g["x"], c), float(g["rms"]), rtol=1e-14)
    assert np.isclose(onp.rms_spot(g["x"], np.zeros(3)), float(g["rms0"]), rtol=1e-14)


def test_known_answer_seed_vector():
    """double-Gauss ray 0 at the image plane."""
    system = onp.system_from_spec(configs.CONFIGS["c2_doublegauss"])
    x0 = np.array([[0.0], [5.0], [0.0]])
    k0 = np.array([[0.0], [0.0], [1.0]])
    e0 = np.array([[1.0], [0.0], [0.0]])
    path = onp.seqtrace(system, x0, k0, e0)[0]
    assert np.allclose(path[-1]["x"][0][:, 0],
                       [0.0, -0.5212471710809123, 172.54859051823328], rtol=1e-13)
    assert np.allclose(path[-1]["k"][0][:, 0],
                       [0.0, -0.04314208775168256, 0.9990689467020913], rtol=1e-13)
    assert np.allclose(path[3]["k"][0][:, 0],
                       [0.0, -0.06039952118111261, 1.522259388291602], rtol=1e-13)


def test_tir_glass_angle_form_equals_vector_form():
    """material_isotropic_tir.py:46-118 (angles) against material_isotropic.py:163-199
    (k_par + xi n) on a bundle where rim rays are totally reflected."""
    import copy
    spec = copy.deepcopy(configs.CONFIGS["x13_tirglass"])
    spec["bundle"]["radius"] = 7.0
    deg = np.pi / 180.0
    (x0, k0, e0) = configs.config_bundle(spec, 10, (0., np.sin(deg), np.cos(deg)), (1., 0., 0.))
    a = onp.seqtrace(onp.system_from_spec(spec), x0, k0, e0, wave=configs.DLINE)[0]
    spec["materials"]["light"] = ("ConstantIndexGlass", spec["materials"]["light"][1])
    b = onp.seqtrace(onp.system_from_spec(spec), x0, k0, e0, wave=configs.DLINE)[0]
    assert a[-1]["x"].shape[2] < x0.shape[1]
    for (ba, bb) in zip(a, b):
        util.compare_bundle(ba, bb, 1e-13, "tir forms")


def _live_reference_api():
    import refshim
    if not refshim.reference_available():
        pytest.skip("reference tree not present (build container only)")
    return refshim.api()


@pytest.mark.parametrize("seed", list(range(16)))
def test_oracle_matches_live_reference_on_random_systems(seed):
    """Beyond the committed fixtures: the restatement against the UNMODIFIED reference,
    imported live from /root/reference (build container only; skipped elsewhere), on
    random decentred / tilted systems with all aperture kinds, mirrors and -- every
    other seed -- mild aspheres / XY polynomials, for a random oblique field."""
    api = _live_reference_api()
    explicit = seed % 2 == 1
    spec = util.random_spec(seed, explicit=explicit)
    rng = np.random.default_rng(500 + seed)
    (ax, ay) = rng.uniform(-0.03, 0.03, 2)
    kdir = (np.sin(ax), np.sin(ay) * np.cos(ax), np.cos(ay) * np.cos(ax))
    (x0, k0, e0) = configs.config_bundle(spec, 4, kdir, (0., 1., 0.))
    (s, seq) = configs.build_system(spec, api)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref_paths = s.seqtrace(api.RayBundle(x0.copy(), k0.copy(), e0.copy(),
                                             wave=configs.DLINE), seq)
        paths = onp.seqtrace(onp.system_from_spec(spec), x0, k0, e0, wave=configs.DLINE)
    assert len(paths) == len(ref_paths) == 1
    tol = 1e-9 if explicit else 1e-12
    assert len(paths[0]) == len(ref_paths[0].raybundles)
    for (ib, (b, rb)) in enumerate(zip(paths[0], ref_paths[0].raybundles)):
        ref = {"x": np.asarray(rb.x), "k": np.asarray(rb.k), "valid": np.asarray(rb.valid),
               "rayID": np.asarray(rb.rayID)}
        util.compare_bundle(b, ref, tol, "seed %d b%d" % (seed, ib))


def test_uniaxial_closed_form_roots_match_reference_eigenvalues():
    """The factorised dispersion relation of a uniaxial crystal against the xi eigenvalues
    of the reference (MaxwellMaterial.calcXiEigenvectorsNorm, material.py:407-454, fixture
    aniso_modes.npz), and against the LAPACK restatement for a rotated axis and oblique
    complex in-plane vectors."""
    g = np.load(util.GOLDEN + "/aniso_modes.npz")
    dec = onp.uniaxial_decomposition(g["uniaxial_y_eps"])
    assert dec is not None and np.allclose(np.abs(dec[2]), [0, 1, 0])
    assert np.isclose(dec[0], 1.658 ** 2) and np.isclose(dec[1], 1.486 ** 2)
    roots = onp.uniaxial_xi_roots(*dec, g["uniaxial_y_n"], g["uniaxial_y_kpa"])
    ref = g["uniaxial_y_xi4"]
    for j in range(ref.shape[1]):
        assert np.allclose(np.sort_complex(roots[:, j]), np.sort_complex(ref[:, j]),
                           rtol=1e-10, atol=1e-12), j
    assert onp.uniaxial_decomposition(g["biaxial_rot_eps"]) is None
    assert onp.uniaxial_decomposition(2.25 * np.eye(3)) is None
    # rotated axis, complex kpa: every root must satisfy the full dispersion relation
    rng = np.random.default_rng(9)
    axis = rng.normal(size=3)
    axis /= np.linalg.norm(axis)
    (eo, ee) = (2.4, 2.9)
    eps = eo * np.eye(3) + (ee - eo) * np.outer(axis, axis)
    dec = onp.uniaxial_decomposition(eps)
    assert np.isclose(dec[0], eo) and np.isclose(dec[1], ee) and np.isclose(abs(dec[2] @ axis), 1.0)
    n = rng.normal(size=(3, 20))
    n /= np.linalg.norm(n, axis=0)
    kin = rng.normal(size=(3, 20)) * 0.4 + 0.05j * rng.normal(size=(3, 20))
    kpa = kin - np.sum(kin * n, axis=0) * n
    roots = onp.uniaxial_xi_roots(*dec, n, kpa)
    for r in range(4):
        k = kpa + roots[r] * n
        for j in range(20):
            kj = k[:, j]
            m = eps - np.dot(kj, kj) * np.eye(3) + np.outer(kj, kj)      # (eps - k.k 1 + k k^T) E = 0
            assert abs(np.linalg.det(m)) < 1e-10, (r, j)


def test_oracle_matches_live_reference_below_the_critical_angles_of_a_crystal():
    """A ray fan through a dense glass -> tilted calcite-like interface that straddles both
    critical angles (the bundle of the GPU test of the crystal kernel's real / complex paths).
    Rays whose four modes propagate agree with the UNMODIFIED reference to rounding.  Rays with
    evanescent modes cannot be pinned: their Poynting sort keys S.n vanish, so WHICH two of the
    four complex roots material.py:122-153 keeps is decided by LAPACK's rounding (the selection
    differs between the reference and any other solver, including this oracle); what both
    engines agree on there is that the ray stays in the bundle (anisotropic refract has no
    validity filter, material_anisotropic.py:87-100)."""
    api = _live_reference_api()
    import warnings
    import test_gpu_parity as tg
    spec = tg._evanescent_spec()
    n = 61
    rng = np.random.default_rng(5)
    ang = np.linspace(-0.43, 0.43, n)
    tx = rng.uniform(-0.02, 0.02, n)
    k0 = np.stack((np.sin(tx), np.cos(tx) * np.sin(ang), np.cos(tx) * np.cos(ang)))
    x0 = np.stack((rng.uniform(-0.5, 0.5, n), rng.uniform(-0.5, 0.5, n), np.full(n, -3.0)))
    e0 = np.cross(k0.T, np.array([0.0, 1.0, 0.3])).T
    e0 /= np.linalg.norm(e0, axis=0)
    (s, seq) = configs.build_system(spec, api)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = s.seqtrace(api.RayBundle(x0.copy(), k0.copy(), e0.copy(), wave=configs.DLINE), seq)[0].raybundles
        got = onp.seqtrace(onp.system_from_spec(spec), x0, k0, e0, wave=configs.DLINE)[0]
    assert len(got) == len(ref)
    (b, rb) = (got[4], ref[4])                       # the bundle born at the crystal surface
    (gk, rk) = (np.asarray(b["k"])[-1], np.asarray(rb.k)[-1])
    (gid, rid) = (np.asarray(b["rayID"]), np.asarray(rb.rayID))
    assert gk.shape == rk.shape == (3, 2 * n) and np.iscomplexobj(rk)
    assert np.asarray(rb.valid).all() and np.asarray(b["valid"]).all()      # nobody is dropped here
    (propagating, evanescent) = (0, 0)
    for ray in range(n):
        (ko, kr) = (gk[:, gid == ray], rk[:, rid == ray])
        if np.abs(kr.imag).max() > 1e-9 or np.abs(ko.imag).max() > 1e-9:
            evanescent += 1
            continue
        propagating += 1
        d = min(max(np.abs(ko[:, 0] - kr[:, 0]).max(), np.abs(ko[:, 1] - kr[:, 1]).max()),
                max(np.abs(ko[:, 0] - kr[:, 1]).max(), np.abs(ko[:, 1] - kr[:, 0]).max()))
        assert d < 1e-12, (ray, d)
    assert propagating >= 40 and evanescent >= 10
