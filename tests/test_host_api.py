"""CPU: the host API mirror (frames, shapes, builders, lowering) against the
reference fixtures, and -- when the reference tree is present (build container)
-- the lowering of the reference's OWN object graph against ours."""
import ctypes

import numpy as np
import pytest

import pyrate_b200 as pb
from pyrate_b200 import _native as nat
from pyrate_b200 import configs, lowering

import util


def test_frames_match_reference_fixture():
    g = np.load(util.GOLDEN + "/frames.npz")
    for i in range(g["params"].shape[0]):
        parent = None
        for lvl in range(3):
            (dx, dy, dz, tx, ty, tz, ttd) = g["params"][i, lvl]
            lc = pb.LocalCoordinates.p(name="l%d_%d" % (i, lvl), decx=dx, decy=dy, decz=dz,
                                       tiltx=tx, tilty=ty, tiltz=tz, tiltThenDecenter=int(ttd))
            if parent is not None:
                parent.addChild(lc)
            parent = lc
            assert np.allclose(lc.localbasis, g["basis"][i, lvl], atol=1e-14)
            assert np.allclose(lc.globalcoordinates, g["origin"][i, lvl], atol=1e-13)
    pts = g["pts"]
    assert np.allclose(parent.returnLocalToGlobalPoints(pts), g["last_l2g_pts"], atol=1e-13)
    assert np.allclose(parent.returnGlobalToLocalPoints(pts), g["last_g2l_pts"], atol=1e-13)
    assert np.allclose(parent.returnLocalToGlobalDirections(pts), g["last_l2g_dir"], atol=1e-13)
    assert np.allclose(parent.returnGlobalToLocalDirections(pts), g["last_g2l_dir"], atol=1e-13)


def test_frame_round_trips_and_invariants():
    """Same properties the reference pins with hypothesis
    (tests/test_localcoordinates.py:63-183): round trips, scalar products and
    tensor contractions are frame independent."""
    rng = np.random.default_rng(3)
    for _ in range(20):
        kw = dict(zip(("decx", "decy", "decz", "tiltx", "tilty", "tiltz"),
                      rng.uniform(-3, 3, 6)))
        root = pb.LocalCoordinates.p(name="root", **kw)
        lc = root.addChild(pb.LocalCoordinates.p(name="child", tiltThenDecenter=1,
                                                 **dict(zip(kw, rng.uniform(-3, 3, 6)))))
        (a, b) = (rng.normal(size=(3, 5)), rng.normal(size=(3, 5)))
        t = rng.normal(size=(3, 3, 5))
        assert np.allclose(lc.returnGlobalToLocalPoints(lc.returnLocalToGlobalPoints(a)), a)
        assert np.allclose(lc.returnLocalToGlobalDirections(lc.returnGlobalToLocalDirections(a)), a)
        assert np.allclose(lc.returnGlobalToLocalTensors(lc.returnLocalToGlobalTensors(t)), t)
        (ag, bg) = (lc.returnLocalToGlobalDirections(a), lc.returnLocalToGlobalDirections(b))
        assert np.allclose(np.sum(ag * bg, axis=0), np.sum(a * b, axis=0))
        tg = lc.returnLocalToGlobalTensors(t)
        assert np.allclose(np.einsum("in,ijn,jn->n", ag, tg, bg),
                           np.einsum("in,ijn,jn->n", a, t, b))
        assert np.allclose(root.returnOtherToActualPoints(
            root.returnActualToOtherPoints(a, lc), lc), a)


def test_shapes_match_reference_fixture():
    g = np.load(util.GOLDEN + "/shapes.npz")
    lc = pb.LocalCoordinates.p(name="s")
    (x, y) = (g["x"], g["y"])
    for (i, (curv, cc)) in enumerate(g["conic_params"]):
        sh = pb.Conic.p(lc, curv=curv, cc=cc)
        assert np.allclose(sh.getSag(x, y), g["conic%d_sag" % i], rtol=1e-14, equal_nan=True)
        assert np.allclose(sh.getGrad(x, y), g["conic%d_grad" % i], rtol=1e-14, atol=1e-16,
                           equal_nan=True)
        assert np.allclose(sh.getNormal(x, y), g["conic%d_normal" % i], rtol=1e-14, atol=1e-16,
                           equal_nan=True)
    ap = g["asph_params"]
    sh = pb.Asphere.p(lc, curv=ap[0], cc=ap[1], coefficients=list(ap[2:]))
    assert np.allclose(sh.getSag(x, y), g["asph_sag"], rtol=1e-14)
    assert np.allclose(sh.getGrad(x, y), g["asph_grad"], rtol=1e-13, atol=1e-16)
    assert np.allclose(sh.getNormal(x, y), g["asph_normal"], rtol=1e-13, atol=1e-16)
    for (nm, cls) in (("zf", pb.ZernikeFringe), ("za", pb.ZernikeANSI)):
        zs = cls.p(lc, normradius=5.0, coefficients=list(g[nm + "_coeffs"]))
        assert np.allclose(zs.getSag(x, y), g[nm + "_sag"], rtol=1e-12, atol=1e-15)
        # gradient: consistent with the sag (central differences), which the
        # reference's polar formula is not for m != 0 (surface_shape.py:1085-1095)
        h = 1e-6
        gnum = -(zs.getSag(x + h, y) - zs.getSag(x - h, y)) / (2 * h)
        assert np.allclose(zs.getGrad(x, y)[0], gnum, rtol=1e-6, atol=1e-8)
    bp = g["bic_params"]
    sh = pb.Biconic.p(lc, curvx=bp[0], ccx=bp[1], curvy=bp[2], ccy=bp[3],
                      coefficients=[(bp[4], bp[5]), (bp[6], bp[7])])
    assert np.allclose(sh.getSag(x, y), g["bic_sag"], rtol=1e-13, atol=1e-16)
    assert np.allclose(sh.getGrad(x, y), g["bic_grad"], rtol=1e-12, atol=1e-15)
    assert np.allclose(sh.getNormal(x, y), g["bic_normal"], rtol=1e-12, atol=1e-15)
    sh = pb.XYPolynomials.p(lc, normradius=float(g["xy_normradius"]),
                            coefficients=[(int(a), int(b), c) for (a, b, c) in g["xy_coeffs"]])
    assert np.allclose(sh.getSag(x, y), g["xy_sag"], rtol=1e-13, atol=1e-16)
    assert np.allclose(sh.getGrad(x, y), g["xy_grad"], rtol=1e-13, atol=1e-16)
    # same evaluators on torch tensors
    import torch
    (xt, yt) = (torch.from_numpy(x), torch.from_numpy(y))
    assert np.allclose(sh.getNormal(xt, yt).numpy(), g["xy_normal"], rtol=1e-13, atol=1e-16)


def test_gridsag_and_combination_mirrors_match_reference_fixture():
    g = np.load(util.GOLDEN + "/shapes2.npz")
    lc = pb.LocalCoordinates.p(name="s2")
    grid = configs.X11_GRIDSAG["surfaces"][2]["shape"][1]["grid"]
    gs = pb.GridSag.p(lc, configs.grid_arrays(grid))
    assert gs.kind == "shape_GridSag"
    assert np.allclose(gs.getSag(g["x"], g["y"]), g["grid_sag"], rtol=1e-14, atol=1e-16)
    assert np.allclose(gs.getGrad(g["x"], g["y"]), g["grid_grad"], rtol=1e-13, atol=1e-16)
    lcd = pb.LocalCoordinates.p(name="s2_dec", decx=0.5, decy=-0.25)
    lc.addChild(lcd)
    asph = pb.Asphere.p(lc, curv=1. / 45.0, cc=-0.8, coefficients=[2e-6, -1e-9])
    xyp = pb.XYPolynomials.p(lcd, normradius=10.0,
                             coefficients=[(2, 0, 0.02), (1, 1, -0.01), (0, 3, 0.004)])
    comb = pb.LinearCombination.p(lc, list_of_coefficients_and_shapes=[(1.0, asph), (0.5, xyp)])
    assert comb.kind == "shape_LinearCombination"
    assert np.allclose(comb.getSag(g["xs"], g["ys"]), g["comb_sag"], rtol=1e-13, atol=1e-16)
    assert np.allclose(comb.getGrad(g["xs"], g["ys"]), g["comb_grad"], rtol=1e-13, atol=1e-16)
    # lowering: FITPACK arrays for the grid, one term record per sub-shape
    from pyrate_b200 import _native as nat, lowering

    class _S(object):
        pass
    (s1, s2) = (_S(), _S())
    (s1.shape, s1.aperture) = (gs, pb.BaseAperture.p(lc))
    (s2.shape, s2.aperture) = (comb, pb.BaseAperture.p(lc))
    st = nat.PyrStep()
    lowering.lower_surface(s1, st)
    assert st.shape_kind == nat.SHAPE_GRIDSAG
    (tx, ty, c) = st._grid
    assert c.size == (tx.size - 4) * (ty.size - 4)
    st = nat.PyrStep()
    lowering.lower_surface(s2, st)
    assert st.shape_kind == nat.SHAPE_COMBINATION and st.n_terms == 2
    assert (st.terms[1].dx, st.terms[1].dy, st.terms[1].dz) == (0.5, -0.25, 0.0)
    assert (st.terms[0].coeff_len, st.terms[1].coeff_off, st.terms[1].coeff_len) == (2, 2, 3)
    rot = pb.LocalCoordinates.p(name="s2_rot", tiltx=0.1)
    lc.addChild(rot)
    bad = pb.LinearCombination.p(lc, [(1.0, pb.Asphere.p(rot, curv=0.01))])
    s2.shape = bad
    with pytest.raises(lowering.LoweringError):
        lowering.lower_surface(s2, nat.PyrStep())


def test_structural_errors_like_the_reference():
    s = pb.OpticalSystem.p()
    lc0 = s.addLocalCoordinateSystem(pb.LocalCoordinates.p(name="obj"),
                                     refname=s.rootcoordinatesystem.name)
    elem = pb.OpticalElement.p(lc0, name="e")
    stray = pb.LocalCoordinates.p(name="stray")
    with pytest.raises(Exception):       # optical_element.py:76
        elem.addSurface("s", pb.Surface.p(stray), (None, None))
    with pytest.raises(Exception):       # optical_element.py:107
        elem.addMaterial("m", pb.ConstantIndexGlass.p(stray, 1.5))
    with pytest.raises(Exception):       # optical_system.py:227
        s.addElement("e2", pb.OpticalElement.p(stray))
    with pytest.raises(lowering.LoweringError):
        lowering.lower(s, [("nope", [])], configs.DLINE)


def test_lowering_follows_material_toggle_and_mirror_rules():
    (s, seq) = configs.build_system(configs.CONFIGS["x1_tilted"], pb.api())
    low = lowering.lower(s, seq, configs.DLINE)
    kinds = [(l.surfkey, round(l.st.before.n, 4), round(l.st.after.n, 4), l.st.interaction)
             for l in low]
    n_mg = round(pb.ModelGlass.p(pb.LocalCoordinates.p()).get_optical_index(None, configs.DLINE), 4)
    assert kinds == [("stop", 1.0, 1.0, 0), ("front", 1.0, n_mg, 0), ("back", n_mg, 1.0, 0),
                     ("mirror", 1.0, 1.0, 1), ("image", 1.0, 1.0, 0)]
    assert low[0].st.dir_mode == nat.DIR_POYNTING and low[1].st.dir_mode == nat.DIR_K
    assert low[2].st.k_norm_hint == pytest.approx(low[2].st.before.n)


def test_grin_profile_is_verified_against_python_source():
    spec = configs.CONFIGS["c5_grin"]
    (s, seq) = configs.build_system(spec, pb.api())
    lowering.lower(s, seq, configs.DLINE)                       # consistent: passes
    grin = s.elements["stdelem"].materials["grin"]
    grin.annotations["device_profile"] = dict(grin.annotations["device_profile"],
                                              params=[1.0, 0.4, 1.0, 4.0])
    with pytest.raises(lowering.LoweringError):
        lowering.lower(s, seq, configs.DLINE)
    del grin.annotations["device_profile"]
    with pytest.raises(lowering.LoweringError):
        lowering.lower(s, seq, configs.DLINE)


def _step_bytes(ls):
    st = nat.PyrStep()
    ctypes.memmove(ctypes.addressof(st), ctypes.addressof(ls.st), ctypes.sizeof(nat.PyrStep))
    (st.out_x, st.out_k, st.out_e, st.out_flags) = (None, None, None, None)
    return bytes(st)


def test_zernike_index_conventions():
    for j in range(1, 38):
        assert pb.ZernikeFringe.nmtoj(pb.ZernikeFringe.jtonm(j)) == j
        assert pb.ZernikeANSI.nmtoj(pb.ZernikeANSI.jtonm(j)) == j
    assert pb.ZernikeFringe.jtonm(4) == (2, 0) and pb.ZernikeFringe.jtonm(9) == (4, 0)


@pytest.mark.parametrize("name", sorted(configs.CONFIGS))
def test_reference_object_graph_lowers_like_ours(name):
    import refshim
    if not refshim.reference_available():
        pytest.skip("reference tree not present (GPU box)")
    (rs, rseq) = configs.build_system(configs.CONFIGS[name], refshim.api())
    (s, seq) = configs.build_system(configs.CONFIGS[name], pb.api())
    a = lowering.lower(rs, rseq, configs.DLINE)
    b = lowering.lower(s, seq, configs.DLINE)
    assert len(a) == len(b)
    for (la, lb) in zip(a, b):
        (sa, sb) = (la.st, lb.st)
        for f in ("shape_kind", "aperture_kind", "interaction", "dir_mode", "n_coeff", "split"):
            assert getattr(sa, f) == getattr(sb, f), f
        for f in ("curv", "cc", "curv2", "cc2", "normradius", "k_norm_hint"):
            assert getattr(sa, f) == pytest.approx(getattr(sb, f), rel=1e-15, abs=0)
        for fr in ("shape_frame", "aperture_frame"):
            assert np.allclose(list(getattr(sa, fr).r), list(getattr(sb, fr).r), atol=1e-15)
            assert np.allclose(list(getattr(sa, fr).o), list(getattr(sb, fr).o), atol=1e-13)
        for m in ("before", "after"):
            (ma, mb) = (getattr(sa, m), getattr(sb, m))
            assert ma.kind == mb.kind and ma.n == pytest.approx(mb.n, rel=1e-15)
            assert list(ma.eps) == list(mb.eps) and list(ma.grin_p) == list(mb.grin_p)
        assert sorted(zip(sa.xpow, sa.ypow, sa.coeff)) == sorted(zip(sb.xpow, sb.ypow, sb.coeff))


def test_convenience_builders_match_spec_builder():
    """build_rotationally_symmetric_optical_system (reference __init__.py:83-121)
    lowers to the same step table as the explicit construction of C2."""
    from pyrate_b200.configs import _DG, _N1, _N2, _N3
    idx = {"g1": _N1, "g2": _N2, "g3": _N3, None: None}
    rows = [(r, 0.0, dz, idx[mt], nm, {"is_stop": True} if nm == "stop" else {})
            for (nm, r, dz, mt) in _DG]
    (s1, seq1) = pb.build_rotationally_symmetric_optical_system(rows)
    (s2, seq2) = configs.build_system(configs.CONFIGS["c2_doublegauss"], pb.api())
    a = lowering.lower(s1, seq1, configs.DLINE)
    b = lowering.lower(s2, seq2, configs.DLINE)
    assert [l.surfkey for l in a] == [l.surfkey for l in b]
    for (la, lb) in zip(a, b):
        assert _step_bytes(la) == _step_bytes(lb)
    with pytest.raises(NotImplementedError):
        pb.build_rotationally_symmetric_optical_system([(10.0, 0, 1.0, "N-BK7", "s", {})])


FDC = (0.4861e-3, 0.5876e-3, 0.6563e-3)          # Fraunhofer F, d, C lines in mm


def test_wavelength_batch_lowering():
    """lower_batch: one step table for several wavelengths; only the media indices
    differ and they equal the single-wavelength lowerings."""
    from pyrate_b200 import _native as nat, lowering
    (s, seq) = configs.build_system(configs.CONFIGS["x14_dispersive"], pb.api())
    (per_wave, batch) = lowering.lower_batch(s, seq, FDC)
    assert len(per_wave) == 3 and len(batch) == 13
    seen_dispersion = False
    for (i, ls) in enumerate(batch):
        for w in range(3):
            assert ls.st.after_n_w[w] == per_wave[w][i].st.after.n
            assert ls.st.before_n_w[w] == per_wave[w][i].st.before.n
        assert ls.st.after.n == per_wave[0][i].st.after.n
        seen_dispersion = seen_dispersion or ls.st.after_n_w[0] != ls.st.after_n_w[2]
        if ls.st.after.n != 1.0:
            assert ls.st.after_n_w[0] > ls.st.after_n_w[1] > ls.st.after_n_w[2] > 1.0   # normal dispersion
    assert seen_dispersion
    with pytest.raises(lowering.LoweringError):
        lowering.lower_batch(s, seq, FDC + (0.7e-3, 0.8e-3))
    for name in ("c5_grin", "c4_anisotropic"):
        (s2, seq2) = configs.build_system(configs.CONFIGS[name], pb.api())
        with pytest.raises(lowering.LoweringError):
            lowering.lower_batch(s2, seq2, FDC)
    assert nat.MAX_WAVES == 4


def test_shape_gradients_are_the_derivatives_of_their_sag():
    """Every host shape mirror: getGrad(x, y) = (-dz/dx, -dz/dy, 1) of getSag by central
    differences (the reference's own test idea, tests/test_surf_shape.py:287-353, applied
    to all shape classes; the conic's implicit gradient is proportional to it)."""
    lc = pb.LocalCoordinates.p(name="fd")
    rng = np.random.default_rng(17)
    x = rng.uniform(-3, 3, 25)
    y = rng.uniform(-3, 3, 25)
    zco = [0.0] * 16
    (zco[3], zco[8], zco[15]) = (-0.05, 0.01, -0.002)       # m = 0 terms
    (xl, yl, zg) = configs.grid_arrays(configs.X11_GRIDSAG["surfaces"][2]["shape"][1]["grid"])
    asph = pb.Asphere.p(lc, curv=-0.03, cc=-0.7, coefficients=[1e-3, -2e-5, 1e-7])
    shapes = {
        "conic": pb.Conic.p(lc, curv=0.07, cc=-0.4),
        "asphere": asph,
        "biconic": pb.Biconic.p(lc, curvx=0.03, ccx=-0.6, curvy=-0.02, ccy=0.4,
                                coefficients=[(1e-3, 0.3), (-2e-5, -0.2)]),
        "xypoly": pb.XYPolynomials.p(lc, normradius=10.0, coefficients=[
            (2, 0, -0.9), (0, 2, -1.1), (1, 1, 0.05), (3, 0, 0.02), (1, 2, -0.03)]),
        "zernike": pb.ZernikeFringe.p(lc, normradius=5.0, coefficients=zco),
        "gridsag": pb.GridSag.p(lc, (xl, yl, zg)),
        "combination": pb.LinearCombination.p(lc, list_of_coefficients_and_shapes=[
            (1.0, asph), (0.5, pb.XYPolynomials.p(lc, normradius=10.0,
                                                  coefficients=[(2, 0, 0.02), (0, 3, 0.004)]))]),
    }
    h = 1e-5
    for (name, sh) in shapes.items():
        sag = lambda a, b: np.asarray(sh.getSag(a.copy(), b.copy()), dtype=float)   # noqa: E731
        dzdx = (sag(x + h, y) - sag(x - h, y)) / (2 * h)
        dzdy = (sag(x, y + h) - sag(x, y - h)) / (2 * h)
        g = np.asarray(sh.getGrad(x.copy(), y.copy()), dtype=float)
        g = g / g[2]                                           # conic: implicit gradient
        assert np.allclose(g[0], -dzdx, rtol=1e-6, atol=1e-8), name
        assert np.allclose(g[1], -dzdy, rtol=1e-6, atol=1e-8), name
        n = np.asarray(sh.getNormal(x.copy(), y.copy()), dtype=float)
        assert np.allclose(np.sum(n * n, axis=0), 1.0, atol=1e-12), name
