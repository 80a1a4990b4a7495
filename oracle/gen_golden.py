"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz from the unmodified
reference (mess42/pyrate at /root/reference, imported through oracle/refshim.py).

Run in the build container:   python oracle/gen_golden.py
The fixtures are committed; the GPU box never needs /root/reference.

What is dumped (all float64 / complex128 / bool / int64, NumPy .npz):
  seqtrace_<config>.npz  every RayBundle of every RayPath returned by
                         OpticalSystem.seqtrace (optical_system.py:73-94):
                         x, k, valid, rayID (+E for anisotropic segments); for
                         GRIN bundles only the first and the last two history
                         rows plus the row count (history is 200+ rows).
  frames.npz             LocalCoordinates chains (localcoordinates.py:238-307)
  shapes.npz             getSag/getGrad/getNormal (surface_shape.py:100-237,
                         :529-555, :785-807)
  aniso_modes.npz        MaxwellMaterial.sortKnormEField (material.py:122-153)
  spot.npz               RayBundleAnalysis centroid / rms (ray_analysis.py:44-86)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import refshim  # noqa: E402
from pyrate_b200 import configs  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# (config, rings, kdir, efield, splitup, tag)
DEG = np.pi / 180.0
TRACES = [
    ("c1_doublet", 18, (0, 0, 1), (0, 1, 0), False, ""),
    ("c2_doublegauss", 6, (0, 0, 1), (0, 1, 0), False, ""),
    ("c2_doublegauss", 4, (0, np.sin(5 * DEG), np.cos(5 * DEG)), (1, 0, 0),
     False, "_field5"),
    # E0 not perpendicular to k0: first segment direction follows the
    # Poynting vector (ray.py:136-152)
    ("c2_doublegauss", 3, (0, np.sin(3 * DEG), np.cos(3 * DEG)), (0, 1, 0),
     False, "_obliqueE"),
    ("c3_asphere", 6, (0, 0, 1), (0, 1, 0), False, ""),
    ("c4_anisotropic", 3, (0, 0, 1), (0, 1, 0), False, ""),
    ("c4_anisotropic", 2, (0, 0, 1), (0, 1, 0), True, "_split"),
    ("c5_grin", 3, (0, 0, 1), (0, 1, 0), False, ""),
    ("x1_tilted", 8, (0, 0, 1), (0, 1, 0), False, ""),
    ("x2_xypoly", 6, (0, 0, 1), (0, 1, 0), False, ""),
    ("x3_vignette", 10, (0, 0, 1), (0, 1, 0), False, ""),
    ("x4_biaxial", 3, (0, np.sin(2 * DEG), np.cos(2 * DEG)), (1, 0, 0), False, ""),
    ("x5_degenerate", 2, (0, 0, 1), (0, 1, 0), False, ""),
    ("x6_biconic", 6, (np.sin(1 * DEG), 0, np.cos(1 * DEG)), (0, 1, 0), False, ""),
    ("x7_two_elements", 5, (0, np.sin(2 * DEG), np.cos(2 * DEG)), (1, 0, 0), False, ""),
    ("x8_crystal_mirror", 2, (0, np.sin(1 * DEG), np.cos(1 * DEG)), (1, 0, 0), False, ""),
    ("x9_zernike", 6, (np.sin(1 * DEG), 0, np.cos(1 * DEG)), (0, 1, 0), False, ""),
    ("x11_gridsag", 5, (np.sin(1 * DEG), 0, np.cos(1 * DEG)), (0, 1, 0), False, ""),
    ("x12_combination", 5, (0, np.sin(1 * DEG), np.cos(1 * DEG)), (1, 0, 0), False, ""),
    ("x13_tirglass", 6, (0, np.sin(1 * DEG), np.cos(1 * DEG)), (1, 0, 0), False, ""),
]


def dump_trace(api, name, rings, kdir, efield, splitup, tag):
    spec = configs.CONFIGS[name]
    (s, seq) = configs.build_system(spec, api)
    (x0, k0, e0) = configs.config_bundle(spec, rings, kdir, efield)
    bundle = api.RayBundle(x0, k0, e0, wave=configs.DLINE)
    paths = s.seqtrace(bundle, seq, splitup=splitup)
    out = {"x0": x0, "k0": k0, "E0": e0, "npaths": np.int64(len(paths)),
           "splitup": np.bool_(splitup)}
    for (ip, path) in enumerate(paths):
        out["p%d_nbundles" % ip] = np.int64(len(path.raybundles))
        for (ib, rb) in enumerate(path.raybundles):
            pre = "p%d_b%d_" % (ip, ib)
            rows = rb.x.shape[0]
            out[pre + "rows"] = np.int64(rows)
            sel = slice(None) if rows <= 3 else [0, rows - 2, rows - 1]
            out[pre + "x"] = np.asarray(rb.x)[sel]
            out[pre + "k"] = np.asarray(rb.k)[sel]
            out[pre + "valid"] = np.asarray(rb.valid)[sel]
            out[pre + "rayID"] = np.asarray(rb.rayID, dtype=np.int64)
            if np.iscomplexobj(rb.Efield) or name.startswith(("c4", "x4", "x5", "x8")):
                out[pre + "E"] = np.asarray(rb.Efield)[sel]
    fn = os.path.join(OUT, "seqtrace_%s%s.npz" % (name, tag))
    np.savez_compressed(fn, **out)
    nlast = paths[0].raybundles[-1].x.shape[2]
    print("%-32s paths=%d bundles=%d last-width=%d" %
          (os.path.basename(fn), len(paths), len(paths[0].raybundles), nlast))
    return paths


def dump_frames(api):
    rng = np.random.default_rng(20260925)
    n = 24
    params = np.zeros((n, 3, 7))
    basis = np.zeros((n, 3, 3, 3))
    origin = np.zeros((n, 3, 3))
    for i in range(n):
        parent = None
        for lvl in range(3):
            dec = rng.uniform(-10, 10, 3)
            tilt = rng.uniform(-np.pi, np.pi, 3)
            ttd = int(rng.integers(0, 2))
            params[i, lvl] = (*dec, *tilt, ttd)
            lc = api.LocalCoordinates.p(name="l%d_%d" % (i, lvl), decx=dec[0],
                                        decy=dec[1], decz=dec[2],
                                        tiltx=tilt[0], tilty=tilt[1],
                                        tiltz=tilt[2], tiltThenDecenter=ttd)
            if parent is not None:
                parent.addChild(lc)
            parent = lc
            basis[i, lvl] = lc.localbasis
            origin[i, lvl] = lc.globalcoordinates
    pts = rng.uniform(-5, 5, (3, 7))
    lc_last = parent
    np.savez_compressed(
        os.path.join(OUT, "frames.npz"), params=params, basis=basis,
        origin=origin, pts=pts,
        last_l2g_pts=lc_last.returnLocalToGlobalPoints(pts),
        last_g2l_pts=lc_last.returnGlobalToLocalPoints(pts),
        last_l2g_dir=lc_last.returnLocalToGlobalDirections(pts),
        last_g2l_dir=lc_last.returnGlobalToLocalDirections(pts))
    print("frames.npz")


def dump_shapes(api):
    rng = np.random.default_rng(7)
    lc = api.LocalCoordinates.p(name="shapes")
    x = rng.uniform(-3, 3, 40)
    y = rng.uniform(-3, 3, 40)
    out = {"x": x, "y": y}
    conic_params = [(0.05, 0.0), (-0.08, -1.0), (0.11, 0.7), (0.0, 0.0),
                    (0.3, 2.0)]   # last one: sag undefined for large r
    out["conic_params"] = np.array(conic_params)
    for (i, (curv, cc)) in enumerate(conic_params):
        sh = api.Conic.p(lc, curv=curv, cc=cc)
        out["conic%d_sag" % i] = sh.getSag(x.copy(), y.copy())
        out["conic%d_grad" % i] = sh.getGrad(x.copy(), y.copy())
        out["conic%d_normal" % i] = sh.getNormal(x.copy(), y.copy())
    asph = api.Asphere.p(lc, curv=-0.02, cc=-1.0,
                         coefficients=[1e-3, 1e-5, -1e-7])
    out["asph_params"] = np.array([-0.02, -1.0, 1e-3, 1e-5, -1e-7])
    out["asph_sag"] = asph.getSag(x, y)
    out["asph_grad"] = asph.getGrad(x, y)
    out["asph_normal"] = asph.getNormal(x, y)
    coeffs = [(2, 0, -0.9), (0, 2, -1.1), (1, 1, 0.05), (3, 0, 0.02),
              (1, 2, -0.03), (0, 0, 0.1), (0, 1, 0.02)]
    xy = api.XYPolynomials.p(lc, normradius=10.0, coefficients=coeffs)
    out["xy_coeffs"] = np.array(coeffs, dtype=float)
    out["xy_normradius"] = np.float64(10.0)
    out["xy_sag"] = xy.getSag(x, y)
    out["xy_grad"] = xy.getGrad(x, y)
    out["xy_normal"] = xy.getNormal(x, y)
    rngz = np.random.default_rng(21)
    for (nm, cls, ncoef) in (("zf", api.ZernikeFringe, 25), ("za", api.ZernikeANSI, 21)):
        co = rngz.normal(size=ncoef) * 0.02
        zs = cls.p(lc, normradius=5.0, coefficients=list(co))
        out[nm + "_coeffs"] = co
        out[nm + "_sag"] = zs.getSag(x, y)
    bic = api.Biconic.p(lc, curvx=0.03, ccx=-0.6, curvy=-0.02, ccy=0.4,
                        coefficients=[(1e-3, 0.3), (-2e-5, -0.2)])
    out["bic_params"] = np.array([0.03, -0.6, -0.02, 0.4, 1e-3, 0.3, -2e-5, -0.2])
    out["bic_sag"] = bic.getSag(x, y)
    out["bic_grad"] = bic.getGrad(x, y)
    out["bic_normal"] = bic.getNormal(x, y)
    np.savez_compressed(os.path.join(OUT, "shapes.npz"), **out)
    print("shapes.npz")


def dump_shapes2(api):
    """GridSag and LinearCombination sag / gradient (separate file: shapes.npz predates
    them)."""
    rng = np.random.default_rng(33)
    lc = api.LocalCoordinates.p(name="shapes2")
    x = rng.uniform(-9.5, 9.5, 60)        # a few points outside the grid (clamped)
    y = rng.uniform(-8.5, 8.5, 60)
    out = {"x": x, "y": y}
    grid = configs.X11_GRIDSAG["surfaces"][2]["shape"][1]["grid"]
    gs = api.GridSag.p(lc, configs.grid_arrays(grid))
    out["grid_sag"] = gs.getSag(x, y)
    out["grid_grad"] = gs.getGrad(x, y)
    xs = x * 0.7
    ys = y * 0.7
    (out["xs"], out["ys"]) = (xs, ys)
    lcd = api.LocalCoordinates.p(name="shapes2_dec", decx=0.5, decy=-0.25)
    lc.addChild(lcd)
    asph = api.Asphere.p(lc, curv=1. / 45.0, cc=-0.8, coefficients=[2e-6, -1e-9])
    xyp = api.XYPolynomials.p(lcd, normradius=10.0,
                              coefficients=[(2, 0, 0.02), (1, 1, -0.01), (0, 3, 0.004)])
    comb = api.LinearCombination.p(lc, list_of_coefficients_and_shapes=[(1.0, asph),
                                                                       (0.5, xyp)])
    out["comb_sag"] = comb.getSag(xs, ys)
    out["comb_grad"] = comb.getGrad(xs, ys)
    np.savez_compressed(os.path.join(OUT, "shapes2.npz"), **out)
    print("shapes2.npz")


def dump_aniso(api):
    rng = np.random.default_rng(11)
    lc = api.LocalCoordinates.p(name="aniso")
    npts = 12
    cases = {
        "uniaxial_y": np.diag([1.658 ** 2, 1.486 ** 2, 1.658 ** 2]),
        "biaxial_rot": None,
    }
    # rotated biaxial tensor (real symmetric, positive definite)
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    cases["biaxial_rot"] = q @ np.diag([2.2, 2.6, 3.1]) @ q.T
    out = {}
    for (nm, eps) in cases.items():
        mat = api.AnisotropicMaterial.p(lc, eps, name=nm)
        n = rng.normal(size=(3, npts))
        n[2] = np.abs(n[2]) + 2.0
        n /= np.linalg.norm(n, axis=0)
        kin = rng.normal(size=(3, npts)) * 0.3
        kin[2] += 1.0
        kpa = kin - np.sum(kin * n, axis=0) * n
        x = np.zeros((3, npts))
        (k4, e4) = mat.sortKnormEField(x, n, kpa, n)
        (xi4, ev4) = mat.calcXiEigenvectorsNorm(x, n, kpa)
        out[nm + "_eps"] = eps
        out[nm + "_n"] = n
        out[nm + "_kpa"] = kpa
        out[nm + "_k4"] = k4
        out[nm + "_e4"] = e4
        out[nm + "_xi4"] = xi4
    np.savez_compressed(os.path.join(OUT, "aniso_modes.npz"), **out)
    print("aniso_modes.npz")


def dump_spot(api, paths):
    from pyrateoptics.raytracer.analysis.ray_analysis import RayBundleAnalysis
    last = paths[0].raybundles[-1]
    ra = RayBundleAnalysis(last)
    c = ra.get_centroid_position()
    np.savez_compressed(os.path.join(OUT, "spot.npz"), x=np.asarray(last.x[-1]),
                        centroid=c, rms=np.float64(ra.get_rms_spot_size(c)),
                        rms0=np.float64(ra.get_rms_spot_size(np.zeros(3))))
    print("spot.npz rms=%r" % ra.get_rms_spot_size(c))


def main():
    os.makedirs(OUT, exist_ok=True)
    api = refshim.api()
    np.random.seed(0)
    only = [a.split("=", 1)[1].split(",") for a in sys.argv if a.startswith("--only=")]
    if only:                 # regenerate selected fixtures: --only=x11_gridsag,shapes2
        for t in TRACES:
            if t[0] + t[5] in only[0]:
                dump_trace(api, *t)
        if "shapes2" in only[0]:
            dump_shapes2(api)
        return
    for t in TRACES:
        paths = dump_trace(api, *t)
        if t[0] == "c2_doublegauss" and t[5] == "":
            dump_spot(api, paths)
    dump_frames(api)
    dump_shapes(api)
    dump_shapes2(api)
    dump_aniso(api)


if __name__ == "__main__" and not ({"--rasters", "--glasscat", "--paraxial", "--pathanalysis"} & set(sys.argv)):
    main()


def dump_glasscat():
    """A handful of refractiveindex.info pages (one per dispersion formula that
    occurs in the glass shelves) with the reference's indices at F, d, C."""
    import glob
    import json
    import yaml
    from pyrateoptics.raytracer.material.material_glasscat import CatalogMaterial
    from pyrateoptics.raytracer.localcoordinates import LocalCoordinates
    base = os.path.join(refshim.REFERENCE_ROOT, "pyrateoptics",
                        "refractiveindex.info-database", "database", "data")
    picks = ["glass/schott/N-BK7.yml", "glass/schott/SF5.yml", "glass/ohara/S-LAH64.yml",
             "glass/hoya/FCD1.yml", "main/SiO2/Malitson.yml", "main/CaF2/Li.yml",
             "main/H2O/Daimon-20.0C.yml", "main/Ar/Bideau-Mehu.yml", "main/Si/Edwards.yml",
             "organic/C3H8O3 - glycerol/Rheims.yml",
             "organic/C8H5KO4 - potassium hydrogen phthalate/Moutzouris-beta.yml",
             "organic/C4H10O - butanol/El-Kashef.yml", "other/mixed gases/air/Ciddor.yml"]
    extra = glob.glob(os.path.join(base, "**", "*.yml"), recursive=True)
    for typ in ("formula 7", "tabulated n"):
        for fn in sorted(extra):
            try:
                txt = open(fn).read()
            except Exception:
                continue
            if ("type: " + typ + "\n") in txt and os.path.relpath(fn, base) not in picks:
                picks.append(os.path.relpath(fn, base))
                if sum(1 for p_ in picks if typ in open(os.path.join(base, p_)).read()) >= 3:
                    break
    lc = LocalCoordinates.p(name="gc")
    out = []
    waves = [0.4861e-3, 0.5876e-3, 0.6563e-3]
    for rel in picks:
        fn = os.path.join(base, rel)
        if not os.path.exists(fn):
            continue
        d = yaml.safe_load(open(fn))
        data = [e for e in d["DATA"]]
        if any(e["type"].startswith("tabulated") and len(e["data"]) > 1500 for e in data):
            continue
        try:
            mat = CatalogMaterial.p(lc, {"DATA": data})
            (lo, hi) = (max(t.waverange[0] for t in mat.nk_table),
                        min(t.waverange[1] for t in mat.nk_table))
            use = waves
            if lo > 1e3 * waves[0] or hi < 1e3 * waves[-1]:
                use = [1e-3 * (lo + f * (hi - lo)) for f in (0.2, 0.5, 0.8)]
            ns = [complex(mat.get_optical_index(None, w)) for w in use]
        except Exception as exc:        # unsupported formula etc.
            continue
        out.append({"page": rel, "DATA": data, "waves_mm": use,
                    "n_real": [v.real for v in ns], "n_imag": [v.imag for v in ns]})
    json.dump(out, open(os.path.join(OUT, "glasscat.json"), "w"), indent=1)
    print("glasscat.json", [(o["page"], o["DATA"][0]["type"]) for o in out])


def dump_rasters():
    """sampling2d rasters and OpticalSystemAnalysis bundle generators of the
    reference (sampling2d/raster.py, analysis/optical_system_analysis.py:83-165)."""
    from pyrateoptics.sampling2d import raster
    from pyrateoptics.raytracer.analysis.optical_system_analysis import OpticalSystemAnalysis
    out = {}
    for (nm, n) in (("RectGrid", 60), ("HexGrid", 60), ("MeridionalFan", 11), ("SagitalFan", 7),
                    ("ChiefAndComa", 6), ("Single", 1), ("CircularGrid", 49)):
        (x, y) = getattr(raster, nm)().getGrid(n)
        out[nm + "_x"] = x
        out[nm + "_y"] = y
        out[nm + "_n"] = np.int64(n)
    api = refshim.api()
    (s, seq) = configs.build_system(configs.CONFIGS["c1_doublet"], api)
    osa = OpticalSystemAnalysis(s, seq)
    props = {"radius": 11.43, "startz": -5.0, "starty": 0.3, "anglex": 0.02, "angley": -0.01,
             "raster": raster.RectGrid()}
    (o, k, e) = osa.collimated_bundle(40, props, wave=configs.DLINE)
    (out["coll_o"], out["coll_k"], out["coll_e"]) = (o, k, e)
    props = {"radius": 0.2, "startz": -50.0, "anglex": 0.01, "raster": raster.HexGrid()}
    (o, k, e) = osa.divergent_bundle(40, props, wave=configs.DLINE)
    (out["div_o"], out["div_k"], out["div_e"]) = (o, k, e)
    np.savez_compressed(os.path.join(OUT, "rasters.npz"), **out)
    print("rasters.npz")


if __name__ == "__main__" and "--rasters" in sys.argv:
    refshim.install()
    dump_rasters()


if __name__ == "__main__" and "--glasscat" in sys.argv:
    refshim.install()
    dump_glasscat()


PARAXIAL_CASES = [("c2_doublegauss", 3.0), ("x1_tilted", 2.0), ("x7_two_elements", 3.0),
                  ("c1_doublet", 4.0)]


def dump_paraxial():
    """The paraxial callers of seqtrace in the unmodified reference: pilot bundles
    (helpers.py:78-316), OpticalSystem.extractXYUV / para_seqtrace
    (optical_system.py:105-214, optical_element.py:165-322, :381-469) and Aimy
    (aim.py:46-320) on systems with the stop inside, tilted frames + a mirror, and two
    elements."""
    from pyrateoptics.raytracer.helpers import build_pilotbundle, build_pilotbundle_complex
    from pyrateoptics.raytracer.aim import Aimy
    api = refshim.api()
    out = {}
    for (name, stopsize) in PARAXIAL_CASES:
        spec = configs.CONFIGS[name]
        (s, seq) = configs.build_system(spec, api)
        objsurf = s.elements[seq[0][0]].surfaces[seq[0][1][0][0]]
        for (gen, fn) in (("real", build_pilotbundle), ("complex", build_pilotbundle_complex)):
            pre = "%s_%s_" % (name, gen)
            pbs = fn(objsurf, s.material_background, (0.1, 0.1), (1 * DEG, 1 * DEG),
                     num_sampling_points=3)
            pb = pbs[-1]
            out[pre + "pilot_x"] = np.array(pb.x[0])
            out[pre + "pilot_k"] = np.array(pb.k[0])
            (m1, m2) = s.extractXYUV(pb, seq, pilotbundle_generation=gen)
            out[pre + "m_obj_stop"] = m1
            out[pre + "m_stop_img"] = m2
            # per-pair matrices of the first element
            pb = fn(objsurf, s.material_background, (0.1, 0.1), (1 * DEG, 1 * DEG),
                    num_sampling_points=3)[-1]
            (elemkey, subseq) = seq[0]
            (ppath, mats) = s.elements[elemkey].calculateXYUV(
                pb, subseq, s.material_background, pilotbundle_generation=gen)
            (hitlist, _) = s.elements[elemkey].sequence_to_hitlist(subseq)
            if name in ("c2_doublegauss", "x1_tilted"):       # (the others repeat these)
                out[pre + "pair_matrices"] = np.array([mats[h] for h in hitlist])
                out[pre + "pair_inverse"] = np.array([mats[(h[1], h[0], h[2])]
                                                      for h in hitlist])
                out[pre + "pilotpath_x"] = np.array(
                    [b.x[-1] for b in ppath.raybundles[:len(hitlist) + 1]]).real
                out[pre + "pilotpath_k"] = np.array(
                    [b.k[-1] for b in ppath.raybundles[:len(hitlist) + 1]])
            # linearised trace of a small fan
            pb = fn(objsurf, s.material_background, (0.1, 0.1), (1 * DEG, 1 * DEG),
                    num_sampling_points=3)[-1]
            rng = np.random.default_rng(5)
            x0 = np.vstack((rng.uniform(-1, 1, (2, 9)), np.zeros((1, 9)))) + \
                np.array(pb.x[0][:, :1]).real
            k0 = np.array(pb.k[0][:, :1]).real + \
                np.vstack((rng.uniform(-0.02, 0.02, (2, 9)), np.zeros((1, 9))))
            e0 = np.zeros((3, 9))
            e0[1] = 1.0
            (pp, rp) = s.para_seqtrace(pb, api.RayBundle(x0, k0, e0, wave=configs.DLINE), seq,
                                       pilotbundle_generation=gen)
            out[pre + "para_x0"] = x0
            out[pre + "para_k0"] = k0
            out[pre + "para_x"] = np.array([b.x[-1] for b in rp.raybundles])
            out[pre + "para_k"] = np.array([b.k[-1] for b in rp.raybundles])
            a = Aimy(s, seq, wave=configs.DLINE, num_pupil_points=24, stopsize=stopsize,
                     pilotbundle_generation=gen)
            out[pre + "aimy_m_obj_stop"] = a.m_obj_stop
            b1 = a.aim(np.array([0.01, -0.02]), fieldtype="angle")
            (out[pre + "aim_angle_x"], out[pre + "aim_angle_k"]) = (np.array(b1.x[0]), np.array(b1.k[0]))
            try:        # the stop is the object surface in some cases: B block singular
                b2 = a.aim(np.array([0.3, 0.1]), fieldtype="objectheight")
                (out[pre + "aim_height_x"], out[pre + "aim_height_k"]) = \
                    (np.array(b2.x[0]), np.array(b2.k[0]))
            except np.linalg.LinAlgError:
                pass
            print(pre, pb.x.shape, np.linalg.cond(m2))
    np.savez_compressed(os.path.join(OUT, "paraxial.npz"), **out)
    print("paraxial.npz")


if __name__ == "__main__" and "--paraxial" in sys.argv:
    refshim.install()
    dump_paraxial()


def dump_pathanalysis():
    """RayPathAnalysis (analysis/ray_analysis.py:170-213) of the unmodified reference on
    the double-Gauss fixture bundle (5 degree field)."""
    from pyrateoptics.raytracer.analysis.ray_analysis import RayPathAnalysis
    api = refshim.api()
    spec = configs.CONFIGS["c2_doublegauss"]
    (s, seq) = configs.build_system(spec, api)
    (x0, k0, e0) = configs.config_bundle(spec, 4, (0, np.sin(5 * DEG), np.cos(5 * DEG)), (1, 0, 0))
    path = s.seqtrace(api.RayBundle(x0, k0, e0, wave=configs.DLINE), seq)[0]
    rpa = RayPathAnalysis(path)
    np.savez_compressed(
        os.path.join(OUT, "pathanalysis.npz"), x0=x0, k0=k0, E0=e0,
        arc=rpa.get_arc_length(), phase=rpa.get_phase_difference(),
        arc_2_9=rpa.get_arc_length(first=2, last=9),
        rel=rpa.get_relative_phase_difference(referenceray=0, wavelength=configs.DLINE))
    print("pathanalysis.npz")


if __name__ == "__main__" and "--pathanalysis" in sys.argv:
    refshim.install()
    dump_pathanalysis()
