"""GPU: the native sm_100a path (through OpticalSystem.seqtrace -> lowering ->
C ABI pyr_trace) against (a) the committed reference fixtures and (b) the NumPy
oracle on seeded bundles.  Tolerances are BASELINE.json's: <= 1e-10 relative for
closed-form isotropic conics, <= 1e-6 for iterated asphere / XY / GRIN."""
import numpy as np
import pytest

import pyrate_b200 as pb
from pyrate_b200 import configs

import util

pytestmark = pytest.mark.gpu

REAL_TAGS = [t for t in util.golden_traces() if not t.startswith(("c4", "x4", "x5", "x8"))]


def _device_paths(name, x0, k0, e0, splitup=False, record_efield=False):
    (s, seq) = configs.build_system(configs.CONFIGS[name], pb.api())
    bundle = pb.RayBundle(x0, k0, e0, wave=configs.DLINE)
    return s.seqtrace(bundle, seq, splitup=splitup, record_efield=record_efield)


@pytest.mark.parametrize("tag", REAL_TAGS)
def test_matches_reference_fixture(tag):
    g = util.load_golden(tag)
    name = util.config_of(tag)
    paths = _device_paths(name, g["x0"], g["k0"], g["E0"])
    ref_paths = util.golden_paths(g)
    assert len(paths) == len(ref_paths) == 1
    tol = util.tolerance_of(name)
    (path, rpath) = (paths[0].raybundles, ref_paths[0])
    assert len(path) == len(rpath)
    assert path[0] is path[1]
    worst = 0.0
    for (ib, (b, rb)) in enumerate(zip(path, rpath)):
        grin_bundle = rb["rows"] > 3
        worst = max(worst, util.compare_bundle(
            b.numpy(), rb, tol, "%s b%d" % (tag, ib), first_last_only=grin_bundle,
            check_k_last=not grin_bundle))
    print("%s worst rel err %.3e" % (tag, worst))


@pytest.mark.parametrize("name,rings", [("c1_doublet", 30), ("c2_doublegauss", 40),
                                        ("x1_tilted", 25), ("x3_vignette", 30),
                                        ("c3_asphere", 12), ("x2_xypoly", 12),
                                        ("x6_biconic", 12), ("x9_zernike", 10),
                                        ("x10_zernike_general", 10),
                                        ("x11_gridsag", 10), ("x12_combination", 10)])
def test_matches_oracle_on_larger_bundles(name, rings):
    import pyrate_np as onp
    spec = configs.CONFIGS[name]
    deg = np.pi / 180.0
    (x0, k0, e0) = configs.config_bundle(spec, rings, (np.sin(1.5 * deg) * 0.6,
                                                       np.sin(1.5 * deg) * 0.8,
                                                       np.cos(1.5 * deg)), (0., 1., 0.))
    paths = _device_paths(name, x0, k0, e0)
    ref = onp.seqtrace(onp.system_from_spec(spec), x0, k0, e0, wave=configs.DLINE)
    tol = util.tolerance_of(name)
    for (ib, (b, rb)) in enumerate(zip(paths[0].raybundles, ref[0])):
        util.compare_bundle(b.numpy(), {"x": rb["x"], "k": rb["k"], "valid": rb["valid"],
                                        "rayID": rb["rayID"]}, tol, "%s b%d" % (name, ib))


def test_tir_glass_total_reflection_branch_matches_oracle():
    """IsotropicMaterialTIR with rays that ARE totally reflected at the cemented
    surface (the reference itself raises there, material_isotropic_tir.py:116): the
    device drops them like the angle-form restatement does."""
    import copy
    import pyrate_np as onp
    spec = copy.deepcopy(configs.CONFIGS["x13_tirglass"])
    spec["bundle"]["radius"] = 7.0
    deg = np.pi / 180.0
    (x0, k0, e0) = configs.config_bundle(spec, 14, (0., np.sin(deg), np.cos(deg)), (1., 0., 0.))
    (s, seq) = configs.build_system(spec, pb.api())
    paths = s.seqtrace(pb.RayBundle(x0, k0, e0, wave=configs.DLINE), seq)
    ref = onp.seqtrace(onp.system_from_spec(spec), x0, k0, e0, wave=configs.DLINE)
    widths = [b.numpy()["x"].shape[2] for b in paths[0].raybundles]
    assert widths[-1] < widths[0] and widths[-1] > 0          # some rays reflected, not all
    for (ib, (b, rb)) in enumerate(zip(paths[0].raybundles, ref[0])):
        util.compare_bundle(b.numpy(), {"x": rb["x"], "k": rb["k"], "valid": rb["valid"],
                                        "rayID": rb["rayID"]}, util.TOL_CLOSED_FORM,
                            "x13 tir b%d" % ib)


def test_grin_history_rows_match_reference_and_oracle():
    """Opt-in integrator history: the GRIN bundle carries one row per integrator
    step like the reference's (material_grin.py:198-205); fixture rows [0, P-2, P-1]
    and the row count are the unmodified reference's, every row is checked
    against the oracle."""
    import pyrate_np as onp
    g = util.load_golden("c5_grin")
    spec = configs.CONFIGS["c5_grin"]
    (s, seq) = configs.build_system(spec, pb.api())
    bundle = pb.RayBundle(g["x0"], g["k0"], g["E0"], wave=configs.DLINE)
    path = s.seqtrace(bundle, seq, grin_history=True)[0].raybundles
    rpath = util.golden_paths(g)[0]
    ref = onp.seqtrace(onp.system_from_spec(spec), g["x0"], g["k0"], g["E0"],
                       wave=configs.DLINE, per_ray_energy=True)[0]
    assert len(path) == len(rpath)
    seen = 0
    for (ib, (b, rb, ob)) in enumerate(zip(path, rpath, ref)):
        d = b.numpy()
        assert d["x"].shape[0] == rb["rows"], (ib, d["x"].shape, rb["rows"])
        if rb["rows"] <= 3:
            continue
        seen += 1
        rows = [0, rb["rows"] - 2, rb["rows"] - 1]
        sub = {"x": d["x"][rows], "k": d["k"][rows], "valid": d["valid"][rows],
               "rayID": d["rayID"]}
        util.compare_bundle(sub, rb, 1e-9, "history vs fixture b%d" % ib)
        util.compare_bundle(d, {"x": ob["x"], "k": ob["k"], "valid": ob["valid"],
                                "rayID": ob["rayID"]}, 1e-9, "history vs oracle b%d" % ib)
    assert seen == 1


def test_efield_invariants_and_input_untouched():
    spec = configs.CONFIGS["c2_doublegauss"]
    (x0, k0, e0) = configs.config_bundle(spec, 12)
    (x0c, k0c, e0c) = (x0.copy(), k0.copy(), e0.copy())
    for rec_e in (False, True):
        paths = _device_paths("c2_doublegauss", x0, k0, e0, record_efield=rec_e)
        for b in paths[0].raybundles[2:]:
            d = b.numpy()
            (k, e) = (d["k"][0], d["Efield"][0])
            assert np.allclose(np.sum(e * e, axis=0), 1.0, atol=1e-12)
            assert np.max(np.abs(np.sum(e * k, axis=0))) < 1e-12
    assert np.array_equal(x0, x0c) and np.array_equal(k0, k0c) and np.array_equal(e0, e0c)


def test_empty_and_tiny_bundles():
    spec = configs.CONFIGS["c1_doublet"]
    for n in (0, 1, 2, 3, 31, 33):
        (x0, k0, e0) = configs.config_bundle(spec, 4)
        (x0, k0, e0) = (x0[:, :n].copy(), k0[:, :n].copy(), e0[:, :n].copy())
        paths = _device_paths("c1_doublet", x0, k0, e0)
        last = paths[0].raybundles[-1].numpy()
        assert last["x"].shape == (1, 3, n)
        if n:
            import pyrate_np as onp
            ref = onp.seqtrace(onp.system_from_spec(spec), x0, k0, e0)[0][-1]
            assert util.relerr(last["x"], ref["x"]) < 1e-10
            assert util.relerr(last["k"], ref["k"]) < 1e-10


def test_full_size_properties_double_gauss():
    """BASELINE size (1e7 rays): size-independent checks -- |k| = n after every
    surface, all hit points on their surface, symmetry of the on-axis bundle,
    and agreement of a strided subsample with the oracle."""
    import torch
    import pyrate_np as onp
    spec = configs.CONFIGS["c2_doublegauss"]
    (x0, k0, e0) = configs.config_bundle(spec)          # 9 997 351 rays
    paths = _device_paths("c2_doublegauss", x0, k0, e0)
    rec = paths[0].record
    nidx = [1.0] + [ls.st.after.n for ls in rec.lowered]
    for (s, k) in enumerate(rec.k):
        alive = (rec.flags[s] & 2) != 0
        kn = torch.sqrt((k * k).sum(0))[alive]
        assert float((kn - nidx[s + 1]).abs().max()) < 1e-12
        assert bool(alive.all())
    # centroid of an on-axis hexapolar bundle stays on the axis
    last = rec.hit[-1]
    assert abs(float(last[0].mean())) < 1e-9 and abs(float(last[1].mean())) < 1e-9
    sub = np.arange(0, x0.shape[1], 9973)
    ref = onp.seqtrace(onp.system_from_spec(spec), x0[:, sub], k0[:, sub], e0[:, sub])
    for s in range(len(rec.hit)):
        got_x = rec.hit[s][:, sub].cpu().numpy()
        got_k = rec.k[s][:, sub].cpu().numpy()
        assert util.relerr(got_x, ref[0][s + 1]["x"][-1]) < 1e-10
        assert util.relerr(got_k, ref[0][s + 2]["k"][0]) < 1e-10


@pytest.mark.parametrize("chunk,default_e", [(4000, False), (4096, False), (4096, True)])
def test_host_entry_matches_device_path(chunk, default_e):
    """pyr_trace_host (chunked H2D -> trace -> D2H pipeline) against the device
    resident path, with a chunk size that does not divide the bundle (4000: plain
    schedule, 4096: ramped first/last chunks) and with the default field."""
    import torch
    from pyrate_b200 import engine, lowering
    spec = configs.CONFIGS["c2_doublegauss"]
    (x0, k0, e0) = configs.config_bundle(spec, 70)        # 14 911 rays, E0 = (0,1,0)
    n = x0.shape[1]
    (s, seq) = configs.build_system(spec, pb.api())
    lowered = lowering.lower(s, seq, configs.DLINE)
    rec = engine.trace(lowered, x0, k0, e0, configs.DLINE)
    ht = engine.HostTracer(lowered, n, chunk_rays=chunk)
    (xp, kp, ep) = (torch.from_numpy(a).pin_memory() for a in (x0, k0, e0))
    (xl, kl, fl, spot8) = ht(xp, kp, None if default_e else ep)
    assert np.array_equal(fl.numpy(), rec.flags[-1].cpu().numpy())
    assert np.array_equal(xl.numpy(), rec.hit[-1].cpu().numpy())
    assert np.array_equal(kl.numpy(), rec.k[-1].cpu().numpy())
    origin = engine.last_surface_origin(lowered)
    dev_sums = engine.spot_sums(rec.hit[-1], rec.flags[-1], shift=origin).cpu().numpy()
    assert np.allclose(spot8.numpy(), dev_sums, rtol=1e-11, atol=1e-9)
    ref = rec.hit[-1].cpu().numpy() - np.asarray(origin)[:, None]
    assert np.allclose(dev_sums[:3], ref.sum(axis=1), rtol=1e-11, atol=1e-8)
    assert dev_sums[3] == n
    assert np.allclose(dev_sums[4:7], (ref ** 2).sum(axis=1), rtol=1e-11, atol=1e-9)
    # centroid / rms with the reference's normalisations, against the oracle
    import pyrate_np as onp
    (c, rms) = engine.spot_from_sums(dev_sums, origin)
    full = rec.hit[-1].cpu().numpy()
    assert np.allclose(c, onp.centroid(full), rtol=1e-12, atol=1e-12)
    assert np.isclose(rms, onp.rms_spot(full, onp.centroid(full)), rtol=1e-10)


ANISO_TAGS = ["c4_anisotropic", "c4_anisotropic_split", "x4_biaxial", "x5_degenerate",
              "x8_crystal_mirror"]


@pytest.mark.parametrize("tag", ANISO_TAGS)
def test_birefringent_matches_reference_fixture(tag):
    """Anisotropic doublet / biaxial lens: o/e ray split (ray doubling or path
    forking), complex k and E, Poynting-vector propagation inside the crystal."""
    g = util.load_golden(tag)
    name = util.config_of(tag)
    splitup = bool(g["splitup"])
    paths = _device_paths(name, g["x0"], g["k0"], g["E0"], splitup=splitup)
    ref_paths = util.golden_paths(g)
    assert len(paths) == len(ref_paths)
    # forked paths: the reference's path order depends on the same arbitrary
    # mode order, so match whole paths by their final k
    remaining = list(range(len(ref_paths)))
    for (ip, path) in enumerate(paths):
        bundles = path.raybundles
        last = bundles[-1].numpy()
        cost = [np.nanmax(np.abs(last["k"] - ref_paths[j][-1]["k"])) +
                np.nanmax(np.abs(last["x"] - ref_paths[j][-1]["x"])) for j in remaining]
        j = remaining.pop(int(np.argmin(cost)))
        rpath = ref_paths[j]
        assert len(bundles) == len(rpath)
        for (ib, (b, rb)) in enumerate(zip(bundles, rpath)):
            d = b.numpy()
            d["E"] = d["Efield"]
            iscomplex = np.iscomplexobj(rb["k"])
            assert np.iscomplexobj(d["k"]) == iscomplex, "dtype of k, bundle %d" % ib
            # E is an eigenpolarisation only right after a crystal deflection;
            # after an isotropic one it is an arbitrary null vector in the reference
            low = path.record.lowered
            from_crystal = ib >= 2 and low[ib - 2].is_aniso_deflect
            # isotropic tensors make every mode a double root of the quartic:
            # both engines are then limited to ~sqrt(eps) and E is arbitrary
            degenerate = tag.startswith("x5")
            if iscomplex:
                util.compare_birefringent_bundle(d, rb, 1e-6 if degenerate else 1e-9,
                                                 "%s p%d b%d" % (tag, ip, ib),
                                                 check_e=from_crystal and not degenerate)
            else:
                util.compare_bundle(d, rb, 1e-10, "%s p%d b%d" % (tag, ip, ib))


def test_birefringent_larger_bundle_against_oracle():
    import pyrate_np as onp
    spec = configs.CONFIGS["c4_anisotropic"]
    (x0, k0, e0) = configs.config_bundle(spec, 7, (0.0, np.sin(0.02), np.cos(0.02)),
                                         (1.0, 0.0, 0.0))
    paths = _device_paths("c4_anisotropic", x0, k0, e0)
    ref = onp.seqtrace(onp.system_from_spec(spec), x0, k0, e0, wave=configs.DLINE)
    for (ib, (b, rb)) in enumerate(zip(paths[0].raybundles, ref[0])):
        d = b.numpy()
        d["E"] = d["Efield"]
        rbd = {"x": rb["x"], "k": rb["k"], "valid": rb["valid"], "rayID": rb["rayID"],
               "E": rb["E"]}
        if np.iscomplexobj(rb["k"]):
            low = paths[0].record.lowered
            util.compare_birefringent_bundle(d, rbd, 1e-9, "c4 b%d" % ib,
                                             check_e=ib >= 2 and low[ib - 2].is_aniso_deflect)
        else:
            util.compare_bundle(d, rbd, 1e-10, "c4 b%d" % ib)


@pytest.mark.parametrize("name", ["c1_doublet", "x1_tilted", "x3_vignette", "c5_grin",
                                  "c4_anisotropic", "x4_biaxial", "x8_crystal_mirror"])
def test_plugin_calls_reproduce_the_fused_trace(name):
    """The reference's plugin-level calls -- Material.propagate(bundle, surface)
    (mutates), Material.refract / reflect(bundle, surface) (fresh bundle) --
    driven by a hand-written copy of the per-surface loop
    (optical_element.py:336-375) must give the same rays as the fused launch."""
    spec = configs.CONFIGS[name]
    (s, seq) = configs.build_system(spec, pb.api())
    (x0, k0, e0) = configs.config_bundle(spec, 6)
    fused = s.seqtrace(pb.RayBundle(x0, k0, e0, wave=configs.DLINE), seq)[0].raybundles
    elem = s.elements["stdelem"]
    background = s.material_background
    current = background
    bundle = pb.RayBundle(x0, k0, e0, wave=configs.DLINE)
    stepwise = [bundle]
    for (surfkey, opts) in seq[0][1]:
        surface = elem.surfaces[surfkey]
        (mn, pn) = elem.annotations["surf_mat_connection"][surfkey]
        mnmat = elem.materials.get(mn, background)
        pnmat = elem.materials.get(pn, background)
        current.propagate(bundle, surface)
        mirror = opts.get("is_mirror", False)
        if not mirror:
            current = elem.findoutWhichMaterial(mnmat, pnmat, current)
        (bundle,) = (current.reflect if mirror else current.refract)(bundle, surface)
        stepwise.append(bundle)
    tol = util.tolerance_of(name)
    assert len(stepwise) == len(fused) - 1
    for (ib, (a, b)) in enumerate(zip(stepwise, fused[1:])):
        (da, db) = (a.numpy(), b.numpy())
        assert np.array_equal(da["rayID"], db["rayID"]), ib
        assert np.array_equal(da["valid"][-1], db["valid"][-1]), ib
        v = db["valid"][-1]
        assert util.relerr(da["x"][-1][:, v], db["x"][-1][:, v]) <= tol, ib
        assert util.relerr(da["x"][0], db["x"][0]) <= tol, ib
        assert util.relerr(da["k"][0], db["k"][0]) <= tol, ib
        assert np.iscomplexobj(da["k"]) == np.iscomplexobj(db["k"]), ib
        if np.iscomplexobj(db["k"]):
            # same kernel, same mode order: the fields agree column by column
            assert util.relerr(da["Efield"][0], db["Efield"][0]) <= 1e-9, ib
            assert a.splitted == b.splitted, ib


@pytest.mark.parametrize("name", ["c4_anisotropic", "x8_crystal_mirror"])
def test_anisotropic_plugin_refract_against_oracle(name):
    """AnisotropicMaterial.refract / reflect as stand-alone calls (material_anisotropic.py
    :70-155), doubled and with `splitup`, against the oracle's restatement: children of a
    ray are matched as an unordered pair (the mode order is LAPACK-arbitrary there)."""
    import pyrate_np as onp
    spec = configs.CONFIGS[name]
    (s, seq) = configs.build_system(spec, pb.api())
    deg = np.pi / 180.0
    (x0, k0, e0) = configs.config_bundle(spec, 3, (0., np.sin(deg), np.cos(deg)), (1., 0., 0.))
    elem = s.elements["stdelem"]
    osys = onp.system_from_spec(spec)
    # up to the first crystal surface with plugin calls
    bundle = pb.RayBundle(x0, k0, e0, wave=configs.DLINE)
    ob = onp.new_bundle(x0.copy(), k0.copy(), e0.copy())
    (stop_key, front_key) = (seq[0][1][0][0], seq[0][1][1][0])
    bg = s.material_background
    bg.propagate(bundle, elem.surfaces[stop_key])
    (bundle,) = bg.refract(bundle, elem.surfaces[stop_key])
    bg.propagate(bundle, elem.surfaces[front_key])
    onp.material_propagate(osys["background"], ob, osys["steps"][0])
    (ob,) = onp.material_deflect(osys["background"], ob, osys["steps"][0], configs.DLINE,
                                 False, False)
    onp.material_propagate(osys["background"], ob, osys["steps"][1])
    crystal = elem.materials[spec["surfaces"][1]["mat"]]
    ocrystal = osys["steps"][1]["mat_plus"]
    for mirror in (False, True):
        call = crystal.reflect if mirror else crystal.refract
        (doubled,) = call(bundle, elem.surfaces[front_key])
        (odoubled,) = onp.material_deflect(ocrystal, ob, osys["steps"][1], configs.DLINE,
                                           mirror, False)
        d = doubled.numpy()
        d["E"] = d["Efield"]
        assert doubled.splitted and d["x"].shape[2] == 2 * x0.shape[1]
        util.compare_birefringent_bundle(d, odoubled, 1e-9, "%s doubled mirror=%s" % (name, mirror))
        pair = call(bundle, elem.surfaces[front_key], splitup=True)
        assert len(pair) == 2
        both = {f: np.concatenate([p.numpy()[f] for p in pair], axis=-1)
                for f in ("x", "k", "Efield", "valid", "rayID")}
        both["E"] = both["Efield"]
        util.compare_birefringent_bundle(both, odoubled, 1e-9,
                                         "%s split mirror=%s" % (name, mirror))


def test_device_resident_unaligned_bundle_uses_plain_loads():
    """Caller-owned CUDA tensors with an odd ray count (rows not 16-byte aligned)
    are traced in place through the plain-load path; host arrays go through the
    row-padded TMA-staged path.  Both must agree bit for bit."""
    import torch
    from pyrate_b200 import engine, lowering
    spec = configs.CONFIGS["x1_tilted"]
    (x0, k0, e0) = configs.config_bundle(spec, 20)              # 1261 rays (odd)
    (s, seq) = configs.build_system(spec, pb.api())
    lowered = lowering.lower(s, seq, configs.DLINE)
    rec_host = engine.trace(lowered, x0, k0, e0, configs.DLINE)
    dev = rec_host.hit[0].device
    (xd, kd, ed) = (torch.from_numpy(a).to(dev) for a in (x0, k0, e0))
    assert xd.stride(0) % 2 == 1
    rec_dev = engine.trace(lowered, xd, kd, ed, configs.DLINE)
    assert rec_dev.x0.data_ptr() == xd.data_ptr()               # used in place
    for s_ in range(len(lowered)):
        assert torch.equal(rec_host.flags[s_], rec_dev.flags[s_])
        assert torch.equal(torch.nan_to_num(rec_host.hit[s_]), torch.nan_to_num(rec_dev.hit[s_]))
        assert torch.equal(torch.nan_to_num(rec_host.k[s_]), torch.nan_to_num(rec_dev.k[s_]))


def test_ray_bundle_analysis_on_device_bundles():
    """RayBundleAnalysis (reference analysis/ray_analysis.py:44-170) on traced
    CUDA bundles: native spot sums and torch path integrals against NumPy."""
    import pyrate_np as onp
    from pyrate_b200.raytracer.analysis.ray_analysis import RayBundleAnalysis
    spec = configs.CONFIGS["x3_vignette"]            # some rays are dropped on the way
    (x0, k0, e0) = configs.config_bundle(spec, 25)
    paths = _device_paths("x3_vignette", x0, k0, e0)
    ref = onp.seqtrace(onp.system_from_spec(spec), x0, k0, e0)[0]
    last = paths[0].raybundles[-1]
    ra = RayBundleAnalysis(last)
    xr = ref[-1]["x"][-1]
    assert xr.shape[1] > 10 and last.x.shape[2] == xr.shape[1]
    c = ra.get_centroid_position().numpy()
    assert np.allclose(c, onp.centroid(xr), rtol=1e-11, atol=1e-12)
    assert np.isclose(ra.get_rms_spot_size_centroid(), onp.rms_spot(xr, onp.centroid(xr)),
                      rtol=1e-10)
    mid = paths[0].raybundles[3]
    got = RayBundleAnalysis(mid).get_arc_length().cpu().numpy()
    xm = ref[3]["x"]
    want = np.sqrt(((xm[1:] - xm[:-1]) ** 2).sum(axis=1)).sum(axis=0)
    v = ref[3]["valid"][-1]
    assert np.allclose(got[v], want[v], rtol=1e-10)


def test_long_sequences_are_chunked_across_launches():
    """More sequence entries than one kernel parameter block holds (40) and more
    aspheres than auxiliary records (10): the engine continues from the last
    recorded state in further launches; results equal the oracle."""
    import pyrate_np as onp
    surfaces = [configs._conic("stop", 0.0, opt={"is_stop": True})]
    for i in range(30):
        if i % 2 == 0:
            surfaces.append({"name": "a%d" % i, "lc": {"decz": 1.5},
                             "shape": ("Asphere", {"curv": 0.01 * (1 if i % 4 == 0 else -1),
                                                   "cc": -0.5, "coefficients": [0.0, 1e-6]}),
                             "aperture": None, "mat": "g" if i % 4 == 0 else None, "opt": {}})
        else:
            surfaces.append(configs._conic("c%d" % i, 1.0, curv=0.004 * (-1) ** i,
                                           mat=None if i % 4 == 1 else "g"))
    for i in range(24):
        surfaces.append(configs._conic("p%d" % i, 0.5))
    spec = {"name": "long", "surfaces": surfaces,
            "materials": {"g": ("ConstantIndexGlass", {"n": 1.5})},
            "bundle": {"rings": 6, "radius": 3.0, "z0": -1.0}}
    (x0, k0, e0) = configs.config_bundle(spec)
    (s, seq) = configs.build_system(spec, pb.api())
    assert len(seq[0][1]) == 55
    paths = s.seqtrace(pb.RayBundle(x0, k0, e0, wave=configs.DLINE), seq)
    ref = onp.seqtrace(onp.system_from_spec(spec), x0, k0, e0, wave=configs.DLINE)
    assert len(paths[0].raybundles) == len(ref[0]) == 57
    for (ib, (b, rb)) in enumerate(zip(paths[0].raybundles, ref[0])):
        util.compare_bundle(b.numpy(), {"x": rb["x"], "k": rb["k"], "valid": rb["valid"],
                                        "rayID": rb["rayID"]}, 1e-6, "long b%d" % ib)


def test_raytrace_convenience_function():
    """pyrateoptics.raytrace (reference __init__.py:457-465): aim a collimated
    bundle with a raster, trace it, read the spot -- against the oracle."""
    import pyrate_np as onp
    from pyrate_b200.sampling2d import raster
    spec = configs.CONFIGS["c1_doublet"]
    (s, seq) = configs.build_system(spec, pb.api())
    props = {"radius": 11.43, "startz": -5.0, "raster": raster.HexGrid(), "anglex": 0.01}
    res = pb.raytrace(s, seq, 200, props, wave=configs.DLINE)
    path = res[0][0]
    osa = pb.OpticalSystemAnalysis(s, seq)
    (o, k, e) = osa.collimated_bundle(200, props, wave=configs.DLINE)
    ref = onp.seqtrace(onp.system_from_spec(spec), o, k, e, wave=configs.DLINE)[0]
    last = path.raybundles[-1].numpy()
    assert util.relerr(last["x"], ref[-1]["x"]) < 1e-10
    (xy, rms) = osa.get_spot(path)
    assert np.isclose(rms, onp.rms_spot(ref[-1]["x"][-1], onp.centroid(ref[-1]["x"][-1])),
                      rtol=1e-9)
    assert xy.shape[0] == 2


_random_spec = util.random_spec


@pytest.mark.parametrize("seed,explicit", [(i, False) for i in range(12)] +
                         [(i, True) for i in range(1, 16, 2)])
def test_random_systems_match_oracle(seed, explicit):
    """Random decentred / tilted chains (and, with `explicit`, mild aspheres / XY
    polynomials in them -- the same systems tests/test_oracle_golden.py checks the oracle
    on against the live reference)."""
    import pyrate_np as onp
    spec = _random_spec(seed, explicit=explicit)
    rng = np.random.default_rng(1000 + seed)
    kdir = rng.normal(size=3) * 0.03
    kdir[2] = 1.0
    kdir /= np.linalg.norm(kdir)
    efield = rng.normal(size=3)                      # generic E0: Poynting first segment
    (x0, k0, e0) = configs.config_bundle(spec, None, tuple(kdir), tuple(efield))
    (s, seq) = configs.build_system(spec, pb.api())
    paths = s.seqtrace(pb.RayBundle(x0, k0, e0, wave=configs.DLINE), seq)
    ref = onp.seqtrace(onp.system_from_spec(spec), x0, k0, e0, wave=configs.DLINE)
    assert len(paths[0].raybundles) == len(ref[0])
    for (ib, (b, rb)) in enumerate(zip(paths[0].raybundles, ref[0])):
        util.compare_bundle(b.numpy(), {"x": rb["x"], "k": rb["k"], "valid": rb["valid"],
                                        "rayID": rb["rayID"]},
                            util.TOL_ITERATED if explicit else util.TOL_CLOSED_FORM,
                            "%s b%d" % (spec["name"], ib))


def test_raypath_analysis_on_device_records():
    """RayPathAnalysis over the lazily materialised device bundles of a traced path."""
    import os
    import test_ray_analysis_api as tra
    g = np.load(os.path.join(util.GOLDEN, "pathanalysis.npz"))
    paths = _device_paths("c2_doublegauss", g["x0"], g["k0"], g["E0"])
    tra._check_path_analysis(paths[0], g)


@pytest.mark.parametrize("n", [1, 2, 7, 1000, 1001, 70001])
def test_spot_sums_vector_and_scalar_paths(n):
    """pyr_spot_sums: rows that allow 128-bit loads (padded records) and rows that do not
    (odd leading dimension), with and without a flag mask, against NumPy."""
    import torch
    from pyrate_b200 import engine
    rng = np.random.default_rng(n)
    xh = rng.normal(size=(3, n)) * 3.0 + np.array([[1.0], [-2.0], [50.0]])
    fh = rng.integers(0, 4, n).astype(np.uint8)
    shift = [0.9, -2.1, 49.0]
    dev = torch.device("cuda", 0)
    ld = (n + 15) // 16 * 16
    padded = torch.zeros((3, ld), dtype=torch.float64, device=dev)
    padded[:, :n] = torch.from_numpy(xh).to(dev)
    fpad = torch.zeros((ld,), dtype=torch.uint8, device=dev)
    fpad[:n] = torch.from_numpy(fh).to(dev)
    tight = torch.from_numpy(xh).to(dev)                         # ld = n
    ftight = torch.from_numpy(fh).to(dev)
    for (x, f) in ((padded[:, :n], fpad[:n]), (tight, ftight), (padded[:, :n], None),
                   (tight, None)):
        got = engine.spot_sums(x, f, shift=shift).cpu().numpy()
        m = np.ones(n, dtype=bool) if f is None else (fh & 2) != 0
        d = xh[:, m] - np.asarray(shift)[:, None]
        assert got[3] == m.sum()
        assert np.allclose(got[0:3], d.sum(1), rtol=1e-12, atol=1e-9)
        assert np.allclose(got[4:7], (d * d).sum(1), rtol=1e-12, atol=1e-9)


FDC = (0.4861e-3, 0.5876e-3, 0.6563e-3)          # Fraunhofer F, d, C lines in mm


@pytest.mark.parametrize("name", ["x14_dispersive", "x15_dispersive_asphere", "x1_tilted"])
def test_wavelength_batch_one_launch(name):
    """OpticalSystem.seqtrace_batch: F / d / C bundles of different sizes through a
    dispersive system in ONE launch (per-ray selection of the media indices) against
    one seqtrace per bundle (bit for bit for closed-form shapes: same kernel arithmetic)
    and against the oracle at each wavelength."""
    import pyrate_np as onp
    import torch
    spec = configs.CONFIGS[name]
    (s, seq) = configs.build_system(spec, pb.api())
    deg = np.pi / 180.0
    bundles = []
    raw = []
    for (i, wave) in enumerate(FDC):
        (x0, k0, e0) = configs.config_bundle(spec, 5 + 2 * i,       # 91, 169, 271 rays: segment
                                             (0., np.sin((i - 1) * deg), np.cos((i - 1) * deg)),
                                             (1., 0., 0.))          # borders inside a tile
        raw.append((x0, k0, e0))
        bundles.append(pb.RayBundle(x0, k0, e0, wave=wave))
    batch = s.seqtrace_batch(bundles, seq)
    assert len(batch) == 3
    rec = batch[0][0].record
    assert batch[1][0].record.hit[0].data_ptr() != rec.hit[0].data_ptr()      # column views
    assert batch[1][0].record.hit[0].data_ptr() - rec.hit[0].data_ptr() == 8 * 91
    tol = util.tolerance_of(name)
    saw_dispersion = False
    for (i, (wave, paths)) in enumerate(zip(FDC, batch)):
        single = s.seqtrace(bundles[i], seq)
        assert len(paths) == len(single) == 1
        assert len(paths[0].raybundles) == len(single[0].raybundles)
        (x0, k0, e0) = raw[i]
        ref = onp.seqtrace(onp.system_from_spec(spec), x0, k0, e0, wave=wave)
        for (ib, (a, b, rb)) in enumerate(zip(paths[0].raybundles, single[0].raybundles, ref[0])):
            assert a.wave == wave
            (da, db) = (a.numpy(), b.numpy())
            for f in ("x", "k", "valid", "rayID"):
                if tol == util.TOL_ITERATED and f in ("x", "k"):
                    # Newton leaves the loop by a warp vote: the iteration count of a ray
                    # depends on its warp neighbours, which differ between the two launches
                    assert util.relerr(da[f], db[f]) < 1e-13, (i, ib, f)
                else:
                    assert np.array_equal(da[f], db[f], equal_nan=f in ("x", "k")), (i, ib, f)
            util.compare_bundle(da, {"x": rb["x"], "k": rb["k"], "valid": rb["valid"],
                                     "rayID": rb["rayID"]}, tol, "%s wave %d b%d" % (name, i, ib))
        if i:
            last = paths[0].raybundles[-1].numpy()["k"]
            saw_dispersion = saw_dispersion or not np.allclose(
                last[..., :5], batch[0][0].raybundles[-1].numpy()["k"][..., :5], atol=1e-9)
    assert saw_dispersion


@pytest.mark.parametrize("name", ["c5_grin", "x11_gridsag", "x12_combination"])
def test_wavelength_batch_falls_back_where_the_kernel_cannot_batch(name):
    """Crystals / GRIN media / grid-sag and combination shapes: one launch per bundle,
    same results as seqtrace."""
    spec = configs.CONFIGS[name]
    (s, seq) = configs.build_system(spec, pb.api())
    bundles = [pb.RayBundle(*configs.config_bundle(spec, 2), wave=w) for w in FDC[:2]]
    batch = s.seqtrace_batch(bundles, seq)
    for (b, paths) in zip(bundles, batch):
        single = s.seqtrace(b, seq)
        (da, db) = (paths[0].raybundles[-1].numpy(), single[0].raybundles[-1].numpy())
        assert np.array_equal(da["x"], db["x"], equal_nan=True)


# ---------------------------------------------------------------------------
# BASELINE configs 3-5 at (or near) their stated sizes
# ---------------------------------------------------------------------------
def test_full_size_asphere_strided_subsample():
    """C3 at BASELINE size (9 997 351 rays through the even asphere, Newton intersect):
    a strided subsample against the oracle at every surface, every ray alive, |k| = n."""
    import torch
    import pyrate_np as onp
    spec = configs.CONFIGS["c3_asphere"]
    (x0, k0, e0) = configs.config_bundle(spec)
    assert x0.shape[1] == 9997351
    paths = _device_paths("c3_asphere", x0, k0, e0)
    rec = paths[0].record
    nidx = [ls.st.after.n for ls in rec.lowered]
    for (s, k) in enumerate(rec.k):
        alive = (rec.flags[s] & 2) != 0
        assert bool(alive.all())
        assert float((torch.sqrt((k * k).sum(0)) - nidx[s]).abs().max()) < 1e-12
    sub = np.arange(0, x0.shape[1], 1009)                       # 9 909 rays
    ref = onp.seqtrace(onp.system_from_spec(spec), x0[:, sub], k0[:, sub], e0[:, sub])
    worst = 0.0
    for s in range(len(rec.hit)):
        ex = util.relerr(rec.hit[s][:, sub].cpu().numpy(), ref[0][s + 1]["x"][-1])
        ek = util.relerr(rec.k[s][:, sub].cpu().numpy(), ref[0][s + 2]["k"][0])
        worst = max(worst, ex, ek)
        assert ex < util.TOL_ITERATED and ek < util.TOL_ITERATED, (s, ex, ek)
    print("c3 full size: worst rel err %.3e" % worst)
    assert worst < 1e-11          # measured ~1e-15; the contract is 1e-6


def test_grin_1e5_rays_against_oracle_including_last_grin_row():
    """C5 with 99 919 rays (hexapolar R = 182) against the oracle with the per-ray energy
    test: every bundle, hit points and k after every surface -- including k at the end of
    the GRIN segment (the frozen p/n of the last integrator row)."""
    import pyrate_np as onp
    spec = configs.CONFIGS["c5_grin"]
    (x0, k0, e0) = configs.config_bundle(spec, 182)
    assert x0.shape[1] == 99919
    paths = _device_paths("c5_grin", x0, k0, e0)
    ref = onp.seqtrace(onp.system_from_spec(spec), x0, k0, e0, wave=configs.DLINE,
                       per_ray_energy=True, history=False)
    assert len(paths[0].raybundles) == len(ref[0])
    for (ib, (b, rb)) in enumerate(zip(paths[0].raybundles, ref[0])):
        d = b.numpy()
        rbd = {"x": rb["x"], "k": rb["k"], "valid": rb["valid"], "rayID": rb["rayID"]}
        if rb["x"].shape[0] == 3:
            # oracle rows of the GRIN bundle: start, last integrator row, intersection;
            # device bundle (history off): start, intersection -- and the NEXT bundle's
            # k is the refraction of the last integrator row's k (checked there)
            rbd = {"x": rb["x"][[0, 2]], "k": rb["k"][[0, 0]], "valid": rb["valid"][[0, 2]],
                   "rayID": rb["rayID"]}
        util.compare_bundle(d, rbd, 1e-9, "c5 1e5 b%d" % ib)


def test_grin_last_row_k_matches_reference_fixture():
    """k at the LAST GRIN row of the reference fixture (material_grin.py:198-199, p/n of
    the frozen state) against the device history rows."""
    g = util.load_golden("c5_grin")
    spec = configs.CONFIGS["c5_grin"]
    (s, seq) = configs.build_system(spec, pb.api())
    path = s.seqtrace(pb.RayBundle(g["x0"], g["k0"], g["E0"], wave=configs.DLINE), seq,
                      grin_history=True)[0].raybundles
    rpath = util.golden_paths(g)[0]
    seen = 0
    for (b, rb) in zip(path, rpath):
        if rb["rows"] <= 3:
            continue
        seen += 1
        d = b.numpy()
        v = rb["valid"][-1].astype(bool)
        assert util.relerr(d["k"][-1][:, v], rb["k"][-1][:, v]) < 1e-9
        assert util.relerr(d["k"][-2][:, v], rb["k"][-2][:, v]) < 1e-9
    assert seen == 1


def test_grin_full_shard_strided_subsample():
    """One GPU's shard of BASELINE config 5 (1e8 rays over 8 GPUs = 12.5e6 rays; here the
    central 12.5e6 rays of the R = 5773 raster would need the host to build 1e8 points, so
    the test uses R = 2041: 12 501 127 rays): strided subsample against the oracle."""
    import pyrate_np as onp
    spec = configs.CONFIGS["c5_grin"]
    (x0, k0, e0) = configs.config_bundle(spec, 2041)
    paths = _device_paths("c5_grin", x0, k0, e0)
    rec = paths[0].record
    sub = np.arange(0, x0.shape[1], 6007)                       # 2 082 rays
    ref = onp.seqtrace(onp.system_from_spec(spec), x0[:, sub], k0[:, sub], e0[:, sub],
                       wave=configs.DLINE, per_ray_energy=True, history=False)[0]
    # ref bundles: [b0, b0, b1, b2(GRIN: 3 rows), b3, b4]; valid rays are compacted, so
    # compare through rayID
    for s in range(len(rec.hit)):
        rb = ref[s + 1]
        ids = rb["rayID"]
        hit = rec.hit[s][:, sub].cpu().numpy()[:, ids]
        v = rb["valid"][-1]
        assert util.relerr(hit[:, v], rb["x"][-1][:, v]) < 1e-9, s
        fl = rec.flags[s][sub].cpu().numpy()
        assert np.array_equal((fl[ids] & 1) != 0, v), s
        nb = ref[s + 2]
        k = rec.k[s][:, sub].cpu().numpy()[:, nb["rayID"]]
        assert util.relerr(k, nb["k"][0]) < 1e-9, s


def test_birefringent_1e4_rays_against_oracle():
    """C4 with 10 267 rays (R = 58) -> 41 068 output rays against the oracle."""
    import pyrate_np as onp
    spec = configs.CONFIGS["c4_anisotropic"]
    (x0, k0, e0) = configs.config_bundle(spec, 58, (0.0, np.sin(0.01), np.cos(0.01)),
                                         (1.0, 0.0, 0.0))
    assert x0.shape[1] == 10267
    paths = _device_paths("c4_anisotropic", x0, k0, e0)
    ref = onp.seqtrace(onp.system_from_spec(spec), x0, k0, e0, wave=configs.DLINE)
    # vectorised comparison: children of a ray are an unordered set -> sort the children
    # of every parent by (Re kx, Re ky, Re kz) on both sides
    for (ib, (b, rb)) in enumerate(zip(paths[0].raybundles, ref[0])):
        d = b.numpy()
        assert d["x"].shape == rb["x"].shape, ib
        if not np.iscomplexobj(rb["k"]):
            util.compare_bundle(d, {"x": rb["x"], "k": rb["k"], "valid": rb["valid"],
                                    "rayID": rb["rayID"]}, 1e-10, "c4 1e4 b%d" % ib)
            continue

        def order(ids, k, x):
            key = np.round(k[0].real / 1e-7) * 1e-7      # modes differ by >> 1e-7 in k
            return np.lexsort((x[-1][0], key[2], key[1], key[0], ids))
        (og, orf) = (order(d["rayID"], d["k"], d["x"]), order(rb["rayID"], rb["k"], rb["x"]))
        assert np.array_equal(d["rayID"][og], rb["rayID"][orf]), ib
        assert np.array_equal(d["valid"][:, og], rb["valid"][:, orf]), ib
        assert util.relerr(d["x"][:, :, og], rb["x"][:, :, orf]) < 1e-9, ib
        assert util.relerr(d["k"][:, :, og], rb["k"][:, :, orf]) < 1e-9, ib


def test_grin_without_boundary_dead_rays_do_not_spin():
    """A GRIN medium with boundary kind 'none', rays that are vignetted before they enter
    it and a ray count that is not a multiple of the tile: dead / out-of-range rays must
    not enter the integrator (NaN never satisfies its exit tests; they would run to the
    1e6-step cap).  The whole trace has to finish in well under a second."""
    import copy
    import time
    import torch
    import pyrate_np as onp
    spec = copy.deepcopy(configs.CONFIGS["c5_grin"])
    mat = spec["materials"]["grin"][1]
    mat["device_profile"] = dict(mat["device_profile"], boundary={"kind": "none", "params": []})
    mat["source"] = mat["source"].replace("return x[0]**2 + x[1]**2 < 10.**2",
                                          "return np.ones_like(x[0], dtype=bool)")
    spec["bundle"] = dict(spec["bundle"], radius=6.0)          # aperture radius is 5: vignetting
    (x0, k0, e0) = configs.config_bundle(spec, 11)             # 397 rays, not a multiple of 256
    (s, seq) = configs.build_system(spec, pb.api())
    bundle = pb.RayBundle(x0, k0, e0, wave=configs.DLINE)
    s.seqtrace(bundle, seq)                                     # warm-up (library load)
    torch.cuda.synchronize()
    t = time.perf_counter()
    paths = s.seqtrace(bundle, seq)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    assert dt < 0.5, "trace took %.3f s: dead rays spin in the GRIN integrator" % dt
    ref = onp.seqtrace(onp.system_from_spec(spec), x0, k0, e0, wave=configs.DLINE,
                       per_ray_energy=True, history=False)
    last = paths[0].raybundles[-1].numpy()
    assert 0 < last["x"].shape[2] < x0.shape[1]
    assert np.array_equal(last["rayID"], ref[0][-1]["rayID"])
    assert util.relerr(last["x"], ref[0][-1]["x"]) < 1e-9


# ---------------------------------------------------------------------------
# reference-exact GRIN mode (lock-step loop, bundle-summed energy test)
# ---------------------------------------------------------------------------
def test_grin_lockstep_reproduces_the_reference_fixture_rows():
    """grin_lockstep + grin_history: the device follows material_grin.py:106-213 literally;
    number of rows, rows [0, P-2, P-1] of x, k and `valid` of the UNMODIFIED reference."""
    g = util.load_golden("c5_grin")
    spec = configs.CONFIGS["c5_grin"]
    (s, seq) = configs.build_system(spec, pb.api())
    bundle = pb.RayBundle(g["x0"], g["k0"], g["E0"], wave=configs.DLINE)
    path = s.seqtrace(bundle, seq, grin_history=True, grin_lockstep=True)[0].raybundles
    rpath = util.golden_paths(g)[0]
    assert len(path) == len(rpath)
    seen = 0
    for (ib, (b, rb)) in enumerate(zip(path, rpath)):
        d = b.numpy()
        assert d["x"].shape[0] == rb["rows"], (ib, d["x"].shape, rb["rows"])
        if rb["rows"] > 3:
            seen += 1
            rows = [0, rb["rows"] - 2, rb["rows"] - 1]
            d = {"x": d["x"][rows], "k": d["k"][rows], "valid": d["valid"][rows], "rayID": d["rayID"]}
        util.compare_bundle(d, rb, 1e-9, "lockstep vs fixture b%d" % ib)
    assert seen == 1


@pytest.mark.parametrize("case", ["summed_energy", "late_boundary", "plain"])
def test_grin_lockstep_cross_ray_semantics_match_the_oracle(case):
    """The two couplings between rays that only the lock-step loop has: (a) the energy test
    is summed over the bundle -- with a tolerance that no single ray violates but the sum
    does, EVERY ray becomes invalid (material_grin.py:164-176); (b) a ray that has already
    reached the surface keeps stepping while slower rays finish and is invalidated when it
    then leaves the boundary (:189-190).  Against the oracle's literal restatement."""
    import copy
    import pyrate_np as onp
    spec = copy.deepcopy(configs.CONFIGS["c5_grin"])
    mat = spec["materials"]["grin"][1]
    (rings, kdir) = (9, (0.0, 0.0, 1.0))                         # 271 rays
    if case == "summed_energy":
        # measured on this bundle: largest single-ray violation 1.2e-4, largest bundle sum 1.7e-3
        mat["energyviolation"] = 5e-4
    elif case == "late_boundary":
        # a strongly tilted exit surface (rays finish up to ~20 iterations apart), a bundle
        # travelling obliquely in y and a box boundary |y| < 2.5: rays that finish early and
        # keep stepping drift out of the box while the others are still on their way
        # (found with the oracle: 20 valid rays in lock-step, 22 when integrated independently)
        mat["device_profile"] = dict(mat["device_profile"],
                                     boundary={"kind": "box", "params": [50.0, 2.5]})
        mat["source"] = mat["source"].replace("return x[0]**2 + x[1]**2 < 10.**2",
                                              "return (np.abs(x[0]) < 50.0) & (np.abs(x[1]) < 2.5)")
        spec["surfaces"][2]["lc"]["tiltx"] = -0.5
        spec["surfaces"][2]["aperture"] = None
        spec["bundle"] = dict(spec["bundle"], radius=1.5)
        (rings, kdir) = (3, (0.0, np.sin(0.05), np.cos(0.05)))
    (x0, k0, e0) = configs.config_bundle(spec, rings, kdir, (1.0, 0.0, 0.0))
    (s, seq) = configs.build_system(spec, pb.api())
    bundle = pb.RayBundle(x0, k0, e0, wave=configs.DLINE)
    lock = s.seqtrace(bundle, seq, grin_lockstep=True)[0].raybundles
    ref = onp.seqtrace(onp.system_from_spec(spec), x0, k0, e0, wave=configs.DLINE,
                       per_ray_energy=False, history=False)[0]
    assert len(lock) == len(ref)
    for (ib, (b, rb)) in enumerate(zip(lock, ref)):
        d = b.numpy()
        rbd = {"x": rb["x"], "k": rb["k"], "valid": rb["valid"], "rayID": rb["rayID"]}
        if rb["x"].shape[0] == 3:          # oracle GRIN bundle: start, frozen state, intersection
            rbd = {"x": rb["x"][[0, 2]], "k": rb["k"][[0, 0]], "valid": rb["valid"][[0, 2]],
                   "rayID": rb["rayID"]}
        util.compare_bundle(d, rbd, 1e-9, "%s b%d" % (case, ib))
    free = s.seqtrace(bundle, seq)[0].raybundles               # per-ray normalisation
    n_lock = lock[-1].numpy()["x"].shape[2]
    n_free = free[-1].numpy()["x"].shape[2]
    if case == "summed_energy":
        assert n_lock == 0 and n_free == x0.shape[1]            # sum violates, no single ray does
    elif case == "late_boundary":
        assert 0 < n_lock < n_free                              # early finishers were invalidated
    else:
        assert n_lock == n_free == x0.shape[1]


def _evanescent_spec():
    """Dense glass -> tilted calcite-like crystal: the in-plane wave vector of a ray fan
    straddles both critical angles (n_glass sin(theta) from 1.1 to 1.8 against n_o = 1.658,
    n_e = 1.486), so one or both crystal modes are evanescent for part of the bundle."""
    eps = np.diag([1.658 ** 2, 1.486 ** 2, 1.658 ** 2]).tolist()
    return {"name": "evanescent",
            "surfaces": [configs._conic("stop", 0.0, opt={"is_stop": True}),
                         configs._conic("glass", 2.0, mat="dense"),
                         configs._conic("crystal", 6.0, mat="crystal", tiltx=38.0 * np.pi / 180.0),
                         configs._conic("exit", 4.0, mat="dense"),
                         configs._conic("image", 6.0, mat="dense")],
            "materials": {"dense": ("ConstantIndexGlass", {"n": 2.4}),
                          "crystal": ("AnisotropicMaterial", {"epstensor": eps})},
            "bundle": {"rings": 5, "radius": 1.0, "z0": -3.0}}


@pytest.mark.parametrize("name", ["c4_anisotropic", "x8_crystal_mirror", "evanescent"])
def test_crystal_real_fast_path_equals_the_complex_code(name):
    """The crystal kernel runs a warp in real arithmetic while all its rays carry real k, E
    (csrc/pyr_aniso.cu: aniso_modes_r) and falls back to the complex code per warp and step
    for evanescent or degenerate modes.  A bundle whose E carries a 1e-30 imaginary part
    takes the complex code everywhere: both runs must agree -- also where the two kinds of
    warps sit side by side (the `evanescent` fan: propagating, half and fully evanescent
    rays in one bundle)."""
    if name == "evanescent":
        spec = _evanescent_spec()
        n = 4099
        rng = np.random.default_rng(5)
        ang = np.linspace(-0.43, 0.43, n)
        tx = rng.uniform(-0.02, 0.02, n)
        k0 = np.stack((np.sin(tx), np.cos(tx) * np.sin(ang), np.cos(tx) * np.cos(ang)))
        x0 = np.stack((rng.uniform(-0.5, 0.5, n), rng.uniform(-0.5, 0.5, n), np.full(n, -3.0)))
        e0 = np.cross(k0.T, np.array([0.0, 1.0, 0.3])).T
        e0 /= np.linalg.norm(e0, axis=0)
    else:
        spec = configs.CONFIGS[name]
        (x0, k0, e0) = configs.config_bundle(spec, 21, (0.0, np.sin(0.03), np.cos(0.03)), (1.0, 0.0, 0.0))
    (s, seq) = configs.build_system(spec, pb.api())
    real = s.seqtrace(pb.RayBundle(x0, k0, e0, wave=configs.DLINE), seq)[0].raybundles
    cplx = s.seqtrace(pb.RayBundle(x0, k0, e0 * (1.0 + 1e-30j), wave=configs.DLINE), seq)[0].raybundles
    assert len(real) == len(cplx)
    evanescent = 0
    for (ib, (a, b)) in enumerate(zip(real, cplx)):
        (da, db) = (a.numpy(), b.numpy())
        assert np.array_equal(da["valid"], db["valid"]), ib
        assert np.array_equal(da["rayID"], db["rayID"]), ib
        for key in ("x", "k"):
            (u, v) = (np.asarray(da[key]), np.asarray(db[key]))
            assert u.shape == v.shape
            assert np.array_equal(np.isnan(u.real), np.isnan(v.real)), (ib, key)
            m = ~np.isnan(u.real)
            scale = max(np.max(np.abs(v[m])), 1e-300) if m.any() else 1.0
            assert np.max(np.abs(u[m] - v[m])) <= 1e-11 * scale if m.any() else True, (ib, key)
            if key == "k" and np.iscomplexobj(u):
                evanescent += int(np.sum(np.abs(u[m].imag) > 1e-6))
    if name == "evanescent":
        assert evanescent > 100          # the fan really has evanescent modes


def test_large_launches_on_concurrent_streams_are_independent():
    """Large launches of the conic-only kernels take their tiles in order from an atomic cursor
    (8 bytes from the library's stream-ordered pool, csrc/pyr_trace.cu launch()).  Two such traces
    in flight at the same time on different streams -- and a third one right behind on the first
    stream -- must each see every tile exactly once: all records equal a lone run's."""
    import torch
    from pyrate_b200 import engine, lowering
    spec = configs.CONFIGS["c2_doublegauss"]
    (x0, k0, e0) = configs.config_bundle(spec, 1020)            # 3 124 261 rays: 6 103 tiles
    (s, seq) = configs.build_system(spec, pb.api())
    low = lowering.lower(s, seq, configs.DLINE)
    dev = torch.device("cuda", 0)
    (x0, k0, e0) = engine.device_bundle(x0, k0, e0, dev)
    ref = engine.trace(low, x0, k0, e0, configs.DLINE, device=dev)
    torch.cuda.synchronize()
    (s1, s2) = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
    recs = []
    for rep in range(3):
        with torch.cuda.stream(s1):
            recs.append(engine.trace(low, x0, k0, e0, configs.DLINE, device=dev))
        with torch.cuda.stream(s2):
            recs.append(engine.trace(low, x0, k0, e0, configs.DLINE, device=dev))
    torch.cuda.synchronize()
    for rec in recs:
        for key in ("hit", "k", "flags"):
            for (a, b) in zip(getattr(rec, key), getattr(ref, key)):
                assert torch.equal(torch.nan_to_num(a) if a.dtype.is_floating_point else a,
                                   torch.nan_to_num(b) if b.dtype.is_floating_point else b), key
        assert int((rec.flags[-1] & 2).ne(0).sum()) == x0.shape[1]


@pytest.mark.parametrize("name,rings,odd", [("c2_doublegauss", 120, False), ("c2_doublegauss", 1020, False),
                                            ("x3_vignette", 150, False), ("c1_doublet", 77, True),
                                            ("c3_asphere", 120, False), ("c5_grin", 30, False)])
def test_trace_with_spot_sums_equals_the_separate_reduction(name, rings, odd):
    """engine.trace(..., spot=...) = pyr_trace_spot without read-back: the conic-only kernels
    accumulate the spot sums of the last entry themselves (POLICY bit 32), everything else falls
    back to a pyr_spot_sums launch behind the trace.  Either way the sums equal the separate
    reduction over the last record -- with dead rays (x3), a generated bundle, rows that are not
    16-byte aligned (odd ray count: plain loads) and large launches (in-order tiles)."""
    import torch
    from pyrate_b200 import bundlegen, engine, lowering
    spec = configs.CONFIGS[name]
    (x0, k0, e0) = configs.config_bundle(spec, rings)
    if odd and x0.shape[1] % 2 == 0:
        (x0, k0, e0) = (x0[:, :-1], k0[:, :-1], e0[:, :-1])
    (s, seq) = configs.build_system(spec, pb.api())
    low = lowering.lower(s, seq, configs.DLINE)
    dev = torch.device("cuda", 0)
    origin = engine.last_surface_origin(low)
    if odd:
        (xd, kd, ed) = (torch.as_tensor(np.ascontiguousarray(a), device=dev) for a in (x0, k0, e0))
    else:
        (xd, kd, ed) = engine.device_bundle(x0, k0, e0, dev)
    fused = torch.full((8,), 7.0, dtype=torch.float64, device=dev)          # overwritten, not accumulated
    rec = engine.trace(low, xd, kd, ed, configs.DLINE, device=dev, spot=(fused, origin))
    ref = torch.zeros(8, dtype=torch.float64, device=dev)
    engine.spot_sums(rec.hit[-1], rec.flags[-1], out=ref, shift=origin)
    (a, b) = (fused.cpu().numpy(), ref.cpu().numpy())
    assert a[3] == b[3] and b[3] > 0
    assert np.allclose(a, b, rtol=1e-11, atol=1e-9 * b[3])
    if name == "c2_doublegauss":
        gen = bundlegen.config_generator(spec, rings)
        g = torch.zeros(8, dtype=torch.float64, device=dev)
        grec = engine.trace(low, None, None, None, configs.DLINE, device=dev, gen=gen, spot=(g, origin))
        gref = torch.zeros(8, dtype=torch.float64, device=dev)
        engine.spot_sums(grec.hit[-1], grec.flags[-1], out=gref, shift=origin)
        assert np.allclose(g.cpu().numpy(), gref.cpu().numpy(), rtol=1e-11, atol=1e-9 * b[3])
