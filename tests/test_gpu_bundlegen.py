"""GPU: bundle generation on the device (reference
raytracer/analysis/optical_system_analysis.py:83-165 over sampling2d/raster.py:36-166).
The generator kernel against the NumPy rasters, the fused generate-and-trace launch
against the trace of the same rays read from memory, and the host entry with a
generator / with every record."""
import numpy as np
import pytest

import pyrate_b200 as pb
from pyrate_b200 import _native as nat
from pyrate_b200 import bundlegen, configs

import util

pytestmark = pytest.mark.gpu


def _specs(n):
    return {"hexapolar": bundlegen.hexapolar_spec(configs.rings_for(n)),
            "rect": bundlegen.rect_spec(n), "hex": bundlegen.hex_spec(n),
            "circular": bundlegen.circular_spec(n),
            "circular_sqrt": bundlegen.circular_spec(n, requidistant=False)}


@pytest.mark.parametrize("n", [7, 1000, 123457])
@pytest.mark.parametrize("kind", ["hexapolar", "rect", "hex", "circular", "circular_sqrt"])
def test_generated_bundles_match_the_numpy_rasters(kind, n):
    """pyr_generate_bundle against the host construction (reference formulas): lattice
    coordinates bit for bit, trigonometric rasters to a few ulp, whole bundle and a shard."""
    spec = _specs(n)[kind]
    lattice = kind in ("rect", "hex")
    for (bundle, radius, start, direction, efield) in (
            (nat.BUNDLE_COLLIMATED, 11.43, (0.1, -0.2, -5.0),
             (np.sin(-0.01) * np.cos(0.02), np.sin(0.02), np.cos(-0.01) * np.cos(0.02)), None),
            (nat.BUNDLE_COLLIMATED, 5.0, (0.0, 0.0, 0.0), (0.0, 0.0, 1.0), (0.0, 1.0, 0.0)),
            (nat.BUNDLE_DIVERGENT, 0.2, (0.0, 0.3, -50.0), (0.0, 0.01, 0.0), None)):
        gen = bundlegen.BundleGen(spec, bundle, radius, start, direction, efield, 1.0003)
        (x, k, e) = (t.cpu().numpy() for t in gen.materialise())
        (hx, hk, he) = gen.arrays_host()
        assert x.shape == hx.shape == (3, spec.total)
        if lattice and bundle == nat.BUNDLE_COLLIMATED:
            assert np.array_equal(x, hx) and np.array_equal(k, hk)
        else:
            # sin / cos on the device (sincospi of the reduced argument for the hexapolar
            # raster) against libm: a few ulp of the pupil radius
            assert np.max(np.abs(x - hx)) <= 2e-15 * max(1.0, radius, abs(start[2]))
            assert np.max(np.abs(k - hk)) <= 2e-15
        assert np.max(np.abs(e - he)) <= 1e-15
        assert np.max(np.abs(np.sum(e * k, axis=0))) < 1e-15
        (lo, hi) = (spec.total // 3, spec.total - 1)
        sh = gen.shard(lo, hi)
        (sx, sk, se) = (t.cpu().numpy() for t in sh.materialise())
        assert np.array_equal(sx, x[:, lo:hi]) and np.array_equal(sk, k[:, lo:hi])


def test_rect_raster_of_1e7_points_matches_numpy():
    """BASELINE-size lattice: the row table + binary search against the O(n^2) NumPy mask."""
    from pyrate_b200.sampling2d.raster import RectGrid
    n = 9997351
    gen = bundlegen.BundleGen(bundlegen.rect_spec(n), radius=5.0)
    (x, k, e) = gen.materialise()
    (px, py) = RectGrid().getGrid(n)
    assert x.shape[1] == px.size
    assert np.array_equal(x[0].cpu().numpy(), 5.0 * px + 0.0)
    assert np.array_equal(x[1].cpu().numpy(), 5.0 * py + 0.0)


@pytest.mark.parametrize("name,rings,shard", [("c2_doublegauss", 200, None), ("c2_doublegauss", 33, (100, 2931)),
                                              ("c1_doublet", 40, None), ("c3_asphere", 150, None),
                                              ("x1_tilted", 60, (7, 9000)), ("x3_vignette", 50, None),
                                              ("c5_grin", 30, None), ("c5_grin", 30, (11, 2000))])
def test_fused_generation_equals_the_trace_of_the_same_rays_from_memory(name, rings, shard):
    """PyrRaysIn.gen: the trace kernel generates the rays in its prologue.  Records must be
    bit-identical to tracing the materialised arrays (same x0, k0, E0 bit for bit; the
    arithmetic downstream is the same code), and the generated rays must equal the host
    construction of configs.config_bundle."""
    import torch
    from pyrate_b200 import engine, lowering
    spec = configs.CONFIGS[name]
    deg = np.pi / 180.0
    kdir = (0.0, np.sin(0.5 * deg), np.cos(0.5 * deg))
    gen = bundlegen.config_generator(spec, rings, kdir, (1.0, 0.0, 0.0))
    (hx, hk, he) = configs.config_bundle(spec, rings, kdir, (1.0, 0.0, 0.0))
    if shard is not None:
        gen = gen.shard(*shard)
        (hx, hk, he) = (a[:, shard[0]:shard[1]] for a in (hx, hk, he))
    (s, seq) = configs.build_system(spec, pb.api())
    lowered = lowering.lower(s, seq, configs.DLINE)
    rec_gen = engine.trace(lowered, None, None, None, configs.DLINE, gen=gen)
    assert rec_gen.gen is gen and rec_gen._x0 is None and not gen.materialised   # nothing was written out
    (x0, k0, e0) = gen.materialise()
    assert np.max(np.abs(x0.cpu().numpy() - hx)) <= 3e-15 * max(1.0, spec["bundle"]["radius"])
    assert np.array_equal(k0.cpu().numpy(), hk)
    rec_mem = engine.trace(lowered, x0, k0, e0, configs.DLINE)
    newton = util.tolerance_of(name) == util.TOL_ITERATED and name != "c5_grin"
    for s_ in range(len(lowered)):
        assert torch.equal(rec_gen.flags[s_], rec_mem.flags[s_]), s_
        for (a, b) in ((rec_gen.hit[s_], rec_mem.hit[s_]), (rec_gen.k[s_], rec_mem.k[s_])):
            if newton:      # warp-voted iteration count: neighbours differ between the kernels
                assert util.relerr(a.cpu().numpy(), b.cpu().numpy()) < 1e-13, s_
            else:
                assert torch.equal(torch.nan_to_num(a), torch.nan_to_num(b)), s_


def test_aim_describes_the_bundle_and_seqtrace_generates_it():
    """OpticalSystemAnalysis.aim + trace (reference :167-260): the initial bundle is a
    generator; seqtrace launches the fused kernel; the path's first bundle materialises on
    demand and everything matches the oracle fed with collimated_bundle's arrays."""
    import pyrate_np as onp
    from pyrate_b200.sampling2d import raster
    spec = configs.CONFIGS["c1_doublet"]
    (s, seq) = configs.build_system(spec, pb.api())
    for (rast, bundletype, props) in (
            (raster.RectGrid(), "collimated", {"radius": 11.43, "startz": -5.0, "anglex": 0.01}),
            (raster.HexGrid(), "collimated", {"radius": 9.0, "startz": -5.0, "angley": -0.01}),
            (raster.CircularGrid(), "divergent", {"radius": 0.05, "startz": -150.0}),
            (raster.HexapolarGrid(), "collimated", {"radius": 11.0, "startz": -5.0})):
        props = dict(props, raster=rast)
        osa = pb.OpticalSystemAnalysis(s, seq)
        osa.aim(3000, props, bundletype=bundletype, wave=configs.DLINE)
        ib = osa.initial_bundles[0]
        assert ib.generator is not None and not ib.generator.materialised
        path = osa.trace()[0][0]
        assert not ib.generator.materialised            # traced without ever being written out
        make = osa.collimated_bundle if bundletype == "collimated" else osa.divergent_bundle
        (o, k, e) = make(3000, props, wave=configs.DLINE)
        ref = onp.seqtrace(onp.system_from_spec(spec), o, k, e, wave=configs.DLINE)[0]
        assert len(path.raybundles) == len(ref)
        for (ib_, (b, rb)) in enumerate(zip(path.raybundles, ref)):
            util.compare_bundle(b.numpy(), {"x": rb["x"], "k": rb["k"], "valid": rb["valid"],
                                            "rayID": rb["rayID"]}, 1e-10,
                                "%s b%d" % (type(rast).__name__, ib_))
        first = path.raybundles[0].numpy()
        assert np.max(np.abs(first["x"][0] - o)) < 1e-13 and np.max(np.abs(first["k"][0] - k)) < 1e-15


@pytest.mark.parametrize("chunk", [4000, 4096])
def test_host_entry_with_generator_and_with_every_record(chunk):
    """pyr_trace_host_io: (a) generator instead of host arrays -- same last record as the
    device-resident trace of the same generator; (b) all_records: every sequence entry's
    hit points / wave vectors / flags in host memory, equal to the device records."""
    import torch
    from pyrate_b200 import engine, lowering
    spec = configs.CONFIGS["c2_doublegauss"]
    gen = bundlegen.config_generator(spec, 70)            # 14 911 rays
    n = gen.n
    (s, seq) = configs.build_system(spec, pb.api())
    lowered = lowering.lower(s, seq, configs.DLINE)
    rec = engine.trace(lowered, None, None, None, configs.DLINE, gen=gen)
    ht = engine.HostTracer(lowered, n, chunk_rays=chunk)
    (xl, kl, fl, spot8) = ht(gen=gen)
    assert ht.h2d_bytes < 64 * 1024
    assert np.array_equal(fl.numpy(), rec.flags[-1].cpu().numpy())
    assert np.array_equal(xl.numpy(), rec.hit[-1].cpu().numpy())
    assert np.array_equal(kl.numpy(), rec.k[-1].cpu().numpy())
    assert spot8[3] == n
    hta = engine.HostTracer(lowered, n, chunk_rays=chunk, all_records=True)
    (x0, k0, e0) = (t.cpu().contiguous().pin_memory() for t in gen.materialise())
    for kw in ({"gen": gen}, {"x0": x0, "k0": k0, "e0": e0}):
        hta.x_all.zero_()
        hta.k_all.zero_()
        hta.flags_all.zero_()
        hta(**kw)
        for s_ in range(len(lowered)):
            assert np.array_equal(hta.x_all[s_].numpy(), rec.hit[s_].cpu().numpy()), s_
            assert np.array_equal(hta.k_all[s_].numpy(), rec.k[s_].cpu().numpy()), s_
            assert np.array_equal(hta.flags_all[s_].numpy(), rec.flags[s_].cpu().numpy()), s_
    assert hta.d2h_bytes == 49 * n * 14 + 64


def test_host_entry_continues_long_sequences():
    """More entries than one launch carries (40 steps / 10 auxiliary records): the host entry
    chains launches through its record buffers; last record equals the engine's."""
    import torch
    from pyrate_b200 import engine, lowering
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
            "bundle": {"rings": 20, "radius": 3.0, "z0": -1.0}}
    (x0, k0, e0) = configs.config_bundle(spec)
    (s, seq) = configs.build_system(spec, pb.api())
    lowered = lowering.lower(s, seq, configs.DLINE)
    assert len(lowered) == 55
    rec = engine.trace(lowered, x0, k0, e0, configs.DLINE)
    n = x0.shape[1]
    for all_records in (False, True):
        ht = engine.HostTracer(lowered, n, chunk_rays=512, all_records=all_records)
        (xp, kp, ep) = (torch.from_numpy(a).pin_memory() for a in (x0, k0, e0))
        (xl, kl, fl, _) = ht(xp, kp, ep)
        assert np.array_equal(fl.numpy(), rec.flags[-1].cpu().numpy())
        assert util.relerr(xl.numpy(), rec.hit[-1].cpu().numpy()) < 1e-13
        assert util.relerr(kl.numpy(), rec.k[-1].cpu().numpy()) < 1e-13
        if all_records:
            for s_ in (0, 17, 39, 40, 54):
                assert util.relerr(ht.x_all[s_].numpy(), rec.hit[s_].cpu().numpy()) < 1e-13, s_


def test_spot_points_compaction_on_device():
    """pyr_spot_points: (x, y) of the surviving rays, compacted without a host round trip;
    as a point SET equal to the NumPy selection, in global and in last-surface coordinates
    (OpticalSystemAnalysis.get_spot, reference :283-303)."""
    import torch
    from pyrate_b200 import distributed as pd
    from pyrate_b200 import engine, lowering
    spec = configs.CONFIGS["x3_vignette"]                       # rays are dropped on the way
    (x0, k0, e0) = configs.config_bundle(spec, 60)              # 10 981 rays
    (s, seq) = configs.build_system(spec, pb.api())
    lowered = lowering.lower(s, seq, configs.DLINE)
    rec = engine.trace(lowered, x0, k0, e0, configs.DLINE)
    (hit, fl) = (rec.hit[-1], rec.flags[-1])
    keep = ((fl & 2) != 0).cpu().numpy()
    assert 0 < keep.sum() < keep.size
    want = hit.cpu().numpy()[:2][:, keep]

    def as_set(a):
        return a[:, np.lexsort((a[1], a[0]))]
    (xy, count) = engine.spot_points(hit, fl)
    assert int(count) == keep.sum()
    assert np.array_equal(as_set(xy[:, :int(count)].cpu().numpy()), as_set(want))
    frame = lowered[-1].st.shape_frame
    (xyl, count2) = engine.spot_points(hit, fl, frame=frame)
    o = np.array(list(frame.o))
    r = np.array(list(frame.r)).reshape(3, 3)
    local = (r.T @ (hit.cpu().numpy() - o[:, None]))[:2][:, keep]
    got = as_set(xyl[:, :int(count2)].cpu().numpy())
    assert np.max(np.abs(got - as_set(local))) < 1e-12
    # too narrow a buffer: counted, not stored beyond the width
    (xy3, count3) = engine.spot_points(hit, fl, width=100)
    assert int(count3) == keep.sum() and xy3.shape == (2, 100)
    # single-process "gather": same container the multi-rank path returns
    pts = pd.gather_spot_points(hit, fl)
    assert pts.xy.shape[0] == 1 and np.array_equal(as_set(pts.points().cpu().numpy()), as_set(want))
    # odd sizes and no flags
    for n in (1, 255, 256, 257, 1000):
        (xy4, c4) = engine.spot_points(hit[:, :n])
        assert int(c4) == n
        assert np.array_equal(as_set(xy4[:, :n].cpu().numpy()), as_set(hit[:2, :n].cpu().numpy()))


@pytest.mark.parametrize("name,chunk", [("c4_anisotropic", 1 << 20), ("c4_anisotropic", 1000),
                                        ("x8_crystal_mirror", 777), ("x4_biaxial", 512)])
def test_host_entry_traces_crystal_sequences(name, chunk):
    """pyr_trace_host_io with birefringent media: real host arrays (or a generator) in, the
    last record of the DOUBLED bundle back (x, complex k and E, flags in the reference's
    hstack order) -- equal to the device-resident trace, chunked or not."""
    import torch
    from pyrate_b200 import engine, lowering
    spec = configs.CONFIGS[name]
    (x0, k0, e0) = configs.config_bundle(spec, 20, (0.0, np.sin(0.01), np.cos(0.01)), (1.0, 0.0, 0.0))
    n = x0.shape[1]                                            # 1261 rays
    (s, seq) = configs.build_system(spec, pb.api())
    lowered = lowering.lower(s, seq, configs.DLINE)
    rec = engine.trace(lowered, x0, k0, e0, configs.DLINE)
    ht = engine.HostTracer(lowered, n, chunk_rays=chunk)
    assert ht.crystal and ht.mult_k == rec.k[-1].shape[1] // n and ht.mult_x == rec.hit[-1].shape[1] // n
    (xp, kp, ep) = (torch.from_numpy(a).pin_memory() for a in (x0, k0, e0))
    (xl, kl, fl, spot8) = ht(xp, kp, ep)
    assert np.array_equal(fl.numpy(), rec.flags[-1].cpu().numpy())
    assert np.array_equal(np.nan_to_num(xl.numpy()), np.nan_to_num(rec.hit[-1].cpu().numpy()))
    assert np.array_equal(np.nan_to_num(kl.numpy()), np.nan_to_num(rec.k[-1].cpu().numpy()))
    assert np.array_equal(np.nan_to_num(ht.e_last.numpy()), np.nan_to_num(rec.e[-1].cpu().numpy()))
    assert spot8[3] == int(((rec.flags[-1] & 2) != 0).sum())
    gen = bundlegen.config_generator(spec, 20, (0.0, np.sin(0.01), np.cos(0.01)), (1.0, 0.0, 0.0))
    (xg, kg, fg, _) = ht(gen=gen)
    assert np.array_equal(fg.numpy(), rec.flags[-1].cpu().numpy())
    assert util.relerr(np.nan_to_num(xg.numpy()), np.nan_to_num(rec.hit[-1].cpu().numpy())) < 1e-13
