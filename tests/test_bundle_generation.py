"""CPU: rasters and OpticalSystemAnalysis bundle generators against outputs of the
reference (tests/golden/rasters.npz, oracle/gen_golden.py --rasters)."""
import os

import numpy as np

import pyrate_b200 as pb
from pyrate_b200 import configs
from pyrate_b200.sampling2d import raster

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rasters.npz"))


def test_rasters_match_reference():
    for nm in ("RectGrid", "HexGrid", "MeridionalFan", "SagitalFan", "ChiefAndComa", "Single",
               "CircularGrid"):
        (x, y) = getattr(raster, nm)().getGrid(int(G[nm + "_n"]))
        assert np.allclose(x, G[nm + "_x"], atol=1e-15) and np.allclose(y, G[nm + "_y"], atol=1e-15)


def test_hexapolar_raster():
    (x, y) = raster.HexapolarGrid().getGrid(1000)
    assert x.size == 1027 and np.max(x * x + y * y) <= 1 + 1e-15
    assert np.isclose(np.hypot(x[-1], y[-1]), 1.0)


def test_bundle_generators_match_reference():
    (s, seq) = configs.build_system(configs.CONFIGS["c1_doublet"], pb.api())
    osa = pb.OpticalSystemAnalysis(s, seq)
    props = {"radius": 11.43, "startz": -5.0, "starty": 0.3, "anglex": 0.02, "angley": -0.01,
             "raster": raster.RectGrid()}
    (o, k, e) = osa.collimated_bundle(40, props, wave=configs.DLINE)
    assert np.allclose(o, G["coll_o"], atol=1e-14) and np.allclose(k, np.real(G["coll_k"]), atol=1e-13)
    assert np.allclose(np.sum(e * k, axis=0), 0, atol=1e-15) and np.allclose(np.sum(e * e, axis=0), 1)
    props = {"radius": 0.2, "startz": -50.0, "anglex": 0.01, "raster": raster.HexGrid()}
    (o, k, e) = osa.divergent_bundle(40, props, wave=configs.DLINE)
    assert np.allclose(o, G["div_o"], atol=1e-14) and np.allclose(k, np.real(G["div_k"]), atol=1e-13)
    assert np.allclose(np.sum(e * k, axis=0), 0, atol=1e-15)
    osa.aim(40, props, bundletype="divergent", wave=configs.DLINE)
    assert osa.initial_bundles[0].x.shape == (1, 3, o.shape[1])


def test_poisson_disk_sampling_properties():
    """PoissonDiskSampling (reference raster.py:106-125 + pds.py) is random: checked by
    its defining properties -- inside the unit disk, hard-core distance, sensible count."""
    import math
    import numpy as np
    from pyrate_b200.sampling2d.raster import PoissonDiskSampling
    nray = 200
    (x, y) = PoissonDiskSampling().getGrid(nray, rng=np.random.default_rng(3))
    assert np.all(x ** 2 + y ** 2 <= 1.0)
    r = 1.0 / int(round(math.sqrt(nray * 4.0 / math.pi)))
    d2 = (x[:, None] - x[None, :]) ** 2 + (y[:, None] - y[None, :]) ** 2
    np.fill_diagonal(d2, np.inf)
    assert d2.min() >= r * r * (1 - 1e-12)
    # maximal Poisson-disk packings hold 0.6-0.9 points per r^2-cell of the disk area
    assert 0.5 * math.pi / r ** 2 * 0.6 < x.size < math.pi / r ** 2


def test_device_raster_index_arithmetic_matches_the_numpy_rasters():
    """pyrate_b200.bundlegen: the (row table, linspace parameters) description the device
    kernel expands, restated in NumPy index by index (BundleGen.points_host), must give the
    rasters bit for bit -- including the rim, where the clipping test is decided by the
    rounding of x*x + y*y."""
    from pyrate_b200 import bundlegen as bg
    for n in (1, 5, 17, 100, 1000, 12345, 200000):
        for (spec_fn, rast) in ((bg.rect_spec, raster.RectGrid()), (bg.hex_spec, raster.HexGrid()),
                                (bg.circular_spec, raster.CircularGrid())):
            if n == 1 and rast.__class__ is not raster.RectGrid:
                continue
            (px, py) = bg.BundleGen(spec_fn(n)).points_host()
            (rx, ry) = rast.getGrid(n)
            assert px.shape == rx.shape, (n, type(rast).__name__)
            assert np.array_equal(px, rx) and np.array_equal(py, ry), (n, type(rast).__name__)
    (px, py) = bg.BundleGen(bg.circular_spec(5000, False)).points_host()
    (rx, ry) = raster.CircularGrid().getGrid(5000, requidistant=False)
    assert np.array_equal(px, rx) and np.array_equal(py, ry)
    for rings in (0, 1, 2, 18, 182):
        g = bg.BundleGen(bg.hexapolar_spec(rings))
        (px, py) = g.points_host()
        (rx, ry) = configs.hexapolar(rings)
        assert np.array_equal(px, rx) and np.array_equal(py, ry)
        if rings > 2:
            (a, b) = g.shard(7, 100).points_host()
            assert np.array_equal(a, rx[7:100]) and np.array_equal(b, ry[7:100])


def test_row_table_against_brute_force_mask():
    from pyrate_b200 import bundlegen as bg
    rng = np.random.default_rng(5)
    for _ in range(20):
        n = int(rng.integers(2, 400))
        xa = np.sort(rng.uniform(-1.3, 1.3, n))
        ya = rng.uniform(-1.5, 1.5, int(rng.integers(1, 50)))
        (prefix, first) = bg.raster_rows(xa, ya)
        mask = xa[None, :] ** 2 + ya[:, None] ** 2 <= 1
        assert np.array_equal(np.diff(prefix), mask.sum(axis=1))
        for (r, row) in enumerate(mask):
            if row.any():
                assert first[r] == int(np.argmax(row))


def test_generated_arrays_equal_the_config_bundles_and_the_osa_bundles():
    from pyrate_b200 import bundlegen as bg
    spec = configs.CONFIGS["c2_doublegauss"]
    (x, k, e) = bg.config_generator(spec, 30).arrays_host()
    (hx, hk, he) = configs.config_bundle(spec, 30)
    assert np.array_equal(x, hx) and np.array_equal(k, hk) and np.array_equal(e, he)
    (s, seq) = configs.build_system(configs.CONFIGS["c1_doublet"], pb.api())
    osa = pb.OpticalSystemAnalysis(s, seq)
    for (bundletype, props) in (("collimated", {"radius": 11.43, "startz": -5.0, "starty": 0.3,
                                                "anglex": 0.02, "angley": -0.01,
                                                "raster": raster.RectGrid()}),
                                ("divergent", {"radius": 0.2, "startz": -50.0, "anglex": 0.01,
                                               "raster": raster.HexGrid()})):
        gen = osa.bundle_generator(40, props, bundletype, wave=configs.DLINE)
        make = osa.collimated_bundle if bundletype == "collimated" else osa.divergent_bundle
        (o, kk, ee) = make(40, props, wave=configs.DLINE)
        (x, k, e) = gen.arrays_host()
        assert np.array_equal(x, o) and np.array_equal(k, kk) and np.allclose(e, ee, atol=1e-15)
