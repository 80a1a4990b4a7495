"""Cylinder shape (reference surface_shape.py:328-388).  The reference's own
Cylinder.intersect is dead code (AttributeError at :380) and its quadratic is the
rotationally symmetric conic's, so there is no reference output to pin against: PARITY
UNPINNED for this shape.  The corrected restatement in oracle/pyrate_np.py is pinned on
what must hold instead -- the surface equation, independence of x, equality with the
Conic for meridional rays (where the reference's formula is right), the sag the reference
does define -- and the device is compared with that oracle."""
import numpy as np
import pytest

import pyrate_b200 as pb
from pyrate_b200 import configs

import util

import pyrate_np as onp


def _trace(spec, x0, k0, e0):
    return onp.seqtrace(onp.system_from_spec(spec), x0, k0, e0, wave=configs.DLINE)[0]


def test_oracle_cylinder_hits_lie_on_the_surface_and_ignore_x():
    spec = configs.CONFIGS["x16_cylinder"]
    osys = onp.system_from_spec(spec)
    d = np.array([0.03, -0.02, 1.0])
    d /= np.linalg.norm(d)
    (x0, k0, e0) = configs.config_bundle(spec, 9, tuple(d), (0., 1., 0.))
    ref = _trace(spec, x0, k0, e0)
    for s in (1, 2):
        sh = osys["steps"][s]["shape"]
        hit = onp.g2l_pts(sh["frame"], ref[s + 1]["x"][-1])
        (c, cc) = (sh["curv"], sh["cc"])
        assert np.max(np.abs(c * (hit[1] ** 2 + (1 + cc) * hit[2] ** 2) - 2 * hit[2])) < 1e-13
        assert np.allclose(hit[2], onp.shape_sag(sh, hit[0], hit[1]), atol=1e-13)
        nrm = onp.shape_normal(sh, hit[0], hit[1])
        assert np.all(nrm[0] == 0) and np.allclose(np.sum(nrm ** 2, axis=0), 1)
    # translating the bundle along the extrusion axis of an UNTILTED cylinder changes nothing
    flat = {"name": "cyl", "surfaces": [configs._conic("stop", 0.0, opt={"is_stop": True}),
                                        {"name": "c", "lc": {"decz": 3.0},
                                         "shape": ("Cylinder", {"curv": 0.04, "cc": -0.3}),
                                         "aperture": None, "mat": "g", "opt": {}},
                                        configs._conic("image", 20.0)],
            "materials": {"g": ("ConstantIndexGlass", {"n": 1.6})},
            "bundle": {"rings": 5, "radius": 4.0, "z0": -2.0}}
    (x0, k0, e0) = configs.config_bundle(flat, 5, tuple(d), (0., 1., 0.))
    a = _trace(flat, x0, k0, e0)
    shift = x0.copy()
    shift[0] += 1.7
    b = _trace(flat, shift, k0, e0)
    assert np.allclose(b[-1]["x"][-1] - a[-1]["x"][-1], np.array([[1.7], [0.], [0.]]), atol=1e-12)
    assert np.allclose(b[-1]["k"][-1], a[-1]["k"][-1], atol=1e-14)


def test_oracle_cylinder_equals_conic_for_meridional_rays():
    """Rays in the plane x = 0 with d_x = 0: the cylinder and the rotationally symmetric
    conic (closed form pinned on the reference) are the same curve."""
    def spec(kind):
        shape = ("Cylinder", {"curv": 1. / 25.0, "cc": -0.6}) if kind == "Cylinder" else None
        surf = {"name": "s", "lc": {"decz": 3.0}, "shape": shape, "aperture": None, "mat": "g",
                "opt": {}} if shape else configs._conic("s", 3.0, curv=1. / 25.0, cc=-0.6, mat="g")
        return {"name": kind, "surfaces": [configs._conic("stop", 0.0, opt={"is_stop": True}), surf,
                                           configs._conic("image", 30.0)],
                "materials": {"g": ("ConstantIndexGlass", {"n": 1.7})},
                "bundle": {"rings": 1, "radius": 1.0, "z0": -2.0}}
    y = np.linspace(-8, 8, 41)
    x0 = np.vstack((np.zeros_like(y), y, np.full_like(y, -2.0)))
    k0 = np.repeat(np.array([[0.], [np.sin(0.05)], [np.cos(0.05)]]), y.size, axis=1)
    e0 = np.repeat(np.array([[1.], [0.], [0.]]), y.size, axis=1)
    (a, b) = (_trace(spec("Cylinder"), x0, k0, e0), _trace(spec("Conic"), x0, k0, e0))
    for (ba, bb) in zip(a, b):
        assert np.allclose(ba["x"], bb["x"], atol=1e-13) and np.allclose(ba["k"], bb["k"], atol=1e-14)


def test_host_mirror_cylinder_sag_and_gradient():
    lc = pb.LocalCoordinates.p(name="cyl_lc")
    cyl = pb.Cylinder.p(lc, curv=1. / 20.0, cc=-0.4)
    con = pb.Conic.p(lc, curv=1. / 20.0, cc=-0.4)
    (x, y) = (np.linspace(-3, 3, 7), np.linspace(-5, 5, 7))
    assert np.allclose(cyl.getSag(x, y), con.getSag(0 * x, y))            # reference :360-367
    h = 1e-6
    g = cyl.getGrad(x, y)
    assert np.allclose(g[0], 0) and np.allclose(
        -g[1] / g[2], (cyl.getSag(x, y + h) - cyl.getSag(x, y - h)) / (2 * h), atol=1e-8)
    assert cyl.kind == "shape_Cylinder"


@pytest.mark.gpu
@pytest.mark.parametrize("rings", [8, 40])
def test_device_cylinder_matches_oracle(rings):
    spec = configs.CONFIGS["x16_cylinder"]
    d = np.array([0.02, 0.015, 1.0])
    d /= np.linalg.norm(d)
    (x0, k0, e0) = configs.config_bundle(spec, rings, tuple(d), (0., 1., 0.))
    (s, seq) = configs.build_system(spec, pb.api())
    paths = s.seqtrace(pb.RayBundle(x0, k0, e0, wave=configs.DLINE), seq)
    ref = _trace(spec, x0, k0, e0)
    assert len(paths[0].raybundles) == len(ref)
    for (ib, (b, rb)) in enumerate(zip(paths[0].raybundles, ref)):
        util.compare_bundle(b.numpy(), {"x": rb["x"], "k": rb["k"], "valid": rb["valid"],
                                        "rayID": rb["rayID"]}, util.TOL_CLOSED_FORM, "cyl b%d" % ib)
