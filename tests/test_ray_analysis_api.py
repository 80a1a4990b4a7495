"""CPU: known-answer tests of the reference (tests/test_ray_analysis.py:34-112)
against the RayBundleAnalysis mirror."""
import math

import numpy as np

from pyrate_b200.raytracer.analysis.ray_analysis import RayBundleAnalysis
from pyrate_b200.raytracer.ray import RayBundle

X5 = np.array([[1, 0, 0, 1, 2], [0, 1, 0, 1, 2], [0, 0, 1, 1, 2]], dtype=float)


def test_centroid_and_rms():
    rb = RayBundle(x0=X5, k0=np.zeros((3, 5)), Efield0=np.zeros((3, 5)))
    ra = RayBundleAnalysis(rb)
    assert np.allclose(ra.get_centroid_position().numpy(), 4. / 5.)
    assert np.isclose(ra.get_rms_spot_size(np.array([0, 0, 0])), math.sqrt(18.0 / 4.0))


def test_arc_length():
    (k0, e0) = (np.zeros((3, 2)), np.zeros((3, 2)))
    rb = RayBundle(x0=np.zeros((3, 2)), k0=k0, Efield0=e0)
    valid = np.ones(2, dtype=bool)
    for x in ([[1, 0], [0, 0], [0, 0]], [[1, 1], [1, 1], [0, 0]],
              [[0, 2], [1, 2], [0, 0]], [[0, 3], [0, 3], [0, 0]]):
        rb.append(np.array(x, dtype=float), k0, e0, valid)
    assert np.allclose(RayBundleAnalysis(rb).get_arc_length().numpy(),
                       np.array([4., 3 * np.sqrt(2)]))


def test_direction_centroid_and_angular_size():
    k0 = np.zeros((3, 5))
    k0[2, :] = 1
    e0 = np.zeros((3, 5))
    e0[1, :] = 1.
    ra = RayBundleAnalysis(RayBundle(x0=X5, k0=k0, Efield0=e0))
    assert np.allclose(ra.get_centroid_direction().numpy(), np.array([0, 0, 1]))
    ang = ra.get_rms_angluar_size(np.array([math.sin(math.pi / 180.0), 0,
                                            math.cos(math.pi / 180.0)]))
    assert np.isclose(ang, math.pi / 180.0)


def test_spot_matches_reference_fixture():
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "spot.npz"))
    n = g["x"].shape[1]
    ra = RayBundleAnalysis(RayBundle(x0=g["x"], k0=np.zeros((3, n)), Efield0=np.zeros((3, n))))
    c = ra.get_centroid_position().numpy()
    assert np.allclose(c, g["centroid"], rtol=1e-13, atol=1e-15)
    assert np.isclose(ra.get_rms_spot_size(c), float(g["rms"]), rtol=1e-12)
    assert np.isclose(ra.get_rms_spot_size_centroid(), float(g["rms"]), rtol=1e-12)


def _path_from_oracle(x0, k0, e0):
    """A RayPath of host bundles with the reference's structure, served by the oracle."""
    import pyrate_np as onp
    import torch
    from pyrate_b200 import configs
    from pyrate_b200.raytracer.ray import RayPath
    ref = onp.seqtrace(onp.system_from_spec(configs.CONFIGS["c2_doublegauss"]), x0, k0, e0,
                       wave=configs.DLINE)[0]
    made = {}
    path = RayPath()
    for b in ref:
        if id(b) not in made:
            made[id(b)] = RayBundle(_lazy={
                "x": torch.from_numpy(b["x"]), "k": torch.from_numpy(b["k"]),
                "Efield": torch.from_numpy(b["E"]), "valid": torch.from_numpy(b["valid"]),
                "rayID": torch.from_numpy(b["rayID"])}, wave=configs.DLINE)
        path.appendRayBundle(made[id(b)])
    return path


def _check_path_analysis(path, g):
    from pyrate_b200.raytracer.analysis.ray_analysis import RayPathAnalysis
    from pyrate_b200 import configs
    rpa = RayPathAnalysis(path)
    assert np.allclose(rpa.get_arc_length().cpu().numpy(), g["arc"], rtol=1e-12)
    assert np.allclose(rpa.get_phase_difference().cpu().numpy(), g["phase"], rtol=1e-12)
    assert np.allclose(rpa.get_arc_length(first=2, last=9).cpu().numpy(), g["arc_2_9"],
                       rtol=1e-12)
    rel = rpa.get_relative_phase_difference(referenceray=0, wavelength=configs.DLINE)
    assert np.allclose(rel.cpu().numpy(), g["rel"], rtol=1e-9, atol=1e-6)   # in waves


def test_raypath_analysis_matches_reference_fixture():
    """RayPathAnalysis (ray_analysis.py:170-213): arc length and optical phase summed
    over the bundles of a path (the duplicated hand-over bundle counts twice, as in
    the reference)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                             "pathanalysis.npz"))
    _check_path_analysis(_path_from_oracle(g["x0"], g["k0"], g["E0"]), g)
