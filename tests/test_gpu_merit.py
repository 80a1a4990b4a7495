"""GPU: the optimiser-loop binding (pyrate_b200.merit.MeritTrace) against the general
API: same RMS spot radius / centroid as seqtrace + RayBundleAnalysis, before and after
optimisable variables change (what optimize/optimize.py:73-91 does between evaluations)."""
import numpy as np
import pytest

import pyrate_b200 as pb
from pyrate_b200 import bundlegen, configs
from pyrate_b200.merit import MeritTrace
from pyrate_b200.raytracer.analysis.ray_analysis import RayBundleAnalysis

pytestmark = pytest.mark.gpu


def _reference_merit(s, seq, bundle):
    path = s.seqtrace(bundle, seq)[0]
    ra = RayBundleAnalysis(path.raybundles[-1])
    return (ra.get_rms_spot_size_centroid(), ra.get_centroid_position().numpy(),
            path.raybundles[-1].x.shape[2])


@pytest.mark.parametrize("name,rings,generated", [("c2_doublegauss", 18, False), ("c2_doublegauss", 18, True),
                                                  ("c3_asphere", 18, True), ("x3_vignette", 20, False),
                                                  ("x1_tilted", 12, True), ("c5_grin", 6, False)])
def test_merit_trace_equals_seqtrace_plus_spot(name, rings, generated):
    spec = configs.CONFIGS[name]
    (s, seq) = configs.build_system(spec, pb.api())
    if generated:
        bundle = pb.RayBundle(generator=bundlegen.config_generator(spec, rings), wave=configs.DLINE)
    else:
        bundle = pb.RayBundle(*configs.config_bundle(spec, rings), wave=configs.DLINE)
    mt = MeritTrace(s, seq, bundle)
    elem = s.elements["stdelem"]
    rng = np.random.default_rng(3)
    seen = set()
    for it in range(4):
        (rms, cen, count) = _reference_merit(s, seq, bundle)
        (c2, rms2, count2) = mt.centroid_and_rms()
        assert count2 == count
        assert np.isclose(rms2, rms, rtol=1e-9), (it, rms2, rms)
        assert np.allclose(c2, cen, rtol=1e-10, atol=1e-10)
        assert np.isclose(mt(refresh=False), rms2, rtol=1e-15)
        seen.add(round(rms, 12))
        # an optimiser step: perturb curvatures and a thickness, update the frames
        for surf in elem.surfaces.values():
            var = getattr(surf.shape, "curvature", None)
            if var is not None and var() != 0.0:
                var.setvalue(var() * (1 + 2e-3 * rng.standard_normal()))
            for (nm, v) in getattr(surf.shape, "params", {}).items():
                if nm == "curv":
                    v.setvalue(v() * (1 + 2e-3 * rng.standard_normal()))
        last_lc = list(elem.surfaces.values())[-1].shape.lc
        last_lc.decz.setvalue(last_lc.decz() + 0.05 * rng.standard_normal())
        s.rootcoordinatesystem.update()
    assert len(seen) > 1                      # the merit function did respond to the changes
    if generated:
        assert mt.gen is not None and not mt.gen.materialised
