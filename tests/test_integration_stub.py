"""GPU: the reference-side binding of INTEGRATION.md section 2, executed VERBATIM.

The code block a pyrate maintainer would paste into pyrateoptics/raytracer/optical_system.py
is read out of INTEGRATION.md, attached to the reference's own OpticalSystem class (the
unmodified package: /root/reference in the build container, the copy staged by
oracle/make_ref.sh on the GPU box) and run on the reference's object graph; every surface's
hit points and wave vectors are compared with what the reference's seqtrace computes for
the same bundle on the CPU."""
import os
import re

import numpy as np
import pytest

from pyrate_b200 import configs

import util

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stub_source():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    (stub,) = [b for b in blocks if "def seqtrace_b200" in b]
    return stub


@pytest.mark.parametrize("name,rings", [("c2_doublegauss", 12), ("c1_doublet", 10), ("x1_tilted", 9)])
def test_reference_object_graph_through_the_documented_stub(name, rings):
    import refshim
    if not refshim.reference_available():
        pytest.skip("no reference package here (run oracle/make_ref.sh in the build container)")
    api = refshim.api()
    from pyrateoptics.raytracer.optical_system import OpticalSystem
    ns = {"OpticalSystem": OpticalSystem}
    exec(compile(_stub_source(), "INTEGRATION.md", "exec"), ns)
    assert hasattr(OpticalSystem, "seqtrace_b200")
    spec = configs.CONFIGS[name]
    (s, seq) = configs.build_system(spec, api)                  # the REFERENCE's classes
    deg = np.pi / 180.0
    (x0, k0, e0) = configs.config_bundle(spec, rings, (0., np.sin(deg), np.cos(deg)), (1., 0., 0.))
    bundle = api.RayBundle(x0, k0, e0, wave=configs.DLINE)
    (X, K, F) = s.seqtrace_b200(bundle, seq)                    # device, through the C ABI
    ref = s.seqtrace(api.RayBundle(x0, k0, e0, wave=configs.DLINE), seq)[0].raybundles   # CPU
    nsteps = X.shape[0]
    assert len(ref) == nsteps + 2
    (X, K, F) = (X.cpu().numpy(), K.cpu().numpy(), F.cpu().numpy())
    for s_ in range(nsteps):
        rb = ref[s_ + 1]                                        # bundle propagated to entry s_
        ids = np.asarray(rb.rayID)
        v = np.asarray(rb.valid[-1], dtype=bool)
        assert np.array_equal((F[s_][ids] & 1) != 0, v), s_
        assert util.relerr(X[s_][:, ids][:, v], np.asarray(rb.x[-1])[:, v]) < 1e-10, s_
        nb = ref[s_ + 2]                                        # bundle after the deflection
        assert util.relerr(K[s_][:, np.asarray(nb.rayID)], np.real(np.asarray(nb.k[0]))) < 1e-10, s_
