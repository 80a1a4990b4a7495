"""User-defined GRIN index functions (reference: Python source in IsotropicGrinMaterial,
core/functionobject.py:99-119, material/material_grin.py:79-104): CUDA expressions in
annotations["device_source"], compiled at run time by NVRTC (pyrate_b200/grin_jit.py)."""
import copy

import numpy as np
import pytest

import pyrate_b200 as pb
from pyrate_b200 import _native as nat
from pyrate_b200 import configs, grin_jit, lowering

import util


def test_user_profile_compiles_for_sm_100a_without_a_gpu():
    src = configs.CONFIGS["x17_user_grin"]["materials"]["rod"][1]["device_source"]
    (key, cubin) = grin_jit.compile_cubin(src)
    assert cubin[:4] == b"\x7fELF" and len(key) == 40
    assert grin_jit.compile_cubin(src)[1] is cubin                    # cached
    bad = dict(src, n="p[0] +* 2")
    with pytest.raises(grin_jit.GrinJitError, match="does not compile"):
        grin_jit.compile_cubin(bad)
    with pytest.raises(grin_jit.GrinJitError):
        grin_jit.render_source({"dndx": "0.0"})


def test_user_profile_lowers_as_user_grin_medium():
    spec = configs.CONFIGS["x17_user_grin"]
    (s, seq) = configs.build_system(spec, pb.api())
    low = lowering.lower(s, seq, configs.DLINE)
    assert low[1].st.after.kind == nat.MEDIUM_ISO_GRIN and low[1].st.after.grin_profile == nat.GRIN_USER
    assert low[2].st.before.grin_profile == nat.GRIN_USER and low[2].st.before.grin_ds == 0.05
    # the C library refuses to integrate a profile it does not know (checked without a GPU)
    import ctypes
    rays = nat.PyrRaysIn()
    dummy = (ctypes.c_double * 8)()
    (rays.x, rays.k) = (ctypes.addressof(dummy), ctypes.addressof(dummy))
    steps = (nat.PyrStep * 1)(low[2].st)
    assert nat.load().pyr_trace(steps, 1, ctypes.byref(rays), 1, 0, None) == nat.E_UNSUPPORTED
    steps = (nat.PyrStep * 1)(low[1].st)                              # entrance without per-ray index
    assert nat.load().pyr_trace(steps, 1, ctypes.byref(rays), 1, 0, None) == nat.E_UNSUPPORTED


@pytest.mark.gpu
@pytest.mark.parametrize("rings", [6, 40])
def test_user_grin_trace_matches_oracle(rings):
    """The sech / tapered rod (outside the device catalogue) through the NVRTC kernels against
    the oracle running the material's Python source."""
    import pyrate_np as onp
    spec = configs.CONFIGS["x17_user_grin"]
    deg = np.pi / 180.0
    (x0, k0, e0) = configs.config_bundle(spec, rings, (0.0, np.sin(0.5 * deg), np.cos(0.5 * deg)),
                                         (1.0, 0.0, 0.0))
    (s, seq) = configs.build_system(spec, pb.api())
    paths = s.seqtrace(pb.RayBundle(x0, k0, e0, wave=configs.DLINE), seq)
    ref = onp.seqtrace(onp.system_from_spec(spec), x0, k0, e0, wave=configs.DLINE,
                       per_ray_energy=True, history=False)[0]
    assert len(paths[0].raybundles) == len(ref)
    for (ib, (b, rb)) in enumerate(zip(paths[0].raybundles, ref)):
        d = b.numpy()
        rbd = {"x": rb["x"], "k": rb["k"], "valid": rb["valid"], "rayID": rb["rayID"]}
        if rb["x"].shape[0] == 3:
            rbd = {"x": rb["x"][[0, 2]], "k": rb["k"][[0, 0]], "valid": rb["valid"][[0, 2]],
                   "rayID": rb["rayID"]}
        util.compare_bundle(d, rbd, 1e-9, "user grin b%d" % ib)
    assert paths[0].raybundles[-1].numpy()["x"].shape[2] > 0.5 * x0.shape[1]


@pytest.mark.gpu
def test_user_grin_source_is_verified_against_the_python_functions():
    spec = copy.deepcopy(configs.CONFIGS["x17_user_grin"])
    src = spec["materials"]["rod"][1]["device_source"]
    src["params"] = [1.55, 0.13, 0.002]                      # not the medium of the Python source
    (x0, k0, e0) = configs.config_bundle(spec, 3)
    (s, seq) = configs.build_system(spec, pb.api())
    with pytest.raises(grin_jit.GrinJitError, match="disagrees"):
        s.seqtrace(pb.RayBundle(x0, k0, e0, wave=configs.DLINE), seq)
