"""CPU: the C-ABI library loads and exports every symbol include/pyrate_b200.h
declares; struct layouts of the ctypes binding match the compiled ones; the
product fails loudly without a device (no fallback)."""
import ctypes
import os
import re

import pytest

from pyrate_b200 import _native as nat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pyrate_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pyr_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = nat.load()
    syms = _declared_symbols()
    assert set(syms) == set(nat.EXPORTS)
    for s in syms:
        assert getattr(lib, s) is not None


def test_struct_layouts_match():
    lib = nat.load()
    assert lib.pyr_sizeof_step() == ctypes.sizeof(nat.PyrStep)
    assert lib.pyr_sizeof_rays_in() == ctypes.sizeof(nat.PyrRaysIn)
    assert lib.pyr_sizeof_bundle_gen() == ctypes.sizeof(nat.PyrBundleGen)
    assert lib.pyr_version() == 4


def test_strerror():
    lib = nat.load()
    assert lib.pyr_strerror(0) == b"ok"
    assert b"bad" in lib.pyr_strerror(-1)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present")
    import numpy as np
    import pyrate_b200 as pb
    from pyrate_b200 import configs, engine
    (s, seq) = configs.build_system(configs.CONFIGS["c1_doublet"], pb.api())
    (x0, k0, e0) = configs.config_bundle(configs.CONFIGS["c1_doublet"], 2)
    with pytest.raises(engine.DeviceRequired):
        s.seqtrace(pb.RayBundle(x0, k0, e0), seq)
    assert np.array_equal(x0[2], np.full(x0.shape[1], -5.0))


def _one_step():
    st = nat.PyrStep()
    for i in range(3):
        st.shape_frame.r[i * 4] = 1.0
        st.aperture_frame.r[i * 4] = 1.0
        st.before.frame.r[i * 4] = 1.0
        st.after.frame.r[i * 4] = 1.0
    st.before.n = 1.0
    st.after.n = 1.5
    return st


def test_argument_validation_happens_before_any_device_work():
    """Structural errors come back as PYR_E_* codes (never exceptions across the
    boundary, never a launch) -- checked here without a GPU."""
    lib = nat.load()
    rays = nat.PyrRaysIn()
    dummy = (ctypes.c_double * 8)()
    rays.x = ctypes.addressof(dummy)
    rays.k = ctypes.addressof(dummy)
    steps = (nat.PyrStep * 1)(_one_step())
    assert lib.pyr_trace(steps, 0, ctypes.byref(rays), 1, 0, None) == -1          # BADARG
    assert lib.pyr_trace(None, 1, ctypes.byref(rays), 1, 0, None) == -1
    bad = (nat.PyrStep * 1)(_one_step())
    bad[0].shape_kind = 99
    assert lib.pyr_trace(bad, 1, ctypes.byref(rays), 1, 0, None) == -2            # UNSUPPORTED
    bad = (nat.PyrStep * 1)(_one_step())
    bad[0].n_coeff = 1000
    assert lib.pyr_trace(bad, 1, ctypes.byref(rays), 1, 0, None) == -1
    many = (nat.PyrStep * 41)(*[_one_step() for _ in range(41)])
    assert lib.pyr_trace(many, 41, ctypes.byref(rays), 1, 0, None) == -3          # TOOLARGE
    split_mid = (nat.PyrStep * 2)(_one_step(), _one_step())
    split_mid[0].split = 1
    assert lib.pyr_trace(split_mid, 2, ctypes.byref(rays), 1, 0, None) == -1
    aniso = (nat.PyrStep * 1)(_one_step())
    aniso[0].after.kind = nat.MEDIUM_ANISO
    assert lib.pyr_trace(aniso, 1, ctypes.byref(rays), 1, 0, None) == -2          # needs PYR_F_COMPLEX
    rays.x = None
    assert lib.pyr_trace(steps, 1, ctypes.byref(rays), 1, 0, None) == -1
    assert lib.pyr_trace(steps, 1, ctypes.byref(rays), 0, 0, None) == 0           # empty bundle
    assert lib.pyr_spot_sums(None, 0, None, 2, 1, None, None, None) == -1
    assert lib.pyr_trace_host(steps, 1, None, None, None, 1, None, None, None, None, None, 0, 1) == -1


def test_wavelength_batch_argument_validation():
    """PyrRaysIn.n_waves / wave_end (ABI v3): malformed segment tables and media a batch
    cannot carry are rejected before any device work."""
    lib = nat.load()
    rays = nat.PyrRaysIn()
    dummy = (ctypes.c_double * 8)()
    rays.x = ctypes.addressof(dummy)
    rays.k = ctypes.addressof(dummy)
    steps = (nat.PyrStep * 1)(_one_step())
    rays.n_waves = nat.MAX_WAVES + 1
    assert lib.pyr_trace(steps, 1, ctypes.byref(rays), 100, 0, None) == -1        # too many segments
    rays.n_waves = -1
    assert lib.pyr_trace(steps, 1, ctypes.byref(rays), 100, 0, None) == -1
    rays.n_waves = 3
    (rays.wave_end[0], rays.wave_end[1]) = (60, 40)                               # not ascending
    assert lib.pyr_trace(steps, 1, ctypes.byref(rays), 100, 0, None) == -1
    (rays.wave_end[0], rays.wave_end[1]) = (40, 160)                              # beyond the bundle
    assert lib.pyr_trace(steps, 1, ctypes.byref(rays), 100, 0, None) == -1
    (rays.wave_end[0], rays.wave_end[1]) = (40, 60)
    grin = (nat.PyrStep * 1)(_one_step())
    grin[0].after.kind = nat.MEDIUM_ISO_GRIN
    assert lib.pyr_trace(grin, 1, ctypes.byref(rays), 100, 0, None) == -2         # UNSUPPORTED
    assert lib.pyr_trace(steps, 1, ctypes.byref(rays), 100, nat.F_COMPLEX, None) in (-1, -2)
    assert ctypes.sizeof(nat.PyrRaysIn) == 6 * 8 + 8 + 8 * nat.MAX_WAVES + 8


def test_bundle_generator_argument_validation():
    """PyrBundleGen (ABI v4): malformed descriptors and sequences the generating kernels do
    not carry are rejected before any device work."""
    from pyrate_b200 import bundlegen
    lib = nat.load()
    gen = bundlegen.hexapolar_collimated(10, 5.0, 0.0).descriptor("cpu")
    n = gen.total
    assert lib.pyr_generate_bundle(None, n, None, None, None, n, None) == -1
    bad = nat.PyrBundleGen.from_buffer_copy(gen)
    bad.raster = 17
    assert lib.pyr_generate_bundle(ctypes.byref(bad), n, None, None, None, n, None) == -2
    bad = nat.PyrBundleGen.from_buffer_copy(gen)
    bad.first = 5                                         # first + n > total
    assert lib.pyr_generate_bundle(ctypes.byref(bad), n, None, None, None, n, None) == -1
    bad = nat.PyrBundleGen.from_buffer_copy(gen)
    bad.total = n + 1                                     # not 1 + 3 R (R + 1)
    assert lib.pyr_generate_bundle(ctypes.byref(bad), n, None, None, None, n, None) == -1
    bad = nat.PyrBundleGen.from_buffer_copy(gen)
    bad.raster = nat.RASTER_RECT                          # clipped lattice without its row table
    assert lib.pyr_generate_bundle(ctypes.byref(bad), n, None, None, None, n, None) == -1
    assert lib.pyr_generate_bundle(ctypes.byref(gen), 0, None, None, None, 0, None) == 0
    # a trace call with a generator needs no ray arrays ...
    rays = nat.PyrRaysIn()
    rays.gen = ctypes.pointer(gen)
    steps = (nat.PyrStep * 1)(_one_step())
    assert lib.pyr_trace(steps, 1, ctypes.byref(rays), 0, 0, None) == 0
    # ... but only real-valued, non-batched sequences without E recording
    assert lib.pyr_trace(steps, 1, ctypes.byref(rays), n, nat.F_RECORD_E, None) == -2
    rays.n_waves = 2
    rays.wave_end[0] = 10
    assert lib.pyr_trace(steps, 1, ctypes.byref(rays), n, 0, None) == -2
    # host entry: generator or arrays, crystals refused
    io = nat.PyrHostIO()
    assert lib.pyr_trace_host_io(steps, 1, ctypes.byref(io), 10, None, 0, 4) == -1
    assert lib.pyr_trace_host_io_workspace(13, 1 << 20, 1) > lib.pyr_trace_host_io_workspace(13, 1 << 20, 0) > 0
