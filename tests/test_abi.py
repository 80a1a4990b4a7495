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
    assert lib.pyr_version() == 1


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
