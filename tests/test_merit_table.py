"""CPU: the live step table of the optimiser-loop binding (pyrate_b200.merit): after
arbitrary changes of optimisable variables (curvatures, conic constants, asphere
coefficients, decenters / tilts, dispersion coefficients) `refresh()` must leave the table
byte-identical to a fresh lowering of the mutated system."""
import ctypes as C

import numpy as np
import pytest

import pyrate_b200 as pb
from pyrate_b200 import _native as nat
from pyrate_b200 import configs, lowering
from pyrate_b200.merit import LiveStepTable

POINTER_FIELDS = ("out_x", "out_k", "out_e", "out_flags", "grid_tx", "grid_ty", "grid_c",
                  "grin_hist_x", "grin_hist_k", "grin_hist_valid", "grin_hist_count")


def _bytes_without_pointers(st):
    copy = nat.PyrStep.from_buffer_copy(st)
    for f in POINTER_FIELDS:
        setattr(copy, f, None)
    copy.ld_out = 0
    return bytes(copy)


@pytest.mark.parametrize("name", ["c2_doublegauss", "c3_asphere", "x1_tilted", "x2_xypoly",
                                  "x14_dispersive", "x16_cylinder", "c5_grin"])
def test_refresh_equals_fresh_lowering(name):
    spec = configs.CONFIGS[name]
    (s, seq) = configs.build_system(spec, pb.api())
    live = LiveStepTable(s, seq, configs.DLINE)
    rng = np.random.default_rng(11)
    elem = s.elements["stdelem"]
    for round_ in range(4):
        for surf in elem.surfaces.values():
            shape = surf.shape
            for var in (getattr(shape, "curvature", None), getattr(shape, "conic", None)):
                if var is not None:
                    var.setvalue(var() * (1 + 0.01 * rng.standard_normal()) + 1e-4 * rng.standard_normal())
            for var in getattr(shape, "params", {}).values():
                var.setvalue(var() * (1 + 0.01 * rng.standard_normal()))
            lc = shape.lc
            for nm in ("decx", "decy", "decz", "tiltx", "tilty"):
                getattr(lc, nm).setvalue(getattr(lc, nm)() + 1e-3 * rng.standard_normal())
        for mat in elem.materials.values():
            for var in getattr(mat, "params", {}).values() if hasattr(mat, "params") else ():
                var.setvalue(var() * (1 + 1e-3 * rng.standard_normal()))
            for nm in ("n0", "A", "B"):
                v = getattr(mat, nm, None)
                if v is not None and hasattr(v, "setvalue"):
                    v.setvalue(v() * (1 + 1e-3 * rng.standard_normal()))
        s.rootcoordinatesystem.update()
        if name == "c5_grin":
            # (the GRIN profile parameters are verified against the Python source at lowering
            # time; this config only varies geometry)
            pass
        live.refresh()
        fresh = lowering.lower(s, seq, configs.DLINE)
        assert len(fresh) == live.n_steps
        for (i, ls) in enumerate(fresh):
            assert _bytes_without_pointers(live.arr[i]) == _bytes_without_pointers(ls.st), (round_, i)
    # something did change
    first = lowering.step_array(lowering.lower(*configs.build_system(spec, pb.api()), configs.DLINE))
    assert any(_bytes_without_pointers(first[i]) != _bytes_without_pointers(live.arr[i])
               for i in range(live.n_steps))
