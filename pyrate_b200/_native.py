"""ctypes binding of the C ABI in include/pyrate_b200.h (libpyrate_b200.so).

The same mechanism the reference uses for its only FFI
(raytracer/surface_shape_zmxdll.py:293, ctypes.CDLL).  There is NO fallback: if
the library is missing or was built for another ABI, loading raises.
"""
import ctypes as C
import os

ABI_VERSION = 4
MAX_COEFF = 80
MAX_GRIN_PARAMS = 8

(SHAPE_CONIC, SHAPE_ASPHERE, SHAPE_XYPOLY, SHAPE_BICONIC, SHAPE_GRIDSAG,
 SHAPE_COMBINATION, SHAPE_CYLINDER) = (0, 1, 2, 3, 4, 5, 6)
MAX_TERMS = 4
MAX_WAVES = 4
(AP_BASE, AP_CIRCULAR, AP_RECTANGULAR) = (0, 1, 2)
(REFRACT, REFLECT) = (0, 1)
(MEDIUM_ISO_CONST, MEDIUM_ISO_GRIN, MEDIUM_ANISO) = (0, 1, 2)
(GRIN_GAUSSIAN_XY, GRIN_POLY_RZ, GRIN_USER) = (0, 1, 100)
(BND_NONE, BND_CYLINDER, BND_BOX, BND_SPHERE) = (0, 1, 2, 3)
(DIR_POYNTING, DIR_K) = (0, 1)
(STEP_FULL, STEP_PROPAGATE_ONLY, STEP_DEFLECT_ONLY) = (0, 1, 2)
(RAY_HIT, RAY_ALIVE) = (1, 2)
(F_COMPLEX, F_RECORD_E) = (1, 2)
(E_BADARG, E_UNSUPPORTED, E_TOOLARGE, E_NODEVICE) = (-1, -2, -3, -4)


class PyrFrame(C.Structure):
    _fields_ = [("r", C.c_double * 9), ("o", C.c_double * 3)]


class PyrMedium(C.Structure):
    _fields_ = [("kind", C.c_int32), ("grin_profile", C.c_int32),
                ("grin_boundary", C.c_int32), ("grin_max_steps", C.c_int32),
                ("n", C.c_double), ("eps", C.c_double * 18),
                ("grin_p", C.c_double * MAX_GRIN_PARAMS),
                ("grin_b", C.c_double * 4),
                ("grin_ds", C.c_double), ("grin_energy_tol", C.c_double),
                ("frame", PyrFrame)]


class PyrShapeTerm(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_coeff", C.c_int32),
                ("coeff_off", C.c_int32), ("coeff_len", C.c_int32),
                ("weight", C.c_double),
                ("dx", C.c_double), ("dy", C.c_double), ("dz", C.c_double),
                ("curv", C.c_double), ("cc", C.c_double),
                ("curv2", C.c_double), ("cc2", C.c_double),
                ("normradius", C.c_double)]


class PyrStep(C.Structure):
    _fields_ = [("shape_kind", C.c_int32), ("aperture_kind", C.c_int32),
                ("interaction", C.c_int32), ("dir_mode", C.c_int32),
                ("n_coeff", C.c_int32), ("newton_maxit", C.c_int32),
                ("split", C.c_int32), ("mode", C.c_int32),
                ("k_norm_hint", C.c_double),
                ("curv", C.c_double), ("cc", C.c_double),
                ("curv2", C.c_double), ("cc2", C.c_double),
                ("normradius", C.c_double), ("newton_tol", C.c_double),
                ("coeff", C.c_double * MAX_COEFF),
                ("xpow", C.c_int8 * MAX_COEFF), ("ypow", C.c_int8 * MAX_COEFF),
                ("aperture_p", C.c_double * 4),
                ("shape_frame", PyrFrame), ("aperture_frame", PyrFrame),
                ("before", PyrMedium), ("after", PyrMedium),
                ("out_x", C.c_void_p), ("out_k", C.c_void_p),
                ("out_e", C.c_void_p), ("out_flags", C.c_void_p),
                ("ld_out", C.c_int64),
                ("grin_hist_x", C.c_void_p), ("grin_hist_k", C.c_void_p),
                ("grin_hist_valid", C.c_void_p), ("grin_hist_count", C.c_void_p),
                ("grin_hist_rows", C.c_int64),
                ("grid_tx", C.c_void_p), ("grid_ty", C.c_void_p),
                ("grid_c", C.c_void_p),
                ("grid_nx", C.c_int32), ("grid_ny", C.c_int32),
                ("n_terms", C.c_int32), ("reserved0", C.c_int32),
                ("terms", PyrShapeTerm * MAX_TERMS),
                ("ld_out2", C.c_int64),
                ("before_n_w", C.c_double * MAX_WAVES),
                ("after_n_w", C.c_double * MAX_WAVES),
                ("after_n_rays", C.c_void_p)]


(RASTER_HEXAPOLAR, RASTER_RECT, RASTER_HEX, RASTER_CIRCULAR) = (0, 1, 2, 3)
(BUNDLE_COLLIMATED, BUNDLE_DIVERGENT) = (0, 1)
(GEN_E_PERP, GEN_SQRT_R) = (1, 2)


class PyrBundleGen(C.Structure):
    _fields_ = [("raster", C.c_int32), ("bundle", C.c_int32), ("flags", C.c_uint32),
                ("reserved0", C.c_int32),
                ("param", C.c_int64), ("first", C.c_int64), ("total", C.c_int64),
                ("lin_start", C.c_double), ("lin_step", C.c_double), ("lin_stop", C.c_double),
                ("aux", C.c_double * 2), ("radius", C.c_double),
                ("start", C.c_double * 3), ("dir", C.c_double * 3), ("e", C.c_double * 3),
                ("n_index", C.c_double), ("rows", C.c_void_p)]


class PyrRaysIn(C.Structure):
    _fields_ = [("x", C.c_void_p), ("k", C.c_void_p), ("e", C.c_void_p),
                ("alive", C.c_void_p), ("ld", C.c_int64), ("n_x", C.c_int64),
                ("n_waves", C.c_int32), ("reserved0", C.c_int32),
                ("wave_end", C.c_int64 * MAX_WAVES),
                ("gen", C.POINTER(PyrBundleGen))]


class PyrHostIO(C.Structure):
    _fields_ = [("x0", C.c_void_p), ("k0", C.c_void_p), ("e0", C.c_void_p),
                ("gen", C.POINTER(PyrBundleGen)),
                ("x_last", C.c_void_p), ("k_last", C.c_void_p), ("flags_last", C.c_void_p),
                ("x_all", C.c_void_p), ("k_all", C.c_void_p), ("flags_all", C.c_void_p),
                ("spot8", C.c_void_p), ("e_last", C.c_void_p)]


LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lib",
                        "libpyrate_b200.so")

EXPORTS = ("pyr_version", "pyr_strerror", "pyr_sizeof_step",
           "pyr_sizeof_rays_in", "pyr_sizeof_bundle_gen", "pyr_device_count", "pyr_trace",
           "pyr_spot_sums", "pyr_spot_points", "pyr_trace_spot", "pyr_generate_bundle",
           "pyr_grin_lockstep", "pyr_grin_lockstep_scratch", "pyr_trace_host_workspace",
           "pyr_trace_host", "pyr_trace_host_io_workspace", "pyr_trace_host_io",
           "pyr_trace_host_crystal_workspace")

_lib = None


def use_tools_library(name="libpyrate_b200_tools.so"):
    """tools/ only: load a measurement build (`make tools`, -DPYR_TOOLS: A/B knobs
    compiled in) instead of the product library.  Must be called before load()."""
    global LIB_PATH
    assert _lib is None, "library already loaded"
    LIB_PATH = os.path.join(os.path.dirname(LIB_PATH), name)


class NativeError(RuntimeError):
    pass


def load():
    """Load (once) and type the library.  Raises if absent or ABI-mismatched."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(
            "pyrate_b200: %s not built (run `make` or __graft_entry__.build()); "
            "there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.pyr_version.restype = C.c_int
    lib.pyr_strerror.restype = C.c_char_p
    lib.pyr_strerror.argtypes = [C.c_int]
    lib.pyr_sizeof_step.restype = C.c_int64
    lib.pyr_sizeof_rays_in.restype = C.c_int64
    lib.pyr_device_count.restype = C.c_int
    lib.pyr_trace.restype = C.c_int
    lib.pyr_trace.argtypes = [C.POINTER(PyrStep), C.c_int32,
                              C.POINTER(PyrRaysIn), C.c_int64, C.c_uint32,
                              C.c_void_p]
    lib.pyr_spot_sums.restype = C.c_int
    lib.pyr_spot_sums.argtypes = [C.c_void_p, C.c_int64, C.c_void_p,
                                  C.c_uint32, C.c_int64, C.POINTER(C.c_double),
                                  C.c_void_p, C.c_void_p]
    lib.pyr_trace_host_workspace.restype = C.c_int64
    lib.pyr_trace_host_workspace.argtypes = [C.c_int32, C.c_int64]
    lib.pyr_trace_host.restype = C.c_int
    lib.pyr_trace_host.argtypes = [C.POINTER(PyrStep), C.c_int32, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_int64,
                                   C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
    lib.pyr_trace_spot.restype = C.c_int
    lib.pyr_trace_spot.argtypes = [C.POINTER(PyrStep), C.c_int32, C.POINTER(PyrRaysIn), C.c_int64,
                                   C.c_uint32, C.POINTER(C.c_double), C.c_void_p, C.c_void_p,
                                   C.c_void_p]
    lib.pyr_grin_lockstep_scratch.restype = C.c_int64
    lib.pyr_grin_lockstep_scratch.argtypes = [C.c_int64]
    lib.pyr_grin_lockstep.restype = C.c_int
    lib.pyr_grin_lockstep.argtypes = [C.POINTER(PyrStep)] + [C.c_void_p] * 4 + [C.c_int64, C.c_int64] + \
        [C.c_void_p] * 8 + [C.c_int64, C.c_void_p]
    lib.pyr_spot_points.restype = C.c_int
    lib.pyr_spot_points.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_uint32, C.c_int64,
                                    C.POINTER(PyrFrame), C.c_void_p, C.c_int64, C.c_void_p,
                                    C.c_void_p]
    lib.pyr_sizeof_bundle_gen.restype = C.c_int64
    lib.pyr_generate_bundle.restype = C.c_int
    lib.pyr_generate_bundle.argtypes = [C.POINTER(PyrBundleGen), C.c_int64, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.pyr_trace_host_io_workspace.restype = C.c_int64
    lib.pyr_trace_host_io_workspace.argtypes = [C.c_int32, C.c_int64, C.c_int32]
    lib.pyr_trace_host_crystal_workspace.restype = C.c_int64
    lib.pyr_trace_host_crystal_workspace.argtypes = [C.POINTER(PyrStep), C.c_int32, C.c_int64,
                                                     C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.pyr_trace_host_io.restype = C.c_int
    lib.pyr_trace_host_io.argtypes = [C.POINTER(PyrStep), C.c_int32, C.POINTER(PyrHostIO),
                                      C.c_int64, C.c_void_p, C.c_int64, C.c_int64]
    if lib.pyr_version() != ABI_VERSION:
        raise NativeError("pyrate_b200: libpyrate_b200.so has ABI version %d, "
                          "_native.py expects %d (rebuild with `make`)" %
                          (lib.pyr_version(), ABI_VERSION))
    if lib.pyr_sizeof_step() != C.sizeof(PyrStep) or \
            lib.pyr_sizeof_rays_in() != C.sizeof(PyrRaysIn) or \
            lib.pyr_sizeof_bundle_gen() != C.sizeof(PyrBundleGen):
        raise NativeError("pyrate_b200: struct layout mismatch between "
                          "_native.py and libpyrate_b200.so (%d vs %d)" %
                          (C.sizeof(PyrStep), lib.pyr_sizeof_step()))
    _lib = lib
    return lib


def check(code):
    if code != 0:
        err = NativeError("pyrate_b200 native call failed (%d): %s" %
                          (code, load().pyr_strerror(code).decode()))
        err.code = code
        raise err
