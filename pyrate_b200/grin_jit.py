"""User-defined GRIN index functions on the device (NVRTC).

The reference lets the user write the index profile as Python source --
`IsotropicGrinMaterial` keeps `nfunc`, `dndx`, `dndy`, `dndz`, `bnd` taken from a source
string (core/functionobject.py:99-119, material/material_grin.py:37-104) and evaluates them
inside its integrator loop.  The fused trace kernels only know the closed catalogue of
include/pyrate_b200.h (PyrGrinProfile).  For anything else the material carries

    annotations["device_source"] = {
        "n":    "p[0] + p[1] * exp(-x * x - 4.0 * y * y)",     # CUDA C++ expressions in
        "dndx": "...", "dndy": "...", "dndz": "0.0",           # x, y, z (material frame)
        "inside": "x * x + y * y < 100.0",                     # and the parameter array p[]
        "params": [1.0, 0.5]}

and this module compiles, ONCE per distinct source, a small kernel pair with NVRTC for
sm_100a: `pyr_user_grin_eval` (index at given global points: the refraction INTO the medium
needs it per ray, PyrStep.after_n_rays) and `pyr_user_grin_propagate` (the symplectic
integrator of material_grin.py:106-213 with the engine's per-ray normalisations, up to a
conic next surface).  Before the first trace the compiled functions are sampled on the
device and compared with the material's Python functions: a device source that does not
describe the same medium raises instead of tracing something else.

cuda-python (cuda.bindings.nvrtc / driver) does the compiling and launching; the kernels
run on torch's current stream and context.
"""
import ctypes as C
import hashlib

import numpy as np

_SOURCE = r'''
typedef unsigned char uint8_t;
struct Frame { double r[9]; double o[3]; };
struct Args {
    const double *x; const double *k; const uint8_t *alive;
    long long ld; long long n;
    double *out_x; double *out_k; uint8_t *out_alive; double *out_n; double *out_g;
    Frame mat; Frame to_shape;
    double curv, cc, ds, energy_tol;
    int max_steps; int pad;
    double p[16];
};

__device__ __forceinline__ void user_index(const double *p, double x, double y, double z,
                                           double &n, double &gx, double &gy, double &gz) {
    n = (@N@);
    gx = (@DNDX@);
    gy = (@DNDY@);
    gz = (@DNDZ@);
}
__device__ __forceinline__ bool user_inside(const double *p, double x, double y, double z) {
    return (@INSIDE@);
}

__device__ __forceinline__ void g2l(const Frame &f, const double *xg, double *xl) {
    const double t0 = xg[0] - f.o[0], t1 = xg[1] - f.o[1], t2 = xg[2] - f.o[2];
    xl[0] = f.r[0] * t0 + f.r[3] * t1 + f.r[6] * t2;
    xl[1] = f.r[1] * t0 + f.r[4] * t1 + f.r[7] * t2;
    xl[2] = f.r[2] * t0 + f.r[5] * t1 + f.r[8] * t2;
}
__device__ __forceinline__ void l2g(const Frame &f, const double *xl, double *xg) {
    xg[0] = f.r[0] * xl[0] + f.r[1] * xl[1] + f.r[2] * xl[2] + f.o[0];
    xg[1] = f.r[3] * xl[0] + f.r[4] * xl[1] + f.r[5] * xl[2] + f.o[1];
    xg[2] = f.r[6] * xl[0] + f.r[7] * xl[1] + f.r[8] * xl[2] + f.o[2];
}
__device__ __forceinline__ void rot_t(const Frame &f, const double *v, double *y) {
    y[0] = f.r[0] * v[0] + f.r[3] * v[1] + f.r[6] * v[2];
    y[1] = f.r[1] * v[0] + f.r[4] * v[1] + f.r[7] * v[2];
    y[2] = f.r[2] * v[0] + f.r[5] * v[1] + f.r[8] * v[2];
}
__device__ __forceinline__ void rot(const Frame &f, const double *v, double *y) {
    y[0] = f.r[0] * v[0] + f.r[1] * v[1] + f.r[2] * v[2];
    y[1] = f.r[3] * v[0] + f.r[4] * v[1] + f.r[5] * v[2];
    y[2] = f.r[6] * v[0] + f.r[7] * v[1] + f.r[8] * v[2];
}

// index (and gradient) at GLOBAL points: out_n[i], out_g[c * ld + i] (either may be null)
extern "C" __global__ void pyr_user_grin_eval(const Args a) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n;
         i += (long long)gridDim.x * blockDim.x) {
        const double xg[3] = {a.x[i], a.x[a.ld + i], a.x[2 * a.ld + i]};
        double q[3], n, gx, gy, gz;
        g2l(a.mat, xg, q);
        user_index(a.p, q[0], q[1], q[2], n, gx, gy, gz);
        if (a.out_n) a.out_n[i] = n;
        if (a.out_g) { a.out_g[i] = gx; a.out_g[a.ld + i] = gy; a.out_g[2 * a.ld + i] = gz; }
        if (a.out_alive) a.out_alive[i] = user_inside(a.p, q[0], q[1], q[2]) ? 1 : 0;
    }
}

// material_grin.py:106-213 per ray (energy test per ray, a ray stops stepping once final):
// from x along d = k / |k| until the conic next surface is crossed; out_x = last position in
// front of the surface, out_k = p / n there (global frame), out_alive = PYR_RAY_ALIVE (2) if valid
extern "C" __global__ void pyr_user_grin_propagate(const Args a) {
    const double cbrt2 = 1.2599210498948732;
    const double c0 = 1.0 / (2.0 * (2.0 - cbrt2)), c1 = (1.0 - cbrt2) / (2.0 * (2.0 - cbrt2));
    const double d0 = 1.0 / (2.0 - cbrt2), d1 = -cbrt2 / (2.0 - cbrt2);
    const double cs[4] = {c0, c1, c1, c0};
    const double ds[4] = {d0, d1, d0, 0.0};
    const double tau2 = 2.0 * a.ds;
    const int cap = a.max_steps > 0 ? a.max_steps : 1000000;
    const double qn = __longlong_as_double(0x7ff8000000000000LL);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n;
         i += (long long)gridDim.x * blockDim.x) {
        const bool enter = a.alive ? (a.alive[i] & 2) != 0 : true;
        if (!enter) {
            for (int c = 0; c < 3; ++c) { a.out_x[c * a.ld + i] = qn; a.out_k[c * a.ld + i] = qn; }
            a.out_alive[i] = 0;
            continue;
        }
        const double xg[3] = {a.x[i], a.x[a.ld + i], a.x[2 * a.ld + i]};
        double kg[3] = {a.k[i], a.k[a.ld + i], a.k[2 * a.ld + i]};
        const double inv = 1.0 / sqrt(kg[0] * kg[0] + kg[1] * kg[1] + kg[2] * kg[2]);
        kg[0] *= inv; kg[1] *= inv; kg[2] *= inv;
        double q[3], p[3], n, g[3];
        g2l(a.mat, xg, q);
        rot_t(a.mat, kg, p);
        user_index(a.p, q[0], q[1], q[2], n, g[0], g[1], g[2]);
        p[0] *= n; p[1] *= n; p[2] *= n;
        double uq[3] = {q[0], q[1], q[2]}, up[3] = {p[0], p[1], p[2]};
        bool valid = true;
        for (int it = 0; it < cap; ++it) {
            double nq = 0.0;
            for (int s = 0; s < 4; ++s) {
                q[0] = fma(tau2 * cs[s], p[0], q[0]);
                q[1] = fma(tau2 * cs[s], p[1], q[1]);
                q[2] = fma(tau2 * cs[s], p[2], q[2]);
                user_index(a.p, q[0], q[1], q[2], nq, g[0], g[1], g[2]);
                const double f = tau2 * ds[s] * nq;
                p[0] = fma(f, g[0], p[0]); p[1] = fma(f, g[1], p[1]); p[2] = fma(f, g[2], p[2]);
            }
            if (!(fabs(p[0] * p[0] + p[1] * p[1] + p[2] * p[2] - nq * nq) <= a.energy_tol)) valid = false;
            double xs[3];
            l2g(a.to_shape, q, xs);
            const double r2 = xs[0] * xs[0] + xs[1] * xs[1];
            const double sarg = 1.0 - (1.0 + a.cc) * a.curv * a.curv * r2;
            const double sag = sarg > 0.0 ? a.curv * r2 / (1.0 + sqrt(sarg)) : qn;
            const double gap = xs[2] - sag;
            const bool crossed = gap > 0.0;
            if (!user_inside(a.p, q[0], q[1], q[2]) || gap != gap) valid = false;
            const bool stop = crossed || !valid;
            if (!stop) {
                for (int c = 0; c < 3; ++c) { uq[c] = q[c]; up[c] = p[c]; }
                if (it == cap - 1) valid = false;
            }
            if (stop) break;
        }
        user_index(a.p, uq[0], uq[1], uq[2], n, g[0], g[1], g[2]);
        const double kl[3] = {up[0] / n, up[1] / n, up[2] / n};
        double xo[3], ko[3];
        l2g(a.mat, uq, xo);
        rot(a.mat, kl, ko);
        for (int c = 0; c < 3; ++c) { a.out_x[c * a.ld + i] = xo[c]; a.out_k[c * a.ld + i] = ko[c]; }
        a.out_alive[i] = valid ? 2 : 0;
    }
}
'''


class _Frame(C.Structure):
    _fields_ = [("r", C.c_double * 9), ("o", C.c_double * 3)]


class _Args(C.Structure):
    _fields_ = [("x", C.c_void_p), ("k", C.c_void_p), ("alive", C.c_void_p),
                ("ld", C.c_longlong), ("n", C.c_longlong),
                ("out_x", C.c_void_p), ("out_k", C.c_void_p), ("out_alive", C.c_void_p),
                ("out_n", C.c_void_p), ("out_g", C.c_void_p),
                ("mat", _Frame), ("to_shape", _Frame),
                ("curv", C.c_double), ("cc", C.c_double), ("ds", C.c_double),
                ("energy_tol", C.c_double), ("max_steps", C.c_int), ("pad", C.c_int),
                ("p", C.c_double * 16)]


class GrinJitError(RuntimeError):
    pass


def render_source(device_source):
    """The CUDA translation unit of one user profile."""
    src = _SOURCE
    for (tag, key, default) in (("@N@", "n", None), ("@DNDX@", "dndx", None), ("@DNDY@", "dndy", None),
                                ("@DNDZ@", "dndz", "0.0"), ("@INSIDE@", "inside", "true")):
        expr = device_source.get(key, default)
        if expr is None:
            raise GrinJitError("device_source lacks the expression %r" % key)
        src = src.replace(tag, str(expr))
    return src


_CUBIN_CACHE = {}       # sha1 of the source -> cubin bytes (compiling needs no GPU)
_MODULE_CACHE = {}      # (sha1, device index) -> (module, eval function, propagate function)


def compile_cubin(device_source):
    """NVRTC: source -> sm_100a cubin (cached per process).  Works without a GPU."""
    from cuda.bindings import nvrtc
    src = render_source(device_source)
    key = hashlib.sha1(src.encode()).hexdigest()
    hit = _CUBIN_CACHE.get(key)
    if hit is not None:
        return key, hit
    (err, prog) = nvrtc.nvrtcCreateProgram(src.encode(), b"pyr_user_grin.cu", 0, [], [])
    if int(err) != 0:
        raise GrinJitError("nvrtcCreateProgram failed: %s" % (err,))
    opts = [b"--gpu-architecture=sm_100a", b"--std=c++17", b"--fmad=true"]
    (err,) = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
    if int(err) != 0:
        (_, size) = nvrtc.nvrtcGetProgramLogSize(prog)
        log = b" " * size
        nvrtc.nvrtcGetProgramLog(prog, log)
        raise GrinJitError("the device_source of the GRIN material does not compile:\n%s" %
                           log.decode(errors="replace"))
    (_, size) = nvrtc.nvrtcGetCUBINSize(prog)
    cubin = b" " * size
    nvrtc.nvrtcGetCUBIN(prog, cubin)
    nvrtc.nvrtcDestroyProgram(prog)
    _CUBIN_CACHE[key] = cubin
    return key, cubin


def _check(res):
    err = res[0]
    if int(err) != 0:
        raise GrinJitError("CUDA driver call failed: %s" % (err,))
    return res[1:] if len(res) > 2 else (res[1] if len(res) == 2 else None)


class UserGrin(object):
    """The compiled kernels of one user profile on one device."""

    def __init__(self, device_source, device):
        import torch
        from cuda.bindings import driver
        self.driver = driver
        self.device = torch.device(device)
        torch.cuda.current_stream(self.device)          # torch has made the primary context current
        (key, cubin) = compile_cubin(device_source)
        hit = _MODULE_CACHE.get((key, self.device.index))
        if hit is None:
            with torch.cuda.device(self.device):
                module = _check(driver.cuModuleLoadData(cubin))
                f_eval = _check(driver.cuModuleGetFunction(module, b"pyr_user_grin_eval"))
                f_prop = _check(driver.cuModuleGetFunction(module, b"pyr_user_grin_propagate"))
            hit = _MODULE_CACHE[(key, self.device.index)] = (module, f_eval, f_prop)
        (self.module, self.f_eval, self.f_prop) = hit
        self.params = [float(v) for v in device_source.get("params", [])]
        if len(self.params) > 16:
            raise GrinJitError("at most 16 parameters")

    def _args(self, frame_mat):
        a = _Args()
        for (i, v) in enumerate(self.params):
            a.p[i] = v
        a.mat.r[:] = list(frame_mat.r)
        a.mat.o[:] = list(frame_mat.o)
        return a

    def _launch(self, fn, args, n):
        import torch
        stream = torch.cuda.current_stream(self.device).cuda_stream
        ptrs = (C.c_void_p * 1)(C.addressof(args))
        grid = max(1, min((n + 255) // 256, 148 * 8))
        with torch.cuda.device(self.device):
            _check(self.driver.cuLaunchKernel(fn, grid, 1, 1, 256, 1, 1, 0, stream, C.addressof(ptrs), 0))

    def index_at(self, frame_mat, x, n, ld, want_inside=False):
        """n(x) at the global points x (3, ld): float64 CUDA tensor (ld,) [and inside flags]."""
        import torch
        out = torch.empty((ld,), dtype=torch.float64, device=self.device)
        ins = torch.empty((ld,), dtype=torch.uint8, device=self.device) if want_inside else None
        a = self._args(frame_mat)
        (a.x, a.ld, a.n, a.out_n) = (x.data_ptr(), ld, n, out.data_ptr())
        if ins is not None:
            a.out_alive = ins.data_ptr()
        self._launch(self.f_eval, a, n)
        return (out, ins) if want_inside else out

    def gradient_at(self, frame_mat, x, n, ld):
        import torch
        val = torch.empty((ld,), dtype=torch.float64, device=self.device)
        grad = torch.empty((3, ld), dtype=torch.float64, device=self.device)
        a = self._args(frame_mat)
        (a.x, a.ld, a.n, a.out_n, a.out_g) = (x.data_ptr(), ld, n, val.data_ptr(), grad.data_ptr())
        self._launch(self.f_eval, a, n)
        return val, grad

    def propagate(self, medium, to_shape, curv, cc, x, k, alive, n, ld):
        """Integrate n rays through the medium up to the conic (curv, cc) whose frame is
        reached from the material frame by `to_shape` (r[9], o[3]).  Returns (x, k, alive)."""
        import torch
        xo = torch.empty((3, ld), dtype=torch.float64, device=self.device)
        ko = torch.empty((3, ld), dtype=torch.float64, device=self.device)
        ao = torch.zeros((ld,), dtype=torch.uint8, device=self.device)
        a = self._args(medium.frame)
        (a.x, a.k, a.ld, a.n) = (x.data_ptr(), k.data_ptr(), ld, n)
        a.alive = alive.data_ptr() if alive is not None else None
        (a.out_x, a.out_k, a.out_alive) = (xo.data_ptr(), ko.data_ptr(), ao.data_ptr())
        a.to_shape.r[:] = [float(v) for v in to_shape[0]]
        a.to_shape.o[:] = [float(v) for v in to_shape[1]]
        (a.curv, a.cc, a.ds, a.energy_tol) = (float(curv), float(cc), float(medium.grin_ds),
                                              float(medium.grin_energy_tol))
        a.max_steps = int(medium.grin_max_steps)
        self._launch(self.f_prop, a, n)
        return xo, ko, ao

    def verify(self, material, frame_mat, samples=256):
        """Compare the compiled index / gradient / boundary with the material's Python
        functions on random points (material frame); raises on disagreement."""
        import torch
        user = material.user_functions() if hasattr(material, "user_functions") else None
        if user is None:
            return
        rng = np.random.default_rng(2468)
        pts = rng.uniform(-3.0, 3.0, (3, samples))
        r = np.asarray(list(frame_mat.r)).reshape(3, 3)
        o = np.asarray(list(frame_mat.o))
        glob = np.ascontiguousarray(r @ pts + o[:, None])
        xg = torch.from_numpy(glob).to(self.device)
        (val, grad) = self.gradient_at(frame_mat, xg, samples, samples)
        (val, grad) = (val.cpu().numpy(), grad.cpu().numpy())
        want = [np.asarray(f(pts), dtype=float) for f in user[:4]]
        for (name, got, ref) in (("n", val, want[0]), ("dndx", grad[0], want[1]),
                                 ("dndy", grad[1], want[2]), ("dndz", grad[2], want[3])):
            if not np.allclose(got, ref, rtol=1e-11, atol=1e-12):
                raise GrinJitError(
                    "GRIN material %r: device_source[%r] disagrees with the Python source (max |diff| "
                    "= %.3e); refusing to trace a different medium"
                    % (getattr(material, "name", "?"), name, float(np.max(np.abs(got - ref)))))
        far = rng.uniform(-30.0, 30.0, (3, samples))
        gfar = torch.from_numpy(np.ascontiguousarray(r @ far + o[:, None])).to(self.device)
        (_, ins) = self.index_at(frame_mat, gfar, samples, samples, want_inside=True)
        if not np.array_equal(ins.cpu().numpy() != 0, np.asarray(user[4](far), dtype=bool)):
            raise GrinJitError("GRIN material %r: device_source['inside'] disagrees with the Python "
                               "boundary function" % (getattr(material, "name", "?"),))
