// Device-side step table and small FP64 helpers of the trace kernels.
//
// The public PyrStep (include/pyrate_b200.h) is a fat, self-describing POD.  For
// the launch it is packed into a compact DStep (hot fields every step needs)
// plus an optional DAux (polynomial coefficients, second frame, GRIN / crystal
// payload).  Both tables travel in the kernel parameter block
// (__grid_constant__, constant bank 0), so per-step constants reach the FP64
// pipe as constant-cache operands instead of LSU traffic, and the library keeps
// no global state (re-entrant per stream).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pyrate_b200.h"

namespace pyr {

constexpr int kMaxSteps = 40;      // per launch (longer sequences are chunked)
constexpr int kMaxAux = 10;

// DStep.bits
constexpr int kMaxWaves = PYR_MAX_WAVES;
constexpr uint32_t kRotIdentity = 1u;      // shape frame rotation is the identity
constexpr uint32_t kApSameFrame = 2u;      // aperture.lc == shape.lc
constexpr uint32_t kSplit = 4u;            // anisotropic ray doubling on this step
constexpr uint32_t kSphere = 8u;           // cc == 0: |grad| == 1, no normalisation
constexpr uint32_t kPlane = 16u;           // curv == 0 (and conic shape)
constexpr uint32_t kOutVec2 = 32u;         // outputs allow 128-bit stores
constexpr uint32_t kNoDeflect = 64u;       // PYR_STEP_PROPAGATE_ONLY
constexpr uint32_t kNoIntersect = 128u;    // PYR_STEP_DEFLECT_ONLY
constexpr uint32_t kRecorded = 256u;       // the step has at least one output pointer
constexpr uint32_t kPlain = 512u;          // untilted frame, no aperture, refraction, cc == 0: step_lean<.., PLAIN>

struct DFrame {
    double r[9];
    double o[3];
};

struct DMedium {
    int32_t kind, profile, boundary, max_steps;
    double n;
    double eps[18];        // rotated into the SHAPE frame of the step at pack time
    double p[PYR_MAX_GRIN_PARAMS];
    double b[4];
    double ds, energy_tol;
    DFrame frame;          // material frame
    DFrame to_shape;       // material frame -> shape frame of this step
};

struct DAux {
    double coeff[PYR_MAX_COEFF];
    int8_t xpow[PYR_MAX_COEFF];
    int8_t ypow[PYR_MAX_COEFF];
    double normradius, newton_tol;
    double curv2, cc2;             // biconic y section
    int32_t n_coeff, newton_maxit;
    DFrame aperture_frame;
    DMedium before, after;
    double *hist_x, *hist_k;       // optional GRIN integrator history
    uint8_t *hist_valid;
    int32_t *hist_count;
    int64_t hist_rows;
    // grid sag (FITPACK bicubic B-spline), device pointers
    const double *grid_tx, *grid_ty, *grid_c;
    int32_t grid_nx, grid_ny;
    // linear combination: this record is term 0 and the following n_terms - 1 records
    // are the other terms (each with its own coefficients / grid)
    int32_t n_terms, term_kind;
    double term_w, term_dx, term_dy, term_dz, term_curv, term_cc;
    // crystal payload of the DEFLECTING medium (`after`), everything in the shape frame:
    // uniaxial / isotropic real tensors eps = eps_o 1 + (eps_e - eps_o) a a^T take the
    // closed-form roots (the general case forms the invariants of the Fresnel quartic on
    // the device: the kernel parameter block has no room for them)
    int32_t uniaxial, pad1;
    double eps_o, eps_e, axis[3];
    const double *after_n_rays;    // per-ray index of the deflecting medium at the hit point (user GRIN)
};

struct DStep {
    DFrame frame;                  // shape frame (local -> global)
    double curv, cc;
    double ap0, ap1;               // circular: min^2, max^2; rectangular: w/2, h/2
    double n2sq[kMaxWaves];        // (index of the deflecting medium)^2, ISO_CONST, per
                                   // wavelength segment (entry 0 for a plain call)
    double inv_knorm[kMaxWaves];   // > 0: |k| known on entry (1/n of `before`)
    double *out_x, *out_k, *out_e;
    uint8_t *out_flags;
    int64_t ld_out, ld_out2;
    uint32_t bits;
    int8_t shape_kind, aperture_kind, interaction, dir_mode;
    int8_t before_kind, after_kind, aux, pad0;
};

// bundle generator (PyrBundleGen packed for the launch, csrc/pyr_gen.cuh)
struct DGen {
    int32_t raster, bundle;
    uint32_t flags;
    int32_t on;                    // 0: no generator (rays are read from memory)
    int64_t param, first;
    double lin_start, lin_step, lin_stop;
    double aux0, aux1;
    double radius;
    double start[3], dir[3], e[3];
    double n_index;
    const int64_t *rows;
};

struct LaunchParams {
    const double *x, *k, *e;
    const uint8_t *alive;
    int64_t ld_in, n_x, n;
    int32_t n_steps;
    uint32_t flags;
    int32_t in_vec2;               // inputs allow 128-bit loads
    int32_t n_waves;               // > 1: wavelength batch, ray i is in segment #{j: i >= wave_end[j]}
    int64_t wave_end[kMaxWaves];
    unsigned long long *tile_ctr;  // in-order tile hand-out (csrc/pyr_trace.cu launch()); nullptr: static schedule
    double *spot8;                 // POLICY bit 32: spot sums of the last entry accumulated by the trace kernel
    double spot_shift[3];
    DGen gen;
    DStep steps[kMaxSteps];
    DAux aux[kMaxAux];
};

static_assert(sizeof(LaunchParams) <= 32000, "kernel parameter block is limited to 32 KB");

// ---------------------------------------------------------------------------
__device__ __forceinline__ double dot3(const double a[3], const double b[3]) {
    return fma(a[0], b[0], fma(a[1], b[1], a[2] * b[2]));
}

// y = R^T v   (global -> local direction)
__device__ __forceinline__ void rot_t(const double r[9], const double v[3], double y[3]) {
    y[0] = fma(r[0], v[0], fma(r[3], v[1], r[6] * v[2]));
    y[1] = fma(r[1], v[0], fma(r[4], v[1], r[7] * v[2]));
    y[2] = fma(r[2], v[0], fma(r[5], v[1], r[8] * v[2]));
}

// y = R v     (local -> global direction)
__device__ __forceinline__ void rot(const double r[9], const double v[3], double y[3]) {
    y[0] = fma(r[0], v[0], fma(r[1], v[1], r[2] * v[2]));
    y[1] = fma(r[3], v[0], fma(r[4], v[1], r[5] * v[2]));
    y[2] = fma(r[6], v[0], fma(r[7], v[1], r[8] * v[2]));
}

__device__ __forceinline__ void g2l_point(const DFrame &f, const double x[3], double y[3]) {
    const double t[3] = {x[0] - f.o[0], x[1] - f.o[1], x[2] - f.o[2]};
    rot_t(f.r, t, y);
}

__device__ __forceinline__ void l2g_point(const DFrame &f, const double x[3], double y[3]) {
    y[0] = fma(f.r[0], x[0], fma(f.r[1], x[1], fma(f.r[2], x[2], f.o[0])));
    y[1] = fma(f.r[3], x[0], fma(f.r[4], x[1], fma(f.r[5], x[2], f.o[1])));
    y[2] = fma(f.r[6], x[0], fma(f.r[7], x[1], fma(f.r[8], x[2], f.o[2])));
}

__device__ __forceinline__ bool finite3(const double v[3]) {
    return isfinite(v[0]) && isfinite(v[1]) && isfinite(v[2]);
}

__device__ __forceinline__ double qnan() { return __longlong_as_double(0x7ff8000000000000LL); }

// ---------------------------------------------------------------------------
// Branch-free FP64 sqrt / rsqrt / reciprocal for the hot path.
// MUFU seed (~2^-22 relative) + one third-order / two second-order Newton steps
// -> ~1 ulp.  No slow-path subroutine (CUDA's sqrt()/division carry a range check
// and a CALL): arguments here are O(1) geometric quantities, never denormal.
// Contract: a < 0 or NaN -> NaN; a == 0 -> NaN (reference: tangent ray / TIR
// boundary, invalid there as well); +-inf not expected.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double fast_rsqrt(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    const double e = fma(-a, y * y, 1.0);            // 1 - a y^2
    return fma(fma(e, 0.375, 0.5), y * e, y);        // y (1 + e/2 + 3 e^2/8)
}

__device__ __forceinline__ double fast_sqrt(double a) {
    const double y = fast_rsqrt(a);
    const double s = a * y;
    return fma(fma(-s, s, a), 0.5 * y, s);           // one Heron correction
}

// sqrt(a) and 1/sqrt(a) together
__device__ __forceinline__ double fast_sqrt_r(double a, double &rs) {
    rs = fast_rsqrt(a);
    const double s = a * rs;
    return fma(fma(-s, s, a), 0.5 * rs, s);
}

__device__ __forceinline__ double fast_rcp(double a) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double e = fma(-a, y, 1.0);
    y = fma(y, e, y);
    e = fma(-a, y, 1.0);
    return fma(y, e, y);
}

// a / b with one residual correction (~1 ulp)
__device__ __forceinline__ double fast_div(double a, double b) {
    const double y = fast_rcp(b);
    const double q = a * y;
    return fma(fma(-q, b, a), y, q);
}

}  // namespace pyr
