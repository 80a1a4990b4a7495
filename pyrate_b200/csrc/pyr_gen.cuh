// Bundle generation in registers: raster point -> start point, wave vector, field.
// Mirrors the callers right above seqtrace in the reference:
//   raytracer/analysis/optical_system_analysis.py:83-125 collimated_bundle,
//   :127-165 divergent_bundle;  sampling2d/raster.py:36-60 RectGrid, :62-91 HexGrid,
//   :150-166 CircularGrid;  plus the hexapolar raster of BASELINE.json's bundles.
// Every operation the host does in NumPy with separately rounded multiply / add
// (linspace, radius * p + start) is done with explicitly rounded __dmul_rn / __dadd_rn
// here, so lattice coordinates are bit-identical to the host rasters; only sin / cos
// (hexapolar, circular, divergent) differ from libm by <= 2 ulp.
#pragma once

#include "pyr_device.cuh"

namespace pyr {

__device__ __forceinline__ double gen_lin(const DGen &g, int64_t i) {
    // numpy.linspace: y = arange(n) * step + start, y[-1] = stop
    if (i == g.param - 1 && g.param > 1) return g.lin_stop;
    return __dadd_rn(__dmul_rn((double)i, g.lin_step), g.lin_start);
}

// largest r in [0, nrows) with rows[r] <= idx  (rows[nrows] = total > idx)
__device__ __forceinline__ int64_t gen_row_of(const int64_t *__restrict__ rows, int64_t nrows,
                                              int64_t idx) {
    int64_t lo = 0, hi = nrows;          // invariant: rows[lo] <= idx < rows[hi]
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(rows + mid) <= idx) lo = mid; else hi = mid;
    }
    return lo;
}

// normalised raster coordinates of point idx
__device__ __forceinline__ void gen_point(const DGen &g, int64_t idx, double &px, double &py) {
    if (g.raster == PYR_RASTER_HEXAPOLAR) {
        if (idx <= 0) { px = 0.0; py = 0.0; return; }
        // ring j holds the indices [1 + 3 (j - 1) j, 1 + 3 j (j + 1))
        // branch-free approximations + exact integer fix-up; the angle 2 pi i / (6 j) is
        // evaluated as sincospi(i / (3 j)) (no Payne-Hanek reduction in the instruction
        // stream: the generating kernels must stay small); <= 2 ulp from the host raster
        int64_t j = (int64_t)((3.0 + fast_sqrt(9.0 + 12.0 * (double)(idx - 1))) * (1.0 / 6.0));
        if (j < 1) j = 1;
        if (idx < 1 + 3 * (j - 1) * j) --j;
        if (idx >= 1 + 3 * j * (j + 1)) ++j;
        const int64_t i = idx - (1 + 3 * (j - 1) * j);
        const double frac = fast_div((double)i, 3.0 * (double)j);
        const double rad = (double)j * g.aux0;                      // aux0 = 1 / rings
        double s, c;
        sincospi(frac, &s, &c);
        px = rad * c;
        py = rad * s;
    } else if (g.raster == PYR_RASTER_RECT) {
        const int64_t nrows = g.param;
        const int64_t r = gen_row_of(g.rows, nrows, idx);
        const int64_t ix = __ldg(g.rows + nrows + 1 + r) + (idx - __ldg(g.rows + r));
        px = gen_lin(g, ix);
        py = gen_lin(g, r);
    } else if (g.raster == PYR_RASTER_HEX) {
        const int64_t nx = g.param, nrows = 2 * nx;
        const int64_t r = gen_row_of(g.rows, nrows, idx);
        const int64_t ix = __ldg(g.rows + nrows + 1 + r) + (idx - __ldg(g.rows + r));
        const bool second = r >= nx;
        px = gen_lin(g, ix);
        py = __dmul_rn(gen_lin(g, second ? r - nx : r), 1.7320508075688772);
        if (second) { px = __dadd_rn(px, g.aux0); py = __dadd_rn(py, g.aux1); }
    } else {   // PYR_RASTER_CIRCULAR: index = iphi * m + ir
        const int64_t m = g.param > 0 ? g.param : 1;
        const int64_t iphi = idx / m, ir = idx - iphi * m;
        double r = (ir == m - 1 && m > 1) ? 1.0 : __dmul_rn((double)ir, g.lin_step);
        if (g.flags & PYR_GEN_SQRT_R) r = sqrt(r);
        const double phi = __dmul_rn((double)iphi, g.aux0);
        double s, c;
        sincos(phi, &s, &c);
        px = __dmul_rn(r, c);
        py = __dmul_rn(r, s);
    }
}

struct GenRay {
    double x[3], k[3], e[3];
};

// ray `i` of the call: x, k, e (global frame) -- every raster / bundle kind
__device__ __forceinline__ void gen_ray_any(const DGen &g, int64_t i, double x[3], double k[3],
                                            double e[3]) {
    double px, py;
    gen_point(g, g.first + i, px, py);
    double d[3];
    if (g.bundle == PYR_BUNDLE_COLLIMATED) {
        x[0] = __dadd_rn(__dmul_rn(g.radius, px), g.start[0]);
        x[1] = __dadd_rn(__dmul_rn(g.radius, py), g.start[1]);
        x[2] = g.start[2];
        d[0] = g.dir[0]; d[1] = g.dir[1]; d[2] = g.dir[2];
    } else {
        x[0] = g.start[0]; x[1] = g.start[1]; x[2] = g.start[2];
        const double ay = __dadd_rn(g.dir[0], __dmul_rn(g.radius, px));
        const double ax = __dadd_rn(g.dir[1], __dmul_rn(g.radius, py));
        double say, cay, sax, cax;
        sincos(ay, &say, &cay);
        sincos(ax, &sax, &cax);
        d[0] = __dmul_rn(say, cax);
        d[1] = sax;
        d[2] = __dmul_rn(cay, cax);
    }
    k[0] = __dmul_rn(g.n_index, d[0]);
    k[1] = __dmul_rn(g.n_index, d[1]);
    k[2] = __dmul_rn(g.n_index, d[2]);
    if (g.flags & PYR_GEN_E_PERP) {
        // axis least aligned with d (first minimum, like numpy.argmin), Gram-Schmidt
        const double a0 = fabs(d[0]), a1 = fabs(d[1]), a2 = fabs(d[2]);
        const int ax = (a0 <= a1 && a0 <= a2) ? 0 : (a1 <= a2 ? 1 : 2);
        const double dax = (ax == 0) ? d[0] : (ax == 1 ? d[1] : d[2]);
        const double c = dax / dot3(d, d);
        double t[3] = {-c * d[0], -c * d[1], -c * d[2]};
        t[0] += (ax == 0) ? 1.0 : 0.0;
        t[1] += (ax == 1) ? 1.0 : 0.0;
        t[2] += (ax == 2) ? 1.0 : 0.0;
        const double inv = 1.0 / sqrt(dot3(t, t));
        e[0] = t[0] * inv; e[1] = t[1] * inv; e[2] = t[2] * inv;
    } else {
        e[0] = g.e[0]; e[1] = g.e[1]; e[2] = g.e[2];
    }
}

// the general generator behind a call (results by value: registers), so that the trace
// kernels inline only the hexapolar / collimated / fixed-field case of BASELINE's bundles
// and stay small (instruction cache)
__device__ __noinline__ GenRay gen_ray_call(const DGen &g, int64_t i) {
    GenRay r;
    gen_ray_any(g, i, r.x, r.k, r.e);
    return r;
}

__device__ __forceinline__ void gen_ray(const DGen &g, int64_t i, double x[3], double k[3],
                                        double e[3]) {
    if (g.raster == PYR_RASTER_HEXAPOLAR && g.bundle == PYR_BUNDLE_COLLIMATED &&
        !(g.flags & PYR_GEN_E_PERP)) {
        const int64_t idx = g.first + i;
        double px = 0.0, py = 0.0;
        if (idx > 0) {
            int64_t j = (int64_t)((3.0 + fast_sqrt(9.0 + 12.0 * (double)(idx - 1))) * (1.0 / 6.0));
            if (j < 1) j = 1;
            if (idx < 1 + 3 * (j - 1) * j) --j;
            if (idx >= 1 + 3 * j * (j + 1)) ++j;
            const int64_t m = idx - (1 + 3 * (j - 1) * j);
            double s, c;
            sincospi(fast_div((double)m, 3.0 * (double)j), &s, &c);
            const double rad = (double)j * g.aux0;
            px = rad * c;
            py = rad * s;
        }
        x[0] = __dadd_rn(__dmul_rn(g.radius, px), g.start[0]);
        x[1] = __dadd_rn(__dmul_rn(g.radius, py), g.start[1]);
        x[2] = g.start[2];
#pragma unroll
        for (int c = 0; c < 3; ++c) { k[c] = __dmul_rn(g.n_index, g.dir[c]); e[c] = g.e[c]; }
        return;
    }
    const GenRay r = gen_ray_call(g, i);
#pragma unroll
    for (int c = 0; c < 3; ++c) { x[c] = r.x[c]; k[c] = r.k[c]; e[c] = r.e[c]; }
}

}  // namespace pyr
