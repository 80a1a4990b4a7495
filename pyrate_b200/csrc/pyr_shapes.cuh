// Surface shapes: sag, gradient and ray intersection in the shape frame.
// Mirrors (does not copy) raytracer/surface_shape.py of the reference:
//   Conic.intersect :289-325, conic_function :206-219, getGrad :221-237
//   ExplicitShape.intersect :448-465 (root of z0 + t dz - F(x0 + t dx, y0 + t dy))
//   Asphere.F :529-537, gradF :539-555;  XYPolynomials.F :785-793, gradF :795-807
//   Biconic.F :618-629;  GridSag.F / gradF :866-880 (scipy RectBivariateSpline.ev);
//   LinearCombination.F / gradF :714-754
// Template parameter EXT: the kernels that carry grid-sag / combination shapes are
// separate instantiations, so the common shapes keep their register budget.
#pragma once

#include "pyr_device.cuh"

namespace pyr {

// Closed-form conic hit.  Returns t; `ok` = (F^2 + H G >= 0) as in :321.
__device__ __forceinline__ double conic_t(double curv, double cc, const double r0[3],
                                          const double d[3], bool &ok) {
    const double cc1 = 1.0 + cc;
    const double F = d[2] - curv * fma(d[0], r0[0], fma(d[1], r0[1], d[2] * r0[2] * cc1));
    const double G = curv * fma(r0[0], r0[0], fma(r0[1], r0[1], r0[2] * r0[2] * cc1)) - 2.0 * r0[2];
    const double H = -curv - cc * curv * d[2] * d[2];
    const double square = fma(F, F, H * G);
    ok = square >= 0.0;
    return fast_div(G, F + fast_sqrt(square));
}

// The same hit from the raw MUFU approximations (~2^-20 relative): a Newton SEED for shapes
// whose base conic is only the leading term (no refinement steps, no slow paths).
__device__ __forceinline__ double conic_t_seed(double curv, double cc, const double r0[3],
                                               const double d[3]) {
    const double cc1 = 1.0 + cc;
    const double F = d[2] - curv * fma(d[0], r0[0], fma(d[1], r0[1], d[2] * r0[2] * cc1));
    const double G = curv * fma(r0[0], r0[0], fma(r0[1], r0[1], r0[2] * r0[2] * cc1)) - 2.0 * r0[2];
    const double H = -curv - cc * curv * d[2] * d[2];
    const double square = fma(F, F, H * G);
    double y, ry;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(square));
    const double den = fma(square, y, F);
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(ry) : "d"(den));
    return G * ry;
}

// Unit normal of a conic at (x, y) on the vertex branch.
//   reference: z = sag(x, y); grad = (-c x, -c y, 1 - c z (1 + cc)); n = grad/|grad|
//   with s = 1 - (1 + cc) c^2 r^2 one has 1 - c z (1 + cc) = sqrt(s) and
//   |grad|^2 = 1 - cc c^2 r^2, which is what is evaluated here (s <= 0 -> NaN,
//   like conic_function :214-216).
__device__ __forceinline__ void conic_normal(double curv, double cc, bool sphere, double x,
                                             double y, double n[3]) {
    const double c2r2 = curv * curv * fma(x, x, y * y);
    const double s = fma(-(1.0 + cc), c2r2, 1.0);
    double gz = fast_sqrt(s);                           // s <= 0 -> NaN
    double gx = -curv * x, gy = -curv * y;
    if (!sphere) {
        const double inv = fast_rsqrt(fma(-cc, c2r2, 1.0));
        gx *= inv; gy *= inv; gz *= inv;
    }
    n[0] = gx; n[1] = gy; n[2] = gz;
}

// Cylinder (conic section in y, extruded along x): the conic formulas with x dropped
//   quadratic in t:  -H t^2 - 2 F t + G = 0,  H = -c (d_y^2 + (1 + cc) d_z^2)
__device__ __forceinline__ double cylinder_t(double curv, double cc, const double r0[3],
                                             const double d[3], bool &ok) {
    const double cc1 = 1.0 + cc;
    const double F = d[2] - curv * fma(d[1], r0[1], d[2] * r0[2] * cc1);
    const double G = curv * fma(r0[1], r0[1], r0[2] * r0[2] * cc1) - 2.0 * r0[2];
    const double H = -curv * fma(d[1], d[1], cc1 * d[2] * d[2]);
    const double square = fma(F, F, H * G);
    ok = square >= 0.0;
    return fast_div(G, F + fast_sqrt(square));
}

__device__ __forceinline__ void cylinder_normal(double curv, double cc, double y, double n[3]) {
    const double c2y2 = curv * curv * y * y;
    const double gz = fast_sqrt(fma(-(1.0 + cc), c2y2, 1.0));        // s <= 0 -> NaN
    const double inv = fast_rsqrt(fma(-cc, c2y2, 1.0));
    n[0] = 0.0; n[1] = -curv * y * inv; n[2] = gz * inv;
}

__device__ __forceinline__ double conic_sag(double curv, double cc, double x, double y) {
    const double r2 = fma(x, x, y * y);
    const double s = fma(-(1.0 + cc) * curv * curv, r2, 1.0);
    return curv * r2 * fast_rcp(1.0 + fast_sqrt(s));    // s <= 0 -> NaN
}

// Asphere: value and gradient (gx, gy; gz = 1) of z - F(x, y) in one pass.
__device__ __forceinline__ void asphere_eval(const DAux &a, double curv, double cc, double x,
                                             double y, double &F, double &Fx, double &Fy) {
    const double r2 = fma(x, x, y * y);
    double rsq;
    const double sq = fast_sqrt_r(fma(-curv * curv * (1.0 + cc), r2, 1.0), rsq);   // NaN outside
    // polynomial part: sum a_n r2^(n+1) and its r2-derivative by Horner
    double p = 0.0, dp = 0.0;
    for (int i = a.n_coeff - 1; i >= 0; --i) {
        dp = fma(dp, r2, p);
        p = fma(p, r2, a.coeff[i]);
    }
    // p(r2) = sum a_i r2^i  ->  poly = r2 p,  d poly / d r2 = p + r2 dp
    F = fma(curv * r2, fast_rcp(1.0 + sq), r2 * p);
    const double dr = fma(curv, rsq, 2.0 * fma(r2, dp, p));            // dF/dx = x * dr
    Fx = x * dr;
    Fy = y * dr;
}

// XY polynomial: F = sum c x^m y^n / R^(m+n)
__device__ __forceinline__ void xypoly_eval(const DAux &a, double x, double y, double &F,
                                            double &Fx, double &Fy) {
    const double ir = 1.0 / a.normradius;
    const double xs = x * ir, ys = y * ir;
    double f = 0.0, fx = 0.0, fy = 0.0;
    for (int i = 0; i < a.n_coeff; ++i) {
        const int m = a.xpow[i], n = a.ypow[i];
        double xm1 = 1.0, yn1 = 1.0;                   // xs^(m-1), ys^(n-1)
        for (int j = 1; j < m; ++j) xm1 *= xs;
        for (int j = 1; j < n; ++j) yn1 *= ys;
        const double xm = (m >= 1) ? xm1 * xs : 1.0;
        const double yn = (n >= 1) ? yn1 * ys : 1.0;
        const double c = a.coeff[i];
        f = fma(c, xm * yn, f);
        if (m >= 1) fx = fma(c * m, xm1 * yn, fx);
        if (n >= 1) fy = fma(c * n, xm * yn1, fy);
    }
    F = f;
    Fx = fx * ir;
    Fy = fy * ir;
}

// Biconic (surface_shape.py:609-706): F = (cx x^2 + cy y^2) / (1 + sq) + sum a_n u_n^(n+1),
// sq = sqrt(1 - cx^2 (1+kx) x^2 - cy^2 (1+ky) y^2), u_n = r^2 - b_n (x^2 - y^2)
__device__ __forceinline__ void biconic_eval(const DAux &a, double cx, double kx, double x, double y,
                                             double &F, double &Fx, double &Fy) {
    const double cy = a.curv2, ky = a.cc2;
    const double x2 = x * x, y2 = y * y;
    const double N = fma(cx, x2, cy * y2);
    double rsq;
    const double sq = fast_sqrt_r(fma(-cx * cx * (1.0 + kx), x2, fma(-cy * cy * (1.0 + ky), y2, 1.0)), rsq);
    const double inv1 = fast_rcp(1.0 + sq);
    const double common = inv1 * inv1 * rsq;
    const double two = 2.0 * (sq + 1.0) * sq;
    double f = N * inv1;
    double fx = cx * x * fma(cx * (kx + 1.0), N, two) * common;
    double fy = cy * y * fma(cy * (ky + 1.0), N, two) * common;
    const double r2 = x2 + y2, ast2 = x2 - y2;
    for (int i = 0; i < a.n_coeff; ++i) {
        const double an = a.coeff[i], bn = a.coeff[16 + i];
        const double u = fma(-bn, ast2, r2);
        double un = 1.0;                                   // u^i
        for (int j = 0; j < i; ++j) un *= u;
        f = fma(an * un, u, f);
        const double g = 2.0 * an * (i + 1) * un;
        fx = fma(g * (1.0 - bn), x, fx);
        fy = fma(g * (1.0 + bn), y, fy);
    }
    F = f; Fx = fx; Fy = fy;
}

// Cubic B-spline basis values h[0..3] and derivatives dh[0..3] of the knot interval
// t[l] <= x < t[l+1] (de Boor / Cox recursion, FITPACK fpbspl); they weight the
// coefficients l-3 .. l.
__device__ __forceinline__ void bspline_basis(const double *__restrict__ t, int l, double x,
                                              double h[4], double dh[4]) {
    double tm[3], tp[3];                               // t[l-2..l], t[l+1..l+3]
#pragma unroll
    for (int i = 0; i < 3; ++i) { tm[i] = __ldg(t + l - 2 + i); tp[i] = __ldg(t + l + 1 + i); }
    // degree 1
    const double f1 = 1.0 / (tp[0] - tm[2]);
    double a0 = f1 * (tp[0] - x), a1 = f1 * (x - tm[2]);
    // degree 2: basis l-2, l-1, l
    const double g0 = a0 / (tp[0] - tm[1]), g1 = a1 / (tp[1] - tm[2]);
    const double q0 = g0 * (tp[0] - x);
    const double q1 = fma(g0, x - tm[1], g1 * (tp[1] - x));
    const double q2 = g1 * (x - tm[2]);
    // degree 3: basis l-3 .. l
    const double r0 = 1.0 / (tp[0] - tm[0]), r1 = 1.0 / (tp[1] - tm[1]), r2 = 1.0 / (tp[2] - tm[2]);
    const double e0 = q0 * r0, e1 = q1 * r1, e2 = q2 * r2;
    h[0] = e0 * (tp[0] - x);
    h[1] = fma(e0, x - tm[0], e1 * (tp[1] - x));
    h[2] = fma(e1, x - tm[1], e2 * (tp[2] - x));
    h[3] = e2 * (x - tm[2]);
    // B'_{i,3} = 3 (B_{i,2} / (t_{i+3} - t_i) - B_{i+1,2} / (t_{i+4} - t_{i+1}))
    dh[0] = -3.0 * e0;
    dh[1] = 3.0 * (e0 - e1);
    dh[2] = 3.0 * (e1 - e2);
    dh[3] = 3.0 * e2;
}

// knot interval of x in t[3 .. n-4] (x already clamped): largest l in [3, n-5] with t[l] <= x
__device__ __forceinline__ int bspline_interval(const double *__restrict__ t, int n, double x) {
    int lo = 3, hi = n - 5;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(t + mid) <= x) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// Grid sag: value and gradient of the bicubic spline, arguments clamped to the grid
// like FITPACK's bispev / parder (what RectBivariateSpline.ev does).
__device__ __forceinline__ void gridsag_eval(const DAux &a, double x, double y, double &F,
                                             double &Fx, double &Fy) {
    const int nx = a.grid_nx, ny = a.grid_ny;
    const double xc = fmin(fmax(x, __ldg(a.grid_tx + 3)), __ldg(a.grid_tx + nx - 4));
    const double yc = fmin(fmax(y, __ldg(a.grid_ty + 3)), __ldg(a.grid_ty + ny - 4));
    const int lx = bspline_interval(a.grid_tx, nx, xc);
    const int ly = bspline_interval(a.grid_ty, ny, yc);
    double hx[4], dhx[4], hy[4], dhy[4];
    bspline_basis(a.grid_tx, lx, xc, hx, dhx);
    bspline_basis(a.grid_ty, ly, yc, hy, dhy);
    const int ncy = ny - 4;
    const double *c = a.grid_c + (int64_t)(lx - 3) * ncy + (ly - 3);
    double f = 0.0, fx = 0.0, fy = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double row = 0.0, drow = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double cij = __ldg(c + (int64_t)i * ncy + j);
            row = fma(cij, hy[j], row);
            drow = fma(cij, dhy[j], drow);
        }
        f = fma(hx[i], row, f);
        fx = fma(dhx[i], row, fx);
        fy = fma(hx[i], drow, fy);
    }
    F = f; Fx = fx; Fy = fy;
}

template <bool EXT>
__device__ __forceinline__ void simple_eval(int kind, const DAux &a, double curv, double cc,
                                            double x, double y, double &F, double &Fx,
                                            double &Fy) {
    if (kind == PYR_SHAPE_ASPHERE) asphere_eval(a, curv, cc, x, y, F, Fx, Fy);
    else if (kind == PYR_SHAPE_BICONIC) biconic_eval(a, curv, cc, x, y, F, Fx, Fy);
    else if (EXT && kind == PYR_SHAPE_GRIDSAG) gridsag_eval(a, x, y, F, Fx, Fy);
    else xypoly_eval(a, x, y, F, Fx, Fy);
}

template <bool EXT>
__device__ __forceinline__ void explicit_eval(int kind, const DAux &a, double curv, double cc,
                                              double x, double y, double &F, double &Fx,
                                              double &Fy) {
    if (EXT && kind == PYR_SHAPE_COMBINATION) {
        // terms live in consecutive auxiliary records (pack() lays them out)
        double f = 0.0, fx = 0.0, fy = 0.0;
        for (int t = 0; t < a.n_terms; ++t) {
            const DAux &at = (&a)[t];
            double tf, tfx, tfy;
            simple_eval<EXT>(at.term_kind, at, at.term_curv, at.term_cc, x - at.term_dx,
                             y - at.term_dy, tf, tfx, tfy);
            f = fma(at.term_w, tf + at.term_dz, f);
            fx = fma(at.term_w, tfx, fx);
            fy = fma(at.term_w, tfy, fy);
        }
        F = f; Fx = fx; Fy = fy;
        return;
    }
    simple_eval<EXT>(kind, a, curv, cc, x, y, F, Fx, Fy);
}

template <bool EXT>
__device__ __forceinline__ double shape_sag(int kind, const DAux *a, double curv, double cc,
                                            double x, double y) {
    if (kind == PYR_SHAPE_CONIC) return conic_sag(curv, cc, x, y);
    if (kind == PYR_SHAPE_CYLINDER) return conic_sag(curv, cc, 0.0, y);
    double F, Fx, Fy;
    explicit_eval<EXT>(kind, *a, curv, cc, x, y, F, Fx, Fy);
    return F;
}

// Newton iteration on f(t) = z0 + t dz - F(x0 + t dx, y0 + t dy), seeded with the
// base-conic hit (plane for XY polynomials), for the N rays a thread owns TOGETHER: one
// loop, N independent dependency chains inside its body (the iteration is a long serial
// FP64 chain; at 16 resident warps per SM a single chain per thread leaves the pipe idle).
// The warp leaves the loop together (__all_sync vote on the step sizes), capped at
// `maxit`.  On return (gx, gy) hold dF/dx, dF/dy of the last evaluation and `grad_ok` tells
// whether that evaluation was within the convergence tolerance of the returned point --
// then the surface normal can reuse it instead of evaluating the shape once more.
template <bool EXT, int N>
__device__ __forceinline__ void explicit_t_n(int kind, const DAux &a, double curv, double cc,
                                             const double (&r0)[N][3], const double (&d)[N][3],
                                             const bool (&active)[N], double (&t)[N],
                                             double (&gx)[N], double (&gy)[N], bool (&grad_ok)[N]) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
        bool ok;
        if (kind == PYR_SHAPE_ASPHERE) t[j] = conic_t(curv, cc, r0[j], d[j], ok);
        else if (kind == PYR_SHAPE_BICONIC) t[j] = conic_t(0.5 * (curv + a.curv2), 0.5 * (cc + a.cc2), r0[j], d[j], ok);
        else if (EXT && kind == PYR_SHAPE_COMBINATION && a.term_kind == PYR_SHAPE_ASPHERE &&
                 a.term_dx == 0.0 && a.term_dy == 0.0)
            t[j] = conic_t(a.term_w * a.term_curv, a.term_cc, r0[j], d[j], ok);   // leading base conic
        else t[j] = -r0[j][2] * fast_rcp(d[j][2]);
        if (!isfinite(t[j])) t[j] = 0.0;
        grad_ok[j] = false;
        gx[j] = gy[j] = 0.0;
    }
    for (int it = 0; it < a.newton_maxit; ++it) {
        bool done = true;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const double x = fma(t[j], d[j][0], r0[j][0]);
            const double y = fma(t[j], d[j][1], r0[j][1]);
            double F;
            explicit_eval<EXT>(kind, a, curv, cc, x, y, F, gx[j], gy[j]);
            const double res = fma(t[j], d[j][2], r0[j][2]) - F;
            const double dres = d[j][2] - fma(gx[j], d[j][0], gy[j] * d[j][1]);
            double step = fast_div(res, dres);
            const bool bad = !isfinite(step);
            if (bad) step = 0.0;
            t[j] -= step;
            const bool conv = fabs(step) <= a.newton_tol * (1.0 + fabs(t[j]));
            grad_ok[j] = conv && !bad;
            done = done && (bad || !active[j] || conv);
        }
        if (__all_sync(__activemask(), done)) break;
    }
}

// one ray (the crystal kernel's explicit-shape path)
template <bool EXT>
__device__ __forceinline__ double explicit_t(int kind, const DAux &a, double curv, double cc,
                                             const double r0[3], const double d[3],
                                             bool active, double &gx, double &gy, bool &grad_ok) {
    const double r0n[1][3] = {{r0[0], r0[1], r0[2]}}, dn[1][3] = {{d[0], d[1], d[2]}};
    const bool act[1] = {active};
    double t[1], gxn[1], gyn[1];
    bool gok[1];
    explicit_t_n<EXT, 1>(kind, a, curv, cc, r0n, dn, act, t, gxn, gyn, gok);
    gx = gxn[0]; gy = gyn[0]; grad_ok = gok[0];
    return t[0];
}

// 1 / a to ~2^-40 (MUFU seed + one Newton step): enough for a Newton step whose own error is
// corrected by the next iteration (the last step of a converged iteration is < 1e-6 |t|)
__device__ __forceinline__ double rcp_nr1(double a) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    return fma(y, fma(-a, y, 1.0), y);
}

// p = sum a_i r2^i and p' = dp / d(r2) by Horner for N rays.  Up to four coefficients (the usual
// even asphere: A2 .. A8) run as a fixed four-step chain -- the record is zero-padded beyond
// n_coeff, and a leading zero coefficient leaves the exact value -- without the loop's branches.
template <int N>
__device__ __forceinline__ void asphere_horner(const DAux &a, int nc, const double (&r2)[N], double (&p)[N],
                                               double (&dp)[N]) {
    if (nc <= 4) {
        const double c3 = a.coeff[3], c2 = a.coeff[2], c1 = a.coeff[1], c0 = a.coeff[0];
#pragma unroll
        for (int j = 0; j < N; ++j) {
            double pj = c3, dj = 0.0;
            dj = fma(dj, r2[j], pj); pj = fma(pj, r2[j], c2);
            dj = fma(dj, r2[j], pj); pj = fma(pj, r2[j], c1);
            dj = fma(dj, r2[j], pj); pj = fma(pj, r2[j], c0);
            p[j] = pj; dp[j] = dj;
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < N; ++j) { p[j] = 0.0; dp[j] = 0.0; }
    for (int i = nc - 1; i >= 0; --i) {
        const double ci = a.coeff[i];
#pragma unroll
        for (int j = 0; j < N; ++j) {
            dp[j] = fma(dp[j], r2[j], p[j]);
            p[j] = fma(p[j], r2[j], ci);
        }
    }
}

// Even asphere (the common explicit shape), N rays of a thread together, WITHOUT a square
// root or a full-precision division inside the iteration: with q(r^2) = sum a_i r^(2i+2) the
// hit satisfies  w = z - q(r^2) = conic sag(r^2), and on the vertex branch of the conic that
// is the root of the conic's implicit function
//     g(t) = c r^2 + c (1 + cc) w^2 - 2 w,      r^2, z along the ray,
// (same root as z - F = 0 of surface_shape.py:448-465: near it g = -2 sq (w - sag) with
// sq = sqrt(1 - (1+cc) c^2 r^2) > 0).  Newton on g, seeded with the base-conic hit, costs one
// Horner pass for q, q' and ~20 FP64 operations per iteration (the seed only needs the MUFU
// approximations: its distance to the asphere's root is the polynomial term anyway).  It
// converges quadratically:
// step_(k+1) ~ C step_k^2 with C estimated from the last two steps, so once the PREDICTED
// next step is two orders below the tolerance the iteration stops without the confirming
// evaluation.  The gradient (dF/dx, dF/dy) = (x, y) dr is then evaluated AT the returned point;
// there sq = 1 - c (1 + cc) w (surface equation), so no square root is needed either, and a
// point on the far branch (sq <= 0: the explicit sag is undefined) returns NaN like F does.
template <int N>
__device__ __forceinline__ void asphere_t_n(const DAux &a, double curv, double cc,
                                            const double (&r0)[N][3], const double (&d)[N][3],
                                            const bool (&active)[N], double (&t)[N],
                                            double (&gx)[N], double (&gy)[N]) {
    const double cK1 = curv * (1.0 + cc);
    double prev[N];                                      // |previous step|, 0 = none yet
#pragma unroll
    for (int j = 0; j < N; ++j) {
        t[j] = conic_t_seed(curv, cc, r0[j], d[j]);
        if (!isfinite(t[j])) t[j] = 0.0;
        prev[j] = 0.0;
    }
    const int nc = a.n_coeff;
    for (int it = 0; it < a.newton_maxit; ++it) {
        double x[N], y[N], r2[N], p[N], dp[N];
#pragma unroll
        for (int j = 0; j < N; ++j) {
            x[j] = fma(t[j], d[j][0], r0[j][0]);
            y[j] = fma(t[j], d[j][1], r0[j][1]);
            r2[j] = fma(x[j], x[j], y[j] * y[j]);
            p[j] = 0.0; dp[j] = 0.0;
        }
        asphere_horner<N>(a, nc, r2, p, dp);
        bool done = true;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const double z = fma(t[j], d[j][2], r0[j][2]);
            const double w = fma(-r2[j], p[j], z);                         // z - q
            const double g = fma(curv, r2[j], w * fma(cK1, w, -2.0));
            const double xy = fma(x[j], d[j][0], y[j] * d[j][1]);          // (d r^2 / dt) / 2
            const double wp = fma(-2.0 * fma(r2[j], dp[j], p[j]), xy, d[j][2]);   // dw/dt
            const double hg = fma(curv, xy, fma(cK1, w, -1.0) * wp);       // g' / 2
            double step = 0.5 * g * rcp_nr1(hg);
            const bool bad = !isfinite(step);
            if (bad) step = 0.0;
            t[j] -= step;
            const double as = fabs(step), lim = a.newton_tol * (1.0 + fabs(t[j]));
            const bool conv = as <= lim;
            // predicted next step |step|^2 C, C = |step| / prev^2; two orders of safety
            const bool early = prev[j] > 0.0 && as * as * as <= 0.01 * lim * prev[j] * prev[j];
            prev[j] = as;
            done = done && (bad || !active[j] || conv || early);
        }
        if (__all_sync(__activemask(), done)) break;
    }
    // gradient factor at the returned point: dr = c / sq + 2 (p + r^2 p')
    {
        double x[N], y[N], r2[N], p[N], dp[N];
#pragma unroll
        for (int j = 0; j < N; ++j) {
            x[j] = fma(t[j], d[j][0], r0[j][0]);
            y[j] = fma(t[j], d[j][1], r0[j][1]);
            r2[j] = fma(x[j], x[j], y[j] * y[j]);
            p[j] = 0.0; dp[j] = 0.0;
        }
        asphere_horner<N>(a, nc, r2, p, dp);
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const double w = fma(-r2[j], p[j], fma(t[j], d[j][2], r0[j][2]));
            const double sq = fma(-cK1, w, 1.0);
            const double dr = fma(curv, fast_rcp(sq), 2.0 * fma(r2[j], dp[j], p[j]));
            gx[j] = x[j] * dr; gy[j] = y[j] * dr;
            if (!(sq > 0.0)) { t[j] = qnan(); gx[j] = gy[j] = qnan(); }
        }
    }
}

// unit normal of an explicit shape: grad = (-Fx, -Fy, 1)/|.|  (FreeShape.getGrad :420-423)
template <bool EXT>
__device__ __forceinline__ void explicit_normal(int kind, const DAux &a, double curv, double cc,
                                                double x, double y, double n[3]) {
    double F, Fx, Fy;
    explicit_eval<EXT>(kind, a, curv, cc, x, y, F, Fx, Fy);
    const double inv = fast_rsqrt(fma(Fx, Fx, fma(Fy, Fy, 1.0)));
    n[0] = -Fx * inv; n[1] = -Fy * inv; n[2] = inv;
}

__device__ __forceinline__ void normal_from_gradient(double Fx, double Fy, double n[3]) {
    const double inv = fast_rsqrt(fma(Fx, Fx, fma(Fy, Fy, 1.0)));
    n[0] = -Fx * inv; n[1] = -Fy * inv; n[2] = inv;
}

}  // namespace pyr
