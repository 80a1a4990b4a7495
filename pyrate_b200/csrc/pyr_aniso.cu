// Complex-valued trace kernel: anisotropic (birefringent) media and everything
// downstream of them (the reference keeps k and E complex128 from the first
// crystal on).
//
// Reference path replaced (pyrateoptics/raytracer/material, file:line):
//   material_anisotropic.py:70-155   refract / reflect (mode selection, ray doubling)
//   material.py:122-153              sortKnormEField (Poynting sort)
//   material.py:353-454              quadratic eigenvalue problem for xi and E
//   material.py:214-223              Poynting vector
//   material_isotropic.py:137-199    isotropic deflection with complex k
//
// The reference solves a 6x6 generalised eigenproblem per ray (scipy.linalg.eig).
// Here the same eigenvalues are the roots of the Fresnel quartic
//   D(k) = det(eps) - (k.k) c2(eps) + k^T adj(eps) k + (k.k)(k^T eps k),  k = kpa + xi n
// (det(-(k.k) I + k k^T + eps) expanded with the matrix determinant lemma; the
// reference's own tests pin eigenvalues == polynomial roots, tests/test_material.py
// :174-327), found with Aberth-Ehrlich iterations, and E is the null vector of the
// 3x3 propagator at each root (cross product of its two most independent rows).
#include <cuda_runtime.h>

#include <cstring>

#include "pyr_device.cuh"
#include "pyr_shapes.cuh"

namespace pyr {

struct cplx {
    double re, im;
};
__device__ __forceinline__ cplx C(double r, double i = 0.0) { return cplx{r, i}; }
__device__ __forceinline__ cplx operator+(cplx a, cplx b) { return {a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ cplx operator-(cplx a, cplx b) { return {a.re - b.re, a.im - b.im}; }
__device__ __forceinline__ cplx operator-(cplx a) { return {-a.re, -a.im}; }
__device__ __forceinline__ cplx operator*(cplx a, cplx b) {
    return {fma(a.re, b.re, -a.im * b.im), fma(a.re, b.im, a.im * b.re)};
}
__device__ __forceinline__ cplx operator*(double s, cplx a) { return {s * a.re, s * a.im}; }
__device__ __forceinline__ cplx conj(cplx a) { return {a.re, -a.im}; }
__device__ __forceinline__ double abs2(cplx a) { return fma(a.re, a.re, a.im * a.im); }
__device__ __forceinline__ cplx operator/(cplx a, cplx b) {
    const double d = 1.0 / abs2(b);
    return {(a.re * b.re + a.im * b.im) * d, (a.im * b.re - a.re * b.im) * d};
}
__device__ __forceinline__ cplx csqrt_(cplx z) {       // principal branch, like numpy
    const double m = hypot(z.re, z.im);
    if (m == 0.0) return {0.0, 0.0};
    double a = sqrt(0.5 * (m + fabs(z.re)));
    double b = 0.5 * z.im / a;
    if (z.re >= 0.0) return {a, b};
    return {fabs(b), copysign(a, z.im)};
}
__device__ __forceinline__ bool cfinite(cplx a) { return isfinite(a.re) && isfinite(a.im); }

// bilinear products (no conjugation), as np.sum(a * b, axis=0)
__device__ __forceinline__ cplx cdot(const cplx a[3], const cplx b[3]) {
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
__device__ __forceinline__ cplx cdotr(const cplx a[3], const double b[3]) {
    return {fma(a[0].re, b[0], fma(a[1].re, b[1], a[2].re * b[2])),
            fma(a[0].im, b[0], fma(a[1].im, b[1], a[2].im * b[2]))};
}
__device__ __forceinline__ double herm2(const cplx a[3]) { return abs2(a[0]) + abs2(a[1]) + abs2(a[2]); }

__device__ __forceinline__ void ccross(const cplx a[3], const cplx b[3], cplx c[3]) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

// y = R^T v / y = R v for complex vectors with a real rotation
__device__ __forceinline__ void crot_t(const double r[9], const cplx v[3], cplx y[3]) {
    for (int i = 0; i < 3; ++i)
        y[i] = {fma(r[i], v[0].re, fma(r[3 + i], v[1].re, r[6 + i] * v[2].re)),
                fma(r[i], v[0].im, fma(r[3 + i], v[1].im, r[6 + i] * v[2].im))};
}
__device__ __forceinline__ void crot(const double r[9], const cplx v[3], cplx y[3]) {
    for (int i = 0; i < 3; ++i)
        y[i] = {fma(r[3 * i], v[0].re, fma(r[3 * i + 1], v[1].re, r[3 * i + 2] * v[2].re)),
                fma(r[3 * i], v[0].im, fma(r[3 * i + 1], v[1].im, r[3 * i + 2] * v[2].im))};
}

// Poynting direction, ray.py:140-152 (complex k, E)
__device__ __forceinline__ void poynting_vec(const cplx k[3], const cplx e[3], double s[3]) {
    const double ee = herm2(e);
    const cplx ek = cdot(e, k);
    for (int i = 0; i < 3; ++i) s[i] = ee * k[i].re - (ek * conj(e[i])).re;
}

// ---------------------------------------------------------------------------
// Fresnel quartic in xi and its roots
// ---------------------------------------------------------------------------
struct EpsInv {           // invariants of eps used by the quartic
    cplx eps[9];
    cplx adj[9];
    cplx c2, det;
};

__device__ __forceinline__ void eps_invariants(const double e18[18], EpsInv &v) {
    for (int i = 0; i < 9; ++i) v.eps[i] = {e18[2 * i], e18[2 * i + 1]};
    const cplx *m = v.eps;
    // adjugate (transpose of the cofactor matrix)
    v.adj[0] = m[4] * m[8] - m[5] * m[7];
    v.adj[1] = m[2] * m[7] - m[1] * m[8];
    v.adj[2] = m[1] * m[5] - m[2] * m[4];
    v.adj[3] = m[5] * m[6] - m[3] * m[8];
    v.adj[4] = m[0] * m[8] - m[2] * m[6];
    v.adj[5] = m[2] * m[3] - m[0] * m[5];
    v.adj[6] = m[3] * m[7] - m[4] * m[6];
    v.adj[7] = m[1] * m[6] - m[0] * m[7];
    v.adj[8] = m[0] * m[4] - m[1] * m[3];
    v.det = m[0] * v.adj[0] + m[1] * v.adj[3] + m[2] * v.adj[6];
    v.c2 = v.adj[0] + v.adj[4] + v.adj[8];          // sum of principal 2x2 minors
}

__device__ __forceinline__ cplx quad_form(const cplx m[9], const cplx a[3], const cplx b[3]) {
    cplx r = C(0.0);
    for (int i = 0; i < 3; ++i) r = r + a[i] * (m[3 * i] * b[0] + m[3 * i + 1] * b[1] + m[3 * i + 2] * b[2]);
    return r;
}

__device__ __forceinline__ cplx poly_eval(const cplx c[5], cplx z, cplx &dp) {
    cplx p = c[4];
    dp = C(0.0);
    for (int i = 3; i >= 0; --i) {
        dp = dp * z + p;
        p = p * z + c[i];
    }
    return p;
}

// roots of c[4] z^4 + ... + c[0]; returns false if the iteration met non-finite data.
// Aberth-Ehrlich sweeps (cubic convergence) from the four starts z0[]; once every
// correction is below 1e-9 relative, exactly one more sweep lands on the rounding
// floor -- no tolerance at the 1e-16 level that rounding noise could keep missing.
__device__ __forceinline__ bool quartic_roots(const cplx c[5], const cplx z0[4], double scale,
                                              cplx z[4]) {
    if (!(abs2(c[4]) > 0.0) || !isfinite(scale)) return false;
    for (int i = 0; i < 4; ++i) z[i] = z0[i];
    bool last = false;
    for (int it = 0; it < 60; ++it) {
        double worst = 0.0;
        for (int i = 0; i < 4; ++i) {
            cplx dp;
            const cplx p = poly_eval(c, z[i], dp);
            if (abs2(p) == 0.0) continue;
            cplx w = p / dp;                                  // Newton correction
            if (!cfinite(w)) { w = cplx{1e-8 * scale, 1e-8 * scale}; }
            cplx rep = C(0.0);
            for (int j = 0; j < 4; ++j)
                if (j != i) {
                    cplx dz = z[i] - z[j];
                    if (abs2(dz) == 0.0) dz = cplx{1e-16 * scale, 1e-16 * scale};
                    rep = rep + C(1.0) / dz;
                }
            cplx step = w / (C(1.0) - w * rep);
            if (!cfinite(step)) step = w;
            z[i] = z[i] - step;
            worst = fmax(worst, abs2(step) / (abs2(z[i]) + 1e-6 * scale * scale));
        }
        if (!(worst == worst)) return false;                  // NaN
        if (last) break;
        if (worst < 1e-18) last = true;                       // |step| < 1e-9 |z|: one more sweep
    }
    return true;
}

// Null vector of A(k) = -(k.k) I + k k^T + eps.  `second` selects the other vector
// of a 2-dimensional null space (degenerate / isotropic eps).
__device__ __forceinline__ void null_vector(const cplx eps[9], const cplx k[3], bool second,
                                            cplx e[3]) {
    const cplx kk = cdot(k, k);
    cplx a[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) a[3 * i + j] = eps[3 * i + j] + k[i] * k[j] - (i == j ? kk : C(0.0));
    cplx c01[3], c02[3], c12[3];
    ccross(&a[0], &a[3], c01);
    ccross(&a[0], &a[6], c02);
    ccross(&a[3], &a[6], c12);
    const double n01 = herm2(c01), n02 = herm2(c02), n12 = herm2(c12);
    double scale = 0.0;
    for (int i = 0; i < 9; ++i) scale += abs2(a[i]);
    const double best = fmax(n01, fmax(n02, n12));
    const cplx *pick = (n01 >= n02 && n01 >= n12) ? c01 : (n02 >= n12 ? c02 : c12);
    if (best > 1e-18 * scale * scale) {
        const double inv = rsqrt(best);
        for (int i = 0; i < 3; ++i) e[i] = inv * pick[i];
        return;
    }
    // rank <= 1: null space = bilinear complement of the dominant row r
    const double r0 = herm2(&a[0]), r1 = herm2(&a[3]), r2 = herm2(&a[6]);
    const cplx *r = (r0 >= r1 && r0 >= r2) ? &a[0] : (r1 >= r2 ? &a[3] : &a[6]);
    const double ax = abs2(r[0]), ay = abs2(r[1]), az = abs2(r[2]);
    cplx axis[3] = {C(0.0), C(0.0), C(0.0)};
    if (ax <= ay && ax <= az) axis[0] = C(1.0); else if (ay <= az) axis[1] = C(1.0); else axis[2] = C(1.0);
    cplx e1[3], e2[3];
    ccross(r, axis, e1);
    ccross(r, e1, e2);
    const cplx *out = second ? e2 : e1;
    const double inv = rsqrt(herm2(out));
    for (int i = 0; i < 3; ++i) e[i] = inv * out[i];
}

struct CRay {
    double x[3];
    cplx k[3], e[3];
    bool alive;
};

// Anisotropic deflection in the shape frame.  kl: incoming k (shape frame), nrm: unit
// normal.  Produces the two selected modes (ka, ea), (kb, eb).
__device__ __forceinline__ void aniso_modes(const EpsInv &ei, const cplx kl[3], const double nrm[3],
                                            bool mirror, cplx ka[3], cplx ea[3], cplx kb[3],
                                            cplx eb[3]) {
    const cplx kn = cdotr(kl, nrm);
    cplx p[3], nc[3];
    for (int i = 0; i < 3; ++i) { p[i] = kl[i] - nrm[i] * kn; nc[i] = C(nrm[i]); }
    bool ok = finite3(nrm) && cfinite(p[0]) && cfinite(p[1]) && cfinite(p[2]);
    cplx xi[4];
    if (ok) {
        const cplx s0 = cdot(p, p), s1 = 2.0 * cdotr(p, nrm), s2 = C(dot3(nrm, nrm));
        const cplx q0 = quad_form(ei.eps, p, p);
        const cplx q1 = quad_form(ei.eps, p, nc) + quad_form(ei.eps, nc, p);
        const cplx q2 = quad_form(ei.eps, nc, nc);
        const cplx a0 = quad_form(ei.adj, p, p);
        const cplx a1 = quad_form(ei.adj, p, nc) + quad_form(ei.adj, nc, p);
        const cplx a2 = quad_form(ei.adj, nc, nc);
        cplx c[5];
        c[4] = s2 * q2;
        c[3] = s1 * q2 + s2 * q1;
        c[2] = s0 * q2 + s1 * q1 + s2 * q0 - ei.c2 * s2 + a2;
        c[1] = s0 * q1 + s1 * q0 - ei.c2 * s1 + a1;
        c[0] = s0 * q0 - ei.c2 * s0 + a0 + ei.det;
        // starts: the roots of the isotropic medium with the mean permittivity,
        // xi = +-sqrt(tr(eps)/3 - p.p), split by a few per cent off the real axis
        const cplx third = {1.0 / 3.0, 0.0};
        const cplx xi2 = third * (ei.eps[0] + ei.eps[4] + ei.eps[8]) - s0;
        cplx r = csqrt_(xi2);
        double scale = sqrt(abs2(r));
        if (!(scale > 1e-3)) { scale = 1.0; r = C(1.0); }
        const cplx z0[4] = {r * cplx{1.03, 0.02}, r * cplx{0.97, -0.02},
                            r * cplx{-1.03, 0.02}, r * cplx{-0.97, -0.02}};
        ok = quartic_roots(c, z0, scale, xi);
    }
    if (!ok) {
        const cplx q = {qnan(), qnan()};
        for (int i = 0; i < 3; ++i) { ka[i] = kb[i] = ea[i] = eb[i] = q; }
        return;
    }
    // modes, Poynting sort key S.n with the reference's eigenvector normalisation
    // (unit 6-vector (xi E, E) -> |E|^2 = 1/(1 + |xi|^2))
    cplx km[4][3], em[4][3];
    double key[4], wgt[4];
    for (int m = 0; m < 4; ++m) {
        for (int i = 0; i < 3; ++i) km[m][i] = p[i] + xi[m] * nc[i];
        bool second = false;
        for (int j = 0; j < m; ++j)
            if (abs2(xi[m] - xi[j]) < 1e-14 * (1.0 + abs2(xi[m]))) second = !second;
        null_vector(ei.eps, km[m], second, em[m]);
        wgt[m] = 1.0 / (1.0 + abs2(xi[m]));
        double s[3];
        poynting_vec(km[m], em[m], s);
        key[m] = wgt[m] * dot3(s, nrm);
    }
    int ord[4] = {0, 1, 2, 3};
    for (int i = 1; i < 4; ++i) {                       // stable insertion sort, ascending
        const int v = ord[i];
        int j = i - 1;
        while (j >= 0 && key[ord[j]] > key[v]) { ord[j + 1] = ord[j]; --j; }
        ord[j + 1] = v;
    }
    const int ia = mirror ? ord[0] : ord[2];
    const int ib = mirror ? ord[1] : ord[3];
    const double sa = (mirror ? -1.0 : 1.0) * sqrt(wgt[ia]);
    const double sb = (mirror ? -1.0 : 1.0) * sqrt(wgt[ib]);
    const double sgn = mirror ? -1.0 : 1.0;
    for (int i = 0; i < 3; ++i) {
        ka[i] = sgn * km[ia][i]; kb[i] = sgn * km[ib][i];
        ea[i] = sa * em[ia][i]; eb[i] = sb * em[ib][i];
    }
}

__device__ __forceinline__ void cstore(double *base, int64_t idx, cplx v) {
    __stcs(reinterpret_cast<double2 *>(base) + idx, make_double2(v.re, v.im));
}
__device__ __forceinline__ cplx cload(const double *base, int64_t idx) {
    const double2 v = __ldcs(reinterpret_cast<const double2 *>(base) + idx);
    return {v.x, v.y};
}

__global__ void __launch_bounds__(128)
trace_complex_kernel(const __grid_constant__ LaunchParams P) {
    const int64_t n = P.n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        CRay r;
        const int64_t ix = (P.n_x == n) ? i : i % P.n_x;
        for (int c = 0; c < 3; ++c) {
            r.x[c] = P.x[c * P.ld_in + ix];
            r.k[c] = cload(P.k, c * P.ld_in + i);
            r.e[c] = P.e ? cload(P.e, c * P.ld_in + i) : C(c == 1 ? 1.0 : 0.0);
        }
        r.alive = P.alive ? (P.alive[ix] & PYR_RAY_ALIVE) != 0 : true;

        for (int s = 0; s < P.n_steps; ++s) {
            const DStep &st = P.steps[s];
            const DAux *aux = st.aux >= 0 ? &P.aux[st.aux] : nullptr;
            const bool ok = r.alive;

            // direction of energy transport: always the Poynting vector here
            double d[3];
            {
                double sv[3];
                poynting_vec(r.k, r.e, sv);
                const double inv = rsqrt(dot3(sv, sv));
                d[0] = sv[0] * inv; d[1] = sv[1] * inv; d[2] = sv[2] * inv;
            }
            double r0[3], dl[3];
            g2l_point(st.frame, r.x, r0);
            rot_t(st.frame.r, d, dl);
            double t;
            bool hit_ok = true;
            if (st.bits & kNoIntersect) t = 0.0;         // stand-alone refract / reflect: x is the hit point
            else if (st.shape_kind == PYR_SHAPE_CONIC) t = conic_t(st.curv, st.cc, r0, dl, hit_ok);
            else { double gfx, gfy; bool gok; t = explicit_t<false>(st.shape_kind, *aux, st.curv, st.cc, r0, dl, ok, gfx, gfy, gok); }
            const double h[3] = {fma(dl[0], t, r0[0]), fma(dl[1], t, r0[1]), fma(dl[2], t, r0[2])};
            double hit_g[3];
            l2g_point(st.frame, h, hit_g);

            bool ap_ok = true;
            if (st.aperture_kind != PYR_AP_BASE) {
                double ax = h[0], ay = h[1];
                if (!(st.bits & kApSameFrame)) {
                    double a[3];
                    g2l_point(aux->aperture_frame, hit_g, a);
                    ax = a[0]; ay = a[1];
                }
                if (st.aperture_kind == PYR_AP_CIRCULAR) {
                    const double rr = fma(ax, ax, ay * ay);
                    ap_ok = (rr >= st.ap0) && (rr <= st.ap1);
                } else {
                    ap_ok = (ax >= -st.ap0) && (ax <= st.ap0) && (ay >= -st.ap1) && (ay <= st.ap1);
                }
            }
            const bool hit = ok && hit_ok && ap_ok;

            double nrm[3];
            if (st.shape_kind == PYR_SHAPE_CONIC)
                conic_normal(st.curv, st.cc, (st.bits & kSphere) != 0, h[0], h[1], nrm);
            else
                explicit_normal<false>(st.shape_kind, *aux, st.curv, st.cc, h[0], h[1], nrm);

            cplx kl[3];
            crot_t(st.frame.r, r.k, kl);
            const bool mirror = st.interaction == PYR_REFLECT;
            cplx k2a[3], e2a[3], k2b[3], e2b[3];
            bool alive;
            const bool no_deflect = (st.bits & kNoDeflect) != 0;   // stand-alone propagate
            const bool aniso = st.after_kind == PYR_MEDIUM_ANISO && !no_deflect;
            if (no_deflect) {
                for (int c = 0; c < 3; ++c) { k2a[c] = kl[c]; }
                alive = hit;                 // k, E unchanged (material_anisotropic.py:58-68)
            } else if (aniso) {
                EpsInv ei;
                eps_invariants(aux->after.eps, ei);      // eps already in the shape frame
                aniso_modes(ei, kl, nrm, mirror, k2a, e2a, k2b, e2b);
                alive = ok;          // no validity filter (material_anisotropic.py:87-100)
            } else {
                // isotropic deflection with complex k (material_isotropic.py:163-236)
                const cplx kn = cdotr(kl, nrm);
                cplx kin[3];
                for (int c = 0; c < 3; ++c) kin[c] = kl[c] - nrm[c] * kn;
                const cplx square = C(st.n2sq[0]) - cdot(kin, kin);
                // numpy orders complex numbers lexicographically: (re, im) > (0, 0)
                const bool refr_ok = (square.re > 0.0 || (square.re == 0.0 && square.im > 0.0)) &&
                                     finite3(nrm);
                const cplx xi = csqrt_(square);
                for (int c = 0; c < 3; ++c) k2a[c] = (mirror ? -kin[c] : kin[c]) + nrm[c] * xi;
                alive = hit && refr_ok;
                // E: project the previous field (shape frame) onto the plane k2.E = 0
                cplx el[3];
                crot_t(st.frame.r, r.e, el);
                const cplx kk = cdot(k2a, k2a);
                cplx cc = cdot(el, k2a) / kk;
                cplx tv[3] = {el[0] - cc * k2a[0], el[1] - cc * k2a[1], el[2] - cc * k2a[2]};
                double tt = herm2(tv);
                if (!(tt > 1e-24 * herm2(el))) {
                    const double ax = abs2(k2a[0]), ay = abs2(k2a[1]), az = abs2(k2a[2]);
                    cplx a[3] = {C(0.0), C(0.0), C(0.0)};
                    if (ax <= ay && ax <= az) a[0] = C(1.0); else if (ay <= az) a[1] = C(1.0); else a[2] = C(1.0);
                    cc = cdot(a, k2a) / kk;
                    for (int c = 0; c < 3; ++c) tv[c] = a[c] - cc * k2a[c];
                    tt = herm2(tv);
                }
                const double inv = rsqrt(tt);
                for (int c = 0; c < 3; ++c) e2a[c] = inv * tv[c];
            }

            // ---- back to the global frame, record ----
            cplx kga[3], ega[3], kgb[3], egb[3];
            crot(st.frame.r, k2a, kga);
            crot(st.frame.r, e2a, ega);
            if (no_deflect) { for (int c = 0; c < 3; ++c) { kga[c] = r.k[c]; ega[c] = r.e[c]; } }
            const bool split = aniso && (st.bits & kSplit);
            if (aniso) { crot(st.frame.r, k2b, kgb); crot(st.frame.r, e2b, egb); }
            const cplx qn = {qnan(), qnan()};
            if (!alive) {
                for (int c = 0; c < 3; ++c) { kga[c] = ega[c] = kgb[c] = egb[c] = qn; }
            }
            const int64_t ld = st.ld_out;
            if (st.out_x)
                for (int c = 0; c < 3; ++c) __stcs(st.out_x + c * ld + i, ok ? hit_g[c] : qnan());
            if (st.out_flags)
                st.out_flags[i] = (uint8_t)((hit ? PYR_RAY_HIT : 0u) | (alive ? PYR_RAY_ALIVE : 0u));
            if (split) {
                const int64_t ld2 = st.ld_out2;
                for (int c = 0; c < 3; ++c) {
                    if (st.out_k) { cstore(st.out_k, c * ld2 + i, kga[c]); cstore(st.out_k, c * ld2 + n + i, kgb[c]); }
                    if (st.out_e) { cstore(st.out_e, c * ld2 + i, ega[c]); cstore(st.out_e, c * ld2 + n + i, egb[c]); }
                }
            } else {
                for (int c = 0; c < 3; ++c) {
                    if (st.out_k) cstore(st.out_k, c * ld + i, kga[c]);
                    if (st.out_e) cstore(st.out_e, c * ld + i, ega[c]);
                }
            }
            // continue with mode a (a split step is always the last of the launch)
            for (int c = 0; c < 3; ++c) {
                r.x[c] = alive ? hit_g[c] : qnan();
                r.k[c] = kga[c];
                r.e[c] = ega[c];
            }
            r.alive = alive;
        }
    }
}

int pack_steps(const PyrStep *steps, int32_t n_steps, const PyrRaysIn *rays, int64_t n_rays,
               uint32_t flags, LaunchParams &P, bool &general, bool &any_aniso);   // pyr_trace.cu
int sm_count();

int trace_complex(const PyrStep *steps, int32_t n_steps, const PyrRaysIn *rays, int64_t n_rays,
                  uint32_t flags, cudaStream_t stream) {
    static thread_local LaunchParams P;
    bool general = false, any_aniso = false;
    int rc = pack_steps(steps, n_steps, rays, n_rays, flags, P, general, any_aniso);
    if (rc != PYR_OK) return rc;
    if (!rays->e) return PYR_E_BADARG;            // E defines the ray direction in crystals
    if (rays->n_waves > 1) return PYR_E_UNSUPPORTED;
    for (int s = 0; s < n_steps; ++s) {
        if (steps[s].before.kind == PYR_MEDIUM_ISO_GRIN || steps[s].after.kind == PYR_MEDIUM_ISO_GRIN)
            return PYR_E_UNSUPPORTED;
    }
    if (n_rays == 0) return PYR_OK;
    const int threads = 128;
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trace_complex_kernel, threads, 0);
    if (e != cudaSuccess) return (int)e;
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (n_rays + threads - 1) / threads;
    const int64_t cap = (int64_t)sm_count() * per_sm;
    if (grid > cap) grid = cap;
    trace_complex_kernel<<<(unsigned)grid, threads, 0, stream>>>(P);
    e = cudaGetLastError();
    return e == cudaSuccess ? PYR_OK : (int)e;
}

}  // namespace pyr
