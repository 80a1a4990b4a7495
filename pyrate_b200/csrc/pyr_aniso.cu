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
// :174-327).  For uniaxial (and isotropic) real tensors eps = eo 1 + (ee - eo) a a^T --
// calcite, quartz, the BASELINE crystals -- the quartic factorises into the ordinary
// sphere k.k = eo and the extraordinary ellipsoid eo k.k + (ee - eo)(k.a)^2 = eo ee: two
// quadratics in xi, closed form (pinned on the reference's eigenvalues in the oracle,
// oracle/pyrate_np.py uniaxial_xi_roots).  General (biaxial / complex) tensors keep the
// Aberth-Ehrlich iteration, in an own kernel instantiation.  E is the null vector of the
// 3x3 propagator at each root (cross product of its two most independent rows).
//
// ONE launch per complex stretch: a birefringent interface doubles the rays
// (material_anisotropic.py:87-100); thread i walks the whole mode tree of its ray depth
// first.  The record of a doubling step IS the stack: mode b's (x, k, E) is what the
// step writes to column w + c of its record anyway, and the same thread reads it back
// when it returns to that branch -- leaf t > 0 resumes at the split whose bit is the
// lowest set bit of t.  All threads follow the same control flow.
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
    const double d = fast_rcp(abs2(b));                // b == 0: NaN (callers test cfinite)
    return {(a.re * b.re + a.im * b.im) * d, (a.im * b.re - a.re * b.im) * d};
}
// Square roots / reciprocals in this file are the branch-free MUFU + Newton forms of
// pyr_device.cuh (~1 ulp): the C library's sqrt / hypot / rsqrt / division each carry a range
// check and an out-of-line slow path -- 14 calls per warp and 7 % of the instructions of the
// C4 trace (profiles/r02_c4.md) -- and the arguments here are O(1) optical quantities.
__device__ __forceinline__ cplx csqrt_(cplx z) {       // principal branch, like numpy
    const double m2 = abs2(z);
    const double m = fast_sqrt(m2);                    // z == 0: NaN, replaced below
    const double a = fast_sqrt(0.5 * (m + fabs(z.re)));
    const double b = fast_div(0.5 * z.im, a);
    cplx r = {a, b};
    if (!(z.re >= 0.0)) r = {fabs(b), copysign(a, z.im)};
    if (m2 == 0.0) r = {0.0, 0.0};
    return r;
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
            worst = fmax(worst, fast_div(abs2(step), abs2(z[i]) + 1e-6 * scale * scale));
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
        const double inv = fast_rsqrt(best);
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
    const double inv = fast_rsqrt(herm2(out));
    for (int i = 0; i < 3; ++i) e[i] = inv * out[i];
}

// general null vector behind a call: the uniaxial kernel needs it only for degenerate
// modes (propagation along the optic axis, isotropic tensors) and must not pay its
// registers on the fast path.  Arguments and result travel BY VALUE (registers): a pointer
// to the caller's k / e arrays would pin those arrays -- the outputs of every deflection --
// in local memory for the whole kernel (it did: ~300 local stores per ray, profiles/r02_c4.md)
struct cvec3 {
    cplx v[3];
};
__device__ __noinline__ cvec3 null_vector_call(const double *eps18, cplx k0, cplx k1, cplx k2, bool second) {
    cplx eps[9];
    for (int i = 0; i < 9; ++i) eps[i] = {eps18[2 * i], eps18[2 * i + 1]};
    const cplx k[3] = {k0, k1, k2};
    cvec3 r;
    null_vector(eps, k, second, r.v);
    return r;
}

// One mode k = p + xi n: unit null vector e of the propagator, weight of the reference's
// eigenvector normalisation (unit 6-vector (xi E, E) -> |E|^2 = 1 / (1 + |xi|^2)) and the
// Poynting sort key S.n in that normalisation (material.py:122-153, :214-223).
//   UNI: uniaxial tensor eps = eo 1 + de a a^T.  Ordinary mode: E = k x a; extraordinary
//   mode: E = eo a - (k.a) k  (A(k) E = 0 with A = -(k.k) 1 + k k^T + eps, given the
//   respective dispersion relation).  Both vanish for k parallel to a: general null vector.
template <bool UNI>
__device__ __forceinline__ double mode_key(const DAux &ax, const cplx p[3], const double nrm[3],
                                           cplx xi, bool extraordinary, bool second, cplx k[3],
                                           cplx e[3], double &wgt) {
    for (int i = 0; i < 3; ++i) k[i] = p[i] + nrm[i] * xi;
    bool closed = false;
    if (UNI) {
        const double a[3] = {ax.axis[0], ax.axis[1], ax.axis[2]};
        cplx ko[3];                                           // k x a
        ko[0] = {k[1].re * a[2] - k[2].re * a[1], k[1].im * a[2] - k[2].im * a[1]};
        ko[1] = {k[2].re * a[0] - k[0].re * a[2], k[2].im * a[0] - k[0].im * a[2]};
        ko[2] = {k[0].re * a[1] - k[1].re * a[0], k[0].im * a[1] - k[1].im * a[0]};
        const double cross2 = herm2(ko);
        closed = ax.eps_e != ax.eps_o && cross2 > 1e-8 * herm2(k);
        if (closed) {
            if (extraordinary) {
                const cplx ka = cdotr(k, a);
                for (int i = 0; i < 3; ++i) e[i] = C(ax.eps_o * a[i]) - ka * k[i];
                const double inv = fast_rsqrt(herm2(e));
                for (int i = 0; i < 3; ++i) e[i] = inv * e[i];
            } else {
                const double inv = fast_rsqrt(cross2);
                for (int i = 0; i < 3; ++i) e[i] = inv * ko[i];
            }
        }
    }
    if (!closed) {
        if (UNI) {
            const cvec3 nv = null_vector_call(ax.after.eps, k[0], k[1], k[2], second);
            for (int i = 0; i < 3; ++i) e[i] = nv.v[i];
        } else {
            cplx eps[9];
            for (int i = 0; i < 9; ++i) eps[i] = {ax.after.eps[2 * i], ax.after.eps[2 * i + 1]};
            null_vector(eps, k, second, e);
        }
    }
    wgt = fast_rcp(1.0 + abs2(xi));
    double s[3];
    poynting_vec(k, e, s);
    return wgt * dot3(s, nrm);
}

// Sort key of a closed-form uniaxial mode WITHOUT forming its field: with E_o = k x a
// (E.k = 0) the Poynting vector is |E|^2 Re k, so the key is w Re(k).n; with
// E_e = eo a - (k.a) k one has E.k = (k.a)(eo - k.k) and E.n = eo (a.n) - (k.a)(k.n), so
// S.n / |E|^2 = Re(k.n) - Re((E.k) conj(E.n)) / |E|^2.  `closed` = false where the closed
// forms degenerate (k parallel to a, isotropic tensor): the caller then uses mode_key.
__device__ __forceinline__ double uni_key(const DAux &ax, const cplx p[3], const double nrm[3], cplx xi,
                                          bool extraordinary, bool &closed) {
    cplx k[3];
    for (int i = 0; i < 3; ++i) k[i] = p[i] + nrm[i] * xi;
    const double a[3] = {ax.axis[0], ax.axis[1], ax.axis[2]};
    const double c0r = k[1].re * a[2] - k[2].re * a[1], c0i = k[1].im * a[2] - k[2].im * a[1];
    const double c1r = k[2].re * a[0] - k[0].re * a[2], c1i = k[2].im * a[0] - k[0].im * a[2];
    const double c2r = k[0].re * a[1] - k[1].re * a[0], c2i = k[0].im * a[1] - k[1].im * a[0];
    const double cross2 = c0r * c0r + c0i * c0i + c1r * c1r + c1i * c1i + c2r * c2r + c2i * c2i;
    closed = closed && ax.eps_e != ax.eps_o && cross2 > 1e-8 * herm2(k);
    const double wgt = fast_rcp(1.0 + abs2(xi));
    const cplx kn = cdotr(k, nrm);
    if (!extraordinary) return wgt * kn.re;
    const cplx ka = cdotr(k, a);
    const cplx kk = cdot(k, k);
    cplx e[3];
    for (int i = 0; i < 3; ++i) e[i] = C(ax.eps_o * a[i]) - ka * k[i];
    const cplx ek = ka * (C(ax.eps_o) - kk);
    const cplx en = C(ax.eps_o * dot3(a, nrm)) - ka * kn;
    return wgt * (kn.re - fast_div((ek * conj(en)).re, herm2(e)));
}

// Anisotropic deflection in the shape frame.  kl: incoming k (shape frame), nrm: unit
// normal.  Produces the two selected modes (ka, ea), (kb, eb).  No dynamically indexed
// local arrays (no stack frame): the four sort keys are computed first, the two selected
// modes are then evaluated again.
template <bool GENERAL_EPS>
__device__ __forceinline__ void aniso_modes(const DAux &ax, const cplx kl[3], const double nrm[3],
                                            bool mirror, cplx ka[3], cplx ea[3], cplx kb[3],
                                            cplx eb[3]) {
    const cplx kn = cdotr(kl, nrm);
    cplx p[3];
    for (int i = 0; i < 3; ++i) p[i] = kl[i] - nrm[i] * kn;
    bool ok = finite3(nrm) && cfinite(p[0]) && cfinite(p[1]) && cfinite(p[2]);
    constexpr bool UNI = !GENERAL_EPS;
    cplx xi0, xi1, xi2, xi3;
    if (UNI) {
        // closed form: ordinary +-, extraordinary +- (p.n = 0 by construction)
        const double eo = ax.eps_o, de = ax.eps_e - ax.eps_o;
        const double a[3] = {ax.axis[0], ax.axis[1], ax.axis[2]};
        const cplx kk = cdot(p, p);
        const cplx pa = cdotr(p, a);
        const double na = dot3(nrm, a);
        const cplx xo = csqrt_(C(eo) - kk);
        const double qa = fma(de * na, na, eo);
        const cplx qb = (de * na) * pa;                          // half the linear coefficient
        const cplx qc = eo * kk + de * (pa * pa) - C(eo * ax.eps_e);
        const cplx disc = csqrt_(qb * qb - qa * qc);
        const double iqa = fast_rcp(qa);
        xi0 = xo; xi1 = -xo;
        xi2 = iqa * (disc - qb); xi3 = iqa * (-disc - qb);
        ok = ok && cfinite(xi0) && cfinite(xi2) && cfinite(xi3);
    } else {
        cplx nc[3];
        for (int i = 0; i < 3; ++i) nc[i] = C(nrm[i]);
        cplx xi[4];
        if (ok) {
            cplx eps[9];
            for (int i = 0; i < 9; ++i) eps[i] = {ax.after.eps[2 * i], ax.after.eps[2 * i + 1]};
            // invariants of eps: adjugate, sum of the principal 2x2 minors, determinant
            cplx adj[9];
            const cplx *mm = eps;
            adj[0] = mm[4] * mm[8] - mm[5] * mm[7];
            adj[1] = mm[2] * mm[7] - mm[1] * mm[8];
            adj[2] = mm[1] * mm[5] - mm[2] * mm[4];
            adj[3] = mm[5] * mm[6] - mm[3] * mm[8];
            adj[4] = mm[0] * mm[8] - mm[2] * mm[6];
            adj[5] = mm[2] * mm[3] - mm[0] * mm[5];
            adj[6] = mm[3] * mm[7] - mm[4] * mm[6];
            adj[7] = mm[1] * mm[6] - mm[0] * mm[7];
            adj[8] = mm[0] * mm[4] - mm[1] * mm[3];
            const cplx det = mm[0] * adj[0] + mm[1] * adj[3] + mm[2] * adj[6];
            const cplx c2 = adj[0] + adj[4] + adj[8];
            const cplx s0 = cdot(p, p), s1 = 2.0 * cdotr(p, nrm), s2 = C(dot3(nrm, nrm));
            const cplx q0 = quad_form(eps, p, p);
            const cplx q1 = quad_form(eps, p, nc) + quad_form(eps, nc, p);
            const cplx q2 = quad_form(eps, nc, nc);
            const cplx a0 = quad_form(adj, p, p);
            const cplx a1 = quad_form(adj, p, nc) + quad_form(adj, nc, p);
            const cplx a2 = quad_form(adj, nc, nc);
            cplx c[5];
            c[4] = s2 * q2;
            c[3] = s1 * q2 + s2 * q1;
            c[2] = s0 * q2 + s1 * q1 + s2 * q0 - c2 * s2 + a2;
            c[1] = s0 * q1 + s1 * q0 - c2 * s1 + a1;
            c[0] = s0 * q0 - c2 * s0 + a0 + det;
            // starts: the roots of the isotropic medium with the mean permittivity,
            // xi = +-sqrt(tr(eps)/3 - p.p), split by a few per cent off the real axis
            const cplx third = {1.0 / 3.0, 0.0};
            const cplx xi2s = third * (eps[0] + eps[4] + eps[8]) - s0;
            cplx r = csqrt_(xi2s);
            double scale = fast_sqrt(abs2(r));          // r == 0: NaN, replaced below
            if (!(scale > 1e-3)) { scale = 1.0; r = C(1.0); }
            const cplx z0[4] = {r * cplx{1.03, 0.02}, r * cplx{0.97, -0.02},
                                r * cplx{-1.03, 0.02}, r * cplx{-0.97, -0.02}};
            ok = quartic_roots(c, z0, scale, xi);
        }
        xi0 = xi[0]; xi1 = xi[1]; xi2 = xi[2]; xi3 = xi[3];
    }
    if (!ok) {
        const cplx q = {qnan(), qnan()};
        for (int i = 0; i < 3; ++i) { ka[i] = kb[i] = ea[i] = eb[i] = q; }
        return;
    }
    // degenerate pairs (isotropic tensors, propagation along the optic axis): the second
    // root of a pair takes the other vector of the two-dimensional null space
    auto same = [](cplx a, cplx b) { return abs2(a - b) < 1e-14 * (1.0 + abs2(a)); };
    const bool sec1 = same(xi1, xi0);
    const bool sec2 = same(xi2, xi0) != same(xi2, xi1);
    const bool sec3 = (same(xi3, xi0) != same(xi3, xi1)) != same(xi3, xi2);
    // pass 1: sort keys S.n with the reference's eigenvector normalisation.  A rolled
    // loop (one copy of the mode evaluation in the instruction stream, not four: the kernel
    // was stalling on instruction fetch, profiles/r02_c4.md)
    double key0 = 0.0, key1 = 0.0, key2 = 0.0, key3 = 0.0, w;
    bool closed = UNI;
    if (UNI) {
        // uniaxial crystal: the four keys in closed form (no fields)
        key0 = uni_key(ax, p, nrm, xi0, false, closed);
        key1 = uni_key(ax, p, nrm, xi1, false, closed);
        key2 = uni_key(ax, p, nrm, xi2, true, closed);
        key3 = uni_key(ax, p, nrm, xi3, true, closed);
    }
    if (!closed)
#pragma unroll 1
    for (int mi = 0; mi < 4; ++mi) {
        const cplx xim = mi == 0 ? xi0 : (mi == 1 ? xi1 : (mi == 2 ? xi2 : xi3));
        const bool secm = mi == 1 ? sec1 : (mi == 2 ? sec2 : (mi == 3 ? sec3 : false));
        const double key = mode_key<UNI>(ax, p, nrm, xim, mi >= 2, secm, ka, ea, w);
        if (mi == 0) key0 = key; else if (mi == 1) key1 = key; else if (mi == 2) key2 = key; else key3 = key;
    }
    // stable ascending rank of every key (what the reference's argsort gives)
    const int r0 = (key1 < key0) + (key2 < key0) + (key3 < key0);
    const int r1 = (key0 <= key1) + (key2 < key1) + (key3 < key1);
    const int r2 = (key0 <= key2) + (key1 <= key2) + (key3 < key2);
    const int r3 = (key0 <= key3) + (key1 <= key3) + (key2 <= key3);
    const int ra = mirror ? 0 : 2, rb = mirror ? 1 : 3;      // material_anisotropic.py:89-91 / :133-134
    cplx xa = xi2, xb = xi3;
    bool seca = sec2, secb = sec3, exa = true, exb = true;
    if (r0 == ra) { xa = xi0; seca = false; exa = false; } else if (r1 == ra) { xa = xi1; seca = sec1; exa = false; }
    else if (r3 == ra) { xa = xi3; seca = sec3; }
    if (r0 == rb) { xb = xi0; secb = false; exb = false; } else if (r1 == rb) { xb = xi1; secb = sec1; exb = false; }
    else if (r2 == rb) { xb = xi2; secb = sec2; }
    // pass 2: the two selected modes (same rolled evaluation; mode b first, it lands in kb / eb)
    double wa = 0.0, wb = 0.0;
#pragma unroll 1
    for (int mi = 0; mi < 2; ++mi) {
        const bool second_mode = mi == 0;
        mode_key<UNI>(ax, p, nrm, second_mode ? xb : xa, second_mode ? exb : exa,
                      second_mode ? secb : seca, ka, ea, w);
        if (second_mode) {
            for (int i = 0; i < 3; ++i) { kb[i] = ka[i]; eb[i] = ea[i]; }
            wb = w;
        } else {
            wa = w;
        }
    }
    const double sgn = mirror ? -1.0 : 1.0;
    const double sa = sgn * fast_sqrt(wa), sb = sgn * fast_sqrt(wb);
    for (int i = 0; i < 3; ++i) {
        ka[i] = sgn * ka[i]; kb[i] = sgn * kb[i];
        ea[i] = sa * ea[i]; eb[i] = sb * eb[i];
    }
}

// ---------------------------------------------------------------------------
// Real-valued fast path.  In a transparent uniaxial crystal below every critical angle --
// BASELINE config 4, and what birefringent systems normally are -- k and E of every mode
// stay REAL: the complex arithmetic above then multiplies by zero imaginary parts in three
// of four products.  A warp whose rays all carry real (k, E) runs the functions below (the
// same formulas, real); a negative radicand (evanescent mode) or a degenerate mode in any
// lane sends the whole warp through the complex code for that step.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void cross3(const double a[3], const double b[3], double c[3]) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

// ray.py:140-152 for real k, E
__device__ __forceinline__ void poynting_vec_r(const double k[3], const double e[3], double s[3]) {
    const double ee = dot3(e, e), ek = dot3(e, k);
    for (int i = 0; i < 3; ++i) s[i] = ee * k[i] - ek * e[i];
}

// uni_key for a real root: `closed` as there
__device__ __forceinline__ double uni_key_r(const DAux &ax, const double p[3], const double nrm[3], double xi,
                                            bool extraordinary, bool &closed) {
    double k[3], ko[3];
    for (int i = 0; i < 3; ++i) k[i] = p[i] + nrm[i] * xi;
    const double a[3] = {ax.axis[0], ax.axis[1], ax.axis[2]};
    cross3(k, a, ko);
    const double kk = dot3(k, k);
    closed = closed && ax.eps_e != ax.eps_o && dot3(ko, ko) > 1e-8 * kk;
    const double wgt = fast_rcp(1.0 + xi * xi);
    const double kn = dot3(k, nrm);
    if (!extraordinary) return wgt * kn;
    const double ka = dot3(k, a);
    double e[3];
    for (int i = 0; i < 3; ++i) e[i] = ax.eps_o * a[i] - ka * k[i];
    const double ek = ka * (ax.eps_o - kk);
    const double en = ax.eps_o * dot3(a, nrm) - ka * kn;
    return wgt * (kn - fast_div(ek * en, dot3(e, e)));
}

// one selected mode (closed forms of mode_key<true>), field scaled like aniso_modes does
__device__ __forceinline__ void uni_mode_r(const DAux &ax, const double p[3], const double nrm[3], double xi,
                                           bool extraordinary, double sgn, double k[3], double e[3]) {
    for (int i = 0; i < 3; ++i) k[i] = p[i] + nrm[i] * xi;
    const double a[3] = {ax.axis[0], ax.axis[1], ax.axis[2]};
    if (extraordinary) {
        const double ka = dot3(k, a);
        for (int i = 0; i < 3; ++i) e[i] = ax.eps_o * a[i] - ka * k[i];
    } else {
        cross3(k, a, e);
    }
    // unit field times sqrt(weight), weight = 1 / (1 + xi^2)
    const double sc = sgn * fast_rsqrt(dot3(e, e) * (1.0 + xi * xi));
    for (int i = 0; i < 3; ++i) { k[i] = sgn * k[i]; e[i] = sc * e[i]; }
}

// aniso_modes<false> for real k: returns false where the complex code is needed (evanescent
// or degenerate mode); NaN outputs for non-finite input like there
__device__ __forceinline__ bool aniso_modes_r(const DAux &ax, const double kl[3], const double nrm[3],
                                              bool mirror, double ka[3], double ea[3], double kb[3],
                                              double eb[3]) {
    const double kn = dot3(kl, nrm);
    double p[3];
    for (int i = 0; i < 3; ++i) p[i] = kl[i] - nrm[i] * kn;
    const double eo = ax.eps_o, de = ax.eps_e - ax.eps_o;
    const double a[3] = {ax.axis[0], ax.axis[1], ax.axis[2]};
    const double kk = dot3(p, p), pa = dot3(p, a), na = dot3(nrm, a);
    const double r1 = eo - kk;                                   // xi_o^2
    const double qa = fma(de * na, na, eo);
    const double qb = (de * na) * pa;
    const double qc = eo * kk + de * (pa * pa) - eo * ax.eps_e;
    const double r2 = qb * qb - qa * qc;
    const bool finite = isfinite(r1) && isfinite(r2) && finite3(nrm);
    if (!finite) {
        for (int i = 0; i < 3; ++i) ka[i] = kb[i] = ea[i] = eb[i] = qnan();
        return true;
    }
    if (r1 < 0.0 || r2 < 0.0) return false;                      // evanescent: complex roots
    const double xo = r1 > 0.0 ? fast_sqrt(r1) : 0.0;
    const double disc = r2 > 0.0 ? fast_sqrt(r2) : 0.0;
    const double iqa = fast_rcp(qa);
    const double xi0 = xo, xi1 = -xo, xi2 = iqa * (disc - qb), xi3 = iqa * (-disc - qb);
    if (!(isfinite(xi2) && isfinite(xi3))) {
        for (int i = 0; i < 3; ++i) ka[i] = kb[i] = ea[i] = eb[i] = qnan();
        return true;
    }
    bool closed = true;
    const double key0 = uni_key_r(ax, p, nrm, xi0, false, closed);
    const double key1 = uni_key_r(ax, p, nrm, xi1, false, closed);
    const double key2 = uni_key_r(ax, p, nrm, xi2, true, closed);
    const double key3 = uni_key_r(ax, p, nrm, xi3, true, closed);
    if (!closed) return false;                                   // degenerate: general null vectors
    // stable ascending rank of every key (what the reference's argsort gives)
    const int r0 = (key1 < key0) + (key2 < key0) + (key3 < key0);
    const int r1k = (key0 <= key1) + (key2 < key1) + (key3 < key1);
    const int r2k = (key0 <= key2) + (key1 <= key2) + (key3 < key2);
    const int r3 = (key0 <= key3) + (key1 <= key3) + (key2 <= key3);
    const int ra = mirror ? 0 : 2, rb = mirror ? 1 : 3;      // material_anisotropic.py:89-91 / :133-134
    double xa = xi2, xb = xi3;
    bool exa = true, exb = true;
    if (r0 == ra) { xa = xi0; exa = false; } else if (r1k == ra) { xa = xi1; exa = false; }
    else if (r3 == ra) { xa = xi3; }
    if (r0 == rb) { xb = xi0; exb = false; } else if (r1k == rb) { xb = xi1; exb = false; }
    else if (r2k == rb) { xb = xi2; }
    const double sgn = mirror ? -1.0 : 1.0;
    uni_mode_r(ax, p, nrm, xa, exa, sgn, ka, ea);
    uni_mode_r(ax, p, nrm, xb, exb, sgn, kb, eb);
    return true;
}

__device__ __forceinline__ void cstore(double *base, int64_t idx, cplx v) {
    __stcs(reinterpret_cast<double2 *>(base) + idx, make_double2(v.re, v.im));
}
__device__ __forceinline__ cplx cload(const double *base, int64_t idx) {
    const double2 v = __ldcg(reinterpret_cast<const double2 *>(base) + idx);
    return {v.x, v.y};
}

// split bookkeeping of a launch: step indices of the doubling steps (at most kMaxSplits)
constexpr int kMaxSplits = 6;

// EXPLICIT: explicit shapes (asphere / XY polynomial / biconic Newton) present; crystals
// between conics -- the common case -- run an instantiation without that code
#ifndef PYR_C4_MINB
#define PYR_C4_MINB 2          // resident CTAs per SM of the uniaxial kernel (tools builds vary it):
                               // 2 (254 registers, no spills) 0.98 ms on C4, 3 (168 registers, 670 B
                               // of spills) 1.20 ms (profiles/r02_c4.md)
#endif
template <bool GENERAL_EPS, bool EXPLICIT>
__global__ void __launch_bounds__(128, GENERAL_EPS ? 2 : PYR_C4_MINB)
trace_complex_kernel(const __grid_constant__ LaunchParams P) {
    const int64_t n = P.n;
    // doubling steps of this launch, in order (uniform over the grid)
    int m = 0;
    int split_step[kMaxSplits];
#pragma unroll
    for (int q = 0; q < kMaxSplits; ++q) split_step[q] = 0;
    for (int s = 0; s < P.n_steps; ++s) {
        if ((P.steps[s].bits & kSplit) && P.steps[s].after_kind == PYR_MEDIUM_ANISO &&
            !(P.steps[s].bits & kNoDeflect)) {
#pragma unroll
            for (int q = 0; q < kMaxSplits; ++q)
                if (q == m) split_step[q] = s;
            ++m;
        }
    }
    const unsigned leaves = 1u << m;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        for (unsigned t = 0; t < leaves; ++t) {
            double x[3];
            cplx k[3], e[3];
            bool alive;
            int s_begin;
            int64_t col, w;                 // column of this branch and width of its level
            if (t == 0) {
                const int64_t ix = (P.n_x == n) ? i : i % P.n_x;
                for (int c = 0; c < 3; ++c) {
                    x[c] = P.x[c * P.ld_in + ix];
                    k[c] = cload(P.k, c * P.ld_in + i);
                    e[c] = P.e ? cload(P.e, c * P.ld_in + i) : C(c == 1 ? 1.0 : 0.0);
                }
                alive = P.alive ? (P.alive[ix] & PYR_RAY_ALIVE) != 0 : true;
                s_begin = 0; col = i; w = n;
            } else {
                // mode b of split j = m - 1 - ctz(t); splits q < j follow the bits of t
                const int j = m - 1 - (__ffs((int)t) - 1);
                col = i; w = n;
                int sj = 0;
#pragma unroll
                for (int q = 0; q < kMaxSplits; ++q) {
                    if (q < j) { if ((t >> (m - 1 - q)) & 1u) col += w; w *= 2; }
                    if (q == j) sj = split_step[q];
                }
                const DStep &st = P.steps[sj];
                for (int c = 0; c < 3; ++c) {
                    x[c] = __ldcg(st.out_x + c * st.ld_out + col);
                    k[c] = cload(st.out_k, c * st.ld_out2 + w + col);
                    e[c] = cload(st.out_e, c * st.ld_out2 + w + col);
                }
                alive = (__ldcg(st.out_flags + col) & PYR_RAY_ALIVE) != 0;
                if (!alive) x[0] = x[1] = x[2] = qnan();
                col += w; w *= 2;
                s_begin = sj + 1;
            }

            for (int s = s_begin; s < P.n_steps; ++s) {
                const DStep &st = P.steps[s];
                const DAux *aux = st.aux >= 0 ? &P.aux[st.aux] : nullptr;
                const bool ok = alive;

                // real-valued fast path of this step: every lane of the warp carries real k, E
                // (dead lanes hold NaN and follow along)
                const bool lane_real = !ok || (k[0].im == 0.0 && k[1].im == 0.0 && k[2].im == 0.0 &&
                                               e[0].im == 0.0 && e[1].im == 0.0 && e[2].im == 0.0);
                const bool warp_real = __all_sync(__activemask(), lane_real);
                // direction of energy transport: always the Poynting vector here
                double d[3];
                {
                    double sv[3];
                    if (warp_real) {
                        const double kr[3] = {k[0].re, k[1].re, k[2].re}, er[3] = {e[0].re, e[1].re, e[2].re};
                        poynting_vec_r(kr, er, sv);
                    } else {
                        poynting_vec(k, e, sv);
                    }
                    const double inv = fast_rsqrt(dot3(sv, sv));
                    d[0] = sv[0] * inv; d[1] = sv[1] * inv; d[2] = sv[2] * inv;
                }
                const bool ident = (st.bits & kRotIdentity) != 0;   // chain systems without tilts
                double r0[3], dl[3];
                if (ident) {
                    for (int c = 0; c < 3; ++c) { r0[c] = x[c] - st.frame.o[c]; dl[c] = d[c]; }
                } else {
                    g2l_point(st.frame, x, r0);
                    rot_t(st.frame.r, d, dl);
                }
                double tt;
                bool hit_ok = true;
                if (st.bits & kNoIntersect) tt = 0.0;        // stand-alone refract / reflect: x is the hit point
                else if (st.shape_kind == PYR_SHAPE_CONIC) tt = conic_t(st.curv, st.cc, r0, dl, hit_ok);
                else if (st.shape_kind == PYR_SHAPE_CYLINDER) tt = cylinder_t(st.curv, st.cc, r0, dl, hit_ok);
                else if (EXPLICIT) { double gfx, gfy; bool gok; tt = explicit_t<false>(st.shape_kind, *aux, st.curv, st.cc, r0, dl, ok, gfx, gfy, gok); }
                else tt = qnan();
                const double h[3] = {fma(dl[0], tt, r0[0]), fma(dl[1], tt, r0[1]), fma(dl[2], tt, r0[2])};
                double hit_g[3];
                if (ident) { for (int c = 0; c < 3; ++c) hit_g[c] = h[c] + st.frame.o[c]; }
                else l2g_point(st.frame, h, hit_g);

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
                else if (st.shape_kind == PYR_SHAPE_CYLINDER)
                    cylinder_normal(st.curv, st.cc, h[1], nrm);
                else if (EXPLICIT)
                    explicit_normal<false>(st.shape_kind, *aux, st.curv, st.cc, h[0], h[1], nrm);
                else
                    nrm[0] = nrm[1] = nrm[2] = qnan();

                cplx kl[3];
                if (ident) { for (int c = 0; c < 3; ++c) kl[c] = k[c]; }
                else crot_t(st.frame.r, k, kl);
                const bool mirror = st.interaction == PYR_REFLECT;
                cplx k2a[3], e2a[3], k2b[3], e2b[3];
                const bool no_deflect = (st.bits & kNoDeflect) != 0;   // stand-alone propagate
                const bool aniso = st.after_kind == PYR_MEDIUM_ANISO && !no_deflect;
                bool done_real = false;
                if (warp_real && !no_deflect) {
                    const double klr[3] = {kl[0].re, kl[1].re, kl[2].re};
                    if (aniso) {
                        if (!GENERAL_EPS) {
                            double ka_[3], ea_[3], kb_[3], eb_[3];
                            const bool fine = aniso_modes_r(*aux, klr, nrm, mirror, ka_, ea_, kb_, eb_);
                            if (__all_sync(__activemask(), fine)) {
                                for (int c = 0; c < 3; ++c) {
                                    k2a[c] = C(ka_[c]); e2a[c] = C(ea_[c]); k2b[c] = C(kb_[c]); e2b[c] = C(eb_[c]);
                                }
                                alive = ok;
                                done_real = true;
                            }
                        }
                    } else {
                        // isotropic deflection, real k (material_isotropic.py:163-236); total
                        // reflection invalidates the ray (its complex k is never used)
                        const double kn = dot3(klr, nrm);
                        double kin[3];
                        for (int c = 0; c < 3; ++c) kin[c] = klr[c] - nrm[c] * kn;
                        const double square = st.n2sq[0] - dot3(kin, kin);
                        const bool refr_ok = square > 0.0 && finite3(nrm);
                        const double xi = fast_sqrt(square);
                        double k2[3];
                        for (int c = 0; c < 3; ++c) k2[c] = (mirror ? -kin[c] : kin[c]) + nrm[c] * xi;
                        alive = hit && refr_ok;
                        double el[3];
                        if (ident) { for (int c = 0; c < 3; ++c) el[c] = e[c].re; }
                        else { const double er[3] = {e[0].re, e[1].re, e[2].re}; rot_t(st.frame.r, er, el); }
                        const double kk = dot3(k2, k2);
                        double cc = fast_div(dot3(el, k2), kk);
                        double tv[3] = {el[0] - cc * k2[0], el[1] - cc * k2[1], el[2] - cc * k2[2]};
                        double t2 = dot3(tv, tv);
                        if (!(t2 > 1e-24 * dot3(el, el))) {
                            const double ax = k2[0] * k2[0], ay = k2[1] * k2[1], az = k2[2] * k2[2];
                            double a[3] = {0.0, 0.0, 0.0};
                            if (ax <= ay && ax <= az) a[0] = 1.0; else if (ay <= az) a[1] = 1.0; else a[2] = 1.0;
                            cc = fast_div(dot3(a, k2), kk);
                            for (int c = 0; c < 3; ++c) tv[c] = a[c] - cc * k2[c];
                            t2 = dot3(tv, tv);
                        }
                        const double inv = fast_rsqrt(t2);
                        for (int c = 0; c < 3; ++c) { k2a[c] = C(k2[c]); e2a[c] = C(inv * tv[c]); }
                        done_real = true;
                    }
                }
                if (done_real) {
                    // (deflected in real arithmetic)
                } else if (no_deflect) {
                    for (int c = 0; c < 3; ++c) { k2a[c] = kl[c]; }
                    alive = hit;                 // k, E unchanged (material_anisotropic.py:58-68)
                } else if (aniso) {
                    aniso_modes<GENERAL_EPS>(*aux, kl, nrm, mirror, k2a, e2a, k2b, e2b);  // eps in the shape frame
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
                    if (ident) { for (int c = 0; c < 3; ++c) el[c] = e[c]; }
                    else crot_t(st.frame.r, e, el);
                    const cplx kk = cdot(k2a, k2a);
                    cplx cc = cdot(el, k2a) / kk;
                    cplx tv[3] = {el[0] - cc * k2a[0], el[1] - cc * k2a[1], el[2] - cc * k2a[2]};
                    double t2 = herm2(tv);
                    if (!(t2 > 1e-24 * herm2(el))) {
                        const double ax = abs2(k2a[0]), ay = abs2(k2a[1]), az = abs2(k2a[2]);
                        cplx a[3] = {C(0.0), C(0.0), C(0.0)};
                        if (ax <= ay && ax <= az) a[0] = C(1.0); else if (ay <= az) a[1] = C(1.0); else a[2] = C(1.0);
                        cc = cdot(a, k2a) / kk;
                        for (int c = 0; c < 3; ++c) tv[c] = a[c] - cc * k2a[c];
                        t2 = herm2(tv);
                    }
                    const double inv = fast_rsqrt(t2);
                    for (int c = 0; c < 3; ++c) e2a[c] = inv * tv[c];
                }

                // ---- back to the global frame, record ----
                // (in place: the common cases -- untilted frames, every lane alive -- move nothing)
                if (no_deflect) {
                    for (int c = 0; c < 3; ++c) { k2a[c] = k[c]; e2a[c] = e[c]; }
                } else if (!ident) {
                    cplx t[3];
                    crot(st.frame.r, k2a, t); for (int c = 0; c < 3; ++c) k2a[c] = t[c];
                    crot(st.frame.r, e2a, t); for (int c = 0; c < 3; ++c) e2a[c] = t[c];
                    if (aniso) {
                        crot(st.frame.r, k2b, t); for (int c = 0; c < 3; ++c) k2b[c] = t[c];
                        crot(st.frame.r, e2b, t); for (int c = 0; c < 3; ++c) e2b[c] = t[c];
                    }
                }
                const bool split = aniso && (st.bits & kSplit);
                if (__any_sync(__activemask(), !alive)) {
                    const cplx qn = {qnan(), qnan()};
                    if (!alive) {
                        for (int c = 0; c < 3; ++c) { k2a[c] = e2a[c] = k2b[c] = e2b[c] = qn; }
                    }
                }
                cplx (&kga)[3] = k2a, (&ega)[3] = e2a, (&kgb)[3] = k2b, (&egb)[3] = e2b;
                const int64_t ld = st.ld_out;
                if (st.out_x)
                    for (int c = 0; c < 3; ++c) __stcs(st.out_x + c * ld + col, ok ? hit_g[c] : qnan());
                if (st.out_flags)
                    st.out_flags[col] = (uint8_t)((hit ? PYR_RAY_HIT : 0u) | (alive ? PYR_RAY_ALIVE : 0u));
                if (split) {
                    const int64_t ld2 = st.ld_out2;
                    for (int c = 0; c < 3; ++c) {
                        if (st.out_k) { cstore(st.out_k, c * ld2 + col, kga[c]); cstore(st.out_k, c * ld2 + w + col, kgb[c]); }
                        if (st.out_e) { cstore(st.out_e, c * ld2 + col, ega[c]); cstore(st.out_e, c * ld2 + w + col, egb[c]); }
                    }
                    w *= 2;             // mode a keeps its column in the doubled level
                } else {
                    for (int c = 0; c < 3; ++c) {
                        if (st.out_k) cstore(st.out_k, c * ld + col, kga[c]);
                        if (st.out_e) cstore(st.out_e, c * ld + col, ega[c]);
                    }
                }
                // continue with mode a
                for (int c = 0; c < 3; ++c) {
                    x[c] = alive ? hit_g[c] : qnan();
                    k[c] = kga[c];
                    e[c] = ega[c];
                }
            }
        }
    }
}

int pack_steps(const PyrStep *steps, int32_t n_steps, const PyrRaysIn *rays, int64_t n_rays,
               uint32_t flags, LaunchParams &P, bool &general, bool &any_aniso);   // pyr_trace.cu
int sm_count();

template <typename K>
static int launch_complex(K kernel, const LaunchParams &P, int64_t n_rays, cudaStream_t stream) {
    const int threads = 128;
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0);
    if (e != cudaSuccess) return (int)e;
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (n_rays + threads - 1) / threads;
    const int64_t cap = (int64_t)sm_count() * per_sm;
    if (grid > cap) grid = cap;
    kernel<<<(unsigned)grid, threads, 0, stream>>>(P);
    e = cudaGetLastError();
    return e == cudaSuccess ? PYR_OK : (int)e;
}

int trace_complex(const PyrStep *steps, int32_t n_steps, const PyrRaysIn *rays, int64_t n_rays,
                  uint32_t flags, cudaStream_t stream) {
    static thread_local LaunchParams P;
    bool general = false, any_aniso = false;
    int rc = pack_steps(steps, n_steps, rays, n_rays, flags, P, general, any_aniso);
    if (rc != PYR_OK) return rc;
    if (rays->gen) return PYR_E_UNSUPPORTED;
    if (!rays->e) return PYR_E_BADARG;            // E defines the ray direction in crystals
    if (rays->n_waves > 1) return PYR_E_UNSUPPORTED;
    int splits = 0;
    bool general_eps = false, explicit_shapes = false;
    for (int s = 0; s < n_steps; ++s) {
        if (steps[s].before.kind == PYR_MEDIUM_ISO_GRIN || steps[s].after.kind == PYR_MEDIUM_ISO_GRIN)
            return PYR_E_UNSUPPORTED;
        const bool deflects = steps[s].after.kind == PYR_MEDIUM_ANISO && steps[s].mode != PYR_STEP_PROPAGATE_ONLY;
        if (steps[s].split && deflects) {
            ++splits;
            // a doubling step that is not the last one is resumed from its own record
            if (s != n_steps - 1 && (!steps[s].out_x || !steps[s].out_k || !steps[s].out_e || !steps[s].out_flags))
                return PYR_E_BADARG;
        }
        if (deflects && P.steps[s].aux >= 0 && !P.aux[P.steps[s].aux].uniaxial) general_eps = true;
        if (steps[s].shape_kind != PYR_SHAPE_CONIC && steps[s].shape_kind != PYR_SHAPE_CYLINDER &&
            steps[s].mode != PYR_STEP_DEFLECT_ONLY)
            explicit_shapes = true;
        if (steps[s].shape_kind != PYR_SHAPE_CONIC && steps[s].shape_kind != PYR_SHAPE_CYLINDER)
            explicit_shapes = true;             // (the normal of a deflect-only step needs it too)
    }
    if (splits > kMaxSplits) return PYR_E_TOOLARGE;
    if (n_rays == 0) return PYR_OK;
    if (explicit_shapes)
        return general_eps ? launch_complex(trace_complex_kernel<true, true>, P, n_rays, stream)
                           : launch_complex(trace_complex_kernel<false, true>, P, n_rays, stream);
    return general_eps ? launch_complex(trace_complex_kernel<true, false>, P, n_rays, stream)
                       : launch_complex(trace_complex_kernel<false, false>, P, n_rays, stream);
}

}  // namespace pyr
