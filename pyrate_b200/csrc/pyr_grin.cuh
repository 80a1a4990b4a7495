// GRIN media: index profiles and the 4th-order symplectic integrator.
// Mirrors raytracer/material/material_grin.py of the reference:
//   symplecticintegrator :106-213 (Forest-Ruth coefficients :116-123, drift/kick
//   loop :139-157, energy test :164-176, "final" test :181-186, boundary :189-193,
//   freeze rule :195-196, k = p/n :198-199), propagate :215-220.
// Normalisations (documented in DESIGN.md): the energy test is per ray instead of
// bundle-summed, and a ray stops stepping once it is final instead of stepping in
// lock-step with the slowest ray of the bundle.
#pragma once

#include "pyr_exp.cuh"
#include "pyr_shapes.cuh"

namespace pyr {

// index n and gradient at material-frame position q
// etab: the 2^(j/512) table of pyr_exp.cuh in shared memory
__device__ __forceinline__ double grin_index(const DMedium &m, const double q[3], double g[3],
                                             bool want_grad, const double *etab) {
    if (m.profile == PYR_GRIN_GAUSSIAN_XY) {
        // n = p0 + p1 E, E = exp(-(p2 x^2 + p3 y^2)); grad n = -2 p1 E (p2 x, p3 y, 0)
        const double w0 = m.p[2] * q[0], w1 = m.p[3] * q[1];
        const double ex = exp_tab_any(-fma(w0, q[0], w1 * q[1]), etab);
        if (want_grad) {
            const double h = -2.0 * m.p[1] * ex;
            g[0] = h * w0;
            g[1] = h * w1;
            g[2] = 0.0;
        }
        return fma(m.p[1], ex, m.p[0]);
    }
    // PYR_GRIN_POLY_RZ
    const double r2 = fma(q[0], q[0], q[1] * q[1]);
    const double z = q[2];
    const double nr = fma(fma(fma(m.p[3], r2, m.p[2]), r2, m.p[1]), r2, m.p[0]);
    const double nz = z * fma(fma(m.p[6], z, m.p[5]), z, m.p[4]);
    if (want_grad) {
        const double dr = 2.0 * fma(fma(3.0 * m.p[3], r2, 2.0 * m.p[2]), r2, m.p[1]);
        g[0] = q[0] * dr;
        g[1] = q[1] * dr;
        g[2] = fma(fma(3.0 * m.p[6], z, 2.0 * m.p[5]), z, m.p[4]);
    }
    return nr + nz;
}

__device__ __forceinline__ bool grin_inside(const DMedium &m, const double q[3]) {
    switch (m.boundary) {
        case PYR_BND_CYLINDER: return fma(q[0], q[0], q[1] * q[1]) < m.b[0] * m.b[0];
        case PYR_BND_BOX: return fabs(q[0]) < m.b[0] && fabs(q[1]) < m.b[1];
        case PYR_BND_SPHERE: return dot3(q, q) < m.b[0] * m.b[0];
        default: return true;
    }
}

// Integrates from global point x along global unit direction d through medium m
// until the next surface (shape of the current step) is crossed.  On return x is
// the last position BEFORE the crossing and k = p/n there (both global);
// returns validity (energy, boundary, step cap).
// History (aux->hist_*, optional): one row per integrator step with the frozen state
// the reference appends (material_grin.py:195-205).
template <bool EXT>
__device__ __forceinline__ bool grin_propagate(const DMedium &m, int shape_kind, const DAux *aux,
                                               double curv, double cc, double x[3],
                                               const double d[3], double k[3],
                                               int64_t ray_index, int64_t ld_hist,
                                               const double *etab) {
    const double c0 = 1.0 / (2.0 * (2.0 - 1.2599210498948732));
    const double c1 = (1.0 - 1.2599210498948732) / (2.0 * (2.0 - 1.2599210498948732));
    const double d0 = 1.0 / (2.0 - 1.2599210498948732);
    const double d1 = -1.2599210498948732 / (2.0 - 1.2599210498948732);
    const double cs[4] = {c0, c1, c1, c0};
    const double ds[4] = {d0, d1, d0, 0.0};

    // dead / out-of-range rays carry NaN: every comparison below would be false for them
    // and the loop would only end at the step cap -- they do not enter the integrator
    if (ray_index < 0) return false;
    double q[3], p[3], g[3];
    g2l_point(m.frame, x, q);
    rot_t(m.frame.r, d, p);
    const double n0 = grin_index(m, q, g, false, etab);
    p[0] *= n0; p[1] *= n0; p[2] *= n0;
    double uq[3] = {q[0], q[1], q[2]}, up[3] = {p[0], p[1], p[2]};
    const double tau2 = 2.0 * m.ds;
    const int cap = m.max_steps > 0 ? m.max_steps : 1000000;
    bool valid = true;
    for (int it = 0; it < cap; ++it) {
        double nq = 0.0;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            q[0] = fma(tau2 * cs[s], p[0], q[0]);
            q[1] = fma(tau2 * cs[s], p[1], q[1]);
            q[2] = fma(tau2 * cs[s], p[2], q[2]);
            nq = grin_index(m, q, g, s < 3, etab);
            if (s < 3) {
                const double f = tau2 * ds[s] * nq;
                p[0] = fma(f, g[0], p[0]);
                p[1] = fma(f, g[1], p[1]);
                p[2] = fma(f, g[2], p[2]);
            }
        }
        if (!(fabs(dot3(p, p) - nq * nq) <= m.energy_tol)) valid = false;     // NaN-safe
        double xs[3];
        l2g_point(m.to_shape, q, xs);
        const double gap = xs[2] - shape_sag<EXT>(shape_kind, aux, curv, cc, xs[0], xs[1]);
        const bool crossed = gap > 0.0;
        if (!grin_inside(m, q) || gap != gap) valid = false;               // NaN sag: off the shape
        const bool stop = crossed || !valid;
        if (!stop) {
            uq[0] = q[0]; uq[1] = q[1]; uq[2] = q[2];
            up[0] = p[0]; up[1] = p[1]; up[2] = p[2];
            if (it == cap - 1) valid = false;
        }
        if (aux->hist_count && ray_index >= 0) aux->hist_count[ray_index] = it + 1;
        if (aux->hist_x && ray_index >= 0 && it < aux->hist_rows) {
            double gq[3], kq[3], kg[3], dummy[3];
            const double invn = 1.0 / grin_index(m, uq, dummy, false, etab);
            kq[0] = up[0] * invn; kq[1] = up[1] * invn; kq[2] = up[2] * invn;
            l2g_point(m.frame, uq, gq);
            rot(m.frame.r, kq, kg);
            for (int c = 0; c < 3; ++c) {
                aux->hist_x[((int64_t)it * 3 + c) * ld_hist + ray_index] = gq[c];
                if (aux->hist_k) aux->hist_k[((int64_t)it * 3 + c) * ld_hist + ray_index] = kg[c];
            }
            if (aux->hist_valid) aux->hist_valid[(int64_t)it * ld_hist + ray_index] = valid ? 1 : 0;
        }
        if (stop) break;
    }
    const double inv = 1.0 / grin_index(m, uq, g, false, etab);
    const double kl[3] = {up[0] * inv, up[1] * inv, up[2] * inv};
    l2g_point(m.frame, uq, x);
    rot(m.frame.r, kl, k);
    return valid;
}

// ---- straight-line building blocks of the specialised integrator ----
// index (and gradient) of N positions, operation by operation over the N rays.
// Gaussian profile: E = exp(-(p2 x^2 + p3 y^2)) and w = (p2 x, p3 y) are returned instead of
// the gradient -- grad n = -2 p1 E (w0, w1, 0); the kick folds the constant factors.
template <int PROFILE, int N>
__device__ __forceinline__ void grin_index_n(const DMedium &m, const double (*q)[3], double *nn,
                                             double (*g)[3], const double *etab) {
    if (PROFILE == PYR_GRIN_GAUSSIAN_XY) {
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const double w0 = m.p[2] * q[j][0], w1 = m.p[3] * q[j][1];
            const double ex = exp_tab_any(-fma(w0, q[j][0], w1 * q[j][1]), etab);
            g[j][0] = w0; g[j][1] = w1; g[j][2] = ex;
            nn[j] = fma(m.p[1], ex, m.p[0]);
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const double r2 = fma(q[j][0], q[j][0], q[j][1] * q[j][1]);
        const double z = q[j][2];
        const double nr = fma(fma(fma(m.p[3], r2, m.p[2]), r2, m.p[1]), r2, m.p[0]);
        const double nz = z * fma(fma(m.p[6], z, m.p[5]), z, m.p[4]);
        const double dr = 2.0 * fma(fma(3.0 * m.p[3], r2, 2.0 * m.p[2]), r2, m.p[1]);
        g[j][0] = q[j][0] * dr;
        g[j][1] = q[j][1] * dr;
        g[j][2] = fma(fma(3.0 * m.p[6], z, 2.0 * m.p[5]), z, m.p[4]);
        nn[j] = nr + nz;
    }
}

// boundary test without a branch on the boundary kind; BND >= 0: kind fixed at compile time
template <int BND>
__device__ __forceinline__ bool grin_inside_nb(const DMedium &m, const double q[3]) {
    const double r2 = fma(q[0], q[0], q[1] * q[1]);
    if (BND == PYR_BND_CYLINDER) return r2 < m.b[0] * m.b[0];
    const bool cyl = r2 < m.b[0] * m.b[0];
    const bool box = (fabs(q[0]) < m.b[0]) & (fabs(q[1]) < m.b[1]);
    const bool sph = fma(q[2], q[2], r2) < m.b[0] * m.b[0];
    const int k = m.boundary;
    return (k == PYR_BND_NONE) | ((k == PYR_BND_CYLINDER) & cyl) | ((k == PYR_BND_BOX) & box) |
           ((k == PYR_BND_SPHERE) & sph);
}

// N rays of one thread integrated TOGETHER: every iteration advances all N rays by one
// integrator step in STRAIGHT-LINE code (profile fixed at compile time, conic next surface,
// no branch between or inside the rays' steps), so that the compiler interleaves the N
// independent dependency chains -- a warp issues in order, and the single-ray loop is bound
// by the latency of its own chain (drift -> exp -> kick, four times per step;
// profiles/r02_grin.md).  A ray that has stopped keeps stepping like in the reference's
// lock-step loop (material_grin.py:139) but its frozen state (uq, up) and validity no longer
// change, so every ray's result is bit-identical to grin_propagate's.  No integrator history
// (the history mode runs the single-ray function).
template <int PROFILE, int BND, int N>
__device__ __forceinline__ void grin_propagate_n(const DMedium &m, double curv, double cc,
                                                 double (*x)[3], const double (*d)[3], double (*k)[3],
                                                 const bool *enter, bool *valid_out,
                                                 const double *etab) {
    const double c0 = 1.0 / (2.0 * (2.0 - 1.2599210498948732));
    const double c1 = (1.0 - 1.2599210498948732) / (2.0 * (2.0 - 1.2599210498948732));
    const double d0 = 1.0 / (2.0 - 1.2599210498948732);
    const double d1 = -1.2599210498948732 / (2.0 - 1.2599210498948732);
    const double cs[4] = {c0, c1, c1, c0};
    const double ds[4] = {d0, d1, d0, 0.0};
    double q[N][3], p[N][3], uq[N][3], up[N][3];
    bool valid[N], done[N];
    double g[N][3], nq[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        g2l_point(m.frame, x[j], q[j]);
        rot_t(m.frame.r, d[j], p[j]);
    }
    grin_index_n<PROFILE, N>(m, q, nq, g, etab);
#pragma unroll
    for (int j = 0; j < N; ++j) {
#pragma unroll
        for (int c = 0; c < 3; ++c) { p[j][c] *= nq[j]; uq[j][c] = q[j][c]; up[j][c] = p[j][c]; }
        valid[j] = enter[j];
        done[j] = !enter[j];          // dead / out-of-range rays (NaN state) never enter
    }
    const double tau2 = 2.0 * m.ds;
    const double kgrad = -2.0 * m.p[1];                  // Gaussian profile: grad n = kgrad E w
    const int cap = m.max_steps > 0 ? m.max_steps : 1000000;
    for (int it = 0; it < cap; ++it) {
        bool all_done = true;
        // stage-major: every operation on all N rays before the next one
#pragma unroll
        for (int s = 0; s < 4; ++s) {
#pragma unroll
            for (int j = 0; j < N; ++j) {
                q[j][0] = fma(tau2 * cs[s], p[j][0], q[j][0]);
                q[j][1] = fma(tau2 * cs[s], p[j][1], q[j][1]);
                // the Gaussian-xy profile has no z gradient: p_z is constant, the profile does
                // not read z, and the four drifts add up to tau2 p_z (c0 + c1 + c1 + c0 = 1)
                if (PROFILE != PYR_GRIN_GAUSSIAN_XY) q[j][2] = fma(tau2 * cs[s], p[j][2], q[j][2]);
                else if (s == 3) q[j][2] = fma(tau2, p[j][2], q[j][2]);
            }
            grin_index_n<PROFILE, N>(m, q, nq, g, etab);
            if (s < 3) {
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    if (PROFILE == PYR_GRIN_GAUSSIAN_XY) {
                        // p += tau2 d_s n grad n = (tau2 d_s (-2 p1) n E) (w0, w1, 0)
                        const double h = (tau2 * ds[s] * kgrad) * nq[j] * g[j][2];
                        p[j][0] = fma(h, g[j][0], p[j][0]);
                        p[j][1] = fma(h, g[j][1], p[j][1]);
                    } else {
                        const double f = tau2 * ds[s] * nq[j];
                        p[j][0] = fma(f, g[j][0], p[j][0]);
                        p[j][1] = fma(f, g[j][1], p[j][1]);
                        p[j][2] = fma(f, g[j][2], p[j][2]);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const bool e_ok = fabs(dot3(p[j], p[j]) - nq[j] * nq[j]) <= m.energy_tol;      // NaN -> false
            double xs[3];
            l2g_point(m.to_shape, q[j], xs);
            // "z > sag(x, y)" (material_grin.py:181) without the square root and the division of
            // the sag: z1 = sag is the vertex-branch root of F(z) = A z^2 - 2 z + c r^2, A = c (1 + cc);
            // z > z1  <=>  F < 0 (or beyond the far root, A z > 1) for A >= 0,  F < 0 and A z < 1 for
            // A < 0.  The sag is undefined (NaN in the reference, conic_function :214-216: the ray
            // can neither finish nor stay valid) where 1 - (1 + cc) c^2 r^2 <= 0.
            const double r2 = fma(xs[0], xs[0], xs[1] * xs[1]);
            const double A = curv * (1.0 + cc);
            const bool defined = fma(-A * curv, r2, 1.0) > 0.0;
            const double F = fma(A * xs[2] - 2.0, xs[2], curv * r2);
            const double t = A * xs[2];
            const bool beyond = (A > 0.0) ? ((F < 0.0) | (t > 1.0)) : ((F < 0.0) & (t < 1.0));
            const bool crossed = defined & beyond;
            const bool v = valid[j] & e_ok & grin_inside_nb<BND>(m, q[j]) & defined & (xs[2] == xs[2]);
            const bool stop = crossed | !v;
            const bool live = !done[j];
            const bool advance = live & !stop;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                uq[j][c] = advance ? q[j][c] : uq[j][c];
                up[j][c] = advance ? p[j][c] : up[j][c];
            }
            valid[j] = live ? (v & !(advance & (it == cap - 1))) : valid[j];
            done[j] = done[j] | stop;
            all_done = all_done & done[j];
        }
        if (all_done) break;
    }
    grin_index_n<PROFILE, N>(m, uq, nq, g, etab);
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const double inv = 1.0 / nq[j];
        const double kl[3] = {up[j][0] * inv, up[j][1] * inv, up[j][2] * inv};
        if (enter[j]) {
            l2g_point(m.frame, uq[j], x[j]);
            rot(m.frame.r, kl, k[j]);
        }
        valid_out[j] = valid[j];
    }
}

// dispatcher of a thread's N rays: the interleaved loop for the catalogue profiles in front
// of a conic surface, else one ray after the other through the general function
template <bool EXT, int N>
__device__ __forceinline__ void grin_propagate_rays(const DMedium &m, int shape_kind, const DAux *aux,
                                                    double curv, double cc, double (*x)[3],
                                                    const double (*d)[3], double (*k)[3],
                                                    const bool *enter, bool *valid_out,
                                                    const double *etab) {
    if (shape_kind == PYR_SHAPE_CONIC && m.profile == PYR_GRIN_GAUSSIAN_XY) {
        // (the cylinder boundary of demos/demo_grin.py has its own copy of the loop)
        if (m.boundary == PYR_BND_CYLINDER)
            grin_propagate_n<PYR_GRIN_GAUSSIAN_XY, PYR_BND_CYLINDER, N>(m, curv, cc, x, d, k, enter, valid_out, etab);
        else
            grin_propagate_n<PYR_GRIN_GAUSSIAN_XY, -1, N>(m, curv, cc, x, d, k, enter, valid_out, etab);
    } else if (shape_kind == PYR_SHAPE_CONIC && m.profile == PYR_GRIN_POLY_RZ) {
        grin_propagate_n<PYR_GRIN_POLY_RZ, -1, N>(m, curv, cc, x, d, k, enter, valid_out, etab);
    } else {
#pragma unroll 1
        for (int j = 0; j < N; ++j)
            valid_out[j] = grin_propagate<EXT>(m, shape_kind, aux, curv, cc, x[j], d[j], k[j],
                                               enter[j] ? 0 : -1, 0, etab);
    }
}

}  // namespace pyr
