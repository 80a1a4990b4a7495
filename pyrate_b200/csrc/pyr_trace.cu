// pyrate_b200 trace kernels (sm_100a) and C ABI (include/pyrate_b200.h).
//
// One persistent launch walks every ray of the bundle through the whole element
// sequence: the ray state (x, k[, E]) lives in registers from the first surface
// to the last, each thread owns two adjacent rays of the (3, n) component-major
// arrays so every HBM access is a coalesced 128-bit load/store, the per-step
// record (hit point, wave vector after deflection, flag byte) is streamed out
// with evict-first stores, and the step table sits in the kernel parameter block
// (constant bank).  FP64 throughout: parity with the reference is <= 1e-10.
// HBM-bound integer/FP64 streaming work; no tensor cores (no dense contraction).
//
// Reference path replaced (pyrateoptics/raytracer, file:line):
//   optical_element.py:336-375   per-surface loop
//   surface.py:116-135           Surface.intersect (+ aperture.py:71-140)
//   surface_shape.py:151-155, 289-325, 448-465   Shape.intersect
//   ray.py:136-161               returnKtoD, getLocalSurfaceNormal
//   material/material_isotropic.py:137-247   Snell via in-plane k, propagate
//   material/material_grin.py:106-220        GRIN symplectic propagate
//   localcoordinates.py:383-413  global <-> local maps
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "pyr_device.cuh"
#include "pyr_gen.cuh"
#include "pyr_grin.cuh"
#include "pyr_shapes.cuh"

namespace pyr {


template <bool WITH_E>
struct Ray {
    double x[3], k[3];
    double e[WITH_E ? 3 : 1];
    bool alive;
};

// ---- mbarrier + 1-D TMA bulk copy (global -> shared), sm_90+/sm_100a ----
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_load_1d(double *smem_dst, const double *gsrc, unsigned bytes,
                                            unsigned long long *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_store_1d(void *gdst, const void *smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// Record stores.  POLICY 0: evict-first (st.global.cs) -- the records are written once
// and never re-read by the kernel; 1: default write-back; 2: cache-global (st.global.cg).
template <int POLICY>
__device__ __forceinline__ void store_stream(double *p, double v) {
    if ((POLICY & 3) == 0 || (POLICY & 3) == 3) __stcs(p, v); else if ((POLICY & 3) == 2) __stcg(p, v); else *p = v;
}
template <int POLICY>
__device__ __forceinline__ void store_stream2(double *p, double a, double b) {
    double2 *q = reinterpret_cast<double2 *>(p);
    const double2 v = make_double2(a, b);
    if ((POLICY & 3) == 0 || (POLICY & 3) == 3) __stcs(q, v); else if ((POLICY & 3) == 2) __stcg(q, v); else *q = v;
}

// d from (k, E): ray.py:140-152 for real k, E
__device__ __forceinline__ void poynting_dir(const double k[3], const double e[3], double d[3]) {
    const double ee = dot3(e, e), ek = dot3(e, k);
    double s[3] = {fma(ee, k[0], -ek * e[0]), fma(ee, k[1], -ek * e[1]), fma(ee, k[2], -ek * e[2])};
    const double inv = fast_rsqrt(dot3(s, s));
    d[0] = s[0] * inv; d[1] = s[1] * inv; d[2] = s[2] * inv;
}

// Deterministic E perpendicular to the new k (the reference's choice is an
// SVD null-space vector, arbitrary in the plane; only |E| = 1, E.k = 0 matter).
__device__ __forceinline__ void reproject_e(const double k[3], double e[3]) {
    const double kk = dot3(k, k);
    double c = dot3(e, k) / kk;
    double t[3] = {fma(-c, k[0], e[0]), fma(-c, k[1], e[1]), fma(-c, k[2], e[2])};
    double tt = dot3(t, t);
    if (!(tt > 1e-24 * dot3(e, e))) {
        // E was parallel to k: restart from the axis least aligned with k
        const double ax = fabs(k[0]), ay = fabs(k[1]), az = fabs(k[2]);
        double a[3] = {0.0, 0.0, 0.0};
        if (ax <= ay && ax <= az) a[0] = 1.0; else if (ay <= az) a[1] = 1.0; else a[2] = 1.0;
        c = dot3(a, k) / kk;
        t[0] = fma(-c, k[0], a[0]); t[1] = fma(-c, k[1], a[1]); t[2] = fma(-c, k[2], a[2]);
        tt = dot3(t, t);
    }
    const double inv = rsqrt(tt);
    e[0] = t[0] * inv; e[1] = t[1] * inv; e[2] = t[2] * inv;
}

// One sequence entry for the N rays a thread owns (real k, E); fl[j] = flag byte of ray j.
// The rays go through every phase together, so the straight-line parts interleave N
// independent dependency chains and the Newton iteration of an explicit shape is ONE loop
// over all of them (csrc/pyr_shapes.cuh).
// ASPH: the only explicit shape of the sequence is the even asphere -- the instantiation
// carries neither the XY-polynomial / biconic evaluators nor the generic Newton loop (small
// enough to stay resident in the instruction cache) and uses asphere_t_n
template <int N, bool WITH_E, bool HAS_GRIN, bool EXT, bool ASPH = false>
__device__ __forceinline__ void step_real_n(const LaunchParams &P, const DStep &st,
                                            Ray<WITH_E> (&r)[N], double (&d)[N][3], double (&hit_g)[N][3],
                                            const int64_t (&ray_index)[N], const int (&w)[N],
                                            const double *etab, bool grin_media, uint32_t (&fl)[N]) {
    constexpr bool GENERAL = true;
    const DAux *aux = (st.aux >= 0) ? &P.aux[st.aux] : nullptr;
    const bool ident = (st.bits & kRotIdentity) != 0;
    bool ok[N];
#pragma unroll
    for (int j = 0; j < N; ++j) ok[j] = r[j].alive;

    // ---- propagate through a GRIN medium (material_grin.py:215-220) ----
    if (HAS_GRIN && st.before_kind == PYR_MEDIUM_ISO_GRIN) {
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const bool v = grin_propagate<EXT>(aux->before, st.shape_kind, aux, st.curv, st.cc, r[j].x, d[j], r[j].k,
                                               ok[j] ? ray_index[j] : -1, st.ld_out, etab);
            ok[j] = ok[j] && v;
            const double inv = fast_rsqrt(dot3(r[j].k, r[j].k));
            d[j][0] = r[j].k[0] * inv; d[j][1] = r[j].k[1] * inv; d[j][2] = r[j].k[2] * inv;
        }
    }

    // ---- into the shape frame (surface_shape.py:151-155) ----
    double r0[N][3], dl[N][3];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        if (ident) {
            r0[j][0] = r[j].x[0] - st.frame.o[0]; r0[j][1] = r[j].x[1] - st.frame.o[1];
            r0[j][2] = r[j].x[2] - st.frame.o[2];
            dl[j][0] = d[j][0]; dl[j][1] = d[j][1]; dl[j][2] = d[j][2];
        } else {
            g2l_point(st.frame, r[j].x, r0[j]);
            rot_t(st.frame.r, d[j], dl[j]);
        }
    }

    // ---- intersect ----
    double t[N], gfx[N], gfy[N];
    bool hit_ok[N], grad_ok[N];
#pragma unroll
    for (int j = 0; j < N; ++j) { hit_ok[j] = true; grad_ok[j] = false; gfx[j] = gfy[j] = 0.0; t[j] = 0.0; }
    if (GENERAL && (st.bits & kNoIntersect)) {
        // t = 0
    } else if (!GENERAL || st.shape_kind == PYR_SHAPE_CONIC) {
#pragma unroll
        for (int j = 0; j < N; ++j) t[j] = conic_t(st.curv, st.cc, r0[j], dl[j], hit_ok[j]);
    } else if (!ASPH && st.shape_kind == PYR_SHAPE_CYLINDER) {
#pragma unroll
        for (int j = 0; j < N; ++j) t[j] = cylinder_t(st.curv, st.cc, r0[j], dl[j], hit_ok[j]);
    } else if (ASPH) {
        asphere_t_n<N>(*aux, st.curv, st.cc, r0, dl, ok, t, gfx, gfy);
#pragma unroll
        for (int j = 0; j < N; ++j) grad_ok[j] = true;     // gradient of the returned point itself
    } else {
        explicit_t_n<EXT, N>(st.shape_kind, *aux, st.curv, st.cc, r0, dl, ok, t, gfx, gfy, grad_ok);
    }

#pragma unroll
    for (int j = 0; j < N; ++j) {
        const double h[3] = {fma(dl[j][0], t[j], r0[j][0]), fma(dl[j][1], t[j], r0[j][1]),
                             fma(dl[j][2], t[j], r0[j][2])};
        if (ident) {
            hit_g[j][0] = h[0] + st.frame.o[0]; hit_g[j][1] = h[1] + st.frame.o[1]; hit_g[j][2] = h[2] + st.frame.o[2];
        } else {
            l2g_point(st.frame, h, hit_g[j]);
        }

        // ---- aperture (surface.py:127-135) ----
        bool ap_ok = true;
        if (st.aperture_kind != PYR_AP_BASE) {
            double ax = h[0], ay = h[1];
            if (GENERAL && !(st.bits & kApSameFrame)) {
                double a[3];
                g2l_point(aux->aperture_frame, hit_g[j], a);
                ax = a[0]; ay = a[1];
            }
            if (st.aperture_kind == PYR_AP_CIRCULAR) {
                const double rr = fma(ax, ax, ay * ay);
                ap_ok = (rr >= st.ap0) && (rr <= st.ap1);
            } else {
                ap_ok = (ax >= -st.ap0) && (ax <= st.ap0) && (ay >= -st.ap1) && (ay <= st.ap1);
            }
        }
        const bool hit = ok[j] && hit_ok[j] && ap_ok;

        // ---- surface normal in the shape frame (ray.py:156-161) ----
        double nrm[3];
        if (!GENERAL || st.shape_kind == PYR_SHAPE_CONIC) {
            conic_normal(st.curv, st.cc, (st.bits & kSphere) != 0, h[0], h[1], nrm);
        } else if (!ASPH && st.shape_kind == PYR_SHAPE_CYLINDER) {
            cylinder_normal(st.curv, st.cc, h[1], nrm);
        } else if (ASPH || grad_ok[j]) {
            // gradient of the converged Newton iterate: within tol of the hit point
            normal_from_gradient(gfx[j], gfy[j], nrm);
        } else {
            explicit_normal<EXT>(st.shape_kind, *aux, st.curv, st.cc, h[0], h[1], nrm);
        }

        // ---- deflection (material_isotropic.py:163-236), done in the shape frame ----
        double kl[3];
        if (ident) { kl[0] = r[j].k[0]; kl[1] = r[j].k[1]; kl[2] = r[j].k[2]; }
        else rot_t(st.frame.r, r[j].k, kl);
        double n2sq = st.n2sq[w[j]];
        if (aux && aux->after_n_rays) {
            // index of a position-dependent medium the caller evaluated at the hit points
            const double nn = ray_index[j] >= 0 ? aux->after_n_rays[ray_index[j]] : qnan();
            n2sq = nn * nn;
        } else if (grin_media && st.after_kind == PYR_MEDIUM_ISO_GRIN) {
            double q[3], g[3];
            g2l_point(aux->after.frame, hit_g[j], q);
            const double nn = grin_index(aux->after, q, g, false, etab);
            n2sq = nn * nn;
        }
        const double kn = dot3(kl, nrm);
        // k_inplane = k - (k.n) n ; square = n^2 - k_inplane.k_inplane
        const double kin[3] = {fma(-kn, nrm[0], kl[0]), fma(-kn, nrm[1], kl[1]), fma(-kn, nrm[2], kl[2])};
        const double square = n2sq - dot3(kin, kin);
        const double xi = fast_sqrt(square);
        // (a non-finite normal makes `square` NaN or -inf: no separate finite check)
        const bool refr_ok = square > 0.0;
        double k2[3];
        if (st.interaction == PYR_REFLECT) {
            k2[0] = fma(xi, nrm[0], -kin[0]); k2[1] = fma(xi, nrm[1], -kin[1]); k2[2] = fma(xi, nrm[2], -kin[2]);
        } else {
            k2[0] = fma(xi, nrm[0], kin[0]); k2[1] = fma(xi, nrm[1], kin[1]); k2[2] = fma(xi, nrm[2], kin[2]);
        }
        bool alive = hit && refr_ok;
        if (GENERAL && (st.bits & kNoDeflect)) {
            alive = hit;                       // k unchanged
        } else if (ident) { r[j].k[0] = k2[0]; r[j].k[1] = k2[1]; r[j].k[2] = k2[2]; }
        else rot(st.frame.r, k2, r[j].k);
        r[j].x[0] = hit_g[j][0]; r[j].x[1] = hit_g[j][1]; r[j].x[2] = hit_g[j][2];
        if (__any_sync(__activemask(), !alive)) {          // (nothing to fill while every lane lives)
            if (!ok[j]) { hit_g[j][0] = hit_g[j][1] = hit_g[j][2] = qnan(); }
            if (!alive) {
                const double q = qnan();
                r[j].x[0] = r[j].x[1] = r[j].x[2] = q;
                r[j].k[0] = r[j].k[1] = r[j].k[2] = q;
            }
        }
        if (WITH_E) {
            if (!alive) r[j].e[0] = r[j].e[1] = r[j].e[2] = qnan();
            else if (!(GENERAL && (st.bits & kNoDeflect))) reproject_e(r[j].k, r[j].e);
        }
        r[j].alive = alive;
        fl[j] = (hit ? PYR_RAY_HIT : 0u) | (alive ? PYR_RAY_ALIVE : 0u);
    }
}

// Tuned step for the common case (conic shape, homogeneous isotropic media,
// aperture in the shape frame): same arithmetic as step_real, but
//   * plane surfaces (curv == 0) take a closed form without the conic machinery,
//   * sqrt / division run branch-free (fast_sqrt / fast_div, ~1 ulp),
//   * validity rides on NaN propagation: a ray that dies gets k = NaN once, every
//     later hit point and wave vector is NaN by arithmetic, and `square > 0`
//     already rejects NaN normals (no separate finite check).
// PLAIN (DStep bit kPlain): untilted frame, no aperture, refraction, sphere or plane -- the
// thirteen entries of the double-Gauss -- without the branches that decide those properties
template <bool WITH_E, bool PLAIN = false>
__device__ __forceinline__ uint32_t step_lean(const DStep &st, Ray<WITH_E> &r, const double d[3],
                                              double hit_g[3], int w = 0) {
    const bool ok = r.alive;
    const bool ident = PLAIN || (st.bits & kRotIdentity) != 0;
    double r0[3], dl[3], kl[3];
    if (ident) {
        r0[0] = r.x[0] - st.frame.o[0]; r0[1] = r.x[1] - st.frame.o[1]; r0[2] = r.x[2] - st.frame.o[2];
        dl[0] = d[0]; dl[1] = d[1]; dl[2] = d[2];
        kl[0] = r.k[0]; kl[1] = r.k[1]; kl[2] = r.k[2];
    } else {
        g2l_point(st.frame, r.x, r0);
        rot_t(st.frame.r, d, dl);
        rot_t(st.frame.r, r.k, kl);
    }
    const double curv = st.curv, cc = st.cc;
    double t, nrm[3];
    bool hit_ok = true;
    double h[3];
    const bool plane = (st.bits & kPlane) != 0;
    if (plane && dl[2] > 0.0) {
        // F = d_z, G = -2 z0, H = 0  ->  t = -z0 / d_z  (surface_shape.py:305-318)
        t = -r0[2] * fast_rcp(dl[2]);
        h[0] = fma(dl[0], t, r0[0]); h[1] = fma(dl[1], t, r0[1]); h[2] = fma(dl[2], t, r0[2]);
        nrm[0] = 0.0; nrm[1] = 0.0; nrm[2] = 1.0;
    } else if (PLAIN || (st.bits & kSphere)) {
        // cc = 0:  F = d_z - c (d.r0),  G = c (r0.r0) - 2 z0,  H = -c
        const double F = fma(-curv, dot3(dl, r0), dl[2]);
        const double G = fma(curv, dot3(r0, r0), -2.0 * r0[2]);
        const double square = fma(F, F, -curv * G);
        hit_ok = square >= 0.0;
        t = fast_div(G, F + fast_sqrt(square));
        h[0] = fma(dl[0], t, r0[0]); h[1] = fma(dl[1], t, r0[1]); h[2] = fma(dl[2], t, r0[2]);
        // unit normal (-c x, -c y, sqrt(1 - c^2 r^2)); on the vertex branch of the sphere
        // sqrt(1 - c^2 r^2) == 1 - c z exactly (surface equation), which saves the sqrt.
        // The far branch (1 - c z <= 0) keeps the reference's sag-based value.
        double gz = fma(-curv, h[2], 1.0);
        if (!(gz > 0.0)) gz = fast_sqrt(fma(-curv * curv, fma(h[0], h[0], h[1] * h[1]), 1.0));
        nrm[0] = -curv * h[0]; nrm[1] = -curv * h[1]; nrm[2] = gz;
    } else {
        const double cc1 = 1.0 + cc;
        const double F = dl[2] - curv * fma(dl[0], r0[0], fma(dl[1], r0[1], dl[2] * r0[2] * cc1));
        const double G = curv * fma(r0[0], r0[0], fma(r0[1], r0[1], r0[2] * r0[2] * cc1)) - 2.0 * r0[2];
        const double H = fma(-cc * curv * dl[2], dl[2], -curv);
        const double square = fma(F, F, H * G);
        hit_ok = square >= 0.0;
        t = fast_div(G, F + fast_sqrt(square));
        h[0] = fma(dl[0], t, r0[0]); h[1] = fma(dl[1], t, r0[1]); h[2] = fma(dl[2], t, r0[2]);
        // grad = (-c x, -c y, sqrt(1 - (1+cc) c^2 r^2)); the z component equals
        // 1 - c (1+cc) z on the vertex branch (see the sphere case)
        double gx = -curv * h[0], gy = -curv * h[1];
        double gz = fma(-curv * cc1, h[2], 1.0);
        if (!(gz > 0.0)) gz = fast_sqrt(fma(-cc1 * curv * curv, fma(h[0], h[0], h[1] * h[1]), 1.0));
        const double inv = fast_rsqrt(fma(gx, gx, fma(gy, gy, gz * gz)));
        nrm[0] = gx * inv; nrm[1] = gy * inv; nrm[2] = gz * inv;
    }
    if (ident) {
        hit_g[0] = h[0] + st.frame.o[0]; hit_g[1] = h[1] + st.frame.o[1]; hit_g[2] = h[2] + st.frame.o[2];
    } else {
        l2g_point(st.frame, h, hit_g);
    }
    bool ap_ok = true;
    if (PLAIN) {
        // no aperture
    } else if (st.aperture_kind == PYR_AP_CIRCULAR) {
        const double rr = fma(h[0], h[0], h[1] * h[1]);
        ap_ok = (rr >= st.ap0) && (rr <= st.ap1);
    } else if (st.aperture_kind == PYR_AP_RECTANGULAR) {
        ap_ok = (fabs(h[0]) <= st.ap0) && (fabs(h[1]) <= st.ap1);
    }
    const bool hit = ok && hit_ok && ap_ok;

    // Snell via in-plane k (material_isotropic.py:175-185 / :224)
    const double kn = dot3(kl, nrm);
    const double kin[3] = {fma(-kn, nrm[0], kl[0]), fma(-kn, nrm[1], kl[1]), fma(-kn, nrm[2], kl[2])};
    const double square2 = st.n2sq[w] - dot3(kin, kin);
    const double xi = fast_sqrt(square2);
    const bool alive = hit && (square2 > 0.0);
    double k2[3];
    if (!PLAIN && st.interaction == PYR_REFLECT) {
        k2[0] = fma(xi, nrm[0], -kin[0]); k2[1] = fma(xi, nrm[1], -kin[1]); k2[2] = fma(xi, nrm[2], -kin[2]);
    } else {
        k2[0] = fma(xi, nrm[0], kin[0]); k2[1] = fma(xi, nrm[1], kin[1]); k2[2] = fma(xi, nrm[2], kin[2]);
    }
    if (ident) { r.k[0] = k2[0]; r.k[1] = k2[1]; r.k[2] = k2[2]; }
    else rot(st.frame.r, k2, r.k);
    r.x[0] = hit_g[0]; r.x[1] = hit_g[1]; r.x[2] = hit_g[2];
    if (!alive) { const double q = qnan(); r.k[0] = q; r.k[1] = q; r.k[2] = q; }
    if (WITH_E) {
        if (alive) reproject_e(r.k, r.e);
        else r.e[0] = r.e[1] = r.e[2] = qnan();
    }
    r.alive = alive;
    return (hit ? PYR_RAY_HIT : 0u) | (alive ? PYR_RAY_ALIVE : 0u);
}

// FEAT: 0 = lean steps only; 1 = + explicit shapes / own-frame apertures / partial
// step modes; 3 = + GRIN media; 7 = + grid-sag / combination shapes; 9 = like 1 with the
// even asphere as the only explicit shape (separate instantiations keep the register
// budget and the code size of the common cases small)
template <int RPT, bool WITH_E, int FEAT, int MINB = 1, int POLICY = 0, int BLOCK = 256>
__global__ void __launch_bounds__(BLOCK, MINB)
trace_real_kernel(const __grid_constant__ LaunchParams P) {
    constexpr bool GENERAL = FEAT != 0;
    // POLICY bit 16: the rays are generated in registers from P.gen (csrc/pyr_gen.cuh): no
    // input rows are read and no input stage exists
    constexpr bool GEN = (POLICY & 16) != 0;
    // FEAT bit 16: GRIN segments integrate the RPT rays of a thread together (no history)
    constexpr bool GRIN_N = (FEAT & 16) != 0;
    const int64_t n = P.n;
    // Step table: one cooperative copy from the parameter block into shared memory;
    // per-step constants are then broadcast LDS reads (short, fixed latency) instead of
    // register-indexed constant loads.
    __shared__ DStep sst[kMaxSteps];
    __shared__ double etab[(FEAT & 2) ? kExpTabSize : 1];          // 2^(j/512) of pyr_exp.cuh (GRIN profiles)
    if (FEAT & 2)
        for (int i = threadIdx.x; i < kExpTabSize; i += blockDim.x) etab[i] = kExp2Tab[i];
    {
        const uint64_t *src = reinterpret_cast<const uint64_t *>(P.steps);
        uint64_t *dst = reinterpret_cast<uint64_t *>(sst);
        const int words = P.n_steps * (int)(sizeof(DStep) / 8);
        for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const bool need_e0 = sst[0].dir_mode == PYR_DIR_POYNTING;
    const bool load_e = !GEN && (WITH_E || need_e0) && P.e != nullptr;

    // Input staging (aligned bundles): a CTA owns tiles of TILE consecutive rays; the
    // rows of the tile two iterations ahead are pulled into shared memory by the TMA
    // (cp.async.bulk, one elected thread, mbarrier completion), so the HBM read latency
    // never sits on any warp's critical path and costs no registers or scoreboards.
    // Bundles whose rows are not 16-byte aligned are read with plain coalesced loads.
    constexpr int TILE = BLOCK * RPT;
    // POLICY 3 (TMA record stores) stages the outputs in shared memory as well and
    // therefore keeps a single input stage (prefetch distance: one tile = 13 steps)
    constexpr bool TMA_OUT = (POLICY & 3) == 3 && RPT == 2;
    // POLICY bit 8: wavelength batch -- per-ray segment index selects the media indices
    // (own instantiations: the plain kernels keep their code)
    constexpr bool MULTI = (POLICY & 8) != 0;
    // POLICY bit 32: the spot sums of the LAST entry (analysis/ray_analysis.py:44-86: sum x, count,
    // sum x^2 of the rays that survive it) are accumulated in registers over the CTA's tiles and
    // reduced once at the end -- pyr_trace_spot without the second pass over the record
    constexpr bool SPOT = (POLICY & 32) != 0;
    double sacc[SPOT ? 7 : 1];
    if (SPOT) {
#pragma unroll
        for (int q = 0; q < 7; ++q) sacc[q] = 0.0;
    }
    constexpr int IN_STAGES = GEN ? 0 : (TMA_OUT ? 1 : 2);
    // record stages: x, k (and E) rows; E recording has room for one stage only
    constexpr int OUT_ROWS = WITH_E ? 9 : 6;
    constexpr int OUT_STAGES = WITH_E ? 1 : 2;
    extern __shared__ __align__(128) double stage_buf[];          // [IN_STAGES][9][TILE] (+ out)
    __shared__ __align__(8) unsigned long long full_bar[2];
    double *out_buf = stage_buf + (size_t)IN_STAGES * 9 * TILE;   // [OUT_STAGES][OUT_ROWS][TILE]
    unsigned char *out_fl = reinterpret_cast<unsigned char *>(out_buf + OUT_STAGES * OUT_ROWS * TILE);
    unsigned store_count = 0;
    const bool staged = !GEN && P.in_vec2 != 0;
    const int rows = load_e ? 9 : 6;
    auto issue_tile = [&](int64_t tile, int stage) {
        const int64_t t0 = tile * TILE;
        const int64_t cnt = (n - t0 < TILE) ? n - t0 : TILE;
        const unsigned bytes = (unsigned)(((cnt + 1) & ~(int64_t)1) * 8);
        mbar_expect_tx(&full_bar[stage], bytes * rows);
        double *dst = stage_buf + (size_t)stage * 9 * TILE;
        for (int c = 0; c < 3; ++c) {
            tma_load_1d(dst + c * TILE, P.x + c * P.ld_in + t0, bytes, &full_bar[stage]);
            tma_load_1d(dst + (3 + c) * TILE, P.k + c * P.ld_in + t0, bytes, &full_bar[stage]);
            if (load_e) tma_load_1d(dst + (6 + c) * TILE, P.e + c * P.ld_in + t0, bytes, &full_bar[stage]);
        }
    };
    // In-order tile hand-out (P.tile_ctr != nullptr: large launches of the lean kernels): the CTAs
    // take their tiles from an atomic cursor instead of the static cyclic schedule, so the set of
    // tiles in flight stays a compact window however the CTAs drift.  The record streams of a
    // memory-bound launch then keep their DRAM page locality: C2 1.33 -> 1.15 ms with the same
    // instruction stream, records bit-identical (profiles/r02_c2_dynamic_tiles.md).
    const bool DYN = IN_STAGES <= 1 && P.tile_ctr != nullptr;
    __shared__ long long dyn_tile[2];                    // DYN: current and next tile of the CTA
    if (DYN && threadIdx.x == 0) {
        dyn_tile[0] = (long long)atomicAdd(P.tile_ctr, 1ull);
        dyn_tile[1] = (long long)atomicAdd(P.tile_ctr, 1ull);
    }
    if (DYN) __syncthreads();
    if (staged && threadIdx.x == 0) {
        mbar_init(&full_bar[0], 1);
        mbar_init(&full_bar[1], 1);
        fence_mbar_init();
        const int64_t first = DYN ? dyn_tile[0] : (int64_t)blockIdx.x;
        if (first * TILE < n) issue_tile(first, 0);
        if (IN_STAGES == 2 && ((int64_t)blockIdx.x + gridDim.x) * TILE < n)
            issue_tile((int64_t)blockIdx.x + gridDim.x, 1);
    }
    __syncthreads();

    int it = 0;
    int64_t dyn_next = DYN ? dyn_tile[1] : 0;
    for (int64_t tile = DYN ? dyn_tile[0] : (int64_t)blockIdx.x; tile * TILE < n;
         tile = DYN ? dyn_next : tile + gridDim.x, dyn_next = DYN ? dyn_tile[it & 1] : 0, ++it) {
        const int64_t base = tile * TILE + (int64_t)threadIdx.x * RPT;
        const int stage = GEN ? 0 : it % (IN_STAGES > 0 ? IN_STAGES : 1);
        // WITH_E == false still needs E for the first segment's Poynting direction
        Ray<true> in[RPT];
        bool in_range[RPT];
#pragma unroll
        for (int j = 0; j < RPT; ++j) in_range[j] = base + j < n;
        if (GEN) {
#pragma unroll
            for (int j = 0; j < RPT; ++j)
                gen_ray(P.gen, in_range[j] ? base + j : 0, in[j].x, in[j].k, in[j].e);
        } else if (staged) {
            mbar_wait(&full_bar[stage], (it / (IN_STAGES > 0 ? IN_STAGES : 1)) & 1);
            const double *src = stage_buf + (size_t)stage * 9 * TILE + threadIdx.x * RPT;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                if (RPT == 2) {
                    const double2 vx = *reinterpret_cast<const double2 *>(src + c * TILE);
                    const double2 vk = *reinterpret_cast<const double2 *>(src + (3 + c) * TILE);
                    in[0].x[c] = vx.x; in[RPT - 1].x[c] = vx.y;
                    in[0].k[c] = vk.x; in[RPT - 1].k[c] = vk.y;
                    double2 ve = make_double2(c == 1 ? 1.0 : 0.0, c == 1 ? 1.0 : 0.0);
                    if (load_e) ve = *reinterpret_cast<const double2 *>(src + (6 + c) * TILE);
                    in[0].e[c] = ve.x; in[RPT - 1].e[c] = ve.y;
                } else {
                    in[0].x[c] = src[c * TILE];
                    in[0].k[c] = src[(3 + c) * TILE];
                    in[0].e[c] = load_e ? src[(6 + c) * TILE] : (c == 1 ? 1.0 : 0.0);
                }
            }
            __syncthreads();                       // every thread has drained this stage
            if (threadIdx.x == 0) {
                const int64_t next = DYN ? dyn_next : tile + IN_STAGES * (int64_t)gridDim.x;
                if (next * TILE < n) issue_tile(next, stage);
            }
        } else {
#pragma unroll
            for (int j = 0; j < RPT; ++j) {
                const int64_t i = in_range[j] ? base + j : 0;
                const int64_t ix = (P.n_x == n) ? i : i % P.n_x;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    in[j].x[c] = __ldcs(P.x + c * P.ld_in + ix);
                    in[j].k[c] = __ldcs(P.k + c * P.ld_in + i);
                    in[j].e[c] = load_e ? __ldcs(P.e + c * P.ld_in + i) : (c == 1 ? 1.0 : 0.0);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            const int64_t i = in_range[j] ? base + j : 0;
            const int64_t ix = (P.n_x == n) ? i : i % P.n_x;
            in[j].alive = in_range[j] && ((!GEN && P.alive) ? (P.alive[ix] & PYR_RAY_ALIVE) != 0 : true);
        }

        int wsel[RPT];
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            wsel[j] = 0;
            if (MULTI) {
#pragma unroll
                for (int q = 0; q < kMaxWaves - 1; ++q)
                    wsel[j] += (q < P.n_waves - 1 && base + j >= P.wave_end[q]) ? 1 : 0;
            }
        }

        Ray<WITH_E> ray[RPT];
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
#pragma unroll
            for (int c = 0; c < 3; ++c) { ray[j].x[c] = in[j].x[c]; ray[j].k[c] = in[j].k[c]; }
            if (WITH_E) {
#pragma unroll
                for (int c = 0; c < 3; ++c) ray[j].e[c] = in[j].e[c];
            }
            ray[j].alive = in[j].alive;
            if (!in[j].alive) { ray[j].k[0] = ray[j].k[1] = ray[j].k[2] = qnan(); }
        }

        for (int s = 0; s < P.n_steps; ++s) {
            const DStep &st = sst[s];
            double hit[RPT][3], dd[RPT][3];
            uint32_t fl[RPT];
#pragma unroll
            for (int j = 0; j < RPT; ++j) {
                // direction of energy transport (ray.py:136-152): Poynting vector of the
                // user's (k, E) on the first segment, k/|k| wherever E.k = 0 is guaranteed
                double *d = dd[j];
                if (st.dir_mode == PYR_DIR_POYNTING && (WITH_E || s == 0)) {
                    if (WITH_E) poynting_dir(ray[j].k, ray[j].e, d);
                    else poynting_dir(ray[j].k, in[j].e, d);
                } else {
                    const double ik = st.inv_knorm[MULTI ? wsel[j] : 0];
                    const double inv = (ik > 0.0) ? ik : fast_rsqrt(dot3(ray[j].k, ray[j].k));
                    d[0] = ray[j].k[0] * inv; d[1] = ray[j].k[1] * inv; d[2] = ray[j].k[2] * inv;
                }
            }
            if (GRIN_N && st.before_kind == PYR_MEDIUM_ISO_GRIN) {
                // GRIN segment: the RPT rays of the thread are integrated together
                // (interleaved dependency chains, csrc/pyr_grin.cuh grin_propagate_n)
                const DAux *ga = &P.aux[st.aux];
                double gx[RPT][3], gk[RPT][3];
                bool enter[RPT], gvalid[RPT];
#pragma unroll
                for (int j = 0; j < RPT; ++j) {
                    enter[j] = ray[j].alive && in_range[j];
#pragma unroll
                    for (int c = 0; c < 3; ++c) { gx[j][c] = ray[j].x[c]; gk[j][c] = ray[j].k[c]; }
                }
                grin_propagate_rays<(FEAT & 4) != 0, RPT>(ga->before, st.shape_kind, ga, st.curv, st.cc, gx, dd, gk,
                                                          enter, gvalid, etab);
#pragma unroll
                for (int j = 0; j < RPT; ++j) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) { ray[j].x[c] = gx[j][c]; ray[j].k[c] = gk[j][c]; }
                    ray[j].alive = ray[j].alive && gvalid[j];
                    const double inv = fast_rsqrt(dot3(ray[j].k, ray[j].k));
                    dd[j][0] = ray[j].k[0] * inv; dd[j][1] = ray[j].k[1] * inv; dd[j][2] = ray[j].k[2] * inv;
                }
            }
            // steps without an auxiliary record (conic shape, homogeneous isotropic
            // media, aperture in the shape frame) always take the tuned path
            if (GENERAL && st.aux >= 0) {
                int64_t ridx[RPT];
                int wj[RPT];
#pragma unroll
                for (int j = 0; j < RPT; ++j) { ridx[j] = in_range[j] ? base + j : -1; wj[j] = MULTI ? wsel[j] : 0; }
                step_real_n<RPT, WITH_E, (FEAT & 2) != 0 && !GRIN_N, (FEAT & 4) != 0, (FEAT & 8) != 0>(
                    P, st, ray, dd, hit, ridx, wj, etab, (FEAT & 2) != 0, fl);
            } else {
                if (st.bits & kPlain) {
#pragma unroll
                    for (int j = 0; j < RPT; ++j)
                        fl[j] = step_lean<WITH_E, true>(st, ray[j], dd[j], hit[j], MULTI ? wsel[j] : 0);
                } else {
#pragma unroll
                    for (int j = 0; j < RPT; ++j)
                        fl[j] = step_lean<WITH_E>(st, ray[j], dd[j], hit[j], MULTI ? wsel[j] : 0);
                }
            }

            if (SPOT && s == P.n_steps - 1) {
#pragma unroll
                for (int j = 0; j < RPT; ++j) {
                    if (fl[j] & PYR_RAY_ALIVE) {
                        const double a = hit[j][0] - P.spot_shift[0], b = hit[j][1] - P.spot_shift[1],
                                     c = hit[j][2] - P.spot_shift[2];
                        sacc[0] += a; sacc[1] += b; sacc[2] += c; sacc[3] += 1.0;
                        sacc[4] = fma(a, a, sacc[4]); sacc[5] = fma(b, b, sacc[5]); sacc[6] = fma(c, c, sacc[6]);
                    }
                }
            }
            // ---- record the step ----
            const int64_t ld = st.ld_out;
            if (!(st.bits & kRecorded)) continue;                 // no output pointer at all
            if (TMA_OUT && (st.bits & kOutVec2)) {
                // The CTA's slice of the record goes through shared memory and leaves
                // as TMA bulk stores issued by one thread: no per-thread global stores,
                // no store back-pressure on the warps that do the arithmetic.
                const int b = store_count % OUT_STAGES;
                ++store_count;
                if (threadIdx.x == 0) tma_store_wait_read<OUT_STAGES - 1>();   // stage b is drained
                __syncthreads();
                double *ob = out_buf + (size_t)b * OUT_ROWS * TILE + threadIdx.x * 2;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    *reinterpret_cast<double2 *>(ob + c * TILE) = make_double2(hit[0][c], hit[RPT - 1][c]);
                    *reinterpret_cast<double2 *>(ob + (3 + c) * TILE) =
                        make_double2(ray[0].k[c], ray[RPT - 1].k[c]);
                    if (WITH_E)
                        *reinterpret_cast<double2 *>(ob + (6 + c) * TILE) =
                            make_double2(ray[0].e[c], ray[RPT - 1].e[c]);
                }
                *reinterpret_cast<uchar2 *>(out_fl + b * TILE + threadIdx.x * 2) =
                    make_uchar2((unsigned char)fl[0], (unsigned char)fl[RPT - 1]);
                fence_proxy_async_smem();
                __syncthreads();
                if (threadIdx.x == 0) {
                    const int64_t t0 = tile * TILE;
                    const int64_t cnt = (n - t0 < TILE) ? n - t0 : TILE;
                    const unsigned bytes = (unsigned)(((cnt + 1) & ~(int64_t)1) * 8);
                    const double *sb = out_buf + (size_t)b * OUT_ROWS * TILE;
                    for (int c = 0; c < 3; ++c) {
                        if (st.out_x) tma_store_1d(st.out_x + c * ld + t0, sb + c * TILE, bytes);
                        if (st.out_k) tma_store_1d(st.out_k + c * ld + t0, sb + (3 + c) * TILE, bytes);
                        if (WITH_E && st.out_e) tma_store_1d(st.out_e + c * ld + t0, sb + (6 + c) * TILE, bytes);
                    }
                    if (st.out_flags)
                        tma_store_1d(st.out_flags + t0, out_fl + b * TILE, (unsigned)((cnt + 15) & ~(int64_t)15));
                    tma_store_commit();
                }
                continue;
            }
            const bool v2 = RPT == 2 && (st.bits & kOutVec2) && in_range[RPT - 1];
            if (st.out_x) {
                if (v2) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) store_stream2<POLICY>(st.out_x + c * ld + base, hit[0][c], hit[RPT - 1][c]);
                } else {
#pragma unroll
                    for (int j = 0; j < RPT; ++j)
                        if (in_range[j]) {
#pragma unroll
                            for (int c = 0; c < 3; ++c) store_stream<POLICY>(st.out_x + c * ld + base + j, hit[j][c]);
                        }
                }
            }
            if (st.out_k) {
                if (v2) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) store_stream2<POLICY>(st.out_k + c * ld + base, ray[0].k[c], ray[RPT - 1].k[c]);
                } else {
#pragma unroll
                    for (int j = 0; j < RPT; ++j)
                        if (in_range[j]) {
#pragma unroll
                            for (int c = 0; c < 3; ++c) store_stream<POLICY>(st.out_k + c * ld + base + j, ray[j].k[c]);
                        }
                }
            }
            if (WITH_E && st.out_e) {
                if (v2) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) store_stream2<POLICY>(st.out_e + c * ld + base, ray[0].e[c], ray[RPT - 1].e[c]);
                } else {
#pragma unroll
                    for (int j = 0; j < RPT; ++j)
                        if (in_range[j]) {
#pragma unroll
                            for (int c = 0; c < 3; ++c) store_stream<POLICY>(st.out_e + c * ld + base + j, ray[j].e[c]);
                        }
                }
            }
            if (st.out_flags) {
                if (v2) {
                    __stcs(reinterpret_cast<uchar2 *>(st.out_flags + base),
                           make_uchar2((unsigned char)fl[0], (unsigned char)fl[RPT - 1]));
                } else {
#pragma unroll
                    for (int j = 0; j < RPT; ++j)
                        if (in_range[j]) st.out_flags[base + j] = (uint8_t)fl[j];
                }
            }
        }
        if (DYN) {
            // slot it & 1 held this tile: it receives the tile after next
            __syncthreads();
            if (threadIdx.x == 0) dyn_tile[it & 1] = (long long)atomicAdd(P.tile_ctr, 1ull);
            __syncthreads();
        }
    }
    if (SPOT) {
        __shared__ double spot_sm[BLOCK / 32][7];
#pragma unroll
        for (int q = 0; q < 7; ++q) {
            double v = sacc[q];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
            if ((threadIdx.x & 31) == 0) spot_sm[threadIdx.x >> 5][q] = v;
        }
        __syncthreads();
        if (threadIdx.x < 7) {
            double v = 0.0;
            for (int w = 0; w < BLOCK / 32; ++w) v += spot_sm[w][threadIdx.x];
            atomicAdd(P.spot8 + threadIdx.x, v);
        }
    }
    if (TMA_OUT && threadIdx.x == 0) tma_store_wait_read<0>();      // smem must outlive the reads
}

// ---------------------------------------------------------------------------
// spot sums (analysis/ray_analysis.py:44-86): sum x, count, sum x^2
// ---------------------------------------------------------------------------
// VEC: rows 16-byte aligned (x aligned, ld even): two rays per thread and iteration with
// 128-bit loads, two iterations in flight -- the kernel is a pure HBM read stream
template <bool VEC>
__global__ void __launch_bounds__(256)
spot_sums_kernel(const double *__restrict__ x, int64_t ld, const uint8_t *__restrict__ flags,
                 uint32_t mask, int64_t n, double sx, double sy, double sz, double *out8) {
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    auto add = [&](double px, double py, double pz, bool on) {
        if (!on) return;
        const double a = px - sx, b = py - sy, c = pz - sz;
        acc[0] += a; acc[1] += b; acc[2] += c; acc[3] += 1.0;
        acc[4] = fma(a, a, acc[4]); acc[5] = fma(b, b, acc[5]); acc[6] = fma(c, c, acc[6]);
    };
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    if (VEC) {
        const int64_t pairs = n / 2;
        const double2 *x0 = reinterpret_cast<const double2 *>(x);
        const double2 *x1 = reinterpret_cast<const double2 *>(x + ld);
        const double2 *x2 = reinterpret_cast<const double2 *>(x + 2 * ld);
        const uchar2 *f2 = reinterpret_cast<const uchar2 *>(flags);
        int64_t p = tid;
        for (; p + nthreads < pairs; p += 2 * nthreads) {        // two independent pairs
            const int64_t q = p + nthreads;
            const double2 a0 = __ldcs(x0 + p), a1 = __ldcs(x1 + p), a2 = __ldcs(x2 + p);
            const double2 b0 = __ldcs(x0 + q), b1 = __ldcs(x1 + q), b2 = __ldcs(x2 + q);
            uchar2 fa = make_uchar2(255, 255), fb = fa;
            if (flags) { fa = f2[p]; fb = f2[q]; }
            add(a0.x, a1.x, a2.x, (fa.x & mask) != 0); add(a0.y, a1.y, a2.y, (fa.y & mask) != 0);
            add(b0.x, b1.x, b2.x, (fb.x & mask) != 0); add(b0.y, b1.y, b2.y, (fb.y & mask) != 0);
        }
        for (; p < pairs; p += nthreads) {
            const double2 a0 = __ldcs(x0 + p), a1 = __ldcs(x1 + p), a2 = __ldcs(x2 + p);
            uchar2 fa = make_uchar2(255, 255);
            if (flags) fa = f2[p];
            add(a0.x, a1.x, a2.x, (fa.x & mask) != 0); add(a0.y, a1.y, a2.y, (fa.y & mask) != 0);
        }
        if (tid == 0 && (n & 1)) {
            const int64_t i = n - 1;
            add(x[i], x[ld + i], x[2 * ld + i], !flags || (flags[i] & mask) != 0);
        }
    } else {
        for (int64_t i = tid; i < n; i += nthreads)
            add(x[i], x[ld + i], x[2 * ld + i], !flags || (flags[i] & mask) != 0);
    }
    __shared__ double sm[8][7];
#pragma unroll
    for (int q = 0; q < 7; ++q) {
        double v = acc[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5][q] = v;
    }
    __syncthreads();
    if (threadIdx.x < 7) {
        double v = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += sm[w][threadIdx.x];
        atomicAdd(out8 + threadIdx.x, v);
    }
}

// ---------------------------------------------------------------------------
// spot-diagram points (analysis/optical_system_analysis.py:283-303 get_spot): (x, y) of the
// selected rays in the frame of the last surface, compacted.  One atomic per CTA and tile
// on a device cursor (block-aggregated: ballot + popc per warp, warp totals scanned in
// shared memory), order = tile arrival order, stable inside a tile.  No host round trip:
// the count stays on the device.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
spot_points_kernel(const double *__restrict__ x, int64_t ld, const uint8_t *__restrict__ flags,
                   uint32_t mask, int64_t n, const __grid_constant__ DFrame frame, int to_local,
                   double *__restrict__ xy, int64_t ld_out, unsigned long long *cursor) {
    __shared__ unsigned warp_count[8];
    __shared__ unsigned long long tile_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t t0 = (int64_t)blockIdx.x * 256; t0 < n; t0 += (int64_t)gridDim.x * 256) {
        const int64_t i = t0 + threadIdx.x;
        bool on = i < n && (!flags || (flags[i] & mask) != 0);
        double px = 0.0, py = 0.0;
        if (on) {
            const double p[3] = {__ldcs(x + i), __ldcs(x + ld + i), __ldcs(x + 2 * ld + i)};
            if (to_local) {
                double q[3];
                g2l_point(frame, p, q);
                px = q[0]; py = q[1];
            } else {
                px = p[0]; py = p[1];
            }
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, on);
        if (lane == 0) warp_count[warp] = __popc(ballot);
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned total = 0;
            for (int w = 0; w < 8; ++w) { const unsigned c = warp_count[w]; warp_count[w] = total; total += c; }
            tile_base = total ? atomicAdd(cursor, (unsigned long long)total) : 0ull;
        }
        __syncthreads();
        if (on) {
            const int64_t pos = (int64_t)tile_base + warp_count[warp] + __popc(ballot & ((1u << lane) - 1u));
            if (pos < ld_out) {
                __stcs(xy + pos, px);
                __stcs(xy + ld_out + pos, py);
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// stand-alone bundle generation (collimated_bundle / divergent_bundle resident on the device)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
generate_bundle_kernel(const __grid_constant__ DGen g, int64_t n, double *__restrict__ x,
                       double *__restrict__ k, double *__restrict__ e, int64_t ld) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        double rx[3], rk[3], re[3];
        gen_ray_any(g, i, rx, rk, re);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (x) x[c * ld + i] = rx[c];
            if (k) k[c * ld + i] = rk[c];
            if (e) e[c * ld + i] = re[c];
        }
    }
}

// ---------------------------------------------------------------------------
// host side: pack the public step table into launch parameters
// ---------------------------------------------------------------------------
static void pack_frame(const PyrFrame &f, DFrame &d) {
    for (int i = 0; i < 9; ++i) d.r[i] = f.r[i];
    for (int i = 0; i < 3; ++i) d.o[i] = f.o[i];
}

static bool frame_equal(const PyrFrame &a, const PyrFrame &b) {
    return std::memcmp(&a, &b, sizeof(PyrFrame)) == 0;
}

static bool rot_is_identity(const PyrFrame &f) {
    const double id[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int i = 0; i < 9; ++i)
        if (f.r[i] != id[i]) return false;
    return true;
}

// frame mapping material-local -> shape-local:  x_s = Rs^T (Rm x_m + om - os)
static void compose_to_shape(const PyrFrame &m, const PyrFrame &s, DFrame &out) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double v = 0.0;
            for (int l = 0; l < 3; ++l) v += s.r[l * 3 + i] * m.r[l * 3 + j];
            out.r[i * 3 + j] = v;
        }
    for (int i = 0; i < 3; ++i) {
        double v = 0.0;
        for (int l = 0; l < 3; ++l) v += s.r[l * 3 + i] * (m.o[l] - s.o[l]);
        out.o[i] = v;
    }
}

static void pack_medium(const PyrMedium &m, const PyrFrame &shape, DMedium &d) {
    d.kind = m.kind; d.profile = m.grin_profile; d.boundary = m.grin_boundary;
    d.max_steps = m.grin_max_steps;
    d.n = m.n;
    compose_to_shape(m.frame, shape, d.to_shape);
    // eps (complex, material frame) -> shape frame: Q eps Q^T with the real rotation
    // Q = Rs^T Rm, so the whole deflection can run in the shape frame
    {
        const double *q = d.to_shape.r;
        for (int part = 0; part < 2; ++part) {
            double t[9], o[9];
            for (int i = 0; i < 9; ++i) t[i] = m.eps[2 * i + part];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) {
                    double v = 0.0;
                    for (int a = 0; a < 3; ++a)
                        for (int b = 0; b < 3; ++b) v += q[i * 3 + a] * t[a * 3 + b] * q[j * 3 + b];
                    o[i * 3 + j] = v;
                }
            for (int i = 0; i < 9; ++i) d.eps[2 * i + part] = o[i];
        }
    }
    for (int i = 0; i < PYR_MAX_GRIN_PARAMS; ++i) d.p[i] = m.grin_p[i];
    for (int i = 0; i < 4; ++i) d.b[i] = m.grin_b[i];
    d.ds = m.grin_ds; d.energy_tol = m.grin_energy_tol;
    pack_frame(m.frame, d.frame);
}

// ---- crystal tensors (host side, at pack time) ----
// eigen-decomposition of a real symmetric 3x3 matrix (cyclic Jacobi): w ascending, v columns
static void jacobi3(const double a_in[9], double w[3], double v[9]) {
    double a[9];
    for (int i = 0; i < 9; ++i) { a[i] = a_in[i]; v[i] = (i % 4 == 0) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 50; ++sweep) {
        const double off = std::fabs(a[1]) + std::fabs(a[2]) + std::fabs(a[5]);
        if (off < 1e-300) break;
        for (int pq = 0; pq < 3; ++pq) {
            const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
            const double apq = a[p * 3 + q];
            if (std::fabs(apq) < 1e-300) continue;
            const double theta = (a[q * 3 + q] - a[p * 3 + p]) / (2.0 * apq);
            const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
            const double c = 1.0 / std::sqrt(t * t + 1.0), sn = t * c;
            for (int k = 0; k < 3; ++k) {          // A <- A J
                const double akp = a[k * 3 + p], akq = a[k * 3 + q];
                a[k * 3 + p] = c * akp - sn * akq;
                a[k * 3 + q] = sn * akp + c * akq;
            }
            for (int k = 0; k < 3; ++k) {          // A <- J^T A
                const double apk = a[p * 3 + k], aqk = a[q * 3 + k];
                a[p * 3 + k] = c * apk - sn * aqk;
                a[q * 3 + k] = sn * apk + c * aqk;
            }
            for (int k = 0; k < 3; ++k) {
                const double vkp = v[k * 3 + p], vkq = v[k * 3 + q];
                v[k * 3 + p] = c * vkp - sn * vkq;
                v[k * 3 + q] = sn * vkp + c * vkq;
            }
        }
    }
    int idx[3] = {0, 1, 2};
    const double d[3] = {a[0], a[4], a[8]};
    for (int i = 0; i < 3; ++i)
        for (int j = i + 1; j < 3; ++j)
            if (d[idx[j]] < d[idx[i]]) { const int t = idx[i]; idx[i] = idx[j]; idx[j] = t; }
    double vv[9];
    for (int i = 0; i < 3; ++i) {
        w[i] = d[idx[i]];
        for (int k = 0; k < 3; ++k) vv[k * 3 + i] = v[k * 3 + idx[i]];
    }
    for (int i = 0; i < 9; ++i) v[i] = vv[i];
}

// Crystal payload of a deflecting medium: eps18 = complex 3x3 already rotated into the shape
// frame.  Real symmetric tensors with a twofold (or threefold) eigenvalue are uniaxial
// (isotropic): eps = eps_o 1 + (eps_e - eps_o) a a^T (oracle/pyrate_np.py
// uniaxial_decomposition); everything else is solved through the Fresnel quartic.
static void pack_crystal(const double eps18[18], DAux &a) {
    a.uniaxial = 0;
    a.eps_o = a.eps_e = 0.0;
    a.axis[0] = a.axis[1] = 0.0; a.axis[2] = 1.0;
    double re[9], scale = 0.0, imag = 0.0;
    for (int i = 0; i < 9; ++i) {
        re[i] = eps18[2 * i];
        scale = std::fmax(scale, std::fabs(re[i]));
        imag = std::fmax(imag, std::fabs(eps18[2 * i + 1]));
    }
    const double tol = 1e-12;
    if (!(scale > 0.0) || imag > tol * scale) return;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            if (std::fabs(re[i * 3 + j] - re[j * 3 + i]) > tol * scale) return;
    double w[3], v[9];
    jacobi3(re, w, v);
    const bool lo_pair = std::fabs(w[0] - w[1]) <= tol * scale;
    const bool hi_pair = std::fabs(w[2] - w[1]) <= tol * scale;
    if (lo_pair && hi_pair) {                      // isotropic tensor: both sheets coincide
        a.uniaxial = 1;
        a.eps_o = a.eps_e = (w[0] + w[1] + w[2]) / 3.0;
    } else if (lo_pair) {
        a.uniaxial = 1;
        a.eps_o = 0.5 * (w[0] + w[1]); a.eps_e = w[2];
        for (int k = 0; k < 3; ++k) a.axis[k] = v[k * 3 + 2];
    } else if (hi_pair) {
        a.uniaxial = 1;
        a.eps_o = 0.5 * (w[1] + w[2]); a.eps_e = w[0];
        for (int k = 0; k < 3; ++k) a.axis[k] = v[k * 3 + 0];
    }
}

static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// auxiliary records (DAux) a step needs in the launch: 0 for the lean case (conic shape,
// aperture in the shape frame, homogeneous isotropic media, full step), one per term of a
// combination, else 1
int step_aux_records(const PyrStep &u) {
    const bool ap_same = u.aperture_kind == PYR_AP_BASE || frame_equal(u.shape_frame, u.aperture_frame);
    if (u.shape_kind == PYR_SHAPE_COMBINATION) return u.n_terms > 0 ? u.n_terms : 1;
    const bool need = u.shape_kind != PYR_SHAPE_CONIC || !ap_same || u.mode != PYR_STEP_FULL ||
                      u.before.kind != PYR_MEDIUM_ISO_CONST || u.after.kind != PYR_MEDIUM_ISO_CONST;
    return need ? 1 : 0;
}

// PyrBundleGen -> DGen (validated)
int pack_gen(const PyrBundleGen *g, int64_t n_rays, DGen &d) {
    std::memset(&d, 0, sizeof(d));
    if (!g) return PYR_OK;
    if (g->raster < PYR_RASTER_HEXAPOLAR || g->raster > PYR_RASTER_CIRCULAR) return PYR_E_UNSUPPORTED;
    if (g->bundle != PYR_BUNDLE_COLLIMATED && g->bundle != PYR_BUNDLE_DIVERGENT) return PYR_E_UNSUPPORTED;
    if (g->param < 0 || g->first < 0 || n_rays < 0 || g->first + n_rays > g->total) return PYR_E_BADARG;
    if ((g->raster == PYR_RASTER_RECT || g->raster == PYR_RASTER_HEX) && (!g->rows || g->param < 1))
        return PYR_E_BADARG;
    if (g->raster == PYR_RASTER_CIRCULAR && g->param < 1) return PYR_E_BADARG;
    if (g->raster == PYR_RASTER_HEXAPOLAR && g->total != 1 + 3 * g->param * (g->param + 1)) return PYR_E_BADARG;
    if (g->raster == PYR_RASTER_CIRCULAR && g->total != g->param * g->param) return PYR_E_BADARG;
    d.raster = g->raster; d.bundle = g->bundle; d.flags = g->flags; d.on = 1;
    d.param = g->param; d.first = g->first;
    d.lin_start = g->lin_start; d.lin_step = g->lin_step; d.lin_stop = g->lin_stop;
    d.aux0 = g->aux[0]; d.aux1 = g->aux[1];
    if (g->raster == PYR_RASTER_HEXAPOLAR) d.aux0 = 1.0 / (double)(g->param > 0 ? g->param : 1);
    d.radius = g->radius; d.n_index = g->n_index;
    for (int i = 0; i < 3; ++i) { d.start[i] = g->start[i]; d.dir[i] = g->dir[i]; d.e[i] = g->e[i]; }
    d.rows = g->rows;
    return PYR_OK;
}

struct Packed {
    LaunchParams P;
    uint32_t shape_mask;   // bit k: a step of shape kind k is present
    bool general;
    bool any_aniso;
    bool extended;      // grid-sag / combination shapes present
};

static int pack(const PyrStep *steps, int32_t n_steps, const PyrRaysIn *rays, int64_t n_rays,
                uint32_t flags, Packed &out) {
    if (!steps || !rays || n_steps <= 0 || n_rays < 0) return PYR_E_BADARG;
    if (n_steps > kMaxSteps) return PYR_E_TOOLARGE;
    if (!rays->gen && (!rays->x || !rays->k)) return PYR_E_BADARG;
    LaunchParams &P = out.P;
    std::memset(&P, 0, sizeof(P));
    if (rays->gen) {
        const int rc = pack_gen(rays->gen, n_rays, P.gen);
        if (rc != PYR_OK) return rc;
    }
    P.x = rays->x; P.k = rays->k; P.e = rays->e; P.alive = rays->alive;
    P.ld_in = rays->ld > 0 ? rays->ld : n_rays;
    P.n = n_rays;
    P.n_x = rays->n_x > 0 ? rays->n_x : n_rays;
    P.n_steps = n_steps;
    P.flags = flags;
    if (rays->n_waves < 0 || rays->n_waves > kMaxWaves) return PYR_E_BADARG;
    P.n_waves = rays->n_waves > 1 ? rays->n_waves : 1;
    for (int w = 0; w < kMaxWaves; ++w) {
        P.wave_end[w] = (w < P.n_waves - 1) ? rays->wave_end[w] : n_rays;
        if (w < P.n_waves - 1 && (rays->wave_end[w] < (w ? rays->wave_end[w - 1] : 0) ||
                                  rays->wave_end[w] > n_rays))
            return PYR_E_BADARG;
    }
    if (rays->gen) { P.x = P.k = P.e = nullptr; P.alive = nullptr; P.n_x = n_rays; P.ld_in = n_rays; }
    P.in_vec2 = !rays->gen && (P.n_x == P.n) && (P.ld_in % 2 == 0) && aligned16(P.x) && aligned16(P.k) &&
                (!P.e || aligned16(P.e));
    out.general = false;
    out.any_aniso = false;
    out.extended = false;
    out.shape_mask = 0;
    int n_aux = 0;
    for (int s = 0; s < n_steps; ++s) {
        const PyrStep &u = steps[s];
        DStep &d = P.steps[s];
        if (u.shape_kind < PYR_SHAPE_CONIC || u.shape_kind > PYR_SHAPE_CYLINDER) return PYR_E_UNSUPPORTED;
        out.shape_mask |= 1u << u.shape_kind;
        const bool grid = u.shape_kind == PYR_SHAPE_GRIDSAG;
        const bool comb = u.shape_kind == PYR_SHAPE_COMBINATION;
        if (comb && (u.n_terms < 1 || u.n_terms > PYR_MAX_TERMS)) return PYR_E_BADARG;
        bool uses_grid = grid;
        if (comb) {
            int grids = 0;
            for (int t = 0; t < u.n_terms; ++t) {
                const PyrShapeTerm &tm = u.terms[t];
                if (tm.kind < PYR_SHAPE_ASPHERE || tm.kind > PYR_SHAPE_GRIDSAG) return PYR_E_UNSUPPORTED;
                if (tm.coeff_off < 0 || tm.coeff_len < 0 || tm.coeff_off + tm.coeff_len > PYR_MAX_COEFF ||
                    tm.n_coeff < 0)
                    return PYR_E_BADARG;
                if (tm.kind == PYR_SHAPE_BICONIC) {
                    if (tm.n_coeff > 16 || (tm.n_coeff > 0 && tm.coeff_len < 16 + tm.n_coeff)) return PYR_E_BADARG;
                } else if (tm.kind != PYR_SHAPE_GRIDSAG && tm.n_coeff > tm.coeff_len) {
                    return PYR_E_BADARG;
                }
                if (tm.kind == PYR_SHAPE_GRIDSAG) ++grids;
            }
            if (grids > 1) return PYR_E_UNSUPPORTED;
            uses_grid = grids == 1;
        }
        if (uses_grid && (!u.grid_tx || !u.grid_ty || !u.grid_c || u.grid_nx < 8 || u.grid_ny < 8))
            return PYR_E_BADARG;
        out.extended = out.extended || grid || comb;
        if (u.shape_kind == PYR_SHAPE_BICONIC && u.n_coeff > 16) return PYR_E_BADARG;
        if (u.aperture_kind < PYR_AP_BASE || u.aperture_kind > PYR_AP_RECTANGULAR) return PYR_E_UNSUPPORTED;
        // position-dependent media: catalogue profiles only; a user profile is integrated by the
        // caller's own kernel and may only appear as the deflecting medium with its per-ray index
        if (u.before.kind == PYR_MEDIUM_ISO_GRIN && u.mode != PYR_STEP_DEFLECT_ONLY &&
            u.before.grin_profile != PYR_GRIN_GAUSSIAN_XY && u.before.grin_profile != PYR_GRIN_POLY_RZ)
            return PYR_E_UNSUPPORTED;
        if (u.after.kind == PYR_MEDIUM_ISO_GRIN && u.mode != PYR_STEP_PROPAGATE_ONLY && !u.after_n_rays &&
            u.after.grin_profile != PYR_GRIN_GAUSSIAN_XY && u.after.grin_profile != PYR_GRIN_POLY_RZ)
            return PYR_E_UNSUPPORTED;
        if (u.n_coeff < 0 || u.n_coeff > PYR_MAX_COEFF) return PYR_E_BADARG;
        if (u.split && s != n_steps - 1 && !(flags & PYR_F_COMPLEX)) return PYR_E_BADARG;
        pack_frame(u.shape_frame, d.frame);
        d.curv = u.curv; d.cc = u.cc;
        if (u.aperture_kind == PYR_AP_CIRCULAR) {
            d.ap0 = u.aperture_p[0] * u.aperture_p[0];
            d.ap1 = u.aperture_p[1] * u.aperture_p[1];
        } else if (u.aperture_kind == PYR_AP_RECTANGULAR) {
            d.ap0 = 0.5 * u.aperture_p[0];
            d.ap1 = 0.5 * u.aperture_p[1];
        }
        const int n_waves = rays->n_waves > 1 ? rays->n_waves : 1;
        for (int w = 0; w < kMaxWaves; ++w) {
            const bool used = w < n_waves;
            const double na = (n_waves > 1 && used) ? u.after_n_w[w] : u.after.n;
            const double nb = (n_waves > 1 && used) ? u.before_n_w[w] : u.k_norm_hint;
            d.n2sq[w] = na * na;
            d.inv_knorm[w] = (u.dir_mode == PYR_DIR_K && u.k_norm_hint > 0.0 && nb > 0.0 &&
                              u.before.kind == PYR_MEDIUM_ISO_CONST) ? 1.0 / nb : 0.0;
        }
        if (n_waves > 1 && (u.before.kind != PYR_MEDIUM_ISO_CONST || u.after.kind != PYR_MEDIUM_ISO_CONST ||
                            u.split))
            return PYR_E_UNSUPPORTED;        // a wavelength batch needs homogeneous isotropic media
        d.out_x = u.out_x; d.out_k = u.out_k; d.out_e = u.out_e; d.out_flags = u.out_flags;
#ifdef PYR_TOOLS
        {   // measurement knob (tools/ builds only, `make tools`): PYR_DEBUG_RECORD_LAST=1 records
            // the last entry only, which times the arithmetic of a trace without its record stream
            static const bool last_only = [] { const char *e = std::getenv("PYR_DEBUG_RECORD_LAST");
                                               return e && std::atoi(e) != 0; }();
            if (last_only && s != n_steps - 1) { d.out_x = d.out_k = d.out_e = nullptr; d.out_flags = nullptr; }
        }
#endif
        d.ld_out = u.ld_out > 0 ? u.ld_out : n_rays;
        d.ld_out2 = u.ld_out2 > 0 ? u.ld_out2 : 2 * d.ld_out;
        d.shape_kind = (int8_t)u.shape_kind; d.aperture_kind = (int8_t)u.aperture_kind;
        d.interaction = (int8_t)u.interaction; d.dir_mode = (int8_t)u.dir_mode;
        d.before_kind = (int8_t)u.before.kind; d.after_kind = (int8_t)u.after.kind;
        uint32_t bits = 0;
        if (rot_is_identity(u.shape_frame)) bits |= kRotIdentity;
        const bool ap_same = u.aperture_kind == PYR_AP_BASE || frame_equal(u.shape_frame, u.aperture_frame);
        if (ap_same) bits |= kApSameFrame;
        if (u.split) bits |= kSplit;
        if (u.cc == 0.0) bits |= kSphere;
        if (u.curv == 0.0 && u.shape_kind == PYR_SHAPE_CONIC) bits |= kPlane;
        if (u.mode == PYR_STEP_PROPAGATE_ONLY) bits |= kNoDeflect;
        else if (u.mode == PYR_STEP_DEFLECT_ONLY) bits |= kNoIntersect;
        else if (u.mode != PYR_STEP_FULL) return PYR_E_BADARG;
        // rows padded to 16 elements: the vector / TMA record path may write up to the
        // next multiple of 2 doubles (16 flag bytes) past the last ray of a row
        if (d.out_x || d.out_k || d.out_e || d.out_flags) bits |= kRecorded;
        // the double-Gauss's kind of entry: nothing to decide inside the step (step_lean<.., PLAIN>)
        if ((bits & kRotIdentity) && u.aperture_kind == PYR_AP_BASE && u.interaction != PYR_REFLECT &&
            u.shape_kind == PYR_SHAPE_CONIC && u.cc == 0.0 && u.mode == PYR_STEP_FULL)
            bits |= kPlain;
        if ((d.ld_out % 16 == 0) && (!d.out_x || aligned16(d.out_x)) && (!d.out_k || aligned16(d.out_k)) &&
            (!d.out_e || aligned16(d.out_e)) &&
            (!d.out_flags || aligned16(d.out_flags)))
            bits |= kOutVec2;
        d.bits = bits;
        const bool aniso = u.before.kind == PYR_MEDIUM_ANISO || u.after.kind == PYR_MEDIUM_ANISO;
        out.any_aniso = out.any_aniso || aniso;
        const int n_rec = step_aux_records(u);
        d.aux = -1;
        if (n_rec > 0) {
            if (n_aux + n_rec > kMaxAux) return PYR_E_TOOLARGE;
            DAux &a = P.aux[n_aux];
            const bool bic = u.shape_kind == PYR_SHAPE_BICONIC;
            for (int i = 0; i < PYR_MAX_COEFF; ++i) {
                const bool used = bic ? (i % 16) < u.n_coeff : i < u.n_coeff;
                a.coeff[i] = used ? u.coeff[i] : 0.0;
                a.xpow[i] = u.xpow[i]; a.ypow[i] = u.ypow[i];
            }
            a.curv2 = u.curv2; a.cc2 = u.cc2;
            a.hist_x = u.grin_hist_x; a.hist_k = u.grin_hist_k; a.hist_valid = u.grin_hist_valid;
            a.hist_count = u.grin_hist_count; a.hist_rows = u.grin_hist_rows;
            a.normradius = u.normradius != 0.0 ? u.normradius : 1.0;
            a.newton_tol = u.newton_tol > 0.0 ? u.newton_tol : 1e-14;
            a.n_coeff = u.n_coeff;
            a.newton_maxit = u.newton_maxit > 0 ? u.newton_maxit : 30;
            pack_frame(u.aperture_frame, a.aperture_frame);
            pack_medium(u.before, u.shape_frame, a.before);
            pack_medium(u.after, u.shape_frame, a.after);
            if (u.after.kind == PYR_MEDIUM_ANISO) pack_crystal(a.after.eps, a);
            a.grid_tx = u.grid_tx; a.grid_ty = u.grid_ty; a.grid_c = u.grid_c;
            a.grid_nx = u.grid_nx; a.grid_ny = u.grid_ny;
            a.after_n_rays = u.after_n_rays;
            a.n_terms = 0;
            if (comb) {
                // one auxiliary record per term, consecutive; record 0 also carries the
                // step's frame / media payload filled in above
                for (int t = 0; t < u.n_terms; ++t) {
                    const PyrShapeTerm &tm = u.terms[t];
                    DAux &r = P.aux[n_aux + t];
                    for (int i = 0; i < PYR_MAX_COEFF; ++i) {
                        const bool used = i < tm.coeff_len;
                        r.coeff[i] = used ? u.coeff[tm.coeff_off + i] : 0.0;
                        r.xpow[i] = used ? u.xpow[tm.coeff_off + i] : 0;
                        r.ypow[i] = used ? u.ypow[tm.coeff_off + i] : 0;
                    }
                    r.n_coeff = tm.n_coeff;
                    r.curv2 = tm.curv2; r.cc2 = tm.cc2;
                    r.normradius = tm.normradius != 0.0 ? tm.normradius : 1.0;
                    r.grid_tx = u.grid_tx; r.grid_ty = u.grid_ty; r.grid_c = u.grid_c;
                    r.grid_nx = u.grid_nx; r.grid_ny = u.grid_ny;
                    r.term_kind = tm.kind; r.term_w = tm.weight;
                    r.term_dx = tm.dx; r.term_dy = tm.dy; r.term_dz = tm.dz;
                    r.term_curv = tm.curv; r.term_cc = tm.cc;
                }
                a.n_terms = u.n_terms;
            }
            d.aux = (int8_t)n_aux;
            n_aux += n_rec;
            out.general = true;
        }
    }
    // the crystal kernel carries the common shapes only
    if (out.extended && (flags & PYR_F_COMPLEX)) return PYR_E_UNSUPPORTED;
    return PYR_OK;
}

int pack_steps(const PyrStep *steps, int32_t n_steps, const PyrRaysIn *rays, int64_t n_rays,
               uint32_t flags, LaunchParams &P, bool &general, bool &any_aniso) {
    static thread_local Packed pk;
    const int rc = pack(steps, n_steps, rays, n_rays, flags, pk);
    if (rc != PYR_OK) return rc;
    std::memcpy(&P, &pk.P, sizeof(LaunchParams));
    general = pk.general;
    any_aniso = pk.any_aniso;
    return PYR_OK;
}

static int g_sm_count = 0;

int sm_count() {
    if (g_sm_count == 0) {
        int dev = 0, sms = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 148;
        g_sm_count = sms > 0 ? sms : 148;
    }
    return g_sm_count;
}

// one private memory pool per device for the tile cursors (created on first use, kept for the
// life of the process: it never holds more than a few bytes)
static cudaMemPool_t tile_cursor_pool() {
    static std::mutex mu;
    static cudaMemPool_t pools[64] = {};
    static bool tried[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (!tried[dev]) {
        tried[dev] = true;
        cudaMemPoolProps props;
        std::memset(&props, 0, sizeof(props));
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        if (cudaMemPoolCreate(&pools[dev], &props) == cudaSuccess) {
            unsigned long long keep = ~0ull;                     // never trim: the next launch reuses the bytes
            cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &keep);
        } else {
            pools[dev] = nullptr;
            cudaGetLastError();
        }
    }
    return pools[dev];
}

template <typename K>
static int launch(K kernel, LaunchParams &P, int rpt, cudaStream_t stream, bool tma_out = false,
                  bool with_e = false, int threads = 256, bool gen = false,
                  bool dynamic_tiles = false) {   // threads == kernel's BLOCK
    // dynamic shared memory: input stages of rpt x 9 doubles per thread (two, or one plus
    // two output stages of rpt x 6 doubles + flag bytes when records leave through the TMA)
    const size_t tile = (size_t)threads * rpt;
    const size_t out_stages = with_e ? 1 : 2, out_rows = with_e ? 9 : 6;
    const size_t in_stages = gen ? 0 : (tma_out ? 1 : 2);      // a generated bundle has no input stage
    const size_t smem = in_stages * tile * 9 * 8 +
                        (tma_out ? out_stages * out_rows * tile * 8 + out_stages * tile : 0) + 16;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem);
    if (e != cudaSuccess) return (int)e;
    if (per_sm < 1) per_sm = 1;
    const int64_t work = (P.n + (int64_t)threads * rpt - 1) / ((int64_t)threads * rpt);
    // persistent grid: a whole number of CTAs per SM (148 SMs on B200)
    int64_t grid = (int64_t)sm_count() * per_sm;
    if (work < grid) grid = work;
    if (grid < 1) return PYR_OK;
    // tile cursor of the in-order hand-out: 8 bytes from a private stream-ordered pool, zeroed,
    // used by this launch and released, all in stream order (re-entrant per stream; nothing
    // persists but the pool itself).  Any failure falls back to the static schedule.
    unsigned long long *ctr = nullptr;
    if (dynamic_tiles && work >= 8 * grid) {
        cudaMemPool_t pool = tile_cursor_pool();
        if (pool && cudaMallocFromPoolAsync(reinterpret_cast<void **>(&ctr), sizeof(unsigned long long), pool,
                                            stream) == cudaSuccess) {
            if (cudaMemsetAsync(ctr, 0, sizeof(unsigned long long), stream) != cudaSuccess) {
                cudaFreeAsync(ctr, stream);
                ctr = nullptr;
            }
        } else {
            ctr = nullptr;
        }
        if (!ctr) cudaGetLastError();
    }
    P.tile_ctr = ctr;
    kernel<<<(unsigned)grid, threads, smem, stream>>>(P);
    e = cudaGetLastError();
    if (ctr) cudaFreeAsync(ctr, stream);
    P.tile_ctr = nullptr;
    return e == cudaSuccess ? PYR_OK : (int)e;
}

int trace_complex(const PyrStep *steps, int32_t n_steps, const PyrRaysIn *rays, int64_t n_rays,
                  uint32_t flags, cudaStream_t stream);   // pyr_aniso.cu

// GRIN kernels: rays per thread and resident CTAs per SM (tools builds vary them).  Measured
// on C5 (12.5e6 rays, profiles/r02_grin.md): 1 ray x 2 CTAs 32.8 ms, 2 rays x 2 CTAs 35.0 ms
// (ptxas re-serialises the two chains under the 128-register cap), 1 x 3 34.9, 1 x 4 36.0,
// 2 x 1 40.9, 2 x 3 39.3 -- more warps or more rays per thread do not pay: the FP64 pipe is
// the limit, so the default is the variant with the fewest instructions
#ifndef PYR_GRIN_RPT
#define PYR_GRIN_RPT 1
#endif
#ifndef PYR_GRIN_MINB
#define PYR_GRIN_MINB 2
#endif
constexpr int kGrinPolicy = PYR_GRIN_RPT == 2 ? 3 : 0;       // TMA record stores need two rays per thread
// asphere-only kernel: CTA size and resident CTAs per SM (tools builds vary them)
#ifndef PYR_ASPH_BLOCK
#define PYR_ASPH_BLOCK 256
#endif
#ifndef PYR_ASPH_MINB
#define PYR_ASPH_MINB 2
#endif

// spot8 / spot_shift / fused: pyr_trace_spot -- where the launch is ONE lean kernel that records the
// last entry, the kernel accumulates the spot sums itself (*fused = true); otherwise nothing changes
static int trace_impl(const PyrStep *steps, int32_t n_steps, const PyrRaysIn *rays, int64_t n_rays,
                      uint32_t flags, cudaStream_t stream, double *spot8 = nullptr,
                      const double *spot_shift = nullptr, bool *fused = nullptr) {
    if (fused) *fused = false;
    if (flags & PYR_F_COMPLEX) return trace_complex(steps, n_steps, rays, n_rays, flags, stream);
    static thread_local Packed pk;
    int rc = pack(steps, n_steps, rays, n_rays, flags, pk);
    if (rc != PYR_OK) return rc;
    pk.P.spot8 = nullptr;
    // (only with a flag record on the last step: without one pyr_spot_sums counts EVERY ray)
    const bool want_spot = spot8 != nullptr && fused != nullptr && !pk.general && !pk.any_aniso && n_rays > 0 &&
                           !(flags & PYR_F_RECORD_E) && pk.P.n_waves <= 1 && steps[n_steps - 1].out_flags != nullptr;
    if (want_spot) {
        pk.P.spot8 = spot8;
        for (int c = 0; c < 3; ++c) pk.P.spot_shift[c] = spot_shift ? spot_shift[c] : 0.0;
    }
    if (pk.any_aniso) return PYR_E_UNSUPPORTED;      // needs PYR_F_COMPLEX
    if (n_rays == 0) return PYR_OK;
    bool with_e = (flags & PYR_F_RECORD_E) != 0;
    // a Poynting-direction step after the first needs E carried along
    for (int s = 1; s < n_steps; ++s) with_e = with_e || steps[s].dir_mode == PYR_DIR_POYNTING;
    // conics + even aspheres only: the specialised general kernel
    const bool asph_only = pk.general && !pk.extended &&
                           (pk.shape_mask & ~((1u << PYR_SHAPE_CONIC) | (1u << PYR_SHAPE_ASPHERE))) == 0;
    bool hist = false;                      // GRIN integrator history requested
    for (int s = 0; s < n_steps; ++s) hist = hist || steps[s].grin_hist_x || steps[s].grin_hist_count;
    if (pk.P.gen.on) {
        // generated bundle: own instantiations (no input stage) of the lean, the general and
        // the GRIN kernel; everything else is traced from arrays (pyr_generate_bundle)
        bool grin = false;
        for (int s = 0; s < n_steps; ++s)
            grin = grin || steps[s].before.kind == PYR_MEDIUM_ISO_GRIN || steps[s].after.kind == PYR_MEDIUM_ISO_GRIN;
        if (with_e || pk.extended || pk.P.n_waves > 1) return PYR_E_UNSUPPORTED;
        if (!pk.general && want_spot) {
            *fused = true;
            return launch(trace_real_kernel<2, false, 0, 2, 19 | 32>, pk.P, 2, stream, true, false, 256, true, true);
        }
        if (!pk.general) return launch(trace_real_kernel<2, false, 0, 2, 19>, pk.P, 2, stream, true, false, 256, true, true);
        if (grin && !hist)
            return launch(trace_real_kernel<PYR_GRIN_RPT, false, 19, PYR_GRIN_MINB, kGrinPolicy | 16>, pk.P,
                          PYR_GRIN_RPT, stream, PYR_GRIN_RPT == 2, false, 256, true);
        if (grin) return launch(trace_real_kernel<1, false, 3, 2, 16>, pk.P, 1, stream, false, false, 256, true);
        if (asph_only)
            return launch(trace_real_kernel<2, false, 9, PYR_ASPH_MINB, 19, PYR_ASPH_BLOCK>, pk.P, 2, stream, true, false,
                          PYR_ASPH_BLOCK, true);
        return launch(trace_real_kernel<2, false, 1, 2, 19>, pk.P, 2, stream, true, false, 256, true);
    }
    if (pk.P.n_waves > 1) {
        // wavelength batch: own instantiations of the two record-streaming kernels
        bool grin_or_ext = pk.extended;
        for (int s = 0; s < n_steps; ++s)
            grin_or_ext = grin_or_ext || steps[s].before.kind == PYR_MEDIUM_ISO_GRIN ||
                          steps[s].after.kind == PYR_MEDIUM_ISO_GRIN;
        if (with_e || grin_or_ext) return PYR_E_UNSUPPORTED;
        if (!pk.general) return launch(trace_real_kernel<2, false, 0, 2, 11>, pk.P, 2, stream, true, false, 256, false, true);
        return launch(trace_real_kernel<2, false, 1, 2, 11>, pk.P, 2, stream, true);
    }
    if (!pk.general) {
        if (with_e) return launch(trace_real_kernel<2, true, 0, 2, 3>, pk.P, 2, stream, true, true, 256, false, true);
#ifdef PYR_TOOLS
        // PYR_LEAN_VARIANT=50 selects per-thread STG records instead of the TMA record
        // path (A/B knob of tools/ builds only).  Other configurations that were measured
        // and rejected are listed in profiles/r01_final_kernels.md.
        static const int variant = [] {
            const char *e = std::getenv("PYR_LEAN_VARIANT");
            return e ? std::atoi(e) : 0;
        }();
        if (variant == 50) return launch(trace_real_kernel<2, false, 0, 2>, pk.P, 2, stream);
#endif
        if (want_spot && !with_e) {
            *fused = true;
            return launch(trace_real_kernel<2, false, 0, 2, 3 | 32>, pk.P, 2, stream, true, false, 256, false, true);
        }
        return launch(trace_real_kernel<2, false, 0, 2, 3>, pk.P, 2, stream, true, false, 256, false, true);
    }
    bool has_grin = false;
    for (int s = 0; s < n_steps; ++s)
        has_grin = has_grin || steps[s].before.kind == PYR_MEDIUM_ISO_GRIN ||
                   steps[s].after.kind == PYR_MEDIUM_ISO_GRIN;
    if (pk.extended)   // grid-sag / combination shapes: the all-features instantiation
        return with_e ? launch(trace_real_kernel<1, true, 7, 2>, pk.P, 1, stream)
                      : launch(trace_real_kernel<1, false, 7, 2>, pk.P, 1, stream);
    if (has_grin && !with_e && !hist)   // two rays per thread, integrated together (interleaved chains)
        return launch(trace_real_kernel<PYR_GRIN_RPT, false, 19, PYR_GRIN_MINB, kGrinPolicy>, pk.P, PYR_GRIN_RPT,
                      stream, PYR_GRIN_RPT == 2);
    if (has_grin)   // history / E recording: one ray per thread
        return with_e ? launch(trace_real_kernel<1, true, 3, 2>, pk.P, 1, stream)
                      : launch(trace_real_kernel<1, false, 3, 2>, pk.P, 1, stream);
    if (asph_only && !with_e)
        return launch(trace_real_kernel<2, false, 9, PYR_ASPH_MINB, 3, PYR_ASPH_BLOCK>, pk.P, 2, stream, true, false,
                      PYR_ASPH_BLOCK);
    return with_e ? launch(trace_real_kernel<2, true, 1, 2, 3>, pk.P, 2, stream, true, true)
                  : launch(trace_real_kernel<2, false, 1, 2, 3>, pk.P, 2, stream, true);
}

int trace_entry(const PyrStep *steps, int32_t n_steps, const PyrRaysIn *rays, int64_t n_rays,
                uint32_t flags, cudaStream_t stream) {
    return trace_impl(steps, n_steps, rays, n_rays, flags, stream);
}

}  // namespace pyr

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" {

int pyr_version(void) { return PYR_ABI_VERSION; }
int64_t pyr_sizeof_step(void) { return (int64_t)sizeof(PyrStep); }
int64_t pyr_sizeof_rays_in(void) { return (int64_t)sizeof(PyrRaysIn); }
int64_t pyr_sizeof_bundle_gen(void) { return (int64_t)sizeof(PyrBundleGen); }

int pyr_generate_bundle(const PyrBundleGen *gen, int64_t n_rays, double *x, double *k, double *e,
                        int64_t ld, void *stream) {
    if (!gen || n_rays < 0) return PYR_E_BADARG;
    pyr::DGen d;
    const int rc = pyr::pack_gen(gen, n_rays, d);
    if (rc != PYR_OK) return rc;
    if (n_rays == 0) return PYR_OK;
    if (ld <= 0) ld = n_rays;
    if (ld < n_rays) return PYR_E_BADARG;
    int64_t grid = (n_rays + 255) / 256;
    const int64_t cap = (int64_t)pyr::sm_count() * 8;
    if (grid > cap) grid = cap;
    pyr::generate_bundle_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(d, n_rays, x, k, e, ld);
    cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? PYR_OK : (int)err;
}

const char *pyr_strerror(int code) {
    switch (code) {
        case PYR_OK: return "ok";
        case PYR_E_BADARG: return "bad argument";
        case PYR_E_UNSUPPORTED: return "unsupported shape / medium / flag combination";
        case PYR_E_TOOLARGE: return "too many steps or auxiliary records for one launch";
        case PYR_E_NODEVICE: return "no CUDA device";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "unknown error";
}

int pyr_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int pyr_trace(const PyrStep *steps, int32_t n_steps, const PyrRaysIn *rays, int64_t n_rays,
              uint32_t flags, void *stream) {
    if (n_rays == 0 && steps && rays && n_steps > 0) return PYR_OK;
    return pyr::trace_impl(steps, n_steps, rays, n_rays, flags, (cudaStream_t)stream);
}

int pyr_spot_sums(const double *x, int64_t ld, const uint8_t *flags, uint32_t mask, int64_t n,
                  const double *shift, double *out8, void *stream) {
    if (!x || !out8 || n < 0) return PYR_E_BADARG;
    if (n == 0) return PYR_OK;
    if (ld <= 0) ld = n;
    int64_t grid = (n + 255) / 256;
    const int64_t cap = (int64_t)pyr::sm_count() * 8;
    if (grid > cap) grid = cap;
    const double sx = shift ? shift[0] : 0.0, sy = shift ? shift[1] : 0.0, sz = shift ? shift[2] : 0.0;
    const bool vec = (reinterpret_cast<uintptr_t>(x) % 16 == 0) && (ld % 2 == 0) &&
                     (!flags || reinterpret_cast<uintptr_t>(flags) % 2 == 0);
    if (vec)
        pyr::spot_sums_kernel<true><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(x, ld, flags, mask, n,
                                                                                    sx, sy, sz, out8);
    else
        pyr::spot_sums_kernel<false><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(x, ld, flags, mask, n,
                                                                                     sx, sy, sz, out8);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? PYR_OK : (int)e;
}

int pyr_trace_spot(const PyrStep *steps, int32_t n_steps, const PyrRaysIn *rays, int64_t n_rays,
                   uint32_t flags, const double *shift, double *spot8_dev, double *spot8_host,
                   void *stream) {
    if (!steps || n_steps <= 0 || !rays || !spot8_dev || n_rays < 0) return PYR_E_BADARG;
    const PyrStep &last = steps[n_steps - 1];
    if (!last.out_x || last.split) return PYR_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(spot8_dev, 0, 64, st);
    if (e != cudaSuccess) return (int)e;
    bool fused = false;
    int rc = PYR_OK;
    if (n_rays > 0) rc = pyr::trace_impl(steps, n_steps, rays, n_rays, flags, st, spot8_dev, shift, &fused);
    if (rc != PYR_OK) return rc;
    if (!fused) {
        rc = pyr_spot_sums(last.out_x, last.ld_out > 0 ? last.ld_out : n_rays, last.out_flags, PYR_RAY_ALIVE,
                           n_rays, shift, spot8_dev, stream);
        if (rc != PYR_OK) return rc;
    }
    if (!spot8_host) return PYR_OK;              // asynchronous form: the sums stay on the device
    e = cudaMemcpyAsync(spot8_host, spot8_dev, 64, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    return e == cudaSuccess ? PYR_OK : (int)e;
}

int pyr_spot_points(const double *x, int64_t ld, const uint8_t *flags, uint32_t mask, int64_t n,
                    const PyrFrame *frame, double *xy, int64_t ld_out, int64_t *count, void *stream) {
    if (!x || !xy || !count || n < 0 || ld_out < 0) return PYR_E_BADARG;
    if (ld <= 0) ld = n;
    cudaError_t e = cudaMemsetAsync(count, 0, sizeof(int64_t), (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
    if (n == 0) return PYR_OK;
    pyr::DFrame f;
    std::memset(&f, 0, sizeof(f));
    if (frame) {
        for (int i = 0; i < 9; ++i) f.r[i] = frame->r[i];
        for (int i = 0; i < 3; ++i) f.o[i] = frame->o[i];
    }
    int64_t grid = (n + 255) / 256;
    const int64_t cap = (int64_t)pyr::sm_count() * 8;
    if (grid > cap) grid = cap;
    pyr::spot_points_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
        x, ld, flags, mask, n, f, frame ? 1 : 0, xy, ld_out, reinterpret_cast<unsigned long long *>(count));
    e = cudaGetLastError();
    return e == cudaSuccess ? PYR_OK : (int)e;
}

}  // extern "C"
