// Reference-exact GRIN propagation for SMALL bundles (opt-in): the lock-step loop of
// IsotropicGrinMaterial.symplecticintegrator (raytracer/material/material_grin.py:106-213)
// with its two cross-ray couplings, which the fused trace kernels deliberately normalise
// away:
//   * every ray keeps stepping until ALL rays are final (:139); a ray that is already final
//     is still moved, and if its over-stepped position leaves the boundary it becomes
//     invalid after the fact (:189-190);
//   * the energy test is BUNDLE-SUMMED: sum(v^2) - sum(n^2) over all rays against
//     annotations["energyviolation"]; a violation invalidates every ray (:164-176).
// One CTA walks the whole bundle (rays strided over 1024 threads, state in a caller-owned
// scratch buffer), two block reductions per integrator step.  Cost O(steps x rays / 1024):
// meant for the bundle sizes at which the reference itself is usable (<= 1e5 rays).
#include <cuda_runtime.h>

#include <cstring>

#include "pyr_device.cuh"
#include "pyr_grin.cuh"

namespace pyr {

int pack_steps(const PyrStep *steps, int32_t n_steps, const PyrRaysIn *rays, int64_t n_rays,
               uint32_t flags, LaunchParams &P, bool &general, bool &any_aniso);   // pyr_trace.cu

constexpr int kLockThreads = 1024;
constexpr uint8_t kExists = 1, kValid = 2, kFinal = 4, kCrossed = 8, kInside = 16;

__device__ __forceinline__ double block_sum(double v, double *sm) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < kLockThreads / 32; ++w) t += sm[w];      // same order in every thread
    return t;
}

__global__ void __launch_bounds__(kLockThreads)
grin_lockstep_kernel(const __grid_constant__ LaunchParams P, const double *__restrict__ x,
                     const double *__restrict__ k, const double *__restrict__ e,
                     const uint8_t *__restrict__ alive, int64_t ld, int64_t n, double *out_x,
                     double *out_k, uint8_t *out_alive, double *scratch, int32_t *iterations,
                     double *hist_x, double *hist_k, uint8_t *hist_valid, int64_t hist_rows) {
    const DStep &st = P.steps[0];
    const DAux *ax = &P.aux[st.aux];
    const DMedium &m = ax->before;
    __shared__ double etab[kExpTabSize];
    __shared__ double red[kLockThreads / 32];
    __shared__ int all_invalid;
    for (int i = threadIdx.x; i < kExpTabSize; i += kLockThreads) etab[i] = kExp2Tab[i];
    if (threadIdx.x == 0) all_invalid = 0;
    double *pos = scratch, *vel = scratch + 3 * ld, *upos = scratch + 6 * ld, *uvel = scratch + 9 * ld;
    uint8_t *flag = reinterpret_cast<uint8_t *>(scratch + 12 * ld);
    __syncthreads();

    // start state (:112-127): position and n * direction in the material frame
    for (int64_t i = threadIdx.x; i < n; i += kLockThreads) {
        const bool exists = alive ? (alive[i] & PYR_RAY_ALIVE) != 0 : true;
        double xg[3], kg[3], d[3], q[3], p[3], g[3];
        for (int c = 0; c < 3; ++c) { xg[c] = x[c * ld + i]; kg[c] = k[c * ld + i]; }
        if (st.dir_mode == PYR_DIR_POYNTING && e) {
            // ray.py:140-152 for real k, E
            double ev[3];
            for (int c = 0; c < 3; ++c) ev[c] = e[c * ld + i];
            const double ee = dot3(ev, ev), ek = dot3(ev, kg);
            double s[3] = {ee * kg[0] - ek * ev[0], ee * kg[1] - ek * ev[1], ee * kg[2] - ek * ev[2]};
            const double inv = 1.0 / sqrt(dot3(s, s));
            for (int c = 0; c < 3; ++c) d[c] = s[c] * inv;
        } else {
            const double inv = 1.0 / sqrt(dot3(kg, kg));
            for (int c = 0; c < 3; ++c) d[c] = kg[c] * inv;
        }
        g2l_point(m.frame, xg, q);
        rot_t(m.frame.r, d, p);
        const double n0 = grin_index(m, q, g, false, etab);
        for (int c = 0; c < 3; ++c) {
            pos[c * ld + i] = q[c]; upos[c * ld + i] = q[c];
            vel[c * ld + i] = n0 * p[c]; uvel[c * ld + i] = n0 * p[c];
        }
        flag[i] = exists ? (kExists | kValid) : 0;
    }
    __syncthreads();

    const double c0 = 1.0 / (2.0 * (2.0 - 1.2599210498948732));
    const double c1 = (1.0 - 1.2599210498948732) / (2.0 * (2.0 - 1.2599210498948732));
    const double d0 = 1.0 / (2.0 - 1.2599210498948732);
    const double d1 = -1.2599210498948732 / (2.0 - 1.2599210498948732);
    const double cs[4] = {c0, c1, c1, c0};
    const double ds[4] = {d0, d1, d0, 0.0};
    const double tau2 = 2.0 * m.ds;
    const int cap = m.max_steps > 0 ? m.max_steps : 1000000;
    int it = 0;
    for (; it < cap; ++it) {
        // ---- one symplectic step of EVERY ray (:144-158), partial energy sums ----
        double sp = 0.0, sn = 0.0;
        for (int64_t i = threadIdx.x; i < n; i += kLockThreads) {
            uint8_t f = flag[i];
            if (!(f & kExists)) continue;
            double q[3], p[3], g[3], nq = 0.0;
            for (int c = 0; c < 3; ++c) { q[c] = pos[c * ld + i]; p[c] = vel[c * ld + i]; }
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                for (int c = 0; c < 3; ++c) q[c] = fma(tau2 * cs[s], p[c], q[c]);
                nq = grin_index(m, q, g, true, etab);
                const double fk = tau2 * ds[s] * nq;
                for (int c = 0; c < 3; ++c) p[c] = fma(fk, g[c], p[c]);
            }
            for (int c = 0; c < 3; ++c) { pos[c * ld + i] = q[c]; vel[c * ld + i] = p[c]; }
            sp += dot3(p, p);
            sn += nq * nq;
            double xs[3];
            l2g_point(m.to_shape, q, xs);
            const bool crossed = xs[2] - shape_sag<false>(st.shape_kind, ax, st.curv, st.cc, xs[0], xs[1]) > 0.0;
            f = (uint8_t)((f & (kExists | kValid)) | (crossed ? kCrossed : 0) | (grin_inside(m, q) ? kInside : 0));
            flag[i] = f;
        }
        // ---- bundle-summed energy test (:164-176): a violation invalidates ALL rays ----
        const double tp = block_sum(sp, red);
        const double tn = block_sum(sn, red);
        if (threadIdx.x == 0 && fabs(tp - tn) > m.energy_tol) all_invalid = 1;
        __syncthreads();
        const bool inval = all_invalid != 0;
        // ---- validity, finality, frozen state (:181-196), history row (:198-205) ----
        double nonfinal = 0.0;
        for (int64_t i = threadIdx.x; i < n; i += kLockThreads) {
            uint8_t f = flag[i];
            if (!(f & kExists)) continue;
            const bool valid = (f & kValid) && !inval && (f & kInside);
            const bool final = (f & kCrossed) || !valid;
            if (!final) {
                for (int c = 0; c < 3; ++c) { upos[c * ld + i] = pos[c * ld + i]; uvel[c * ld + i] = vel[c * ld + i]; }
                nonfinal += 1.0;
            }
            flag[i] = (uint8_t)(kExists | (valid ? kValid : 0) | (final ? kFinal : 0));
            if (hist_x && it < hist_rows) {
                double uq[3], up[3], g[3], gq[3], kq[3], kgl[3];
                for (int c = 0; c < 3; ++c) { uq[c] = upos[c * ld + i]; up[c] = uvel[c * ld + i]; }
                const double invn = 1.0 / grin_index(m, uq, g, false, etab);
                for (int c = 0; c < 3; ++c) kq[c] = up[c] * invn;
                l2g_point(m.frame, uq, gq);
                rot(m.frame.r, kq, kgl);
                for (int c = 0; c < 3; ++c) {
                    hist_x[((int64_t)it * 3 + c) * ld + i] = gq[c];
                    if (hist_k) hist_k[((int64_t)it * 3 + c) * ld + i] = kgl[c];
                }
                if (hist_valid) hist_valid[(int64_t)it * ld + i] = valid ? 1 : 0;
            }
        }
        const double left = block_sum(nonfinal, red);
        if (left == 0.0) { ++it; break; }              // while not all(final)  (:139)
    }
    if (threadIdx.x == 0 && iterations) *iterations = it;
    // ---- frozen state back to the global frame; k = v / n (:198-199) ----
    for (int64_t i = threadIdx.x; i < n; i += kLockThreads) {
        const uint8_t f = flag[i];
        double uq[3], up[3], g[3], gq[3], kq[3], kgl[3];
        for (int c = 0; c < 3; ++c) { uq[c] = upos[c * ld + i]; up[c] = uvel[c * ld + i]; }
        const double invn = 1.0 / grin_index(m, uq, g, false, etab);
        for (int c = 0; c < 3; ++c) kq[c] = up[c] * invn;
        l2g_point(m.frame, uq, gq);
        rot(m.frame.r, kq, kgl);
        const bool ok = (f & kExists) && (f & kValid);
        for (int c = 0; c < 3; ++c) {
            out_x[c * ld + i] = (f & kExists) ? gq[c] : qnan();
            out_k[c * ld + i] = (f & kExists) ? kgl[c] : qnan();
        }
        out_alive[i] = ok ? PYR_RAY_ALIVE : 0;
    }
}

}  // namespace pyr

extern "C" {

int64_t pyr_grin_lockstep_scratch(int64_t ld) { return ld > 0 ? 12 * ld * 8 + ((ld + 255) / 256) * 256 : 0; }

int pyr_grin_lockstep(const PyrStep *step, const double *x, const double *k, const double *e,
                      const uint8_t *alive, int64_t ld, int64_t n, double *out_x, double *out_k,
                      uint8_t *out_alive, void *scratch, int32_t *iterations, double *hist_x,
                      double *hist_k, uint8_t *hist_valid, int64_t hist_rows, void *stream) {
    if (!step || !x || !k || !out_x || !out_k || !out_alive || !scratch || n < 0 || ld < n) return PYR_E_BADARG;
    if (step->before.kind != PYR_MEDIUM_ISO_GRIN) return PYR_E_BADARG;
    if (step->shape_kind == PYR_SHAPE_GRIDSAG || step->shape_kind == PYR_SHAPE_COMBINATION) return PYR_E_UNSUPPORTED;
    static thread_local pyr::LaunchParams P;
    PyrRaysIn rin;
    std::memset(&rin, 0, sizeof(rin));
    rin.x = x; rin.k = k; rin.e = e; rin.ld = ld; rin.n_x = n;
    bool general = false, aniso = false;
    PyrStep one = *step;
    one.out_x = one.out_k = one.out_e = nullptr;
    one.out_flags = nullptr;
    one.split = 0;
    const int rc = pyr::pack_steps(&one, 1, &rin, n, 0u, P, general, aniso);
    if (rc != PYR_OK) return rc;
    if (aniso) return PYR_E_UNSUPPORTED;
    if (n == 0) return PYR_OK;
    pyr::grin_lockstep_kernel<<<1, pyr::kLockThreads, 0, (cudaStream_t)stream>>>(
        P, x, k, e, alive, ld, n, out_x, out_k, out_alive, static_cast<double *>(scratch), iterations, hist_x,
        hist_k, hist_valid, hist_rows);
    const cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? PYR_OK : (int)err;
}

}  // extern "C"
