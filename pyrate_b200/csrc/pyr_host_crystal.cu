// Host-buffer entry for sequences with birefringent media (the complex-valued part of
// pyr_trace_host_io): real start points / wave vectors / fields in host memory in, the LAST
// record of the doubled bundle back in host memory -- x (3, n m_x), k and E (3, n m_k)
// complex128, flags (n m_x), m_x / m_k = 2^(doubling steps ahead of / up to the last entry)
// -- in the reference's hstack order (material_anisotropic.py:89-100: mode a of column c
// stays in c, mode b goes to w + c, w = width of the level).  The bundle is cut into chunks
// like in pyr_host.cu; a chunk's doubled record is a set of m column blocks, block b of
// chunk [off, off + cn) lands at host columns b n + off.
#include <cuda_runtime.h>

#include <cstring>
#include <vector>

#include "pyr_device.cuh"

namespace pyr {
int trace_entry(const PyrStep *steps, int32_t n_steps, const PyrRaysIn *rays, int64_t n_rays,
                uint32_t flags, cudaStream_t stream);
int step_aux_records(const PyrStep &u);

__global__ void fill_rows_kernel(double *p, int64_t ld, int64_t n, double v0, double v1, double v2) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        p[i] = v0; p[ld + i] = v1; p[2 * ld + i] = v2;
    }
}
}  // namespace pyr

namespace {
int64_t up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

struct Plan {
    int first_aniso = -1;
    std::vector<int64_t> mult_in, mult_out;     // width multipliers of every step
    int64_t mult_x = 1, mult_k = 1;
};

int make_plan(const PyrStep *steps, int32_t n_steps, Plan &pl) {
    pl.mult_in.assign(n_steps, 1);
    pl.mult_out.assign(n_steps, 1);
    int64_t m = 1;
    int aux = 0, splits = 0;
    for (int s = 0; s < n_steps; ++s) {
        const bool an = steps[s].before.kind == PYR_MEDIUM_ANISO || steps[s].after.kind == PYR_MEDIUM_ANISO;
        if (an && pl.first_aniso < 0) pl.first_aniso = s;
        if (steps[s].before.kind == PYR_MEDIUM_ISO_GRIN || steps[s].after.kind == PYR_MEDIUM_ISO_GRIN)
            return PYR_E_UNSUPPORTED;
        pl.mult_in[s] = m;
        const bool split = steps[s].split && steps[s].after.kind == PYR_MEDIUM_ANISO;
        if (split) { m *= 2; ++splits; }
        pl.mult_out[s] = m;
        if (pl.first_aniso >= 0) aux += pyr::step_aux_records(steps[s]);
    }
    if (pl.first_aniso < 0) return PYR_E_BADARG;
    if (n_steps - pl.first_aniso > pyr::kMaxSteps || pl.first_aniso > pyr::kMaxSteps || aux > pyr::kMaxAux ||
        splits > 6)
        return PYR_E_TOOLARGE;
    pl.mult_x = pl.mult_in[n_steps - 1];
    pl.mult_k = pl.mult_out[n_steps - 1];
    return PYR_OK;
}

// bytes of one chunk slot
int64_t slot_bytes(const PyrStep *steps, int32_t n_steps, const Plan &pl, int64_t chunk) {
    const int64_t ld = up(chunk, 32);
    int64_t b = 9 * ld * 8;                                   // x0, k0, e0
    b += 9 * ld * 8 + up(ld, 256);                            // last record of the real prefix (x, k, e, flags)
    b += 2 * 3 * ld * 16;                                     // promoted k, e
    for (int s = pl.first_aniso; s < n_steps; ++s) {
        const bool need = (steps[s].split != 0) || s == n_steps - 1;
        if (!need) continue;
        const int64_t lw = up(chunk * pl.mult_in[s], 32), lw2 = up(chunk * pl.mult_out[s], 32);
        b += 3 * lw * 8 + up(lw, 256) + 2 * 3 * lw2 * 16;
    }
    return b + 4096;
}
}  // namespace

extern "C" int64_t pyr_trace_host_crystal_workspace(const PyrStep *steps, int32_t n_steps, int64_t chunk_rays,
                                                     int64_t *mult_x, int64_t *mult_k) {
    Plan pl;
    if (!steps || n_steps <= 0 || chunk_rays <= 0 || make_plan(steps, n_steps, pl) != PYR_OK) return 0;
    if (mult_x) *mult_x = pl.mult_x;
    if (mult_k) *mult_k = pl.mult_k;
    return 4 * slot_bytes(steps, n_steps, pl, chunk_rays) + 256;
}

// called by pyr_trace_host_io (pyr_host.cu) when the sequence holds anisotropic media
int pyr_trace_host_crystal(const PyrStep *steps, int32_t n_steps, const PyrHostIO *io, int64_t n_rays,
                           void *workspace, int64_t workspace_bytes, int64_t chunk_rays, cudaStream_t *st) {
    Plan pl;
    int rc = make_plan(steps, n_steps, pl);
    if (rc != PYR_OK) return rc;
    if (io->x_all || io->k_all || io->flags_all) return PYR_E_UNSUPPORTED;
    if (workspace_bytes < pyr_trace_host_crystal_workspace(steps, n_steps, chunk_rays, nullptr, nullptr))
        return PYR_E_BADARG;
    const int64_t ld = up(chunk_rays, 32);
    const int64_t sb = slot_bytes(steps, n_steps, pl, chunk_rays);
    char *ws = static_cast<char *>(workspace);
    double *spot_dev = reinterpret_cast<double *>(ws + 4 * sb);
    cudaError_t ce = cudaSuccess;
    if (io->spot8) {
        ce = cudaMemsetAsync(spot_dev, 0, 64, st[0]);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(st[0]);
        if (ce != cudaSuccess) return (int)ce;
    }
    std::vector<PyrStep> local(steps, steps + n_steps);
    const int fa = pl.first_aniso;
    const int64_t hx = n_rays * pl.mult_x, hk = n_rays * pl.mult_k;      // host leading dimensions
    int64_t chunk_idx = 0;
    for (int64_t off = 0; off < n_rays && rc == PYR_OK; off += chunk_rays, ++chunk_idx) {
        const int64_t cn = n_rays - off < chunk_rays ? n_rays - off : chunk_rays;
        cudaStream_t s = st[chunk_idx % 4];
        char *p = ws + (chunk_idx % 4) * sb;
        auto take = [&](int64_t bytes) { char *q = p; p += up(bytes, 256); return q; };
        double *dx = reinterpret_cast<double *>(take(3 * ld * 8));
        double *dk = reinterpret_cast<double *>(take(3 * ld * 8));
        double *de = reinterpret_cast<double *>(take(3 * ld * 8));
        double *px = reinterpret_cast<double *>(take(3 * ld * 8));
        double *pk = reinterpret_cast<double *>(take(3 * ld * 8));
        double *pe = reinterpret_cast<double *>(take(3 * ld * 8));
        uint8_t *pf = reinterpret_cast<uint8_t *>(take(ld));
        double *kc = reinterpret_cast<double *>(take(3 * ld * 16));
        double *ec = reinterpret_cast<double *>(take(3 * ld * 16));
        // ---- inputs ----
        if (io->gen) {
            PyrBundleGen g = *io->gen;
            g.first += off;
            rc = pyr_generate_bundle(&g, cn, dx, dk, de, ld, s);
            if (rc != PYR_OK) break;
        } else {
            ce = cudaMemcpy2DAsync(dx, ld * 8, io->x0 + off, n_rays * 8, cn * 8, 3, cudaMemcpyHostToDevice, s);
            if (ce == cudaSuccess)
                ce = cudaMemcpy2DAsync(dk, ld * 8, io->k0 + off, n_rays * 8, cn * 8, 3, cudaMemcpyHostToDevice, s);
            if (ce == cudaSuccess && io->e0)
                ce = cudaMemcpy2DAsync(de, ld * 8, io->e0 + off, n_rays * 8, cn * 8, 3, cudaMemcpyHostToDevice, s);
            if (ce != cudaSuccess) { rc = (int)ce; break; }
            if (!io->e0) pyr::fill_rows_kernel<<<64, 256, 0, s>>>(de, ld, cn, 0.0, 1.0, 0.0);   // ray.py:71-73
        }
        // ---- real-valued prefix [0, fa): E is transported and recorded at its last step ----
        const double *sx = dx, *sk = dk, *se = de;
        const uint8_t *salive = nullptr;
        if (fa > 0) {
            for (int i = 0; i < fa; ++i) {
                PyrStep &u = local[i];
                u.out_x = u.out_k = u.out_e = nullptr; u.out_flags = nullptr; u.ld_out = ld;
            }
            PyrStep &u = local[fa - 1];
            u.out_x = px; u.out_k = pk; u.out_e = pe; u.out_flags = pf;
            PyrRaysIn in;
            std::memset(&in, 0, sizeof(in));
            in.x = dx; in.k = dk; in.e = de; in.ld = ld; in.n_x = cn;
            rc = pyr::trace_entry(local.data(), fa, &in, cn, PYR_F_RECORD_E, s);
            if (rc != PYR_OK) break;
            sx = px; sk = pk; se = pe; salive = pf;
        }
        // ---- promote k, E to complex128 (imaginary part 0) ----
        ce = cudaMemsetAsync(kc, 0, 3 * ld * 16, s);
        if (ce == cudaSuccess) ce = cudaMemsetAsync(ec, 0, 3 * ld * 16, s);
        if (ce == cudaSuccess)
            ce = cudaMemcpy2DAsync(kc, 16, sk, 8, 8, 3 * ld, cudaMemcpyDeviceToDevice, s);
        if (ce == cudaSuccess)
            ce = cudaMemcpy2DAsync(ec, 16, se, 8, 8, 3 * ld, cudaMemcpyDeviceToDevice, s);
        if (ce != cudaSuccess) { rc = (int)ce; break; }
        // ---- the complex stretch [fa, n_steps): one launch, records of the doubling steps
        //      (the kernel's stack) and of the last entry ----
        double *lx = nullptr, *lk = nullptr, *le = nullptr;
        uint8_t *lf = nullptr;
        int64_t llw = 0, llw2 = 0;
        for (int i = fa; i < n_steps; ++i) {
            PyrStep &u = local[i];
            u.out_x = u.out_k = u.out_e = nullptr; u.out_flags = nullptr;
            const bool need = (u.split != 0) || i == n_steps - 1;
            const int64_t lw = up(cn * pl.mult_in[i], 32), lw2 = up(cn * pl.mult_out[i], 32);
            u.ld_out = lw; u.ld_out2 = lw2;
            if (!need) continue;
            u.out_x = reinterpret_cast<double *>(take(3 * lw * 8));
            u.out_flags = reinterpret_cast<uint8_t *>(take(lw));
            u.out_k = reinterpret_cast<double *>(take(3 * lw2 * 16));
            u.out_e = reinterpret_cast<double *>(take(3 * lw2 * 16));
            if (i == n_steps - 1) { lx = u.out_x; lf = u.out_flags; lk = u.out_k; le = u.out_e; llw = lw; llw2 = lw2; }
        }
        {
            PyrRaysIn in;
            std::memset(&in, 0, sizeof(in));
            in.x = sx; in.k = kc; in.e = ec; in.alive = salive; in.ld = ld; in.n_x = cn;
            rc = pyr::trace_entry(local.data() + fa, n_steps - fa, &in, cn, PYR_F_COMPLEX | PYR_F_RECORD_E, s);
            if (rc != PYR_OK) break;
        }
        const int64_t wx = cn * pl.mult_x;
        if (io->spot8) {
            rc = pyr_spot_sums(lx, llw, lf, PYR_RAY_ALIVE, wx, steps[n_steps - 1].shape_frame.o, spot_dev, s);
            if (rc != PYR_OK) break;
        }
        // ---- D2H: block b of the chunk -> host columns b n + off ----
        for (int64_t b = 0; b < pl.mult_x && ce == cudaSuccess; ++b) {
            if (io->x_last)
                ce = cudaMemcpy2DAsync(io->x_last + b * n_rays + off, hx * 8, lx + b * cn, llw * 8, cn * 8, 3,
                                       cudaMemcpyDeviceToHost, s);
            if (ce == cudaSuccess && io->flags_last)
                ce = cudaMemcpyAsync(io->flags_last + b * n_rays + off, lf + b * cn, cn, cudaMemcpyDeviceToHost, s);
        }
        for (int64_t b = 0; b < pl.mult_k && ce == cudaSuccess; ++b) {
            if (io->k_last)
                ce = cudaMemcpy2DAsync(io->k_last + 2 * (b * n_rays + off), hk * 16, lk + 2 * b * cn, llw2 * 16,
                                       cn * 16, 3, cudaMemcpyDeviceToHost, s);
            if (ce == cudaSuccess && io->e_last)
                ce = cudaMemcpy2DAsync(io->e_last + 2 * (b * n_rays + off), hk * 16, le + 2 * b * cn, llw2 * 16,
                                       cn * 16, 3, cudaMemcpyDeviceToHost, s);
        }
        if (ce != cudaSuccess) rc = (int)ce;
    }
    for (int i = 0; i < 4; ++i) {
        ce = cudaStreamSynchronize(st[i]);
        if (ce != cudaSuccess && rc == PYR_OK) rc = (int)ce;
    }
    if (rc == PYR_OK && io->spot8) {
        ce = cudaMemcpy(io->spot8, spot_dev, 64, cudaMemcpyDeviceToHost);
        if (ce != cudaSuccess) rc = (int)ce;
    }
    return rc;
}
