// End-to-end host entry of the C ABI: bundle in (pinned) host memory in, final
// surface record + spot sums back in host memory.  The bundle is cut into chunks
// that cycle through four device slots on four streams, so the H2D copy of
// chunk c+1, the trace of chunk c and the D2H copy of chunk c-1 overlap.  The first
// and last chunks are short (1/8, 1/4, 1/2 of the nominal size): the pipeline's
// fill (first H2D with nothing to overlap) and drain (last D2H) shrink with them.
#include <cuda_runtime.h>

#include <vector>

#include "pyr_device.cuh"

namespace pyr {
int trace_entry(const PyrStep *steps, int32_t n_steps, const PyrRaysIn *rays, int64_t n_rays,
                uint32_t flags, cudaStream_t stream);
}

namespace {
constexpr int kSlots = 4;

int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

// three rows of a (3, n) array as ONE 2-D copy (measured: three 1-D copies per array
// cost 16.95 ms instead of 15.44 ms per 1e7-ray pass, profiles/r01_e2e_pcie.txt)
cudaError_t copy_rows(double *dst, size_t dst_ld, const double *src, size_t src_ld, size_t count,
                      cudaMemcpyKind kind, cudaStream_t s) {
    return cudaMemcpy2DAsync(dst, dst_ld * 8, src, src_ld * 8, count * 8, 3, kind, s);
}

struct SlotLayout {
    int64_t ld;          // padded chunk width (doubles)
    int64_t in_bytes;    // x, k, e : 3 x (3, ld)
    int64_t out_bytes;   // x_last, k_last : 2 x (3, ld)
    int64_t flag_bytes;
    int64_t slot_bytes;
};

SlotLayout layout(int64_t chunk) {
    SlotLayout L;
    L.ld = round_up(chunk, 32);
    L.in_bytes = 9 * L.ld * 8;
    L.out_bytes = 6 * L.ld * 8;
    L.flag_bytes = round_up(L.ld, 256);
    L.slot_bytes = L.in_bytes + L.out_bytes + L.flag_bytes;
    return L;
}
}  // namespace

extern "C" {

int64_t pyr_trace_host_workspace(int32_t n_steps, int64_t chunk_rays) {
    (void)n_steps;
    if (chunk_rays <= 0) return 0;
    return kSlots * layout(chunk_rays).slot_bytes + 256;
}

int pyr_trace_host(const PyrStep *steps, int32_t n_steps, const double *x0, const double *k0,
                   const double *e0, int64_t n_rays, double *x_last, double *k_last,
                   uint8_t *flags_last, double *spot8, void *workspace, int64_t workspace_bytes,
                   int64_t chunk_rays) {
    if (!steps || n_steps <= 0 || !x0 || !k0 || n_rays < 0 || !workspace || chunk_rays <= 0)
        return PYR_E_BADARG;
    if (workspace_bytes < pyr_trace_host_workspace(n_steps, chunk_rays)) return PYR_E_BADARG;
    if ((reinterpret_cast<uintptr_t>(workspace) & 255u) != 0) return PYR_E_BADARG;
    const SlotLayout L = layout(chunk_rays);
    char *ws = static_cast<char *>(workspace);
    double *spot_dev = reinterpret_cast<double *>(ws + kSlots * L.slot_bytes);

    cudaStream_t st[kSlots];
    for (int i = 0; i < kSlots; ++i) {
        cudaError_t e = cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking);
        if (e != cudaSuccess) return (int)e;
    }
    int rc = PYR_OK;
    cudaError_t ce = cudaMemsetAsync(spot_dev, 0, 64, st[0]);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st[0]);
    if (ce != cudaSuccess) rc = (int)ce;

    std::vector<PyrStep> local(steps, steps + n_steps);
    // chunk schedule: ramp up, nominal chunks, ramp down (multiples of 32 rays)
    std::vector<int64_t> sizes;
    {
        int64_t ramp[3] = {round_up(chunk_rays / 8, 32), round_up(chunk_rays / 4, 32),
                           round_up(chunk_rays / 2, 32)};
        int64_t ramp_total = 2 * (ramp[0] + ramp[1] + ramp[2]);
        if (chunk_rays >= 4096 && n_rays >= ramp_total + chunk_rays) {
            int64_t body = n_rays - ramp_total;
            for (int i = 0; i < 3; ++i) sizes.push_back(ramp[i]);
            while (body > 0) { sizes.push_back(body < chunk_rays ? body : chunk_rays); body -= chunk_rays; }
            for (int i = 2; i >= 0; --i) sizes.push_back(ramp[i]);
        } else {
            for (int64_t off = 0; off < n_rays; off += chunk_rays)
                sizes.push_back(n_rays - off < chunk_rays ? n_rays - off : chunk_rays);
        }
    }
    int64_t off = 0;
    for (size_t chunk_idx = 0; chunk_idx < sizes.size() && rc == PYR_OK; off += sizes[chunk_idx], ++chunk_idx) {
        const int64_t cn = sizes[chunk_idx];
        const int slot = (int)(chunk_idx % kSlots);
        cudaStream_t s = st[slot];
        char *base = ws + slot * L.slot_bytes;
        double *dx = reinterpret_cast<double *>(base);
        double *dk = dx + 3 * L.ld;
        double *de = dk + 3 * L.ld;
        double *ox = de + 3 * L.ld;
        double *ok = ox + 3 * L.ld;
        uint8_t *of = reinterpret_cast<uint8_t *>(ok + 3 * L.ld);
        ce = copy_rows(dx, L.ld, x0 + off, n_rays, cn, cudaMemcpyHostToDevice, s);
        if (ce == cudaSuccess)
            ce = copy_rows(dk, L.ld, k0 + off, n_rays, cn, cudaMemcpyHostToDevice, s);
        if (ce == cudaSuccess && e0)
            ce = copy_rows(de, L.ld, e0 + off, n_rays, cn, cudaMemcpyHostToDevice, s);
        if (ce != cudaSuccess) { rc = (int)ce; break; }
        for (int i = 0; i < n_steps; ++i) {
            local[i].out_x = nullptr; local[i].out_k = nullptr; local[i].out_e = nullptr;
            local[i].out_flags = nullptr; local[i].ld_out = L.ld;
        }
        local[n_steps - 1].out_x = ox;
        local[n_steps - 1].out_k = ok;
        local[n_steps - 1].out_flags = of;
        PyrRaysIn in;
        in.x = dx; in.k = dk; in.e = e0 ? de : nullptr; in.alive = nullptr; in.ld = L.ld; in.n_x = cn;
        in.n_waves = 0; in.reserved0 = 0;
        for (int w = 0; w < PYR_MAX_WAVES; ++w) in.wave_end[w] = 0;
        rc = pyr::trace_entry(local.data(), n_steps, &in, cn, 0u, s);
        if (rc != PYR_OK) break;
        if (spot8) {
            rc = pyr_spot_sums(ox, L.ld, of, PYR_RAY_ALIVE, cn, steps[n_steps - 1].shape_frame.o,
                               spot_dev, s);
            if (rc != PYR_OK) break;
        }
        if (x_last)
            ce = copy_rows(x_last + off, n_rays, ox, L.ld, cn, cudaMemcpyDeviceToHost, s);
        if (ce == cudaSuccess && k_last)
            ce = copy_rows(k_last + off, n_rays, ok, L.ld, cn, cudaMemcpyDeviceToHost, s);
        if (ce == cudaSuccess && flags_last)
            ce = cudaMemcpyAsync(flags_last + off, of, (size_t)cn, cudaMemcpyDeviceToHost, s);
        if (ce != cudaSuccess) rc = (int)ce;
    }
    for (int i = 0; i < kSlots; ++i) {
        ce = cudaStreamSynchronize(st[i]);
        if (ce != cudaSuccess && rc == PYR_OK) rc = (int)ce;
    }
    if (rc == PYR_OK && spot8) {
        ce = cudaMemcpy(spot8, spot_dev, 64, cudaMemcpyDeviceToHost);
        if (ce != cudaSuccess) rc = (int)ce;
    }
    for (int i = 0; i < kSlots; ++i) cudaStreamDestroy(st[i]);
    return rc;
}

}  // extern "C"
