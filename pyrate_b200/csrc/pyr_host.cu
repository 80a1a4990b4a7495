// End-to-end host entries of the C ABI: bundle in (pinned) host memory -- or described by
// a generator, then nothing is uploaded -- in; last record or every record + spot sums back
// in host memory.  The bundle is cut into chunks that cycle through four device slots on
// four streams, so the H2D copy of chunk c+1, the trace of chunk c and the D2H copy of
// chunk c-1 overlap.  The first and last chunks are short (1/8, 1/4, 1/2 of the nominal
// size): the pipeline's fill (first H2D with nothing to overlap) and drain (last D2H)
// shrink with them.  Sequences longer than one launch carries (40 steps / 10 auxiliary
// records) continue from the last record of the previous launch inside the slot.
//
// Reference boundary: OpticalSystem.seqtrace (raytracer/optical_system.py:73-94) called
// with NumPy bundles; RayPath of S + 2 bundles (ray.py:207-260) = the `*_all` outputs.
#include <cuda_runtime.h>

#include <cstring>
#include <mutex>
#include <vector>

#include "pyr_device.cuh"

namespace pyr {
int trace_entry(const PyrStep *steps, int32_t n_steps, const PyrRaysIn *rays, int64_t n_rays,
                uint32_t flags, cudaStream_t stream);
int step_aux_records(const PyrStep &u);     // pyr_trace.cu: auxiliary records pack() will use
}

int pyr_trace_host_crystal(const PyrStep *steps, int32_t n_steps, const PyrHostIO *io, int64_t n_rays,
                           void *workspace, int64_t workspace_bytes, int64_t chunk_rays, cudaStream_t *st);

namespace {
constexpr int kSlots = 4;

int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

// `rows` rows of a row-major array as ONE 2-D copy (measured: three 1-D copies per (3, n)
// array cost 16.95 ms instead of 15.44 ms per 1e7-ray pass, profiles/r01_e2e_pcie.txt)
cudaError_t copy_rows(double *dst, size_t dst_ld, const double *src, size_t src_ld, size_t count,
                      size_t rows, cudaMemcpyKind kind, cudaStream_t s) {
    return cudaMemcpy2DAsync(dst, dst_ld * 8, src, src_ld * 8, count * 8, rows, kind, s);
}

struct SlotLayout {
    int64_t ld;          // padded chunk width (doubles)
    int64_t in_bytes;    // x, k, e : 3 x (3, ld)
    int64_t n_rec;       // record buffers of a slot: all steps, or 1 (2 when launches chain)
    int64_t rec_bytes;   // x, k of one record: 2 x (3, ld)
    int64_t flag_bytes;  // flags of one record
    int64_t slot_bytes;
};

SlotLayout layout(int64_t chunk, int64_t n_rec) {
    SlotLayout L;
    L.ld = round_up(chunk, 32);
    L.in_bytes = 9 * L.ld * 8;
    L.n_rec = n_rec;
    L.rec_bytes = 6 * L.ld * 8;
    L.flag_bytes = round_up(L.ld, 256);
    L.slot_bytes = L.in_bytes + n_rec * (L.rec_bytes + L.flag_bytes);
    return L;
}

// Streams are created once per device and kept (creating and destroying four streams per
// call cost ~40 us and leaked the earlier ones when a later creation failed).
struct StreamSet {
    cudaStream_t s[kSlots];
    bool ok = false;
};
std::mutex g_mu;
StreamSet g_streams[64];

int get_streams(cudaStream_t **out) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev < 0 || dev >= 64) return PYR_E_BADARG;
    std::lock_guard<std::mutex> lock(g_mu);
    StreamSet &ss = g_streams[dev];
    if (!ss.ok) {
        int made = 0;
        for (; made < kSlots; ++made) {
            e = cudaStreamCreateWithFlags(&ss.s[made], cudaStreamNonBlocking);
            if (e != cudaSuccess) break;
        }
        if (made < kSlots) {
            for (int i = 0; i < made; ++i) cudaStreamDestroy(ss.s[i]);
            return (int)e;
        }
        ss.ok = true;
    }
    *out = ss.s;
    return PYR_OK;
}

// launches of a sequence: [cut[i], cut[i+1]) obeys the per-launch limits
std::vector<int> launch_cuts(const PyrStep *steps, int32_t n_steps) {
    std::vector<int> cuts{0};
    int count = 0, aux = 0;
    for (int i = 0; i < n_steps; ++i) {
        const int a = pyr::step_aux_records(steps[i]);
        if (count + 1 > pyr::kMaxSteps || aux + a > pyr::kMaxAux) {
            cuts.push_back(i);
            count = 0;
            aux = 0;
        }
        ++count;
        aux += a;
    }
    cuts.push_back(n_steps);
    return cuts;
}
}  // namespace

extern "C" {

int64_t pyr_trace_host_io_workspace(int32_t n_steps, int64_t chunk_rays, int32_t all_records) {
    if (chunk_rays <= 0 || n_steps <= 0) return 0;
    // last-record mode keeps two record buffers so that chained launches can ping-pong
    return kSlots * layout(chunk_rays, all_records ? n_steps : 2).slot_bytes + 256;
}

int64_t pyr_trace_host_workspace(int32_t n_steps, int64_t chunk_rays) {
    return pyr_trace_host_io_workspace(n_steps > 0 ? n_steps : 1, chunk_rays, 0);
}

int pyr_trace_host_io(const PyrStep *steps, int32_t n_steps, const PyrHostIO *io, int64_t n_rays,
                      void *workspace, int64_t workspace_bytes, int64_t chunk_rays) {
    if (!steps || n_steps <= 0 || !io || n_rays < 0 || !workspace || chunk_rays <= 0)
        return PYR_E_BADARG;
    if (!io->gen && (!io->x0 || !io->k0)) return PYR_E_BADARG;
    const bool all = io->x_all || io->k_all || io->flags_all;
    if ((reinterpret_cast<uintptr_t>(workspace) & 255u) != 0) return PYR_E_BADARG;
    for (int i = 0; i < n_steps; ++i)
        if (steps[i].before.kind == PYR_MEDIUM_ANISO || steps[i].after.kind == PYR_MEDIUM_ANISO) {
            // birefringent media: the complex-valued pipeline (pyr_host_crystal.cu)
            cudaStream_t *cst = nullptr;
            const int crc = get_streams(&cst);
            if (crc != PYR_OK) return crc;
            return pyr_trace_host_crystal(steps, n_steps, io, n_rays, workspace, workspace_bytes, chunk_rays, cst);
        }
    if (workspace_bytes < pyr_trace_host_io_workspace(n_steps, chunk_rays, all ? 1 : 0)) return PYR_E_BADARG;
    const SlotLayout L = layout(chunk_rays, all ? n_steps : 2);
    char *ws = static_cast<char *>(workspace);
    double *spot_dev = reinterpret_cast<double *>(ws + kSlots * L.slot_bytes);
    const std::vector<int> cuts = launch_cuts(steps, n_steps);

    cudaStream_t *st = nullptr;
    int rc = get_streams(&st);
    if (rc != PYR_OK) return rc;
    cudaError_t ce = cudaSuccess;
    if (io->spot8) {
        ce = cudaMemsetAsync(spot_dev, 0, 64, st[0]);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(st[0]);
        if (ce != cudaSuccess) return (int)ce;
    }

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
        char *rec0 = base + L.in_bytes;
        auto rec_x = [&](int64_t r) { return reinterpret_cast<double *>(rec0 + r * L.rec_bytes); };
        auto rec_k = [&](int64_t r) { return rec_x(r) + 3 * L.ld; };
        auto rec_f = [&](int64_t r) {
            return reinterpret_cast<uint8_t *>(rec0 + L.n_rec * L.rec_bytes + r * L.flag_bytes);
        };
        if (!io->gen) {
            ce = copy_rows(dx, L.ld, io->x0 + off, n_rays, cn, 3, cudaMemcpyHostToDevice, s);
            if (ce == cudaSuccess)
                ce = copy_rows(dk, L.ld, io->k0 + off, n_rays, cn, 3, cudaMemcpyHostToDevice, s);
            if (ce == cudaSuccess && io->e0)
                ce = copy_rows(de, L.ld, io->e0 + off, n_rays, cn, 3, cudaMemcpyHostToDevice, s);
            if (ce != cudaSuccess) { rc = (int)ce; break; }
        }
        // record buffer of step i: its own (all records) or ping-pong per launch (last only)
        int64_t last_rec = 0;
        for (size_t li = 0; li + 1 < cuts.size() && rc == PYR_OK; ++li) {
            const int lo = cuts[li], hi = cuts[li + 1];
            for (int i = lo; i < hi; ++i) {
                PyrStep &u = local[i];
                u.out_e = nullptr; u.ld_out = L.ld;
                if (all) {
                    u.out_x = rec_x(i); u.out_k = rec_k(i); u.out_flags = rec_f(i);
                } else if (i == hi - 1) {
                    const int64_t r = (int64_t)(li & 1);
                    u.out_x = rec_x(r); u.out_k = rec_k(r); u.out_flags = rec_f(r);
                } else {
                    u.out_x = nullptr; u.out_k = nullptr; u.out_flags = nullptr;
                }
            }
            PyrRaysIn in;
            std::memset(&in, 0, sizeof(in));
            PyrBundleGen gen;
            if (li == 0) {
                if (io->gen) {
                    gen = *io->gen;
                    gen.first += off;
                    in.gen = &gen;
                } else {
                    in.x = dx; in.k = dk; in.e = io->e0 ? de : nullptr;
                }
            } else {
                in.x = rec_x(last_rec); in.k = rec_k(last_rec); in.alive = rec_f(last_rec);
            }
            in.ld = L.ld; in.n_x = cn;
            rc = pyr::trace_entry(local.data() + lo, hi - lo, &in, cn, 0u, s);
            last_rec = all ? hi - 1 : (int64_t)(li & 1);
        }
        if (rc != PYR_OK) break;
        double *ox = rec_x(last_rec), *ok = rec_k(last_rec);
        uint8_t *of = rec_f(last_rec);
        if (io->spot8) {
            rc = pyr_spot_sums(ox, L.ld, of, PYR_RAY_ALIVE, cn, steps[n_steps - 1].shape_frame.o,
                               spot_dev, s);
            if (rc != PYR_OK) break;
        }
        if (io->x_last)
            ce = copy_rows(io->x_last + off, n_rays, ox, L.ld, cn, 3, cudaMemcpyDeviceToHost, s);
        if (ce == cudaSuccess && io->k_last)
            ce = copy_rows(io->k_last + off, n_rays, ok, L.ld, cn, 3, cudaMemcpyDeviceToHost, s);
        if (ce == cudaSuccess && io->flags_last)
            ce = cudaMemcpyAsync(io->flags_last + off, of, (size_t)cn, cudaMemcpyDeviceToHost, s);
        if (all && ce == cudaSuccess) {
            // every record: the device layout is (n_steps, [x: 3 rows | k: 3 rows], ld), the
            // host arrays are (n_steps, 3, n_rays): one 2-D copy per record and array
            for (int i = 0; i < n_steps && ce == cudaSuccess; ++i) {
                if (io->x_all)
                    ce = copy_rows(io->x_all + (int64_t)i * 3 * n_rays + off, n_rays, rec_x(i), L.ld, cn, 3,
                                   cudaMemcpyDeviceToHost, s);
                if (ce == cudaSuccess && io->k_all)
                    ce = copy_rows(io->k_all + (int64_t)i * 3 * n_rays + off, n_rays, rec_k(i), L.ld, cn, 3,
                                   cudaMemcpyDeviceToHost, s);
            }
            if (ce == cudaSuccess && io->flags_all)
                ce = cudaMemcpy2DAsync(io->flags_all + off, (size_t)n_rays, rec_f(0), (size_t)L.flag_bytes,
                                       (size_t)cn, (size_t)n_steps, cudaMemcpyDeviceToHost, s);
        }
        if (ce != cudaSuccess) rc = (int)ce;
    }
    for (int i = 0; i < kSlots; ++i) {
        ce = cudaStreamSynchronize(st[i]);
        if (ce != cudaSuccess && rc == PYR_OK) rc = (int)ce;
    }
    if (rc == PYR_OK && io->spot8) {
        ce = cudaMemcpy(io->spot8, spot_dev, 64, cudaMemcpyDeviceToHost);
        if (ce != cudaSuccess) rc = (int)ce;
    }
    return rc;
}

int pyr_trace_host(const PyrStep *steps, int32_t n_steps, const double *x0, const double *k0,
                   const double *e0, int64_t n_rays, double *x_last, double *k_last,
                   uint8_t *flags_last, double *spot8, void *workspace, int64_t workspace_bytes,
                   int64_t chunk_rays) {
    PyrHostIO io;
    std::memset(&io, 0, sizeof(io));
    io.x0 = x0; io.k0 = k0; io.e0 = e0;
    io.x_last = x_last; io.k_last = k_last; io.flags_last = flags_last; io.spot8 = spot8;
    return pyr_trace_host_io(steps, n_steps, &io, n_rays, workspace, workspace_bytes, chunk_rays);
}

}  // extern "C"
