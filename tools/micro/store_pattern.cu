// Microbenchmark: how fast can 296 persistent CTAs write the C2 record volume
// (13 entries x (6 rows of doubles + 1 flag byte per ray)) for different HBM layouts
// and tile-to-CTA assignments?  nvcc -O3 -arch=sm_100a -o store_pattern store_pattern.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

constexpr int S = 13, ROWS = 6;

// mode 0: row layout [s][row][ld], cyclic tiles     (what the engine does)
// mode 1: row layout, blocked tiles
// mode 2: tile-major layout [tile][s][row][TILE] (+ flags), cyclic tiles
// mode 3: step-tile-major [s][tile][row][TILE], cyclic tiles
template <int TILE_T>
__global__ void __launch_bounds__(256, 2) pattern(double *out, uint8_t *fl, int64_t n, int64_t ld, int mode,
                                                  int reps_compute) {
    constexpr int TILE = TILE_T;
    const int64_t ntiles = (n + TILE - 1) / TILE;
    const int64_t per = (ntiles + gridDim.x - 1) / gridDim.x;
    int64_t t_begin = blockIdx.x, t_end = ntiles, t_step = gridDim.x;
    if (mode >= 4) { mode -= 4; if (mode == 1) mode = 2; }   // non-persistent: grid == ntiles (t_step skips to the end)
    if (mode == 1) { t_begin = blockIdx.x * per; t_end = min(ntiles, t_begin + per); t_step = 1; }
    double v = threadIdx.x * 1e-3;
    for (int64_t tile = t_begin; tile < t_end; tile += t_step) {
        for (int s = 0; s < S; ++s) {
            for (int r = 0; r < reps_compute; ++r) v = fma(v, 1.0000001, 1e-9);
            for (int e = threadIdx.x * 2; e < TILE; e += blockDim.x * 2) {
                const int64_t i = tile * TILE + e;
                if (i + 1 >= n + 1) continue;
#pragma unroll
                for (int row = 0; row < ROWS; ++row) {
                    double *p;
                    if (mode <= 1) p = out + ((int64_t)s * ROWS + row) * ld + i;
                    else if (mode == 2) p = out + ((tile * S + s) * ROWS + row) * TILE + e;
                    else p = out + (((int64_t)s * ntiles + tile) * ROWS + row) * TILE + e;
                    __stcs(reinterpret_cast<double2 *>(p), make_double2(v, v + row));
                }
                uint8_t *q;
                if (mode <= 1) q = fl + (int64_t)s * ld + i;
                else if (mode == 2) q = fl + (tile * S + s) * TILE + e;
                else q = fl + ((int64_t)s * ntiles + tile) * TILE + e;
                __stcs(reinterpret_cast<uchar2 *>(q), make_uchar2(3, 3));
            }
        }
    }
}

// persistent CTAs, tiles handed out in order by an atomic counter
template <int TILE_T>
__global__ void __launch_bounds__(256, 2) pattern_dyn(double *out, uint8_t *fl, int64_t n, int64_t ld,
                                                      unsigned *counter) {
    constexpr int TILE = TILE_T;
    const int64_t ntiles = (n + TILE - 1) / TILE;
    __shared__ unsigned next_tile;
    double v = threadIdx.x * 1e-3;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) next_tile = atomicAdd(counter, 1u);
        __syncthreads();
        const int64_t tile = next_tile;
        if (tile >= ntiles) break;
        for (int s = 0; s < S; ++s) {
            for (int e = threadIdx.x * 2; e < TILE; e += blockDim.x * 2) {
                const int64_t i = tile * TILE + e;
#pragma unroll
                for (int row = 0; row < ROWS; ++row)
                    __stcs(reinterpret_cast<double2 *>(out + ((int64_t)s * ROWS + row) * ld + i), make_double2(v, v + row));
                __stcs(reinterpret_cast<uchar2 *>(fl + (int64_t)s * ld + i), make_uchar2(3, 3));
            }
        }
    }
}

// engine-like: persistent CTAs, read 9 input rows per tile, write 13 x (6 rows + flags)
template <int TILE_T, bool DYN>
__global__ void __launch_bounds__(256, 2) pattern_rw(const double *in, double *out, uint8_t *fl, int64_t n,
                                                     int64_t ld, unsigned *counter, int fmas) {
    constexpr int TILE = TILE_T;
    const int64_t ntiles = (n + TILE - 1) / TILE;
    __shared__ long long next_tile;
    int64_t tile = blockIdx.x;
    while (tile < ntiles) {
        if (DYN) { if (threadIdx.x == 0) next_tile = gridDim.x + atomicAdd(counter, 1u); }
        const int64_t i = tile * TILE + threadIdx.x * 2;
        double2 acc = make_double2(0.0, 0.0);
        if (i < n) {
#pragma unroll
            for (int row = 0; row < 9; ++row) {
                const double2 t = __ldcs(reinterpret_cast<const double2 *>(in + (int64_t)row * ld + i));
                acc.x += t.x; acc.y += t.y;
            }
        }
        for (int s = 0; s < S; ++s) {
            for (int r = 0; r < fmas; ++r) { acc.x = fma(acc.x, 1.0000001, 1e-9); acc.y = fma(acc.y, 0.9999999, 1e-9); }
            if (i < n) {
#pragma unroll
                for (int row = 0; row < ROWS; ++row)
                    __stcs(reinterpret_cast<double2 *>(out + ((int64_t)s * ROWS + row) * ld + i), make_double2(acc.x, acc.y + row));
                __stcs(reinterpret_cast<uchar2 *>(fl + (int64_t)s * ld + i), make_uchar2(3, 3));
            }
        }
        if (DYN) { __syncthreads(); tile = next_tile; __syncthreads(); } else tile += gridDim.x;
    }
}

// like pattern_rw (dynamic in-order), but a CTA takes GROUP consecutive tiles at a time and
// reads all their inputs in one burst before it writes their records
template <int TILE_T, int GROUP>
__global__ void __launch_bounds__(256, 2) pattern_rw_group(const double *in, double *out, uint8_t *fl, int64_t n,
                                                           int64_t ld, unsigned *counter) {
    constexpr int TILE = TILE_T;
    const int64_t ngroups = ((n + TILE - 1) / TILE + GROUP - 1) / GROUP;
    __shared__ long long next_group;
    int64_t grp = blockIdx.x;
    while (grp < ngroups) {
        if (threadIdx.x == 0) next_group = gridDim.x + atomicAdd(counter, 1u);
        double2 acc[GROUP];
#pragma unroll
        for (int g = 0; g < GROUP; ++g) {
            acc[g] = make_double2(0.0, 0.0);
            const int64_t i = (grp * GROUP + g) * TILE + threadIdx.x * 2;
            if (i < n) {
#pragma unroll
                for (int row = 0; row < 9; ++row) {
                    const double2 t = __ldcs(reinterpret_cast<const double2 *>(in + (int64_t)row * ld + i));
                    acc[g].x += t.x; acc[g].y += t.y;
                }
            }
        }
#pragma unroll
        for (int g = 0; g < GROUP; ++g) {
            const int64_t i = (grp * GROUP + g) * TILE + threadIdx.x * 2;
            for (int s = 0; s < S; ++s) {
                if (i < n) {
#pragma unroll
                    for (int row = 0; row < ROWS; ++row)
                        __stcs(reinterpret_cast<double2 *>(out + ((int64_t)s * ROWS + row) * ld + i), make_double2(acc[g].x, acc[g].y + row));
                    __stcs(reinterpret_cast<uchar2 *>(fl + (int64_t)s * ld + i), make_uchar2(3, 3));
                }
            }
        }
        __syncthreads(); grp = next_group; __syncthreads();
    }
}

template <int POL>
__device__ __forceinline__ void st2(double2 *p, double2 v) {
    if (POL == 0) *p = v;
    else if (POL == 1) __stcs(p, v);
    else if (POL == 2) __stwt(p, v);
    else __stcg(p, v);
}
template <int POL>
__global__ void fill(double2 *p, int64_t n2) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x)
        st2<POL>(p + i, make_double2(1.0, 2.0));
}
// one CTA per contiguous chunk (non-persistent), like torch's elementwise kernels
template <int POL>
__global__ void fill_chunk(double2 *p, int64_t n2) {
    const int64_t base = (int64_t)blockIdx.x * blockDim.x * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int64_t i = base + j * blockDim.x + threadIdx.x;
        if (i < n2) st2<POL>(p + i, make_double2(1.0, 2.0));
    }
}

int main(int argc, char **argv) {
    const int64_t n = 9997351, ld = (n + 15) / 16 * 16;
    const int64_t ldt = (n + 2047) / 2048 * 2048;
    double *out; uint8_t *fl;
    CK(cudaMalloc(&out, (size_t)S * ROWS * ldt * 8 + (1 << 20)));
    CK(cudaMalloc(&fl, (size_t)S * ldt + (1 << 20)));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const double bytes = (double)n * S * 49;
    auto timeit = [&](const char *name, auto launch) {
        float best = 1e9;
        for (int it = 0; it < 8; ++it) {
            cudaEventRecord(a); launch(); cudaEventRecord(b); CK(cudaEventSynchronize(b));
            float ms; cudaEventElapsedTime(&ms, a, b); if (it >= 2 && ms < best) best = ms;
        }
        CK(cudaGetLastError());
        printf("%-46s %.4f ms  %.0f GB/s\n", name, best, bytes / best * 1e-6);
    };
    const int64_t n2 = (int64_t)(bytes / 16);
    timeit("fill default st, 148x8 persistent", [&] { fill<0><<<148 * 8, 256>>>((double2 *)out, n2); });
    timeit("fill st.cs,      148x8 persistent", [&] { fill<1><<<148 * 8, 256>>>((double2 *)out, n2); });
    timeit("fill st.wt,      148x8 persistent", [&] { fill<2><<<148 * 8, 256>>>((double2 *)out, n2); });
    timeit("fill st.cg,      148x8 persistent", [&] { fill<3><<<148 * 8, 256>>>((double2 *)out, n2); });
    timeit("fill default st, 148x2 persistent", [&] { fill<0><<<148 * 2, 256>>>((double2 *)out, n2); });
    timeit("fill default st, 148x32 persistent", [&] { fill<0><<<148 * 32, 256>>>((double2 *)out, n2); });
    timeit("fill default st, chunked CTAs", [&] { fill_chunk<0><<<(unsigned)((n2 + 1023) / 1024), 256>>>((double2 *)out, n2); });
    timeit("fill st.cs,      chunked CTAs", [&] { fill_chunk<1><<<(unsigned)((n2 + 1023) / 1024), 256>>>((double2 *)out, n2); });
    {
        const unsigned nt512 = (unsigned)((n + 511) / 512), nt2048 = (unsigned)((n + 2047) / 2048);
        timeit("NON-persistent TILE 512  row layout", [&] { pattern<512><<<nt512, 256>>>(out, fl, n, ld, 4, 0); });
        timeit("NON-persistent TILE 512  tile-major", [&] { pattern<512><<<nt512, 256>>>(out, fl, n, ld, 5, 0); });
        timeit("NON-persistent TILE 512  step-tile-major", [&] { pattern<512><<<nt512, 256>>>(out, fl, n, ld, 7, 0); });
        timeit("NON-persistent TILE 2048 row layout", [&] { pattern<2048><<<nt2048, 256>>>(out, fl, n, ld, 4, 0); });
        timeit("NON-persistent TILE 2048 tile-major", [&] { pattern<2048><<<nt2048, 256>>>(out, fl, n, ld, 5, 0); });
        CK(cudaFuncSetAttribute(pattern<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        timeit("NON-persistent TILE 512 row layout, 2 CTAs/SM (95 KB smem)", [&] { pattern<512><<<nt512, 256, 95 * 1024>>>(out, fl, n, ld, 4, 0); });
        timeit("NON-persistent TILE 512 step-tile-major, 2 CTAs/SM", [&] { pattern<512><<<nt512, 256, 95 * 1024>>>(out, fl, n, ld, 7, 0); });
        timeit("NON-persistent TILE 512 row layout, 3 CTAs/SM (70 KB smem)", [&] { pattern<512><<<nt512, 256, 70 * 1024>>>(out, fl, n, ld, 4, 0); });
        timeit("NON-persistent TILE 512 row layout, 4 CTAs/SM (50 KB smem)", [&] { pattern<512><<<nt512, 256, 50 * 1024>>>(out, fl, n, ld, 4, 0); });
        timeit("persistent 296 CTAs row layout cyclic, 2 CTAs/SM (95 KB smem)", [&] { pattern<512><<<296, 256, 95 * 1024>>>(out, fl, n, ld, 0, 0); });
        unsigned *counter; CK(cudaMalloc(&counter, 4));
        CK(cudaFuncSetAttribute(pattern_dyn<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        timeit("persistent 1184 CTAs row layout, ATOMIC in-order tiles (8/SM)", [&] { cudaMemsetAsync(counter, 0, 4); pattern_dyn<512><<<1184, 256>>>(out, fl, n, ld, counter); });
        timeit("persistent 296 CTAs row layout, ATOMIC in-order tiles (2/SM)", [&] { cudaMemsetAsync(counter, 0, 4); pattern_dyn<512><<<296, 256, 95 * 1024>>>(out, fl, n, ld, counter); });
        double *in; CK(cudaMalloc(&in, (size_t)9 * ld * 8)); CK(cudaMemset(in, 0, (size_t)9 * ld * 8));
        CK(cudaFuncSetAttribute(pattern_rw<512, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        CK(cudaFuncSetAttribute(pattern_rw<512, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        for (int fm = 0; fm <= 60; fm += 30) {
            char nm[160];
            snprintf(nm, sizeof nm, "R+W persistent 296 static cyclic, %d fma pairs/step", fm);
            timeit(nm, [&] { pattern_rw<512, false><<<296, 256, 95 * 1024>>>(in, out, fl, n, ld, counter, fm); });
            snprintf(nm, sizeof nm, "R+W persistent 296 dynamic in-order, %d fma pairs/step", fm);
            timeit(nm, [&] { cudaMemsetAsync(counter, 0, 4); pattern_rw<512, true><<<296, 256, 95 * 1024>>>(in, out, fl, n, ld, counter, fm); });
            snprintf(nm, sizeof nm, "R+W NON-persistent, %d fma pairs/step", fm);
            timeit(nm, [&] { pattern_rw<512, false><<<nt512, 256, 95 * 1024>>>(in, out, fl, n, ld, counter, fm); });
        }
        CK(cudaFuncSetAttribute(pattern_rw_group<512, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        CK(cudaFuncSetAttribute(pattern_rw_group<512, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        CK(cudaFuncSetAttribute(pattern_rw_group<512, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        timeit("R+W dynamic, input bursts of 2 tiles", [&] { cudaMemsetAsync(counter, 0, 4); pattern_rw_group<512, 2><<<296, 256, 95 * 1024>>>(in, out, fl, n, ld, counter); });
        timeit("R+W dynamic, input bursts of 4 tiles", [&] { cudaMemsetAsync(counter, 0, 4); pattern_rw_group<512, 4><<<296, 256, 95 * 1024>>>(in, out, fl, n, ld, counter); });
        timeit("R+W dynamic, input bursts of 8 tiles", [&] { cudaMemsetAsync(counter, 0, 4); pattern_rw_group<512, 8><<<296, 256, 95 * 1024>>>(in, out, fl, n, ld, counter); });
        timeit("persistent 1184 CTAs row layout cyclic (8/SM)", [&] { pattern<512><<<1184, 256>>>(out, fl, n, ld, 0, 0); });
        timeit("NON-persistent TILE 512  row layout + 400 FMAs/step", [&] { pattern<512><<<nt512, 256>>>(out, fl, n, ld, 4, 400); });
    }
    const char *names[4] = {"row layout, cyclic tiles (engine)", "row layout, blocked tiles", "tile-major [tile][s][row]", "step-tile-major [s][tile][row]"};
    for (int mode = 0; mode < 4; ++mode) {
        char nm[128];
        snprintf(nm, sizeof nm, "TILE 512  %s", names[mode]);
        timeit(nm, [&] { pattern<512><<<296, 256>>>(out, fl, n, ld, mode, 0); });
    }
    for (int mode = 0; mode < 4; ++mode) {
        char nm[128];
        snprintf(nm, sizeof nm, "TILE 2048 %s", names[mode]);
        timeit(nm, [&] { pattern<2048><<<296, 256>>>(out, fl, n, ld, mode, 0); });
    }
    for (int mode = 0; mode < 4; mode += 3) {
        char nm[128];
        snprintf(nm, sizeof nm, "TILE 512 + 400 dependent FMAs/step %s", names[mode]);
        timeit(nm, [&] { pattern<512><<<296, 256>>>(out, fl, n, ld, mode, 400); });
    }
    return 0;
}
