// exp(x) in FP64 for the GRIN integrator: table-driven (Tang 1989):
//   x = (32 k + j) ln2/32 + r, |r| <= ln2/64;  exp(x) = 2^k * 2^(j/32) * (1 + P(r)),
//   P = r + r^2/2 + ... + r^7/5040 (truncation 4e-19 relative).
// 12 FP64 instructions and one shared-memory load instead of the ~18 FP64 + ~12 uniform
// moves of CUDA's exp() (whose degree-11 polynomial carries its coefficients as 64-bit
// immediates): the GRIN profile n = n0 + g exp(-a x^2 - b y^2) evaluates one exp per
// integrator stage, four per step, 800 per ray of BASELINE config 5 -- it was 60 % of the
// FP64 work of that kernel (profiles/r02_grin_exp.md).  Measured against std::exp on the
// host (tools/micro/test_exp.cu, tests/test_exp_host.py): <= 1 ulp over [-700, 700].
// Out of that range (and NaN) the caller falls back to exp().
#pragma once

#include <cuda_runtime.h>

namespace pyr {

#ifdef __CUDA_ARCH__
#define PYR_EXP_CONST __constant__
#else
#define PYR_EXP_CONST static const
#endif

// 2^(j/32), j = 0..31, correctly rounded
PYR_EXP_CONST double kExp2Tab[32] = {
    0x1.0000000000000p+0, 0x1.059b0d3158574p+0, 0x1.0b5586cf9890fp+0, 0x1.11301d0125b51p+0,
    0x1.172b83c7d517bp+0, 0x1.1d4873168b9aap+0, 0x1.2387a6e756238p+0, 0x1.29e9df51fdee1p+0,
    0x1.306fe0a31b715p+0, 0x1.371a7373aa9cbp+0, 0x1.3dea64c123422p+0, 0x1.44e086061892dp+0,
    0x1.4bfdad5362a27p+0, 0x1.5342b569d4f82p+0, 0x1.5ab07dd485429p+0, 0x1.6247eb03a5585p+0,
    0x1.6a09e667f3bcdp+0, 0x1.71f75e8ec5f74p+0, 0x1.7a11473eb0187p+0, 0x1.82589994cce13p+0,
    0x1.8ace5422aa0dbp+0, 0x1.93737b0cdc5e5p+0, 0x1.9c49182a3f090p+0, 0x1.a5503b23e255dp+0,
    0x1.ae89f995ad3adp+0, 0x1.b7f76f2fb5e47p+0, 0x1.c199bdd85529cp+0, 0x1.cb720dcef9069p+0,
    0x1.d5818dcfba487p+0, 0x1.dfc97337b9b5fp+0, 0x1.ea4afa2a490dap+0, 0x1.f50765b6e4540p+0};

// constants of exp_tab in a constant-bank array: the FP64 instructions take them as c[][]
// operands (literals would be materialised with two uniform moves each, per use)
PYR_EXP_CONST double kExpK[9] = {
    0x1.71547652b82fep+5,        // 32 / ln 2
    0x1.62e42fe000000p-6,        // ln 2 / 32, 24 trailing zero bits: n * hi exact
    0x1.f473de6af278fp-35,       // ln 2 / 32 - hi
    0x1.8p52,                    // 2^52 + 2^51: rounds to nearest integer
    1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0};

// `tab`: the table above (the kernels keep a copy in shared memory: a per-thread index into
// constant memory would serialise)
__host__ __device__ __forceinline__ double exp_tab(double x, const double *tab) {
    const double kInvL = kExpK[0], kLhi = kExpK[1], kLlo = kExpK[2], kMagic = kExpK[3];
    const double t = fma(x, kInvL, kMagic);
    const double nf = t - kMagic;
#ifdef __CUDA_ARCH__
    const int n = __double2loint(t);
#else
    long long bits;
    __builtin_memcpy(&bits, &t, 8);
    const int n = (int)(unsigned)(bits & 0xffffffffll);
#endif
    double r = fma(-nf, kLhi, x);
    r = fma(-nf, kLlo, r);
    // P(r) = r + r^2 (1/2 + r (1/6 + r (1/24 + r (1/120 + r (1/720 + r / 5040)))))
    double p = fma(r, kExpK[4], kExpK[5]);
    p = fma(p, r, kExpK[6]);
    p = fma(p, r, kExpK[7]);
    p = fma(p, r, kExpK[8]);
    p = fma(p, r, 0.5);
    p = fma(p * r, r, r);
    const double tj = tab[n & 31];
    const double res = fma(tj, p, tj);
    const int k = n >> 5;                             // arithmetic shift: floor
#ifdef __CUDA_ARCH__
    return __hiloint2double(__double2hiint(res) + (k << 20), __double2loint(res));
#else
    long long rb;
    __builtin_memcpy(&rb, &res, 8);
    rb += (long long)k << 52;
    double out;
    __builtin_memcpy(&out, &rb, 8);
    return out;
#endif
}

// N independent arguments, operation by operation (source-level interleave: a warp issues
// in order, so the N Horner chains only overlap if their instructions alternate)
template <int N>
__device__ __forceinline__ void exp_tab_n(const double *x, double *out, const double *tab) {
    const double kInvL = kExpK[0], kLhi = kExpK[1], kLlo = kExpK[2], kMagic = kExpK[3];
    double t[N], r[N], p[N], tj[N];
    int n[N];
#pragma unroll
    for (int j = 0; j < N; ++j) t[j] = fma(x[j], kInvL, kMagic);
#pragma unroll
    for (int j = 0; j < N; ++j) { n[j] = __double2loint(t[j]); t[j] = t[j] - kMagic; }
#pragma unroll
    for (int j = 0; j < N; ++j) tj[j] = tab[n[j] & 31];
#pragma unroll
    for (int j = 0; j < N; ++j) r[j] = fma(-t[j], kLhi, x[j]);
#pragma unroll
    for (int j = 0; j < N; ++j) r[j] = fma(-t[j], kLlo, r[j]);
#pragma unroll
    for (int j = 0; j < N; ++j) p[j] = fma(r[j], kExpK[4], kExpK[5]);
#pragma unroll
    for (int j = 0; j < N; ++j) p[j] = fma(p[j], r[j], kExpK[6]);
#pragma unroll
    for (int j = 0; j < N; ++j) p[j] = fma(p[j], r[j], kExpK[7]);
#pragma unroll
    for (int j = 0; j < N; ++j) p[j] = fma(p[j], r[j], kExpK[8]);
#pragma unroll
    for (int j = 0; j < N; ++j) p[j] = fma(p[j], r[j], 0.5);
#pragma unroll
    for (int j = 0; j < N; ++j) p[j] = fma(p[j] * r[j], r[j], r[j]);
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const double res = fma(tj[j], p[j], tj[j]);
        out[j] = __hiloint2double(__double2hiint(res) + ((n[j] >> 5) << 20), __double2loint(res));
    }
}

// valid range of exp_tab (result and 2^k normal)
__host__ __device__ __forceinline__ bool exp_tab_ok(double x) { return x > -700.0 && x < 700.0; }

}  // namespace pyr
