// exp(x) in FP64 for the GRIN integrator: table-driven (Tang 1989):
//   x = (128 k + j) ln2/128 + r, |r| <= ln2/256;  exp(x) = 2^k * 2^(j/128) * (1 + P(r)),
//   P = r + r^2/2 + ... + r^5/120 (truncation r^6/720 <= 6e-19 relative).
// 10 FP64 instructions and one shared-memory load instead of the ~18 FP64 + ~12 uniform
// moves of CUDA's exp() (whose degree-11 polynomial carries its coefficients as 64-bit
// immediates): the GRIN profile n = n0 + g exp(-a x^2 - b y^2) evaluates one exp per
// integrator stage, four per step, 800 per ray of BASELINE config 5 -- it was 60 % of the
// FP64 work of that kernel (profiles/r02_grin.md).  Measured against expl on the host
// (tools/micro/test_exp.cu, tests/test_exp_host.py): <= 1 ulp over [-700, 700].
// exp_tab_any covers every argument (0 below the range, NaN above it and for NaN) with an
// integer test of the argument's high word that stays off the arithmetic's dependency chain.
#pragma once

#include <cuda_runtime.h>

namespace pyr {

#ifdef __CUDA_ARCH__
#define PYR_EXP_CONST __constant__
#else
#define PYR_EXP_CONST static const
#endif

constexpr int kExpTabSize = 128;

// 2^(j/128), j = 0..127, correctly rounded
PYR_EXP_CONST double kExp2Tab[kExpTabSize] = {
    0x1.0000000000000p+0, 0x1.0163da9fb3335p+0, 0x1.02c9a3e778061p+0, 0x1.04315e86e7f85p+0,
    0x1.059b0d3158574p+0, 0x1.0706b29ddf6dep+0, 0x1.0874518759bc8p+0, 0x1.09e3ecac6f383p+0,
    0x1.0b5586cf9890fp+0, 0x1.0cc922b7247f7p+0, 0x1.0e3ec32d3d1a2p+0, 0x1.0fb66affed31bp+0,
    0x1.11301d0125b51p+0, 0x1.12abdc06c31ccp+0, 0x1.1429aaea92de0p+0, 0x1.15a98c8a58e51p+0,
    0x1.172b83c7d517bp+0, 0x1.18af9388c8deap+0, 0x1.1a35beb6fcb75p+0, 0x1.1bbe084045cd4p+0,
    0x1.1d4873168b9aap+0, 0x1.1ed5022fcd91dp+0, 0x1.2063b88628cd6p+0, 0x1.21f49917ddc96p+0,
    0x1.2387a6e756238p+0, 0x1.251ce4fb2a63fp+0, 0x1.26b4565e27cddp+0, 0x1.284dfe1f56381p+0,
    0x1.29e9df51fdee1p+0, 0x1.2b87fd0dad990p+0, 0x1.2d285a6e4030bp+0, 0x1.2ecafa93e2f56p+0,
    0x1.306fe0a31b715p+0, 0x1.32170fc4cd831p+0, 0x1.33c08b26416ffp+0, 0x1.356c55f929ff1p+0,
    0x1.371a7373aa9cbp+0, 0x1.38cae6d05d866p+0, 0x1.3a7db34e59ff7p+0, 0x1.3c32dc313a8e5p+0,
    0x1.3dea64c123422p+0, 0x1.3fa4504ac801cp+0, 0x1.4160a21f72e2ap+0, 0x1.431f5d950a897p+0,
    0x1.44e086061892dp+0, 0x1.46a41ed1d0057p+0, 0x1.486a2b5c13cd0p+0, 0x1.4a32af0d7d3dep+0,
    0x1.4bfdad5362a27p+0, 0x1.4dcb299fddd0dp+0, 0x1.4f9b2769d2ca7p+0, 0x1.516daa2cf6642p+0,
    0x1.5342b569d4f82p+0, 0x1.551a4ca5d920fp+0, 0x1.56f4736b527dap+0, 0x1.58d12d497c7fdp+0,
    0x1.5ab07dd485429p+0, 0x1.5c9268a5946b7p+0, 0x1.5e76f15ad2148p+0, 0x1.605e1b976dc09p+0,
    0x1.6247eb03a5585p+0, 0x1.6434634ccc320p+0, 0x1.6623882552225p+0, 0x1.68155d44ca973p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6c012750bdabfp+0, 0x1.6dfb23c651a2fp+0, 0x1.6ff7df9519484p+0,
    0x1.71f75e8ec5f74p+0, 0x1.73f9a48a58174p+0, 0x1.75feb564267c9p+0, 0x1.780694fde5d3fp+0,
    0x1.7a11473eb0187p+0, 0x1.7c1ed0130c132p+0, 0x1.7e2f336cf4e62p+0, 0x1.80427543e1a12p+0,
    0x1.82589994cce13p+0, 0x1.8471a4623c7adp+0, 0x1.868d99b4492edp+0, 0x1.88ac7d98a6699p+0,
    0x1.8ace5422aa0dbp+0, 0x1.8cf3216b5448cp+0, 0x1.8f1ae99157736p+0, 0x1.9145b0b91ffc6p+0,
    0x1.93737b0cdc5e5p+0, 0x1.95a44cbc8520fp+0, 0x1.97d829fde4e50p+0, 0x1.9a0f170ca07bap+0,
    0x1.9c49182a3f090p+0, 0x1.9e86319e32323p+0, 0x1.a0c667b5de565p+0, 0x1.a309bec4a2d33p+0,
    0x1.a5503b23e255dp+0, 0x1.a799e1330b358p+0, 0x1.a9e6b5579fdbfp+0, 0x1.ac36bbfd3f37ap+0,
    0x1.ae89f995ad3adp+0, 0x1.b0e07298db666p+0, 0x1.b33a2b84f15fbp+0, 0x1.b59728de5593ap+0,
    0x1.b7f76f2fb5e47p+0, 0x1.ba5b030a1064ap+0, 0x1.bcc1e904bc1d2p+0, 0x1.bf2c25bd71e09p+0,
    0x1.c199bdd85529cp+0, 0x1.c40ab5fffd07ap+0, 0x1.c67f12e57d14bp+0, 0x1.c8f6d9406e7b5p+0,
    0x1.cb720dcef9069p+0, 0x1.cdf0b555dc3fap+0, 0x1.d072d4a07897cp+0, 0x1.d2f87080d89f2p+0,
    0x1.d5818dcfba487p+0, 0x1.d80e316c98398p+0, 0x1.da9e603db3285p+0, 0x1.dd321f301b460p+0,
    0x1.dfc97337b9b5fp+0, 0x1.e264614f5a129p+0, 0x1.e502ee78b3ff6p+0, 0x1.e7a51fbc74c83p+0,
    0x1.ea4afa2a490dap+0, 0x1.ecf482d8e67f1p+0, 0x1.efa1bee615a27p+0, 0x1.f252b376bba97p+0,
    0x1.f50765b6e4540p+0, 0x1.f7bfdad9cbe14p+0, 0x1.fa7c1819e90d8p+0, 0x1.fd3c22b8f71f1p+0};

// constants of exp_tab in a constant-bank array: the FP64 instructions take them as c[][]
// operands (literals would be materialised with two uniform moves each, per use)
PYR_EXP_CONST double kExpK[7] = {
    0x1.71547652b82fep+7,        // 128 / ln 2
    0x1.62e42fe000000p-8,        // ln 2 / 128, 24 trailing zero bits: n * hi exact
    0x1.f473de6af278fp-37,       // ln 2 / 128 - hi
    0x1.8p52,                    // 2^52 + 2^51: rounds to nearest integer
    1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0};

// `tab`: the table above (the kernels keep a copy in shared memory: a per-thread index into
// constant memory would serialise).  Valid for |x| < 700.
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
    // P(r) = r + r^2 (1/2 + r (1/6 + r (1/24 + r / 120)))
    double p = fma(r, kExpK[4], kExpK[5]);
    p = fma(p, r, kExpK[6]);
    p = fma(p, r, 0.5);
    p = fma(p * r, r, r);
    const double tj = tab[n & (kExpTabSize - 1)];
    const double res = fma(tj, p, tj);
    const int k = n >> 7;                             // arithmetic shift: floor
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

// valid range of exp_tab (result and 2^k normal)
__host__ __device__ __forceinline__ bool exp_tab_ok(double x) { return x > -700.0 && x < 700.0; }

// exp for ANY argument: exp_tab inside (-700, 700); 0 below (e^-700 = 1e-304 stands for 0 next
// to any index profile), NaN above and for NaN (either way the ray cannot stay valid).  The
// range test reads the argument's high word with integer instructions, in parallel with the
// arithmetic (whose out-of-range result is discarded), so it adds nothing to the FP64 chain.
__host__ __device__ __forceinline__ double exp_tab_any(double x, const double *tab) {
    const double r = exp_tab(x, tab);
#ifdef __CUDA_ARCH__
    const int hi = __double2hiint(x);
#else
    long long xb;
    __builtin_memcpy(&xb, &x, 8);
    const int hi = (int)(xb >> 32);
#endif
    const int mag = hi & 0x7fffffff;
    if (mag < 0x4085e000) return r;                   // |x| < 700
    const bool below = hi < 0 && mag <= 0x7ff00000;   // -inf .. -700
#ifdef __CUDA_ARCH__
    return below ? 0.0 : __longlong_as_double(0x7ff8000000000000LL);
#else
    return below ? 0.0 : __builtin_nan("");
#endif
}

}  // namespace pyr
