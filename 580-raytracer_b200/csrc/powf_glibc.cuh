// powf exactly as the reference's toolchain evaluates it.
//
// The reference calls std::powf / std::pow(float,float) in two places: the Phong
// specular term (Raytracer.cpp:253) and the gamma encode (Raytracer.cpp:816-818).  Its
// result is truncated to an integer right afterwards (Pixel(Vector3), Raytracer.h:376-379),
// so "a few ulp" is not good enough: the product must return the same float as the libm
// the oracle links (glibc >= 2.28, sysdeps/ieee754/flt-32/e_powf.c - the Szabolcs Nagy /
// ARM "optimized routines" algorithm; glibc is a system dependency that is not vendored
// in /root/reference, version here: Ubuntu GLIBC 2.39).  That routine is NOT correctly
// rounded (documented bound 0.82 ULP), so it is restated here operation for operation:
//   log2(x)  : 16-entry table {invc, logc} + degree-5 polynomial, in double
//   y*log2(x): one double product
//   exp2     : 32-entry table 2^(i/32) + degree-3 polynomial, round with the 0x1.8p+52/32 shift
// The constants are the published ones of that algorithm; they were checked against
// the .rodata of this image's libm.so.6, and tests/test_powf.py checks the restatement
// against the host powf on ~10^7 inputs (CPU, via powf_host_check.cpp) and the device
// version against the host on the GPU.  x86-64 glibc selects its FMA build of this
// routine (__powf_fma) on every CPU this runs on, so a*b+c is a fused multiply-add here.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define RT580_HD __host__ __device__ __forceinline__
#else
#define RT580_HD static inline
#endif

namespace rt580 {

// __powf_log2_data.tab : {invc, logc}, 16 sub-intervals of [0x1.66p-1, 0x1.66p0)
#define RT580_POWF_LOG2_TAB \
    { 0x1.661ec79f8f3bep+0, -0x1.efec65b963019p-2 }, { 0x1.571ed4aaf883dp+0, -0x1.b0b6832d4fca4p-2 }, \
    { 0x1.49539f0f010bp+0, -0x1.7418b0a1fb77bp-2 },  { 0x1.3c995b0b80385p+0, -0x1.39de91a6dcf7bp-2 }, \
    { 0x1.30d190c8864a5p+0, -0x1.01d9bf3f2b631p-2 }, { 0x1.25e227b0b8eap+0, -0x1.97c1d1b3b7afp-3 },   \
    { 0x1.1bb4a4a1a343fp+0, -0x1.2f9e393af3c9fp-3 }, { 0x1.12358f08ae5bap+0, -0x1.960cbbf788d5cp-4 }, \
    { 0x1.0953f419900a7p+0, -0x1.a6f9db6475fcep-5 }, { 0x1p+0, 0x0p+0 },                               \
    { 0x1.e608cfd9a47acp-1, 0x1.338ca9f24f53dp-4 },  { 0x1.ca4b31f026aap-1, 0x1.476a9543891bap-3 },   \
    { 0x1.b2036576afce6p-1, 0x1.e840b4ac4e4d2p-3 },  { 0x1.9c2d163a1aa2dp-1, 0x1.40645f0c6651cp-2 },  \
    { 0x1.886e6037841edp-1, 0x1.88e9c2c1b9ff8p-2 },  { 0x1.767dcf5534862p-1, 0x1.ce0a44eb17bccp-2 }
// __exp2f_data.tab[i] = bits(2^(i/32)) - (i << 47)
#define RT580_POWF_EXP2_TAB \
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull, \
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull, \
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull, \
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull, \
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull, \
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull, \
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull, \
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull

static const double h_powf_log2_tab[16][2] = { RT580_POWF_LOG2_TAB };
static const uint64_t h_powf_exp2_tab[32] = { RT580_POWF_EXP2_TAB };
#if defined(__CUDACC__)
static __constant__ double c_powf_log2_tab[16][2] = { RT580_POWF_LOG2_TAB };
static __constant__ uint64_t c_powf_exp2_tab[32] = { RT580_POWF_EXP2_TAB };
#endif

RT580_HD void powf_log2_tab(int i, double& invc, double& logc) {
#if defined(__CUDA_ARCH__)
    invc = c_powf_log2_tab[i][0]; logc = c_powf_log2_tab[i][1];
#else
    invc = h_powf_log2_tab[i][0]; logc = h_powf_log2_tab[i][1];
#endif
}
RT580_HD uint64_t powf_exp2_tab(int i) {
#if defined(__CUDA_ARCH__)
    return c_powf_exp2_tab[i];
#else
    return h_powf_exp2_tab[i];
#endif
}

RT580_HD uint32_t powf_asuint(float f) { union { float f; uint32_t u; } c; c.f = f; return c.u; }
RT580_HD float powf_asfloat(uint32_t u) { union { float f; uint32_t u; } c; c.u = u; return c.f; }
RT580_HD uint64_t powf_asuint64(double f) { union { double f; uint64_t u; } c; c.f = f; return c.u; }
RT580_HD double powf_asdouble(uint64_t u) { union { double f; uint64_t u; } c; c.u = u; return c.f; }

// Main path of __powf for finite x > 0 (normal or subnormal) and finite y; the
// reference only ever calls it with x = max(V.R, 0) in [0,1] and x = c/255 in [0,1].
RT580_HD float powf_glibc_pos(float x, float y) {
    uint32_t ix = powf_asuint(x);
    if (ix < 0x00800000u) {                      // subnormal x: normalise (e_powf.c "ix &= 0x7fffffff; ix -= 23 << 23")
        ix = powf_asuint(x * 0x1p23f);
        ix &= 0x7fffffffu;
        ix -= 23u << 23;
    }
    // log2_inline
    uint32_t tmp = ix - 0x3f330000u;
    int i = (int)((tmp >> (23 - 4)) % 16u);
    uint32_t top = tmp & 0xff800000u;
    uint32_t iz = ix - top;
    int k = (int32_t)top >> 23;
    double invc, logc;
    powf_log2_tab(i, invc, logc);
    double z = (double)powf_asfloat(iz);
    double r = fma(z, invc, -1.0);
    double y0 = logc + (double)k;
    double r2 = r * r;
    double yy = fma(0x1.27616c9496e0bp-2, r, -0x1.71969a075c67ap-2);
    double p = fma(0x1.ec70a6ca7baddp-2, r, -0x1.7154748bef6c8p-1);
    double r4 = r2 * r2;
    double q = fma(0x1.71547652ab82bp0, r, y0);
    q = fma(p, r2, q);
    double logx = fma(yy, r4, q);
    double ylogx = (double)y * logx;
    if (((powf_asuint64(ylogx) >> 47) & 0xffff) >= (powf_asuint64(126.0) >> 47)) {
        if (ylogx > 0x1.fffffffd1d571p+6) return INFINITY;           // overflow
        if (ylogx <= -150.0) return 0.0f;                            // underflow
    }
    // exp2_inline
    const double SHIFT = 0x1.8p+52 / 32;
    double kd = ylogx + SHIFT;
    uint64_t ki = powf_asuint64(kd);
    kd -= SHIFT;
    double rr = ylogx - kd;
    uint64_t t = powf_exp2_tab((int)(ki % 32u));
    t += ki << (52 - 5);
    double s = powf_asdouble(t);
    double zz = fma(0x1.c6af84b912394p-5, rr, 0x1.ebfce50fac4f3p-3);
    double rr2 = rr * rr;
    double out = fma(0x1.62e42ff0c52d6p-1, rr, 1.0);
    out = fma(zz, rr2, out);
    out = out * s;
    return (float)out;
}

// powf for the domain the hot path uses: x >= 0 (x = fmax(dot, 0)), any finite y.
RT580_HD float powf_glibc(float x, float y) {
    if (y == 0.0f) return 1.0f;                       // pow(x, +-0) = 1 for every x, NaN included
    if (x == 1.0f) return 1.0f;
    if (x != x || y != y) return x + y;
    if (x == 0.0f) return (y > 0.0f) ? 0.0f : INFINITY;   // x = +0 (x is never -0: fmax(.,0))
    if (isinf(x)) return (y > 0.0f) ? INFINITY : 0.0f;
    if (isinf(y)) {
        if (x < 1.0f) return (y > 0.0f) ? 0.0f : INFINITY;
        return (y > 0.0f) ? INFINITY : 0.0f;
    }
    return powf_glibc_pos(x, y);
}

}  // namespace rt580
