// Device restatement of the shading half of the hot path:
//   CalculateLocalColor   Raytracer.cpp:213-267 (+ InterpolateVector3 cpp:333-338, Clipf cpp:206-210)
//   ComputeFresnel        Raytracer.cpp:131-166
//   CalculateRefraction   Raytracer.cpp:168-203
//   RandomUnitVector / RandomInHemisphere  Raytracer.cpp:269-292 over the libstdc++ stream
//   (minstd_rand0, one engine step per uniform_real_distribution<float> draw)
#pragma once
#include "rt_math.cuh"
#include "powf_glibc.cuh"

namespace rt580 {

struct Material { V3 Cs; float Ka, Kd, Ks, Kt, n; };
struct Light { int type; V3 color; float intensity; V3 position, direction; };

__device__ __forceinline__ Material load_material(const float* __restrict__ m, int id) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(m + 8 * (size_t)id));
    const float4 b = __ldg(reinterpret_cast<const float4*>(m + 8 * (size_t)id + 4));
    Material r; r.Cs = mk(a.x, a.y, a.z); r.Ka = a.w; r.Kd = b.x; r.Ks = b.y; r.Kt = b.z; r.n = b.w;
    return r;
}
__device__ __forceinline__ Light load_light(const int32_t* __restrict__ types, const float* __restrict__ f, int i) {
    Light L; const float* p = f + 10 * i;
    L.type = __ldg(types + i);
    L.color = mk(__ldg(p), __ldg(p + 1), __ldg(p + 2)); L.intensity = __ldg(p + 3);
    L.position = mk(__ldg(p + 4), __ldg(p + 5), __ldg(p + 6));
    L.direction = mk(__ldg(p + 7), __ldg(p + 8), __ldg(p + 9));
    return L;
}

// cpp:131-166 (the exact unpolarised formula; Q21)
__device__ __forceinline__ void compute_fresnel(float ior, V3 normal, V3 incident, float& Kr, float& Kt) {
    float cosi = clipf(dot(incident, normal), -1, 1);
    const bool inside = cosi > 0;
    float eta_i = 1, eta_t = ior;
    if (inside) { float tmp = eta_i; eta_i = eta_t; eta_t = tmp; cosi = -cosi; }
    const float sint = eta_i / eta_t * sqrtf(fmaxf(0.f, 1 - cosi * cosi));
    if (sint >= 1) { Kr = 1; Kt = 0; }
    else {
        const float cost = sqrtf(fmaxf(0.f, 1 - sint * sint));
        cosi = fabsf(cosi);
        const float Rs = ((eta_t * cosi) - (eta_i * cost)) / ((eta_t * cosi) + (eta_i * cost));
        const float Rp = ((eta_i * cosi) - (eta_t * cost)) / ((eta_i * cosi) + (eta_t * cost));
        Kr = (Rs * Rs + Rp * Rp) / 2;
        Kt = 1 - Kr;
    }
}

// cpp:168-203: zero vector on total internal reflection (Q20)
__device__ __forceinline__ V3 calculate_refraction(V3 I, V3 N, float indexM2) {
    float cosi = dot(I, N);
    if (cosi < -1) cosi = -1; else if (cosi > 1) cosi = 1;
    float m1 = 1, m2 = indexM2;
    V3 n = N;
    if (cosi < 0) cosi = -1 * cosi;
    else { float tmp = m1; m1 = m2; m2 = tmp; n = -N; }
    const float eta = m1 / m2;
    const float k = 1 - eta * eta * (1 - cosi * cosi);
    if (k < 0) return mk(0, 0, 0);
    return I * eta + n * (eta * cosi - sqrtf(k));
}

// cpp:213-267.  `normal` = normalize(shading normal) (cpp:237: the interpolated vertex normal of a triangle,
// normalised once already by InterpolateVector3, or the geometric hit normal of a sphere) and `view` =
// normalize(camera.from - hitPoint) (cpp:249-250, Q8) do not depend on the light: the caller computes them
// once per node (phong_frame) instead of once per light.
struct PhongFrame { V3 normal, view; };
__device__ __forceinline__ PhongFrame phong_frame(V3 hitPoint, V3 shading_normal, V3 cam_from) {
    PhongFrame f;
    f.normal = normalize(shading_normal);                                // cpp:237
    f.view = normalize(cam_from - hitPoint);                             // cpp:249-250 (Q8)
    return f;
}
__device__ __forceinline__ Pix calculate_local_color(V3 hitPoint, const PhongFrame& f, const Light& L, const Material& M) {
    V3 lightVector;
    if (L.type == 1) lightVector = normalize(L.position - hitPoint);     // cpp:215-218
    else lightVector = normalize(L.direction * -1.0f);                   // cpp:220-221
    // fmax(double(x), 0) back to float == x > 0 ? x : +0 (fmax drops a NaN)
    const float ldn = dot(lightVector, f.normal);
    const float diffuseStrength = (ldn > 0.0f) ? ldn : 0.0f;             // cpp:242
    const V3 diffuse = (L.color * diffuseStrength) * L.intensity;        // cpp:243
    const V3 reflection = normalize(reflect(lightVector, f.normal));     // cpp:246-247 (Q9)
    const float vdr = dot(f.view, reflection);
    float spec = (vdr > 0.0f) ? vdr : 0.0f;                              // cpp:252
    spec = powf_glibc(spec, M.n);                                        // cpp:253
    const V3 specular = (L.color * spec) * L.intensity;                  // cpp:254
    const V3 lighting = diffuse * M.Kd + specular * M.Ks;                // cpp:256
    V3 color = M.Cs * lighting;                                          // cpp:258
    color.x = clipf(color.x, 0, 1); color.y = clipf(color.y, 0, 1); color.z = clipf(color.z, 0, 1);
    return pix_from_v3(color);                                           // cpp:264
}
__device__ __forceinline__ Pix calculate_local_color(V3 hitPoint, V3 shading_normal, const Light& L,
                                                     const Material& M, V3 cam_from) {
    return calculate_local_color(hitPoint, phong_frame(hitPoint, shading_normal, cam_from), L, M);
}

// ---- the AO sample stream ----------------------------------------------------------------
#define RT_LCG_M 2147483647u
__device__ __forceinline__ uint32_t lcg_mulmod(uint32_t a, uint32_t b) {
    // a*b mod (2^31-1) for a,b < 2^31-1, Mersenne folding (no division)
    const uint64_t p = (uint64_t)a * b;
    uint32_t r = (uint32_t)(p & RT_LCG_M) + (uint32_t)(p >> 31);
    r = (r & RT_LCG_M) + (r >> 31);
    return (r == RT_LCG_M) ? 0u : r;
}
// engine state after `steps` steps from seed 1: 16807^steps mod (2^31-1)
__device__ __forceinline__ uint32_t lcg_state_at(uint64_t steps) {
    uint32_t e = (uint32_t)(steps % (uint64_t)(RT_LCG_M - 1u));
    uint32_t base = 16807u, r = 1u;
    while (e) { if (e & 1u) r = lcg_mulmod(r, base); base = lcg_mulmod(base, base); e >>= 1; }
    return r;
}
// same through a table of 16807^(d * 256^k), d < 256, k < 4 (tab[k * 256 + d]): 3 modular products instead of ~45
__device__ __forceinline__ uint32_t lcg_state_at_tab(uint64_t steps, const uint32_t* __restrict__ tab) {
    const uint32_t e = (uint32_t)(steps % (uint64_t)(RT_LCG_M - 1u));
    uint32_t r = __ldg(tab + (e & 255u));
    r = lcg_mulmod(r, __ldg(tab + 256 + ((e >> 8) & 255u)));
    r = lcg_mulmod(r, __ldg(tab + 512 + ((e >> 16) & 255u)));
    return lcg_mulmod(r, __ldg(tab + 768 + (e >> 24)));
}
__device__ __forceinline__ float lcg_canonical(uint32_t& st) {
    st = lcg_mulmod(st, 16807u);
    float u = __uint2float_rn(st - 1u) / 2147483648.0f;      // generate_canonical<float,24>
    if (u >= 1.0f) u = 0.99999994f;                           // nextafterf(1, 0)
    return u;
}
// cpp:269-292
__device__ __forceinline__ V3 random_in_hemisphere(uint32_t& st, V3 normal) {
    const float two_pi = 6.2831853f;                          // (float)(2 * 3.14159265)
    const float z = lcg_canonical(st) * 2.0f + (-1.0f);       // cpp:273 (zDist first)
    const float a = lcg_canonical(st) * two_pi + 0.0f;        // cpp:274
    const float r = sqrtf(1 - z * z);                         // cpp:275
    double sn, cs;
    sincos((double)a, &sn, &cs);                              // libm double cos/sin (Q27)
    const float x = (float)((double)r * cs);                  // cpp:277
    const float y = (float)((double)r * sn);                  // cpp:278
    const V3 v = normalize(mk(x, y, z));                      // cpp:285
    if (dot(v, normal) > 0.0f) return v;                      // cpp:286
    return -v;
}

}  // namespace rt580
