// Device restatement of the reference's value types (Raytracer.h:39-149 Vector3,
// h:373-418 Pixel) with the reference's exact operation order.  The translation unit is
// compiled with -fmad=false -prec-div=true -prec-sqrt=true -ftz=false, so every * + / sqrt
// below is one correctly rounded IEEE fp32 operation, exactly like the reference's
// /fp:precise (or g++ -ffp-contract=off) build.  Anything that may use FMA says so
// explicitly with __fmaf_rn.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rt580 {

// (double)x < 1e-6  <=>  x <= RT_EPS_F ;  (double)x > 1e-6  <=>  x > RT_EPS_F
// because float(1e-6) = 0x358637BD = 9.99999997e-7 is the largest float below the double
// literal EPSILON (Raytracer.h:12).
#define RT_EPS_F 9.99999997e-7f
#define RT_SHADOW_OFFSET 0.2f   // h:13, converted to float at Vector3::operator*(float)
#define RT_IOR 2.5f             // h:460 (Q5)

struct V3 { float x, y, z; };

__host__ __device__ __forceinline__ V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__host__ __device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }   // h:97
__host__ __device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }   // h:93
__host__ __device__ __forceinline__ V3 operator*(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }     // h:83
__host__ __device__ __forceinline__ V3 operator*(V3 a, V3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }  // h:88
__host__ __device__ __forceinline__ V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }                       // h:101
__host__ __device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }       // h:131
__host__ __device__ __forceinline__ V3 cross(V3 a, V3 b) {                                                     // h:122
    return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__host__ __device__ __forceinline__ float length(V3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }   // h:118
__host__ __device__ __forceinline__ V3 normalize(V3 a) {                                                       // h:109-116
    float len = sqrtf(a.x * a.x + a.y * a.y + a.z * a.z);
    if (len > 0) { a.x /= len; a.y /= len; a.z /= len; }
    return a;
}
__host__ __device__ __forceinline__ V3 reflect(V3 I, V3 N) {                                                   // h:143-148
    float IDotN = dot(I, N);
    IDotN *= 2;
    return I - N * IDotN;
}

// ---- Pixel (h:373-418): 16-bit integer colour algebra ---------------------------------
struct Pix { int16_t r, g, b; };

// static_cast<short>(float) as x86 does it: cvttss2si to int32 (0x80000000 when out of
// range or NaN), then the low 16 bits.
__device__ __forceinline__ int16_t f2short(float v) {
    int i = (fabsf(v) < 2147483648.0f) ? __float2int_rz(v) : (int)0x80000000;
    return (int16_t)i;
}
__device__ __forceinline__ Pix mkpix(int r, int g, int b) { Pix p; p.r = (int16_t)r; p.g = (int16_t)g; p.b = (int16_t)b; return p; }
__device__ __forceinline__ Pix pix_from_v3(V3 c) {       // h:376-381: no clamp (Q2)
    return mkpix(f2short(c.x * 255), f2short(c.y * 255), f2short(c.z * 255));
}
__device__ __forceinline__ int16_t clamp255(int16_t v) { return (v > 255) ? (int16_t)255 : (v < 0 ? (int16_t)0 : v); }
__device__ __forceinline__ Pix pix_clamp(Pix p) { return mkpix(clamp255(p.r), clamp255(p.g), clamp255(p.b)); }   // h:411-417
__device__ __forceinline__ Pix pix_muls(Pix p, float s) {                                                      // h:394-400
    return pix_clamp(mkpix(f2short((float)p.r * s), f2short((float)p.g * s), f2short((float)p.b * s)));
}
__device__ __forceinline__ Pix pix_add(Pix a, Pix b) {                                                         // h:403-409
    return mkpix((int16_t)(a.r + b.r), (int16_t)(a.g + b.g), (int16_t)(a.b + b.b));
}

__device__ __forceinline__ float clipf(float input, int mn, int mx) {   // cpp:206-210
    if (input < (float)mn) return (float)mn;
    if (input > (float)mx) return (float)mx;
    return input;
}

}  // namespace rt580
