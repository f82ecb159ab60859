// Point-light clearance maps: a conservative "nothing is nearer than this" cube map per point light.
//
// The reference's shadow ray for a point light (Raytracer.cpp:62-75) starts at P + 0.2 L^ and is
// accepted as "lit" unless a primitive is hit at t <= |light - P|; it runs along the line through the
// light.  Every hit the reference can accept near a primitive lies inside that primitive's padded
// leaf box (bvh_build.cu: that is what makes the tree exact).  So if no leaf box reaches into the
// segment between the ray origin and the light, the tree cannot contribute an occluder and the
// traversal can be skipped.  The map stores, per direction texel seen from the light, a lower bound
// of the distance to the nearest leaf box whose projection touches the texel (dilated by one texel
// for the rounding of directions); "box distance > |origin - light|" then proves the segment clear.
// The 0.2 units by which the ray overshoots the light are covered by requiring that no box comes
// nearer than SMAP_CLEARANCE to the light at all (otherwise the light gets no map).
// Large primitives (walls) are not in the tree and not in the map: they keep their own exact test.
#pragma once
#include "device_scene.h"
#include "rt_math.cuh"

namespace rt580 {

#define SMAP_RES_DEFAULT 512   // texels per cube-face edge (RT580_SMAP_RES overrides): 0.18 degrees per texel
#define SMAP_CLEARANCE 0.25f
#define SMAP_MAX 8           // point lights that get a map

__device__ __forceinline__ float smap_axis(const V3& v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }

// one thread per (inner node, child slot): leaf children rasterise their box into the six faces
__global__ void __launch_bounds__(256)
k_smap_raster(const BvhNode* __restrict__ nodes, int n_nodes, float lx, float ly, float lz, float* __restrict__ map,
              unsigned int* __restrict__ min_clear_bits, int SMAP_RES)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = tid >> 1, c = tid & 1;
    if (i >= n_nodes) return;
    const int4 kids = __ldg(&nodes[i].kids);
    if ((c == 0 ? kids.x : kids.y) >= 0) return;                     // inner child
    const float4 xy = __ldg(c == 0 ? &nodes[i].xy0 : &nodes[i].xy1), z01 = __ldg(&nodes[i].z01);
    V3 lo = mk(xy.x - lx, xy.z - ly, (c == 0 ? z01.x : z01.z) - lz);
    V3 hi = mk(xy.y - lx, xy.w - ly, (c == 0 ? z01.y : z01.w) - lz);
    if (lo.x > hi.x || lo.y > hi.y || lo.z > hi.z) return;           // the empty box of a one-primitive tree
    // subtraction rounds: widen by an ulp-scale margin so the box relative to the light still contains the true one
    const float m = 1e-6f * (fabsf(lx) + fabsf(ly) + fabsf(lz) + fabsf(hi.x) + fabsf(hi.y) + fabsf(hi.z) + fabsf(lo.x) + fabsf(lo.y) + fabsf(lo.z));
    lo = mk(lo.x - m, lo.y - m, lo.z - m); hi = mk(hi.x + m, hi.y + m, hi.z + m);
    const float dx = fmaxf(fmaxf(lo.x, -hi.x), 0.f), dy = fmaxf(fmaxf(lo.y, -hi.y), 0.f), dz = fmaxf(fmaxf(lo.z, -hi.z), 0.f);
    const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
    atomicMin(min_clear_bits, __float_as_uint(dist));
    const float depth = dist * 0.99999f;
    const unsigned dbits = __float_as_uint(depth);
    for (int face = 0; face < 6; face++) {
        const int a = face >> 1, ua = (a + 1) % 3, va = (a + 2) % 3;
        const bool neg = face & 1;
        float zlo = neg ? -smap_axis(hi, a) : smap_axis(lo, a);
        const float zhi = neg ? -smap_axis(lo, a) : smap_axis(hi, a);
        if (!(zhi > 0.f)) continue;
        zlo = fmaxf(zlo, 1e-30f);
        const float ulo = smap_axis(lo, ua), uhi = smap_axis(hi, ua), vlo = smap_axis(lo, va), vhi = smap_axis(hi, va);
        float umin = ulo >= 0.f ? ulo / zhi : ulo / zlo, umax = uhi >= 0.f ? uhi / zlo : uhi / zhi;
        float vmin = vlo >= 0.f ? vlo / zhi : vlo / zlo, vmax = vhi >= 0.f ? vhi / zlo : vhi / zhi;
        if (umin > 1.f || umax < -1.f || vmin > 1.f || vmax < -1.f) continue;
        umin = fmaxf(umin, -1.f); umax = fminf(umax, 1.f); vmin = fmaxf(vmin, -1.f); vmax = fminf(vmax, 1.f);
        const int iu0 = max(0, (int)floorf((umin + 1.f) * (0.5f * SMAP_RES)) - 1), iu1 = min(SMAP_RES - 1, (int)floorf((umax + 1.f) * (0.5f * SMAP_RES)) + 1);
        const int iv0 = max(0, (int)floorf((vmin + 1.f) * (0.5f * SMAP_RES)) - 1), iv1 = min(SMAP_RES - 1, (int)floorf((vmax + 1.f) * (0.5f * SMAP_RES)) + 1);
        unsigned int* f = reinterpret_cast<unsigned int*>(map) + (size_t)face * SMAP_RES * SMAP_RES;
        for (int iv = iv0; iv <= iv1; iv++)
            for (int iu = iu0; iu <= iu1; iu++) atomicMin(f + (size_t)iv * SMAP_RES + iu, dbits);
    }
}

// Does the map prove that no tree primitive lies between `origin` (the shadow ray's start) and the light?
__device__ __forceinline__ bool smap_clear(const float* __restrict__ map, int SMAP_RES, V3 light, V3 origin)
{
    const V3 v = origin - light;
    const float ax = fabsf(v.x), ay = fabsf(v.y), az = fabsf(v.z);
    const int a = (ax >= ay && ax >= az) ? 0 : (ay >= az ? 1 : 2);
    const float za = smap_axis(v, a);
    const float z = fabsf(za);
    if (!(z > 0.f)) return false;
    const int ua = (a + 1) % 3, va = (a + 2) % 3;
    const float u = smap_axis(v, ua) / z, w = smap_axis(v, va) / z;
    const int iu = min(SMAP_RES - 1, max(0, (int)floorf((u + 1.f) * (0.5f * SMAP_RES))));
    const int iv = min(SMAP_RES - 1, max(0, (int)floorf((w + 1.f) * (0.5f * SMAP_RES))));
    const int face = a * 2 + (za < 0.f ? 1 : 0);
    const float depth = __ldg(map + ((size_t)face * SMAP_RES + iv) * SMAP_RES + iu);
    const float q = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
    return depth > q * 1.00001f;
}

}  // namespace rt580
