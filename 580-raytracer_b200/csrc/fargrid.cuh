// Far-field direction grid: which triangles can a ray that leaves the scene still "hit"?
//
// The reference's triangle test (Raytracer.cpp:348-409) accepts a plane hit P when three float-evaluated
// signed areas are non-negative (cpp:392-396).  Far from the triangle the first of them,
//     da = dot(cross(v1 - P, v2 - P), N)                                              (cpp:392, 937-942)
// is the cross product of two long, almost parallel vectors: rounding noise.  So a ray that runs almost
// parallel to a triangle's plane "hits" it 10^5..10^8 units away about every other time, provided P lies
// in the infinite wedge at v0 (cpp:393-394 are reliable).  The reference reports those hits (AO rays,
// shadow rays of a directional light, rays that see the sky), hence so must this library - exactly.
//
// Bound used here (derivation in DESIGN.md section 2.1; checked against the float arithmetic on 2.4e7
// adversarial samples, tools/far_bound_check.py).  With u = 2^-24, q = P - v1, e = v2 - v1, m = e x N,
//     da = -q.m + err,   |err| <= 6.0001 u G(q, N) + 9.3 u |q| |e|,   G(q, N) = |qy qz Nx| + |qz qx Ny| + |qx qy Nz|
// and P beyond the edge v1v2 is accepted only if |q.m| <= |err|.  With w = q / |q| (which is the ray
// direction up to |O - v1| / |q|) that is
//     |q| >= L(w) = (|w.m| - 9.3 u |e|) / (6.0001 u g(w, N)),        g(w, N) = G(w, N) <= 0.57741
// i.e. a lower bound T on the ray parameter that depends on the triangle and on the DIRECTION only.
// A far-field hit at parameter t >= T needs |N.d| = |N.O + D| / t <= dmax / T: the ray direction lies
// in a band of half-width thr = dmax / T around the great circle of the triangle's plane, on the arc of
// the wedge.  thr is 1e-5..1e-3: the set of directions that can far-hit one triangle is a thin strip.
//
// The grid: the sphere of ray directions as a cube map of 6 K^2 cells; every cell lists the triangles
// whose strip crosses it, with the bound T evaluated for the cell (6-bit factor over the triangle's
// direction-independent bound).  Great circles are straight lines on a cube face, so the strips are
// rasterised row by row.  A ray that found nothing nearer than far_tmin looks up ITS cell only and
// runs a two-stage filter over that list (|N.d| <= thr, then |N.d| T <= |N.O + D| with the ray's own
// origin); what survives gets the reference's exact test (trace.cuh prim_test).  Everything before the
// exact test is a necessary condition with explicit slack: it prunes, it never decides.
//
// Lists hold ~0.24 n / K triangles (n = 10^6, K = 1024: ~240) instead of the 10^6 filter records the
// O(n) scan of round 1 walked per escaping ray.
#pragma once
#include "device_scene.h"
#include "rt_math.cuh"

namespace rt580 {

#define FG_ID_BITS 26
#define FG_ID_MASK ((1u << FG_ID_BITS) - 1u)
#define FG_MAX_PRIMS (1 << FG_ID_BITS)
#define FG_U 5.9604644775390625e-8          /* 2^-24 */
#define FG_ND_SLACK 4.0e-7f                 /* |N.d| evaluated with FMA here vs unfused in the reference: <= 6u apart */
#define FG_ND_MIN 6.0e-7f                   /* |N.d| below this is below EPSILON in the reference as well (cpp:371) */
#define FG_WIDE_FACTOR 8.0f                 /* triangles whose far field begins nearer than this many scene diagonals: "wide" list */

// in-scene ray origins: the box around all primitives (+ padding + the 0.2 offset of cpp:67/98/110/322) and the camera
__device__ __forceinline__ bool in_scene(const DeviceScene& sc, V3 O) {
    const bool in_box = O.x >= sc.ob_lo[0] && O.x <= sc.ob_hi[0] && O.y >= sc.ob_lo[1] && O.y <= sc.ob_hi[1] &&
                        O.z >= sc.ob_lo[2] && O.z <= sc.ob_hi[2];
    return in_box || (O.x == sc.ob_cam[0] && O.y == sc.ob_cam[1] && O.z == sc.ob_cam[2]);
}

__host__ __device__ __forceinline__ float fg_axis(const V3& v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }

// cell of a direction: face = 2 * major axis + (negative ? 1 : 0); (u, v) = the other two components over |major|
__device__ __forceinline__ int fg_cell_of_dir(V3 d, int K) {
    const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    const int a = (ax >= ay && ax >= az) ? 0 : (ay >= az ? 1 : 2);
    const float w = fg_axis(d, a);
    const float aw = fabsf(w);
    if (!(aw > 0.f)) return -1;                                   // zero / NaN direction: no cell (such rays never need the far field)
    const float fu = fg_axis(d, (a + 1) % 3) / aw, fv = fg_axis(d, (a + 2) % 3) / aw;
    const float half = 0.5f * (float)K;
    const int iu = min(K - 1, max(0, (int)floorf((fu + 1.f) * half)));
    const int iv = min(K - 1, max(0, (int)floorf((fv + 1.f) * half)));
    return ((a * 2 + (w < 0.f ? 1 : 0)) * K + iv) * K + iu;
}

// the same in double, for the walk along the arc of a ray from outside the scene (rt580_core.cu k_fg_arc): the cell must agree
// with the walls the walk computes in double
__device__ __forceinline__ int fg_cell_of_point_d(double x, double y, double z, int K) {
    const double ax = fabs(x), ay = fabs(y), az = fabs(z);
    const int a = (ax >= ay && ax >= az) ? 0 : (ay >= az ? 1 : 2);
    const double w = a == 0 ? x : (a == 1 ? y : z);
    const double aw = fabs(w);
    if (!(aw > 0.0)) return -1;
    const double fu = (a == 0 ? y : (a == 1 ? z : x)) / aw, fv = (a == 0 ? z : (a == 1 ? x : y)) / aw;
    const double half = 0.5 * (double)K;
    const int iu = min(K - 1, max(0, (int)floor((fu + 1.0) * half)));
    const int iv = min(K - 1, max(0, (int)floor((fv + 1.0) * half)));
    return ((a * 2 + (w < 0.0 ? 1 : 0)) * K + iv) * K + iu;
}

// T for a list entry: the triangle's bound times 2^(k6 / 4), a hair below
__device__ __forceinline__ float fg_entry_T(float Ti, unsigned k6) {
    // 2^(k6 / 4) = 2^(k6 >> 2) * 2^((k6 & 3) / 4): a power of two built from its exponent bits, times one of four constants
    // (the build rounds k6 down with 1e-3 to spare in the exponent, fargrid_build.cu: the constants' last bit does not matter)
    const float frac = (k6 & 2u) ? ((k6 & 1u) ? 1.6817928f : 1.4142135f) : ((k6 & 1u) ? 1.1892071f : 1.0f);
    return Ti * (frac * __uint_as_float((127u + (k6 >> 2)) << 23)) * 0.999998f;
}

}  // namespace rt580
