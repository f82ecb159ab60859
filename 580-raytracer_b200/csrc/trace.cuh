// Ray / primitive tests and BVH traversal.  Device restatement of
//   IntersectTriangle  Raytracer.cpp:348-409 (+ CalcTriangleAreaSigned cpp:937-942, NearlyEquals cpp:16-18)
//   IntersectSphere    Raytracer.cpp:419-464 (+ GreaterThanZero Raytracer.h:558-560)
//   IntersectScene     Raytracer.cpp:473-526
// The leaf tests are the reference's arithmetic operation for operation (the TU is built
// with -fmad=false); the LBVH above them only prunes.  Closest hit = lexicographic minimum
// of (t, primitive order index), which is what the reference's first-wins strict '<' loop
// returns (cpp:494, cpp:513).
#pragma once
#include "device_scene.h"
#include "rt_math.cuh"
#include <cfloat>

namespace rt580 {

// The LBVH's depth is bounded by its keys: 63 Morton bits, then the 32 index bits that order equal keys (bvh_build.cu delta()):
// a stack of 96 entries holds any tree the build can produce (clustered centroids, thousands of coincident instances).
#define RT_STACK_SIZE 96
#define RT_SLAB_WIDEN 4.76837158e-7f   // 2^-21: > 3 roundings of (plane - o) * (1/d), see DESIGN.md

// 32 bytes per lane in one instruction (LDG.E.256, sm_100): a 64-byte node or primitive record is two
// requests to the L1 instead of four.  ncu had the any-hit traversal at 72 % of the L1's wavefront rate.
__device__ __forceinline__ void ldg256(const void* p, float4& a, float4& b) {
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
// one inner node: both child boxes and the child indices
__device__ __forceinline__ void load_node(const BvhNode* __restrict__ nd, float4& xy0, float4& xy1, float4& z01, int4& kids) {
    float4 k;
    ldg256(&nd->xy0, xy0, xy1);
    ldg256(&nd->z01, z01, k);
    kids = make_int4(__float_as_int(k.x), __float_as_int(k.y), 0, 0);
}

struct HitRec {
    float t;
    int leaf;    // index into DeviceScene::prims
    int prim;    // primitive order index (tie-break key, shading lookup)
};

// true iff (float)(0.5 * (double)dval) / total < 0, i.e. the reference's barycentric sign test
// (cpp:392-396 with cpp:941's double 0.5).  Fast path: no underflow possible => sign logic.
__device__ __forceinline__ bool bary_negative(float dval, float total, bool slow) {
    if (!slow && fabsf(dval) >= 1e-30f) return (dval < 0.0f) != (total < 0.0f);
    float h = 0.5f * dval;          // == (float)(0.5 * (double)dval)
    return (h / total) < 0.0f;
}

// Does the reference accept this primitive for the ray, and at which t?  (no best-so-far logic)
// t_limit: candidates with t > t_limit (or, at t == t_limit, prim >= prim_limit) are not worth
// the barycentric work because the caller would discard them anyway.
template <bool GLOBAL> __device__ __forceinline__ float4 ld4(const float4* p) {
    if (GLOBAL) return __ldg(p);      // read-only path (LDG.E.128.CONSTANT)
    return *p;                        // shared-memory staged records
}

template <bool GLOBAL>
__device__ __forceinline__ bool prim_test(const PrimRec* __restrict__ p, V3 O, V3 d, float t_limit, int prim_limit,
                                          float& t_out, int& prim_out)
{
    const float4 ra = ld4<GLOBAL>(&p->a);
    const float4 rd = ld4<GLOBAL>(&p->d);
    const float4 rc = ld4<GLOBAL>(&p->c);
    const int flags = __float_as_int(rd.w);
    const int prim = __float_as_int(rc.w);
    if (!(flags & RT_PRIM_SPHERE)) {
        const V3 N = mk(rd.x, rd.y, rd.z);
        const float nd = dot(N, d);                               // cpp:367
        if (fabsf(nd - 0.0f) <= RT_EPS_F) return false;           // cpp:371
        const float num = -(dot(N, O) + ra.w);
        // two rejections that need no division (an IEEE division keeps the sign, and is within half an
        // ulp of the quotient): the plane lies behind the ray, t <= 0 <= EPSILON (cpp:382); or so far
        // beyond t_limit that the rounded t exceeds it as well (the caller would discard the hit)
        if (num == 0.0f || ((num < 0.0f) != (nd < 0.0f))) return false;
        if (fabsf(num) > (t_limit * fabsf(nd)) * 1.000001f) return false;
        const float t = num / nd;                                 // cpp:381
        if (t <= RT_EPS_F) return false;                          // cpp:382
        if (!(t < t_limit || (t == t_limit && prim < prim_limit))) return false;
        const float4 rb = ld4<GLOBAL>(&p->b);
        const V3 v0 = mk(ra.x, ra.y, ra.z), v1 = mk(rb.x, rb.y, rb.z), v2 = mk(rc.x, rc.y, rc.z);
        const V3 P = O + d * t;                                   // cpp:387
        const bool slow = (flags & RT_PRIM_SLOWPATH) != 0;
        // CalcTriangleAreaSigned(A,B,C,N) = 0.5 * dot(cross(B-A, C-A), N)   (cpp:937-942)
        const float da = dot(cross(v1 - P, v2 - P), N);           // alpha: (P, v1, v2)   cpp:392
        if (bary_negative(da, rb.w, slow)) return false;
        const float db = dot(cross(P - v0, v2 - v0), N);          // beta : (v0, P, v2)   cpp:393
        if (bary_negative(db, rb.w, slow)) return false;
        const float dg = dot(cross(v1 - v0, P - v0), N);          // gamma: (v0, v1, P)   cpp:394
        if (bary_negative(dg, rb.w, slow)) return false;          // cpp:396
        t_out = t; prim_out = prim;
        return true;
    } else {
        const V3 C = mk(ra.x, ra.y, ra.z);
        const V3 oc = O - C;                                      // cpp:421
        const float b = 2.0f * dot(d, oc);                        // cpp:422
        const float c = dot(oc, oc) - (ra.w * ra.w);              // cpp:423
        const float disc = (b * b) - (4.0f * c);                  // cpp:426
        if (disc <= RT_EPS_F) return false;                       // cpp:427
        const float sq = sqrtf(disc);                             // cpp:429
        const float t0 = (-b + sq) / 2.0f;                        // cpp:430
        const float t1 = (-b - sq) / 2.0f;                        // cpp:431
        const bool g0 = t0 > RT_EPS_F, g1 = t1 > RT_EPS_F;        // h:558-560
        if (!g0 && !g1) return false;                             // cpp:433
        const float t = !g0 ? t1 : (!g1 ? t0 : fminf(t0, t1));    // cpp:438-453
        if (!(t < t_limit || (t == t_limit && prim < prim_limit))) return false;
        t_out = t; prim_out = prim;
        return true;
    }
}

// The two halves of the triangle branch of prim_test, for callers that test several triangles of ONE plane
// (bit-identical N and D: device_scene.h big_planes): the plane half gives the same t and P for all of them.
// Same operations in the same order as prim_test; any-hit form (a hit at t == t_limit counts, cpp:75).
__device__ __forceinline__ bool plane_point_any(float4 pl, V3 O, V3 d, float t_limit, V3& P_out)
{
    const V3 N = mk(pl.x, pl.y, pl.z);
    const float nd = dot(N, d);                                   // cpp:367
    if (fabsf(nd - 0.0f) <= RT_EPS_F) return false;               // cpp:371
    const float num = -(dot(N, O) + pl.w);
    if (num == 0.0f || ((num < 0.0f) != (nd < 0.0f))) return false;
    if (fabsf(num) > (t_limit * fabsf(nd)) * 1.000001f) return false;
    const float t = num / nd;                                     // cpp:381
    if (t <= RT_EPS_F) return false;                              // cpp:382
    if (!(t <= t_limit)) return false;
    P_out = O + d * t;                                            // cpp:387
    return true;
}
template <bool GLOBAL>
__device__ __forceinline__ bool tri_bary_accept(const PrimRec* __restrict__ p, V3 P)
{
    const float4 ra = ld4<GLOBAL>(&p->a), rb = ld4<GLOBAL>(&p->b), rc = ld4<GLOBAL>(&p->c), rd = ld4<GLOBAL>(&p->d);
    const V3 N = mk(rd.x, rd.y, rd.z);
    const V3 v0 = mk(ra.x, ra.y, ra.z), v1 = mk(rb.x, rb.y, rb.z), v2 = mk(rc.x, rc.y, rc.z);
    const bool slow = (__float_as_int(rd.w) & RT_PRIM_SLOWPATH) != 0;
    const float da = dot(cross(v1 - P, v2 - P), N);               // alpha: (P, v1, v2)   cpp:392
    if (bary_negative(da, rb.w, slow)) return false;
    const float db = dot(cross(P - v0, v2 - v0), N);              // beta : (v0, P, v2)   cpp:393
    if (bary_negative(db, rb.w, slow)) return false;
    const float dg = dot(cross(v1 - v0, P - v0), N);              // gamma: (v0, v1, P)   cpp:394
    return !bary_negative(dg, rb.w, slow);                        // cpp:396
}

// Conservative ray/box slab test.  (plane - o) * inv has <= 3 roundings, so widening the
// interval by 2^-21 relative makes the float result a superset of the exact one.
__device__ __forceinline__ bool slab(float lox, float hix, float loy, float hiy, float loz, float hiz,
                                     V3 O, V3 inv, float tcull, float& tnear)
{
    const float ax = (lox - O.x) * inv.x, bx = (hix - O.x) * inv.x;
    const float ay = (loy - O.y) * inv.y, by = (hiy - O.y) * inv.y;
    const float az = (loz - O.z) * inv.z, bz = (hiz - O.z) * inv.z;
    float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
    float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
    tn = __fmaf_rn(fabsf(tn), -RT_SLAB_WIDEN, tn);
    tf = __fmaf_rn(fabsf(tf), RT_SLAB_WIDEN, tf);
    tnear = tn;
    return (tn <= tf) && (tf >= 0.0f) && (tn <= tcull);
}

// Closest hit (ANY == false): best = lexicographic min (t, prim) over accepted primitives.
// Any hit     (ANY == true) : true as soon as one primitive is accepted with t <= tmax
//                             (cpp:75 and cpp:325 use only that boolean).
// `best` comes in initialised (t = tmax or +inf / the best of the large-primitive list).
template <bool ANY>
__device__ __forceinline__ bool traverse_bvh(const DeviceScene& sc, V3 O, V3 d, float tmax, HitRec& best,
                                             unsigned* cnt = nullptr)
{
    if (sc.n_leaf <= 0) return false;
    const V3 inv = mk(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    int stack[RT_STACK_SIZE];
    int sp = 0;
    int node = 0;
    bool found = false;
    for (;;) {
        if (cnt) cnt[0]++;
        float4 xy0, xy1, z01; int4 kids;
        load_node(sc.nodes + node, xy0, xy1, z01, kids);
        float tn0, tn1;
        const bool h0 = slab(xy0.x, xy0.y, xy0.z, xy0.w, z01.x, z01.y, O, inv, best.t, tn0);
        const bool h1 = slab(xy1.x, xy1.y, xy1.z, xy1.w, z01.z, z01.w, O, inv, best.t, tn1);
        int next;
        if (h0 && h1) {
            int nearc = kids.x, farc = kids.y;
            if (tn1 < tn0) { nearc = kids.y; farc = kids.x; }
            if (sp < RT_STACK_SIZE) stack[sp++] = farc;   // depth is checked at build time (<= RT_STACK_SIZE)
            next = nearc;
        } else if (h0) next = kids.x;
        else if (h1) next = kids.y;
        else { if (sp == 0) break; next = stack[--sp]; }
        // leaves
        bool done = false;
        while (next < 0) {
            const int leaf = ~next;
            float t; int prim;
            if (cnt) cnt[1]++;
            if (prim_test<true>(sc.prims + leaf, O, d, best.t, ANY ? 0x7fffffff : best.prim, t, prim)) {
                if (ANY) return true;   // prim_test's limit is "t < tmax || t == tmax" (prim_limit = INT_MAX)
                best.t = t; best.leaf = leaf; best.prim = prim; found = true;
            }
            if (sp == 0) { done = true; break; }
            next = stack[--sp];
        }
        if (done) break;
        node = next;
    }
    return found;
}

// Slow path, warp-cooperative.  Two kinds of rays cannot be answered by the tree alone
// (bvh_build.cu header):
//   far scan  a ray that found nothing nearer than far_tmin must replay the reference's
//             far-field acceptances: every triangle whose filter record passes |N.d| <= thr
//             (thr >= |N.O + D| / T_far, so t_plane >= T_far implies the filter passes) gets
//             the reference's exact test;
//   linear    a ray that starts outside the extent the boxes were padded for (the child of a
//             far-field hit, 10^4..10^7 units away) takes the reference's own linear loop.
// Such rays are rare and scattered, so instead of one lane walking ~10^6 records while 31 idle,
// the whole warp serves them one at a time: the ray is broadcast, lane l tests records
// l, l+32, ... (coalesced 512-byte loads), and a lexicographic (t, prim) warp reduction
// returns what the reference's first-wins loop would.  Must be called by all 32 lanes.
// One slow ray served by a whole warp (all 32 lanes call with the same ray): returns whether
// something was accepted; for closest hits `best` holds the lexicographic (t, prim) minimum.
template <bool ANY>
__device__ __forceinline__ bool warp_scan_one(const DeviceScene& sc, bool lin, V3 Ob, V3 db, HitRec& best)
{
    const int lane = threadIdx.x & 31;
    const float4* __restrict__ far = sc.far;
    const int n = sc.n_all;
    best.leaf = -1;
    bool f = false;
    if (!lin) {
        // far scan: 4 filter records per lane in flight (independent loads), exact test on the survivors
        for (int base = 0; base < n && !(ANY && f); base += 128) {
            float4 fr[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int i = base + 32 * k + lane;
                fr[k] = (i < n) ? __ldg(far + i) : make_float4(0.f, 0.f, 0.f, -1.0f);
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float nd = __fmaf_rn(fr[k].x, db.x, __fmaf_rn(fr[k].y, db.y, fr[k].z * db.z));   // filter only: FMA is fine
                if (fabsf(nd) <= fr[k].w) {
                    const int i = base + 32 * k + lane;
                    float t; int prim;
                    if (prim_test<true>(sc.prims + i, Ob, db, best.t, ANY ? 0x7fffffff : best.prim, t, prim)) {
                        best.t = t; best.leaf = i; best.prim = prim; f = true;
                    }
                }
            }
        }
    } else {
        for (int i = lane; i < n; i += 32) {
            float t; int prim;
            if (prim_test<true>(sc.prims + i, Ob, db, best.t, ANY ? 0x7fffffff : best.prim, t, prim)) {
                best.t = t; best.leaf = i; best.prim = prim; f = true;
                if (ANY) break;
            }
        }
    }
    const bool anyf = __any_sync(0xffffffffu, f);
    if (!ANY) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float t2 = __shfl_xor_sync(0xffffffffu, best.t, o);
            const int p2 = __shfl_xor_sync(0xffffffffu, best.prim, o);
            const int l2 = __shfl_xor_sync(0xffffffffu, best.leaf, o);
            if (t2 < best.t || (t2 == best.t && (p2 < best.prim || (p2 == best.prim && l2 > best.leaf)))) {
                best.t = t2; best.prim = p2; best.leaf = l2;
            }
        }
    }
    return anyf;
}

// In-kernel form: the lanes that need the slow path are served one after the other by the
// whole warp (used when no deferred queue is available, or it is full).
template <bool ANY>
__device__ __forceinline__ bool warp_slow_path(const DeviceScene& sc, bool need, bool linear, V3 O, V3 d,
                                               HitRec& hit, bool found)
{
    unsigned pending = __ballot_sync(0xffffffffu, need);
    if (pending == 0u) return found;
    const int lane = threadIdx.x & 31;
    while (pending) {
        const int src = __ffs(pending) - 1;
        pending &= pending - 1u;
        const V3 Ob = mk(__shfl_sync(0xffffffffu, O.x, src), __shfl_sync(0xffffffffu, O.y, src), __shfl_sync(0xffffffffu, O.z, src));
        const V3 db = mk(__shfl_sync(0xffffffffu, d.x, src), __shfl_sync(0xffffffffu, d.y, src), __shfl_sync(0xffffffffu, d.z, src));
        const bool lin = __shfl_sync(0xffffffffu, linear ? 1 : 0, src) != 0;
        HitRec best;
        best.t = __shfl_sync(0xffffffffu, hit.t, src);
        best.prim = __shfl_sync(0xffffffffu, hit.prim, src);
        const bool anyf = warp_scan_one<ANY>(sc, lin, Ob, db, best);
        if (lane == src && anyf) {
            found = true;
            if (!ANY) hit = best;
        }
    }
    return found;
}

// The plane part of the reference's triangle test (cpp:367-382) for one plane record (N.xyz, D), with
// the two division-free rejections of prim_test: can a triangle in this plane still be accepted with
// t within t_limit?  Same operations on the same values as prim_test, hence the same answer for
// every triangle of the plane.
__device__ __forceinline__ bool plane_ahead(float4 pl, V3 O, V3 d, float t_limit) {
    const V3 N = mk(pl.x, pl.y, pl.z);
    const float nd = dot(N, d);                                   // cpp:367
    if (fabsf(nd - 0.0f) <= RT_EPS_F) return false;               // cpp:371
    const float num = -(dot(N, O) + pl.w);
    if (num == 0.0f || ((num < 0.0f) != (nd < 0.0f))) return false;
    return !(fabsf(num) > (t_limit * fabsf(nd)) * 1.000001f);
}

// The large primitives kept out of the tree (DeviceScene::n_big): exact test, first, for every ray,
// plane by plane (device_scene.h).
template <bool ANY>
__device__ __forceinline__ bool big_scan(const DeviceScene& sc, V3 O, V3 d, HitRec& best)
{
    bool found = false;
    for (unsigned long long m = sc.big_sphere_mask; m; m &= m - 1ull) {
        const int i = sc.n_leaf + (__ffsll((long long)m) - 1);
        float t; int prim;
        if (prim_test<true>(sc.prims + i, O, d, best.t, ANY ? 0x7fffffff : best.prim, t, prim)) {
            if (ANY) return true;
            best.t = t; best.leaf = i; best.prim = prim; found = true;
        }
    }
    // Closest hit: the plane the ray reaches first goes first (found with approximate arithmetic, it only
    // orders the exact tests); a hit there bounds t, and plane_ahead then turns away the planes behind it
    // without a triangle test.  In list order a ray met several walls "ahead" and tested all their triangles.
    int first = -1;
    if (!ANY && sc.n_big_planes > 2) {
        unsigned m1 = 0xffffffffu;
        for (int j = 0; j < sc.n_big_planes; j++) {
            const float4 pl = __ldg(sc.big_planes + j);
            const float nd = __fmaf_rn(pl.x, d.x, __fmaf_rn(pl.y, d.y, pl.z * d.z));
            const float num = -__fmaf_rn(pl.x, O.x, __fmaf_rn(pl.y, O.y, __fmaf_rn(pl.z, O.z, pl.w)));
            float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(nd));
            const float ta = num * r;
            const unsigned key = (ta > 0.0f) ? ((__float_as_uint(ta) & ~63u) | (unsigned)j) : 0xffffffffu;
            m1 = min(m1, key);
        }
        if (m1 != 0xffffffffu) first = (int)(m1 & 63u);
    }
    for (int jj = (first >= 0 ? -1 : 0); jj < sc.n_big_planes; jj++) {
        const int j = jj < 0 ? first : jj;
        if (jj >= 0 && j == first) continue;
        if (!plane_ahead(__ldg(sc.big_planes + j), O, d, best.t)) continue;
        for (unsigned long long m = __ldg(sc.big_masks + j); m; m &= m - 1ull) {
            const int i = sc.n_leaf + (__ffsll((long long)m) - 1);
            float t; int prim;
            if (prim_test<true>(sc.prims + i, O, d, best.t, ANY ? 0x7fffffff : best.prim, t, prim)) {
                if (ANY) return true;
                best.t = t; best.leaf = i; best.prim = prim; found = true;
            }
        }
    }
    return found;
}

// Sliver triangles whose far field begins inside the scene (bvh_build.cu: thr >= 4): exact test
// for every ray, whatever the BVH found.
template <bool ANY>
__device__ __forceinline__ bool always_scan(const DeviceScene& sc, V3 O, V3 d, HitRec& best, bool found)
{
    for (int k = 0; k < sc.n_always; k++) {
        const int i = __ldg(sc.always_idx + k);
        float t; int prim;
        if (prim_test<true>(sc.prims + i, O, d, best.t, ANY ? 0x7fffffff : best.prim, t, prim)) {
            if (ANY) return true;
            best.t = t; best.leaf = i; best.prim = prim; found = true;
        }
    }
    return found;
}

// The reference's own linear loop (cpp:476-521) over `n` records, e.g. staged in shared memory.
template <bool ANY, bool GLOBAL>
__device__ __forceinline__ bool traverse_linear(const PrimRec* __restrict__ prims, int n, V3 O, V3 d, float tmax,
                                                HitRec& best)
{
    best.t = ANY ? tmax : __int_as_float(0x7f800000);
    best.leaf = -1;
    best.prim = 0x7fffffff;
    bool found = false;
    for (int i = 0; i < n; i++) {
        float t; int prim;
        if (prim_test<GLOBAL>(prims + i, O, d, best.t, ANY ? 0x7fffffff : best.prim, t, prim)) {
            if (ANY) return true;
            best.t = t; best.leaf = i; best.prim = prim; found = true;
        }
    }
    return found;
}

}  // namespace rt580
