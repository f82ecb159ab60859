// GPU scene preparation + LBVH build (replaces the reference's per-ray linear loop over
// shapes/triangles, Raytracer.cpp:473-526, and hoists the per-ray, per-triangle constants of
// cpp:362-365 / 377 / 389 to load time).  The tree only prunes; every leaf test is the
// reference's own float arithmetic, so the closest hit is the one the linear loop returns
// (same t bits, same primitive at ties).
//
// Pipeline (all on the device):
//   k_scene_bounds   world bounds of every vertex / sphere                 (atomics)
//   k_prim_setup     PrimRec (N, D, totalArea ...), padded AABB, 63-bit Morton key
//   cub radix sort   (key, id) pairs                                        [library: CUB]
//   k_gather         records + boxes into Morton order
//   k_hierarchy      Karras 2012 radix tree over the sorted keys
//   k_refit          bottom-up AABB union with per-node arrival counters
//   k_pack           64-byte traversal nodes with both child boxes inline
//   k_depth          max leaf depth (bounds the traversal stack)
//
// Why the tree returns the reference's hit (DESIGN.md has the derivation):
// the reference accepts a hit from float arithmetic (cpp:392-396), so the set of points P the
// test accepts is not the triangle but, with u = 2^-24, L = distance of P from the triangle,
// r = its diameter, h = the smaller altitude onto the two edges meeting at v0:
//   NEAR  L <= L_near = 2048 u r^2 / h     rounding slop around the triangle (beta/gamma at
//                                          cpp:393-394 confine P to the wedge at v0 up to a
//                                          relative 10u; alpha at cpp:392 cuts the wedge off
//                                          with an absolute error <= 9u |v-P|^2 / |v1v2|)
//   FAR   L >= R_safe = h / (13u)          cpp:392's cross product of two long, almost parallel
//                                          vectors is rounding noise there: the wedge is no
//                                          longer cut off
//   and nothing in between.
// NEAR is covered by padding every box with pad = 2^-18 E + L_near (E bounds every coordinate a
// ray can start from or hit; the first term covers the off-plane error of t at cpp:381-387).
// FAR cannot be covered by any box: those are rays almost parallel to the triangle's plane that
// "hit" it 10^4..10^7 units away; the reference reports them (they decide hit/miss of rays that
// leave the scene, i.e. tree structure and AO), so the device replays them: each triangle gets
// a filter record (N, thr) with thr = 4E / (R_safe - 4E) >= |N.O + D| / T_far, and a ray that
// found nothing nearer than far_tmin tests every triangle with |N.d| <= thr exactly (trace.cuh).
// Sphere boxes use r_box^2 = r^2 + 2^-18 (2E+1)^2: the cancellation error of the
// discriminant at cpp:422-426 for the ray origins possible in the scene.
#include "device_scene.h"
#include "rt_math.cuh"
#include "build.h"
#include <cub/device/device_radix_sort.cuh>
#include <cfloat>
#include <cstdio>
#include <cstring>

namespace rt580 {

// ---- ordered-int encoding so float min/max can use integer atomics ---------------------
__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__host__ __device__ __forceinline__ float ord2f(int i) {
    int j = i >= 0 ? i : i ^ 0x7fffffff;
#if defined(__CUDA_ARCH__)
    return __int_as_float(j);
#else
    float f; memcpy(&f, &j, 4); return f;
#endif
}

__global__ void k_scene_bounds(const float4* __restrict__ v0, const float4* __restrict__ v1,
                               const float4* __restrict__ v2, int64_t n_tris,
                               const float4* __restrict__ sph, int64_t n_sph, int* __restrict__ bounds)
{
    float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_tris + n_sph; i += stride) {
        if (i < n_tris) {
            float4 a = v0[i], b = v1[i], c = v2[i];
            lo[0] = fminf(lo[0], fminf(a.x, fminf(b.x, c.x))); hi[0] = fmaxf(hi[0], fmaxf(a.x, fmaxf(b.x, c.x)));
            lo[1] = fminf(lo[1], fminf(a.y, fminf(b.y, c.y))); hi[1] = fmaxf(hi[1], fmaxf(a.y, fmaxf(b.y, c.y)));
            lo[2] = fminf(lo[2], fminf(a.z, fminf(b.z, c.z))); hi[2] = fmaxf(hi[2], fmaxf(a.z, fmaxf(b.z, c.z)));
        } else {
            float4 s = sph[i - n_tris];
            float r = fabsf(s.w);
            lo[0] = fminf(lo[0], s.x - r); hi[0] = fmaxf(hi[0], s.x + r);
            lo[1] = fminf(lo[1], s.y - r); hi[1] = fmaxf(hi[1], s.y + r);
            lo[2] = fminf(lo[2], s.z - r); hi[2] = fmaxf(hi[2], s.z + r);
        }
    }
    for (int k = 0; k < 3; k++) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
        for (int k = 0; k < 3; k++) {
            if (lo[k] <= hi[k]) { atomicMin(&bounds[k], f2ord(lo[k])); atomicMax(&bounds[3 + k], f2ord(hi[k])); }
        }
    }
}

// 21 bits -> every third bit of a 63-bit word
__device__ __forceinline__ uint64_t spread21(uint32_t v) {
    uint64_t x = v & 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

struct SetupParams {
    float blo[3];       // scene bounds (padded)
    float inv_ext[3];   // 2^21 / extent
    float pad;          // box padding (see file header)
    float sph_extra;    // added to r^2 for sphere boxes
    float E;            // bound on |coordinate| of ray origins and hit points
    float big_diam;     // primitives with a larger diameter stay out of the tree (RT_MAX_BIG of them at most)
};
#define RT_MAX_BIG 64
#define KEY_DROPPED (~0ull)
#define KEY_BIG (~0ull - 1ull)

__global__ void k_prim_setup(const float4* __restrict__ v0, const float4* __restrict__ v1,
                             const float4* __restrict__ v2, const int32_t* __restrict__ tri_prim, int64_t n_tris,
                             const float4* __restrict__ sph, const int32_t* __restrict__ sph_prim, int64_t n_sph,
                             SetupParams sp, PrimRec* __restrict__ rec, float4* __restrict__ box_lo,
                             float4* __restrict__ box_hi, uint64_t* __restrict__ keys, uint32_t* __restrict__ ids,
                             float4* __restrict__ far, unsigned int* __restrict__ n_dropped)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tris + n_sph) return;
    PrimRec r;
    float lo[3], hi[3];
    bool drop = false, big = false;
    if (i < n_tris) {
        float4 a = v0[i], b = v1[i], c = v2[i];
        V3 p0 = mk(a.x, a.y, a.z), p1 = mk(b.x, b.y, b.z), p2 = mk(c.x, c.y, c.z);
        V3 e1 = p1 - p0, e2 = p2 - p0;                        // cpp:362-363
        V3 N = normalize(cross(e1, e2));                      // cpp:364-365
        float D = -dot(N, p0);                                // cpp:377
        // cpp:389 + cpp:937-942: 0.5 * dot in double, rounded to float on return.  0.5f * x in
        // fp32 is the same correctly rounded value (exact halving, single rounding if subnormal).
        float total = 0.5f * dot(cross(p1 - p0, p2 - p0), N);
        int flags = RT_PRIM_TRIANGLE;
        float at = fabsf(total);
        if (!(at >= 1e-30f && at <= 1e30f)) flags |= RT_PRIM_SLOWPATH;
        drop = (N.x == 0.0f && N.y == 0.0f && N.z == 0.0f);  // never intersectable: |N.d| = 0 < EPSILON
        r.a = make_float4(p0.x, p0.y, p0.z, D);
        r.b = make_float4(p1.x, p1.y, p1.z, total);
        r.c = make_float4(p2.x, p2.y, p2.z, __int_as_float(tri_prim[i]));
        r.d = make_float4(N.x, N.y, N.z, __int_as_float(flags));
        // near / far split of the reference's acceptance region (file header):
        // hmin = the smaller of the altitudes onto the two edges that meet at v0
        const V3 e0 = p2 - p1;
        const float la = length(e0), lb = length(e2), lc = length(e1);          // |v1v2|, |v0v2|, |v0v1|
        const float diam = fmaxf(la, fmaxf(lb, lc));
        const float area2 = length(cross(e1, e2));
        const float hmin = 0.99f * area2 / fmaxf(lb, lc);
        float l_near = 1.2207031e-4f * diam * diam / hmin;                       // 2048 u diam^2 / hmin
        const float r_safe = hmin * 1290555.0f;                                  // hmin / (13 u)
        float thr = 4.0f;                                                        // > |N.d| always: candidate for every ray
        if (hmin > 0.0f && r_safe > 8.0f * sp.E + 12.0f * diam && l_near <= diam) {
            const float t_far = r_safe - 4.0f * sp.E - 2.0f * diam;
            thr = 4.0f * sp.E / t_far;
        } else l_near = fminf(l_near, diam);
        if (!(l_near >= 0.0f)) l_near = diam;                                    // NaN guard
        const float pad = sp.pad + l_near;
        if (!drop) atomicMax(n_dropped + 5, (unsigned)__float_as_int(l_near));          // >= 0: orders like its bits
        far[i] = make_float4(N.x, N.y, N.z, drop ? -1.0f : thr);
        big = !drop && diam > sp.big_diam;
        if (!drop) {
            if (thr < 4.0f) atomicMax(n_dropped + 2, (unsigned)__float_as_int(thr));   // positive floats order like ints
            else if (!big) atomicAdd(n_dropped + 3, 1u);
        }
        lo[0] = fminf(p0.x, fminf(p1.x, p2.x)) - pad; hi[0] = fmaxf(p0.x, fmaxf(p1.x, p2.x)) + pad;
        lo[1] = fminf(p0.y, fminf(p1.y, p2.y)) - pad; hi[1] = fmaxf(p0.y, fmaxf(p1.y, p2.y)) + pad;
        lo[2] = fminf(p0.z, fminf(p1.z, p2.z)) - pad; hi[2] = fmaxf(p0.z, fmaxf(p1.z, p2.z)) + pad;
    } else {
        far[i] = make_float4(0.f, 0.f, 0.f, -1.0f);                              // spheres have no far field
        float4 s = sph[i - n_tris];
        big = 2.0f * fabsf(s.w) > sp.big_diam;
        r.a = s;
        r.b = make_float4(s.w * s.w, 0.f, 0.f, 0.f);          // cpp:423 radius*radius
        r.c = make_float4(0.f, 0.f, 0.f, __int_as_float(sph_prim[i - n_tris]));
        r.d = make_float4(0.f, 0.f, 0.f, __int_as_float(RT_PRIM_SPHERE));
        float rb = sqrtf(s.w * s.w + sp.sph_extra) * 1.000001f + sp.pad;
        lo[0] = s.x - rb; hi[0] = s.x + rb; lo[1] = s.y - rb; hi[1] = s.y + rb; lo[2] = s.z - rb; hi[2] = s.z + rb;
    }
    rec[i] = r;
    box_lo[i] = make_float4(lo[0], lo[1], lo[2], 0.f);
    box_hi[i] = make_float4(hi[0], hi[1], hi[2], 0.f);
    uint64_t key;
    if (drop) { key = KEY_DROPPED; atomicAdd(n_dropped, 1u); }
    else if (big) { key = KEY_BIG; atomicAdd(n_dropped + 4, 1u); }
    else {
        uint32_t q[3];
        for (int k = 0; k < 3; k++) {
            float c = 0.5f * (lo[k] + hi[k]);
            float f = (c - sp.blo[k]) * sp.inv_ext[k];
            f = fminf(fmaxf(f, 0.0f), 2097151.0f);
            q[k] = (uint32_t)f;
        }
        key = (spread21(q[0]) << 2) | (spread21(q[1]) << 1) | spread21(q[2]);
    }
    keys[i] = key;
    ids[i] = (uint32_t)i;
}

__global__ void k_gather(const PrimRec* __restrict__ rec, const float4* __restrict__ box_lo,
                         const float4* __restrict__ box_hi, const float4* __restrict__ far_in,
                         const uint32_t* __restrict__ ids, int n_leaf, int n_all,
                         PrimRec* __restrict__ out, float4* __restrict__ far_out, float4* __restrict__ nb_lo,
                         float4* __restrict__ nb_hi)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_all) return;
    uint32_t i = ids[j];
    out[j] = rec[i];
    far_out[j] = far_in[i];
    if (j >= n_leaf) return;      // large primitives: record only, no place in the tree
    // node boxes: inner nodes [0, n_leaf-1), leaves [n_leaf-1, 2 n_leaf-1)
    nb_lo[n_leaf - 1 + j] = box_lo[i];
    nb_hi[n_leaf - 1 + j] = box_hi[i];
}

__global__ void k_leaf_of_prim(const PrimRec* __restrict__ prims, int n, int32_t* __restrict__ out) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) out[__float_as_int(prims[j].c.w)] = j;
}

__global__ void k_collect_always(const float4* __restrict__ far, int n, int32_t* __restrict__ out,
                                 unsigned int* __restrict__ count)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n && far[j].w >= 4.0f) out[atomicAdd(count, 1u)] = j;
}

// Karras 2012: length of the common prefix of keys i and j, ties broken by index
__device__ __forceinline__ int delta(const uint64_t* __restrict__ k, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint64_t a = k[i], b = k[j];
    if (a == b) return 64 + __clz(i ^ j);
    return __clzll((long long)(a ^ b));
}

__global__ void k_hierarchy(const uint64_t* __restrict__ keys, int n, int2* __restrict__ kids,
                            int* __restrict__ parent)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta(keys, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + min(d, 0);
    int left = (min(i, j) == gamma) ? ~gamma : gamma;          // leaf encoded as ~index
    int right = (max(i, j) == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    kids[i] = make_int2(left, right);
    // parent array: inner nodes [0,n-1), leaves [n-1, 2n-1)
    parent[left >= 0 ? left : (n - 1 + ~left)] = i;
    parent[right >= 0 ? right : (n - 1 + ~right)] = i;
    if (i == 0) parent[0] = -1;
}

__global__ void k_refit(int n, const int2* __restrict__ kids, const int* __restrict__ parent,
                        float4* __restrict__ nb_lo, float4* __restrict__ nb_hi, unsigned int* __restrict__ arrive)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    int node = parent[n - 1 + j];
    while (node >= 0) {
        __threadfence();
        if (atomicAdd(&arrive[node], 1u) == 0u) return;        // first child to arrive stops
        __threadfence();
        int2 k = kids[node];
        int a = k.x >= 0 ? k.x : (n - 1 + ~k.x), b = k.y >= 0 ? k.y : (n - 1 + ~k.y);
        float4 la = __ldcg(&nb_lo[a]), lb = __ldcg(&nb_lo[b]), ha = __ldcg(&nb_hi[a]), hb = __ldcg(&nb_hi[b]);
        nb_lo[node] = make_float4(fminf(la.x, lb.x), fminf(la.y, lb.y), fminf(la.z, lb.z), 0.f);
        nb_hi[node] = make_float4(fmaxf(ha.x, hb.x), fmaxf(ha.y, hb.y), fmaxf(ha.z, hb.z), 0.f);
        node = parent[node];
    }
}

__global__ void k_pack(int n, const int2* __restrict__ kids, const float4* __restrict__ nb_lo,
                       const float4* __restrict__ nb_hi, BvhNode* __restrict__ nodes)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int2 k = kids[i];
    int a = k.x >= 0 ? k.x : (n - 1 + ~k.x), b = k.y >= 0 ? k.y : (n - 1 + ~k.y);
    float4 la = nb_lo[a], ha = nb_hi[a], lb = nb_lo[b], hb = nb_hi[b];
    BvhNode nd;
    nd.xy0 = make_float4(la.x, ha.x, la.y, ha.y);
    nd.xy1 = make_float4(lb.x, hb.x, lb.y, hb.y);
    nd.z01 = make_float4(la.z, ha.z, lb.z, hb.z);
    nd.kids = make_int4(k.x, k.y, 0, 0);
    nodes[i] = nd;
}

__global__ void k_depth(int n, const int* __restrict__ parent, unsigned int* __restrict__ max_depth) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    unsigned int d = 0;
    int node = parent[n - 1 + j];
    while (node >= 0) { d++; node = parent[node]; }
    for (int o = 16; o > 0; o >>= 1) d = max(d, __shfl_xor_sync(0xffffffffu, d, o));
    if ((threadIdx.x & 31) == 0) atomicMax(max_depth, d);
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { snprintf(err, errlen, "%s:%d %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); return false; } } while (0)

#define TAKE(ptr, arena, T, n) do { (ptr) = (arena).take<T>(n); if (!(ptr)) { snprintf(err, errlen, "%s:%d arena too small for %s", __FILE__, __LINE__, #ptr); return false; } } while (0)

bool arena_reserve(DevArena& a, size_t bytes, char* err, size_t errlen)
{
    a.off = 0;
    if (bytes <= a.cap) return true;
    if (a.base) cudaFree(a.base);
    a.base = nullptr; a.cap = 0;
    const size_t want = bytes + bytes / 8 + (1u << 20);
    CK(cudaMalloc((void**)&a.base, want));
    a.cap = want;
    return true;
}
void arena_release(DevArena& a) { if (a.base) cudaFree(a.base); a.base = nullptr; a.cap = 0; a.off = 0; }

static size_t sort_tmp_bytes(int64_t n_in)
{
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int)n_in, 0, 64, (cudaStream_t)0);
    return bytes;
}
size_t build_tmp_bytes(int64_t n_in)
{
    const size_t n = (size_t)(n_in > 0 ? n_in : 1);
    // rec 64, boxes 32, far 16, keys 16, ids 8 | node boxes 64, kids 8, parent 8, arrive 4 | 32 buffers x 256 B rounding
    return n * (64 + 32 + 16 + 16 + 8 + 64 + 8 + 8 + 4) + sort_tmp_bytes(n_in) + 64 * 256 + 4096;
}
size_t build_out_bytes(int64_t n_in, int64_t n_prims)
{
    const size_t n = (size_t)(n_in > 0 ? n_in : 1), np = (size_t)(n_prims > 0 ? n_prims : 1);
    return n * (64 + 64 + 16 + 4) + np * 4 + 16 * 256 + 4096 + 64;   // prims (+ 1), nodes, far, always_idx | leaf_of_prim
}

bool build_bvh(const BuildInput& in, BuildOutput* out, DevArena& tmpa, DevArena& outa, cudaStream_t stream, char* err,
               size_t errlen)
{
    const int64_t n_in = in.n_tris + in.n_spheres;
    out->prims = nullptr; out->nodes = nullptr; out->far = nullptr; out->far_tmin = 0.f; out->n_always = 0; out->always_idx = nullptr; out->leaf_of_prim = nullptr;
    out->n_leaf = 0; out->n_big = 0; out->max_depth = 0; out->launches = 0; out->pad_max = 0.f; out->nan_leaf = -1;
    if (n_in > 0x7ffffff0ll) { snprintf(err, errlen, "too many primitives (%lld)", (long long)n_in); return false; }

    int* d_bounds = nullptr; unsigned int* d_counters = nullptr;
    TAKE(d_bounds, tmpa, int, 6); TAKE(d_counters, tmpa, unsigned int, 8);
    CK(cudaMemsetAsync(d_counters, 0, 8 * sizeof(unsigned int), stream));   // dropped, depth, max thr bits, always-candidates, big
    {
        const int init[6] = { 0x7f7fffff, 0x7f7fffff, 0x7f7fffff,                     // +FLT_MAX, ordered encoding
                              (int)0x80800000, (int)0x80800000, (int)0x80800000 };    // -FLT_MAX, ordered encoding
        CK(cudaMemcpyAsync(d_bounds, init, sizeof init, cudaMemcpyHostToDevice, stream));
    }
    if (n_in > 0) {
        int blocks = (int)((n_in + 255) / 256); if (blocks > 148 * 16) blocks = 148 * 16;
        k_scene_bounds<<<blocks, 256, 0, stream>>>(in.tri_v0, in.tri_v1, in.tri_v2, in.n_tris, in.sph, in.n_spheres, d_bounds);
        out->launches++;
    }
    int hb[6];
    CK(cudaMemcpyAsync(hb, d_bounds, sizeof hb, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    float lo[3], hi[3], E = 1.0f;
    for (int k = 0; k < 3; k++) {
        lo[k] = ord2f(hb[k]); hi[k] = ord2f(hb[3 + k]);
        if (!(lo[k] <= hi[k])) { lo[k] = 0.f; hi[k] = 0.f; }
        E = fmaxf(E, fmaxf(fabsf(lo[k]), fabsf(hi[k])));
        E = fmaxf(E, fabsf(in.origin_hint[k]));
    }
    E += 1.0f;   // + SHADOW_CLIPPING_OFFSET steps and slack
    SetupParams sp;
    sp.pad = E * (1.0f / 262144.0f);                                   // 2^-18 E
    sp.E = E;
    sp.sph_extra = (2.f * E + 1.f) * (2.f * E + 1.f) * (1.0f / 262144.0f);   // 2^-18 (2E+1)^2
    {
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        sp.big_diam = 0.125f * sqrtf(dx * dx + dy * dy + dz * dz);
        if (!(sp.big_diam > 0.f) || n_in <= RT_MAX_BIG) sp.big_diam = 3.0e38f;   // tiny scenes: nothing to separate
    }
    for (int k = 0; k < 3; k++) {
        float l = lo[k] - 2.f * sp.pad, h = hi[k] + 2.f * sp.pad;
        sp.blo[k] = l;
        sp.inv_ext[k] = 2097152.0f / fmaxf(h - l, 1e-20f);
    }
    out->pad = sp.pad; out->extent = E; out->pad_max = sp.pad;
    for (int k = 0; k < 3; k++) { out->bounds_lo[k] = lo[k]; out->bounds_hi[k] = hi[k]; }
    if (n_in == 0) return true;

    PrimRec* rec = nullptr; float4 *blo = nullptr, *bhi = nullptr, *far_in = nullptr;
    uint64_t *keys = nullptr, *keys2 = nullptr; uint32_t *ids = nullptr, *ids2 = nullptr;
    TAKE(rec, tmpa, PrimRec, n_in); TAKE(blo, tmpa, float4, n_in); TAKE(bhi, tmpa, float4, n_in); TAKE(far_in, tmpa, float4, n_in);
    TAKE(keys, tmpa, uint64_t, n_in); TAKE(keys2, tmpa, uint64_t, n_in); TAKE(ids, tmpa, uint32_t, n_in); TAKE(ids2, tmpa, uint32_t, n_in);
    {
        int blocks = (int)((n_in + 255) / 256);
        k_prim_setup<<<blocks, 256, 0, stream>>>(in.tri_v0, in.tri_v1, in.tri_v2, in.tri_prim, in.n_tris, in.sph,
                                                 in.sph_prim, in.n_spheres, sp, rec, blo, bhi, keys, ids, far_in, d_counters);
        out->launches++;
    }
    size_t tmp_bytes = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, ids, ids2, (int)n_in, 0, 64, stream));
    char* tmp = nullptr;
    TAKE(tmp, tmpa, char, tmp_bytes);
    CK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, ids, ids2, (int)n_in, 0, 64, stream));
    out->launches += 4;   // cub onesweep: histogram + scan + passes (approximate, library)
    unsigned int hc[8];
    CK(cudaMemcpyAsync(hc, d_counters, sizeof hc, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    if (hc[4] > RT_MAX_BIG) {
        // too many "large" primitives for a list that every ray walks: keep them all in the tree
        sp.big_diam = 3.0e38f;
        CK(cudaMemsetAsync(d_counters, 0, 8 * sizeof(unsigned int), stream));
        k_prim_setup<<<(int)((n_in + 255) / 256), 256, 0, stream>>>(in.tri_v0, in.tri_v1, in.tri_v2, in.tri_prim, in.n_tris, in.sph,
                                                                    in.sph_prim, in.n_spheres, sp, rec, blo, bhi, keys, ids, far_in, d_counters);
        CK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, ids, ids2, (int)n_in, 0, 64, stream));
        out->launches += 5;
        CK(cudaMemcpyAsync(hc, d_counters, sizeof hc, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
    }
    const int n_big = (int)hc[4];
    const int n = (int)(n_in - hc[0]) - n_big;
    const int n_all = n + n_big;
    out->n_leaf = n; out->n_big = n_big;
    out->n_dropped = (int)hc[0];
    out->n_always = (int)hc[3];
    { float ln; memcpy(&ln, &hc[5], 4); out->pad_max = sp.pad + ln; }
    {
        float max_thr; memcpy(&max_thr, &hc[2], 4);
        // far-field acceptance needs t >= T_far = 4E / thr; with sliver triangles (candidates for
        // every ray) there is no lower bound
        out->far_tmin = max_thr > 0.f ? 4.0f * E / max_thr : 3.0e38f;
    }
    if (n_all > 0) {
        PrimRec* prims = nullptr; BvhNode* nodes = nullptr; float4* far = nullptr;
        float4 *nlo = nullptr, *nhi = nullptr; int2* kids = nullptr; int* parent = nullptr; unsigned int* arrive = nullptr;
        TAKE(prims, outa, PrimRec, (size_t)n_all + 1); TAKE(nodes, outa, BvhNode, n > 1 ? n - 1 : 1); TAKE(far, outa, float4, n_all);
        TAKE(nlo, tmpa, float4, 2 * (size_t)n + 2); TAKE(nhi, tmpa, float4, 2 * (size_t)n + 2);
        TAKE(kids, tmpa, int2, n + 1); TAKE(parent, tmpa, int, 2 * (size_t)n + 2); TAKE(arrive, tmpa, unsigned int, n + 1);
        int blocks = (n_all + 255) / 256;
        k_gather<<<blocks, 256, 0, stream>>>(rec, blo, bhi, far_in, ids2, n, n_all, prims, far, nlo, nhi); out->launches++;
        if (in.first_tri >= 0 && in.first_tri < in.n_tris) {
            CK(cudaMemcpyAsync(prims + n_all, rec + in.first_tri, sizeof(PrimRec), cudaMemcpyDeviceToDevice, stream));
            out->nan_leaf = n_all;
        }
        if (n > 1) {
            CK(cudaMemsetAsync(arrive, 0, sizeof(unsigned int) * n, stream));
            k_hierarchy<<<blocks, 256, 0, stream>>>(keys2, n, kids, parent); out->launches++;
            k_refit<<<blocks, 256, 0, stream>>>(n, kids, parent, nlo, nhi, arrive); out->launches++;
            k_pack<<<blocks, 256, 0, stream>>>(n, kids, nlo, nhi, nodes); out->launches++;
            k_depth<<<blocks, 256, 0, stream>>>(n, parent, d_counters + 1); out->launches++;
            CK(cudaMemcpyAsync(hc, d_counters, sizeof hc, cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            out->max_depth = hc[1];
        } else if (n == 1) {
            // single primitive: one node, child1 = empty box that no ray can enter
            float4 l, h;
            CK(cudaMemcpyAsync(&l, nlo + (n - 1), sizeof l, cudaMemcpyDeviceToHost, stream));
            CK(cudaMemcpyAsync(&h, nhi + (n - 1), sizeof h, cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            BvhNode nd;
            nd.xy0 = make_float4(l.x, h.x, l.y, h.y);
            nd.xy1 = make_float4(FLT_MAX, -FLT_MAX, FLT_MAX, -FLT_MAX);
            nd.z01 = make_float4(l.z, h.z, FLT_MAX, -FLT_MAX);
            nd.kids = make_int4(~0, ~0, 0, 0);
            CK(cudaMemcpyAsync(nodes, &nd, sizeof nd, cudaMemcpyHostToDevice, stream));
            CK(cudaStreamSynchronize(stream));
            out->max_depth = 1;
        }
        CK(cudaGetLastError());
        out->prims = prims; out->nodes = nodes; out->far = far;
        {
            int32_t* lop = nullptr;
            TAKE(lop, outa, int32_t, (size_t)in.n_prims);
            CK(cudaMemsetAsync(lop, 0xff, sizeof(int32_t) * (size_t)(in.n_prims ? in.n_prims : 1), stream));
            k_leaf_of_prim<<<blocks, 256, 0, stream>>>(prims, n_all, lop); out->launches++;
            out->leaf_of_prim = lop;
        }
        if (out->n_always > 0) {
            int32_t* idx = nullptr;
            TAKE(idx, outa, int32_t, out->n_always);
            CK(cudaMemsetAsync(d_counters, 0, sizeof(unsigned int), stream));
            k_collect_always<<<blocks, 256, 0, stream>>>(far, n, idx, d_counters); out->launches++;
            CK(cudaStreamSynchronize(stream));
            out->always_idx = idx;
        }
    }
    CK(cudaStreamSynchronize(stream));
    CK(cudaGetLastError());
    return true;
}

}  // namespace rt580
