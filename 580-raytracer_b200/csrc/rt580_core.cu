// librt580 core: context, scene upload, and the wavefront frame pipeline that replaces the
// loop body of Raytracer::Render (Raytracer.cpp:921-932) and the recursion of
// Raytracer::Raycast (cpp:28-129).
//
// The recursive reflect/refract Raycast becomes a breadth-first wavefront over "hit nodes":
//
//   level 0   k_trace<primary>     GenerateRay (cpp:832-858) + closest hit per pixel
//   level L   main chain           k_closest (persistent closest-hit traversal of the level's reflection / refraction
//                                  rays) -> k_commit (large primitives, far field, node creation): level L+1 nodes.
//                                  The kernel that creates a node spawns its children (cpp:94-112) into the next queue.
//             side streams         k_shade_gen (shadow rays cpp:53-81: large primitives, clearance map) -> k_anyhit
//                                  (persistent any-hit traversal) -> k_shade_local (CalculateLocalColor of the lit lights)
//   order     k_subtree (bottom-up), scan over pixels, k_preorder (top-down): gives every hit
//             node its ordinal in the reference's traversal order (scanline pixels x pre-order
//             nodes), i.e. its position in the single AO random stream
//   AO        k_ao_gen             one thread per (node, ambient light, sample): stream position by modular
//                                  exponentiation (Appendix C), hemisphere direction, large primitives first;
//             k_anyhit             the rays those did not stop
//   resolve   k_resolve (bottom-up per level): the integer Pixel algebra of cpp:39-51 and
//             cpp:114-128, child colour written into the parent's slot, roots into the frame
//
// Tiny scenes (<= 64 primitives) and the brute-force checker run the reference's linear loop instead, one thread
// per ray / node: k_trace<MODE>, k_shade<MODE>, k_ao<MODE>.
//
// AO never influences ray geometry (cpp:45 only scales a colour), so the tree of every pixel
// is known before a single AO ray is traced; that is what makes the stream addressable.
#include "../../include/rt580.h"
#include "device_scene.h"
#include "build.h"
#include "trace.cuh"
#include "shade.cuh"
#include "smap.cuh"
#include "fargrid.cuh"
#include <cuda_runtime.h>
#include <cstdio>
#include <ctime>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <string>
#include <mutex>

using namespace rt580;

// ---------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
extern "C" const char* rt580_last_error(void) { return g_err; }
void rt580_set_error(const char* msg) { snprintf(g_err, sizeof g_err, "%s", msg); }
#define FAIL(code, ...) do { snprintf(g_err, sizeof g_err, __VA_ARGS__); return (code); } while (0)
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { snprintf(g_err, sizeof g_err, "%s:%d %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); return RT580_FAILURE; } } while (0)

// ---------------------------------------------------------------------------------------
// frame data
// ---------------------------------------------------------------------------------------
struct __align__(16) Node {       // 64 B per hit Raycast node
    float4 P;   // hitPoint.xyz                      w: bits(primitive order index)
    float4 N;   // RaycastHitInfo::normal.xyz        w: bits(parent node, -1 for a root)
    float4 D;   // direction of the ray that hit     w: bits(local pixel index)
    float4 B;   // alpha, beta, gamma (triangles)    w: bits(flags)
};
#define NF_SPHERE   1u            // flags bit 0: hit a sphere
#define NF_REFR     2u            // flags bit 1: this node is its parent's refraction child
#define NF_BOUNCE_SHIFT 8         // flags bits 8..15: bounces left at this node

struct __align__(16) NodeAux {    // 32 B
    short local[4];               // sum over non-ambient lights of lit ? CalculateLocalColor : 0
    short refl[4];                // reflectionColor (cpp:91,103): rgb; [3] = 0 none, 1 miss(BG), 2 hit child
    short refr[4];                // refractionColor (cpp:92,111)
    uint32_t sub_refl;            // hit nodes in the reflection subtree
    uint32_t sub_refr;            // hit nodes in the refraction subtree
};

struct __align__(16) QRay {       // 32 B queue entry
    float4 o;   // origin.xyz    w: bits(parent node)
    float4 d;   // direction.xyz w: bits(flags for the child: NF_REFR, bounces left)
};

// Deferred slow rays (trace.cuh: far scan / linear fallback): recorded by the kernel that meets
// them, answered by k_slow with one warp per ray over the whole GPU, applied by a finish kernel.
struct __align__(16) SlowRay {
    float4 o;   // origin.xyz     w: bits(best t so far / tmax)
    float4 d;   // direction.xyz  w: bits(best prim so far)
    int4 c;     // x: bit0 = linear fallback, y/z: consumer ids, w: best leaf so far (-1 none)
};
// answer of a slow ray: closest hit = lexicographic minimum of (t, prim) packed so that an integer
// atomicMin over record slices gives it (t > 0, so the float bits order like the value);
// any hit = the flag
struct __align__(16) SlowRes { unsigned long long key; int found; int pad; };
struct SlowQ { SlowRay* rays; SlowRes* res; unsigned int* count; unsigned cap; };
__host__ __device__ __forceinline__ unsigned long long slow_key(float t, int prim) {
#if defined(__CUDA_ARCH__)
    return ((unsigned long long)__float_as_uint(t) << 32) | (unsigned)prim;
#else
    unsigned u; memcpy(&u, &t, 4); return ((unsigned long long)u << 32) | (unsigned)prim;
#endif
}
#define SLOW_CAP_MAX (16u << 20)   // deferred closest-hit rays / any-hit rays of a leaky scene: one entry per ray of a chunk
#define SLOW_ANY_CAP (4u << 20)    // deferred any-hit rays of a scene that hardly leaks (on overflow the pass is repeated "leaky")
#define RT580_INTERNAL_OVERFLOW 100
#define AH_CHUNK_TIGHT (32u << 20) // any-hit rays generated per chunk
#define LEAKY_ANY_CAP (384u << 20) // deferred any-hit rays of one occlusion pass of a leaky scene
#define OCCL_PENDING 0x40000000u   // shadow ray deferred to the end of the structure pass (k_shadow_finish decides)

// k_anyhit (persistent any-hit traversal) constants
#ifndef AH_MIN_BLOCKS
#define AH_MIN_BLOCKS 12
#endif
#define AH_STEPS 48        // at most this many inner-node steps between two leaf / refill phases
#define AH_MIN_SEARCH 16   // leave the inner-node phase when fewer lanes than this still have an inner node
#define AH_NONE 0x7fffffff // traversal cursor: nothing left
#define AH_BATCH 512  // most rays a warp reserves per atomic on the queue counter (fewer when the queue is short)

struct FrameParams {
    int W, H;
    int row_first, row_step, n_rows;
    int depth, spp, rng_mode;
    int spp_shift;        // log2(spp) when spp is a power of two, else -1
    float cam[3];
    float inv[9];
    const float* ndc_x;   // [W]  (float)(NDCX * aspect * tan(fov/2))   cpp:834-839, 846
    const float* ndc_y;   // [H]  (float)(NDCY * tan(fov/2))            cpp:835, 840, 846
    const uint32_t* lcg_pow;   // [spp] 16807^(2k) mod (2^31-1): advances an AO call's engine state to its sample k
    const uint32_t* lcg_tab;   // [4][256] 16807^(d * 256^k) mod (2^31-1): engine state at any step in 3 modular products
};

struct rt580_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    // the shadow rays of a level run beside the spawn -> closest hit -> commit chain: even levels on side[0], odd
    // levels on side[1] (each with its own ray chunk, occluder counts and queue counters: lanes 0 and 1)
    cudaStream_t side[2] = { nullptr, nullptr };
    cudaEvent_t ev_join[2] = { nullptr, nullptr };
    cudaEvent_t ev_level = nullptr;    // recorded on the context's stream when a level's nodes exist: the side stream starts after it
    DBuf<struct ARay> arays2; DBuf<uint32_t> occl2;
    bool overlap = true;               // RT580_NO_OVERLAP=1: everything on one stream (A/B)
    cudaDeviceProp prop;
    // scene
    DeviceScene sc{};
    PrimRec* d_prims = nullptr; BvhNode* d_nodes = nullptr; float4* d_far = nullptr;
    int n_always = 0, n_dropped = 0; int32_t* d_always = nullptr; int32_t* d_leaf_of_prim = nullptr;
    float4* d_vn = nullptr; int32_t* d_prim_material = nullptr; float* d_materials = nullptr;
    int32_t* d_light_type = nullptr; float* d_light_f = nullptr;
    DevArena scene_arena, build_arena;   // scene buffers / upload + build temporaries (grow-only)
    // far-field direction grid (fargrid.cuh): lists of the scene, and the per-flush sort of the deferred rays by direction cell
    DBuf<unsigned int> fg_counts; DBuf<unsigned long long> fg_start, fg_bsum; DBuf<uint32_t> fg_entries; DBuf<unsigned int> fg_cell_tmin;
    FgBuildInput fg_in{}; bool fg_pending = false;   // the lists are built when a frame first needs them (far_grid_ensure)
    int fg_dense_rays = 2;                    // RT580_FG_DENSE: rays per cell from which k_fg_scan takes 32 rays per warp
    unsigned arc_max_cells = 4096u;           // (ARC_MAX_CELLS) RT580_ARC_MAX_CELLS (tests: a small value forces the k_far_linear overflow path)
    int fg_K_env = -1;                   // RT580_FAR_GRID: -1 default (by triangle count), 0 off, else cells per cube-face edge
    unsigned long long fg_n_entries = 0; float fg_build_ms = 0.f;
    DBuf<unsigned int> fgq_hist, fgq_start, fgq_cellof, fgq_rank, fgq_order, fgq_lin, fgq_first;
    DBuf<struct ArcItem> arc_items; DBuf<unsigned char> arc_pre;
    bool have_scene = false;
    float build_ms = 0.f; unsigned bvh_depth = 0; float pad_extent = 0.f;
    // frame
    FrameParams fp{};
    int traversal = 0;
    DBuf<float> ndc;
    DBuf<Node> nodes; DBuf<NodeAux> aux; DBuf<QRay> queue, queue2;   // the ray queues of two consecutive levels (wavefront path)
    DBuf<uint64_t> pre;            // per node: ordinal of its first AO call / n_ambient
    DBuf<uint32_t> ao_state;       // per AO call: engine state at its first draw
    DBuf<uint32_t> ao_hits;        // per AO call: occluded samples
    DBuf<uint32_t> pix_hits;       // per local pixel: hit nodes
    DBuf<uint32_t> pix_scan;       // exclusive scan of pix_hits
    DBuf<uint32_t> scan_tmp;
    DBuf<uint64_t> row_vals;       // per local row: hit nodes / base
    DBuf<int16_t> fb;              // [n_rows][W][3]
    DBuf<unsigned int> counters;   // see read_counters
    DBuf<SlowRay> slow_rays; DBuf<SlowRes> slow_res;   // deferred closest-hit rays (answered level by level)
    DBuf<SlowRay> any_rays; DBuf<SlowRes> any_res;     // deferred any-hit rays (flushed once per pass when the scene hardly leaks)
    unsigned any_cap = 0;
    unsigned long long slow_seen = 0;                  // slow rays of the frame so far (host copy of counters[4] + [5])
    unsigned long long rays_structure = 0;
    unsigned syncs = 0;                                // host round trips of the frame
    bool force_leaky = false;                          // the small any-hit queue overflowed once for this scene
    int ndc_w = 0, ndc_h = 0; float ndc_fov = 0.f;     // what the primary-ray tables were built for
    DBuf<uint32_t> lcg_pow; int lcg_pow_spp = 0;
    DBuf<uint32_t> lcg_tab;
    DBuf<uint8_t> rgb8; unsigned long long fb_pixels = 0;   // gamma-encoded copy of the last frame (rt580_frame_rgb8)
    // the whole W x H frame of a multi-GPU render: rank 0's own allocation, or that allocation mapped
    // into this process over NVLink (cudaIpc); every rank stores its rows there after the resolve pass
    int16_t* frame = nullptr; bool frame_imported = false; int frame_w = 0, frame_h = 0;
    DBuf<struct ARay> arays;       // one chunk of generated any-hit rays (AO samples / shadow rays)
    DBuf<uint32_t> occl;           // per shadow ray of the current level: occluders found
    DBuf<struct CHit> chits;       // per queued secondary ray: what the tree answered (k_closest -> k_commit)
    int ch_blocks_per_sm = 12; bool one_thread_per_ray = false;   // RT580_CH_BLOCKS_PER_SM, RT580_ONE_THREAD_PER_RAY (A/B)
    uint64_t slow_total = 0;
    int ah_batch_div = 4;
    int smap_res = SMAP_RES_DEFAULT;        // RT580_SMAP_RES
    unsigned slow_any_cap = SLOW_ANY_CAP;   // RT580_SLOW_ANY_CAP: shrink it to exercise the overflow -> repeat path
    int ah_steps = AH_STEPS, ah_min_search = AH_MIN_SEARCH, ah_blocks_per_sm = 12;  // k_anyhit tuning (env RT580_AH_*)
    // the same for a scene that leaks (open: half of the any-hit rays walk the whole tree and leave): measured best on c4_open,
    // 37.3 against 39.9 ms for the AO rays; the closed room prefers the values above (11.8 against 12.8 ms for its shadow rays)
    int ah_steps_leaky = 24, ah_min_search_leaky = 6;
    std::vector<size_t> level_off; // node index where each level starts (+ end)
    std::vector<uint64_t> level_rays;
    bool frame_begun = false;
    cudaEvent_t ev[12];
    rt580_stats stats{};
    uint32_t launches = 0;
    std::vector<uint64_t> last_ao_base;   // host copy for the checker
    // per-class device timers of the frame (rt580_frame_profile): pairs of CUDA events on the launching stream
    std::vector<cudaEvent_t> tm_ev; std::vector<int> tm_cls; size_t tm_n = 0;
    rt580_profile prof{};
    DBuf<unsigned long long> visit_counts;   // [4] nodes / leaves of k_anyhit, nodes / leaves of k_closest
    bool count_visits = false;            // rt580_set_profiling: the traversal kernels count node visits and leaf tests
};
static int tm_begin(rt580_context* c, int cls, cudaStream_t st) {
    if (c->tm_n + 2 > c->tm_ev.size()) {
        for (int k = 0; k < 2; k++) { cudaEvent_t e; CU(cudaEventCreate(&e)); c->tm_ev.push_back(e); }
        c->tm_cls.resize(c->tm_ev.size() / 2);
    }
    c->tm_cls[c->tm_n / 2] = cls;
    CU(cudaEventRecord(c->tm_ev[c->tm_n], st));
    return RT580_SUCCESS;
}
static int tm_end(rt580_context* c, cudaStream_t st) {
    CU(cudaEventRecord(c->tm_ev[c->tm_n + 1], st));
    c->tm_n += 2;
    return RT580_SUCCESS;
}
static void tm_resolve(rt580_context* c) {
    for (size_t k = 0; k + 1 < c->tm_n; k += 2) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c->tm_ev[k], c->tm_ev[k + 1]) == cudaSuccess) { c->prof.ms[c->tm_cls[k / 2]] += ms; c->prof.launches[c->tm_cls[k / 2]]++; }
    }
    c->tm_n = 0;
}

static void frame_release(rt580_context* c);

// ---------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------
// warp-aggregated slot allocation; must be reached by all 32 lanes of the warp
__device__ __forceinline__ unsigned warp_alloc(unsigned int* counter, bool want) {
    const unsigned mask = __ballot_sync(0xffffffffu, want);
    if (mask == 0u) return 0u;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    unsigned base = 0u;
    if (lane == leader) base = atomicAdd(counter, (unsigned)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + (unsigned)__popc(mask & ((1u << lane) - 1u));
}

#define RT_SMEM_PRIMS 64   // scenes up to this many primitives are staged in shared memory

// One ray per lane, called by ALL 32 lanes of a warp (inactive lanes pass active = false).
// The tree answers almost every ray.  The rare ones it cannot (trace.cuh) are pushed to the
// deferred queue `q` when there is one (returns TR_PENDING: the caller's finish kernel applies
// the answer), otherwise - or when the queue is full - the warp serves them in place.
// a NaN anywhere in the ray (device_scene.h nan_leaf)
__device__ __forceinline__ bool ray_has_nan(V3 O, V3 d) { return !(O.x == O.x && O.y == O.y && O.z == O.z && d.x == d.x && d.y == d.y && d.z == d.z); }
#define TR_MISS 0
#define TR_HIT 1
#define TR_PENDING 2
template <int MODE /*0 bvh, 1 linear from smem, 2 linear from global*/, bool ANY>
__device__ __forceinline__ int trace_ray(const DeviceScene& sc, const PrimRec* smem_prims, bool active, V3 O, V3 d,
                                         float tmax, HitRec& hit, SlowQ q, int ca, int cb, unsigned* cnt = nullptr) {
    const bool nan_ray = active && ray_has_nan(O, d);
    if (nan_ray) {
        hit.t = __int_as_float(0x7fc00000); hit.leaf = sc.nan_leaf; hit.prim = sc.nan_prim;
        active = false;                        // (the search below runs for the other lanes of the warp)
    }
    if (MODE == 1) return nan_ray ? (sc.nan_leaf >= 0 ? TR_HIT : TR_MISS) : ((active && traverse_linear<ANY, false>(smem_prims, sc.n_all, O, d, tmax, hit)) ? TR_HIT : TR_MISS);
    if (MODE == 2) return nan_ray ? (sc.nan_leaf >= 0 ? TR_HIT : TR_MISS) : ((active && traverse_linear<ANY, true>(sc.prims, sc.n_all, O, d, tmax, hit)) ? TR_HIT : TR_MISS);
    bool found = false, need = false, linear = false, pending = false;
    if (!nan_ray) { hit.t = ANY ? tmax : __int_as_float(0x7f800000); hit.leaf = -1; hit.prim = 0x7fffffff; }
    if (active) {
        if (sc.farfield && !in_scene(sc, O)) {
            // starts outside the extent the boxes were padded for: only the child of a far-field
            // "hit" can (one float ulp out there is larger than a triangle) -> reference's linear loop
            need = true; linear = true;
            if (cnt) cnt[3]++;
            if (sc.diag) atomicAdd(sc.diag + 1, 1u);
        } else {
            // large-primitive list: first for closest hits (it bounds t for the tree) and for unbounded
            // any-hit rays (the likeliest occluders); last for bounded shadow rays, whose occluder, if
            // any, is usually a small object between the surface and the light
            const bool big_first = !ANY || !(tmax < 3.0e38f);
            if (big_first) found = big_scan<ANY>(sc, O, d, hit);
            if (!(ANY && found)) found = traverse_bvh<ANY>(sc, O, d, tmax, hit, cnt) || found;
            if (!big_first && !found) found = big_scan<ANY>(sc, O, d, hit);
            // a zero direction (total internal reflection, cpp:197-199, Q20) fails |N.d| >= EPSILON for
            // every triangle (cpp:371): only spheres can "hit" it, and those are all in the tree
            const bool zero_dir = (d.x == 0.0f && d.y == 0.0f && d.z == 0.0f);
            if (!(ANY && found) && sc.farfield && !zero_dir) {
                if (sc.n_always) found = always_scan<ANY>(sc, O, d, hit, found);
                if (ANY) need = !found && tmax >= sc.far_tmin;
                else need = !found || hit.t >= sc.far_tmin;
                if (need) { if (cnt) cnt[2]++; if (sc.diag) atomicAdd(sc.diag, 1u); }
            }
        }
        if (need && q.rays) {
            const unsigned slot = atomicAdd(q.count, 1u);
            if (slot < q.cap) {
                SlowRay r;
                r.o = make_float4(O.x, O.y, O.z, hit.t);
                r.d = make_float4(d.x, d.y, d.z, __int_as_float(hit.prim));
                r.c = make_int4(linear ? 1 : 0, ca, cb, found ? hit.leaf : -1);
                q.rays[slot] = r;
                SlowRes a; a.key = slow_key(hit.t, hit.prim); a.found = 0; a.pad = 0;
                q.res[slot] = a;
                pending = true; need = false;
            }
        }
    }
    if (sc.farfield) found = warp_slow_path<ANY>(sc, need, linear, O, d, hit, found);
    if (nan_ray) return sc.nan_leaf >= 0 ? TR_HIT : TR_MISS;
    return pending ? TR_PENDING : (found ? TR_HIT : TR_MISS);
}

// Closest hit, second half, for rays whose tree traversal ran elsewhere (k_closest): `hit` comes in
// with the tree's answer (leaf -1: nothing, leaf -2: the ray starts outside the padded extent and was
// not traversed) and leaves with the closest hit over the tree, the large-primitive list, the sliver
// list and - deferred or in place, as in trace_ray - the far field.  Called by all 32 lanes.
__device__ __forceinline__ int finish_closest(const DeviceScene& sc, bool active, V3 O, V3 d, HitRec& hit, SlowQ q, int ca)
{
    bool found = false, need = false, linear = false, pending = false;
    const bool nan_hit = active && hit.t != hit.t;      // a NaN ray: k_closest has the answer (device_scene.h nan_leaf)
    if (active && !nan_hit) {
        if (hit.leaf == -2) {
            need = true; linear = true;
            hit.t = __int_as_float(0x7f800000); hit.leaf = -1; hit.prim = 0x7fffffff;
            if (sc.diag) atomicAdd(sc.diag + 1, 1u);
        } else {
            found = hit.leaf >= 0;
            found = big_scan<false>(sc, O, d, hit) || found;
            const bool zero_dir = (d.x == 0.0f && d.y == 0.0f && d.z == 0.0f);
            if (sc.farfield && !zero_dir) {
                if (sc.n_always) found = always_scan<false>(sc, O, d, hit, found);
                need = !found || hit.t >= sc.far_tmin;
                if (need && sc.diag) atomicAdd(sc.diag, 1u);
            }
        }
        if (need && q.rays) {
            const unsigned slot = atomicAdd(q.count, 1u);
            if (slot < q.cap) {
                SlowRay r;
                r.o = make_float4(O.x, O.y, O.z, hit.t);
                r.d = make_float4(d.x, d.y, d.z, __int_as_float(hit.prim));
                r.c = make_int4(linear ? 1 : 0, ca, 0, found ? hit.leaf : -1);
                q.rays[slot] = r;
                SlowRes a; a.key = slow_key(hit.t, hit.prim); a.found = 0; a.pad = 0;
                q.res[slot] = a;
                pending = true; need = false;
            }
        }
    }
    if (sc.farfield) found = warp_slow_path<false>(sc, need, linear, O, d, hit, found);
    if (nan_hit) return hit.leaf >= 0 ? TR_HIT : TR_MISS;
    return pending ? TR_PENDING : (found ? TR_HIT : TR_MISS);
}

// Deferred slow rays.  Every slow ray must see every primitive's filter record (far scan) or every
// primitive (linear fallback), so the kernel is bound by record traffic unless rays share it: a block
// takes SLOW_RPB rays, streams the 16-byte filter records through shared memory in tiles of SLOW_TILE,
// and each of its 8 warps runs its 4 rays over the tile (lane l takes records l, l+32, ...).  The
// exact test (the reference's, prim_test, ~100 instructions) is needed for ~0.5 % of the (ray, record)
// pairs only; run where the filter passes it kept 1-2 lanes of a warp busy and cost more than the
// filtering itself.  So the survivors go to a per-warp queue of (ray, record) pairs and are tested 32 at
// a time, one per lane, with the record read from global memory.  blockIdx.y slices the record range
// (few slow rays -> more slices, so the GPU stays busy); slices combine through atomicMin on the
// packed (t, prim) key / the found flag.
#define SLOW_RPB 32
#define SLOW_TILE 256
#define SLOW_RPW (SLOW_RPB / 8)
template <bool ANY>
__device__ __forceinline__ void slow_exact(const DeviceScene& sc, unsigned long long entry, const float4* __restrict__ s_O,
                                           const float4* __restrict__ s_D, unsigned long long* s_key, int* s_found)
{
    const int j = (int)(entry >> 32);
    const unsigned i = (unsigned)entry;
    const float4 o = s_O[j], dd = s_D[j];
    const unsigned long long key = *reinterpret_cast<volatile unsigned long long*>(s_key + j);
    float t; int prim;
    if (prim_test<true>(sc.prims + i, mk(o.x, o.y, o.z), mk(dd.x, dd.y, dd.z), __uint_as_float((unsigned)(key >> 32)),
                        ANY ? 0x7fffffff : (int)(unsigned)(key & 0xffffffffull), t, prim)) {
        if (ANY) s_found[j] = 1;
        else atomicMin(s_key + j, slow_key(t, prim));
    }
}
template <bool ANY>
__global__ void __launch_bounds__(256)
k_slow(DeviceScene sc, const SlowRay* __restrict__ rays, unsigned n, SlowRes* __restrict__ res, int chunk,
       const unsigned int* __restrict__ idx)     // idx != nullptr: the rays to answer are rays[idx[0 .. n)]
{
    __shared__ float4 s_far[SLOW_TILE];
    __shared__ float4 s_O[8][SLOW_RPW], s_D[8][SLOW_RPW];
    __shared__ unsigned long long s_key[8][SLOW_RPW];      // best (t, prim) so far; any hit: (tmax, INT_MAX)
    __shared__ int s_found[8][SLOW_RPW];
    __shared__ unsigned long long s_q[8][64];              // survivors of the filter: (ray << 32) | record
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    // ray j of the block's batch goes to warp j % 8: a batch of 5 rays keeps 5 warps busy, not 2
    const unsigned first = blockIdx.x * SLOW_RPB + warp;
    bool my_live = false, my_lin = false;
    unsigned long long key0 = 0ull;
    if (lane < SLOW_RPW) {
        const unsigned e0 = first + lane * 8u;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f), d = o;
        if (e0 < n) {
            const unsigned e = idx ? __ldg(idx + e0) : e0;
            o = __ldg(&rays[e].o); d = __ldg(&rays[e].d);
            my_lin = (__ldg(&rays[e].c).x & 1) != 0;
            my_live = true;
        }
        s_O[warp][lane] = o; s_D[warp][lane] = d;
        key0 = ((unsigned long long)__float_as_uint(o.w) << 32) | (unsigned)__float_as_int(d.w);
        s_key[warp][lane] = key0;
        s_found[warp][lane] = 0;
    }
    unsigned live_mask = __ballot_sync(0xffffffffu, my_live) & ((1u << SLOW_RPW) - 1u);
    const unsigned lin_mask = __ballot_sync(0xffffffffu, my_lin) & ((1u << SLOW_RPW) - 1u);
    __syncwarp();
    float4 dj[SLOW_RPW];
#pragma unroll
    for (int j = 0; j < SLOW_RPW; j++) dj[j] = s_D[warp][j];
    unsigned q_len = 0;
    const int nl = min(sc.n_all, ((int)blockIdx.y + 1) * chunk);
    for (int base = (int)blockIdx.y * chunk; base < nl; base += SLOW_TILE) {
        {   // stage one tile of filter records: thread t brings record base + t
            const int i = base + (int)threadIdx.x;
            s_far[threadIdx.x] = (i < nl) ? __ldg(sc.far + i) : make_float4(0.f, 0.f, 0.f, -1.0f);
        }
        __syncthreads();
        if (ANY) {
#pragma unroll
            for (int j = 0; j < SLOW_RPW; j++)                               // rays that are answered leave
                if (((live_mask >> j) & 1u) && *reinterpret_cast<volatile int*>(&s_found[warp][j])) live_mask &= ~(1u << j);
        }
        // record in the outer loop, the warp's rays in the inner one: one shared-memory load serves four filter tests
#pragma unroll 2
        for (int k = 0; k < SLOW_TILE / 32; k++) {
            if (ANY && lin_mask) {
                // a ray from outside the scene is accepted by about every fortieth triangle it meets (float noise, fargrid.cuh):
                // stop testing it as soon as one has
                __syncwarp();
#pragma unroll
                for (int j = 0; j < SLOW_RPW; j++)
                    if (((live_mask >> j) & 1u) && *reinterpret_cast<volatile int*>(&s_found[warp][j])) live_mask &= ~(1u << j);
                if (!live_mask) break;
            }
            const int sl = k * 32 + lane, i = base + sl;
            const bool in_range = i < nl;
            const float4 fr = s_far[sl];
#pragma unroll
            for (int j = 0; j < SLOW_RPW; j++) {
                if (!((live_mask >> j) & 1u)) continue;
                // |N.d| <= thr (filter only: FMA is fine); a linear-fallback ray takes every record
                const float nd = __fmaf_rn(fr.x, dj[j].x, __fmaf_rn(fr.y, dj[j].y, fr.z * dj[j].z));
                const bool pass = in_range && (((lin_mask >> j) & 1u) || fabsf(nd) <= fr.w);
                const unsigned mask = __ballot_sync(0xffffffffu, pass);
                if (mask == 0u) continue;
                if (pass) s_q[warp][q_len + (unsigned)__popc(mask & lt_mask)] = ((unsigned long long)j << 32) | (unsigned)i;
                q_len += (unsigned)__popc(mask);
                if (q_len >= 32u) {
                    __syncwarp();
                    slow_exact<ANY>(sc, s_q[warp][lane], s_O[warp], s_D[warp], s_key[warp], s_found[warp]);
                    __syncwarp();
                    const unsigned long long tail = (lane + 32u < q_len) ? s_q[warp][lane + 32] : 0ull;
                    __syncwarp();
                    if (lane + 32u < q_len) s_q[warp][lane] = tail;
                    q_len -= 32u;
                    __syncwarp();
                }
            }
        }
        if (!__syncthreads_or(live_mask ? 1 : 0)) break;                     // also the barrier before the next tile
    }
    __syncwarp();
    if ((unsigned)lane < q_len) slow_exact<ANY>(sc, s_q[warp][lane], s_O[warp], s_D[warp], s_key[warp], s_found[warp]);
    __syncwarp();
    if (lane < SLOW_RPW && my_live) {
        const unsigned e0 = first + lane * 8u;
        const unsigned e = idx ? __ldg(idx + e0) : e0;
        if (ANY) { if (s_found[warp][lane]) res[e].found = 1; }
        else { const unsigned long long k = s_key[warp][lane]; if (k < key0) atomicMin(&res[e].key, k); }
    }
}

// ---- deferred far-scan rays through the far-field direction grid (fargrid.cuh) ----------------------------
// A flush of deferred rays (slow_launch) goes three ways:
//   rays that start in the scene            sorted by the cell of their direction (counting sort: k_fg_bin, scan, k_fg_order),
//                                           then k_fg_scan: a warp takes 32 (or 8) consecutive rays of that order, the rays of
//                                           one cell share the gathers of its list (lane = entry), a two-stage filter, the
//                                           survivors queued per warp for the reference's exact test, 32 at a time;
//   any-hit rays from outside, unbounded    k_fg_lin_first: the first acceptor in the cell of their direction, one thread per ray;
//   every other ray from outside            the arc kernels: k_fg_arc_pre -> k_fg_arc_first -> k_fg_arc -> k_fg_arc_items,
//                                           and k_lin_near for the part of the ray that comes back to the scene.
__global__ void __launch_bounds__(256)
k_fg_bin(const SlowRay* __restrict__ rays, unsigned n, int K, unsigned int* __restrict__ hist, unsigned int* __restrict__ cellof,
         unsigned int* __restrict__ rank, unsigned int* __restrict__ lin_idx, unsigned int* __restrict__ lin_count, int bin_lin,
         unsigned int* __restrict__ first_idx, unsigned int* __restrict__ first_count)
{
    const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = e < n;
    // where the ray goes: 0 nowhere (no cell: no triangle accepts a zero / NaN direction, cpp:371), 1 the sorted order of
    // k_fg_scan, 2 the list of rays from outside for the arc kernels, 3 any-hit from outside, unbounded: k_fg_lin_first
    int to = 0, cell = -1;
    if (valid) {
        const bool lin = (__ldg(&rays[e].c).x & 1) != 0;
        // a ray from outside the scene: closest hit -> its own kernels (k_fg_arc* walk its cells in order of t); any hit (bin_lin)
        // -> first the cell of its direction (k_fg_lin_first: out there, far along the ray, almost every entry of that cell accepts
        // it), unless it is bounded (the shadow ray of a point light ends long before it is "far along d")
        if (lin && !bin_lin) to = 2;
        else {
            const float4 d = __ldg(&rays[e].d);
            cell = fg_cell_of_dir(mk(d.x, d.y, d.z), K);
            if (cell < 0) to = lin ? 2 : 0;                      // (spheres may still accept it: k_lin_near)
            else if (!lin) to = 1;
            else to = (__ldg(&rays[e].o).w < 3.0e38f) ? 2 : 3;
        }
    }
    // (one atomic per warp and list: 86 M rays of a frame go to the same two counters)
    const unsigned s2 = warp_alloc(lin_count, to == 2);
    const unsigned s3 = warp_alloc(first_count, to == 3);
    if (!valid) return;
    if (to == 2) lin_idx[s2] = e;
    if (to == 3) first_idx[s3] = e;
    cellof[e] = (to == 1 || to == 3) ? (unsigned)cell : 0xffffffffu;
    if (hist) rank[e] = to == 1 ? atomicAdd(hist + cell, 1u) : 0xffffffffu;
}
__global__ void __launch_bounds__(256)
k_fg_order(unsigned n, const unsigned int* __restrict__ cellof, const unsigned int* __restrict__ rank,
           const unsigned int* __restrict__ cstart, unsigned int* __restrict__ order)
{
    const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const unsigned cell = cellof[e];
    if (cell != 0xffffffffu && rank[e] != 0xffffffffu) order[cstart[cell] + rank[e]] = e;
}

template <bool ANY>
__device__ __forceinline__ void fg_exact(const DeviceScene& sc, unsigned long long entry, const float4* __restrict__ s_O,
                                         const float4* __restrict__ s_D, unsigned long long* s_key, int* s_found)
{
    const int j = (int)(entry >> 32);
    const unsigned i = (unsigned)entry;
    const float4 o = s_O[j], dd = s_D[j];
    const unsigned long long key = *reinterpret_cast<volatile unsigned long long*>(s_key + j);
    float t; int prim;
    if (prim_test<true>(sc.prims + i, mk(o.x, o.y, o.z), mk(dd.x, dd.y, dd.z), __uint_as_float((unsigned)(key >> 32)),
                        ANY ? 0x7fffffff : (int)(unsigned)(key & 0xffffffffull), t, prim)) {
        if (ANY) s_found[j] = 1;
        else atomicMin(s_key + j, slow_key(t, prim));
    }
}

// One warp takes FG_G consecutive rays of the sorted order; the rays of one cell among them form a segment that runs over the
// cell's list 32 entries at a time (lane = entry: its record is gathered once and serves every ray of the segment).
// rays per warp (template G): the rays of one cell among them share the gathers of that cell's list - 32 where cells hold many
// rays (the AO flush of a whole 4K frame: ~12 per cell); 8 where they hold one or two (closest-hit flushes, a rank's share of a
// frame), so that four times as many warps wait on their gathers side by side
#define FG_WARPS 4
#define FG_U4 1      // list entries per lane and iteration (4 was measured slower: the rays from outside the scene stop after a few entries)
template <int G>
struct FgWarp {
    float4 O[G], D[G];
    unsigned long long key[G];
    int found[G];
    unsigned cell[G], e[G];
    unsigned long long q[64];
};

template <bool ANY, int G>
__device__ __forceinline__ void fg_flush(const DeviceScene& sc, FgWarp<G>& sh, unsigned& q_len, bool all_of_it)
{
    const int lane = threadIdx.x & 31;
    while (q_len >= 32u || (all_of_it && q_len > 0u)) {
        __syncwarp();
        if ((unsigned)lane < q_len) fg_exact<ANY>(sc, sh.q[lane], sh.O, sh.D, sh.key, sh.found);
        __syncwarp();
        const unsigned long long tail = (lane + 32u < q_len) ? sh.q[lane + 32] : 0ull;
        __syncwarp();
        if (lane + 32u < q_len) sh.q[lane] = tail;
        q_len = q_len > 32u ? q_len - 32u : 0u;
        __syncwarp();
    }
}

// rays [j0, j1) of the warp against one list
template <bool ANY, int G>
__device__ __forceinline__ void fg_segment(const DeviceScene& sc, FgWarp<G>& sh, int j0, int j1, const uint32_t* __restrict__ list,
                                           unsigned long long len, bool has_k6, float dno)
{
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    unsigned live = (j1 >= 32 ? 0xffffffffu : ((1u << j1) - 1u)) & ~((1u << j0) - 1u);
    // any hit: the rays that are answered leave (found[] changes in fg_flush only)
    if (ANY) { live &= ~__ballot_sync(0xffffffffu, (lane < G && *reinterpret_cast<volatile int*>(&sh.found[lane < G ? lane : 0]) != 0)); if (!live) return; }
    unsigned q_len = 0;
    for (unsigned long long base = 0; base < len; base += 32 * FG_U4) {
        // this lane's FG_U4 entries (independent gathers in flight together: the loop is bound by their latency):
        // (N, thr) for stage 1, (D, T) for stage 2
        float4 frs[FG_U4]; float eDs[FG_U4], eTs[FG_U4]; unsigned ids[FG_U4], ents[FG_U4];
#pragma unroll
        for (int u = 0; u < FG_U4; u++) { const unsigned long long idx = base + 32 * u + lane; ents[u] = idx < len ? __ldg(list + idx) : 0xffffffffu; }
#pragma unroll
        for (int u = 0; u < FG_U4; u++) {
            frs[u] = make_float4(0.f, 0.f, 0.f, -1.0f); eDs[u] = 0.f; eTs[u] = 0.f; ids[u] = 0u;
            if (base + 32 * u + lane < len) {
                const unsigned ent = ents[u];
                ids[u] = has_k6 ? (ent & FG_ID_MASK) : ent;
                const float4 fa = __ldg(sc.fg_A + ids[u]);
                const float2 fb = __ldg(sc.fg_B + ids[u]);
                eTs[u] = fg_entry_T(fb.x, has_k6 ? (ent >> FG_ID_BITS) : 0u);
                frs[u] = make_float4(fa.x, fa.y, fa.z, __fdividef(fb.y, eTs[u]) * 1.00001f + FG_ND_SLACK);    // (2 ulp: within the factor)
                eDs[u] = fa.w;
            }
        }
#pragma unroll
        for (int u = 0; u < FG_U4; u++) {
            if (base + 32 * u >= len) break;
            const float4 fr = frs[u]; const float eD = eDs[u], eT = eTs[u] * 0.999998f; const unsigned id = ids[u];
            for (unsigned rest = live; rest;) {
                const int j = __ffs((int)rest) - 1;
                rest &= rest - 1u;
                // both stages for all 32 entries at once, without a branch (one entry in seven passes stage 1: some lane always does)
                const float4 dj = sh.D[j], oj = sh.O[j];
                const float nd = __fmaf_rn(fr.x, dj.x, __fmaf_rn(fr.y, dj.y, fr.z * dj.z));
                const float no = __fmaf_rn(fr.x, oj.x, __fmaf_rn(fr.y, oj.y, __fmaf_rn(fr.z, oj.z, eD)));
                const float and_ = fabsf(nd);
                // stage 1: the band of the cell; stage 2: t = -(N.O + D) / (N.d) >= T with this ray's origin, and t > 0
                const float x = __fmaf_rn(and_ - FG_ND_SLACK, eT, -dno);         // (eT: T * 0.999998)
                const bool pass = and_ <= fr.w && and_ > FG_ND_MIN && fabsf(no) >= x && (x <= dno || ((no < 0.f) != (nd < 0.f)));
                const unsigned mask = __ballot_sync(0xffffffffu, pass);
                if (mask == 0u) continue;
                if (pass) sh.q[q_len + (unsigned)__popc(mask & lt_mask)] = ((unsigned long long)j << 32) | id;
                q_len += (unsigned)__popc(mask);
                if (q_len >= 32u) {
                    fg_flush<ANY, G>(sc, sh, q_len, false);
                    if (ANY) { live &= ~__ballot_sync(0xffffffffu, (lane < G && *reinterpret_cast<volatile int*>(&sh.found[lane < G ? lane : 0]) != 0)); rest &= live; }
                }
            }
            if (ANY && !live) break;
        }
        if (ANY && !live) break;
    }
    fg_flush<ANY, G>(sc, sh, q_len, true);
}

template <bool ANY, int G>
__global__ void __launch_bounds__(32 * FG_WARPS)
k_fg_scan(DeviceScene sc, const SlowRay* __restrict__ rays, SlowRes* __restrict__ res, const unsigned int* __restrict__ order,
          const unsigned int* __restrict__ cellof, const unsigned int* __restrict__ total_ptr, unsigned n_direct,
          unsigned int* __restrict__ lin_idx, unsigned int* __restrict__ lin_count)
{
    __shared__ FgWarp<G> shw[FG_WARPS];
    FgWarp<G>& sh = shw[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    // order == nullptr: a small flush, not worth the sort by cell - the rays in queue order (n_direct of them)
    const unsigned total = order ? __ldg(total_ptr) : n_direct;
    const unsigned long long pos0 = ((unsigned long long)blockIdx.x * FG_WARPS + (threadIdx.x >> 5)) * G;
    if (pos0 >= total) return;
    const int nb = (int)min((unsigned long long)G, total - pos0);
    bool my_lin = false;
    if (lane < nb) {
        const unsigned e = order ? order[pos0 + lane] : (unsigned)(pos0 + lane);
        const float4 o = __ldg(&rays[e].o), d = __ldg(&rays[e].d);
        sh.e[lane] = e;
        sh.O[lane] = o; sh.D[lane] = d;
        sh.key[lane] = ((unsigned long long)__float_as_uint(o.w) << 32) | (unsigned)__float_as_int(d.w);
        sh.found[lane] = 0;
        // (unsorted flush, any hit: the rays from outside the scene are k_fg_lin_first's)
        my_lin = ANY && !order && (__ldg(&rays[e].c).x & 1) != 0;
        sh.cell[lane] = my_lin ? 0xffffffffu : cellof[e];
    }
    __syncwarp();
    // the float evaluation of N.O + D here vs in the reference (cpp:377, 381): both within 28 u E of the true value
    const float dno = 3.4e-6f * sc.extent;
    for (int j0 = 0; j0 < nb;) {
        const unsigned cell = sh.cell[j0];
        int j1 = j0 + 1;
        while (j1 < nb && sh.cell[j1] == cell) j1++;
        if (cell != 0xffffffffu) {                       // (unsorted flush: rays without a cell - k_fg_bin has routed them elsewhere)
            const unsigned long long b = __ldg(sc.fg_start + cell), en = __ldg(sc.fg_start + cell + 1);
            fg_segment<ANY, G>(sc, sh, j0, j1, sc.fg_entries + b, en - b, true, dno);
        } else for (int j = j0; j < j1; j++) sh.found[j] = -1;
        j0 = j1;
    }
    if (sc.fg_n_wide > 0) fg_segment<ANY, G>(sc, sh, 0, nb, sc.fg_wide, (unsigned long long)sc.fg_n_wide, false, dno);
    __syncwarp();
    if (lane < nb) {
        const unsigned e = sh.e[lane];
        if (sh.found[lane] == -1) { }                                     // no cell: not this kernel's ray
        else if (ANY) {
            if (sh.found[lane]) res[e].found = 1;
        } else {
            const unsigned long long k = sh.key[lane];
            if (k < res[e].key) res[e].key = k;
        }
    }
}

// Any-hit rays from outside the scene, first attempt: far along such a ray almost every triangle listed in the cell of its
// direction accepts the plane hit (the bounds of the list assume an in-scene origin, so they are not used: every entry gets the
// reference's test, whose plane half rejects cheaply).  One thread per ray walks the list of that one cell until the first
// acceptor - two or three entries on average.  A ray that finds none goes on to k_fg_arc / k_lin_near, which are complete.
__global__ void __launch_bounds__(128)
k_fg_lin_first(DeviceScene sc, const SlowRay* __restrict__ rays, SlowRes* __restrict__ res, const unsigned int* __restrict__ first_idx,
               const unsigned int* __restrict__ first_count, const unsigned int* __restrict__ cellof,
               unsigned int* __restrict__ lin_idx, unsigned int* __restrict__ lin_count)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= __ldg(first_count)) return;
    const unsigned e = first_idx[i];
    const float4 o = __ldg(&rays[e].o), d = __ldg(&rays[e].d);
    const V3 O = mk(o.x, o.y, o.z), D = mk(d.x, d.y, d.z);
    const unsigned cell = cellof[e];
    const unsigned long long b = __ldg(sc.fg_start + cell), en = __ldg(sc.fg_start + cell + 1);
    bool found = false;
    for (unsigned long long k = b; k < en && !found; k++) {
        const unsigned id = __ldg(sc.fg_entries + k) & FG_ID_MASK;
        float t; int prim;
        found = prim_test<true>(sc.prims + id, O, D, o.w, 0x7fffffff, t, prim);
    }
    if (found) res[e].found = 1;
    else lin_idx[atomicAdd(lin_count, 1u)] = e;
}

// ---- rays that start outside the scene (children of far-field hits, 10^5..10^8 units away) ------------------------
// The reference sends them through the same loop over every triangle (cpp:476-521).  Out there they meet two kinds of
// acceptors, answered by two kernels over the list of such rays:
//   far regime   (k_fg_arc)   a triangle accepts the plane hit P far from itself: the direction of P seen from the scene
//                             lies in the triangle's strip and |P| >= T (fargrid.cuh), i.e. the triangle is in the list
//                             of the direction cell of P.  While t grows, the direction of P(t) = O + t d walks along a
//                             great-circle arc from the direction of O to d: the cells on that arc are visited in order
//                             of t, each one's entries filtered by T <= |P|, the rest tested exactly.
//   near regime  (k_lin_near) the ray comes back to the scene and hits a triangle for real: a tree traversal with the
//                             leaf boxes inflated by the rounding of P = O + t d for such an origin; plus the handful of
//                             primitives the tree and the grid do not hold (large ones, slivers), and all
//                             spheres when the ray is aimed at the scene (cpp:426's discriminant is noise out there).
#define ARC_WARPS 4
#define ARC_MAX_CELLS 4096u
struct ArcItem { unsigned e; int cell; float rmax; float t_in; };   // t_in: the ray is in the cell from this parameter on (a lower bound)

// one ray of a warp against the lists of cells of the direction grid: the state both arc kernels share
template <bool ANY>
struct ArcRay {
    const DeviceScene& sc;
    unsigned* q;                       // the warp's queue of candidates [64]
    V3 O, d;
    float tlim; int plim;              // any hit: tmax; closest hit: the best (t, prim) so far
    bool found;
    unsigned q_len, n_cells, n_exact;
    double aa, ad, dd;                 // |P(t)|^2 = aa + 2 t ad + t^2 dd, P relative to the grid's centre
    float Olen, dno_far;
    // list entries per lane and iteration, their gathers in flight together (the loop is bound by that latency): a closest-hit
    // ray reads the whole list anyway; an any-hit ray mostly stops within the first slab
    static constexpr int ARC_U = ANY ? 1 : 2;

    __device__ __forceinline__ ArcRay(const DeviceScene& sc_, unsigned* q_, float4 ro, float4 rd) : sc(sc_), q(q_) {
        O = mk(ro.x, ro.y, ro.z); d = mk(rd.x, rd.y, rd.z);
        tlim = ro.w; plim = ANY ? 0x7fffffff : __float_as_int(rd.w);
        found = false; q_len = 0; n_cells = 0; n_exact = 0;
        const double Ax = (double)O.x - sc.fg_center[0], Ay = (double)O.y - sc.fg_center[1], Az = (double)O.z - sc.fg_center[2];
        aa = Ax * Ax + Ay * Ay + Az * Az; ad = Ax * (double)d.x + Ay * (double)d.y + Az * (double)d.z;
        dd = (double)d.x * d.x + (double)d.y * d.y + (double)d.z * d.z;
        Olen = fabsf(O.x) + fabsf(O.y) + fabsf(O.z);
        dno_far = 1e-6f * (Olen + sc.extent);      // float evaluation of N.O + D here and in the reference: <= 8 u (|O|_1 + |D|) each
    }
    // the queued candidates get the reference's exact test, 32 at a time
    __device__ __forceinline__ void flush(bool all_of_it) {
        const int lane = threadIdx.x & 31;
        const float inf = __int_as_float(0x7f800000);
        while (q_len >= 32u || (all_of_it && q_len > 0u)) {
            __syncwarp();
            float t = inf; int prim = 0x7fffffff;
            if ((unsigned)lane < q_len) {
                float tt; int pp;
                if (prim_test<true>(sc.prims + q[lane], O, d, tlim, plim, tt, pp)) { t = tt; prim = pp; }
            }
            n_exact += min(q_len, 32u);
            if (ANY) { if (__any_sync(0xffffffffu, prim != 0x7fffffff)) found = true; }
            else {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float t2 = __shfl_xor_sync(0xffffffffu, t, o); const int p2 = __shfl_xor_sync(0xffffffffu, prim, o);
                    if (t2 < t || (t2 == t && p2 < prim)) { t = t2; prim = p2; }
                }
                if (prim != 0x7fffffff) { tlim = t; plim = prim; found = true; }
            }
            __syncwarp();
            const unsigned tail = (lane + 32u < q_len) ? q[lane + 32] : 0u;
            __syncwarp();
            if (lane + 32u < q_len) q[lane] = tail;
            q_len = q_len > 32u ? q_len - 32u : 0u;
            __syncwarp();
        }
    }
    // The triangle (fa = (N, D), far field from T on) is listed in every cell its strip crosses, but this ray meets its plane at
    // one parameter t = -(N.O + D) / (N.d) only: bracket it ([t_lo, t_hi] covers the float evaluation here and in cpp:367-381)
    // and ask for t > 0, t within the limit, and |P(t)| >= T somewhere in the bracket (|P(t)| is convex).
    __device__ __forceinline__ bool entry_can_accept(float4 fa, float T) const {
        const float nd = __fmaf_rn(fa.x, d.x, __fmaf_rn(fa.y, d.y, fa.z * d.z));
        const float no = __fmaf_rn(fa.x, O.x, __fmaf_rn(fa.y, O.y, __fmaf_rn(fa.z, O.z, fa.w)));
        const float and_ = fabsf(nd), ano = fabsf(no);
        bool pass = and_ > FG_ND_MIN && !(ano > dno_far && ((no < 0.f) == (nd < 0.f)));
        if (pass) {
            // (approximate divisions, 2 ulp: the factors leave room for them)
            const float t_lo = __fdividef(fmaxf(ano - dno_far, 0.f), and_ + FG_ND_SLACK) * 0.999998f;
            pass = t_lo <= tlim;
            if (pass && and_ > 2.0f * FG_ND_SLACK) {
                const double t_hi = (double)(__fdividef(ano + dno_far, and_ - FG_ND_SLACK) * 1.000002f), tl = t_lo;
                const double p_lo = aa + tl * (2.0 * ad + tl * dd), p_hi = aa + t_hi * (2.0 * ad + t_hi * dd);
                const double need = fmax((double)T * 0.9999 - 2e-6 * Olen, 0.0);
                pass = fmax(p_lo, p_hi) >= need * need;
            }
        }
        return pass;
    }
    // one candidate through the reference's test, by this thread alone
    __device__ __forceinline__ void exact_one(unsigned id) {
        float tt; int pp;
        n_exact++;
        if (prim_test<true>(sc.prims + id, O, d, tlim, plim, tt, pp)) { found = true; if (!ANY) { tlim = tt; plim = pp; } }
    }
    // the entries of one cell that can accept at |P| <= rmax
    __device__ __forceinline__ bool cell_can_accept(int cell, float rmax) const {      // some entry of the cell has T <= rmax
        return __uint_as_float(__ldg(sc.fg_cell_tmin + cell)) <= rmax;                    // (empty cell: NaN)
    }
    __device__ __forceinline__ void cell(int cell, float rmax) {
        const unsigned long long b = __ldg(sc.fg_start + cell), en = __ldg(sc.fg_start + cell + 1);
        n_cells++;
        list(sc.fg_entries, b, en, true, rmax);
    }
    // the same over any list of primitive indices (has_k6: entries of the grid, with their 6-bit factor on T)
    __device__ __forceinline__ void list(const uint32_t* __restrict__ entries, unsigned long long b, unsigned long long en, bool has_k6, float rmax) {
        const int lane = threadIdx.x & 31;
        const unsigned lt_mask = (1u << lane) - 1u;
        for (unsigned long long base = b; base < en; base += 32 * ARC_U) {
            // ARC_U entries per lane: their gathers are in flight together (the loop is bound by that latency)
            unsigned ents[ARC_U], ids[ARC_U]; float Ts[ARC_U]; float4 fas[ARC_U];
#pragma unroll
            for (int u = 0; u < ARC_U; u++) { const unsigned long long idx = base + 32 * u + lane; ents[u] = idx < en ? __ldg(entries + idx) : 0xffffffffu; }
#pragma unroll
            for (int u = 0; u < ARC_U; u++) {
                ids[u] = has_k6 ? (ents[u] & FG_ID_MASK) : ents[u]; Ts[u] = __int_as_float(0x7f800000); fas[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (base + 32 * u + lane < en) { Ts[u] = fg_entry_T(__ldg(sc.fg_B + ids[u]).x, has_k6 ? (ents[u] >> FG_ID_BITS) : 0u); fas[u] = __ldg(sc.fg_A + ids[u]); }
            }
#pragma unroll
            for (int u = 0; u < ARC_U; u++) {
                if (base + 32 * u >= en) break;
                bool pass = false; const unsigned id = ids[u];
                if (base + 32 * u + lane < en) pass = Ts[u] <= rmax && entry_can_accept(fas[u], Ts[u]);
                const unsigned mask = __ballot_sync(0xffffffffu, pass);
                if (mask) {
                    if (pass) q[q_len + (unsigned)__popc(mask & lt_mask)] = id;
                    q_len += (unsigned)__popc(mask);
                    // any hit: the first acceptor ends the ray, but a batch of exact tests costs the same for 2 candidates as
                    // for 32 (and a far-away ray is accepted by one candidate in ten or so): wait for a dozen
                    if (q_len >= (ANY ? 12u : 32u)) flush(ANY);
                    if (ANY && found) return;
                }
            }
        }
        flush(true);
    }
};

// The walk of one ray over the cells of the direction grid, in order of t.  It runs in double: far out, a float t no longer
// resolves a cell (ulp(t) / |P| exceeds the cells' slack when the ray has come a long way and passes near the scene); the cells
// it names are a superset of those the reference's float P visits.
struct ArcWalk {
    double Ax, Ay, Az, dx, dy, dz;     // P(t) = A + t d relative to the grid's centre
    double hole_lo, hole_hi;           // where |P(t)| < 0.9 T_min no far-field acceptor exists
    double t_end, t;
    // after leaving a cell through a wall, the next cell is looked up a hair BEYOND that wall (relative nudge along one axis), so
    // a path that runs along a wall cannot bounce between the lookup and the wall arithmetic.  0: none, 1 + 2 axis + (negative)
    int nudge;
    unsigned n_it;
    int K;

    __device__ __forceinline__ void init(const DeviceScene& sc, V3 O, V3 d, double aa, double ad, double dd, double t_end_) {
        Ax = (double)O.x - sc.fg_center[0]; Ay = (double)O.y - sc.fg_center[1]; Az = (double)O.z - sc.fg_center[2];
        dx = d.x; dy = d.y; dz = d.z;
        K = sc.fg_K;
        const double dinf = (double)__int_as_float(0x7f800000);
        hole_lo = dinf; hole_hi = -dinf;
        const double Tm = 0.9 * (double)sc.fg_tmin;
        const double disc = ad * ad - dd * (aa - Tm * Tm);
        if (disc > 0.0 && dd > 0.0 && Tm < 1e18) { const double sq = sqrt(disc); hole_lo = (-ad - sq) / dd; hole_hi = (-ad + sq) / dd; }
        t_end = t_end_; t = 0.0; nudge = 0; n_it = 0;
    }
    // the next cell of the arc, the parameter at which the ray enters it and the largest |P| inside it; false: the arc is over
    __device__ __forceinline__ bool next(int& cell_out, double& t_in, float& rmax) {
        const float inf = __int_as_float(0x7f800000);
        const double dinf = (double)inf;
        const double h = 2.0 / (double)K;
        while (n_it < 200000u) {
            n_it++;
            if (!(t < t_end)) return false;
            if (t > hole_lo && t < hole_hi) { t = hole_hi; nudge = 0; continue; }
            const double Px = Ax + dx * t, Py = Ay + dy * t, Pz = Az + dz * t;
            const double p2 = Px * Px + Py * Py + Pz * Pz;
            const double plen = sqrt(p2);
            const double nv = nudge ? (((nudge - 1) & 1) ? 1e-9 : -1e-9) * plen : 0.0;
            const int nax = nudge ? ((nudge - 1) >> 1) : -1;
            const int cell = fg_cell_of_point_d(Px + (nax == 0 ? nv : 0.0), Py + (nax == 1 ? nv : 0.0), Pz + (nax == 2 ? nv : 0.0), K);
            const double adv = 3e-7 * plen + 1e-30;
            if (cell < 0) { t = t + adv + 1e-9 * fmax(1.0, t); continue; }
            const int face = cell / (K * K), iv = (cell / K) % K, iu = cell % K;
            const int ax = face >> 1;
            const double sg = (face & 1) ? -1.0 : 1.0;
            const double Aw = (ax == 0 ? Ax : (ax == 1 ? Ay : Az)) * sg, dw = (ax == 0 ? dx : (ax == 1 ? dy : dz)) * sg;
            const double Au = (ax == 0 ? Ay : (ax == 1 ? Az : Ax)), du = (ax == 0 ? dy : (ax == 1 ? dz : dx));
            const double Av = (ax == 0 ? Az : (ax == 1 ? Ax : Ay)), dv = (ax == 0 ? dz : (ax == 1 ? dx : dy));
            // where does P(t) leave the cell: f(t) = P_u - b P_w changes sign at the walls b = u0 (f >= 0 inside), u1 (f <= 0 inside),
            // likewise v.  Only walls the ray moves OUT through count; one it is already beyond (rounding) means "leave now".
            double t_exit = dinf;
            int k_exit = -1;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const bool upper = (k & 1) != 0;
                const double b = ((double)((k < 2 ? iu : iv) + (upper ? 1 : 0))) * h - 1.0;
                const double f0 = (k < 2 ? Au : Av) - b * Aw, f1 = (k < 2 ? du : dv) - b * dw;
                if (upper ? (f1 > 0.0) : (f1 < 0.0)) { const double tc = fmax(-f0 / f1, t); if (tc < t_exit) { t_exit = tc; k_exit = k; } }
            }
            // the nudge for the next lookup: across wall k_exit, i.e. along axis u (k < 2) or v, up (odd k) or down
            nudge = 0;
            if (k_exit >= 0) nudge = 1 + 2 * (k_exit < 2 ? (ax + 1) % 3 : (ax + 2) % 3) + ((k_exit & 1) ? 1 : 0);
            double t_out = fmin(t_exit, t_end);
            if (t < hole_lo) t_out = fmin(t_out, hole_lo);
            rmax = inf;
            if (t_out < 1e300) {
                const double Qx = Ax + dx * t_out, Qy = Ay + dy * t_out, Qz = Az + dz * t_out;
                const double r = sqrt(fmax(p2, Qx * Qx + Qy * Qy + Qz * Qz)) * 1.00001 + 1.0;      // |P(t)| is convex in t
                rmax = r < 3.0e38 ? (float)r * 1.000001f : inf;
            }
            cell_out = cell; t_in = t;
            t = (t_out < 1e300) ? t_out + adv : dinf;                        // (inf: the next call ends the arc)
            return true;
        }
        return false;
    }
};

// The first cells of every arc, one THREAD per ray (the walk is serial arithmetic: a warp per ray would run it 32 times over).
// k_fg_arc takes the cells from here and goes on walking by itself only when these did not settle the ray.
#define ARC_PRE 4
struct ArcPre { double t; double t_in[ARC_PRE]; int cell[ARC_PRE]; float rmax[ARC_PRE]; int n; int nudge; unsigned n_it; int done; };

template <bool ANY>
__global__ void __launch_bounds__(128)
k_fg_arc_pre(DeviceScene sc, const SlowRay* __restrict__ rays, const unsigned int* __restrict__ lin_idx, unsigned n_lin, ArcPre* __restrict__ pre)
{
    const unsigned w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_lin) return;
    const unsigned e = __ldg(lin_idx + w);
    const float4 ro = __ldg(&rays[e].o), rd = __ldg(&rays[e].d);
    ArcPre out;
    out.n = 0; out.done = 1; out.t = 0.0; out.nudge = 0; out.n_it = 0;
    if (!(rd.x == 0.0f && rd.y == 0.0f && rd.z == 0.0f)) {
        ArcRay<ANY> R(sc, nullptr, ro, rd);
        ArcWalk W;
        W.init(sc, R.O, R.d, R.aa, R.ad, R.dd, ANY ? (double)R.tlim : (double)__int_as_float(0x7f800000));
        out.done = 0;
        while (out.n < ARC_PRE) {
            int cell; double t_in; float rmax;
            if (!W.next(cell, t_in, rmax)) { out.done = 1; break; }
            if (!R.cell_can_accept(cell, rmax)) continue;
            out.cell[out.n] = cell; out.t_in[out.n] = t_in; out.rmax[out.n] = rmax; out.n++;
        }
        out.t = W.t; out.nudge = W.nudge; out.n_it = W.n_it;
    }
    pre[w] = out;
}

// The cells k_fg_arc_pre has found, one warp per ray (lane = list entry: the gathers of a slab of 32 entries are in flight
// together; one thread per ray, walking its ~150-entry lists alone, was measured 1.4x slower - a chain of dependent gathers).
// This kernel holds no walk state, so that twice as many warps fit an SM as in k_fg_arc: the loop is bound by gather latency.
// Rays whose arc goes on after these cells, and all unanswered rays when the scene has a list of triangles too small for the
// direction index, are listed for k_fg_arc (`rest`).
// (any hit, unbounded: the cell of the ray's direction, where |P| exceeds every T, has been through k_fg_lin_first already)
#define ARC_FIRST_WARPS 4
template <bool ANY>
__global__ void __launch_bounds__(32 * ARC_FIRST_WARPS, 8)
k_fg_arc_first(DeviceScene sc, const SlowRay* __restrict__ rays, SlowRes* __restrict__ res, const unsigned int* __restrict__ lin_idx,
               unsigned n_lin, unsigned int* __restrict__ stat, const ArcPre* __restrict__ pre, unsigned int* __restrict__ rest_idx,
               unsigned int* __restrict__ rest_count)
{
    __shared__ unsigned s_q[ARC_FIRST_WARPS][64];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned w = blockIdx.x * ARC_FIRST_WARPS + wib;
    if (w >= n_lin) return;
    const unsigned e = __ldg(lin_idx + w);
    const float4 ro = __ldg(&rays[e].o), rd = __ldg(&rays[e].d);
    if (rd.x == 0.0f && rd.y == 0.0f && rd.z == 0.0f) return;                // no triangle accepts a zero direction (cpp:371)
    ArcRay<ANY> R(sc, s_q[wib], ro, rd);
    double t_end = ANY ? (double)R.tlim : (double)__int_as_float(0x7f800000);
    const int pre_n = pre[w].n;
    for (int c = 0; c < pre_n && !(ANY && R.found); c++) {
        if (!(pre[w].t_in[c] < t_end)) break;
        R.cell(pre[w].cell[c], pre[w].rmax[c]);
        if (!ANY) t_end = fmin(t_end, (double)R.tlim * 1.000001);            // (ties at equal t: the lower primitive index wins)
    }
    if (lane != 0) return;
    if (R.found) {
        if (ANY) res[e].found = 1;
        else atomicMin(&res[e].key, slow_key(R.tlim, R.plim));
    }
    const bool walk_on = !pre[w].done && pre[w].t < t_end;
    if (!(ANY && R.found) && (walk_on || sc.fg_n_wide > 0)) rest_idx[atomicAdd(rest_count, 1u)] = w;
    // (statistics from one ray in 64: atomics per ray on a few addresses cost more than the walk itself)
    if (stat && (w & 63u) == 0u) { atomicAdd(stat, 64u); atomicAdd(stat + 2, 64u * R.n_cells); atomicAdd(stat + 3, 64u * R.n_exact); atomicMax(stat + 1, pre[w].n_it); }
}

// What the first cells did not settle: the rest of the arc, one warp per ray.  The walk is serial arithmetic (~2 us per cell,
// an arc of 90 degrees is ~650 cells of a 1024-cell face), so the warp splits the arc by ANGLE into 32 stretches, one per lane;
// every lane walks its stretch and emits the cells that can accept as work items (k_fg_arc_items: one warp per (ray, cell)).
// Neighbouring stretches may both name the cell they meet in: a cell tested twice costs time, nothing else.
template <bool ANY>
__global__ void __launch_bounds__(32 * ARC_WARPS)
k_fg_arc(DeviceScene sc, const SlowRay* __restrict__ rays, SlowRes* __restrict__ res, const unsigned int* __restrict__ lin_idx,
         const unsigned int* __restrict__ rest_idx, unsigned n_rest, unsigned int* __restrict__ stat, const ArcPre* __restrict__ pre,
         ArcItem* __restrict__ items, unsigned int* __restrict__ n_items,
         unsigned item_cap, unsigned int* __restrict__ heavy_idx, unsigned int* __restrict__ heavy_count,     // heavy_*: the rays given up (-> k_far_linear)
         unsigned max_cells)
{
    __shared__ unsigned s_q[ARC_WARPS][64];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const unsigned wr = blockIdx.x * ARC_WARPS + wib;
    if (wr >= n_rest) return;
    const unsigned w = __ldg(rest_idx + wr);
    const unsigned e = __ldg(lin_idx + w);
    const float4 ro = __ldg(&rays[e].o), rd = __ldg(&rays[e].d);
    ArcRay<ANY> R(sc, s_q[wib], ro, rd);
    const float inf = __int_as_float(0x7f800000);
    const double dinf = (double)inf;
    double t_end = ANY ? (double)R.tlim : dinf;
    if (!ANY) {                                                              // what the first cells have found bounds the search
        const unsigned long long key = *reinterpret_cast<volatile unsigned long long*>(&res[e].key);
        const float kt = __uint_as_float((unsigned)(key >> 32)); const int kp = (int)(unsigned)(key & 0xffffffffull);
        if (kt < R.tlim || (kt == R.tlim && kp < R.plim)) { R.tlim = kt; R.plim = kp; t_end = fmin(t_end, (double)kt * 1.000001); }
    }
    unsigned n_emitted = 0, n_it = 0;
    bool given_up = false;
    const double t0 = pre[w].t;
    if (!pre[w].done && t0 < t_end) {
        // P(t) = A + t d seen from the centre: at angle theta(t) = atan2(t |d| s, |A| + t |d| c) from A, (c, s) = cos, sin of (A, d);
        // t(theta) = |A| sin(theta) / (|d| sin(theta_max - theta)), theta_max = angle(A, d)
        const double la = sqrt(R.aa), ld = sqrt(R.dd);
        const double c = fmax(-1.0, fmin(1.0, R.ad / (la * ld))), sn = sqrt(fmax(0.0, 1.0 - c * c));
        const double th_max = atan2(sn, c);
        const double th0 = atan2(t0 * ld * sn, la + t0 * ld * c);
        const double th1 = t_end < 1e300 ? atan2(t_end * ld * sn, la + t_end * ld * c) : th_max;
        double ta = t0, tb = t_end;                                          // this lane's stretch [ta, tb)
        if (lane > 0) { const double th = th0 + (th1 - th0) * (double)lane / 32.0; ta = la * sin(th) / (ld * sin(th_max - th)); }
        if (lane < 31) { const double th = th0 + (th1 - th0) * (double)(lane + 1) / 32.0; tb = la * sin(th) / (ld * sin(th_max - th)); }
        // a ray along A (or through the centre) has no angle to split; nor has an arc of a few cells: lane 0 walks all of it
        const bool ok = th1 - th0 > 64.0 * (2.0 / (double)sc.fg_K) && ta >= t0 && tb <= t_end && ta <= tb;
        const bool split = __all_sync(0xffffffffu, ok);
        if (!split) { ta = t0; tb = lane == 0 ? t_end : t0; }
        ArcWalk W;
        W.init(sc, R.O, R.d, R.aa, R.ad, R.dd, tb);
        W.t = ta; W.nudge = (lane == 0) ? pre[w].nudge : 0;
        bool more = ta < tb;
        while (__any_sync(0xffffffffu, more)) {
            int cell = 0; double t_in = 0.0; float rmax = 0.f;
            bool have = false;
            if (more) {
                more = W.next(cell, t_in, rmax);
                have = more && R.cell_can_accept(cell, rmax);                // (no: every T of this cell exceeds |P| on this stretch)
            }
            const unsigned m = __ballot_sync(0xffffffffu, have);
            if (!m) continue;
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(n_items, (unsigned)__popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            const unsigned slot = base + (unsigned)__popc(m & lt_mask);
            if (have && slot < item_cap) {
                ArcItem itx; itx.e = e; itx.cell = cell; itx.rmax = rmax; itx.t_in = (float)t_in * 0.99999f;
                items[slot] = itx;
            }
            n_emitted += (unsigned)__popc(m);
            // an arc that runs on and on (or a full item list): the block-wide scan of every record (k_far_linear) is complete
            if (base + (unsigned)__popc(m) > item_cap || (heavy_idx && n_emitted >= max_cells)) { given_up = true; break; }
        }
        n_it = W.n_it;
        if (given_up && lane == 0) heavy_idx[atomicAdd(heavy_count, 1u)] = e;
    }
    // the triangles too small for the direction index (fg_wide): anywhere along the ray, same filter
    if (sc.fg_n_wide > 0) R.list(sc.fg_wide, 0ull, (unsigned long long)sc.fg_n_wide, false, inf);
    if (lane == 0) {
        if (stat && (w & 63u) == 0u) { atomicAdd(stat + 2, 64u * n_emitted); atomicAdd(stat + 3, 64u * R.n_exact); atomicMax(stat + 1, n_it); }
        if (R.found) {
            if (ANY) res[e].found = 1;
            else atomicMin(&res[e].key, slow_key(R.tlim, R.plim));
        }
    }
}

// the cells of the long arcs, one warp per (ray, cell)
template <bool ANY>
__global__ void __launch_bounds__(32 * ARC_WARPS)
k_fg_arc_items(DeviceScene sc, const SlowRay* __restrict__ rays, SlowRes* __restrict__ res, const ArcItem* __restrict__ items, unsigned n_items)
{
    __shared__ unsigned s_q[ARC_WARPS][64];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned w = blockIdx.x * ARC_WARPS + wib;
    if (w >= n_items) return;
    const ArcItem itx = items[w];
    const unsigned e = itx.e;
    if (ANY && *reinterpret_cast<volatile int*>(&res[e].found)) return;
    const float4 ro = __ldg(&rays[e].o), rd = __ldg(&rays[e].d);
    ArcRay<ANY> R(sc, s_q[wib], ro, rd);
    if (!ANY) {                                                              // what has been found so far bounds the search
        const unsigned long long key = *reinterpret_cast<volatile unsigned long long*>(&res[e].key);
        const float kt = __uint_as_float((unsigned)(key >> 32)); const int kp = (int)(unsigned)(key & 0xffffffffull);
        if (kt < R.tlim || (kt == R.tlim && kp < R.plim)) { R.tlim = kt; R.plim = kp; }
        if (itx.t_in > R.tlim) return;                                       // the ray has been answered nearer than this cell
    }
    R.cell(itx.cell, itx.rmax);
    if (lane == 0 && R.found) {
        if (ANY) res[e].found = 1;
        else atomicMin(&res[e].key, slow_key(R.tlim, R.plim));
    }
}

// conservative ray / box test with the box inflated by `infl`
__device__ __forceinline__ bool slab_inflated(float lox, float hix, float loy, float hiy, float loz, float hiz, float infl,
                                              V3 O, V3 inv, float tcull, float& tnear)
{
    return slab(lox - infl, hix + infl, loy - infl, hiy + infl, loz - infl, hiz + infl, O, inv, tcull, tnear);
}

// (a ray from 10^8 units away is off by tens of units when it comes back: its inflated boxes cover half the scene.  One thread
// walking all of that would hold its warp for milliseconds: past this many exact tests the ray goes to k_slow's block-wide scan)
#define LIN_NEAR_BUDGET 192u
template <bool ANY>
__global__ void __launch_bounds__(128)
k_lin_near(DeviceScene sc, const SlowRay* __restrict__ rays, SlowRes* __restrict__ res, const unsigned int* __restrict__ lin_idx, unsigned n_lin,
           unsigned int* __restrict__ stat, unsigned int* __restrict__ heavy_idx, unsigned int* __restrict__ heavy_count)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_lin) return;
    const unsigned e = __ldg(lin_idx + i);
    if (ANY && res[e].found) return;
    const float4 ro = __ldg(&rays[e].o), rd = __ldg(&rays[e].d);
    const V3 O = mk(ro.x, ro.y, ro.z), d = mk(rd.x, rd.y, rd.z);
    HitRec best; best.t = ro.w; best.leaf = -1; best.prim = ANY ? 0x7fffffff : __float_as_int(rd.w);
    if (!ANY) {                                                            // what k_fg_arc found bounds the search
        const unsigned long long key = res[e].key;
        const float kt = __uint_as_float((unsigned)(key >> 32)); const int kp = (int)(unsigned)(key & 0xffffffffull);
        if (kt < best.t || (kt == best.t && kp < best.prim)) { best.t = kt; best.prim = kp; }
    }
    bool found = false, heavy = false;
    unsigned n_tests = 0;
    auto test = [&](int idx) {
        float t; int prim;
        n_tests++;
        if (prim_test<true>(sc.prims + idx, O, d, best.t, ANY ? 0x7fffffff : best.prim, t, prim)) {
            found = true;
            if (!ANY) { best.t = t; best.prim = prim; }
            return true;
        }
        return false;
    };
    // the primitives neither the tree nor the grid holds
    for (int k = sc.n_leaf; k < sc.n_all; k++) if (test(k) && ANY) { res[e].found = 1; return; }
    for (int k = 0; k < sc.n_always; k++) if (test(__ldg(sc.always_idx + k)) && ANY) { res[e].found = 1; return; }
    const float Olen = sqrtf(O.x * O.x + O.y * O.y + O.z * O.z);
    // spheres: |oc|^2 - r^2 and b^2 are ~R^2 in float, the discriminant b^2 - 4c (cpp:426) carries an error of ~32 u R^2: it can be
    // positive only if the ray passes within ~1.4e-3 R + r of the centre, i.e. is aimed at the scene within that angle
    if (sc.fg_n_sph > 0) {
        const V3 A = mk(O.x - sc.fg_center[0], O.y - sc.fg_center[1], O.z - sc.fg_center[2]);
        const float R = sqrtf(A.x * A.x + A.y * A.y + A.z * A.z);
        bool aimed = true;
        if (R > 1.0f && R < 1e30f) {
            const float th = 4e-3f + (3.5f * sc.extent + sc.fg_rmax) / R * 1.1f;
            if (th < 1.0f) aimed = (d.x * A.x + d.y * A.y + d.z * A.z) <= -(1.0f - 0.5f * th * th) * R + 1e-6f * R;
        }
        if (aimed)
            for (int k = 0; k < sc.fg_n_sph; k++) if (test((int)__ldg(sc.fg_sph + k)) && ANY) { res[e].found = 1; return; }
    }
    // The tree, boxes inflated by how far the reference's float P = O + d t can lie from a triangle it accepts near itself
    // when the origin is that far out (the leaf boxes are padded for in-scene origins only): off the plane by the rounding
    // of t = num / nd (cpp:367-381: |num| carries <= 3 u |O|, nd <= 3 u, the quotient u: <= 7 u |O| for t ~ |O|), sideways by
    // the rounding of d t (<= u t); 8.5 u (|O| + 2 E) covers both.
    if (sc.n_leaf > 0 && Olen < 1e30f && !(d.x == 0.0f && d.y == 0.0f && d.z == 0.0f && sc.fg_n_sph == 0)) {
        const float infl = 5.1e-7f * (Olen + 2.0f * sc.extent) + 1e-4f;
        const V3 inv = mk(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
        int stack[RT_STACK_SIZE];
        int sp = 0, node = 0;
        for (;;) {
            float4 xy0, xy1, z01; int4 kids;
            load_node(sc.nodes + node, xy0, xy1, z01, kids);
            float tn0, tn1;
            const bool h0 = slab_inflated(xy0.x, xy0.y, xy0.z, xy0.w, z01.x, z01.y, infl, O, inv, best.t, tn0);
            const bool h1 = slab_inflated(xy1.x, xy1.y, xy1.z, xy1.w, z01.z, z01.w, infl, O, inv, best.t, tn1);
            int next;
            if (h0 && h1) { if (sp < RT_STACK_SIZE) stack[sp++] = kids.y; next = kids.x; }
            else if (h0) next = kids.x;
            else if (h1) next = kids.y;
            else { if (sp == 0) break; next = stack[--sp]; }
            bool done = false;
            while (next < 0) {
                if (test(~next) && ANY) { res[e].found = 1; return; }
                if (sp == 0) { done = true; break; }
                next = stack[--sp];
            }
            if (done) break;
            if (n_tests > LIN_NEAR_BUDGET) { heavy = true; break; }
            node = next;
        }
    }
    if (heavy) heavy_idx[atomicAdd(heavy_count, 1u)] = e;       // the block-wide linear scan (k_slow) finishes this one
    if (stat && n_tests > 64u) { atomicAdd(stat, 1u); atomicAdd(stat + 1, n_tests >> 6); }
    if (found) {
        if (ANY) res[e].found = 1;
        else atomicMin(&res[e].key, slow_key(best.t, best.prim));
    }
}

// The far regime of a ray from outside the scene by a scan of every triangle's far-field constants (N, D, T): for the rays
// whose arc over the direction grid runs on and on without an acceptor (k_fg_arc gives them up after ARC_MAX_CELLS cells).
// Same block layout as k_slow (32 rays per block, the records streamed through shared memory in tiles, slices of the record
// range over blockIdx.y), but the filter is the bracket test of ArcRay::list - plane ahead, t within the limit, |P(t)| >= T
// somewhere in the bracket of t - so only a few records per ray reach the exact test.  The near regime, spheres and the
// primitives outside the tree are k_lin_near's.
template <bool ANY>
__global__ void __launch_bounds__(256)
k_far_linear(DeviceScene sc, const SlowRay* __restrict__ rays, unsigned n, SlowRes* __restrict__ res, int chunk,
             const unsigned int* __restrict__ idx)
{
    __shared__ float4 s_A[SLOW_TILE];
    __shared__ float s_T[SLOW_TILE];
    __shared__ float4 s_O[8][SLOW_RPW], s_D[8][SLOW_RPW];
    __shared__ unsigned long long s_key[8][SLOW_RPW];
    __shared__ int s_found[8][SLOW_RPW];
    __shared__ unsigned long long s_q[8][64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const unsigned first = blockIdx.x * SLOW_RPB + warp;
    bool my_live = false;
    unsigned long long key0 = 0ull;
    if (lane < SLOW_RPW) {
        const unsigned e0 = first + lane * 8u;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f), d = o;
        if (e0 < n) {
            const unsigned e = __ldg(idx + e0);
            o = __ldg(&rays[e].o); d = __ldg(&rays[e].d);
            my_live = !(d.x == 0.f && d.y == 0.f && d.z == 0.f);
            if (ANY) { if (*reinterpret_cast<volatile int*>(&res[e].found)) my_live = false; }
            else {
                const unsigned long long k = *reinterpret_cast<volatile unsigned long long*>(&res[e].key);     // what has been found so far bounds the search
                const unsigned long long kr = ((unsigned long long)__float_as_uint(o.w) << 32) | (unsigned)__float_as_int(d.w);
                if (k < kr) { o.w = __uint_as_float((unsigned)(k >> 32)); d.w = __int_as_float((int)(unsigned)(k & 0xffffffffull)); }
            }
        }
        s_O[warp][lane] = o; s_D[warp][lane] = d;
        key0 = ((unsigned long long)__float_as_uint(o.w) << 32) | (unsigned)__float_as_int(d.w);
        s_key[warp][lane] = key0;
        s_found[warp][lane] = 0;
    }
    unsigned live_mask = __ballot_sync(0xffffffffu, my_live) & ((1u << SLOW_RPW) - 1u);
    __syncwarp();
    float4 dj[SLOW_RPW], oj[SLOW_RPW];
    float aa[SLOW_RPW], ad[SLOW_RPW], dd[SLOW_RPW], dl[SLOW_RPW];
#pragma unroll
    for (int j = 0; j < SLOW_RPW; j++) {
        dj[j] = s_D[warp][j]; oj[j] = s_O[warp][j];
        const float ax = oj[j].x - sc.fg_center[0], ay = oj[j].y - sc.fg_center[1], az = oj[j].z - sc.fg_center[2];
        aa[j] = ax * ax + ay * ay + az * az; ad[j] = ax * dj[j].x + ay * dj[j].y + az * dj[j].z;
        dd[j] = dj[j].x * dj[j].x + dj[j].y * dj[j].y + dj[j].z * dj[j].z;
        dl[j] = 1e-6f * (fabsf(oj[j].x) + fabsf(oj[j].y) + fabsf(oj[j].z) + sc.extent);
    }
    unsigned q_len = 0;
    const int nl = min(sc.n_all, ((int)blockIdx.y + 1) * chunk);
    for (int base = (int)blockIdx.y * chunk; base < nl; base += SLOW_TILE) {
        {
            const int i = base + (int)threadIdx.x;
            s_A[threadIdx.x] = (i < nl) ? __ldg(sc.fg_A + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            s_T[threadIdx.x] = (i < nl) ? __ldg(sc.fg_B + i).x : -1.0f;        // (< 0: no far field - spheres)
        }
        __syncthreads();
        if (ANY) {
#pragma unroll
            for (int j = 0; j < SLOW_RPW; j++)
                if (((live_mask >> j) & 1u) && *reinterpret_cast<volatile int*>(&s_found[warp][j])) live_mask &= ~(1u << j);
        }
#pragma unroll 2
        for (int k = 0; k < SLOW_TILE / 32; k++) {
            const int sl = k * 32 + lane, i = base + sl;
            const float4 fa = s_A[sl];
            const float T = s_T[sl];
#pragma unroll
            for (int j = 0; j < SLOW_RPW; j++) {
                if (!((live_mask >> j) & 1u)) continue;
                const float nd = __fmaf_rn(fa.x, dj[j].x, __fmaf_rn(fa.y, dj[j].y, fa.z * dj[j].z));
                const float no = __fmaf_rn(fa.x, oj[j].x, __fmaf_rn(fa.y, oj[j].y, __fmaf_rn(fa.z, oj[j].z, fa.w)));
                const float and_ = fabsf(nd), ano = fabsf(no);
                bool pass = T > 0.f && and_ > FG_ND_MIN && !(ano > dl[j] && ((no < 0.f) == (nd < 0.f)));
                if (pass) {
                    const float tlim = __uint_as_float((unsigned)(*reinterpret_cast<volatile unsigned long long*>(&s_key[warp][j]) >> 32));
                    const float t_lo = fmaxf(ano - dl[j], 0.f) / (and_ + FG_ND_SLACK) * 0.999999f;
                    pass = t_lo <= tlim;
                    if (pass && and_ > 2.0f * FG_ND_SLACK) {
                        // |P(t)|^2 = aa + 2 t ad + t^2 dd in float: every term carries <= 3e-7 of its magnitude
                        const float t_hi = (ano + dl[j]) / (and_ - FG_ND_SLACK) * 1.000001f;
                        const float e_lo = 4e-7f * (aa[j] + 2.0f * t_lo * fabsf(ad[j]) + t_lo * t_lo * dd[j]);
                        const float e_hi = 4e-7f * (aa[j] + 2.0f * t_hi * fabsf(ad[j]) + t_hi * t_hi * dd[j]);
                        const float p_lo = aa[j] + t_lo * (2.0f * ad[j] + t_lo * dd[j]) + e_lo, p_hi = aa[j] + t_hi * (2.0f * ad[j] + t_hi * dd[j]) + e_hi;
                        const float need = fmaxf(T * 0.9999f - 2e-6f * (fabsf(oj[j].x) + fabsf(oj[j].y) + fabsf(oj[j].z)), 0.f);
                        pass = !(fmaxf(p_lo, p_hi) < need * need);           // (an overflow to inf or a NaN passes)
                    }
                }
                const unsigned mask = __ballot_sync(0xffffffffu, pass);
                if (mask == 0u) continue;
                if (pass) s_q[warp][q_len + (unsigned)__popc(mask & lt_mask)] = ((unsigned long long)j << 32) | (unsigned)i;
                q_len += (unsigned)__popc(mask);
                if (q_len >= 32u) {
                    __syncwarp();
                    slow_exact<ANY>(sc, s_q[warp][lane], s_O[warp], s_D[warp], s_key[warp], s_found[warp]);
                    __syncwarp();
                    const unsigned long long tail = (lane + 32u < q_len) ? s_q[warp][lane + 32] : 0ull;
                    __syncwarp();
                    if (lane + 32u < q_len) s_q[warp][lane] = tail;
                    q_len -= 32u;
                    __syncwarp();
                }
            }
        }
        if (!__syncthreads_or(live_mask ? 1 : 0)) break;
    }
    __syncwarp();
    if ((unsigned)lane < q_len) slow_exact<ANY>(sc, s_q[warp][lane], s_O[warp], s_D[warp], s_key[warp], s_found[warp]);
    __syncwarp();
    if (lane < SLOW_RPW && first + lane * 8u < n) {
        const unsigned e = __ldg(idx + first + lane * 8u);
        if (ANY) { if (s_found[warp][lane]) res[e].found = 1; }
        else { const unsigned long long k = s_key[warp][lane]; if (k < key0) atomicMin(&res[e].key, k); }
    }
}

template <int MODE>
__device__ __forceinline__ const PrimRec* stage_prims(const DeviceScene& sc, PrimRec* smem) {
    if (MODE == 1) {
        const float4* src = reinterpret_cast<const float4*>(sc.prims);
        float4* dst = reinterpret_cast<float4*>(smem);
        for (int i = threadIdx.x; i < sc.n_all * 4; i += blockDim.x) dst[i] = src[i];
        __syncthreads();
    }
    return smem;
}

// Fill the node of a closest hit: what IntersectTriangle / IntersectSphere leave in
// RaycastHitInfo (cpp:399-407, cpp:456-462) for the winning primitive.
__device__ __forceinline__ void fill_node(const PrimRec* __restrict__ prims, const HitRec& h, V3 O, V3 d,
                                          int parent, int pixel, unsigned flags, Node& nd) {
    const PrimRec* p = prims + h.leaf;
    const float4 ra = p->a, rb = p->b, rc = p->c, rd = p->d;
    const V3 P = O + d * h.t;                                     // cpp:387 / cpp:456
    V3 Nn; float al = 0.f, be = 0.f, ga = 0.f;
    if (__float_as_int(rd.w) & RT_PRIM_SPHERE) {
        Nn = normalize(P - mk(ra.x, ra.y, ra.z));                 // cpp:459-460
        flags |= NF_SPHERE;
    } else {
        const V3 N = mk(rd.x, rd.y, rd.z);
        const V3 v0 = mk(ra.x, ra.y, ra.z), v1 = mk(rb.x, rb.y, rb.z), v2 = mk(rc.x, rc.y, rc.z);
        al = (0.5f * dot(cross(v1 - P, v2 - P), N)) / rb.w;       // cpp:392
        be = (0.5f * dot(cross(P - v0, v2 - v0), N)) / rb.w;      // cpp:393
        ga = (0.5f * dot(cross(v1 - v0, P - v0), N)) / rb.w;      // cpp:394
        Nn = normalize(N);                                        // cpp:402-403 (normalised twice, Q25)
    }
    nd.P = make_float4(P.x, P.y, P.z, __int_as_float(h.prim));
    nd.N = make_float4(Nn.x, Nn.y, Nn.z, __int_as_float(parent));
    nd.D = make_float4(d.x, d.y, d.z, __int_as_float(pixel));
    nd.B = make_float4(al, be, ga, __uint_as_float(flags));
}

// ---------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------
// What cpp:30-32 / cpp:925 do with the answer of a closest-hit ray: a hit becomes a node of the
// next level, a miss is the background colour (in the frame for a primary ray, in the parent's
// child slot otherwise).  `slot` comes from warp_alloc (all lanes).
// The wavefront path spawns the children of a hit node (cpp:87-112) in the kernel that creates the node: its
// material decides whether a reflection / a refraction ray exists, the rays go to the NEXT level's queue.
struct Spawn { QRay* queue; unsigned int* count; };      // queue == nullptr: the caller spawns elsewhere (k_shade)
// slots in the next queue; called by all 32 lanes
__device__ __forceinline__ void spawn_slots(const DeviceScene& sc, const Spawn& so, bool hit, unsigned flags, int prim,
                                            bool& want_refl, bool& want_refr, unsigned& s0, unsigned& s1) {
    want_refl = want_refr = false; s0 = s1 = 0u;
    if (!so.queue) return;
    if (hit && ((flags >> NF_BOUNCE_SHIFT) & 0xffu) > 0u) {                     // cpp:87
        const Material M = load_material(sc.materials, __ldg(sc.prim_material + prim));
        want_refl = M.Ks > 0; want_refr = M.Kt > 0;                             // cpp:94, 108
    }
    s0 = warp_alloc(so.count, want_refl);
    s1 = warp_alloc(so.count, want_refr);
}
__device__ __forceinline__ void spawn_rays(const Spawn& so, bool want_refl, bool want_refr, unsigned s0, unsigned s1,
                                           const float4& nP, const float4& nN, V3 D, unsigned flags, unsigned node) {
    const V3 P = mk(nP.x, nP.y, nP.z), N = mk(nN.x, nN.y, nN.z);
    const unsigned child_flags = (((flags >> NF_BOUNCE_SHIFT) & 0xffu) - 1u) << NF_BOUNCE_SHIFT;
    if (want_refl) {
        const V3 rdir = normalize(reflect(D, N));                             // cpp:96-97
        const V3 ro = P + rdir * RT_SHADOW_OFFSET;                            // cpp:98
        const V3 rd = normalize(rdir);                                        // Ray ctor
        QRay q; q.o = make_float4(ro.x, ro.y, ro.z, __int_as_float((int)node));
        q.d = make_float4(rd.x, rd.y, rd.z, __uint_as_float(child_flags));
        so.queue[s0] = q;
    }
    if (want_refr) {
        const V3 tdir = calculate_refraction(D, N, RT_IOR);                   // cpp:109
        const V3 to = P + tdir * RT_SHADOW_OFFSET;                            // cpp:110
        const V3 td = normalize(tdir);                                        // Ray ctor (zero stays zero, Q20)
        QRay q; q.o = make_float4(to.x, to.y, to.z, __int_as_float((int)node));
        q.d = make_float4(td.x, td.y, td.z, __uint_as_float(child_flags | NF_REFR));
        so.queue[s1] = q;
    }
}

template <bool PRIMARY>
__device__ __forceinline__ void commit_closest(const PrimRec* __restrict__ prims, bool hit, const HitRec& h, V3 O, V3 d,
                                               int parent, int pixel, unsigned flags, unsigned slot, unsigned node_cap,
                                               Node* __restrict__ nodes, NodeAux* __restrict__ aux,
                                               uint32_t* __restrict__ pix_hits, int16_t* __restrict__ fb, Node& nd)
{
    if (hit) {
        if (slot >= node_cap) return;   // cannot happen: capacity is ensured before the launch
        if (!PRIMARY) pixel = __float_as_int(nodes[parent].D.w);
        fill_node(prims, h, O, d, parent, pixel, flags, nd);
        nodes[slot] = nd;
        uint2* a = reinterpret_cast<uint2*>(aux + slot);          // [0] local (the shading kernel's), [1] refl, [2] refr, [3] subtree sizes
        a[1] = make_uint2(0u, 0u); a[2] = make_uint2(0u, 0u); a[3] = make_uint2(0u, 0u);   // Pixel() (cpp:91-92)
        if (!PRIMARY) {
            short* s = (flags & NF_REFR) ? aux[parent].refr : aux[parent].refl;
            s[3] = 2;
        }
    } else {
        if (PRIMARY) {
            fb[3 * (size_t)pixel + 0] = 254; fb[3 * (size_t)pixel + 1] = 64; fb[3 * (size_t)pixel + 2] = 205;   // BG_COLOR h:597
            pix_hits[pixel] = 0u;
        } else {
            short* s = (flags & NF_REFR) ? aux[parent].refr : aux[parent].refl;
            s[0] = 254; s[1] = 64; s[2] = 205; s[3] = 1;     // cpp:30-32 at depth > 0
        }
    }
}

// Closest-hit wavefront step.  PRIMARY: one thread per local pixel, ray from GenerateRay.
// Otherwise one thread per queued reflection / refraction ray.
template <int MODE, bool PRIMARY>
__global__ void __launch_bounds__(128)
k_trace(DeviceScene sc, FrameParams fp, const QRay* __restrict__ queue, unsigned n_items,
        const unsigned int* __restrict__ n_items_dev, Node* __restrict__ nodes, NodeAux* __restrict__ aux,
        unsigned int* __restrict__ counters, uint32_t* __restrict__ pix_hits, int16_t* __restrict__ fb, unsigned node_cap, SlowQ sq,
        Spawn spawn)
{
    __shared__ PrimRec s_prims[MODE == 1 ? RT_SMEM_PRIMS : 1];
    const PrimRec* sp = stage_prims<MODE>(sc, s_prims);
    // the queue length is still on the device when the launch is issued (no host round trip between
    // k_shade and this kernel): the grid covers the upper bound n_items, the count trims it
    if (n_items_dev) n_items = min(n_items, __ldg(n_items_dev));
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < n_items;
    V3 O = mk(0, 0, 0), d = mk(0, 0, 0);
    int parent = -1, pixel = 0; unsigned flags = 0;
    if (active) {
        if (PRIMARY) {
            pixel = (int)i;
            const int x = pixel % fp.W, row = pixel / fp.W;
            const int y = fp.row_first + row * fp.row_step;
            // cpp:846-852: direction(NDCX, NDCY, -1) through the inverse view matrix, normalised
            const float dx = __ldg(fp.ndc_x + x), dy = __ldg(fp.ndc_y + y), dz = -1.0f;
            const V3 w = mk(fp.inv[0] * dx + fp.inv[1] * dy + fp.inv[2] * dz,
                            fp.inv[3] * dx + fp.inv[4] * dy + fp.inv[5] * dz,
                            fp.inv[6] * dx + fp.inv[7] * dy + fp.inv[8] * dz);      // h:227-232
            d = normalize(w);
            O = mk(fp.cam[0], fp.cam[1], fp.cam[2]);                                // cpp:843
            flags = (unsigned)fp.depth << NF_BOUNCE_SHIFT;
        } else {
            const float4 qo = queue[i].o, qd = queue[i].d;
            O = mk(qo.x, qo.y, qo.z); d = mk(qd.x, qd.y, qd.z);
            parent = __float_as_int(qo.w); flags = __float_as_uint(qd.w);
        }
    }
    HitRec h;
    const int tr = trace_ray<MODE, false>(sc, sp, active, O, d, 0.f, h, sq, (int)i, 0);
    const unsigned slot = warp_alloc(&counters[0], tr == TR_HIT);
    bool wr, wt; unsigned s0, s1;
    spawn_slots(sc, spawn, active && tr == TR_HIT, flags, h.prim, wr, wt, s0, s1);
    if (!active || tr == TR_PENDING) return;
    Node nd;
    commit_closest<PRIMARY>(MODE == 1 ? sp : sc.prims, tr == TR_HIT, h, O, d, parent, pixel, flags, slot, node_cap, nodes, aux,
                            pix_hits, fb, nd);
    if (wr || wt) spawn_rays(spawn, wr, wt, s0, s1, nd.P, nd.N, d, flags, slot);
}

// second half of k_trace for the rays that went through the deferred slow path
template <bool PRIMARY>
__global__ void __launch_bounds__(128)
k_trace_finish(DeviceScene sc, FrameParams fp, const QRay* __restrict__ queue, const SlowRay* __restrict__ rays,
               const SlowRes* __restrict__ res, unsigned n_slow, Node* __restrict__ nodes, NodeAux* __restrict__ aux,
               unsigned int* __restrict__ counters, uint32_t* __restrict__ pix_hits, int16_t* __restrict__ fb, unsigned node_cap,
               Spawn spawn)
{
    const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = e < n_slow;
    bool hit = false; HitRec h; h.t = 0.f; h.leaf = -1; h.prim = 0;
    V3 O = mk(0, 0, 0), d = mk(0, 0, 0);
    int parent = -1, pixel = 0; unsigned flags = 0;
    if (active) {
        const SlowRay r = rays[e];
        const SlowRes a = res[e];
        O = mk(r.o.x, r.o.y, r.o.z); d = mk(r.d.x, r.d.y, r.d.z);
        h.t = __uint_as_float((unsigned)(a.key >> 32)); h.prim = (int)(unsigned)(a.key & 0xffffffffull);
        hit = h.prim != 0x7fffffff;
        h.leaf = hit ? __ldg(sc.leaf_of_prim + h.prim) : -1;
        const int i = r.c.y;
        if (PRIMARY) { pixel = i; flags = (unsigned)fp.depth << NF_BOUNCE_SHIFT; }
        else { parent = __float_as_int(queue[i].o.w); flags = __float_as_uint(queue[i].d.w); }
    }
    const unsigned slot = warp_alloc(&counters[0], hit);
    bool wr, wt; unsigned s0, s1;
    spawn_slots(sc, spawn, hit, flags, h.prim, wr, wt, s0, s1);
    if (!active) return;
    Node nd;
    commit_closest<PRIMARY>(sc.prims, hit, h, O, d, parent, pixel, flags, slot, node_cap, nodes, aux, pix_hits, fb, nd);
    if (wr || wt) spawn_rays(spawn, wr, wt, s0, s1, nd.P, nd.N, d, flags, slot);
}

// ---- closest hit of the queued reflection / refraction rays, wavefront form ----------------
// k_trace<secondary> maps one thread to one ray: the rays of a warp scatter over the scene and end
// after very different numbers of node visits (ncu: 13-17 of 32 lanes active, the leaf tests at 9).
// As in k_anyhit the traversal becomes a persistent kernel whose lanes fetch the next ray when theirs
// is done and which runs inner nodes and leaf tests as two phases; it only walks the tree and leaves
// (t, leaf, prim) per ray.  k_commit (one thread per ray, convergent) then adds the large-primitive
// list, the sliver list and the far field, and creates the nodes exactly as k_trace does.
struct __align__(16) CHit { float t; int leaf; int prim; int pad; };

template <bool COUNT>
__global__ void __launch_bounds__(128, AH_MIN_BLOCKS)
k_closest(DeviceScene sc, const QRay* __restrict__ queue, const unsigned int* __restrict__ n_ptr, unsigned n_bound,
          unsigned int* __restrict__ next_ray, CHit* __restrict__ out, int ah_steps, int ah_min_search, int batch_div,
          unsigned long long* __restrict__ visit_counts)
{
    unsigned cnt_nodes = 0, cnt_leaves = 0;
    const unsigned n = min(__ldg(n_ptr), n_bound);
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    unsigned batch = n / (gridDim.x * (blockDim.x >> 5) * (unsigned)batch_div);
    batch = batch > (unsigned)AH_BATCH ? (unsigned)AH_BATCH : (batch < 32u ? 32u : (batch & ~31u));
    bool active = false;
    V3 O = mk(0, 0, 0), d = mk(0, 0, 0), inv = mk(0, 0, 0);
    float best_t = 0.f; int best_leaf = -1, best_prim = 0x7fffffff; unsigned idx = 0;
    int stack[RT_STACK_SIZE];
    int sp = 0, cur = AH_NONE;
    bool last_batch = false;
    unsigned wnext = 0, wend = 0;
    for (;;) {
        const bool exhausted = last_batch && wnext == wend;
        if (!exhausted) {
            const unsigned idle = __ballot_sync(0xffffffffu, !active);
            if (idle) {
                const unsigned want = (unsigned)__popc(idle);
                if (wend - wnext < want) {
                    if (wnext == wend && !last_batch) {
                        unsigned b = 0;
                        if (lane == 0) b = atomicAdd(next_ray, batch);
                        wnext = __shfl_sync(0xffffffffu, b, 0);
                        wend = wnext + batch;
                        if (wend >= n) { last_batch = true; if (wend > n) wend = n; if (wnext > n) wnext = n; }
                    }
                }
                const unsigned base = wnext;
                const unsigned give = min(want, wend - wnext);
                wnext += give;
                if (!active) {
                    const unsigned rank = (unsigned)__popc(idle & lt_mask);
                    if (rank < give) {
                        idx = base + rank;
                        const float4 qo = __ldg(&queue[idx].o), qd = __ldg(&queue[idx].d);
                        O = mk(qo.x, qo.y, qo.z); d = mk(qd.x, qd.y, qd.z);
                        best_t = __int_as_float(0x7f800000); best_leaf = -1; best_prim = 0x7fffffff;
                        if (ray_has_nan(O, d)) {
                            CHit h; h.t = __int_as_float(0x7fc00000); h.leaf = sc.nan_leaf; h.prim = sc.nan_leaf >= 0 ? sc.nan_prim : best_prim; h.pad = 0;
                            out[idx] = h;                 // the first triangle of the scene at t = NaN (device_scene.h), or nothing
                        } else if (sc.farfield && !in_scene(sc, O)) {
                            CHit h; h.t = best_t; h.leaf = -2; h.prim = best_prim; h.pad = 0;   // not traversed: k_commit takes the linear loop
                            out[idx] = h;
                        } else if (sc.n_leaf > 0) {
                            inv = mk(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
                            cur = 0; sp = 0; active = true;
                        } else {
                            CHit h; h.t = best_t; h.leaf = -1; h.prim = best_prim; h.pad = 0;
                            out[idx] = h;
                        }
                    }
                }
            }
        }
        if (!__any_sync(0xffffffffu, active)) { if (last_batch && wnext == wend) break; continue; }
        // phase 1: inner nodes
#pragma unroll 1
        for (int it = 0; it < ah_steps; it++) {
            const bool searching = active && cur >= 0 && cur != AH_NONE;
            const int n_search = __popc(__ballot_sync(0xffffffffu, searching));
            const bool leaf_work = __any_sync(0xffffffffu, active && !searching);
            if (n_search == 0 || (n_search < ah_min_search && leaf_work)) break;
            if (searching) {
                if (COUNT) cnt_nodes++;
                float4 xy0, xy1, z01; int4 kids;
                load_node(sc.nodes + cur, xy0, xy1, z01, kids);
                float tn0, tn1;
                const bool h0 = slab(xy0.x, xy0.y, xy0.z, xy0.w, z01.x, z01.y, O, inv, best_t, tn0);
                const bool h1 = slab(xy1.x, xy1.y, xy1.z, xy1.w, z01.z, z01.w, O, inv, best_t, tn1);
                if (h0 && h1) {
                    int nearc = kids.x, farc = kids.y;
                    if (tn1 < tn0) { nearc = kids.y; farc = kids.x; }
                    if (sp < RT_STACK_SIZE) stack[sp++] = farc;
                    cur = nearc;
                } else if (h0) cur = kids.x;
                else if (h1) cur = kids.y;
                else if (sp > 0) cur = stack[--sp];
                else cur = AH_NONE;
            }
        }
        // phase 2: one leaf per lane, or the end of the ray
        if (active) {
            if (cur != AH_NONE && cur < 0) {
                if (COUNT) cnt_leaves++;
                float t; int prim;
                if (prim_test<true>(sc.prims + (~cur), O, d, best_t, best_prim, t, prim)) { best_t = t; best_leaf = ~cur; best_prim = prim; }
                cur = (sp > 0) ? stack[--sp] : AH_NONE;
            }
            if (cur == AH_NONE) {
                CHit h; h.t = best_t; h.leaf = best_leaf; h.prim = best_prim; h.pad = 0;
                out[idx] = h;
                active = false;
            }
        }
    }
    if (COUNT) {
        unsigned long long a = cnt_nodes, b = cnt_leaves;
        for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
        if (lane == 0) { atomicAdd(visit_counts, a); atomicAdd(visit_counts + 1, b); }
    }
}

// one thread per queued ray: the rest of the closest-hit search and what cpp:30-32 does with the answer
__global__ void __launch_bounds__(128)
k_commit(DeviceScene sc, const QRay* __restrict__ queue, unsigned n_items, const unsigned int* __restrict__ n_items_dev,
         const CHit* __restrict__ chits, Node* __restrict__ nodes, NodeAux* __restrict__ aux, unsigned int* __restrict__ counters,
         uint32_t* __restrict__ pix_hits, int16_t* __restrict__ fb, unsigned node_cap, SlowQ sq, Spawn spawn)
{
    if (n_items_dev) n_items = min(n_items, __ldg(n_items_dev));
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < n_items;
    V3 O = mk(0, 0, 0), d = mk(0, 0, 0);
    int parent = -1; unsigned flags = 0;
    HitRec h; h.t = __int_as_float(0x7f800000); h.leaf = -1; h.prim = 0x7fffffff;
    if (active) {
        const float4 qo = queue[i].o, qd = queue[i].d;
        O = mk(qo.x, qo.y, qo.z); d = mk(qd.x, qd.y, qd.z);
        parent = __float_as_int(qo.w); flags = __float_as_uint(qd.w);
        const CHit c = chits[i];
        h.t = c.t; h.leaf = c.leaf; h.prim = c.prim;
    }
    const int tr = finish_closest(sc, active, O, d, h, sq, (int)i);
    const unsigned slot = warp_alloc(&counters[0], tr == TR_HIT);
    bool wr, wt; unsigned s0, s1;
    spawn_slots(sc, spawn, active && tr == TR_HIT, flags, h.prim, wr, wt, s0, s1);
    if (!active || tr == TR_PENDING) return;
    Node nd;
    commit_closest<false>(sc.prims, tr == TR_HIT, h, O, d, parent, 0, flags, slot, node_cap, nodes, aux, pix_hits, fb, nd);
    if (wr || wt) spawn_rays(spawn, wr, wt, s0, s1, nd.P, nd.N, d, flags, slot);
}

// shading normal of a node (cpp:225-236): interpolated object-space vertex normals for a triangle
// (InterpolateVector3 cpp:333-338, normalised once there), the geometric hit normal for a sphere
__device__ __forceinline__ V3 shading_normal(const DeviceScene& sc, const Node& nd, unsigned flags, int prim) {
    if (flags & NF_SPHERE) return mk(nd.N.x, nd.N.y, nd.N.z);
    const float4 a = __ldg(sc.vn + 3 * (size_t)prim), b = __ldg(sc.vn + 3 * (size_t)prim + 1),
                 c = __ldg(sc.vn + 3 * (size_t)prim + 2);
    return normalize((mk(a.x, a.y, a.z) * nd.B.x + mk(b.x, b.y, b.z) * nd.B.y) + mk(c.x, c.y, c.z) * nd.B.z);   // cpp:334-336
}

// the shadow ray of cpp:55-71 for one light: origin, direction (normalised twice, Ray ctor), and the
// distance beyond which a hit no longer shadows (point light: |light - P|, cpp:71-75; else unbounded)
__device__ __forceinline__ void shadow_ray(const Light& L, V3 P, V3& so, V3& sd, float& tmax) {
    V3 lightDir = mk(0, 0, 0);
    if (L.type == RT580_LIGHT_DIRECTIONAL) lightDir = -L.direction;          // cpp:56-60
    else if (L.type == RT580_LIGHT_POINT) lightDir = L.position - P;         // cpp:62-64
    lightDir = normalize(lightDir);                                          // cpp:65
    so = P + lightDir * RT_SHADOW_OFFSET;                                    // cpp:67
    sd = normalize(lightDir);                                                // Ray ctor h:431-433
    const float distToLight = length(L.position - P);                        // cpp:71
    tmax = (L.type == RT580_LIGHT_POINT) ? distToLight : __int_as_float(0x7f800000);
}

// Per hit node of one level, one thread per node: direct lighting with inline shadow rays, then spawn the children.
// The form the linear modes use (tiny scenes staged in shared memory, the brute-force checker); the wavefront path over
// the LBVH splits this work over k_shade_gen / k_anyhit / k_shade_local and the node-creating kernels (below).
template <int MODE>
__global__ void __launch_bounds__(128)
k_shade(DeviceScene sc, FrameParams fp, unsigned n0, unsigned n1, const Node* __restrict__ nodes,
        NodeAux* __restrict__ aux, QRay* __restrict__ queue, unsigned int* __restrict__ counters, SlowQ sq)
{
    __shared__ PrimRec s_prims[MODE == 1 ? RT_SMEM_PRIMS : 1];
    const PrimRec* sp = stage_prims<MODE>(sc, s_prims);
    const unsigned i = n0 + blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < n1;
    bool want_refl = false, want_refr = false;
    V3 P = mk(0, 0, 0), N = mk(0, 0, 0), D = mk(0, 0, 0), sn = mk(0, 0, 0);
    unsigned flags = 0; int bounces = 0;
    Material M{};
    if (active) {
        const Node nd = nodes[i];
        P = mk(nd.P.x, nd.P.y, nd.P.z); N = mk(nd.N.x, nd.N.y, nd.N.z); D = mk(nd.D.x, nd.D.y, nd.D.z);
        flags = __float_as_uint(nd.B.w);
        bounces = (int)((flags >> NF_BOUNCE_SHIFT) & 0xffu);
        const int prim = __float_as_int(nd.P.w);
        M = load_material(sc.materials, __ldg(sc.prim_material + prim));
        sn = shading_normal(sc, nd, flags, prim);
    }
    Pix local = mkpix(0, 0, 0);                                       // SHADOW_COLOR h:598
    const V3 cam = mk(fp.cam[0], fp.cam[1], fp.cam[2]);
    const PhongFrame pf = phong_frame(P, sn, cam);
    // the light loop is uniform over the warp (trace_ray is warp-collective)
    int j = 0;                                                        // index among the non-ambient lights
    for (int li = 0; li < sc.n_lights; li++) {                        // cpp:39
        const Light L = load_light(sc.light_type, sc.light_f, li);
        if (L.type == RT580_LIGHT_AMBIENT) continue;                  // handled by k_ao / k_resolve
        V3 so, sd; float tmax;
        shadow_ray(L, P, so, sd, tmax);
        HitRec sh;
        // cpp:75: lit unless something is hit (point light: at distance <= distToLight)
        const int tr = trace_ray<MODE, true>(sc, sp, active, so, sd, tmax, sh, sq, (int)i, li);
        // TR_PENDING: k_shade_finish adds this light's term once the deferred ray is answered
        if (active && tr == TR_MISS) local = pix_add(local, calculate_local_color(P, pf, L, M));   // cpp:77
        j++;
    }
    if (active) {
        NodeAux a;
        a.local[0] = local.r; a.local[1] = local.g; a.local[2] = local.b; a.local[3] = 0;
        a.refl[0] = a.refl[1] = a.refl[2] = a.refl[3] = 0;            // Pixel() (cpp:91-92)
        a.refr[0] = a.refr[1] = a.refr[2] = a.refr[3] = 0;
        a.sub_refl = 0; a.sub_refr = 0;
        aux[i] = a;
        if (bounces > 0) { want_refl = M.Ks > 0; want_refr = M.Kt > 0; }     // cpp:87, 94, 108
    }
    // reflection first, refraction second: queue order is irrelevant to the result (ordinals
    // come from the tree), compaction only keeps the next trace launch dense
    const unsigned s0 = warp_alloc(&counters[1], want_refl);
    const unsigned s1 = warp_alloc(&counters[1], want_refr);
    if (!active) return;
    const unsigned child_flags = (unsigned)(bounces - 1) << NF_BOUNCE_SHIFT;
    if (want_refl) {
        const V3 rdir = normalize(reflect(D, N));                             // cpp:96-97
        const V3 ro = P + rdir * RT_SHADOW_OFFSET;                            // cpp:98
        const V3 rd = normalize(rdir);                                        // Ray ctor
        QRay q; q.o = make_float4(ro.x, ro.y, ro.z, __int_as_float((int)i));
        q.d = make_float4(rd.x, rd.y, rd.z, __uint_as_float(child_flags));
        queue[s0] = q;
    }
    if (want_refr) {
        const V3 tdir = calculate_refraction(D, N, RT_IOR);                   // cpp:109
        const V3 to = P + tdir * RT_SHADOW_OFFSET;                            // cpp:110
        const V3 td = normalize(tdir);                                        // Ray ctor (zero stays zero, Q20)
        QRay q; q.o = make_float4(to.x, to.y, to.z, __int_as_float((int)i));
        q.d = make_float4(td.x, td.y, td.z, __uint_as_float(child_flags | NF_REFR));
        queue[s1] = q;
    }
}

// k_shade in two halves for the wavefront path, so that they can run on different streams: spawning the
// reflection / refraction rays needs nothing from the shadow rays, and the chain spawn -> closest hit ->
// commit -> next level is the critical path of the structure pass, while the shadow rays of a level and the
// Phong terms they gate (the bulk of the work) only have to be done before the resolve pass.
//   spawn_slots / spawn_rays  in the kernel that creates a hit node: NodeAux except `local`, children into the next
//                             level's ray queue                                              (cpp:87-112)
//   k_shade_local             per hit node: sum of the unoccluded lights' Phong terms -> NodeAux::local (cpp:53-81)
// The two write disjoint bytes of NodeAux.
__global__ void __launch_bounds__(128)
k_shade_local(DeviceScene sc, FrameParams fp, unsigned n0, unsigned n1, const Node* __restrict__ nodes,
              NodeAux* __restrict__ aux, const uint32_t* __restrict__ occl)
{
    const unsigned i = n0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n1) return;
    const Node nd = nodes[i];
    const V3 P = mk(nd.P.x, nd.P.y, nd.P.z);
    const unsigned flags = __float_as_uint(nd.B.w);
    const int prim = __float_as_int(nd.P.w);
    const Material M = load_material(sc.materials, __ldg(sc.prim_material + prim));
    const PhongFrame pf = phong_frame(P, shading_normal(sc, nd, flags, prim), mk(fp.cam[0], fp.cam[1], fp.cam[2]));
    Pix local = mkpix(0, 0, 0);                                       // SHADOW_COLOR h:598
    int j = 0;                                                        // index among the non-ambient lights
    for (int li = 0; li < sc.n_lights; li++) {                        // cpp:39
        if (__ldg(sc.light_type + li) == RT580_LIGHT_AMBIENT) continue;          // handled by the occlusion pass / k_resolve
        // occluders of this ray: tree (k_anyhit), large primitives (k_shade_gen), or still pending (k_shadow_finish)
        if (__ldg(occl + (size_t)(i - n0) * sc.n_nonambient + j) == 0u)
            local = pix_add(local, calculate_local_color(P, pf, load_light(sc.light_type, sc.light_f, li), M));   // cpp:77
        j++;
    }
    *reinterpret_cast<uint2*>(aux + i) = make_uint2((unsigned)(unsigned short)local.r | ((unsigned)(unsigned short)local.g << 16),
                                                   (unsigned)(unsigned short)local.b);
}

// deferred shadow rays of k_shade: an unoccluded one adds its light's Phong term (cpp:77).  Pixel
// addition is plain integer addition (h:403-409), so the order of the terms does not matter; two
// lights of one node may finish concurrently, hence the packed 16-bit atomics (terms are in
// [0,255], sums stay far below 2^16).
__global__ void __launch_bounds__(128)
k_shade_finish(DeviceScene sc, FrameParams fp, const SlowRay* __restrict__ rays, const SlowRes* __restrict__ res,
               unsigned n_slow, const Node* __restrict__ nodes, NodeAux* __restrict__ aux)
{
    const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_slow) return;
    if (res[e].found) return;                                          // occluded: + SHADOW_COLOR (cpp:80)
    const int i = rays[e].c.y, li = rays[e].c.z;
    const Node nd = nodes[i];
    const unsigned flags = __float_as_uint(nd.B.w);
    const int prim = __float_as_int(nd.P.w);
    const Material M = load_material(sc.materials, __ldg(sc.prim_material + prim));
    const V3 sn = shading_normal(sc, nd, flags, prim);
    const Light L = load_light(sc.light_type, sc.light_f, li);
    const Pix c = calculate_local_color(mk(nd.P.x, nd.P.y, nd.P.z), sn, L, M, mk(fp.cam[0], fp.cam[1], fp.cam[2]));
    unsigned int* w = reinterpret_cast<unsigned int*>(aux[i].local);
    atomicAdd(w, (unsigned)(unsigned short)c.r | ((unsigned)(unsigned short)c.g << 16));
    atomicAdd(w + 1, (unsigned)(unsigned short)c.b);
}

// Deferred shadow rays of the wavefront path (k_shade_gen -> k_anyhit), answered by k_slow after
// k_shade has run: c.y = node * n_nonambient + j.  k_shade saw OCCL_PENDING and left the light's
// term out; an unoccluded ray adds it now.
__global__ void __launch_bounds__(128)
k_shadow_finish(DeviceScene sc, FrameParams fp, const SlowRay* __restrict__ rays, const SlowRes* __restrict__ res,
                unsigned n_slow, const Node* __restrict__ nodes, NodeAux* __restrict__ aux)
{
    const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_slow) return;
    if (res[e].found) return;                                          // occluded: + SHADOW_COLOR (cpp:80)
    const SlowRay r = rays[e];
    const unsigned gid = (unsigned)r.c.y;
    const unsigned i = gid / (unsigned)sc.n_nonambient;
    int j = (int)(gid % (unsigned)sc.n_nonambient), li = 0;
    for (;; li++) { if (__ldg(sc.light_type + li) != RT580_LIGHT_AMBIENT) { if (j == 0) break; j--; } }
    // (the large primitives were tested by k_shade_gen before the ray was queued; a linear-fallback ray saw every record)
    const Node nd = nodes[i];
    const unsigned flags = __float_as_uint(nd.B.w);
    const int prim = __float_as_int(nd.P.w);
    const Material M = load_material(sc.materials, __ldg(sc.prim_material + prim));
    const V3 sn = shading_normal(sc, nd, flags, prim);
    const Light L = load_light(sc.light_type, sc.light_f, li);
    const Pix c = calculate_local_color(mk(nd.P.x, nd.P.y, nd.P.z), sn, L, M, mk(fp.cam[0], fp.cam[1], fp.cam[2]));
    unsigned int* w = reinterpret_cast<unsigned int*>(aux[i].local);
    atomicAdd(w, (unsigned)(unsigned short)c.r | ((unsigned)(unsigned short)c.g << 16));
    atomicAdd(w + 1, (unsigned)(unsigned short)c.b);
}

// shadow rays of one level as a ray queue for k_anyhit: ray id = (node - n0) * n_nonambient + j
__global__ void __launch_bounds__(512)
k_shade_gen(DeviceScene sc, unsigned n0, unsigned n_level, unsigned long long first, unsigned n, const Node* __restrict__ nodes,
            struct ARay* __restrict__ out, unsigned int* __restrict__ n_out, uint32_t* __restrict__ occl);

// deferred AO rays of k_ao: a hit is one more occluded sample of its AO call (cpp:325-326)
__global__ void k_ao_finish(const SlowRay* __restrict__ rays, const SlowRes* __restrict__ res, unsigned n_slow,
                            uint32_t* __restrict__ ao_hits)
{
    const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n_slow && res[e].found) atomicAdd(ao_hits + rays[e].c.y, 1u);
}

// bottom-up: hit nodes per subtree -> parent slot, or per pixel for roots
__global__ void k_subtree(unsigned n0, unsigned n1, const Node* __restrict__ nodes, NodeAux* __restrict__ aux,
                          uint32_t* __restrict__ pix_hits)
{
    const unsigned i = n0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n1) return;
    const uint32_t cnt = 1u + aux[i].sub_refl + aux[i].sub_refr;
    const int parent = __float_as_int(nodes[i].N.w);
    if (parent < 0) pix_hits[__float_as_int(nodes[i].D.w)] = cnt;
    else if (__float_as_uint(nodes[i].B.w) & NF_REFR) aux[parent].sub_refr = cnt;
    else aux[parent].sub_refl = cnt;
}

// ---- exclusive scan of uint32 (three small kernels, 1024 elements per block) -------------
#define SCAN_BLOCK 256
#define SCAN_ITEMS 4
__global__ void __launch_bounds__(SCAN_BLOCK)
k_scan_block(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t* __restrict__ block_sums, unsigned n)
{
    __shared__ uint32_t warp_sums[SCAN_BLOCK / 32];
    const unsigned base = (blockIdx.x * SCAN_BLOCK + threadIdx.x) * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = (base + k < n) ? in[base + k] : 0u; sum += v[k]; }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = (lane < SCAN_BLOCK / 32) ? warp_sums[lane] : 0u;
#pragma unroll
        for (int o = 1; o < SCAN_BLOCK / 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
        if (lane < SCAN_BLOCK / 32) warp_sums[lane] = w;
    }
    __syncthreads();
    uint32_t excl = incl - sum + (warp ? warp_sums[warp - 1] : 0u);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { if (base + k < n) out[base + k] = excl; excl += v[k]; }
    if (threadIdx.x == SCAN_BLOCK - 1) block_sums[blockIdx.x] = excl;
}
__global__ void k_scan_add(uint32_t* __restrict__ out, const uint32_t* __restrict__ block_offs, unsigned n) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] += block_offs[i / (SCAN_BLOCK * SCAN_ITEMS)];
}

// per local row: hit nodes of the row = scan[end] - scan[begin] (+ last element)
__global__ void k_row_counts(const uint32_t* __restrict__ pix_hits, const uint32_t* __restrict__ pix_scan, int W,
                             int n_rows, uint64_t* __restrict__ row_counts)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const size_t first = (size_t)r * W, last = first + W - 1;
    row_counts[r] = (uint64_t)(pix_scan[last] + pix_hits[last] - pix_scan[first]);
}

// top-down: ordinal of every hit node in the reference's traversal order
__global__ void k_preorder(unsigned n0, unsigned n1, const Node* __restrict__ nodes, const NodeAux* __restrict__ aux,
                           const uint32_t* __restrict__ pix_scan, const uint64_t* __restrict__ row_base, FrameParams fp,
                           int n_ambient, uint64_t* __restrict__ pre, uint32_t* __restrict__ ao_state)
{
    const unsigned i = n0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n1) return;
    const int parent = __float_as_int(nodes[i].N.w);
    uint64_t ord;
    if (parent < 0) {
        const int pixel = __float_as_int(nodes[i].D.w);
        const int row = pixel / fp.W;
        if (fp.rng_mode == RT580_RNG_REFERENCE_LCG)
            ord = row_base[row] + (uint64_t)(pix_scan[pixel] - pix_scan[(size_t)row * fp.W]);
        else {
            const uint64_t gp = (uint64_t)(fp.row_first + row * fp.row_step) * fp.W + (pixel % fp.W);
            ord = gp * 32ull;     // counter mode: at most 31 nodes per pixel (depth <= 4)
        }
    } else {
        ord = pre[parent] + 1ull;                                              // node, then reflection subtree,
        if (__float_as_uint(nodes[i].B.w) & NF_REFR) ord += aux[parent].sub_refl;   // then refraction subtree
    }
    pre[i] = ord;
    for (int a = 0; a < n_ambient; a++) {
        const uint64_t call = ord * (uint64_t)n_ambient + a;      // one AO call per ambient light (cpp:41-45, Q6)
        ao_state[(size_t)i * n_ambient + a] = lcg_state_at_tab(2ull * (uint64_t)fp.spp * call, fp.lcg_tab);
    }
}

// Occlusion pass: one thread per AO sample ray (cpp:320-328).
template <int MODE>
__global__ void __launch_bounds__(128)
k_ao(DeviceScene sc, FrameParams fp, unsigned long long n_rays, int n_ambient, const Node* __restrict__ nodes,
     const uint32_t* __restrict__ ao_state, uint32_t* __restrict__ ao_hits, SlowQ sq)
{
    __shared__ PrimRec s_prims[MODE == 1 ? RT_SMEM_PRIMS : 1];
    const PrimRec* sp = stage_prims<MODE>(sc, s_prims);
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < n_rays;
    unsigned call = 0xffffffffu;
    V3 org = mk(0, 0, 0), rd = mk(0, 0, 0);
    if (active) {
        call = (unsigned)(i / (unsigned)fp.spp);
        const unsigned k = (unsigned)(i % (unsigned)fp.spp);
        const unsigned node = call / (unsigned)n_ambient;
        // state after the 2k draws of the preceding samples of this call: s0 * 16807^(2k)
        uint32_t st = lcg_mulmod(__ldg(ao_state + call), __ldg(fp.lcg_pow + k));
        const float4 nP = __ldg(&nodes[node].P), nN = __ldg(&nodes[node].N);
        const V3 P = mk(nP.x, nP.y, nP.z), N = mk(nN.x, nN.y, nN.z);
        const V3 dir = random_in_hemisphere(st, N);                            // cpp:321
        org = P + dir * RT_SHADOW_OFFSET;                                      // cpp:322
        rd = normalize(dir);                                                   // Ray ctor h:431-433
    }
    HitRec h;
    const bool hit = trace_ray<MODE, true>(sc, sp, active, org, rd, __int_as_float(0x7f800000), h, sq, (int)call, 0) == TR_HIT;   // cpp:325
    // count the occluded samples of each AO call inside the warp, one atomic per (warp, call)
    const unsigned peers = __match_any_sync(0xffffffffu, call);
    const unsigned votes = __ballot_sync(0xffffffffu, hit);
    if (active && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) {
        const unsigned c = __popc(votes & peers);
        if (c) atomicAdd(ao_hits + call, c);
    }
}

// ---- occlusion pass, wavefront form -------------------------------------------------------
// k_ao (above) maps one thread to one AO ray: rays of a warp end after very different numbers of
// node visits (any-hit), and ncu shows 9.5 of 32 lanes active on average.  The wavefront form
// splits it: k_ao_gen writes the sample rays (the exact RNG / hemisphere arithmetic of
// cpp:269-292, 32 bytes per ray), k_anyhit is a persistent kernel in which a lane that finishes
// its ray immediately fetches the next one (warp-aggregated atomic on the ray counter), so the
// warps stay full until the queue runs dry.
struct __align__(16) ARay {
    float4 a;   // origin.xyz, dir.x
    float4 b;   // dir.y, dir.z, bits(consumer id: AO call), tmax
};

// Any hit of an unbounded ray against the large-primitive list (staged in shared memory).  "Any hit" is
// an OR over the list, so the order of the exact tests is free: a first pass over the distinct PLANES
// of the list finds the two nearest ahead of the ray with approximate arithmetic (FMA, approximate
// reciprocal: it only ORDERS the tests), which in a closed room is the wall the ray leaves through;
// the exact tests (the reference's, prim_test) run on the triangles of those two planes and only walk
// the rest of the list if all of them miss.  With the tests in list order the lanes of a warp left
// the loop after 1..n_big iterations and ncu showed 6-14 of 32 lanes active over 74 % of k_ao_gen's
// instructions.
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ bool plane_members_any(const PrimRec* __restrict__ s_big, float4 pl, unsigned long long members, V3 O, V3 d)
{
    V3 P;
    if (!plane_point_any(pl, O, d, __int_as_float(0x7f800000), P)) return false;     // t and P are the same for every triangle of the plane
    for (unsigned long long m = members; m; m &= m - 1ull)
        if (tri_bary_accept<false>(&s_big[__ffsll((long long)m) - 1], P)) return true;
    return false;
}
__device__ __forceinline__ bool big_any_nearest_first(const PrimRec* __restrict__ s_big, int n_big, const float4* __restrict__ s_plane,
                                                      const unsigned long long* __restrict__ s_mask, const unsigned char* __restrict__ s_newn,
                                                      int n_planes, unsigned long long sphere_mask, V3 O, V3 d)
{
    const float inf = __int_as_float(0x7f800000);
    float t; int prim;
    for (unsigned long long m = sphere_mask; m; m &= m - 1ull)
        if (prim_test<false>(&s_big[__ffsll((long long)m) - 1], O, d, inf, 0x7fffffff, t, prim)) return true;
    // keys: approximate t (a positive float orders like its bits) with the plane index in the low 6 bits;
    // planes are sorted by normal, parallel ones share N.d, N.O and the reciprocal (the branch is warp-uniform)
    unsigned m1 = 0xffffffffu, m2 = 0xffffffffu;
    float nO = 0.f, r = 0.f;
#pragma unroll 4
    for (int j = 0; j < n_planes; j++) {
        const float4 pl = s_plane[j];
        if (s_newn[j]) {
            nO = __fmaf_rn(pl.x, O.x, __fmaf_rn(pl.y, O.y, pl.z * O.z));
            r = rcp_approx(__fmaf_rn(pl.x, d.x, __fmaf_rn(pl.y, d.y, pl.z * d.z)));
        }
        const float ta = -(nO + pl.w) * r;
        const unsigned key = (ta > 0.0f) ? ((__float_as_uint(ta) & ~63u) | (unsigned)j) : 0xffffffffu;   // behind, parallel, NaN: last
        const unsigned hi = max(key, m1);
        m1 = min(key, m1);
        m2 = min(m2, hi);
    }
    unsigned long long tested = sphere_mask;
    if (m1 != 0xffffffffu) {
        const unsigned long long mm = s_mask[m1 & 63u];
        if (plane_members_any(s_big, s_plane[m1 & 63u], mm, O, d)) return true;
        tested |= mm;
    }
    if (m2 != 0xffffffffu) {
        const unsigned long long mm = s_mask[m2 & 63u];
        if (plane_members_any(s_big, s_plane[m2 & 63u], mm, O, d)) return true;
        tested |= mm;
    }
    for (int k = 0; k < n_big; k++)
        if (!((tested >> k) & 1ull) && prim_test<false>(&s_big[k], O, d, inf, 0x7fffffff, t, prim)) return true;
    return false;
}

__global__ void __launch_bounds__(256, 6)
k_ao_gen(DeviceScene sc, FrameParams fp, unsigned long long first, unsigned n, int n_ambient, const Node* __restrict__ nodes,
         const uint32_t* __restrict__ ao_state, ARay* __restrict__ out, unsigned int* __restrict__ n_out,
         uint32_t* __restrict__ ao_hits)
{
    // the large-primitive list (<= 64 records) is walked by every ray: keep it in shared memory
    __shared__ PrimRec s_big[64];
    __shared__ float4 s_plane[64];
    __shared__ unsigned long long s_mask[64];
    __shared__ unsigned char s_newn[64];
    {
        const float4* src = reinterpret_cast<const float4*>(sc.prims + sc.n_leaf);
        float4* dst = reinterpret_cast<float4*>(s_big);
        for (int i = threadIdx.x; i < sc.n_big * 4; i += blockDim.x) dst[i] = __ldg(src + i);
        for (int i = threadIdx.x; i < sc.n_big_planes; i += blockDim.x) {
            s_plane[i] = __ldg(sc.big_planes + i); s_mask[i] = __ldg(sc.big_masks + i); s_newn[i] = (unsigned char)__ldg(sc.big_plane_newn + i);
        }
        __syncthreads();
    }
    const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = j < n;
    unsigned call = 0xffffffffu;
    bool emit = false, hit = false;
    ARay r;
    r.a = make_float4(0.f, 0.f, 0.f, 0.f); r.b = r.a;
    if (active) {
        const unsigned long long i = first + j;
        unsigned k;
        if (fp.spp_shift >= 0) { call = (unsigned)(i >> fp.spp_shift); k = (unsigned)i & ((unsigned)fp.spp - 1u); }   // spp a power of two: no 64-bit division
        else { call = (unsigned)(i / (unsigned)fp.spp); k = (unsigned)(i % (unsigned)fp.spp); }
        const unsigned node = (n_ambient == 1) ? call : call / (unsigned)n_ambient;
        uint32_t st = lcg_mulmod(__ldg(ao_state + call), __ldg(fp.lcg_pow + k));    // state after the 2k draws before sample k
        const float4 nP = __ldg(&nodes[node].P), nN = __ldg(&nodes[node].N);
        const V3 P = mk(nP.x, nP.y, nP.z), N = mk(nN.x, nN.y, nN.z);
        const V3 dir = random_in_hemisphere(st, N);                                  // cpp:321
        const V3 org = P + dir * RT_SHADOW_OFFSET;                                   // cpp:322
        const V3 rd = normalize(dir);                                                // Ray ctor h:431-433
        r.a = make_float4(org.x, org.y, org.z, rd.x);
        r.b = make_float4(rd.y, rd.z, __uint_as_float(call), __int_as_float(0x7f800000));
        emit = true;
        // the large primitives first: most likely occluders, and the ray is not queued at all if one is hit
        const bool far_origin = sc.farfield && !in_scene(sc, org);
        if (!far_origin && sc.n_big > 0) {
            hit = big_any_nearest_first(s_big, sc.n_big, s_plane, s_mask, s_newn, sc.n_big_planes, sc.big_sphere_mask, org, rd);
            emit = !hit;
        }
    }
    // occluded samples of one AO call inside the warp: one atomic per (warp, call)
    const unsigned peers = __match_any_sync(0xffffffffu, call);
    const unsigned votes = __ballot_sync(0xffffffffu, hit);
    if (active && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) {
        const unsigned c = __popc(votes & peers);
        if (c) atomicAdd(ao_hits + call, c);
    }
    const unsigned slot = warp_alloc(n_out, emit);
    if (emit) out[slot] = r;
}

__global__ void __launch_bounds__(512)
k_shade_gen(DeviceScene sc, unsigned n0, unsigned n_level, unsigned long long first, unsigned n, const Node* __restrict__ nodes,
            ARay* __restrict__ out, unsigned int* __restrict__ n_out, uint32_t* __restrict__ occl)
{
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    bool emit = false;
    ARay r;
    r.a = make_float4(0.f, 0.f, 0.f, 0.f); r.b = r.a;
    if (t < n) {
        // queue order = light-major: the lanes of a warp take neighbouring nodes and the SAME light, so their
        // rays run to one point and walk the tree together (node-major order put three lights in adjacent lanes)
        const unsigned q = (unsigned)first + t;                                // a level has < 2^32 shadow rays (checked by the host)
        int j = (int)(q / n_level);
        const unsigned m = q - (unsigned)j * n_level;
        const unsigned id = m * (unsigned)sc.n_nonambient + (unsigned)j;      // ray id = (node - n0) * n_nonambient + j
        int li = 0;
        for (;; li++) { if (__ldg(sc.light_type + li) != RT580_LIGHT_AMBIENT) { if (j == 0) break; j--; } }
        const Light L = load_light(sc.light_type, sc.light_f, li);
        const float4 nP = __ldg(&nodes[n0 + m].P);
        V3 so, sd; float tmax;
        shadow_ray(L, mk(nP.x, nP.y, nP.z), so, sd, tmax);
        r.a = make_float4(so.x, so.y, so.z, sd.x);
        r.b = make_float4(sd.y, sd.z, __uint_as_float(id), tmax);
        emit = true;
        const bool far_origin = sc.farfield && !in_scene(sc, so);
        // the large primitives here (plane by plane, exact), so that k_shade only has to read the verdict; a ray one
        // of them stops is not queued (a far origin takes the reference's linear loop over ALL records instead)
        const bool in_free_box = sc.big_free_on && __ldg(sc.big_free_light + li) != 0 &&
                                 so.x >= sc.big_free_lo[0] && so.x <= sc.big_free_hi[0] && so.y >= sc.big_free_lo[1] && so.y <= sc.big_free_hi[1] &&
                                 so.z >= sc.big_free_lo[2] && so.z <= sc.big_free_hi[2];
        if (sc.n_big > 0 && !far_origin && !in_free_box) {
            HitRec sh; sh.t = tmax; sh.leaf = -1; sh.prim = 0x7fffffff;
            if (big_scan<true>(sc, so, sd, sh)) { occl[id] = 1u; emit = false; }
        }
        // the light's clearance map may prove that no tree primitive lies between the ray origin and the
        // light (smap.cuh): such a ray is not queued and its occluder count stays 0
        // (not for rays whose hit / miss the tree does not decide alone: a far origin, a light beyond far_tmin)
        const int mi = (emit && sc.n_smap) ? __ldg(sc.smap_of_light + li) : -1;
        if (mi >= 0 && !(sc.farfield && (tmax >= sc.far_tmin || far_origin)) &&
            smap_clear(sc.smap + (size_t)mi * 6 * sc.smap_res * sc.smap_res, sc.smap_res, L.position, so)) emit = false;
    }
    // order-preserving compaction with ONE global atomic per block (one per warp, 3.7 M on the same
    // address for a 4K level, cost as much as the traversal they saved)
    __shared__ unsigned s_cnt[32];
    __shared__ unsigned s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const unsigned mask = __ballot_sync(0xffffffffu, emit);
    if (lane == 0) s_cnt[warp] = (unsigned)__popc(mask);
    __syncthreads();
    if (warp == 0) {
        const unsigned v = lane < n_warps ? s_cnt[lane] : 0u;
        unsigned incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
        s_cnt[lane] = incl - v;
        const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
        if (lane == 0) s_base = total ? atomicAdd(n_out, total) : 0u;
    }
    __syncthreads();
    if (emit) out[s_base + s_cnt[warp] + (unsigned)__popc(mask & ((1u << lane) - 1u))] = r;
}

// Put an any-hit ray on the deferred queue (far-field scan or, `linear`, the reference's own loop,
// trace.cuh).  A full queue drops the ray but still counts it: the host sees count > cap when it
// flushes and repeats the pass with a queue that takes every ray (it never happens in scenes
// that do not leak; keeping the in-place service out of this kernel keeps it at 51 registers).
__device__ __forceinline__ bool defer_any(const SlowQ& sq, V3 O, V3 d, float tmax, bool linear, unsigned gid) {
    if (!sq.rays) return false;
    const unsigned slot = atomicAdd(sq.count, 1u);
    if (slot >= sq.cap) return false;
    SlowRay r; r.o = make_float4(O.x, O.y, O.z, tmax); r.d = make_float4(d.x, d.y, d.z, __int_as_float(0x7fffffff));
    r.c = make_int4(linear ? 1 : 0, (int)gid, 0, -1);
    sq.rays[slot] = r;
    SlowRes z; z.key = 0ull; z.found = 0; z.pad = 0; sq.res[slot] = z;
    return true;
}

// Persistent any-hit traversal.  hit_count[id] += 1 for every ray that is occluded.  A ray the tree
// cannot answer alone goes to the deferred queue `sq` under the id `id + id_offset`; with
// `pending_mark` its hit_count entry is flagged so that the consumer knows the answer comes later.
template <bool COUNT>
__global__ void __launch_bounds__(128, AH_MIN_BLOCKS)
k_anyhit(DeviceScene sc, const ARay* __restrict__ rays, const unsigned int* __restrict__ n_ptr,
         unsigned int* __restrict__ next_ray, uint32_t* __restrict__ hit_count, SlowQ sq, unsigned id_offset,
         unsigned pending_mark, unsigned long long* __restrict__ traversed_acc, int ah_steps, int ah_min_search, int batch_div,
         unsigned long long* __restrict__ visit_counts)
{
    unsigned cnt_nodes = 0, cnt_leaves = 0;
    const unsigned n = __ldg(n_ptr);
    if (traversed_acc && blockIdx.x == 0 && threadIdx.x == 0 && n) atomicAdd(traversed_acc, (unsigned long long)n);
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    // rays a warp reserves per atomic on the queue counter: large for long queues (a single hot counter
    // would serialise the GPU), small for short ones (the last batches are the tail of the kernel)
    unsigned batch = n / (gridDim.x * (blockDim.x >> 5) * (unsigned)batch_div);
    batch = batch > (unsigned)AH_BATCH ? (unsigned)AH_BATCH : (batch < 32u ? 32u : (batch & ~31u));
    bool active = false;
    V3 O = mk(0, 0, 0), d = mk(0, 0, 0), inv = mk(0, 0, 0);
    float tmax = 0.f; unsigned id = 0;
    int stack[RT_STACK_SIZE];
    int sp = 0, cur = AH_NONE;      // cursor: >= 0 inner node, < 0 leaf (~index), AH_NONE nothing left
    bool last_batch = false;        // warp-uniform: the queue counter has run past n, no further batch exists
    unsigned wnext = 0, wend = 0;   // warp-uniform: this warp's current batch [wnext, wend) of the ray queue
    for (;;) {
        const bool exhausted = last_batch && wnext == wend;
        if (!exhausted) {
            const unsigned idle = __ballot_sync(0xffffffffu, !active);
            if (idle) {
                const unsigned want = (unsigned)__popc(idle);
                if (wend - wnext < want) {
                    // the unused tail of the old batch (< 32 rays) is handed out first
                    if (wnext == wend && !last_batch) {
                        unsigned b = 0;
                        if (lane == 0) b = atomicAdd(next_ray, batch);
                        wnext = __shfl_sync(0xffffffffu, b, 0);
                        wend = wnext + batch;
                        if (wend >= n) { last_batch = true; if (wend > n) wend = n; if (wnext > n) wnext = n; }
                    }
                }
                const unsigned base = wnext;
                const unsigned give = min(want, wend - wnext);
                wnext += give;
                if (!active) {
                    const unsigned rank = (unsigned)__popc(idle & lt_mask);
                    const unsigned idx = base + rank;
                    if (rank < give) {
                        const float4 a = __ldg(&rays[idx].a), b = __ldg(&rays[idx].b);
                        O = mk(a.x, a.y, a.z); d = mk(a.w, b.x, b.y);
                        id = __float_as_uint(b.z); tmax = b.w;
                        if (ray_has_nan(O, d)) {
                            if (sc.nan_leaf >= 0) atomicAdd(hit_count + id, 1u);     // "hits" the first triangle of the scene (device_scene.h)
                        } else if (sc.farfield && !in_scene(sc, O)) {
                            // child of a far-field hit: the reference's linear loop, deferred (trace.cuh)
                            if (sc.diag) atomicAdd(sc.diag + 1, 1u);
                            if (defer_any(sq, O, d, tmax, true, id + id_offset) && pending_mark) atomicOr(hit_count + id, pending_mark);
                        } else if (sc.n_leaf > 0) {
                            inv = mk(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
                            cur = 0; sp = 0; active = true;
                        }
                    }
                }
            }
        }
        if (!__any_sync(0xffffffffu, active)) { if (last_batch && wnext == wend) break; continue; }
        // phase 1: inner nodes.  A lane whose next item is a leaf waits here, so that the (long) leaf
        // tests below run with many lanes at once instead of one or two (ncu, profiles/: with leaf
        // tests inline only ~10 of 32 lanes were active per issued instruction).
#pragma unroll 1
        for (int it = 0; it < ah_steps; it++) {
            const bool searching = active && cur >= 0 && cur != AH_NONE;
            const int n_search = __popc(__ballot_sync(0xffffffffu, searching));
            const bool leaf_work = __any_sync(0xffffffffu, active && !searching);
            if (n_search == 0 || (n_search < ah_min_search && leaf_work)) break;
            if (searching) {
                if (COUNT) cnt_nodes++;
                float4 xy0, xy1, z01; int4 kids;
                load_node(sc.nodes + cur, xy0, xy1, z01, kids);
                float tn0, tn1;
                const bool h0 = slab(xy0.x, xy0.y, xy0.z, xy0.w, z01.x, z01.y, O, inv, tmax, tn0);
                const bool h1 = slab(xy1.x, xy1.y, xy1.z, xy1.w, z01.z, z01.w, O, inv, tmax, tn1);
                if (h0 && h1) {
                    int nearc = kids.x, farc = kids.y;
                    if (tn1 < tn0) { nearc = kids.y; farc = kids.x; }
                    if (sp < RT_STACK_SIZE) stack[sp++] = farc;
                    cur = nearc;
                } else if (h0) cur = kids.x;
                else if (h1) cur = kids.y;
                else if (sp > 0) cur = stack[--sp];
                else cur = AH_NONE;
            }
        }
        // phase 2: leaves, and rays that ran out of nodes
        if (active) {
            bool found = false, done = false;
            if (cur == AH_NONE) done = true;
            else if (cur < 0) {
                if (COUNT) cnt_leaves++;
                float t; int prim;
                if (prim_test<true>(sc.prims + (~cur), O, d, tmax, 0x7fffffff, t, prim)) { found = true; done = true; }
                else if (sp > 0) cur = stack[--sp];
                else done = true;
            }
            if (done) {
                active = false;
                if (!found && sc.farfield) {
                    HitRec h; h.t = tmax; h.leaf = -1; h.prim = 0x7fffffff;
                    if (sc.n_always) found = always_scan<true>(sc, O, d, h, false);
                    if (!found && tmax >= sc.far_tmin) {
                        // found nothing nearer than far_tmin: deferred far-field scan (trace.cuh)
                        if (sc.diag) atomicAdd(sc.diag, 1u);
                        if (defer_any(sq, O, d, tmax, false, id + id_offset) && pending_mark) atomicOr(hit_count + id, pending_mark);
                    }
                }
                if (found) atomicAdd(hit_count + id, 1u);
            }
        }
    }
    if (COUNT) {
        unsigned long long a = cnt_nodes, b = cnt_leaves;
        for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
        if (lane == 0) { atomicAdd(visit_counts, a); atomicAdd(visit_counts + 1, b); }
    }
}

// bottom-up per level: cpp:39-51 (ambient term), cpp:85-128 (Fresnel blend in Pixel algebra)
__global__ void k_resolve(DeviceScene sc, FrameParams fp, unsigned n0, unsigned n1, const Node* __restrict__ nodes,
                          NodeAux* __restrict__ aux, const uint32_t* __restrict__ ao_hits, int n_ambient,
                          int16_t* __restrict__ fb)
{
    const unsigned i = n0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n1) return;
    const Node nd = nodes[i];
    const NodeAux a = aux[i];
    const unsigned flags = __float_as_uint(nd.B.w);
    const int bounces = (int)((flags >> NF_BOUNCE_SHIFT) & 0xffu);
    const Material M = load_material(sc.materials, __ldg(sc.prim_material + __float_as_int(nd.P.w)));
    Pix local = mkpix(a.local[0], a.local[1], a.local[2]);
    int amb = 0;
    for (int li = 0; li < sc.n_lights; li++) {
        if (__ldg(sc.light_type + li) != RT580_LIGHT_AMBIENT) continue;
        const Light L = load_light(sc.light_type, sc.light_f, li);
        V3 c = ((M.Cs * M.Ka) * L.color) * L.intensity;                              // cpp:42
        const float occlusion = (float)ao_hits[(size_t)i * n_ambient + amb];         // cpp:326 summed
        const float ao = 1.0f - (occlusion / (float)fp.spp);                         // cpp:329
        c = c * ao;                                                                  // cpp:45
        local = pix_add(local, pix_from_v3(c));                                      // cpp:48-49
        amb++;
    }
    if (bounces > 0) {                                                               // cpp:87
        float kr, kt;
        const V3 N = mk(nd.N.x, nd.N.y, nd.N.z), D = mk(nd.D.x, nd.D.y, nd.D.z);
        compute_fresnel(RT_IOR, N, D, kr, kt);                                       // cpp:114
        const Pix reflC = mkpix(a.refl[0], a.refl[1], a.refl[2]);
        const Pix refrC = mkpix(a.refr[0], a.refr[1], a.refr[2]);
        const Pix fr = pix_muls(pix_muls(reflC, kr), M.Ks);                          // cpp:116
        const Pix ft = pix_muls(pix_muls(refrC, kt), M.Kt);                          // cpp:117
        float albedo = 1 - M.Ks - M.Kt;                                              // cpp:120
        albedo = fmaxf(albedo, 0.0f);                                                // cpp:121
        local = pix_add(pix_add(pix_muls(local, albedo), pix_muls(fr, M.Ks)), pix_muls(ft, M.Kt));   // cpp:124
    }
    const Pix out = pix_clamp(local);                                                // cpp:128
    const int parent = __float_as_int(nd.N.w);
    if (parent < 0) {
        const size_t px = (size_t)__float_as_int(nd.D.w);
        fb[3 * px] = out.r; fb[3 * px + 1] = out.g; fb[3 * px + 2] = out.b;          // cpp:925
    } else {
        short* s = (flags & NF_REFR) ? aux[parent].refr : aux[parent].refl;
        s[0] = out.r; s[1] = out.g; s[2] = out.b;
    }
}

// FlushFrameBufferToPPM's per-channel work (cpp:809-823) on the device: out = lut[value], where lut is the
// caller's table of the reference expression u8(powf(c / 255.0f, 1 / 2.2f) * 255.0f), c = 0..255 (built on
// the host with the host's powf, so the bytes are the reference's by construction; the path only emits
// values in [0,255]: cpp:128 clamps, the background is a constant).  3 B per pixel leave the device, not 6.
__global__ void __launch_bounds__(256)
k_gamma_rgb8(const int16_t* __restrict__ fb, unsigned long long n, const uint8_t* __restrict__ lut_g, uint8_t* __restrict__ out)
{
    __shared__ uint8_t lut[256];
    lut[threadIdx.x] = lut_g[threadIdx.x];
    __syncthreads();
    const unsigned long long i4 = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) * 4ull;
    if (i4 + 4ull <= n && (reinterpret_cast<uintptr_t>(out) & 3u) == 0u) {        // (a caller's device pointer need not be 4-byte aligned)
        const short4 v = *reinterpret_cast<const short4*>(fb + i4);
        uchar4 o;
        o.x = lut[min(max((int)v.x, 0), 255)]; o.y = lut[min(max((int)v.y, 0), 255)];
        o.z = lut[min(max((int)v.z, 0), 255)]; o.w = lut[min(max((int)v.w, 0), 255)];
        *reinterpret_cast<uchar4*>(out + i4) = o;
    } else {
        for (unsigned long long i = i4; i < n && i < i4 + 4ull; i++) out[i] = lut[min(max((int)fb[i], 0), 255)];
    }
}

// ---- multi-GPU, device side ---------------------------------------------------------------
// AO-stream prefix of this rank's rows from the all-gathered per-row hit-node counts of every rank
// (rows interleaved: row y belongs to rank y % world, where it is row y / world).  One block; the
// frame has a few thousand rows.  Replaces a D2H copy, a host prefix sum and an H2D copy per frame.
__global__ void __launch_bounds__(1024)
k_row_bases_interleaved(const uint64_t* __restrict__ all_counts, int world, int rank, int max_rows, int H,
                        uint64_t* __restrict__ row_base)
{
    __shared__ uint64_t warp_sums[32];
    __shared__ uint64_t carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0ull;
    __syncthreads();
    for (int y0 = 0; y0 < H; y0 += 1024) {
        const int y = y0 + (int)threadIdx.x;
        const uint64_t v = (y < H) ? all_counts[(size_t)(y % world) * max_rows + y / world] : 0ull;
        uint64_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint64_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint64_t w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint64_t t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const uint64_t carry = carry_s;
        const uint64_t excl = carry + incl - v + (warp ? warp_sums[warp - 1] : 0ull);
        if (y < H && y % world == rank) row_base[y / world] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_sums[31];
        __syncthreads();
    }
}

// This rank's rows -> the whole frame (rank 0's memory; for the other ranks a peer mapping, so these
// are stores over NVLink: whole rows, 16 bytes per lane, no NCCL gather and no staging buffer).
template <typename T>
__global__ void __launch_bounds__(256)
k_store_band(const T* __restrict__ band, T* __restrict__ frame, unsigned row_units, int n_rows, int row_first, int row_step)
{
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long total = (unsigned long long)row_units * (unsigned)n_rows;
    if (i >= total) return;
    const unsigned row = (unsigned)(i / row_units), u = (unsigned)(i % row_units);
    frame[(size_t)(row_first + (int)row * row_step) * row_units + u] = band[i];
}

// checker kernels: arbitrary rays
template <int MODE, bool ANY>
__global__ void __launch_bounds__(128)
k_trace_rays(DeviceScene sc, long long n, const float* __restrict__ org, const float* __restrict__ dir,
             const float* __restrict__ tmax, int32_t* __restrict__ prim_out, float* __restrict__ t_out,
             uint8_t* __restrict__ hit_out, SlowQ q)
{
    __shared__ PrimRec s_prims[MODE == 1 ? RT_SMEM_PRIMS : 1];
    const PrimRec* sp = stage_prims<MODE>(sc, s_prims);
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < n;
    V3 O = mk(0, 0, 0), d = mk(0, 0, 0);
    if (active) { O = mk(org[3 * i], org[3 * i + 1], org[3 * i + 2]); d = mk(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]); }
    HitRec h;
    // with a deferred queue (the frame path's: far-field grid, rays from outside the scene) the rays the tree cannot answer
    // alone come back TR_PENDING and k_trace_rays_apply writes their answer; without one the warp serves them in place
    const int tr = trace_ray<MODE, ANY>(sc, sp, active, O, d, (ANY && active) ? tmax[i] : 0.f, h, q, (int)i, 0);
    if (!active || tr == TR_PENDING) return;
    const bool hit = tr == TR_HIT;
    if (ANY) hit_out[i] = hit ? 1 : 0;
    else { prim_out[i] = hit ? h.prim : -1; t_out[i] = hit ? h.t : 0.f; }
}
template <bool ANY>
__global__ void k_trace_rays_apply(const SlowRay* __restrict__ rays, const SlowRes* __restrict__ res, unsigned n_slow,
                                   int32_t* __restrict__ prim_out, float* __restrict__ t_out, uint8_t* __restrict__ hit_out)
{
    const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_slow) return;
    const int i = rays[e].c.y;
    if (ANY) hit_out[i] = res[e].found ? 1 : 0;
    else {
        const unsigned long long key = res[e].key;
        const int prim = (int)(unsigned)(key & 0xffffffffull);
        prim_out[i] = prim != 0x7fffffff ? prim : -1;
        t_out[i] = prim != 0x7fffffff ? __uint_as_float((unsigned)(key >> 32)) : 0.f;
    }
}

// traversal profile of arbitrary rays (BVH path): per ray node visits, leaf tests, far-field scans,
// linear fallbacks -> tuning data, not part of the frame path
template <bool ANY>
__global__ void __launch_bounds__(128)
k_trace_profile(DeviceScene sc, long long n, const float* __restrict__ org, const float* __restrict__ dir,
                const float* __restrict__ tmax, unsigned* __restrict__ counts4)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < n;
    V3 O = mk(0, 0, 0), d = mk(0, 0, 0);
    if (active) { O = mk(org[3 * i], org[3 * i + 1], org[3 * i + 2]); d = mk(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]); }
    HitRec h;
    unsigned cnt[4] = { 0, 0, 0, 0 };
    const SlowQ noq = { nullptr, nullptr, nullptr, 0u };
    trace_ray<0, ANY>(sc, nullptr, active, O, d, (ANY && active) ? tmax[i] : 0.f, h, noq, 0, 0, cnt);
    if (!active) return;
    for (int k = 0; k < 4; k++) counts4[4 * i + k] = cnt[k];
}

__global__ void k_hemisphere(float nx, float ny, float nz, unsigned long long step, int n, float* __restrict__ out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    uint32_t st = lcg_state_at(step);
    for (int k = 0; k < n; k++) {
        const V3 v = random_in_hemisphere(st, mk(nx, ny, nz));
        out[3 * k] = v.x; out[3 * k + 1] = v.y; out[3 * k + 2] = v.z;
    }
}
__global__ void k_powf(long long n, const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = powf_glibc(x[i], y[i]);
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
// Shading tables addressed by primitive order index (vertex normals, material index), scattered on the
// device from the flat arrays as they were uploaded: the host used to assemble them (a 48 MB loop per upload).
__global__ void __launch_bounds__(256)
k_scatter_shading(long long n_tris, const float4* __restrict__ n0, const float4* __restrict__ n1, const float4* __restrict__ n2,
                  const int32_t* __restrict__ tri_prim, const int32_t* __restrict__ tri_material, long long n_spheres,
                  const int32_t* __restrict__ sph_prim, const int32_t* __restrict__ sph_material, long long n_prims,
                  int n_materials, float4* __restrict__ vn, int32_t* __restrict__ prim_material, unsigned int* __restrict__ bad)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_tris) {
        const int32_t p = tri_prim[i], m = tri_material[i];
        if (p < 0 || p >= n_prims) { atomicOr(bad, 1u); return; }
        if (m < 0 || m >= n_materials) { atomicOr(bad, 2u); return; }
        vn[3 * (size_t)p] = n0[i]; vn[3 * (size_t)p + 1] = n1[i]; vn[3 * (size_t)p + 2] = n2[i];
        prim_material[p] = m;
    } else if (i < n_tris + n_spheres) {
        const long long k = i - n_tris;
        const int32_t p = sph_prim[k], m = sph_material[k];
        if (p < 0 || p >= n_prims) { atomicOr(bad, 4u); return; }
        if (m < 0 || m >= n_materials) { atomicOr(bad, 2u); return; }
        prim_material[p] = m;
    }
}

static int pick_mode(const rt580_context* c, int traversal) {
    if (traversal == RT580_TRAVERSAL_BVH) return 0;
    if (traversal == RT580_TRAVERSAL_BRUTE_FORCE) return c->sc.n_all <= RT_SMEM_PRIMS ? 1 : 2;
    return c->sc.n_all <= RT_SMEM_PRIMS ? 1 : 0;     // AUTO
}

// Host staging memory (rt580.h): a 64-byte header in front of the block remembers how it was obtained.
extern "C" void* rt580_host_alloc(uint64_t bytes)
{
    static const bool have_gpu = [] { int n = 0; return cudaGetDeviceCount(&n) == cudaSuccess && n > 0; }();
    void* base = nullptr;
    uint64_t pinned = 0;
    if (have_gpu && cudaHostAlloc(&base, (size_t)bytes + 64, cudaHostAllocPortable) == cudaSuccess) pinned = 1;
    else { cudaGetLastError(); base = malloc((size_t)bytes + 64); }
    if (!base) return nullptr;
    uint64_t* h = static_cast<uint64_t*>(base);
    h[0] = 0x52543538304d454dull; h[1] = pinned;
    return static_cast<char*>(base) + 64;
}
extern "C" void rt580_host_free(void* p)
{
    if (!p) return;
    uint64_t* h = reinterpret_cast<uint64_t*>(static_cast<char*>(p) - 64);
    if (h[0] != 0x52543538304d454dull) return;          // not ours
    h[0] = 0;
    if (h[1]) cudaFreeHost(h); else free(h);
}

extern "C" int rt580_create(int device, rt580_context** out)
{
    if (!out) FAIL(RT580_INVALID_ARG, "rt580_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        FAIL(RT580_FAILURE, "rt580_create: no CUDA device (%s); this library has no CPU fallback",
             e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) FAIL(RT580_INVALID_ARG, "rt580_create: device %d out of range [0,%d)", device, n);
    CU(cudaSetDevice(device));
    rt580_context* c = new rt580_context();
    c->device = device;
    CU(cudaGetDeviceProperties(&c->prop, device));
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c->ev_level, cudaEventDisableTiming));
    for (int k = 0; k < 2; k++) {
        CU(cudaStreamCreateWithFlags(&c->side[k], cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&c->ev_join[k], cudaEventDisableTiming));
    }
    if (getenv("RT580_NO_OVERLAP")) c->overlap = false;
    for (auto& ev : c->ev) CU(cudaEventCreate(&ev));
    if (const char* e = getenv("RT580_AH_STEPS")) { if (atoi(e) > 0) c->ah_steps = c->ah_steps_leaky = atoi(e); }
    if (const char* e = getenv("RT580_AH_MIN_SEARCH")) { if (atoi(e) > 0) c->ah_min_search = c->ah_min_search_leaky = atoi(e); }
    if (const char* e = getenv("RT580_AH_BLOCKS_PER_SM")) c->ah_blocks_per_sm = atoi(e) > 0 ? atoi(e) : c->ah_blocks_per_sm;
    if (const char* e = getenv("RT580_SMAP_RES")) c->smap_res = (atoi(e) >= 16 && atoi(e) <= 4096) ? atoi(e) : c->smap_res;
    if (const char* e = getenv("RT580_SLOW_ANY_CAP")) c->slow_any_cap = atoi(e) > 0 ? (unsigned)atoi(e) : c->slow_any_cap;
    if (const char* e = getenv("RT580_AH_BATCH_DIV")) c->ah_batch_div = atoi(e) > 0 ? atoi(e) : c->ah_batch_div;
    if (const char* e = getenv("RT580_CH_BLOCKS_PER_SM")) c->ch_blocks_per_sm = atoi(e) > 0 ? atoi(e) : c->ch_blocks_per_sm;
    if (const char* e = getenv("RT580_ONE_THREAD_PER_RAY")) c->one_thread_per_ray = atoi(e) != 0;
    if (const char* e = getenv("RT580_FAR_GRID")) c->fg_K_env = atoi(e) >= 0 ? atoi(e) : -1;
    if (const char* e = getenv("RT580_FG_DENSE")) { if (atoi(e) > 0) c->fg_dense_rays = atoi(e); }
    if (const char* e = getenv("RT580_ARC_MAX_CELLS")) { if (atoi(e) > 0) c->arc_max_cells = (unsigned)atoi(e); }
    *out = c;
    return RT580_SUCCESS;
}

static void free_scene(rt580_context* c) {
    // every scene pointer lives in c->scene_arena, which is kept for the next upload
    c->d_prims = nullptr; c->d_nodes = nullptr; c->d_far = nullptr; c->d_always = nullptr; c->d_leaf_of_prim = nullptr;
    c->d_vn = nullptr; c->d_prim_material = nullptr; c->d_materials = nullptr; c->d_light_type = nullptr; c->d_light_f = nullptr;
    c->have_scene = false;
}

extern "C" void rt580_destroy(rt580_context* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    free_scene(c);
    frame_release(c);
    arena_release(c->scene_arena); arena_release(c->build_arena);
    c->ndc.release(); c->lcg_pow.release(); c->lcg_tab.release(); c->rgb8.release(); c->nodes.release(); c->aux.release(); c->queue.release(); c->queue2.release(); c->pre.release();
    c->ao_state.release(); c->ao_hits.release(); c->pix_hits.release(); c->pix_scan.release();
    c->scan_tmp.release(); c->row_vals.release(); c->fb.release(); c->counters.release();
    c->slow_rays.release(); c->slow_res.release(); c->any_rays.release(); c->any_res.release(); c->arays.release(); c->occl.release(); c->chits.release();
    for (auto& ev : c->ev) cudaEventDestroy(ev);
    for (auto& ev : c->tm_ev) cudaEventDestroy(ev);
    for (int k = 0; k < 2; k++) { cudaStreamSynchronize(c->side[k]); cudaEventDestroy(c->ev_join[k]); cudaStreamDestroy(c->side[k]); }
    c->arays2.release(); c->occl2.release();
    c->visit_counts.release();
    c->fg_counts.release(); c->fg_start.release(); c->fg_bsum.release(); c->fg_entries.release(); c->fg_cell_tmin.release();
    c->arc_items.release(); c->arc_pre.release();
    c->fgq_hist.release(); c->fgq_start.release(); c->fgq_cellof.release(); c->fgq_rank.release(); c->fgq_order.release(); c->fgq_lin.release(); c->fgq_first.release();
    cudaEventDestroy(c->ev_level);
    cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" int rt580_device_info(rt580_context* c, int32_t* sm_count, int32_t* sm_clock_mhz, uint64_t* hbm_bytes)
{
    if (!c) FAIL(RT580_INVALID_ARG, "rt580_device_info: ctx is NULL");
    if (sm_count) *sm_count = c->prop.multiProcessorCount;
    if (sm_clock_mhz) { int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, c->device); *sm_clock_mhz = khz / 1000; }
    if (hbm_bytes) *hbm_bytes = (uint64_t)c->prop.totalGlobalMem;
    return RT580_SUCCESS;
}

extern "C" int rt580_get_stream(rt580_context* c, void** cuda_stream)
{
    if (!c || !cuda_stream) FAIL(RT580_INVALID_ARG, "rt580_get_stream: NULL argument");
    *cuda_stream = (void*)c->stream;
    return RT580_SUCCESS;
}

static inline unsigned nblk_ll(long long n, unsigned b) { return (unsigned)((n + b - 1) / b); }
template <typename T> static cudaError_t upload(DevArena& a, T** dst, const void* src, size_t count, cudaStream_t s) {
    *dst = a.take<T>(count);
    if (!*dst) return cudaErrorMemoryAllocation;
    if (count) return cudaMemcpyAsync(*dst, src, count * sizeof(T), cudaMemcpyHostToDevice, s);
    return cudaSuccess;
}

// Build the lists of the far-field direction grid for the uploaded scene, if they are not there yet.
static int far_grid_ensure(rt580_context* c)
{
    if (!c->fg_pending) return RT580_SUCCESS;
    c->fg_pending = false;
    cudaStream_t st = c->stream;
    char err[256] = "";
    FgBuildOutput fo{};
    CU(cudaEventRecord(c->ev[10], st));
    if (!fg_build(c->fg_in, &fo, st, err, sizeof err)) FAIL(RT580_FAILURE, "far-field grid: %s", err);
    CU(cudaEventRecord(c->ev[11], st));
    CU(cudaStreamSynchronize(st));
    CU(cudaEventElapsedTime(&c->fg_build_ms, c->ev[10], c->ev[11]));
    c->sc.fg_K = fo.K; c->sc.fg_start = c->fg_start.p; c->sc.fg_entries = c->fg_entries.p; c->sc.fg_cell_tmin = c->fg_cell_tmin.p; c->sc.fg_n_wide = fo.n_wide;
    c->fg_n_entries = fo.n_entries;
    if (getenv("RT580_DEBUG_TIMING"))
        fprintf(stderr, "[rt580] far-field grid: K %d, %llu entries (%.1f per cell), %d wide, t_min %.4g, diag %.4g, build %.2f ms\n", fo.K, fo.n_entries,
                fo.K ? (double)fo.n_entries / (6.0 * fo.K * fo.K) : 0.0, fo.n_wide, fo.t_min, fo.diag, c->fg_build_ms);
    return RT580_SUCCESS;
}

// ---- FlattenScene on the device ----------------------------------------------------------------------------
// Matrix::TransformPoint (h:234-248): ((m0 x + m1 y) + m2 z) + m3 per row, unfused (this file is built -fmad=false), and the
// division by w when w != 1 - the operations and the order of the reference, so the arrays equal the host's bit for bit.
__device__ __forceinline__ float4 transform_point_ref(const float* __restrict__ M, float px, float py, float pz)
{
    float x = M[0] * px + M[1] * py + M[2] * pz + M[3];
    float y = M[4] * px + M[5] * py + M[6] * pz + M[7];
    float z = M[8] * px + M[9] * py + M[10] * pz + M[11];
    const float w = M[12] * px + M[13] * py + M[14] * pz + M[15];
    if (w != 1.0f) { x /= w; y /= w; z /= w; }
    return make_float4(x, y, z, 0.0f);
}
struct FlatDev {
    float4 *v0, *v1, *v2, *n0, *n1, *n2, *sph;
    int32_t *tprim, *tmat, *sprim, *smat;
};
__global__ void __launch_bounds__(256)
k_flatten_tris(long long n_tris, int n_shapes, const long long* __restrict__ shape_tri_first, const int32_t* __restrict__ shape_prim_first,
               const int32_t* __restrict__ shape_mesh, const long long* __restrict__ mesh_first, const float* __restrict__ mesh_tris,
               const float* __restrict__ shape_matrix, FlatDev out)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tris) return;
    int lo = 0, hi = n_shapes;                       // the last shape whose first triangle is <= t (spheres and empty meshes: zero length)
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (shape_tri_first[mid] <= t) lo = mid; else hi = mid; }
    const int k = lo;
    const long long local = t - shape_tri_first[k];
    const float* src = mesh_tris + (mesh_first[shape_mesh[k]] + local) * 18;
    const float* M = shape_matrix + (size_t)k * 16;
    out.v0[t] = transform_point_ref(M, src[0], src[1], src[2]);            // cpp:353
    out.v1[t] = transform_point_ref(M, src[3], src[4], src[5]);            // cpp:354
    out.v2[t] = transform_point_ref(M, src[6], src[7], src[8]);            // cpp:355
    out.n0[t] = make_float4(src[9], src[10], src[11], 0.0f);               // object space
    out.n1[t] = make_float4(src[12], src[13], src[14], 0.0f);
    out.n2[t] = make_float4(src[15], src[16], src[17], 0.0f);
    out.tprim[t] = shape_prim_first[k] + (int32_t)local;
    out.tmat[t] = k;
}
__global__ void __launch_bounds__(256)
k_flatten_spheres(int n_shapes, const int32_t* __restrict__ shape_mesh, const int32_t* __restrict__ shape_prim_first,
                  const int32_t* __restrict__ shape_sph_index, const float* __restrict__ shape_matrix, const float* __restrict__ shape_radius, FlatDev out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_shapes || shape_mesh[k] >= 0) return;
    const float* M = shape_matrix + (size_t)k * 16;
    const int si = shape_sph_index[k];
    out.sph[si] = make_float4(M[3], M[7], M[11], shape_radius[k]);         // h:212-214, cpp:423
    out.sprim[si] = shape_prim_first[k];
    out.smat[si] = k;
}
// host tables of an instanced scene + the kernels; the arrays come out of `ta`
struct InstTables { std::vector<long long> tri_first; std::vector<int32_t> prim_first, sph_index; long long n_tris = 0, n_spheres = 0, n_prims = 0; int64_t first_tri = -1; int32_t first_prim = 0x7fffffff; };
static int inst_tables(const rt580_instanced_scene* s, InstTables& T)
{
    if (!s || s->n_meshes < 0 || s->n_shapes < 0 || s->n_lights < 0) FAIL(RT580_INVALID_ARG, "instanced scene: negative count");
    if ((s->n_meshes && (!s->mesh_first || (s->mesh_first[s->n_meshes] && !s->mesh_tris))) || (s->n_shapes && (!s->shape_mesh || !s->shape_matrix || !s->shape_radius || !s->materials)))
        FAIL(RT580_INVALID_ARG, "instanced scene: NULL table");
    T.tri_first.assign((size_t)s->n_shapes + 1, 0); T.prim_first.assign((size_t)s->n_shapes + 1, 0); T.sph_index.assign((size_t)s->n_shapes + 1, 0);
    long long nt = 0, np = 0, ns = 0;
    for (int k = 0; k < s->n_shapes; k++) {
        T.tri_first[k] = nt; T.prim_first[k] = (int32_t)np; T.sph_index[k] = (int32_t)ns;
        const int m = s->shape_mesh[k];
        if (m >= s->n_meshes) FAIL(RT580_INVALID_ARG, "instanced scene: shape %d names mesh %d of %d", k, m, s->n_meshes);
        if (m < 0) { ns++; np++; }
        else {
            const long long cnt = s->mesh_first[m + 1] - s->mesh_first[m];
            if (cnt < 0) FAIL(RT580_INVALID_ARG, "instanced scene: mesh_first decreases");
            if (cnt > 0 && T.first_tri < 0) { T.first_tri = nt; T.first_prim = (int32_t)np; }
            nt += cnt; np += cnt;
        }
        if (np > 0x7ffffff0ll) FAIL(RT580_INVALID_ARG, "instanced scene: too many primitives");
    }
    T.tri_first[s->n_shapes] = nt;
    T.n_tris = nt; T.n_spheres = ns; T.n_prims = np;
    return RT580_SUCCESS;
}
static int flatten_on_device(const rt580_instanced_scene* s, const InstTables& T, DevArena& ta, cudaStream_t st, FlatDev* out)
{
    FlatDev f{};
    const size_t nt = (size_t)T.n_tris, ns = (size_t)T.n_spheres;
    f.v0 = ta.take<float4>(nt); f.v1 = ta.take<float4>(nt); f.v2 = ta.take<float4>(nt);
    f.n0 = ta.take<float4>(nt); f.n1 = ta.take<float4>(nt); f.n2 = ta.take<float4>(nt);
    f.tprim = ta.take<int32_t>(nt); f.tmat = ta.take<int32_t>(nt);
    f.sph = ta.take<float4>(ns); f.sprim = ta.take<int32_t>(ns); f.smat = ta.take<int32_t>(ns);
    long long *d_tri_first = nullptr, *d_mesh_first = nullptr; int32_t *d_prim_first = nullptr, *d_sph_index = nullptr, *d_shape_mesh = nullptr;
    float *d_mesh_tris = nullptr, *d_matrix = nullptr, *d_radius = nullptr;
    const size_t nsh = (size_t)s->n_shapes;
    CU(upload(ta, &d_tri_first, T.tri_first.data(), nsh + 1, st));
    CU(upload(ta, &d_prim_first, T.prim_first.data(), nsh + 1, st));
    CU(upload(ta, &d_sph_index, T.sph_index.data(), nsh + 1, st));
    CU(upload(ta, &d_shape_mesh, s->shape_mesh, nsh, st));
    CU(upload(ta, &d_mesh_first, s->mesh_first, (size_t)s->n_meshes + (s->n_meshes ? 1 : 0), st));
    CU(upload(ta, &d_mesh_tris, s->mesh_tris, s->n_meshes ? (size_t)s->mesh_first[s->n_meshes] * 18 : 0, st));
    CU(upload(ta, &d_matrix, s->shape_matrix, nsh * 16, st));
    CU(upload(ta, &d_radius, s->shape_radius, nsh, st));
    if ((nt && (!f.v0 || !f.v1 || !f.v2 || !f.n0 || !f.n1 || !f.n2 || !f.tprim || !f.tmat)) || (ns && (!f.sph || !f.sprim || !f.smat)))
        FAIL(RT580_FAILURE, "instanced scene: arena exhausted");
    if (nt) k_flatten_tris<<<nblk_ll(T.n_tris, 256), 256, 0, st>>>(T.n_tris, s->n_shapes, d_tri_first, d_prim_first, d_shape_mesh, d_mesh_first, d_mesh_tris, d_matrix, f);
    if (ns) k_flatten_spheres<<<nblk_ll(s->n_shapes, 256), 256, 0, st>>>(s->n_shapes, d_shape_mesh, d_prim_first, d_sph_index, d_matrix, d_radius, f);
    CU(cudaGetLastError());
    *out = f;
    return RT580_SUCCESS;
}
static size_t inst_input_bytes(const rt580_instanced_scene* s) {
    return (size_t)s->n_shapes * (8 + 4 + 4 + 4 + 64 + 4 + 64) + (size_t)(s->n_meshes + 1) * 8 + (s->n_meshes ? (size_t)s->mesh_first[s->n_meshes] * 72 : 0) + 32 * 256;
}

static int upload_scene_impl(rt580_context* c, const rt580_flat_scene* s, const rt580_instanced_scene* inst, const InstTables* IT)
{
    if (!c || !s) FAIL(RT580_INVALID_ARG, "rt580_upload_scene: NULL argument");
    if (s->n_tris < 0 || s->n_spheres < 0 || s->n_prims != s->n_tris + s->n_spheres || s->n_lights < 0 || s->n_materials < 0)
        FAIL(RT580_INVALID_ARG, "rt580_upload_scene: inconsistent counts");
    if (s->n_prims > 0x7ffffff0ll) FAIL(RT580_INVALID_ARG, "rt580_upload_scene: too many primitives");
    CU(cudaSetDevice(c->device));
    free_scene(c);
    cudaStream_t st = c->stream;
    CU(cudaStreamSynchronize(st));
    {
        char aerr[256] = "";
        const size_t in_bytes = (size_t)s->n_tris * (6 * 16 + 8) + (size_t)s->n_spheres * (16 + 8) + (size_t)s->n_prims * 4 + 8192 + 24 * 256 + (inst ? inst_input_bytes(inst) : 0);
        int n_point = 0;
        for (int i = 0; i < s->n_lights; i++) if (s->light_type && s->light_type[i] == RT580_LIGHT_POINT) n_point++;
        if (n_point > SMAP_MAX) n_point = SMAP_MAX;
        const size_t shade_bytes = (size_t)s->n_prims * (48 + 4 + 28 + 8) + 6 * 256 * 256 * 4 + 4096 + (size_t)s->n_materials * 32 + (size_t)s->n_lights * 52 + 48 * 256 +
                                   (size_t)n_point * 6 * c->smap_res * c->smap_res * sizeof(float);
        if (!arena_reserve(c->build_arena, in_bytes + build_tmp_bytes(s->n_prims), aerr, sizeof aerr) ||
            !arena_reserve(c->scene_arena, shade_bytes + build_out_bytes(s->n_prims, s->n_prims), aerr, sizeof aerr))
            FAIL(RT580_FAILURE, "rt580_upload_scene: %s", aerr);
    }
    DevArena& ta = c->build_arena; DevArena& sa = c->scene_arena;
    float4 *v0 = nullptr, *v1 = nullptr, *v2 = nullptr, *sph = nullptr; int32_t *tprim = nullptr, *sprim = nullptr;
    float4 *tn0 = nullptr, *tn1 = nullptr, *tn2 = nullptr; int32_t *tmat = nullptr, *smat = nullptr; unsigned int* d_bad = nullptr;
    if (inst) {
        FlatDev f{};
        const int fst = flatten_on_device(inst, *IT, ta, st, &f);
        if (fst != RT580_SUCCESS) return fst;
        v0 = f.v0; v1 = f.v1; v2 = f.v2; tprim = f.tprim; sph = f.sph; sprim = f.sprim;
        tn0 = f.n0; tn1 = f.n1; tn2 = f.n2; tmat = f.tmat; smat = f.smat;
    } else {
    CU(upload(ta, &v0, s->tri_v0, (size_t)s->n_tris, st));
    CU(upload(ta, &v1, s->tri_v1, (size_t)s->n_tris, st));
    CU(upload(ta, &v2, s->tri_v2, (size_t)s->n_tris, st));
    CU(upload(ta, &tprim, s->tri_prim, (size_t)s->n_tris, st));
    CU(upload(ta, &sph, s->sph_center_r, (size_t)s->n_spheres, st));
    CU(upload(ta, &sprim, s->sph_prim, (size_t)s->n_spheres, st));
    // shading tables addressed by primitive order index: scattered on the device (k_scatter_shading)
    CU(upload(ta, &tn0, s->tri_n0, (size_t)s->n_tris, st));
    CU(upload(ta, &tn1, s->tri_n1, (size_t)s->n_tris, st));
    CU(upload(ta, &tn2, s->tri_n2, (size_t)s->n_tris, st));
    CU(upload(ta, &tmat, s->tri_material, (size_t)s->n_tris, st));
    CU(upload(ta, &smat, s->sph_material, (size_t)s->n_spheres, st));
    }
    d_bad = ta.take<unsigned int>(1);
    c->d_vn = sa.take<float4>((size_t)s->n_prims * 3);
    c->d_prim_material = sa.take<int32_t>((size_t)s->n_prims);
    if (!d_bad || !c->d_vn || !c->d_prim_material) FAIL(RT580_FAILURE, "rt580_upload_scene: arena exhausted");
    CU(cudaMemsetAsync(d_bad, 0, sizeof(unsigned int), st));
    CU(cudaMemsetAsync(c->d_vn, 0, sizeof(float4) * 3 * (size_t)(s->n_prims ? s->n_prims : 1), st));
    CU(cudaMemsetAsync(c->d_prim_material, 0, sizeof(int32_t) * (size_t)(s->n_prims ? s->n_prims : 1), st));
    if (s->n_prims > 0) {
        k_scatter_shading<<<nblk_ll(s->n_tris + s->n_spheres, 256), 256, 0, st>>>(s->n_tris, tn0, tn1, tn2, tprim, tmat, s->n_spheres, sprim, smat,
                                                                                s->n_prims, s->n_materials, c->d_vn, c->d_prim_material, d_bad);
        unsigned int bad = 0;
        CU(cudaMemcpyAsync(&bad, d_bad, sizeof bad, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (bad & 1u) FAIL(RT580_INVALID_ARG, "rt580_upload_scene: tri_prim out of range");
        if (bad & 4u) FAIL(RT580_INVALID_ARG, "rt580_upload_scene: sph_prim out of range");
        if (bad & 2u) FAIL(RT580_INVALID_ARG, "rt580_upload_scene: material index out of range");
    }
    CU(upload(sa, &c->d_materials, s->materials, (size_t)s->n_materials * 8, st));
    CU(upload(sa, &c->d_light_type, s->light_type, (size_t)s->n_lights, st));
    CU(upload(sa, &c->d_light_f, s->light_f, (size_t)s->n_lights * 10, st));

    BuildInput in{};
    in.tri_v0 = v0; in.tri_v1 = v1; in.tri_v2 = v2; in.tri_prim = tprim; in.n_tris = s->n_tris;
    in.sph = sph; in.sph_prim = sprim; in.n_spheres = s->n_spheres; in.n_prims = s->n_prims;
    // ray origins never leave the hull of the scene and the camera (bvh_build.cu header)
    for (int k = 0; k < 3; k++) in.origin_hint[k] = s->origin_hint[k];
    in.first_tri = -1;
    int32_t first_prim = 0x7fffffff;
    if (inst) { in.first_tri = IT->first_tri; first_prim = IT->first_prim; }
    else for (int64_t i = 0; i < s->n_tris; i++) if (s->tri_prim[i] < first_prim) { first_prim = s->tri_prim[i]; in.first_tri = i; }
    BuildOutput bo{};
    CU(cudaEventRecord(c->ev[10], st));
    char err[256] = "";
    if (!build_bvh(in, &bo, ta, sa, st, err, sizeof err)) FAIL(RT580_FAILURE, "rt580_upload_scene: %s", err);
    CU(cudaEventRecord(c->ev[11], st));
    CU(cudaStreamSynchronize(st));
    CU(cudaEventElapsedTime(&c->build_ms, c->ev[10], c->ev[11]));
    if (bo.max_depth > RT_STACK_SIZE)
        FAIL(RT580_FAILURE, "rt580_upload_scene: LBVH depth %u exceeds the traversal stack (%d)", bo.max_depth, RT_STACK_SIZE);
    c->d_prims = bo.prims; c->d_nodes = bo.nodes; c->d_far = bo.far; c->bvh_depth = bo.max_depth; c->pad_extent = bo.extent;
    c->sc.far = bo.far; c->sc.far_tmin = bo.far_tmin; c->sc.farfield = 1; c->sc.extent = bo.extent;
    c->n_always = bo.n_always; c->n_dropped = bo.n_dropped;
    c->d_always = bo.always_idx; c->sc.always_idx = bo.always_idx; c->sc.n_always = bo.n_always;
    c->d_leaf_of_prim = bo.leaf_of_prim; c->sc.leaf_of_prim = bo.leaf_of_prim;
    c->sc.prims = bo.prims; c->sc.nodes = bo.nodes; c->sc.n_leaf = bo.n_leaf; c->sc.n_big = bo.n_big; c->sc.n_all = bo.n_leaf + bo.n_big;
    c->sc.nan_leaf = bo.nan_leaf; c->sc.nan_prim = bo.nan_leaf >= 0 ? first_prim : -1;
    {
        // in-scene ray origins: every near-field hit lies within pad_max of its primitive's bounds, the rays of cpp:67 / 98 /
        // 110 / 322 start 0.2 further along a unit vector; whatever starts outside that box (the children of far-field hits)
        // takes the reference's own linear loop
        const float margin = bo.pad_max + 0.2001f + 1e-5f * bo.extent;
        for (int k = 0; k < 3; k++) {
            c->sc.ob_lo[k] = bo.bounds_lo[k] - margin; c->sc.ob_hi[k] = bo.bounds_hi[k] + margin; c->sc.ob_cam[k] = s->origin_hint[k];
        }
        // far-field direction grid (fargrid.cuh): the per-triangle constants now, the lists when a frame first needs them
        // (far_grid_ensure: a closed scene never does, and would pay ~65 ms per upload of a million triangles for nothing)
        c->sc.fg_A = nullptr; c->sc.fg_B = nullptr; c->sc.fg_start = nullptr; c->sc.fg_entries = nullptr; c->sc.fg_cell_tmin = nullptr; c->sc.fg_wide = nullptr;
        c->sc.fg_n_wide = 0; c->sc.fg_K = 0; c->sc.fg_dmax = 0.f; c->fg_n_entries = 0; c->fg_build_ms = 0.f;
        c->sc.fg_sph = nullptr; c->sc.fg_n_sph = 0; c->sc.fg_rmax = 0.f; c->sc.fg_tmin = 3.0e38f;
        c->fg_pending = false;
        const int n_all = bo.n_leaf + bo.n_big;
        if (n_all > 0) {
            FgBuildInput& fi = c->fg_in;
            fi = FgBuildInput{};
            fi.prims = bo.prims; fi.far_old = bo.far; fi.n_all = n_all;
            fi.K = 0;                                    // constants only
            fi.extent = bo.extent;
            for (int k = 0; k < 3; k++) { fi.ob_lo[k] = c->sc.ob_lo[k]; fi.ob_hi[k] = c->sc.ob_hi[k]; fi.cam[k] = c->sc.ob_cam[k]; }
            fi.fgA = sa.take<float4>((size_t)n_all); fi.fgB = sa.take<float2>((size_t)n_all); fi.wide = sa.take<uint32_t>((size_t)n_all); fi.sph = sa.take<uint32_t>((size_t)n_all);
            fi.counters = sa.take<unsigned int>(4);
            if (!fi.fgA || !fi.fgB || !fi.wide || !fi.sph || !fi.counters) FAIL(RT580_FAILURE, "rt580_upload_scene: arena exhausted (far-field grid)");
            fi.counts = &c->fg_counts; fi.start = &c->fg_start; fi.bsum = &c->fg_bsum; fi.entries = &c->fg_entries; fi.cell_tmin = &c->fg_cell_tmin;
            FgBuildOutput fo{};
            if (!fg_build(fi, &fo, st, err, sizeof err)) FAIL(RT580_FAILURE, "rt580_upload_scene: %s", err);
            c->sc.fg_A = fi.fgA; c->sc.fg_B = fi.fgB; c->sc.fg_wide = fi.wide; c->sc.fg_n_wide = fo.n_wide;
            c->sc.fg_sph = fi.sph; c->sc.fg_n_sph = fo.n_sph; c->sc.fg_tmin = fo.t_min;
            {
                float rmax = 0.f;
                if (inst) { for (int k = 0; k < inst->n_shapes; k++) if (inst->shape_mesh[k] < 0) rmax = fmaxf(rmax, fabsf(inst->shape_radius[k])); }
                else for (int64_t i = 0; i < s->n_spheres; i++) rmax = fmaxf(rmax, fabsf(s->sph_center_r[4 * i + 3]));
                c->sc.fg_rmax = rmax;
                for (int k = 0; k < 3; k++) c->sc.fg_center[k] = 0.5f * (c->sc.ob_lo[k] + c->sc.ob_hi[k]);
            }
            // no far-field hit nearer than the smallest of the tightened bounds (sliver list aside)
            if (fo.t_min > c->sc.far_tmin) c->sc.far_tmin = fo.t_min;
            fi.K = c->fg_K_env >= 0 ? c->fg_K_env : fg_default_K(n_all);
            if (fi.K > 4096) fi.K = 4096;
            c->fg_pending = fi.K > 0;
            if (c->fg_pending && getenv("RT580_FAR_GRID_EAGER")) { if (far_grid_ensure(c)) return RT580_FAILURE; }
        }
    }
    {
        // distinct planes of the large triangles (device_scene.h); at most 64 records, grouped on the host
        c->sc.big_planes = nullptr; c->sc.big_masks = nullptr; c->sc.big_plane_newn = nullptr; c->sc.big_sphere_mask = 0ull; c->sc.n_big_planes = 0;
        c->sc.big_free_on = 0; c->sc.big_free_light = nullptr;
        if (bo.n_big > 64) FAIL(RT580_FAILURE, "rt580_upload_scene: %d large primitives (at most 64)", bo.n_big);
        if (bo.n_big > 0) {
            std::vector<PrimRec> big((size_t)bo.n_big);
            CU(cudaMemcpy(big.data(), bo.prims + bo.n_leaf, sizeof(PrimRec) * bo.n_big, cudaMemcpyDeviceToHost));
            std::vector<float4> planes; std::vector<unsigned long long> masks;
            for (int k = 0; k < bo.n_big; k++) {
                int flags; memcpy(&flags, &big[k].d.w, 4);
                if (flags & RT_PRIM_SPHERE) { c->sc.big_sphere_mask |= 1ull << k; continue; }
                const float4 pl = make_float4(big[k].d.x, big[k].d.y, big[k].d.z, big[k].a.w);
                size_t j = 0;
                for (; j < planes.size(); j++) if (memcmp(&planes[j], &pl, sizeof pl) == 0) break;
                if (j == planes.size()) { planes.push_back(pl); masks.push_back(0ull); }
                masks[j] |= 1ull << k;
            }
            if (!planes.empty()) {
                // planes with the same normal next to each other (parallel walls), first occurrence order otherwise
                std::vector<float4> sp; std::vector<unsigned long long> sm; std::vector<int32_t> newn;
                std::vector<char> used(planes.size(), 0);
                for (size_t a = 0; a < planes.size(); a++) {
                    if (used[a]) continue;
                    for (size_t b = a; b < planes.size(); b++) {
                        if (used[b] || memcmp(&planes[a], &planes[b], 3 * sizeof(float)) != 0) continue;
                        used[b] = 1;
                        sp.push_back(planes[b]); sm.push_back(masks[b]); newn.push_back(b == a ? 1 : 0);
                    }
                }
                float4* dp = nullptr; unsigned long long* dm = nullptr; int32_t* dn = nullptr;
                CU(upload(sa, &dp, sp.data(), sp.size(), st));
                CU(upload(sa, &dm, sm.data(), sm.size(), st));
                CU(upload(sa, &dn, newn.data(), newn.size(), st));
                CU(cudaStreamSynchronize(st));
                c->sc.big_planes = dp; c->sc.big_masks = dm; c->sc.big_plane_newn = dn; c->sc.n_big_planes = (int)sp.size();
            }
            // the free box (device_scene.h): seeded with the bounds of the tree (its root's two child boxes) and the point
            // lights, each face then pushed outwards as far as no large triangle is met
            c->sc.big_free_on = 0; c->sc.big_free_light = nullptr;
            if (c->sc.big_sphere_mask == 0ull && bo.n_leaf > 1 && !getenv("RT580_NO_FREE_BOX")) {
                BvhNode root;
                CU(cudaMemcpy(&root, bo.nodes, sizeof root, cudaMemcpyDeviceToHost));
                double lo[3] = { fmin(root.xy0.x, root.xy1.x), fmin(root.xy0.z, root.xy1.z), fmin(root.z01.x, root.z01.z) };
                double hi[3] = { fmax(root.xy0.y, root.xy1.y), fmax(root.xy0.w, root.xy1.w), fmax(root.z01.y, root.z01.w) };
                for (int i = 0; i < s->n_lights; i++) if (s->light_type[i] == RT580_LIGHT_POINT)
                    for (int k = 0; k < 3; k++) { lo[k] = fmin(lo[k], (double)s->light_f[10 * i + 4 + k] - 0.3); hi[k] = fmax(hi[k], (double)s->light_f[10 * i + 4 + k] + 0.3); }
                const double margin = fmax(64.0 * (double)bo.pad, 1e-4 * (double)bo.extent);
                // a large triangle leaves the box alone if its plane keeps all 8 corners on one side by `margin`,
                // or if its own bounding box stays `margin` away from the box
                auto clear_of = [&](const double* blo, const double* bhi) {
                    for (int k = 0; k < bo.n_big; k++) {
                        const PrimRec& r = big[k];
                        const double N[3] = { r.d.x, r.d.y, r.d.z }, D = r.a.w;
                        double smin = 1e300, smax = -1e300;
                        for (int cidx = 0; cidx < 8; cidx++) {
                            const double x = (cidx & 1) ? bhi[0] : blo[0], y = (cidx & 2) ? bhi[1] : blo[1], z = (cidx & 4) ? bhi[2] : blo[2];
                            const double sd = N[0] * x + N[1] * y + N[2] * z + D;
                            smin = fmin(smin, sd); smax = fmax(smax, sd);
                        }
                        if (smin >= margin || smax <= -margin) continue;
                        const float* v[3] = { &r.a.x, &r.b.x, &r.c.x };
                        bool apart = false;
                        for (int ax = 0; ax < 3; ax++) {
                            const double tlo = fmin(fmin(v[0][ax], v[1][ax]), v[2][ax]) - margin, thi = fmax(fmax(v[0][ax], v[1][ax]), v[2][ax]) + margin;
                            if (tlo > bhi[ax] || thi < blo[ax]) apart = true;
                        }
                        if (!apart) return false;
                    }
                    return true;
                };
                if (clear_of(lo, hi)) {
                    const double E = (double)bo.extent;
                    for (int face = 0; face < 6; face++) {
                        const int ax = face >> 1; const bool up = face & 1;
                        double good = up ? hi[ax] : lo[ax], bad = up ? E : -E;
                        for (int it = 0; it < 48; it++) {
                            const double mid = 0.5 * (good + bad);
                            double tl[3] = { lo[0], lo[1], lo[2] }, th[3] = { hi[0], hi[1], hi[2] };
                            (up ? th[ax] : tl[ax]) = mid;
                            if (clear_of(tl, th)) good = mid; else bad = mid;
                        }
                        (up ? hi[ax] : lo[ax]) = good;
                    }
                    std::vector<int32_t> lin((size_t)s->n_lights, 0);
                    for (int i = 0; i < s->n_lights; i++) {
                        if (s->light_type[i] != RT580_LIGHT_POINT) continue;
                        bool in = true;
                        for (int k = 0; k < 3; k++) { const double x = s->light_f[10 * i + 4 + k]; if (!(x - 0.25 >= lo[k] && x + 0.25 <= hi[k])) in = false; }
                        lin[i] = in ? 1 : 0;
                    }
                    int32_t* dl = nullptr;
                    CU(upload(sa, &dl, lin.data(), lin.size(), st));
                    CU(cudaStreamSynchronize(st));
                    for (int k = 0; k < 3; k++) {
                        // inwards by an ulp-scale step: the float box lies inside the double one
                        c->sc.big_free_lo[k] = (float)lo[k] + 1e-6f * (float)bo.extent;
                        c->sc.big_free_hi[k] = (float)hi[k] - 1e-6f * (float)bo.extent;
                    }
                    c->sc.big_free_light = dl; c->sc.big_free_on = 1;
                }
            }
        }
    }
    {
        // clearance maps of the point lights (smap.cuh)
        c->sc.smap = nullptr; c->sc.smap_of_light = nullptr; c->sc.n_smap = 0;
        std::vector<int32_t> of_light((size_t)s->n_lights, -1);
        int n_maps = 0;
        // (sliver triangles are tested for every ray whatever the tree says: a scene that has them gets no maps)
        const bool use_maps = bo.n_leaf > RT_SMEM_PRIMS && bo.n_always == 0 && !getenv("RT580_NO_SMAP");
        if (use_maps) for (int i = 0; i < s->n_lights && n_maps < SMAP_MAX; i++) if (s->light_type[i] == RT580_LIGHT_POINT) of_light[i] = n_maps++;
        if (n_maps) {
            const size_t per = (size_t)6 * c->smap_res * c->smap_res;
            c->sc.smap_res = c->smap_res;
            float* maps = sa.take<float>(per * n_maps);
            unsigned int* clear = ta.take<unsigned int>(SMAP_MAX);
            if (!maps || !clear) FAIL(RT580_FAILURE, "rt580_upload_scene: arena exhausted (clearance maps)");
            CU(cudaMemsetAsync(maps, 0x7f, sizeof(float) * per * n_maps, st));        // 0x7f7f7f7f = 3.4e38: nothing there
            CU(cudaMemsetAsync(clear, 0x7f, sizeof(unsigned int) * SMAP_MAX, st));
            const int n_nodes = bo.n_leaf > 1 ? bo.n_leaf - 1 : 1;
            for (int i = 0; i < s->n_lights; i++) {
                if (of_light[i] < 0) continue;
                const float* lf = s->light_f + 10 * (size_t)i;
                k_smap_raster<<<nblk_ll(2ll * n_nodes, 256), 256, 0, st>>>(bo.nodes, n_nodes, lf[4], lf[5], lf[6], maps + per * of_light[i],
                                                                        clear + of_light[i], c->smap_res);
            }
            float hclear[SMAP_MAX];
            CU(cudaMemcpyAsync(hclear, clear, sizeof hclear, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            CU(cudaGetLastError());
            // a primitive within reach of the 0.2 units by which the reference's shadow ray overshoots the light: no map
            for (int i = 0; i < s->n_lights; i++) if (of_light[i] >= 0 && !(hclear[of_light[i]] >= SMAP_CLEARANCE)) of_light[i] = -1;
            int32_t* dl = nullptr;
            CU(upload(sa, &dl, of_light.data(), of_light.size(), st));
            CU(cudaStreamSynchronize(st));
            c->sc.smap = maps; c->sc.smap_of_light = dl; c->sc.n_smap = n_maps;
        }
    }
    c->sc.n_prims = (int32_t)s->n_prims;
    c->sc.vn = c->d_vn; c->sc.prim_material = c->d_prim_material; c->sc.materials = c->d_materials;
    c->sc.n_materials = s->n_materials; c->sc.light_type = c->d_light_type; c->sc.light_f = c->d_light_f;
    c->sc.n_lights = s->n_lights;
    c->sc.n_ambient = 0;
    for (int i = 0; i < s->n_lights; i++) if (s->light_type[i] == RT580_LIGHT_AMBIENT) c->sc.n_ambient++;
    c->sc.n_nonambient = s->n_lights - c->sc.n_ambient;
    c->have_scene = true;
    c->frame_begun = false;
    c->force_leaky = false;
    return RT580_SUCCESS;
}

extern "C" int rt580_upload_scene(rt580_context* c, const rt580_flat_scene* s)
{
    return upload_scene_impl(c, s, nullptr, nullptr);
}
extern "C" int rt580_upload_instanced_scene(rt580_context* c, const rt580_instanced_scene* s)
{
    if (!c || !s) FAIL(RT580_INVALID_ARG, "rt580_upload_instanced_scene: NULL argument");
    InstTables T;
    const int st = inst_tables(s, T);
    if (st != RT580_SUCCESS) return st;
    rt580_flat_scene fs;
    memset(&fs, 0, sizeof fs);
    fs.n_prims = T.n_prims; fs.n_tris = T.n_tris; fs.n_spheres = T.n_spheres;
    fs.n_materials = s->n_shapes; fs.materials = s->materials;
    fs.n_lights = s->n_lights; fs.light_type = s->light_type; fs.light_f = s->light_f;
    for (int k = 0; k < 3; k++) fs.origin_hint[k] = s->origin_hint[k];
    return upload_scene_impl(c, &fs, s, &T);
}
extern "C" int rt580_flatten_instanced(rt580_context* c, const rt580_instanced_scene* s, float* tri_v0, float* tri_v1, float* tri_v2,
                                       float* tri_n0, float* tri_n1, float* tri_n2, int32_t* tri_prim, int32_t* tri_material,
                                       float* sph_center_r, int32_t* sph_prim, int32_t* sph_material)
{
    if (!c || !s) FAIL(RT580_INVALID_ARG, "rt580_flatten_instanced: NULL argument");
    InstTables T;
    const int ist = inst_tables(s, T);
    if (ist != RT580_SUCCESS) return ist;
    CU(cudaSetDevice(c->device));
    free_scene(c);                                   // (the build arena is the scratch space: whatever scene was uploaded is gone)
    cudaStream_t st = c->stream;
    CU(cudaStreamSynchronize(st));
    char aerr[256] = "";
    if (!arena_reserve(c->build_arena, (size_t)T.n_tris * (6 * 16 + 8) + (size_t)T.n_spheres * 24 + inst_input_bytes(s) + 8192, aerr, sizeof aerr))
        FAIL(RT580_FAILURE, "rt580_flatten_instanced: %s", aerr);
    FlatDev f{};
    const int fst = flatten_on_device(s, T, c->build_arena, st, &f);
    if (fst != RT580_SUCCESS) return fst;
    const size_t nt = (size_t)T.n_tris, ns = (size_t)T.n_spheres;
    struct { void* dst; const void* src; size_t bytes; } cp[] = {
        { tri_v0, f.v0, nt * 16 }, { tri_v1, f.v1, nt * 16 }, { tri_v2, f.v2, nt * 16 }, { tri_n0, f.n0, nt * 16 }, { tri_n1, f.n1, nt * 16 },
        { tri_n2, f.n2, nt * 16 }, { tri_prim, f.tprim, nt * 4 }, { tri_material, f.tmat, nt * 4 }, { sph_center_r, f.sph, ns * 16 },
        { sph_prim, f.sprim, ns * 4 }, { sph_material, f.smat, ns * 4 } };
    for (auto& x : cp) if (x.dst && x.bytes) CU(cudaMemcpyAsync(x.dst, x.src, x.bytes, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return RT580_SUCCESS;
}

extern "C" int rt580_scene_info_get(rt580_context* c, rt580_scene_info* out) {
    if (!c || !out) FAIL(RT580_INVALID_ARG, "rt580_scene_info_get: NULL argument");
    if (!c->have_scene) FAIL(RT580_FAILURE, "rt580_scene_info_get: no scene uploaded");
    memset(out, 0, sizeof *out);
    out->n_leaf = c->sc.n_all; out->n_dropped = c->n_dropped; out->n_always = c->n_always;
    out->far_tmin = c->sc.far_tmin; out->pad = c->pad_extent / 262144.0f; out->extent = c->pad_extent;
    out->build_ms = c->build_ms; out->bvh_max_depth = c->bvh_depth;
    return RT580_SUCCESS;
}

extern "C" int rt580_build_ms(rt580_context* c, float* ms) {
    if (!c || !ms) FAIL(RT580_INVALID_ARG, "rt580_build_ms: NULL argument");
    *ms = c->build_ms;
    return RT580_SUCCESS;
}

static inline unsigned nblk(unsigned long long n, unsigned b) { return (unsigned)((n + b - 1) / b); }

// the context's stream waits for what the side streams have been given so far
static int join_side(rt580_context* c) {
    for (int k = 0; k < 2; k++) { CU(cudaEventRecord(c->ev_join[k], c->side[k])); CU(cudaStreamWaitEvent(c->stream, c->ev_join[k], 0)); }
    return RT580_SUCCESS;
}
static int sync_side(rt580_context* c) {
    for (int k = 0; k < 2; k++) CU(cudaStreamSynchronize(c->side[k]));
    return RT580_SUCCESS;
}
static SlowQ slowq(rt580_context* c, unsigned cap) {
    SlowQ q; q.rays = c->slow_rays.p; q.res = c->slow_res.p; q.count = c->counters.p + 2; q.cap = cap;
    return q;
}
// the second deferred queue: any-hit rays (shadow rays of all levels of a frame, then its AO rays)
static SlowQ slowq_any(rt580_context* c) {
    SlowQ q; q.rays = c->any_cap ? c->any_rays.p : nullptr; q.res = c->any_res.p; q.count = c->counters.p + 3; q.cap = c->any_cap;
    return q;
}
template <int MODE> static void launch_trace(rt580_context* c, bool primary, const QRay* queue, unsigned n_items, const unsigned* n_items_dev,
                                             unsigned node_cap, unsigned slow_cap, Spawn spawn) {
    if (primary)
        k_trace<MODE, true><<<nblk(n_items, 128), 128, 0, c->stream>>>(c->sc, c->fp, queue, n_items, n_items_dev, c->nodes.p, c->aux.p,
                                                                      c->counters.p, c->pix_hits.p, c->fb.p, node_cap, slowq(c, slow_cap), spawn);
    else
        k_trace<MODE, false><<<nblk(n_items, 128), 128, 0, c->stream>>>(c->sc, c->fp, queue, n_items, n_items_dev, c->nodes.p, c->aux.p,
                                                                       c->counters.p, c->pix_hits.p, c->fb.p, node_cap, slowq(c, slow_cap), spawn);
    c->launches++;
}
template <int MODE> static void launch_shade(rt580_context* c, unsigned n0, unsigned n1, unsigned slow_cap) {
    k_shade<MODE><<<nblk(n1 - n0, 128), 128, 0, c->stream>>>(c->sc, c->fp, n0, n1, c->nodes.p, c->aux.p, c->queue.p,
                                                             c->counters.p, slowq(c, slow_cap));
    c->launches++;
}
template <int MODE> static void launch_ao(rt580_context* c, unsigned long long n_rays, unsigned slow_cap) {
    k_ao<MODE><<<nblk(n_rays, 128), 128, 0, c->stream>>>(c->sc, c->fp, n_rays, c->sc.n_ambient, c->nodes.p, c->ao_state.p, c->ao_hits.p,
                                                         slowq(c, slow_cap));
    c->launches++;
}
#define DISPATCH_MODE(mode, fn, ...) do { if ((mode) == 0) fn<0>(__VA_ARGS__); else if ((mode) == 1) fn<1>(__VA_ARGS__); else fn<2>(__VA_ARGS__); } while (0)

static int exclusive_scan_u32(rt580_context* c, const uint32_t* in, uint32_t* out, unsigned n)
{
    // recursive block scan; levels live in scan_tmp
    const unsigned per = SCAN_BLOCK * SCAN_ITEMS;
    std::vector<unsigned> sizes; sizes.push_back(n);
    while (sizes.back() > 1) sizes.push_back(nblk(sizes.back(), per));
    size_t total = 0; for (size_t l = 1; l < sizes.size(); l++) total += 2 * (size_t)sizes[l];
    CU(c->scan_tmp.ensure(total + 2, 0, c->stream));
    std::vector<uint32_t*> sums(sizes.size(), nullptr), offs(sizes.size(), nullptr);
    uint32_t* p = c->scan_tmp.p;
    for (size_t l = 1; l < sizes.size(); l++) { sums[l] = p; p += sizes[l]; offs[l] = p; p += sizes[l]; }
    const uint32_t* src = in; uint32_t* dst = out;
    // up-sweep
    for (size_t l = 0; l + 1 < sizes.size(); l++) {
        k_scan_block<<<sizes[l + 1], SCAN_BLOCK, 0, c->stream>>>(src, dst, sums[l + 1], sizes[l]); c->launches++;
        src = sums[l + 1]; dst = offs[l + 1];
    }
    if (sizes.size() == 1) { CU(cudaMemsetAsync(out, 0, sizeof(uint32_t) * n, c->stream)); return RT580_SUCCESS; }
    // the top level has one element: its exclusive scan is 0
    CU(cudaMemsetAsync(offs.back(), 0, sizeof(uint32_t), c->stream));
    // down-sweep
    for (size_t l = sizes.size() - 1; l >= 1; l--) {
        uint32_t* target = (l == 1) ? out : offs[l - 1];
        k_scan_add<<<nblk(sizes[l - 1], 256), 256, 0, c->stream>>>(target, offs[l], sizes[l - 1]); c->launches++;
    }
    return RT580_SUCCESS;
}

// Device counters (uint32[N_COUNTERS]): [0] hit nodes, [1] queued secondary rays, [2] deferred closest-hit
// rays, [3] deferred any-hit rays, [4] far-field scans, [5] linear fallbacks (diagnostics, whole frame),
// [6] any-hit rays generated for the running chunk, [7] any-hit rays fetched, [8..9] uint64: AO rays that
// went through the tree.  All of them come back in one 64-byte copy + one stream sync (every host round
// trip idles the GPU for ~10-20 us; with 8 ranks a frame is only ~10 ms long).
#define N_COUNTERS 32     // [16..19] / [20..23] k_fg_arc any / closest: rays, -, cells visited, exact tests
static int read_counters(rt580_context* c, unsigned out[N_COUNTERS]) {
    CU(cudaMemcpyAsync(out, c->counters.p, N_COUNTERS * sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->slow_seen = (unsigned long long)out[4] + out[5];
    c->syncs++;
    return RT580_SUCCESS;
}

// Reserve the deferred closest-hit queue for a launch that may record up to `max_rays` slow rays.
static int slow_prepare(rt580_context* c, unsigned long long max_rays, unsigned* cap_out)
{
    unsigned cap = (unsigned)(max_rays < (unsigned long long)SLOW_CAP_MAX ? max_rays : (unsigned long long)SLOW_CAP_MAX);
    if (!c->sc.farfield || c->sc.n_all <= RT_SMEM_PRIMS) cap = 0;        // linear modes never defer
    if (cap) {
        CU(c->slow_rays.ensure(cap, 0, c->stream));
        CU(c->slow_res.ensure(cap, 0, c->stream));
    }
    CU(cudaMemsetAsync(c->counters.p + 2, 0, sizeof(unsigned), c->stream));
    *cap_out = cap;
    return RT580_SUCCESS;
}
// Same for the any-hit queue, which lives across launches until it is flushed.
static int any_prepare(rt580_context* c, unsigned long long max_rays, cudaStream_t st)
{
    unsigned cap = (unsigned)(max_rays < 0xfffffff0ull ? max_rays : 0xfffffff0ull);
    if (!c->sc.farfield || c->sc.n_all <= RT_SMEM_PRIMS) cap = 0;
    if (cap) {
        CU(c->any_rays.ensure(cap, 0, st));
        CU(c->any_res.ensure(cap, 0, st));
    }
    CU(cudaMemsetAsync(c->counters.p + 3, 0, sizeof(unsigned), st));
    c->any_cap = cap;
    return RT580_SUCCESS;
}
// Answer n recorded slow rays.  With a far-field direction grid: the rays are sorted by the cell of their direction
// and answered from that cell's list (k_fg_scan); only the rays that start outside the scene ("linear") go through
// k_slow's scan of every record.  Without a grid (small scenes) k_slow answers all of them.
static int exclusive_scan_u32(rt580_context* c, const uint32_t* in, uint32_t* out, unsigned n);
static int slow_launch(rt580_context* c, bool any, const SlowRay* rays, SlowRes* res, unsigned n)
{
    if (!n) return RT580_SUCCESS;
    cudaStream_t st = c->stream;
    const unsigned int* lin_idx = nullptr;
    unsigned n_lin = n;
    if (tm_begin(c, any ? RT580_CLASS_FAR_ANY : RT580_CLASS_FAR_CLOSEST, st)) return RT580_FAILURE;
    c->prof.rays[any ? RT580_CLASS_FAR_ANY : RT580_CLASS_FAR_CLOSEST] += n;
    // the lists of the far-field grid are built when a flush is first large enough to need them (a closed scene defers a few
    // hundred rays per frame: the scan of every record is cheaper than ~65 ms of build per million triangles)
    if (c->fg_pending && n >= 2048u) { if (far_grid_ensure(c)) return RT580_FAILURE; }
    if (c->sc.fg_K > 0) {
        const size_t n_cells = (size_t)6 * c->sc.fg_K * c->sc.fg_K;
        const bool sorted = n >= 32768u;                 // (below that the sort by cell - a histogram over 6 K^2 cells - costs more than it saves)
        CU(c->fgq_hist.ensure(n_cells + 3, 0, st)); CU(c->fgq_start.ensure(n_cells + 3, 0, st));
        if (any) CU(c->fgq_first.ensure(n, 0, st));
        CU(c->fgq_cellof.ensure(n, 0, st)); CU(c->fgq_rank.ensure(n, 0, st)); CU(c->fgq_order.ensure(n, 0, st)); CU(c->fgq_lin.ensure((size_t)n + 1, 0, st));
        unsigned int* lin_count = c->fgq_hist.p + n_cells + 1;       // (the scan below covers n_cells + 1 elements: [n_cells] stays 0)
        unsigned int* first_count = lin_count + 1;
        if (sorted) CU(cudaMemsetAsync(c->fgq_hist.p, 0, sizeof(unsigned) * (n_cells + 3), st));
        else CU(cudaMemsetAsync(lin_count, 0, 2 * sizeof(unsigned), st));
        k_fg_bin<<<nblk(n, 256), 256, 0, st>>>(rays, n, c->sc.fg_K, sorted ? c->fgq_hist.p : nullptr, c->fgq_cellof.p, c->fgq_rank.p, c->fgq_lin.p, lin_count, any ? 1 : 0,
                                               c->fgq_first.p, first_count);
        if (sorted) {
            if (exclusive_scan_u32(c, c->fgq_hist.p, c->fgq_start.p, (unsigned)(n_cells + 1))) return RT580_FAILURE;
            k_fg_order<<<nblk(n, 256), 256, 0, st>>>(n, c->fgq_cellof.p, c->fgq_rank.p, c->fgq_start.p, c->fgq_order.p);
        }
        const unsigned int* order = sorted ? c->fgq_order.p : nullptr;
        const bool dense = sorted && (unsigned long long)n >= (unsigned long long)c->fg_dense_rays * n_cells;      // rays per cell of the direction grid
        if (any) {
            if (dense) k_fg_scan<true, 32><<<nblk(n, 32 * FG_WARPS), 32 * FG_WARPS, 0, st>>>(c->sc, rays, res, order, c->fgq_cellof.p, c->fgq_start.p + n_cells, n, c->fgq_lin.p, lin_count);
            else k_fg_scan<true, 8><<<nblk(n, 8 * FG_WARPS), 32 * FG_WARPS, 0, st>>>(c->sc, rays, res, order, c->fgq_cellof.p, c->fgq_start.p + n_cells, n, c->fgq_lin.p, lin_count);
            k_fg_lin_first<<<nblk(n, 128), 128, 0, st>>>(c->sc, rays, res, c->fgq_first.p, first_count, c->fgq_cellof.p, c->fgq_lin.p, lin_count);
            c->launches++;
        }
        else if (dense) k_fg_scan<false, 32><<<nblk(n, 32 * FG_WARPS), 32 * FG_WARPS, 0, st>>>(c->sc, rays, res, order, c->fgq_cellof.p, c->fgq_start.p + n_cells, n, c->fgq_lin.p, lin_count);
        else k_fg_scan<false, 8><<<nblk(n, 8 * FG_WARPS), 32 * FG_WARPS, 0, st>>>(c->sc, rays, res, order, c->fgq_cellof.p, c->fgq_start.p + n_cells, n, c->fgq_lin.p, lin_count);
        c->launches += 3;
        CU(cudaMemcpyAsync(&n_lin, lin_count, sizeof n_lin, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        c->syncs++;
        lin_idx = c->fgq_lin.p;
    }
    if (n_lin && lin_idx) {
        // rays from outside the scene: far regime along the arc of the direction grid, near regime through the inflated tree;
        // the few whose inflated boxes cover much of the scene come back in a list for the linear scan below
        CU(c->fgq_rank.ensure(3 * (size_t)n_lin + 3, 0, st));    // (free again: k_fg_order has consumed the ranks)
        unsigned int* heavy_idx = c->fgq_rank.p;                 // k_lin_near's: inflated boxes cover much of the scene -> k_slow
        unsigned int* farheavy_idx = c->fgq_rank.p + n_lin + 1;  // k_fg_arc's: the arc runs on without an acceptor -> k_far_linear
        unsigned int* rest_idx = c->fgq_rank.p + 2 * (size_t)n_lin + 2;   // k_fg_arc_first's: the arc goes on after its first cells -> k_fg_arc
        unsigned int* heavy_count = c->fgq_hist.p;               // (free again as well) [0] heavy, [1] arc items, [2] far-heavy, [3] rest
        CU(cudaMemsetAsync(heavy_count, 0, 4 * sizeof(unsigned), st));
        // (the cells of long arcs become work items: up to 8 per ray of the list, the rest is walked in place)
        const unsigned item_cap = n_lin > (1u << 26) ? (1u << 29) : n_lin * 8u + (1u << 21);
        CU(c->arc_items.ensure(item_cap, 0, st));
        unsigned int* n_items = heavy_count + 1;
        CU(c->arc_pre.ensure((size_t)n_lin * sizeof(ArcPre), 0, st));
        ArcPre* pre = reinterpret_cast<ArcPre*>(c->arc_pre.p);
        unsigned int* stat = c->counters.p + (any ? 16 : 20);
        if (any) {
            k_fg_arc_pre<true><<<nblk(n_lin, 128), 128, 0, st>>>(c->sc, rays, lin_idx, n_lin, pre);
            k_fg_arc_first<true><<<nblk(n_lin, ARC_FIRST_WARPS), 32 * ARC_FIRST_WARPS, 0, st>>>(c->sc, rays, res, lin_idx, n_lin, stat, pre, rest_idx, heavy_count + 3);
            k_lin_near<true><<<nblk(n_lin, 128), 128, 0, st>>>(c->sc, rays, res, lin_idx, n_lin, c->counters.p + 24, heavy_idx, heavy_count);
        } else {
            k_fg_arc_pre<false><<<nblk(n_lin, 128), 128, 0, st>>>(c->sc, rays, lin_idx, n_lin, pre);
            k_fg_arc_first<false><<<nblk(n_lin, ARC_FIRST_WARPS), 32 * ARC_FIRST_WARPS, 0, st>>>(c->sc, rays, res, lin_idx, n_lin, stat, pre, rest_idx, heavy_count + 3);
            k_lin_near<false><<<nblk(n_lin, 128), 128, 0, st>>>(c->sc, rays, res, lin_idx, n_lin, c->counters.p + 26, heavy_idx, heavy_count);
        }
        c->launches += 3;
        unsigned hc2[4] = { 0u, 0u, 0u, 0u };
        CU(cudaMemcpyAsync(hc2, heavy_count, sizeof hc2, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        c->syncs++;
        const unsigned n_rest = hc2[3];
        if (n_rest) {
            if (any) k_fg_arc<true><<<nblk(n_rest, ARC_WARPS), 32 * ARC_WARPS, 0, st>>>(c->sc, rays, res, lin_idx, rest_idx, n_rest, stat, pre,
                                                                              c->arc_items.p, n_items, item_cap, farheavy_idx, heavy_count + 2, c->arc_max_cells);
            else k_fg_arc<false><<<nblk(n_rest, ARC_WARPS), 32 * ARC_WARPS, 0, st>>>(c->sc, rays, res, lin_idx, rest_idx, n_rest, stat, pre,
                                                                               c->arc_items.p, n_items, item_cap, farheavy_idx, heavy_count + 2, c->arc_max_cells);
            c->launches++;
            CU(cudaMemcpyAsync(hc2 + 1, heavy_count + 1, 2 * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            c->syncs++;
        }
        const unsigned n_it = hc2[1] < item_cap ? hc2[1] : item_cap;
        if (getenv("RT580_DEBUG_TIMING")) fprintf(stderr, "[rt580] %s flush of %u: %u from outside after the first cell; %u walk on, %u to k_slow, %u arc items, %u to k_far_linear\n", any ? "any-hit" : "closest-hit", n, n_lin, n_rest, hc2[0], hc2[1], hc2[2]);
        if (n_it) {
            if (any) k_fg_arc_items<true><<<nblk(n_it, ARC_WARPS), 32 * ARC_WARPS, 0, st>>>(c->sc, rays, res, c->arc_items.p, n_it);
            else k_fg_arc_items<false><<<nblk(n_it, ARC_WARPS), 32 * ARC_WARPS, 0, st>>>(c->sc, rays, res, c->arc_items.p, n_it);
            c->launches++;
        }
        if (hc2[2]) {
            const unsigned nf = hc2[2];
            const unsigned batches = nblk(nf, SLOW_RPB);
            unsigned slices = nblk(8u * (unsigned)c->prop.multiProcessorCount, batches);
            const unsigned max_slices = nblk((unsigned)c->sc.n_all, SLOW_TILE);
            if (slices > max_slices) slices = max_slices;
            if (slices > 1024u) slices = 1024u;
            if (slices < 1u) slices = 1u;
            const int chunk = (int)(nblk(nblk((unsigned)c->sc.n_all, slices), SLOW_TILE) * SLOW_TILE);
            const dim3 grid(batches, slices);
            if (any) k_far_linear<true><<<grid, 256, 0, st>>>(c->sc, rays, nf, res, chunk, farheavy_idx);
            else k_far_linear<false><<<grid, 256, 0, st>>>(c->sc, rays, nf, res, chunk, farheavy_idx);
            c->launches++;
        }
        n_lin = hc2[0];
        lin_idx = heavy_idx;
    }
    if (n_lin) {
        const unsigned batches = nblk(n_lin, SLOW_RPB);
        unsigned slices = nblk(8u * (unsigned)c->prop.multiProcessorCount, batches);
        const unsigned max_slices = nblk((unsigned)c->sc.n_all, SLOW_TILE);
        if (slices > max_slices) slices = max_slices;
        if (slices > 1024u) slices = 1024u;
        if (slices < 1u) slices = 1u;
        const int chunk = (int)(nblk(nblk((unsigned)c->sc.n_all, slices), SLOW_TILE) * SLOW_TILE);
        const dim3 grid(batches, slices);
        if (any) k_slow<true><<<grid, 256, 0, st>>>(c->sc, rays, n_lin, res, chunk, lin_idx);
        else k_slow<false><<<grid, 256, 0, st>>>(c->sc, rays, n_lin, res, chunk, lin_idx);
        c->launches++;
    }
    if (tm_end(c, st)) return RT580_FAILURE;
    c->slow_total += n;
    return RT580_SUCCESS;
}
// Closest-hit queue: `n` rays were recorded (counter [2], read by the caller).
static int slow_run(rt580_context* c, bool any, unsigned cap, unsigned n, unsigned* n_out)
{
    if (n > cap) n = cap;
    *n_out = n;
    return slow_launch(c, any, c->slow_rays.p, c->slow_res.p, n);
}

// A scene "leaks" when a noticeable share of its rays needs the slow path (open scenes: every ray that
// escapes).  Then the deferred queue must be able to take every ray of a launch; otherwise a small
// queue, flushed once, is enough (should it overflow, the pass is repeated in the leaky form).
static bool is_leaky(const rt580_context* c, unsigned long long rays_so_far) {
    if (c->force_leaky) return true;
    const unsigned long long thr = rays_so_far / 10000ull;
    return c->slow_seen > (thr > 4096ull ? thr : 4096ull);
}

// Flush the any-hit queue: answer its rays (k_slow) and hand the answers to `finish(n)`.
template <typename Fin>
static int any_flush(rt580_context* c, Fin finish)
{
    if (!c->any_cap) return RT580_SUCCESS;
    unsigned cnt[N_COUNTERS];
    if (read_counters(c, cnt)) return RT580_FAILURE;
    if (cnt[3] > c->any_cap) return RT580_INTERNAL_OVERFLOW;      // rays were dropped: the caller repeats the pass
    const unsigned n = cnt[3];
    if (n) {
        if (slow_launch(c, true, c->any_rays.p, c->any_res.p, n)) return RT580_FAILURE;
        finish(n);
        c->launches++;
        CU(cudaMemsetAsync(c->counters.p + 3, 0, sizeof(unsigned), c->stream));
    }
    return RT580_SUCCESS;
}

// Any-hit pass in wavefront form over `total` rays, in chunks: `gen(first, n)` launches a kernel that
// writes the chunk's rays into c->arays and their number into counters[6]; k_anyhit (persistent, lanes
// refill) adds 1 to hits[ray id] for every occluded ray and defers what the tree cannot answer to the
// any-hit queue (prepared by the caller).  No host round trip between the chunks unless the scene leaks
// and the pass may flush (`ao`: the answers only add to hits[], k_ao_finish).
template <typename Gen>
static int anyhit_queue_pass(rt580_context* c, cudaStream_t st, int lane, unsigned long long total, uint32_t* hits, unsigned id_offset,
                             unsigned pending_mark, bool leaky, bool ao, Gen gen)
{
    DBuf<ARay>& rays = lane ? c->arays2 : c->arays;
    unsigned int* ctr = c->counters.p + (lane ? 13 : 6);           // [0] rays emitted, [1] rays fetched
    const unsigned long long chunk = (unsigned long long)AH_CHUNK_TIGHT;
    CU(rays.ensure((size_t)(total < chunk ? total : chunk), 0, st));
    const unsigned blocks = (unsigned)c->prop.multiProcessorCount * (unsigned)c->ah_blocks_per_sm;
    for (unsigned long long first = 0; first < total; first += chunk) {
        const unsigned n = (unsigned)((total - first) < chunk ? (total - first) : chunk);
        CU(cudaMemsetAsync(ctr, 0, 2 * sizeof(unsigned), st));
        if (tm_begin(c, ao ? RT580_CLASS_AO_GEN : RT580_CLASS_SHADOW_GEN, st)) return RT580_FAILURE;
        gen(first, n, rays.p, ctr);
        if (tm_end(c, st)) return RT580_FAILURE;
        c->launches++;
        if (tm_begin(c, ao ? RT580_CLASS_AO_TREE : RT580_CLASS_SHADOW_TREE, st)) return RT580_FAILURE;
        const int ah_steps = leaky ? c->ah_steps_leaky : c->ah_steps, ah_min = leaky ? c->ah_min_search_leaky : c->ah_min_search;
        if (c->count_visits)
            k_anyhit<true><<<blocks, 128, 0, st>>>(c->sc, rays.p, ctr, ctr + 1, hits, slowq_any(c), id_offset,
                                                   pending_mark, reinterpret_cast<unsigned long long*>(c->counters.p + (ao ? 8 : 10)),
                                                   ah_steps, ah_min, c->ah_batch_div, c->visit_counts.p);
        else
            k_anyhit<false><<<blocks, 128, 0, st>>>(c->sc, rays.p, ctr, ctr + 1, hits, slowq_any(c), id_offset,
                                                    pending_mark, reinterpret_cast<unsigned long long*>(c->counters.p + (ao ? 8 : 10)),
                                                    ah_steps, ah_min, c->ah_batch_div, nullptr);
        if (tm_end(c, st)) return RT580_FAILURE;
        c->launches++;
        const unsigned long long rest = total - first - n;
        if (leaky && ao && c->any_cap && rest && (unsigned long long)c->any_cap < total) {
            // the queue must be able to take every ray of the next chunk (a queue that holds the whole pass needs no check)
            unsigned cnt[N_COUNTERS];
            if (read_counters(c, cnt)) return RT580_FAILURE;
            const unsigned long long next_n = rest < chunk ? rest : chunk;
            if (cnt[3] && (unsigned long long)cnt[3] + next_n > c->any_cap) {
                const unsigned q = cnt[3] > c->any_cap ? c->any_cap : cnt[3];
                if (slow_launch(c, true, c->any_rays.p, c->any_res.p, q)) return RT580_FAILURE;
                k_ao_finish<<<nblk(q, 256), 256, 0, st>>>(c->any_rays.p, c->any_res.p, q, hits); c->launches++;
                CU(cudaMemsetAsync(c->counters.p + 3, 0, sizeof(unsigned), st));
            }
        }
    }
    CU(cudaGetLastError());
    return RT580_SUCCESS;
}

static int render_begin_impl(rt580_context* c, const rt580_render_params* p, uint64_t* row_hit_nodes);
extern "C" int rt580_render_begin(rt580_context* c, const rt580_render_params* p, uint64_t* row_hit_nodes)
{
    int rc = render_begin_impl(c, p, row_hit_nodes);
    if (rc == RT580_INTERNAL_OVERFLOW) {
        // the scene leaks after all: once more, with a deferred queue that takes every shadow ray of a level
        c->force_leaky = true;
        rc = render_begin_impl(c, p, row_hit_nodes);
        if (rc == RT580_INTERNAL_OVERFLOW) FAIL(RT580_FAILURE, "rt580_render_begin: deferred-ray queue overflow");
    }
    return rc;
}
static int render_begin_impl(rt580_context* c, const rt580_render_params* p, uint64_t* row_hit_nodes)
{
    if (!c || !p) FAIL(RT580_INVALID_ARG, "rt580_render_begin: NULL argument");
    if (!c->have_scene) FAIL(RT580_FAILURE, "rt580_render_begin: no scene uploaded");
    if (p->width <= 0 || p->height <= 0) FAIL(RT580_INVALID_ARG, "rt580_render_begin: bad resolution %dx%d", p->width, p->height);
    if (p->depth < 0 || p->depth > 4) FAIL(RT580_INVALID_ARG, "rt580_render_begin: depth %d outside [0,4] (reference: 4, h:563)", p->depth);
    if (p->ao_spp < 1 || p->ao_spp > 65536) FAIL(RT580_INVALID_ARG, "rt580_render_begin: ao_spp %d outside [1,65536]", p->ao_spp);
    if (p->rng_mode != RT580_RNG_REFERENCE_LCG && p->rng_mode != RT580_RNG_COUNTER) FAIL(RT580_INVALID_ARG, "rt580_render_begin: bad rng_mode");
    if (p->traversal < 0 || p->traversal > 2) FAIL(RT580_INVALID_ARG, "rt580_render_begin: bad traversal");
    {
        // the far-field bounds of the scene build hold for ray origins inside the scene's origin box and at the camera
        // the scene was uploaded with
        bool same = true, inside = true;
        for (int k = 0; k < 3; k++) {
            same = same && p->camera_from[k] == c->sc.ob_cam[k];
            inside = inside && p->camera_from[k] >= c->sc.ob_lo[k] && p->camera_from[k] <= c->sc.ob_hi[k];
        }
        if (!same && !inside)
            FAIL(RT580_INVALID_ARG, "rt580_render_begin: camera_from (%g, %g, %g) is neither the rt580_flat_scene::origin_hint the scene was "
                 "uploaded with nor inside the scene's bounds; upload the scene with this camera as origin_hint",
                 p->camera_from[0], p->camera_from[1], p->camera_from[2]);
    }
    CU(cudaSetDevice(c->device));
    if (sync_side(c)) return RT580_FAILURE;        // (idle unless an earlier frame was abandoned on an error)
    const bool dbg_t = getenv("RT580_DEBUG_TIMING") != nullptr;
    auto now_ms = []() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    const double t_enter = now_ms();
    FrameParams& fp = c->fp;
    fp.W = p->width; fp.H = p->height;
    if (p->n_rows == 0) { fp.row_first = 0; fp.row_step = 1; fp.n_rows = p->height; }
    else if (p->n_rows < 0) { fp.row_first = 0; fp.row_step = p->row_step > 0 ? p->row_step : 1; fp.n_rows = 0; }     // this context renders no row (more ranks than rows)
    else { fp.row_first = p->row_first; fp.row_step = p->row_step; fp.n_rows = p->n_rows; }
    if (fp.n_rows < 0 || fp.row_step < 1 || fp.row_first < 0 || (fp.n_rows > 0 && fp.row_first + (long long)(fp.n_rows - 1) * fp.row_step >= fp.H))
        FAIL(RT580_INVALID_ARG, "rt580_render_begin: rows (first %d step %d count %d) outside the %d-row frame", fp.row_first, fp.row_step, fp.n_rows, fp.H);
    const unsigned long long npix64 = (unsigned long long)fp.n_rows * fp.W;
    if (npix64 > 0x7fffff00ull) FAIL(RT580_INVALID_ARG, "rt580_render_begin: too many pixels for one context");
    const unsigned npix = (unsigned)npix64;
    fp.depth = p->depth; fp.spp = p->ao_spp; fp.rng_mode = p->rng_mode;
    fp.spp_shift = -1;
    for (int b = 0; b < 17; b++) if (fp.spp == (1 << b)) fp.spp_shift = b;
    memcpy(fp.cam, p->camera_from, sizeof fp.cam);
    memcpy(fp.inv, p->inv_view3x3, sizeof fp.inv);
    c->traversal = p->traversal;
    if (p->farfield != RT580_FARFIELD_EXACT && p->farfield != RT580_FARFIELD_OFF) FAIL(RT580_INVALID_ARG, "rt580_render_begin: bad farfield");
    c->sc.farfield = (p->farfield == RT580_FARFIELD_EXACT) ? 1 : 0;
    const int mode = pick_mode(c, p->traversal);
    cudaStream_t st = c->stream;
    c->launches = 0; c->syncs = 0;
    memset(&c->stats, 0, sizeof c->stats);
    memset(&c->prof, 0, sizeof c->prof);
    c->tm_n = 0;
    CU(c->visit_counts.ensure(4, 0, st));
    CU(cudaMemsetAsync(c->visit_counts.p, 0, 4 * sizeof(unsigned long long), st));

    // primary-ray tables: cpp:834-846 evaluated in double on the host, exactly as the reference
    // does per pixel (tan is libm's; hoisting is bit-exact because it is a pure function of x / y);
    // kept across frames of the same resolution
    if (c->ndc_w != fp.W || c->ndc_h != fp.H || c->ndc_fov != p->fov_degrees) {
        std::vector<float> t((size_t)fp.W + fp.H);
        const float half = p->fov_degrees / 2;
        const float rad = (float)(half * (3.14159265 / 180));                 // ToRadian h:581-583
        const float aspect = (float)fp.W / (float)fp.H;                       // cpp:836
        // the reference's unqualified tan() is the C library's DOUBLE function of the float radian (Q27);
        // in a .cu file a bare tan(float) would bind to CUDA's float overload, hence the explicit form
        const double tan_half = ::tan((double)rad);
        for (int x = 0; x < fp.W; x++) { double n = (2.0 * x) / fp.W - 1; n *= (double)aspect * tan_half; t[x] = (float)n; }
        for (int y = 0; y < fp.H; y++) { double n = 1 - (2.0 * y) / fp.H; n *= tan_half; t[(size_t)fp.W + y] = (float)n; }
        CU(c->ndc.ensure(t.size(), 0, st));
        CU(cudaMemcpyAsync(c->ndc.p, t.data(), t.size() * sizeof(float), cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));
        c->ndc_w = fp.W; c->ndc_h = fp.H; c->ndc_fov = p->fov_degrees;
    }
    fp.ndc_x = c->ndc.p; fp.ndc_y = c->ndc.p + fp.W;
    if (c->lcg_pow_spp != fp.spp) {
        // 16807^(2k) mod (2^31-1), k < spp: sample k of an AO call starts 2k engine steps after the call (cpp:320-321)
        std::vector<uint32_t> t((size_t)fp.spp);
        const uint64_t m = 2147483647ull, a2 = (16807ull * 16807ull) % m;
        uint64_t v = 1;
        for (int k = 0; k < fp.spp; k++) { t[k] = (uint32_t)v; v = (v * a2) % m; }
        CU(c->lcg_pow.ensure(t.size(), 0, st));
        CU(cudaMemcpyAsync(c->lcg_pow.p, t.data(), t.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));
        c->lcg_pow_spp = fp.spp;
    }
    fp.lcg_pow = c->lcg_pow.p;
    if (!c->lcg_tab.p) {
        std::vector<uint32_t> t(1024);
        const uint64_t m = 2147483647ull;
        uint64_t base = 16807ull;                                  // 16807^(256^k)
        for (int k = 0; k < 4; k++) {
            uint64_t v = 1;
            for (int d = 0; d < 256; d++) { t[(size_t)k * 256 + d] = (uint32_t)v; v = (v * base) % m; }
            base = v;                                              // = base^256
        }
        CU(c->lcg_tab.ensure(t.size(), 0, st));
        CU(cudaMemcpyAsync(c->lcg_tab.p, t.data(), t.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));
    }
    fp.lcg_tab = c->lcg_tab.p;
    CU(c->counters.ensure(N_COUNTERS, 0, st));
    CU(cudaMemsetAsync(c->counters.p, 0, N_COUNTERS * sizeof(unsigned), st));
    c->sc.diag = c->counters.p + 4;
    c->slow_seen = 0; c->any_cap = 0;
    CU(c->pix_hits.ensure(npix + 1, 0, st));
    CU(c->pix_scan.ensure(npix + 1, 0, st));
    CU(c->fb.ensure((size_t)npix * 3 + 1, 0, st));
    CU(c->nodes.ensure(npix + 1, 0, st));
    CU(c->aux.ensure(npix + 1, 0, st));
    CU(c->row_vals.ensure((size_t)fp.n_rows + 1, 0, st));

    const double t_setup = now_ms();
    CU(cudaEventRecord(c->ev[0], st));
    c->level_off.clear(); c->level_rays.clear();
    c->level_off.push_back(0);
    unsigned n_nodes = 0, slow_cap = 0, n_slow = 0;
    unsigned cnt[N_COUNTERS];
    c->slow_total = 0;
    unsigned long long rays_so_far = npix;
    bool any_open = false;      // shadow rays of earlier levels wait in the any-hit queue
    auto shadow_finish = [&](unsigned n) {
        k_shadow_finish<<<nblk(n, 128), 128, 0, st>>>(c->sc, c->fp, c->any_rays.p, c->any_res.p, n, c->nodes.p, c->aux.p);
    };
    const Spawn no_spawn = { nullptr, nullptr };
    if (mode == 0) {
        // ---- wavefront path over the LBVH -------------------------------------------------------------------
        // The kernel that creates a hit node also spawns its children into the NEXT level's queue (two queues,
        // two counters, alternating), so one host read after a level's closest-hit step returns both the new
        // node count and the next queue's length.
        DBuf<QRay>* Q[2] = { &c->queue, &c->queue2 };
        unsigned int* qcnt[2] = { c->counters.p + 1, c->counters.p + 15 };
        const int qidx[2] = { 1, 15 };
        int cur = 0;
        unsigned q = 0;
        if (npix) {
            CU(Q[0]->ensure(2 * (size_t)npix + 1, 0, st));
            if (slow_prepare(c, npix, &slow_cap)) return RT580_FAILURE;
            if (tm_begin(c, RT580_CLASS_PRIMARY, st)) return RT580_FAILURE;
            launch_trace<0>(c, true, nullptr, npix, nullptr, (unsigned)c->nodes.cap, slow_cap, Spawn{ Q[0]->p, qcnt[0] });
            if (tm_end(c, st)) return RT580_FAILURE;
            if (read_counters(c, cnt)) return RT580_FAILURE;
            n_nodes = cnt[0]; q = cnt[qidx[0]];
            if (slow_run(c, false, slow_cap, cnt[2], &n_slow)) return RT580_FAILURE;
            if (n_slow) {
                k_trace_finish<true><<<nblk(n_slow, 128), 128, 0, st>>>(c->sc, c->fp, nullptr, c->slow_rays.p, c->slow_res.p, n_slow,
                                                                       c->nodes.p, c->aux.p, c->counters.p, c->pix_hits.p, c->fb.p,
                                                                       (unsigned)c->nodes.cap, Spawn{ Q[0]->p, qcnt[0] });
                c->launches++;
                if (read_counters(c, cnt)) return RT580_FAILURE;
                n_nodes = cnt[0]; q = cnt[qidx[0]];
            }
        }
        c->level_rays.push_back(npix);
        c->level_off.push_back(n_nodes);
        for (int L = 0; L <= fp.depth; L++) {
            const unsigned n0 = (unsigned)c->level_off[L], n1 = (unsigned)c->level_off[L + 1];
            if (n1 == n0) break;
            const unsigned long long n_sh = (unsigned long long)(n1 - n0) * (unsigned)c->sc.n_nonambient;
            if (n_sh > 0xfffffff0ull) FAIL(RT580_FAILURE, "rt580_render_begin: too many shadow rays in one level");
            const bool more = L < fp.depth && q > 0;
            const bool leaky = is_leaky(c, rays_so_far);
            // Two chains per level.  Critical path, on the context's stream: the closest hits of the level's children
            // and with them the next level's nodes and rays.  Beside it, on a side stream: the level's shadow rays
            // (k_shade_gen -> k_anyhit) and the Phong terms they gate; nothing needs those before the resolve pass.
            // Alone, each persistent traversal kernel ends in a tail that leaves most of the GPU idle, which is what
            // a rank of an 8-GPU run (1/8 of the rays per launch) spent a quarter of its structure pass on.
            const int lane = (c->overlap && !leaky) ? (L & 1) : 0;
            cudaStream_t sb = (c->overlap && !leaky) ? c->side[lane] : st;
            DBuf<uint32_t>& occl = lane ? c->occl2 : c->occl;
            if (sb == st) {
                // one stream from here on: whatever the side streams still do for earlier levels comes first
                if (join_side(c)) return RT580_FAILURE;
            }
            if (leaky && any_open) {
                // the small queue of the levels before is flushed before the big one takes over
                const int fr = any_flush(c, shadow_finish); if (fr) return fr; any_open = false;
            }
            if (!leaky && !any_open) {
                // the small any-hit queue of the frame, shared by both side streams
                if (any_prepare(c, c->slow_any_cap, st)) return RT580_FAILURE;
                any_open = c->any_cap != 0;
            }
            if (sb != st) {
                // everything the context's stream has done so far (the level's nodes) comes first
                CU(cudaEventRecord(c->ev_level, st)); CU(cudaStreamWaitEvent(sb, c->ev_level, 0));
            }
            if (more) {
                // (a buffer that has to move must not be in use on a side stream)
                if ((size_t)n1 + q > c->nodes.cap || (size_t)n1 + q > c->aux.cap) { if (sync_side(c)) return RT580_FAILURE; }
                CU(c->nodes.ensure((size_t)n1 + q, n1, st));
                CU(c->aux.ensure((size_t)n1 + q, n1, st));
                CU(Q[cur ^ 1]->ensure(2 * (size_t)q + 1, 0, st));
                CU(cudaMemsetAsync(qcnt[cur ^ 1], 0, sizeof(unsigned), st));
                if (slow_prepare(c, q, &slow_cap)) return RT580_FAILURE;
                const Spawn sp = { Q[cur ^ 1]->p, qcnt[cur ^ 1] };
                if (!c->one_thread_per_ray) {
                    CU(c->chits.ensure((size_t)q + 1, 0, st));
                    CU(cudaMemsetAsync(c->counters.p + 12, 0, sizeof(unsigned), st));
                    const unsigned blocks = (unsigned)c->prop.multiProcessorCount * (unsigned)c->ch_blocks_per_sm;
                    if (tm_begin(c, RT580_CLASS_CLOSEST, st)) return RT580_FAILURE;
                    if (c->count_visits)
                        k_closest<true><<<blocks, 128, 0, st>>>(c->sc, Q[cur]->p, qcnt[cur], q, c->counters.p + 12, c->chits.p,
                                                                c->ah_steps, c->ah_min_search, c->ah_batch_div, c->visit_counts.p + 2);
                    else
                        k_closest<false><<<blocks, 128, 0, st>>>(c->sc, Q[cur]->p, qcnt[cur], q, c->counters.p + 12, c->chits.p,
                                                                 c->ah_steps, c->ah_min_search, c->ah_batch_div, nullptr);
                    k_commit<<<nblk(q, 128), 128, 0, st>>>(c->sc, Q[cur]->p, q, nullptr, c->chits.p, c->nodes.p, c->aux.p,
                                                           c->counters.p, c->pix_hits.p, c->fb.p, (unsigned)c->nodes.cap, slowq(c, slow_cap), sp);
                    if (tm_end(c, st)) return RT580_FAILURE;
                    c->launches += 2;
                } else {
                    launch_trace<0>(c, false, Q[cur]->p, q, nullptr, (unsigned)c->nodes.cap, slow_cap, sp);
                }
            }
            CU(occl.ensure((size_t)n_sh + 1, 0, sb));
            CU(cudaMemsetAsync(occl.p, 0, sizeof(uint32_t) * ((size_t)n_sh + 1), sb));
            if (n_sh) {
                if (leaky) {
                    // the queue takes every shadow ray of the level and is flushed right after the level
                    if (any_prepare(c, n_sh, sb)) return RT580_FAILURE;
                }
                const int rc = anyhit_queue_pass(c, sb, lane, n_sh, occl.p, n0 * (unsigned)c->sc.n_nonambient, OCCL_PENDING, leaky, false,
                    [&](unsigned long long first, unsigned n, ARay* rays, unsigned int* ctr) {
                        k_shade_gen<<<nblk(n, 512), 512, 0, sb>>>(c->sc, n0, n1 - n0, first, n, c->nodes.p, rays, ctr, occl.p);
                    });
                if (rc) return rc;
            }
            if (tm_begin(c, RT580_CLASS_SHADOW_GEN, sb)) return RT580_FAILURE;
            k_shade_local<<<nblk(n1 - n0, 128), 128, 0, sb>>>(c->sc, c->fp, n0, n1, c->nodes.p, c->aux.p, occl.p);
            if (tm_end(c, sb)) return RT580_FAILURE;
            c->launches++;
            if (leaky && n_sh) { const int fr = any_flush(c, shadow_finish); if (fr) return fr; }
            rays_so_far += n_sh;
            if (!more) break;
            if (read_counters(c, cnt)) return RT580_FAILURE;
            n_nodes = cnt[0];
            unsigned q_next = cnt[qidx[cur ^ 1]];
            if (slow_run(c, false, slow_cap, cnt[2], &n_slow)) return RT580_FAILURE;
            if (n_slow) {
                k_trace_finish<false><<<nblk(n_slow, 128), 128, 0, st>>>(c->sc, c->fp, Q[cur]->p, c->slow_rays.p, c->slow_res.p, n_slow,
                                                                        c->nodes.p, c->aux.p, c->counters.p, c->pix_hits.p, c->fb.p,
                                                                        (unsigned)c->nodes.cap, Spawn{ Q[cur ^ 1]->p, qcnt[cur ^ 1] });
                c->launches++;
                if (read_counters(c, cnt)) return RT580_FAILURE;
                n_nodes = cnt[0]; q_next = cnt[qidx[cur ^ 1]];
            }
            c->level_rays.push_back(q);
            c->level_off.push_back(n_nodes);
            rays_so_far += q;
            q = q_next; cur ^= 1;
        }
    } else {
        // ---- tiny scenes / checker: the reference's linear loop, one thread per ray, k_shade spawns the children ----
        if (npix) {
            if (slow_prepare(c, npix, &slow_cap)) return RT580_FAILURE;
            DISPATCH_MODE(mode, launch_trace, c, true, nullptr, npix, nullptr, (unsigned)c->nodes.cap, slow_cap, no_spawn);
            if (read_counters(c, cnt)) return RT580_FAILURE;
            n_nodes = cnt[0];
            if (slow_run(c, false, slow_cap, cnt[2], &n_slow)) return RT580_FAILURE;
            if (n_slow) {
                k_trace_finish<true><<<nblk(n_slow, 128), 128, 0, st>>>(c->sc, c->fp, c->queue.p, c->slow_rays.p, c->slow_res.p, n_slow,
                                                                       c->nodes.p, c->aux.p, c->counters.p, c->pix_hits.p, c->fb.p,
                                                                       (unsigned)c->nodes.cap, no_spawn);
                c->launches++;
                if (read_counters(c, cnt)) return RT580_FAILURE;
                n_nodes = cnt[0];
            }
        }
        c->level_rays.push_back(npix);
        c->level_off.push_back(n_nodes);
        for (int L = 0; L <= fp.depth; L++) {
            const unsigned n0 = (unsigned)c->level_off[L], n1 = (unsigned)c->level_off[L + 1];
            if (n1 == n0) break;
            const unsigned q_max = (L < fp.depth) ? 2u * (n1 - n0) : 0u;       // every node spawns at most two rays
            CU(c->queue.ensure((size_t)q_max + 1, 0, st));
            if (q_max) {
                CU(c->nodes.ensure((size_t)n1 + q_max, n1, st));
                CU(c->aux.ensure((size_t)n1 + q_max, n1, st));
            }
            CU(cudaMemsetAsync(c->counters.p + 1, 0, sizeof(unsigned), st));
            const unsigned long long n_sh = (unsigned long long)(n1 - n0) * (unsigned)c->sc.n_nonambient;
            if (slow_prepare(c, n_sh, &slow_cap)) return RT580_FAILURE;
            DISPATCH_MODE(mode, launch_shade, c, n0, n1, slow_cap);
            if (slow_cap) {
                if (read_counters(c, cnt)) return RT580_FAILURE;
                if (slow_run(c, true, slow_cap, cnt[2], &n_slow)) return RT580_FAILURE;
                if (n_slow) {
                    k_shade_finish<<<nblk(n_slow, 128), 128, 0, st>>>(c->sc, c->fp, c->slow_rays.p, c->slow_res.p, n_slow, c->nodes.p, c->aux.p);
                    c->launches++;
                }
            }
            rays_so_far += n_sh;
            if (L == fp.depth) break;
            // the queued reflection / refraction rays: their number stays on the device (counters[1])
            if (slow_prepare(c, q_max, &slow_cap)) return RT580_FAILURE;
            DISPATCH_MODE(mode, launch_trace, c, false, c->queue.p, q_max, c->counters.p + 1, (unsigned)c->nodes.cap, slow_cap, no_spawn);
            if (read_counters(c, cnt)) return RT580_FAILURE;
            const unsigned q = cnt[1];
            if (q == 0) break;
            n_nodes = cnt[0];
            if (slow_run(c, false, slow_cap, cnt[2], &n_slow)) return RT580_FAILURE;
            if (n_slow) {
                k_trace_finish<false><<<nblk(n_slow, 128), 128, 0, st>>>(c->sc, c->fp, c->queue.p, c->slow_rays.p, c->slow_res.p, n_slow,
                                                                        c->nodes.p, c->aux.p, c->counters.p, c->pix_hits.p, c->fb.p,
                                                                        (unsigned)c->nodes.cap, no_spawn);
                c->launches++;
                if (read_counters(c, cnt)) return RT580_FAILURE;
                n_nodes = cnt[0];
            }
            c->level_rays.push_back(q);
            c->level_off.push_back(n_nodes);
            rays_so_far += q;
        }
    }
    c->rays_structure = rays_so_far;
    const double t_struct = now_ms();
    CU(cudaEventRecord(c->ev[1], st));
    // order: subtree sizes bottom-up, per-pixel exclusive scan, per-row totals
    if (tm_begin(c, RT580_CLASS_ORDER, st)) return RT580_FAILURE;
    const int n_levels = (int)c->level_off.size() - 1;
    for (int L = n_levels - 1; L >= 0; L--) {
        const unsigned n0 = (unsigned)c->level_off[L], n1 = (unsigned)c->level_off[L + 1];
        if (n1 > n0) { k_subtree<<<nblk(n1 - n0, 256), 256, 0, st>>>(n0, n1, c->nodes.p, c->aux.p, c->pix_hits.p); c->launches++; }
    }
    if (npix) {
        if (exclusive_scan_u32(c, c->pix_hits.p, c->pix_scan.p, npix)) return RT580_FAILURE;
        k_row_counts<<<nblk(fp.n_rows, 128), 128, 0, st>>>(c->pix_hits.p, c->pix_scan.p, fp.W, fp.n_rows, c->row_vals.p); c->launches++;
    }
    if (tm_end(c, st)) return RT580_FAILURE;
    CU(cudaEventRecord(c->ev[7], st));
    // The side streams' work (shadow rays and Phong terms of the last levels) belongs to the structure pass, but the
    // order kernels above need none of it (they read the subtree sizes, the side streams write NodeAux::local): the
    // context's stream waits for it only now, then the deferred shadow rays are answered.
    if (join_side(c)) return RT580_FAILURE;
    if (any_open) { const int fr = any_flush(c, shadow_finish); if (fr) return fr; }
    c->any_cap = 0;
    CU(cudaGetLastError());
    CU(cudaEventRecord(c->ev[6], st));
    if (row_hit_nodes) {
        std::vector<uint64_t> rows((size_t)fp.n_rows);
        if (fp.n_rows) CU(cudaMemcpyAsync(rows.data(), c->row_vals.p, sizeof(uint64_t) * fp.n_rows, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        c->syncs++;
        CU(cudaGetLastError());
        for (int r = 0; r < fp.n_rows; r++) row_hit_nodes[r] = rows[r];
    }
    c->frame_begun = true; c->fb_pixels = 0;
    if (dbg_t) fprintf(stderr, "[rt580] render_begin host ms: setup %.2f structure %.2f order %.2f (levels %d, launches %u, host syncs %u)\n",
                       t_setup - t_enter, t_struct - t_setup, now_ms() - t_struct, (int)c->level_off.size() - 1, c->launches, c->syncs);
    // ray accounting: one ray == one IntersectScene call of the reference (cpp:30, cpp:75, cpp:325)
    c->stats.rays_primary = npix;
    for (size_t l = 1; l < c->level_rays.size(); l++) c->stats.rays_secondary += c->level_rays[l];
    c->stats.hit_nodes = n_nodes;
    c->stats.rays_shadow = (uint64_t)n_nodes * c->sc.n_nonambient;
    c->stats.ao_calls = (uint64_t)n_nodes * c->sc.n_ambient;
    c->stats.rays_ao = c->stats.ao_calls * (uint64_t)fp.spp;
    c->stats.bvh_max_depth = c->bvh_depth;
    return RT580_SUCCESS;
}

static int render_finish_impl(rt580_context* c, const uint64_t* row_ao_base, bool bases_on_device, int16_t* fb_out,
                              int fb_on_device, rt580_stats* stats);
extern "C" int rt580_render_finish(rt580_context* c, const uint64_t* row_ao_base, int16_t* fb_out, int fb_on_device,
                                   rt580_stats* stats)
{
    if (!c) FAIL(RT580_INVALID_ARG, "rt580_render_finish: ctx is NULL");
    if (!c->frame_begun) FAIL(RT580_FAILURE, "rt580_render_finish: no frame begun");
    if (c->fp.rng_mode == RT580_RNG_REFERENCE_LCG && !row_ao_base && c->fp.n_rows)
        FAIL(RT580_INVALID_ARG, "rt580_render_finish: row_ao_base required in RT580_RNG_REFERENCE_LCG mode");
    return render_finish_impl(c, row_ao_base, false, fb_out, fb_on_device, stats);
}

extern "C" int rt580_row_counts_to_device(rt580_context* c, uint64_t* dst_device, int32_t max_rows)
{
    if (!c || !dst_device) FAIL(RT580_INVALID_ARG, "rt580_row_counts_to_device: NULL argument");
    if (!c->frame_begun) FAIL(RT580_FAILURE, "rt580_row_counts_to_device: no frame begun");
    if (max_rows < c->fp.n_rows) FAIL(RT580_INVALID_ARG, "rt580_row_counts_to_device: max_rows %d < %d rows of this context", max_rows, c->fp.n_rows);
    CU(cudaSetDevice(c->device));
    if (c->fp.n_rows)
        CU(cudaMemcpyAsync(dst_device, c->row_vals.p, sizeof(uint64_t) * c->fp.n_rows, cudaMemcpyDeviceToDevice, c->stream));
    if (max_rows > c->fp.n_rows)
        CU(cudaMemsetAsync(dst_device + c->fp.n_rows, 0, sizeof(uint64_t) * (size_t)(max_rows - c->fp.n_rows), c->stream));
    return RT580_SUCCESS;
}

extern "C" int rt580_render_finish_interleaved(rt580_context* c, const uint64_t* all_counts_device, int32_t world, int32_t rank,
                                               int32_t max_rows, int16_t* fb_out, int fb_on_device, rt580_stats* stats)
{
    if (!c || !all_counts_device) FAIL(RT580_INVALID_ARG, "rt580_render_finish_interleaved: NULL argument");
    if (!c->frame_begun) FAIL(RT580_FAILURE, "rt580_render_finish_interleaved: no frame begun");
    const FrameParams& fp = c->fp;
    if (world < 1 || rank < 0 || rank >= world || fp.row_step != world || (fp.n_rows && fp.row_first != rank) ||
        max_rows < (fp.H + world - 1) / world)
        FAIL(RT580_INVALID_ARG, "rt580_render_finish_interleaved: the frame begun is not rank %d's share of rows interleaved over %d ranks", rank, world);
    CU(cudaSetDevice(c->device));
    if (fp.n_rows) {
        k_row_bases_interleaved<<<1, 1024, 0, c->stream>>>(all_counts_device, world, rank, max_rows, fp.H, c->row_vals.p);
        c->launches++;
    }
    return render_finish_impl(c, nullptr, true, fb_out, fb_on_device, stats);
}

static int render_finish_impl(rt580_context* c, const uint64_t* row_ao_base, bool bases_on_device, int16_t* fb_out,
                              int fb_on_device, rt580_stats* stats)
{
    CU(cudaSetDevice(c->device));
    FrameParams& fp = c->fp;
    cudaStream_t st = c->stream;
    const int mode = pick_mode(c, c->traversal);
    const unsigned npix = (unsigned)fp.n_rows * fp.W;
    const unsigned n_nodes = (unsigned)c->level_off.back();
    const int n_amb = c->sc.n_ambient;
    const int n_levels = (int)c->level_off.size() - 1;
    if (!bases_on_device && row_ao_base && fp.n_rows)
        CU(cudaMemcpyAsync(c->row_vals.p, row_ao_base, sizeof(uint64_t) * fp.n_rows, cudaMemcpyHostToDevice, st));
    CU(cudaEventRecord(c->ev[2], st));
    const unsigned long long n_calls = (unsigned long long)n_nodes * n_amb;
    CU(c->pre.ensure((size_t)n_nodes + 1, 0, st));
    CU(c->ao_state.ensure((size_t)n_calls + 1, 0, st));
    CU(c->ao_hits.ensure((size_t)n_calls + 1, 0, st));
    CU(cudaMemsetAsync(c->ao_hits.p, 0, sizeof(uint32_t) * (n_calls + 1), st));
    if (tm_begin(c, RT580_CLASS_ORDER, st)) return RT580_FAILURE;
    for (int L = 0; L < n_levels; L++) {
        const unsigned n0 = (unsigned)c->level_off[L], n1 = (unsigned)c->level_off[L + 1];
        if (n1 > n0) {
            k_preorder<<<nblk(n1 - n0, 128), 128, 0, st>>>(n0, n1, c->nodes.p, c->aux.p, c->pix_scan.p, c->row_vals.p, fp, n_amb,
                                                          c->pre.p, c->ao_state.p);
            c->launches++;
        }
    }
    if (tm_end(c, st)) return RT580_FAILURE;
    CU(cudaEventRecord(c->ev[3], st));
    const unsigned long long n_ao = n_calls * (unsigned long long)fp.spp;
    if (n_ao > 0xffffffffull * 128ull) FAIL(RT580_FAILURE, "rt580_render_finish: AO ray count exceeds one launch");
    if (n_ao) {
        unsigned slow_cap = 0, n_slow = 0;
        CU(cudaEventRecord(c->ev[8], st));
        if (mode == 0) {
            for (int attempt = 0; ; attempt++) {
                const bool leaky = is_leaky(c, c->rays_structure);
                // leaky (open scene): one queue for the escaping rays of the whole pass - the more rays a flush sorts by direction
                // cell, the more of them share a cell's list (64 B per entry: 180 GB of HBM take it)
                if (any_prepare(c, leaky ? (n_ao < (unsigned long long)LEAKY_ANY_CAP ? n_ao : (unsigned long long)LEAKY_ANY_CAP) : c->slow_any_cap, st)) return RT580_FAILURE;
                const int rc = anyhit_queue_pass(c, st, 0, n_ao, c->ao_hits.p, 0u, 0u, leaky, true,
                    [&](unsigned long long first, unsigned n, ARay* rays, unsigned int* ctr) {
                        k_ao_gen<<<nblk(n, 256), 256, 0, st>>>(c->sc, fp, first, n, n_amb, c->nodes.p, c->ao_state.p, rays, ctr, c->ao_hits.p);
                    });
                if (rc) return rc;
                CU(cudaEventRecord(c->ev[9], st));
                const int fr = any_flush(c, [&](unsigned n) { k_ao_finish<<<nblk(n, 256), 256, 0, st>>>(c->any_rays.p, c->any_res.p, n, c->ao_hits.p); });
                if (fr == RT580_INTERNAL_OVERFLOW && attempt == 0) {
                    // the scene leaks after all: once more with a queue that takes every ray of a chunk
                    c->force_leaky = true;
                    CU(cudaMemsetAsync(c->ao_hits.p, 0, sizeof(uint32_t) * (n_calls + 1), st));
                    CU(cudaMemsetAsync(c->counters.p + 8, 0, 2 * sizeof(unsigned), st));
                    continue;
                }
                if (fr) FAIL(RT580_FAILURE, "rt580_render_finish: deferred-ray queue overflow");
                break;
            }
            c->any_cap = 0;
        } else {
            if (slow_prepare(c, n_ao, &slow_cap)) return RT580_FAILURE;
            DISPATCH_MODE(mode, launch_ao, c, n_ao, slow_cap);
            c->stats.ao_rays_traversed = n_ao;
            CU(cudaEventRecord(c->ev[9], st));
            if (slow_cap) {
                unsigned cnt[N_COUNTERS];
                if (read_counters(c, cnt)) return RT580_FAILURE;
                if (slow_run(c, true, slow_cap, cnt[2], &n_slow)) return RT580_FAILURE;
                if (n_slow) { k_ao_finish<<<nblk(n_slow, 256), 256, 0, st>>>(c->slow_rays.p, c->slow_res.p, n_slow, c->ao_hits.p); c->launches++; }
            }
        }
    }
    CU(cudaEventRecord(c->ev[4], st));
    if (tm_begin(c, RT580_CLASS_RESOLVE, st)) return RT580_FAILURE;
    for (int L = n_levels - 1; L >= 0; L--) {
        const unsigned n0 = (unsigned)c->level_off[L], n1 = (unsigned)c->level_off[L + 1];
        if (n1 > n0) {
            k_resolve<<<nblk(n1 - n0, 128), 128, 0, st>>>(c->sc, fp, n0, n1, c->nodes.p, c->aux.p, c->ao_hits.p, n_amb, c->fb.p);
            c->launches++;
        }
    }
    if (tm_end(c, st)) return RT580_FAILURE;
    CU(cudaEventRecord(c->ev[5], st));
    if (c->frame && npix) {
        // multi-GPU: this rank's rows straight into the whole frame on rank 0 (own memory or peer mapping)
        if (c->frame_w != fp.W || c->frame_h != fp.H)
            FAIL(RT580_FAILURE, "rt580_render_finish: the shared frame is %dx%d, the render %dx%d", c->frame_w, c->frame_h, fp.W, fp.H);
        const size_t row_bytes = (size_t)fp.W * 6;
        if (row_bytes % 16 == 0) {
            const unsigned units = (unsigned)(row_bytes / 16);
            k_store_band<uint4><<<nblk((unsigned long long)units * fp.n_rows, 256), 256, 0, st>>>(
                reinterpret_cast<const uint4*>(c->fb.p), reinterpret_cast<uint4*>(c->frame), units, fp.n_rows, fp.row_first, fp.row_step);
        } else {
            const unsigned units = (unsigned)(row_bytes / 2);
            k_store_band<uint16_t><<<nblk((unsigned long long)units * fp.n_rows, 256), 256, 0, st>>>(
                reinterpret_cast<const uint16_t*>(c->fb.p), reinterpret_cast<uint16_t*>(c->frame), units, fp.n_rows, fp.row_first, fp.row_step);
        }
        c->launches++;
    }
    if (fb_out && npix)
        CU(cudaMemcpyAsync(fb_out, c->fb.p, sizeof(int16_t) * 3 * (size_t)npix,
                           fb_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    unsigned cnt[N_COUNTERS];
    if (read_counters(c, cnt)) return RT580_FAILURE;
    CU(cudaGetLastError());
    float ms = 0.f;
    // structure = the main chain up to the order kernels + what the context's stream then still waited for the side
    // streams and the deferred shadow rays; order = the order kernels of both halves
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]); c->stats.ms_structure = ms;
    cudaEventElapsedTime(&ms, c->ev[7], c->ev[6]); c->stats.ms_structure += ms;
    cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]); c->stats.ms_order = ms;
    cudaEventElapsedTime(&ms, c->ev[1], c->ev[7]); c->stats.ms_order += ms;
    cudaEventElapsedTime(&ms, c->ev[3], c->ev[4]); c->stats.ms_ao = ms;
    cudaEventElapsedTime(&ms, c->ev[4], c->ev[5]); c->stats.ms_resolve = ms;
    if (n_ao) { cudaEventElapsedTime(&ms, c->ev[8], c->ev[9]); c->stats.ms_ao_kernel = ms; }
    c->stats.ms_total = c->stats.ms_structure + c->stats.ms_order + c->stats.ms_ao + c->stats.ms_resolve;
    c->stats.kernel_launches = c->launches;
    c->stats.far_scans = cnt[4]; c->stats.linear_fallbacks = cnt[5];
    if (getenv("RT580_DEBUG_TIMING"))
        fprintf(stderr, "[rt580] rays from outside the scene, far regime: any hit %u (%.1f cells, %.0f exact tests each), closest hit %u (%.1f cells, %.0f exact tests each)\n",
                cnt[16], cnt[16] ? (double)cnt[18] / cnt[16] : 0.0, cnt[16] ? (double)cnt[19] / cnt[16] : 0.0,
                cnt[20], cnt[20] ? (double)cnt[22] / cnt[20] : 0.0, cnt[20] ? (double)cnt[23] / cnt[20] : 0.0);
    if (getenv("RT580_DEBUG_TIMING"))
        fprintf(stderr, "[rt580] rays from outside the scene, near regime: any hit %u rays with > 64 exact tests (%.0f each), closest hit %u (%.0f each)\n",
                cnt[24], cnt[24] ? 64.0 * cnt[25] / cnt[24] : 0.0, cnt[26], cnt[26] ? 64.0 * cnt[27] / cnt[26] : 0.0);
    if (mode == 0) {
        c->stats.ao_rays_traversed = (uint64_t)cnt[8] | ((uint64_t)cnt[9] << 32);
        c->stats.shadow_rays_traversed = (uint64_t)cnt[10] | ((uint64_t)cnt[11] << 32);
    } else c->stats.shadow_rays_traversed = c->stats.rays_shadow;
    if (stats) *stats = c->stats;
    {
        tm_resolve(c);
        rt580_profile& pr = c->prof;
        pr.rays[RT580_CLASS_PRIMARY] = c->stats.rays_primary; pr.rays[RT580_CLASS_CLOSEST] = c->stats.rays_secondary;
        pr.rays[RT580_CLASS_SHADOW_GEN] = c->stats.rays_shadow; pr.rays[RT580_CLASS_SHADOW_TREE] = c->stats.shadow_rays_traversed;
        pr.rays[RT580_CLASS_AO_GEN] = c->stats.rays_ao; pr.rays[RT580_CLASS_AO_TREE] = c->stats.ao_rays_traversed;
        pr.rays[RT580_CLASS_ORDER] = c->stats.hit_nodes; pr.rays[RT580_CLASS_RESOLVE] = c->stats.hit_nodes;
        if (c->count_visits && c->visit_counts.p) {
            unsigned long long v[4] = { 0, 0, 0, 0 };
            CU(cudaMemcpy(v, c->visit_counts.p, sizeof v, cudaMemcpyDeviceToHost));
            pr.nodes_any = v[0]; pr.leaves_any = v[1]; pr.nodes_closest = v[2]; pr.leaves_closest = v[3];
        }
    }
    c->frame_begun = false;
    c->fb_pixels = npix;
    return RT580_SUCCESS;
}

extern "C" int rt580_set_profiling(rt580_context* c, int count_visits)
{
    if (!c) FAIL(RT580_INVALID_ARG, "rt580_set_profiling: ctx is NULL");
    c->count_visits = count_visits != 0;
    return RT580_SUCCESS;
}
extern "C" int rt580_frame_profile(rt580_context* c, rt580_profile* out)
{
    if (!c || !out) FAIL(RT580_INVALID_ARG, "rt580_frame_profile: NULL argument");
    if (c->frame_begun) FAIL(RT580_FAILURE, "rt580_frame_profile: the frame is not finished");
    *out = c->prof;
    return RT580_SUCCESS;
}

extern "C" int rt580_frame_rgb8(rt580_context* c, const uint8_t* lut256, uint8_t* rgb_out, int out_on_device)
{
    if (!c || !lut256 || !rgb_out) FAIL(RT580_INVALID_ARG, "rt580_frame_rgb8: NULL argument");
    if (c->frame_begun || !c->fb_pixels) FAIL(RT580_FAILURE, "rt580_frame_rgb8: no finished frame on the device");
    CU(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const unsigned long long n = 3ull * c->fb_pixels;
    CU(c->rgb8.ensure((size_t)n + 256, 0, st));
    CU(cudaMemcpyAsync(c->rgb8.p + n, lut256, 256, cudaMemcpyHostToDevice, st));
    uint8_t* dst = out_on_device ? rgb_out : c->rgb8.p;
    k_gamma_rgb8<<<nblk((n + 3ull) / 4ull, 256), 256, 0, st>>>(c->fb.p, n, c->rgb8.p + n, dst);
    if (!out_on_device) CU(cudaMemcpyAsync(rgb_out, c->rgb8.p, (size_t)n, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    return RT580_SUCCESS;
}

// ---- the shared frame of a multi-GPU render ------------------------------------------------
static void frame_release(rt580_context* c) {
    if (!c->frame) return;
    cudaStreamSynchronize(c->stream);
    if (c->frame_imported) cudaIpcCloseMemHandle(c->frame); else cudaFree(c->frame);
    c->frame = nullptr; c->frame_imported = false; c->frame_w = c->frame_h = 0;
}
extern "C" int rt580_frame_export(rt580_context* c, int32_t width, int32_t height, void* ipc_handle64)
{
    if (!c || width <= 0 || height <= 0) FAIL(RT580_INVALID_ARG, "rt580_frame_export: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    CU(cudaSetDevice(c->device));
    frame_release(c);
    // a dedicated cudaMalloc (not the arena): IPC handles export whole allocations
    CU(cudaMalloc((void**)&c->frame, (size_t)width * height * 3 * sizeof(int16_t)));
    c->frame_w = width; c->frame_h = height; c->frame_imported = false;
    if (ipc_handle64) {
        cudaIpcMemHandle_t h;
        CU(cudaIpcGetMemHandle(&h, c->frame));
        memcpy(ipc_handle64, &h, sizeof h);
    }
    return RT580_SUCCESS;
}
extern "C" int rt580_frame_import(rt580_context* c, const void* ipc_handle64, int32_t width, int32_t height)
{
    if (!c || !ipc_handle64 || width <= 0 || height <= 0) FAIL(RT580_INVALID_ARG, "rt580_frame_import: bad argument");
    CU(cudaSetDevice(c->device));
    frame_release(c);
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle64, sizeof h);
    void* p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->frame = (int16_t*)p; c->frame_imported = true; c->frame_w = width; c->frame_h = height;
    return RT580_SUCCESS;
}
extern "C" int rt580_frame_release(rt580_context* c)
{
    if (!c) FAIL(RT580_INVALID_ARG, "rt580_frame_release: ctx is NULL");
    CU(cudaSetDevice(c->device));
    frame_release(c);
    return RT580_SUCCESS;
}
extern "C" int rt580_frame_read(rt580_context* c, int16_t* fb_out)
{
    if (!c || !fb_out) FAIL(RT580_INVALID_ARG, "rt580_frame_read: NULL argument");
    if (!c->frame || c->frame_imported) FAIL(RT580_FAILURE, "rt580_frame_read: this context does not own a shared frame (rt580_frame_export)");
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(fb_out, c->frame, (size_t)c->frame_w * c->frame_h * 3 * sizeof(int16_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return RT580_SUCCESS;
}

extern "C" int rt580_render(rt580_context* c, const rt580_render_params* p, int16_t* fb_out, rt580_stats* stats)
{
    if (!c || !p) FAIL(RT580_INVALID_ARG, "rt580_render: NULL argument");
    const int n_rows = p->n_rows == 0 ? p->height : (p->n_rows < 0 ? 0 : p->n_rows);
    std::vector<uint64_t> rows((size_t)(n_rows > 0 ? n_rows : 0) + 1);
    int st = rt580_render_begin(c, p, rows.data());
    if (st != RT580_SUCCESS) return st;
    // single context: the rows it owns are the whole stream
    uint64_t run = 0;
    for (int r = 0; r < n_rows; r++) { uint64_t v = rows[r]; rows[r] = run; run += v; }
    return rt580_render_finish(c, rows.data(), fb_out, 0, stats);
}

extern "C" int rt580_last_frame_ao_base(rt580_context* c, uint64_t* out)
{
    if (!c || !out) FAIL(RT580_INVALID_ARG, "rt580_last_frame_ao_base: NULL argument");
    CU(cudaSetDevice(c->device));
    const FrameParams& fp = c->fp;
    const size_t npix = (size_t)fp.n_rows * fp.W;
    if (!npix) return RT580_SUCCESS;
    std::vector<uint32_t> scan(npix); std::vector<uint64_t> rows((size_t)fp.n_rows);
    CU(cudaMemcpy(scan.data(), c->pix_scan.p, sizeof(uint32_t) * npix, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(rows.data(), c->row_vals.p, sizeof(uint64_t) * fp.n_rows, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < npix; i++) {
        const size_t r = i / fp.W;
        out[i] = (rows[r] + (uint64_t)(scan[i] - scan[r * fp.W])) * (uint64_t)c->sc.n_ambient;
    }
    return RT580_SUCCESS;
}

// ---- checkers ---------------------------------------------------------------------------
template <int MODE> static void launch_rays(rt580_context* c, const DeviceScene& sc, bool any, long long n, const float* o, const float* d,
                                            const float* tmax, int32_t* prim, float* t, uint8_t* hit, SlowQ q) {
    if (any) k_trace_rays<MODE, true><<<nblk(n, 128), 128, 0, c->stream>>>(sc, n, o, d, tmax, prim, t, hit, q);
    else k_trace_rays<MODE, false><<<nblk(n, 128), 128, 0, c->stream>>>(sc, n, o, d, tmax, prim, t, hit, q);
}

static int trace_rays_common(rt580_context* c, bool any, int64_t n, const float* org3, const float* dir3, const float* tmax,
                             int traversal, int32_t* prim_out, float* t_out, uint8_t* hit_out)
{
    if (!c || !org3 || !dir3 || n < 0) FAIL(RT580_INVALID_ARG, "rt580_trace: bad argument");
    if (!c->have_scene) FAIL(RT580_FAILURE, "rt580_trace: no scene uploaded");
    if (c->frame_begun) FAIL(RT580_FAILURE, "rt580_trace: a frame is in progress (rt580_render_begin without rt580_render_finish)");
    if (n == 0) return RT580_SUCCESS;
    if (n > 0x7ffffff0ll) FAIL(RT580_INVALID_ARG, "rt580_trace: too many rays");
    CU(cudaSetDevice(c->device));
    struct Tmp {      // freed on every exit path
        float *o = nullptr, *d = nullptr, *tm = nullptr, *t = nullptr; int32_t* pr = nullptr; uint8_t* h = nullptr;
        ~Tmp() { cudaFree(o); cudaFree(d); cudaFree(tm); cudaFree(t); cudaFree(pr); cudaFree(h); }
    } m;
    CU(cudaMalloc(&m.o, sizeof(float) * 3 * n)); CU(cudaMalloc(&m.d, sizeof(float) * 3 * n));
    CU(cudaMemcpy(m.o, org3, sizeof(float) * 3 * n, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(m.d, dir3, sizeof(float) * 3 * n, cudaMemcpyHostToDevice));
    if (any) { CU(cudaMalloc(&m.tm, sizeof(float) * n)); CU(cudaMemcpy(m.tm, tmax, sizeof(float) * n, cudaMemcpyHostToDevice)); CU(cudaMalloc(&m.h, n)); }
    else { CU(cudaMalloc(&m.t, sizeof(float) * n)); CU(cudaMalloc(&m.pr, sizeof(int32_t) * n)); }
    const int mode = pick_mode(c, traversal);
    DeviceScene sc = c->sc;         // the checkers always run the exact path, on a copy: the context's settings stay as they are
    sc.farfield = 1; sc.diag = nullptr;
    SlowQ q = { nullptr, nullptr, nullptr, 0u };
    if (mode == 0 && far_grid_ensure(c)) return RT580_FAILURE;
    const bool deferred = mode == 0 && c->sc.fg_K > 0;      // the frame path's machinery for the rays the tree cannot answer alone
    if (deferred) {
        CU(c->counters.ensure(N_COUNTERS, 0, c->stream));
        CU(c->slow_rays.ensure((size_t)n, 0, c->stream)); CU(c->slow_res.ensure((size_t)n, 0, c->stream));
        CU(cudaMemsetAsync(c->counters.p, 0, N_COUNTERS * sizeof(unsigned), c->stream));
        q.rays = c->slow_rays.p; q.res = c->slow_res.p; q.count = c->counters.p + 2; q.cap = (unsigned)n;
    }
    DISPATCH_MODE(mode, launch_rays, c, sc, any, (long long)n, m.o, m.d, m.tm, m.pr, m.t, m.h, q);
    if (deferred) {
        unsigned n_slow = 0;
        CU(cudaMemcpyAsync(&n_slow, q.count, sizeof n_slow, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        if (n_slow > q.cap) n_slow = q.cap;
        if (n_slow) {
            if (slow_launch(c, any, q.rays, q.res, n_slow)) return RT580_FAILURE;
            if (any) k_trace_rays_apply<true><<<nblk(n_slow, 256), 256, 0, c->stream>>>(q.rays, q.res, n_slow, m.pr, m.t, m.h);
            else k_trace_rays_apply<false><<<nblk(n_slow, 256), 256, 0, c->stream>>>(q.rays, q.res, n_slow, m.pr, m.t, m.h);
        }
    }
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    if (any) CU(cudaMemcpy(hit_out, m.h, n, cudaMemcpyDeviceToHost));
    else { CU(cudaMemcpy(prim_out, m.pr, sizeof(int32_t) * n, cudaMemcpyDeviceToHost)); CU(cudaMemcpy(t_out, m.t, sizeof(float) * n, cudaMemcpyDeviceToHost)); }
    return RT580_SUCCESS;
}

extern "C" int rt580_trace_closest(rt580_context* c, int64_t n, const float* org3, const float* dir3, int traversal,
                                   int32_t* prim_out, float* t_out)
{
    if (!prim_out || !t_out) FAIL(RT580_INVALID_ARG, "rt580_trace_closest: NULL output");
    return trace_rays_common(c, false, n, org3, dir3, nullptr, traversal, prim_out, t_out, nullptr);
}
extern "C" int rt580_trace_any(rt580_context* c, int64_t n, const float* org3, const float* dir3, const float* tmax,
                               int traversal, uint8_t* hit_out)
{
    if (!hit_out || !tmax) FAIL(RT580_INVALID_ARG, "rt580_trace_any: NULL argument");
    return trace_rays_common(c, true, n, org3, dir3, tmax, traversal, nullptr, nullptr, hit_out);
}

extern "C" int rt580_trace_profile(rt580_context* c, int64_t n, const float* org3, const float* dir3, const float* tmax,
                                   uint32_t* counts4)
{
    if (!c || !org3 || !dir3 || !counts4 || n < 0) FAIL(RT580_INVALID_ARG, "rt580_trace_profile: bad argument");
    if (!c->have_scene) FAIL(RT580_FAILURE, "rt580_trace_profile: no scene uploaded");
    if (n == 0) return RT580_SUCCESS;
    CU(cudaSetDevice(c->device));
    struct Tmp { float *o = nullptr, *d = nullptr, *tm = nullptr; unsigned* cn = nullptr; ~Tmp() { cudaFree(o); cudaFree(d); cudaFree(tm); cudaFree(cn); } } m;
    CU(cudaMalloc(&m.o, sizeof(float) * 3 * n)); CU(cudaMalloc(&m.d, sizeof(float) * 3 * n)); CU(cudaMalloc(&m.cn, sizeof(unsigned) * 4 * n));
    CU(cudaMemcpy(m.o, org3, sizeof(float) * 3 * n, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(m.d, dir3, sizeof(float) * 3 * n, cudaMemcpyHostToDevice));
    if (tmax) { CU(cudaMalloc(&m.tm, sizeof(float) * n)); CU(cudaMemcpy(m.tm, tmax, sizeof(float) * n, cudaMemcpyHostToDevice)); }
    DeviceScene sc = c->sc;         // a copy: the context's far-field / diagnostic settings stay as they are
    sc.farfield = 1; sc.diag = nullptr;
    if (tmax) k_trace_profile<true><<<nblk(n, 128), 128, 0, c->stream>>>(sc, (long long)n, m.o, m.d, m.tm, m.cn);
    else k_trace_profile<false><<<nblk(n, 128), 128, 0, c->stream>>>(sc, (long long)n, m.o, m.d, m.tm, m.cn);
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    CU(cudaMemcpy(counts4, m.cn, sizeof(unsigned) * 4 * n, cudaMemcpyDeviceToHost));
    return RT580_SUCCESS;
}

extern "C" int rt580_hemisphere_stream(rt580_context* c, const float normal[3], uint64_t step, int32_t n, float* out3)
{
    if (!c || !normal || !out3 || n < 0) FAIL(RT580_INVALID_ARG, "rt580_hemisphere_stream: bad argument");
    if (n == 0) return RT580_SUCCESS;
    CU(cudaSetDevice(c->device));
    float* d = nullptr;
    CU(cudaMalloc(&d, sizeof(float) * 3 * n));
    k_hemisphere<<<1, 32, 0, c->stream>>>(normal[0], normal[1], normal[2], step, n, d);
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    CU(cudaMemcpy(out3, d, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return RT580_SUCCESS;
}

extern "C" int rt580_powf(rt580_context* c, int64_t n, const float* x, const float* y, float* out)
{
    if (!c || !x || !y || !out || n < 0) FAIL(RT580_INVALID_ARG, "rt580_powf: bad argument");
    if (n == 0) return RT580_SUCCESS;
    CU(cudaSetDevice(c->device));
    float *dx = nullptr, *dy = nullptr, *dz = nullptr;
    CU(cudaMalloc(&dx, sizeof(float) * n)); CU(cudaMalloc(&dy, sizeof(float) * n)); CU(cudaMalloc(&dz, sizeof(float) * n));
    CU(cudaMemcpy(dx, x, sizeof(float) * n, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dy, y, sizeof(float) * n, cudaMemcpyHostToDevice));
    k_powf<<<nblk(n, 256), 256, 0, c->stream>>>((long long)n, dx, dy, dz);
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    CU(cudaMemcpy(out, dz, sizeof(float) * n, cudaMemcpyDeviceToHost));
    cudaFree(dx); cudaFree(dy); cudaFree(dz);
    return RT580_SUCCESS;
}
