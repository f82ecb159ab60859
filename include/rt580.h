/* rt580.h - C ABI of the B200-native 580-Raytracer hot path (librt580.so).
 *
 * The reference has no plugin / FFI interface: its API is the public surface of
 * `class Raytracer` (Raytracer.h:557-588) as driven by main() (Raytracer.cpp:944-953).
 * This repo keeps that class (580-raytracer_b200/csrc/raytracer.h) and cuts the C ABI
 * INSIDE Raytracer::Render (Raytracer.cpp:916-935), between InitializeRenderer()
 * (cpp:917) and FlushFrameBufferToPPM() (cpp:934): everything the double loop at
 * cpp:921-932 does - GenerateRay, Raycast, IntersectScene/Triangle/Sphere,
 * CalculateLocalColor, CalculateAmbientOcclusion, ComputeFresnel, CalculateRefraction -
 * happens behind rt580_render*().
 *
 * Conventions (match the reference): every call returns RT580_SUCCESS / RT580_FAILURE /
 * RT580_INVALID_ARG (Raytracer.h:8-10); diagnostics go to rt580_last_error(); calls are
 * synchronous and a context is not re-entrant (the reference Render is neither,
 * Raytracer.h:592).  All pointers are host memory borrowed for the call unless a
 * parameter says "device".  The library owns all device memory.  There is no CPU
 * fallback: without a CUDA device every compute entry point fails.
 *
 * All "cpp:" / "h:" citations are /root/reference/580 Raytracer/Raytracer.{cpp,h}.
 */
#ifndef RT580_H
#define RT580_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT580_SUCCESS     0   /* h:8  RT_SUCCESS     */
#define RT580_FAILURE     1   /* h:9  RT_FAILURE     */
#define RT580_INVALID_ARG 2   /* h:10 RT_INVALID_ARG */

/* Light::Type, h:520-524 */
#define RT580_LIGHT_DIRECTIONAL 0
#define RT580_LIGHT_POINT       1
#define RT580_LIGHT_AMBIENT     2

/* How the ambient-occlusion sample stream (cpp:269-292, cpp:315-330) is addressed. */
#define RT580_RNG_REFERENCE_LCG 0   /* replays the reference's single sequential
                                       std::default_random_engine stream (h:592, cpp:787):
                                       AO call j of the frame in scanline/pre-order starts at
                                       engine step 2*spp*j.  Needs the global prefix of AO-call
                                       counts, i.e. one exchange between ranks.              */
#define RT580_RNG_COUNTER       1   /* same generator, position = f(pixel, node): no
                                       cross-pixel dependency, no exchange, different noise. */

/* How closest/any hits are searched (cpp:473-526 is a linear loop; results are identical). */
#define RT580_TRAVERSAL_AUTO        0   /* brute force from shared memory for tiny scenes, LBVH otherwise */
#define RT580_TRAVERSAL_BVH         1
#define RT580_TRAVERSAL_BRUTE_FORCE 2   /* the reference's own linear loop, on the GPU (checker / tiny scenes) */

/* The reference's float triangle test (cpp:392-396) also accepts points 10^4..10^7 units away
 * when a ray is almost parallel to a triangle's plane (the cross product at cpp:392 is rounding
 * noise there).  Those "hits" decide hit/miss for rays that leave the scene, so they are part of
 * the reference's image.  EXACT replays them (bit-identical to the linear loop, costs a scan of
 * the filter records for every ray that finds nothing nearer); OFF treats them as misses. */
#define RT580_FARFIELD_EXACT 0
#define RT580_FARFIELD_OFF   1

typedef struct rt580_context rt580_context;

/* Scene after the load-time flatten (replaces the per-ray work of cpp:477-480 and
 * cpp:353-355): world-space triangle vertices = ComputeModelMatrix(shape).TransformPoint(v)
 * (cpp:528-586, h:234-248), sphere centre = translation column of that matrix, radius
 * unscaled (cpp:421-423).  Primitive order = (shape order, triangle order) - it decides
 * ties (cpp:494, cpp:513).  Every float4 array is 16-byte records. */
typedef struct rt580_flat_scene {
    int64_t        n_prims;        /* triangles + spheres, in reference order            */
    int64_t        n_tris;
    const float*   tri_v0;         /* [n_tris][4] world xyz, w ignored                   */
    const float*   tri_v1;         /* [n_tris][4]                                         */
    const float*   tri_v2;         /* [n_tris][4]                                         */
    const float*   tri_n0;         /* [n_tris][4] OBJECT-space vertex normals (cpp:227)  */
    const float*   tri_n1;
    const float*   tri_n2;
    const int32_t* tri_prim;       /* [n_tris] index of the triangle in primitive order  */
    const int32_t* tri_material;   /* [n_tris] material (= shape) index                  */
    int64_t        n_spheres;
    const float*   sph_center_r;   /* [n_spheres][4] world centre xyz, radius            */
    const int32_t* sph_prim;       /* [n_spheres]                                         */
    const int32_t* sph_material;   /* [n_spheres]                                         */
    int32_t        n_materials;
    const float*   materials;      /* [n_materials][8] Cs.rgb Ka Kd Ks Kt n (h:442-463)  */
    int32_t        n_lights;
    const int32_t* light_type;     /* [n_lights] RT580_LIGHT_*, JSON order (cpp:39)      */
    const float*   light_f;        /* [n_lights][10] color.rgb intensity position.xyz direction.xyz */
    float          origin_hint[3]; /* Camera::from (h:506): together with the scene bounds it bounds
                                      every ray origin, which sizes the conservative padding of the
                                      BVH boxes; a render whose camera lies outside is refused */
} rt580_flat_scene;

/* The same scene BEFORE FlattenScene: meshes in object space once, one model matrix per shape.  The
 * library applies Matrix::TransformPoint (h:234-248, the reference's operation order, unfused) on the device and produces
 * exactly the arrays of rt580_flat_scene - c4: 150 KB over PCIe instead of 104 MB.  Shape k has material k; primitive order =
 * (shape order, triangle order) as above. */
typedef struct rt580_instanced_scene {
    int32_t        n_meshes;
    const int64_t* mesh_first;     /* [n_meshes + 1] first triangle of each mesh in mesh_tris                     */
    const float*   mesh_tris;      /* [mesh_first[n_meshes]][18] pos[3].xyz then nrm[3].xyz, object space (h:436) */
    int32_t        n_shapes;
    const int32_t* shape_mesh;     /* [n_shapes] mesh index, or -1: a sphere (cpp:421-423)                        */
    const float*   shape_matrix;   /* [n_shapes][16] row-major model matrix (ComputeModelMatrix, cpp:528-586)     */
    const float*   shape_radius;   /* [n_shapes] sphere radius, unscaled (cpp:423); ignored for meshes            */
    const float*   materials;      /* [n_shapes][8] Cs.rgb Ka Kd Ks Kt n                                          */
    int32_t        n_lights;
    const int32_t* light_type;
    const float*   light_f;
    float          origin_hint[3];
} rt580_instanced_scene;

typedef struct rt580_render_params {
    int32_t width, height;         /* Display xRes,yRes (h:420-425) - the ctor's, not the JSON's (Q23) */
    float   fov_degrees;           /* Display::fov, 60 in the reference (cpp:786)          */
    float   camera_from[3];        /* Camera::from (h:506): ray origin + Phong view point  */
    float   inv_view3x3[9];        /* upper-left 3x3 of Inverse(viewMatrix), row major (cpp:849-851) */
    int32_t depth;                 /* Raycast bounces, 4 in the reference (h:563)          */
    int32_t ao_spp;                /* CalculateAmbientOcclusion samples, 128 in the reference (cpp:317) */
    int32_t rng_mode;              /* RT580_RNG_*                                          */
    int32_t traversal;             /* RT580_TRAVERSAL_*                                    */
    /* rows rendered by this context: row_first + k*row_step, k in [0,n_rows).
     * n_rows == 0 means the whole frame (row_first 0, step 1); n_rows < 0 means no row at all (a rank beyond the
     * frame's height: it still takes part in the exchange, with row_step = the number of ranks). */
    int32_t row_first, row_step, n_rows;
    int32_t farfield;              /* RT580_FARFIELD_*                                     */
} rt580_render_params;

typedef struct rt580_scene_info {
    int64_t  n_leaf;               /* primitives in the BVH                                */
    int64_t  n_dropped;            /* zero-area triangles (never intersectable, cpp:365-373) */
    int64_t  n_always;             /* sliver triangles that are far-field candidates for every ray */
    float    far_tmin;             /* smallest t at which a far-field acceptance is possible */
    float    pad;                  /* global part of the box padding                       */
    float    extent;               /* E: bound on |coordinate| of ray origins / hit points */
    float    build_ms;             /* device time of the build kernels                     */
    uint32_t bvh_max_depth;
    uint32_t reserved;
} rt580_scene_info;

typedef struct rt580_stats {
    uint64_t rays_primary;         /* IntersectScene calls from cpp:30 at depth 0          */
    uint64_t rays_secondary;       /* cpp:30 via cpp:103 / cpp:111                         */
    uint64_t rays_shadow;          /* cpp:75                                               */
    uint64_t rays_ao;              /* cpp:325                                              */
    uint64_t hit_nodes;            /* Raycast nodes that hit something                     */
    uint64_t ao_calls;
    float    ms_structure;         /* primary+secondary closest-hit + shading + shadow     */
    float    ms_order;             /* subtree counts, scan, pre-order ordinals             */
    float    ms_ao;                /* occlusion pass                                       */
    float    ms_resolve;           /* integer Pixel algebra bottom-up                      */
    float    ms_total;             /* device time of the frame (CUDA events)               */
    float    ms_ao_kernel;         /* the dominant kernel alone                            */
    uint32_t kernel_launches;      /* kernels launched for the frame                       */
    uint32_t bvh_max_depth;
    uint32_t far_scans;            /* rays that needed the far-field scan (found nothing nearer) */
    uint32_t linear_fallbacks;     /* rays that started outside the padded extent (children of far hits) */
    uint64_t ao_rays_traversed;    /* AO rays that went through the tree; the others were already occluded by
                                      one of the scene's few very large primitives (tested first) */
    uint64_t shadow_rays_traversed;/* shadow rays that went through the tree; for the others the point light's
                                      clearance map proved that no tree primitive lies before the light */
} rt580_stats;

/* Where the device time of a frame goes: the kernels of the hot path grouped by the kind of ray they serve.  ms = sum of the
 * launches' durations, CUDA events on the launching stream (classes on different streams overlap: the sum can exceed
 * the frame time); rays = what the class processed.  One ray = one IntersectScene call of the reference (cpp:473). */
#define RT580_N_CLASSES        10
#define RT580_CLASS_PRIMARY     0   /* GenerateRay + closest hit of the camera rays (cpp:832-858, cpp:30)                 */
#define RT580_CLASS_CLOSEST     1   /* closest hit of the reflection / refraction rays through the tree (cpp:103, 111)    */
#define RT580_CLASS_SHADOW_GEN  2   /* shadow rays: generation, large primitives, clearance maps, Phong terms (cpp:53-81) */
#define RT580_CLASS_SHADOW_TREE 3   /* any hit of the shadow rays through the tree                                         */
#define RT580_CLASS_AO_GEN      4   /* AO sample rays: random stream, hemisphere directions, large primitives (cpp:269-292, 320-322) */
#define RT580_CLASS_AO_TREE     5   /* any hit of the AO rays through the tree (cpp:325)                                   */
#define RT580_CLASS_FAR_ANY     6   /* far-field replay for any-hit rays that left the scene / start outside it           */
#define RT580_CLASS_FAR_CLOSEST 7   /* the same for closest-hit rays                                                        */
#define RT580_CLASS_ORDER       8   /* subtree sizes, scans, pre-order ordinals, stream seeds          */
#define RT580_CLASS_RESOLVE     9   /* integer Pixel algebra bottom-up, frame store (cpp:39-51, 114-128)                   */
typedef struct rt580_profile {
    float    ms[RT580_N_CLASSES];
    uint64_t rays[RT580_N_CLASSES];
    uint32_t launches[RT580_N_CLASSES];
    uint32_t reserved;
    /* with rt580_set_profiling(ctx, 1): inner-node visits and leaf (primitive) tests of the tree kernels, summed over their rays */
    uint64_t nodes_any, leaves_any, nodes_closest, leaves_closest;
} rt580_profile;

/* ---- context ------------------------------------------------------------------------- */
int  rt580_create(int device, rt580_context** out);
void rt580_destroy(rt580_context* ctx);
const char* rt580_last_error(void);
/* device properties the benchmark needs for its FP32 roofline: SM count, max SM MHz */
int  rt580_device_info(rt580_context* ctx, int32_t* sm_count, int32_t* sm_clock_mhz, uint64_t* hbm_bytes);
/* the cudaStream_t every kernel of this context is launched on (for CUDA-event timing by the caller) */
int  rt580_get_stream(rt580_context* ctx, void** cuda_stream);

/* Host staging memory: page-locked when a CUDA device is present (copies to / from the device then run
 * at PCIe speed instead of through the driver's bounce buffers), plain malloc otherwise.  The host
 * class keeps the flattened scene and the frame buffer in it; any host pointer works with the calls
 * below, this is only faster. */
void* rt580_host_alloc(uint64_t bytes);
void  rt580_host_free(void* p);

/* ---- scene: H2D + per-triangle constants (cpp:362-365, 377, 389) + LBVH build ----------- */
int  rt580_upload_scene(rt580_context* ctx, const rt580_flat_scene* scene);
/* the same from the un-flattened scene: FlattenScene (cpp:348-365's TransformPoint calls, hoisted) runs on the device */
int  rt580_upload_instanced_scene(rt580_context* ctx, const rt580_instanced_scene* scene);
/* the device flatten alone, results copied back to host arrays laid out as in rt580_flat_scene (a test hook: the arrays
 * must equal the host FlattenScene's byte for byte).  Any output pointer may be NULL. */
int  rt580_flatten_instanced(rt580_context* ctx, const rt580_instanced_scene* scene, float* tri_v0, float* tri_v1, float* tri_v2,
                             float* tri_n0, float* tri_n1, float* tri_n2, int32_t* tri_prim, int32_t* tri_material,
                             float* sph_center_r, int32_t* sph_prim, int32_t* sph_material);
/* device ms of the last upload's build kernels (setup, morton, sort, hierarchy, refit, pack) */
int  rt580_build_ms(rt580_context* ctx, float* ms);
int  rt580_scene_info_get(rt580_context* ctx, rt580_scene_info* out);

/* ---- frame: replaces the loop body of Raytracer::Render (cpp:921-932) ------------------ */
/* One call = whole frame (or the rows in params) on this context's GPU.
 * fb_out: [n_rows][width][3] int16 raw Pixel{short r,g,b} (h:373-374), host memory. */
int  rt580_render(rt580_context* ctx, const rt580_render_params* params, int16_t* fb_out, rt580_stats* stats);

/* The last finished frame (its rows on this context) as 8-bit RGB, the body of the reference's PPM
 * (cpp:809-823): rgb = lut256[value] with the caller's table of u8(powf(c / 255.0f, 1 / 2.2f) * 255.0f),
 * c = 0..255, evaluated with the host's powf.  rgb_out: [n_rows][width][3] bytes, host or device (any alignment).
 * Values outside [0, 255] are clamped to the table's ends; the hot path only stores values in [0, 255] (cpp:128 clamps,
 * the background is a constant), so this never differs from the host's FlushFrameBufferToPPM on a rendered frame. */
int  rt580_frame_rgb8(rt580_context* ctx, const uint8_t* lut256, uint8_t* rgb_out, int out_on_device);

/* Split form for several ranks (one context per GPU, rows partitioned):
 *   begin  : structure pass; row_hit_nodes[n_rows] = AO-relevant hit nodes per owned row
 *   finish : row_ao_base[n_rows] = for each owned row the number of hit nodes in ALL rows
 *            before it in scanline order (the exchange: all-gather of row_hit_nodes, then a
 *            prefix sum); ignored (may be NULL) in RT580_RNG_COUNTER mode.
 *            fb_out may be host memory, or device memory when fb_on_device != 0 (so the
 *            caller can hand it to NCCL without staging). */
int  rt580_render_begin(rt580_context* ctx, const rt580_render_params* params, uint64_t* row_hit_nodes);
int  rt580_render_finish(rt580_context* ctx, const uint64_t* row_ao_base, int16_t* fb_out, int fb_on_device,
                         rt580_stats* stats);

/* ---- several GPUs, one frame: the exchange and the gather without leaving the device -----
 * The reference renders its rows in one loop (cpp:921-922) and shares one random stream between
 * them (h:592).  With rows interleaved over `world` contexts (row y -> rank y % world), what the
 * ranks must exchange is the per-row hit-node count (for the stream position of every row) and,
 * at the end, the rows themselves.  Both stay on the device:
 *   rt580_row_counts_to_device     after rt580_render_begin (row_hit_nodes may then be NULL): the
 *                                  counts of this context's rows -> dst_device[max_rows] (zero padded),
 *                                  on the context's stream; the caller all-gathers them (NCCL) into
 *                                  all_counts_device[world][max_rows] on the same stream;
 *   rt580_render_finish_interleaved computes this rank's row_ao_base from all_counts_device on the
 *                                  device and finishes the frame like rt580_render_finish;
 *   rt580_frame_export / _import   rank 0 allocates the whole width x height frame and exports a
 *                                  64-byte CUDA IPC handle; the other ranks' contexts map it.  While
 *                                  a context has a shared frame, every finish stores its rows into
 *                                  it (for the importing ranks these are stores over NVLink to rank
 *                                  0's memory): no NCCL gather, no staging.  The caller orders rank
 *                                  0's read after the other ranks' stores (one tiny collective on
 *                                  the contexts' streams);
 *   rt580_frame_read               rank 0: the whole frame -> host [height][width][3] int16. */
int  rt580_row_counts_to_device(rt580_context* ctx, uint64_t* dst_device, int32_t max_rows);
int  rt580_render_finish_interleaved(rt580_context* ctx, const uint64_t* all_counts_device, int32_t world, int32_t rank,
                                     int32_t max_rows, int16_t* fb_out, int fb_on_device, rt580_stats* stats);
/* Release order of a shared frame: every importing context calls rt580_frame_release (cudaIpcCloseMemHandle) BEFORE the
 * exporting context releases (or re-exports) the allocation; freeing an exported region that is still mapped elsewhere is
 * undefined behaviour in CUDA.  A barrier between the two is the caller's (bench.py: Rig.release_frame). */
int  rt580_frame_export(rt580_context* ctx, int32_t width, int32_t height, void* ipc_handle64);
int  rt580_frame_import(rt580_context* ctx, const void* ipc_handle64, int32_t width, int32_t height);
int  rt580_frame_release(rt580_context* ctx);
int  rt580_frame_read(rt580_context* ctx, int16_t* fb_out);

/* The class breakdown of the last finished frame.  rt580_set_profiling(ctx, 1) makes the tree kernels of the following frames
 * count node visits and leaf tests as well (a few percent slower: not for timed runs). */
int  rt580_set_profiling(rt580_context* ctx, int count_visits);
int  rt580_frame_profile(rt580_context* ctx, rt580_profile* out);

/* ---- checkers (used by the parity tests; same kernels as the frame path) ---------------- */
/* Closest hit of n arbitrary rays: prim_out = primitive order index or -1, t_out = distance. */
int  rt580_trace_closest(rt580_context* ctx, int64_t n, const float* org3, const float* dir3, int traversal,
                         int32_t* prim_out, float* t_out);
/* Any hit with t <= tmax (cpp:75 / cpp:325 use only the bool). */
int  rt580_trace_any(rt580_context* ctx, int64_t n, const float* org3, const float* dir3, const float* tmax,
                     int traversal, uint8_t* hit_out);
/* Traversal profile of n rays through the LBVH path: counts4[4*i..] = node visits, leaf tests,
 * far-field scans, linear fallbacks of ray i.  tmax == NULL: closest hit, else any hit. */
int  rt580_trace_profile(rt580_context* ctx, int64_t n, const float* org3, const float* dir3, const float* tmax,
                         uint32_t* counts4);
/* Per-pixel global ordinal of the first AO call of the last rendered frame (local rows). */
int  rt580_last_frame_ao_base(rt580_context* ctx, uint64_t* pixel_ao_base /* [n_rows*width] */);
/* n AO directions of the stream starting at engine step `step` (cpp:283-292). */
int  rt580_hemisphere_stream(rt580_context* ctx, const float normal[3], uint64_t step, int32_t n, float* out3);
/* device powf used by the Phong term (cpp:253), element-wise */
int  rt580_powf(rt580_context* ctx, int64_t n, const float* x, const float* y, float* out);

/* ---- class mirror: the reference's public API over the C ABI (h:557-588) ---------------- */
typedef struct rt580_raytracer rt580_raytracer;
rt580_raytracer* rt580_raytracer_new(int width, int height);                 /* h:588 ctor            */
void rt580_raytracer_delete(rt580_raytracer* rt);
int  rt580_raytracer_set_assets_path(rt580_raytracer* rt, const char* dir);  /* ASSETS_PATH, h:15     */
int  rt580_raytracer_set_options(rt580_raytracer* rt, int depth, int ao_spp, int rng_mode, int traversal,
                                 int device, int farfield);                  /* h:563, cpp:317        */
int  rt580_raytracer_set_quiet(rt580_raytracer* rt, int quiet);              /* mute cpp:592 / cpp:772 prints */
/* rows of the frame interleaved over GPUs device .. device + n_gpus - 1 of this process: the frame is bit-identical */
int  rt580_raytracer_set_gpus(rt580_raytracer* rt, int n_gpus);
/* binary cache of parsed meshes, "<dir>/<mesh>.rt580mesh", validated against the JSON's size and hash;
 * NULL or "" = off.  The loaded meshes are bit-identical with or without it (cpp:568-643 stays the source of truth). */
int  rt580_raytracer_set_mesh_cache(rt580_raytracer* rt, const char* dir);
int  rt580_raytracer_mesh_cache_hits(rt580_raytracer* rt);
/* FlattenScene on the device: Render uploads meshes + one model matrix per shape (rt580_upload_instanced_scene) */
int  rt580_raytracer_set_device_flatten(rt580_raytracer* rt, int on);
int  rt580_raytracer_instanced_scene(rt580_raytracer* rt, rt580_instanced_scene* out);   /* borrows the object's buffers */
int  rt580_raytracer_load_scene_json(rt580_raytracer* rt, const char* scene);/* h:572 LoadSceneJSON   */
int  rt580_raytracer_render(rt580_raytracer* rt, const char* output_ppm);    /* h:586 Render          */
int  rt580_raytracer_flush_ppm(rt580_raytracer* rt, const char* output_ppm); /* h:573                 */
const int16_t* rt580_raytracer_framebuffer(rt580_raytracer* rt);             /* Display::frameBuffer  */
int  rt580_raytracer_stats(rt580_raytracer* rt, rt580_stats* stats);
/* the flattened scene + camera the class hands to rt580_upload_scene / rt580_render */
int  rt580_raytracer_flat_scene(rt580_raytracer* rt, rt580_flat_scene* out);
int  rt580_raytracer_render_params(rt580_raytracer* rt, rt580_render_params* out);

#ifdef __cplusplus
}
#endif
#endif /* RT580_H */
