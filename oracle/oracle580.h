/* TEST INFRASTRUCTURE ONLY - oracle tier T1: a CPU restatement of the reference's
 * per-pixel hot path (Raytracer::Render -> Raycast -> IntersectScene / ... ), written
 * fresh for this repo, with the loop-invariant work hoisted (model matrices, world
 * vertices, plane constants) so that it can reach scenes the verbatim reference
 * (tier T0, oracle/_ref) cannot.  Pinned bit-for-bit against T0 by
 * tests/test_oracle_vs_ref.py on every input T0 can finish.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this
 * library.  The product (580-raytracer_b200/) never links, loads or calls it.
 *
 * All file:line citations are into /root/reference/580 Raytracer/ (cpp = Raytracer.cpp,
 * h = Raytracer.h).
 */
#ifndef ORACLE580_H
#define ORACLE580_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Scene exactly as the reference loader leaves it in memory (h:436-555), i.e.
 * object-space meshes + per-shape material/transform, NOT flattened. */
typedef struct orc_scene {
    int32_t n_shapes;
    const int32_t* shape_mesh;     /* [n_shapes] index into meshes                      */
    const float*   shape_material; /* [n_shapes][8] Cs.rgb, Ka, Kd, Ks, Kt, n (h:442-463) */
    const float*   shape_srt;      /* [n_shapes][9] S.xyz, R.xyz (degrees), T.xyz (h:532-537) */
    int32_t n_meshes;
    const int32_t* mesh_type;      /* [n_meshes] 0 = RT_POLYGON, 1 = RT_SPHERE (h:476-485) */
    const int32_t* mesh_tri_begin; /* [n_meshes+1] prefix into tri_pos / tri_nrm          */
    const float*   mesh_radius;    /* [n_meshes]                                          */
    const float*   tri_pos;        /* [n_tris][9] object-space v0,v1,v2 positions         */
    const float*   tri_nrm;        /* [n_tris][9] object-space vertex normals             */
    int32_t n_lights;
    const int32_t* light_type;     /* [n_lights] 0 Directional, 1 Point, 2 Ambient (h:520-524) */
    const float*   light_f;        /* [n_lights][10] color.rgb, intensity, position.xyz, direction.xyz */
    float cam_from[3];
    float cam_to[3];
} orc_scene;

/* World-space restatement of the scene (hoists cpp:480, cpp:353-365, cpp:377, cpp:389
 * out of the per-ray loop; bit-identical because those are pure functions). */
typedef struct orc_world orc_world;
orc_world* orc580_prepare(const orc_scene* s);
void orc580_free(orc_world* w);
int64_t orc580_num_prims(const orc_world* w);

/* cpp:528-586 */
void orc580_model_matrix(const float srt[9], float m16[16]);
/* cpp:131-166 + cpp:168-203 */
void orc580_fresnel(float ior, const float n[3], const float i[3], float* kr, float* kt, float refr[3]);
/* cpp:269-292 with the libstdc++ stream of SURVEY Appendix C; engine position `step`
 * (number of engine steps already consumed). Writes n directions. */
void orc580_hemisphere_stream(const float normal[3], uint64_t step, int n, float* out3);
/* engine state after `steps` steps of minstd_rand0 from seed 1 (closed form). */
uint32_t orc580_lcg_state(uint64_t steps);
/* cpp:832-858 (+ cpp:895-915): primary ray for pixel (x,y). returns 0 on success */
int orc580_primary_ray(const orc_world* w, int W, int H, int x, int y, float org[3], float dir[3]);

/* Render pixels.
 *   pix        : npix pixel ids (y*W+x); NULL = the full frame in scanline order (npix = W*H)
 *   ao_base    : per listed pixel, the global ordinal of its first AO call (SURVEY
 *                Appendix C); NULL = derive it: serial running stream if nthreads==1,
 *                structure pre-pass + prefix sum otherwise (full-frame / in-order lists only)
 *   out        : [npix][3] int16 raw Pixel values (h:373-418)
 *   rays       : number of IntersectScene calls (cpp:473) summed over the pixels
 *   hit_nodes  : optional [npix] number of hit Raycast nodes per pixel
 * returns RT_SUCCESS(0) / RT_FAILURE(1) / RT_INVALID_ARG(2) (h:8-10). */
int orc580_render(const orc_world* w, int W, int H, int spp, int depth,
                  int64_t npix, const int32_t* pix, const uint64_t* ao_base, int nthreads,
                  int16_t* out, uint64_t* rays, uint32_t* hit_nodes);

/* Closest hit of n rays by the reference's linear loop (cpp:473-526).  dir is used
 * as given (the Ray ctor's normalisation, h:431-433, is the caller's business).
 * prim_out = index in (shape order, triangle order; a sphere shape counts 1) or -1. */
void orc580_intersect_batch(const orc_world* w, int64_t n, const float* org3, const float* dir3,
                            int64_t* prim_out, float* t_out, int nthreads);

/* Serial render of the listed pixels (NULL = whole frame) that also records every
 * IntersectScene call in call order: origin, direction, closest primitive (-1 = miss), t,
 * kind (0 primary/secondary cpp:30, 1 shadow cpp:75, 2 AO cpp:325).  Returns the number of
 * rays logged (capped at max_rays), -1 on failure. */
int64_t orc580_render_log(const orc_world* w, int W, int H, int spp, int depth, int64_t npix, const int32_t* pix,
                          const uint64_t* ao_base, int64_t max_rays, float* org3, float* dir3, int64_t* prim,
                          float* t, int32_t* kind, int16_t* out);

/* cpp:796-830 gamma encode of a raw frame buffer into 8-bit RGB. */
void orc580_gamma_encode(const int16_t* fb, int64_t n_channels, uint8_t* out);

#ifdef __cplusplus
}
#endif
#endif
