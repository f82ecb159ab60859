/* TEST INFRASTRUCTURE ONLY - oracle tier T1 (see oracle580.h for the rules).
 *
 * CPU restatement of the reference's per-pixel hot path, every function citing the
 * reference file:line it follows (cpp = /root/reference/580 Raytracer/Raytracer.cpp,
 * h = Raytracer.h).  Build: gcc -O2 -ffp-contract=off -fopenmp (no FMA contraction:
 * the reference's arithmetic is plain IEEE fp32 mul/add/div/sqrt).
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py compares this file bit-for-bit
 * with the reference itself (oracle/_ref/libref580.so, built from the reference's own
 * sources by oracle/build_ref.sh) on whole frames and on single functions, and
 * tests/golden/ holds the reference-generated frame buffers for the GPU box.
 */
#include "oracle580.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define RT_SUCCESS 0
#define RT_FAILURE 1
#define RT_INVALID_ARG 2
#define EPSILON 1e-6                 /* h:12  - a DOUBLE literal; floats are promoted for compares */
#define SHADOW_CLIPPING_OFFSET 0.2f  /* h:13  - 0.2 converted to float by Vector3::operator*(float) h:83 */
#define PI_REF 3.14159265            /* h:11 */
#define REFRACTIVE_INDEX 2.5f        /* h:460 - never loaded from JSON (Q5) */

typedef struct { float x, y, z; } v3;
typedef struct { int16_t r, g, b; } pix;

/* ---- h:39-149 Vector3 ---------------------------------------------------------- */
static inline v3 V(float x, float y, float z) { v3 r = { x, y, z }; return r; }
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }      /* h:97  */
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }      /* h:93  */
static inline v3 vmuls(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }        /* h:83  */
static inline v3 vmulv(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }     /* h:88  */
static inline v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }                           /* h:101 */
static inline float vdot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }    /* h:131 */
static inline v3 vcross(v3 a, v3 b) {                                                 /* h:122 */
    return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline float vlength(v3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); } /* h:118 */
static inline v3 vnormalize(v3 a) {                                                   /* h:109-116 */
    float length = sqrtf(a.x * a.x + a.y * a.y + a.z * a.z);
    if (length > 0) { a.x /= length; a.y /= length; a.z /= length; }
    return a;
}
static inline v3 vreflect(v3 I, v3 N) {                                               /* h:143-148 */
    float IDotN = vdot(I, N);
    IDotN *= 2;
    return vsub(I, vmuls(N, IDotN));
}

/* ---- h:373-418 Pixel ------------------------------------------------------------ */
static inline int16_t f2short(float v) {
    /* static_cast<short>(float): gcc emits cvttss2si (32 bit) and keeps the low 16 bits */
    return (int16_t)(int32_t)v;
}
static inline pix pix_from_v3(v3 c) {   /* h:376-381: clamp() result is discarded => NO clamp (Q2) */
    pix p = { f2short(c.x * 255), f2short(c.y * 255), f2short(c.z * 255) };
    return p;
}
static inline int16_t clamp255(int16_t v) { return (v > 255) ? 255 : (v < 0 ? 0 : v); }
static inline pix pix_clamp(pix p) { pix r = { clamp255(p.r), clamp255(p.g), clamp255(p.b) }; return r; } /* h:411-417 */
static inline pix pix_muls(pix p, float s) {   /* h:394-400: truncate, then clamp */
    pix r = { f2short(p.r * s), f2short(p.g * s), f2short(p.b * s) };
    return pix_clamp(r);
}
static inline pix pix_add(pix a, pix b) {      /* h:403-409: no clamp, short wrap */
    pix r = { (int16_t)(a.r + b.r), (int16_t)(a.g + b.g), (int16_t)(a.b + b.b) };
    return r;
}

/* ---- h:168-371 Matrix ----------------------------------------------------------- */
typedef struct { float m[4][4]; } mat4;

static mat4 mat_mul(const mat4* a, const mat4* b) {   /* h:179-190 */
    mat4 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            r.m[i][j] = 0;
            for (int k = 0; k < 4; ++k) r.m[i][j] += a->m[i][k] * b->m[k][j];
        }
    return r;
}
static void mat_identity(mat4* m) {   /* cpp:872-878 */
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) m->m[i][j] = (i == j) ? 1.0f : 0.0f;
}
static v3 mat_transform_point(const mat4* M, v3 p) {   /* h:234-248 */
    float x = M->m[0][0] * p.x + M->m[0][1] * p.y + M->m[0][2] * p.z + M->m[0][3];
    float y = M->m[1][0] * p.x + M->m[1][1] * p.y + M->m[1][2] * p.z + M->m[1][3];
    float z = M->m[2][0] * p.x + M->m[2][1] * p.y + M->m[2][2] * p.z + M->m[2][3];
    float w = M->m[3][0] * p.x + M->m[3][1] * p.y + M->m[3][2] * p.z + M->m[3][3];
    if (w != 1.0f) { x /= w; y /= w; z /= w; }
    return V(x, y, z);
}
static v3 mat_transform_dir(const mat4* M, v3 d) {   /* h:227-232 */
    float x = M->m[0][0] * d.x + M->m[0][1] * d.y + M->m[0][2] * d.z;
    float y = M->m[1][0] * d.x + M->m[1][1] * d.y + M->m[1][2] * d.z;
    float z = M->m[2][0] * d.x + M->m[2][1] * d.y + M->m[2][2] * d.z;
    return V(x, y, z);
}
static float det3(const mat4* a) {   /* h:251-255 (upper-left 3x3 of a scratch Matrix) */
    return a->m[0][0] * (a->m[1][1] * a->m[2][2] - a->m[1][2] * a->m[2][1]) -
           a->m[0][1] * (a->m[1][0] * a->m[2][2] - a->m[1][2] * a->m[2][0]) +
           a->m[0][2] * (a->m[1][0] * a->m[2][1] - a->m[1][1] * a->m[2][0]);
}
static float det4(const mat4* a) {   /* h:257-274 */
    float det = 0;
    for (int i = 0; i < 4; i++) {
        mat4 sub;
        memset(&sub, 0, sizeof sub);
        for (int j = 1; j < 4; j++)
            for (int k = 0; k < 4; k++) {
                if (k < i) sub.m[j - 1][k] = a->m[j][k];
                else if (k > i) sub.m[j - 1][k - 1] = a->m[j][k];
            }
        det += (i % 2 == 0 ? 1 : -1) * a->m[0][i] * det3(&sub);
    }
    return det;
}
static void adjoint4(const mat4* a, mat4* adj) {   /* h:276-296 */
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            mat4 sub;
            memset(&sub, 0, sizeof sub);
            int subi = 0;
            for (int k = 0; k < 4; k++) {
                if (k == i) continue;
                int subj = 0;
                for (int l = 0; l < 4; l++) {
                    if (l == j) continue;
                    sub.m[subi][subj] = a->m[k][l];
                    subj++;
                }
                subi++;
            }
            float cof = det3(&sub);
            if ((i + j) % 2 != 0) cof = -cof;
            adj->m[j][i] = cof;
        }
}
static int mat_inverse(const mat4* a, mat4* out) {   /* h:354-370 */
    float det = det4(a);
    if (fabs(det) < 1e-10) return RT_FAILURE;
    mat4 adj;
    adjoint4(a, &adj);
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) out->m[i][j] = adj.m[i][j] / det;
    return RT_SUCCESS;
}

static inline float to_radian(float degrees) { return degrees * (PI_REF / 180); }   /* h:581-583 */

/* cpp:528-586: S * (Rz*Ry*Rx) * T (Q11); cos/sin are the double libm functions of a
 * float radian, rounded on the store into the float matrix (Q27). */
static mat4 model_matrix(const float* srt) {
    mat4 S; mat_identity(&S);
    S.m[0][0] = srt[0]; S.m[1][1] = srt[1]; S.m[2][2] = srt[2]; S.m[3][3] = 1.0f;
    float radX = to_radian(srt[3]), radY = to_radian(srt[4]), radZ = to_radian(srt[5]);
    mat4 RX; mat_identity(&RX);
    RX.m[1][1] = cos(radX); RX.m[1][2] = -sin(radX); RX.m[2][1] = sin(radX); RX.m[2][2] = cos(radX);
    RX.m[0][0] = 1.0f; RX.m[3][3] = 1.0f;
    mat4 RY; mat_identity(&RY);
    RY.m[0][0] = cos(radY); RY.m[0][2] = sin(radY); RY.m[2][0] = -sin(radY); RY.m[2][2] = cos(radY);
    RY.m[1][1] = 1.0f; RY.m[3][3] = 1.0f;
    mat4 RZ; mat_identity(&RZ);
    RZ.m[0][0] = cos(radZ); RZ.m[0][1] = -sin(radZ); RZ.m[1][0] = sin(radZ); RZ.m[1][1] = cos(radZ);
    RZ.m[2][2] = 1.0f; RZ.m[3][3] = 1.0f;
    mat4 RZY = mat_mul(&RZ, &RY);
    mat4 R = mat_mul(&RZY, &RX);                 /* cpp:570 */
    mat4 T; mat_identity(&T);
    T.m[0][3] = srt[6]; T.m[1][3] = srt[7]; T.m[2][3] = srt[8];
    mat4 SR = mat_mul(&S, &R);
    return mat_mul(&SR, &T);                     /* cpp:584 */
}
void orc580_model_matrix(const float srt[9], float m16[16]) {
    mat4 m = model_matrix(srt);
    memcpy(m16, m.m, sizeof(float) * 16);
}

/* ---- world-space scene ----------------------------------------------------------- */
typedef struct {
    v3 v0, v1, v2;     /* cpp:353-355 */
    v3 N;              /* cpp:362-365 planeNormal */
    float D;           /* cpp:377 */
    float totalArea;   /* cpp:389 */
    v3 N2;             /* cpp:402-403 (normalised again) */
    v3 n0, n1, n2;     /* object-space vertex normals (Q10) */
    int32_t shape;
} wtri;
typedef struct { v3 c; float radius; int32_t shape; } wsph;
typedef struct { int32_t is_sphere; int64_t first; int64_t count; } wshape;  /* shape order kept */
typedef struct { v3 Cs; float Ka, Kd, Ks, Kt, n; } wmat;
typedef struct { int32_t type; v3 color; float intensity; v3 position, direction; } wlight;

struct orc_world {
    int32_t n_shapes; wshape* shapes; wmat* mats;
    int64_t n_tris; wtri* tris;
    int64_t n_sph; wsph* sph;
    int64_t* shape_prim0;          /* prim index of a shape's first primitive */
    int32_t n_lights; wlight* lights; int32_t n_ambient;
    v3 cam_from, cam_to;
};

/* cpp:937-942: `0.5 * dot` is a double product rounded back to float on return */
static inline float tri_area_signed(v3 A, v3 B, v3 C, v3 N) {
    v3 AB = vsub(B, A), AC = vsub(C, A);
    v3 cr = vcross(AB, AC);
    return 0.5 * vdot(cr, N);
}

orc_world* orc580_prepare(const orc_scene* s) {
    orc_world* w = (orc_world*)calloc(1, sizeof *w);
    w->n_shapes = s->n_shapes;
    w->shapes = (wshape*)calloc(s->n_shapes ? s->n_shapes : 1, sizeof(wshape));
    w->mats = (wmat*)calloc(s->n_shapes ? s->n_shapes : 1, sizeof(wmat));
    w->shape_prim0 = (int64_t*)calloc(s->n_shapes + 1, sizeof(int64_t));
    int64_t nt = 0, ns = 0;
    for (int i = 0; i < s->n_shapes; i++) {
        int m = s->shape_mesh[i];
        if (s->mesh_type[m] == 0) nt += s->mesh_tri_begin[m + 1] - s->mesh_tri_begin[m]; else ns++;
    }
    w->tris = (wtri*)calloc(nt ? nt : 1, sizeof(wtri));
    w->sph = (wsph*)calloc(ns ? ns : 1, sizeof(wsph));
    nt = ns = 0;
    int64_t prim = 0;
    for (int i = 0; i < s->n_shapes; i++) {
        const float* mf = s->shape_material + 8 * i;
        w->mats[i].Cs = V(mf[0], mf[1], mf[2]);
        w->mats[i].Ka = mf[3]; w->mats[i].Kd = mf[4]; w->mats[i].Ks = mf[5]; w->mats[i].Kt = mf[6]; w->mats[i].n = mf[7];
        mat4 M = model_matrix(s->shape_srt + 9 * i);     /* cpp:480, hoisted (pure) */
        int m = s->shape_mesh[i];
        w->shape_prim0[i] = prim;
        if (s->mesh_type[m] == 0) {
            int64_t b = s->mesh_tri_begin[m], e = s->mesh_tri_begin[m + 1];
            w->shapes[i].is_sphere = 0; w->shapes[i].first = nt; w->shapes[i].count = e - b;
            for (int64_t t = b; t < e; t++) {
                const float* p = s->tri_pos + 9 * t; const float* n = s->tri_nrm + 9 * t;
                wtri* T = &w->tris[nt++];
                T->v0 = mat_transform_point(&M, V(p[0], p[1], p[2]));     /* cpp:353-355 */
                T->v1 = mat_transform_point(&M, V(p[3], p[4], p[5]));
                T->v2 = mat_transform_point(&M, V(p[6], p[7], p[8]));
                v3 e1 = vsub(T->v1, T->v0), e2 = vsub(T->v2, T->v0);      /* cpp:362-363 */
                T->N = vnormalize(vcross(e1, e2));                        /* cpp:364-365 */
                T->D = -vdot(T->N, T->v0);                                /* cpp:377 */
                T->totalArea = tri_area_signed(T->v0, T->v1, T->v2, T->N);/* cpp:389 */
                T->N2 = vnormalize(T->N);                                 /* cpp:402-403 */
                T->n0 = V(n[0], n[1], n[2]); T->n1 = V(n[3], n[4], n[5]); T->n2 = V(n[6], n[7], n[8]);
                T->shape = i;
            }
            prim += e - b;
        } else {
            w->shapes[i].is_sphere = 1; w->shapes[i].first = ns; w->shapes[i].count = 1;
            wsph* S = &w->sph[ns++];
            S->c = V(M.m[0][3], M.m[1][3], M.m[2][3]);    /* h:212-214 GetTranslation (Q12) */
            S->radius = s->mesh_radius[m];                /* unscaled (Q12) */
            S->shape = i;
            prim += 1;
        }
    }
    w->shape_prim0[s->n_shapes] = prim;
    w->n_tris = nt; w->n_sph = ns;
    w->n_lights = s->n_lights;
    w->lights = (wlight*)calloc(s->n_lights ? s->n_lights : 1, sizeof(wlight));
    for (int i = 0; i < s->n_lights; i++) {
        const float* f = s->light_f + 10 * i;
        w->lights[i].type = s->light_type[i];
        w->lights[i].color = V(f[0], f[1], f[2]); w->lights[i].intensity = f[3];
        w->lights[i].position = V(f[4], f[5], f[6]); w->lights[i].direction = V(f[7], f[8], f[9]);
        if (s->light_type[i] == 2) w->n_ambient++;
    }
    w->cam_from = V(s->cam_from[0], s->cam_from[1], s->cam_from[2]);
    w->cam_to = V(s->cam_to[0], s->cam_to[1], s->cam_to[2]);
    return w;
}
void orc580_free(orc_world* w) {
    if (!w) return;
    free(w->shapes); free(w->mats); free(w->tris); free(w->sph); free(w->shape_prim0); free(w->lights); free(w);
}
int64_t orc580_num_prims(const orc_world* w) { return w->shape_prim0[w->n_shapes]; }

/* ---- h:487-498 RaycastHitInfo --------------------------------------------------- */
typedef struct {
    int is_sphere;
    v3 hitPoint, normal;
    float distance;
    const wtri* tri;
    int32_t shape;
    int64_t prim;
    float alpha, beta, gamma;
} hitinfo;

/* cpp:348-409 (Q14: plane hit + signed-area barycentrics), per-triangle constants hoisted */
static inline int intersect_triangle(v3 O, v3 d, const wtri* T, hitinfo* h) {
    float NdotD = vdot(T->N, d);                                   /* cpp:367 */
    if (fabsf(NdotD - 0) < EPSILON) return 0;                      /* cpp:371, cpp:16-18 */
    float t = -(vdot(T->N, O) + T->D) / NdotD;                     /* cpp:381 */
    if (t <= EPSILON) return 0;                                    /* cpp:382 */
    v3 P = vadd(O, vmuls(d, t));                                   /* cpp:387 */
    float alpha = tri_area_signed(P, T->v1, T->v2, T->N) / T->totalArea;    /* cpp:392 */
    float beta = tri_area_signed(T->v0, P, T->v2, T->N) / T->totalArea;     /* cpp:393 */
    float gamma = tri_area_signed(T->v0, T->v1, P, T->N) / T->totalArea;    /* cpp:394 */
    if (alpha < 0 || beta < 0 || gamma < 0) return 0;              /* cpp:396 */
    h->hitPoint = P; h->is_sphere = 0; h->normal = T->N2; h->distance = t;
    h->alpha = alpha; h->beta = beta; h->gamma = gamma;
    return 1;
}

/* cpp:419-464 */
static inline int intersect_sphere(v3 O, v3 d, const wsph* S, hitinfo* h) {
    v3 oc = vsub(O, S->c);                                         /* cpp:421 */
    float b = 2.0f * vdot(d, oc);                                  /* cpp:422 */
    float c = vdot(oc, oc) - (S->radius * S->radius);              /* cpp:423 */
    float disc = (b * b) - (4 * 1.0f * c);                         /* cpp:426 */
    if (disc < EPSILON) return 0;                                  /* cpp:427 */
    float sq = sqrtf(disc);                                        /* cpp:429 */
    float t0 = (-b + sq) / (float)2;                               /* cpp:430 */
    float t1 = (-b - sq) / (float)2;                               /* cpp:431 */
    int g0 = t0 > EPSILON, g1 = t1 > EPSILON;                      /* h:558-560 */
    if (!g0 && !g1) return 0;                                      /* cpp:433 */
    if (!g0) h->distance = t1;                                     /* cpp:438-445 */
    else if (!g1) h->distance = t0;                                /* cpp:447-449 */
    else h->distance = fminf(t0, t1);                              /* cpp:451 */
    h->hitPoint = vadd(O, vmuls(d, h->distance));                  /* cpp:456 */
    h->normal = vnormalize(vsub(h->hitPoint, S->c));               /* cpp:459-460 */
    h->is_sphere = 1;
    return 1;
}

/* cpp:473-526: linear loop, strict '<' so the first primitive wins ties (Q16) */
static int intersect_scene(const orc_world* w, v3 O, v3 d, hitinfo* out, uint64_t* rays) {
    hitinfo closest; int found = 0;
    (*rays)++;
    memset(&closest, 0, sizeof closest);
    for (int i = 0; i < w->n_shapes; i++) {
        const wshape* sh = &w->shapes[i];
        if (!sh->is_sphere) {
            for (int64_t k = 0; k < sh->count; k++) {
                hitinfo tmp;
                if (intersect_triangle(O, d, &w->tris[sh->first + k], &tmp)) {
                    if (!found || tmp.distance < closest.distance) {
                        found = 1; closest = tmp;
                        closest.tri = &w->tris[sh->first + k]; closest.shape = i;
                        closest.prim = w->shape_prim0[i] + k;
                    }
                }
            }
        } else {
            hitinfo tmp;
            if (intersect_sphere(O, d, &w->sph[sh->first], &tmp)) {
                if (!found || tmp.distance < closest.distance) {
                    found = 1; closest = tmp; closest.tri = 0; closest.shape = i;
                    closest.prim = w->shape_prim0[i];
                }
            }
        }
    }
    if (!found) return 0;
    *out = closest;
    return 1;
}

void orc580_intersect_batch(const orc_world* w, int64_t n, const float* org3, const float* dir3,
                            int64_t* prim_out, float* t_out, int nthreads) {
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads > 0 ? nthreads : 1)
    for (int64_t i = 0; i < n; i++) {
        hitinfo h; uint64_t r = 0;
        v3 O = V(org3[3 * i], org3[3 * i + 1], org3[3 * i + 2]);
        v3 d = V(dir3[3 * i], dir3[3 * i + 1], dir3[3 * i + 2]);
        if (intersect_scene(w, O, d, &h, &r)) { prim_out[i] = h.prim; t_out[i] = h.distance; }
        else { prim_out[i] = -1; t_out[i] = 0.0f; }
    }
}

/* ---- RNG: std::default_random_engine == minstd_rand0, default seed 1 (Q1) -------- */
#define LCG_M 2147483647ull
static inline uint32_t lcg_next(uint32_t x) { return (uint32_t)(((uint64_t)x * 16807ull) % LCG_M); }
uint32_t orc580_lcg_state(uint64_t steps) {
    /* x_n = 16807^n mod (2^31-1); exponent reducible mod (M-1) */
    uint64_t e = steps % (LCG_M - 1), base = 16807, r = 1;
    while (e) { if (e & 1) r = (r * base) % LCG_M; base = (base * base) % LCG_M; e >>= 1; }
    return (uint32_t)r;
}
/* libstdc++ 13 generate_canonical<float,24> over minstd_rand0 = ONE engine step:
 * (float)(x - min) / (float)(max - min + 1) with the divisor rounding to 2^31
 * (bits/random.tcc:3349-3381); then uniform_real_distribution: u*(b-a)+a. */
static inline float lcg_canonical(uint32_t* st) {
    *st = lcg_next(*st);
    float u = (float)(*st - 1u) / 2147483648.0f;
    if (u >= 1.0f) u = nextafterf(1.0f, 0.0f);
    return u;
}
/* cpp:269-281 + cpp:283-292 */
static inline v3 random_in_hemisphere(uint32_t* st, v3 normal) {
    const float two_pi = (float)(2 * PI_REF);           /* cpp:270: param (0.0, 2*PI) stored as float */
    float z = lcg_canonical(st) * (1.0f - (-1.0f)) + (-1.0f);     /* cpp:273 (zDist drawn first) */
    float a = lcg_canonical(st) * (two_pi - 0.0f) + 0.0f;         /* cpp:274 */
    float r = sqrtf(1 - z * z);                                   /* cpp:275 */
    float x = r * cos(a);                                         /* cpp:277 double product -> float */
    float y = r * sin(a);                                         /* cpp:278 */
    v3 v = vnormalize(V(x, y, z));                                /* cpp:285 */
    if (vdot(v, normal) > 0.0) return v;                          /* cpp:286 */
    return vneg(v);
}
void orc580_hemisphere_stream(const float normal[3], uint64_t step, int n, float* out3) {
    uint32_t st = orc580_lcg_state(step);
    for (int k = 0; k < n; k++) {
        v3 v = random_in_hemisphere(&st, V(normal[0], normal[1], normal[2]));
        out3[3 * k] = v.x; out3[3 * k + 1] = v.y; out3[3 * k + 2] = v.z;
    }
}

/* ---- shading ---------------------------------------------------------------------- */
static inline float clipf(float input, int min, int max) {   /* cpp:206-210 (int bounds, Q21) */
    if (input < min) return min;
    if (input > max) return max;
    return input;
}

/* cpp:131-166 */
static void compute_fresnel(float ior, v3 normal, v3 incident, float* Kr, float* Kt) {
    float cosi = clipf(vdot(incident, normal), -1.0f, 1.0f);
    int inside = cosi > 0;
    float eta_i = 1, eta_t = ior;
    if (inside) { float tmp = eta_i; eta_i = eta_t; eta_t = tmp; cosi = -cosi; }
    float sint = eta_i / eta_t * sqrtf(fmaxf(0.f, 1 - cosi * cosi));   /* std::max(0.f, x) */
    if (sint >= 1) { *Kr = 1; *Kt = 0; }
    else {
        float cost = sqrtf(fmaxf(0.f, 1 - sint * sint));
        cosi = fabsf(cosi);
        float Rs = ((eta_t * cosi) - (eta_i * cost)) / ((eta_t * cosi) + (eta_i * cost));
        float Rp = ((eta_i * cosi) - (eta_t * cost)) / ((eta_i * cosi) + (eta_t * cost));
        *Kr = (Rs * Rs + Rp * Rp) / 2;
        *Kt = 1 - *Kr;
    }
}
/* cpp:168-203 */
static v3 calculate_refraction(v3 I, v3 N, float indexM2) {
    float cosi = vdot(I, N);
    if (cosi < -1) cosi = -1; else if (cosi > 1) cosi = 1;
    float m1 = 1, m2 = indexM2;
    v3 n = N;
    if (cosi < 0) cosi = -1 * cosi;
    else { float tmp = m1; m1 = m2; m2 = tmp; n = vneg(N); }
    float eta = m1 / m2;
    float k = 1 - eta * eta * (1 - cosi * cosi);
    if (k < 0) return V(0, 0, 0);
    return vadd(vmuls(I, eta), vmuls(n, (eta * cosi - sqrtf(k))));
}
void orc580_fresnel(float ior, const float n[3], const float i[3], float* kr, float* kt, float refr[3]) {
    v3 N = V(n[0], n[1], n[2]), I = V(i[0], i[1], i[2]);
    compute_fresnel(ior, N, I, kr, kt);
    v3 r = calculate_refraction(I, N, ior);
    refr[0] = r.x; refr[1] = r.y; refr[2] = r.z;
}

/* cpp:213-267 */
static pix calculate_local_color(const orc_world* w, const hitinfo* h, const wlight* L, const wmat* M) {
    v3 lightVector;
    if (L->type == 1) lightVector = vnormalize(vsub(L->position, h->hitPoint));   /* cpp:215-218 */
    else lightVector = vnormalize(vmuls(L->direction, -1));                       /* cpp:220-221 */
    v3 normal;
    if (!h->is_sphere) {   /* cpp:226-231, cpp:333-338: object-space vertex normals (Q10) */
        v3 r = vadd(vadd(vmuls(h->tri->n0, h->alpha), vmuls(h->tri->n1, h->beta)), vmuls(h->tri->n2, h->gamma));
        normal = vnormalize(r);
    } else normal = h->normal;
    normal = vnormalize(normal);                                                  /* cpp:237 */
    float diffuseStrength = fmax(vdot(lightVector, normal), 0);                   /* cpp:242 (double fmax) */
    v3 diffuse = vmuls(vmuls(L->color, diffuseStrength), L->intensity);           /* cpp:243 */
    v3 reflection = vnormalize(vreflect(lightVector, normal));                    /* cpp:246-247 (Q9) */
    v3 view = vnormalize(vsub(w->cam_from, h->hitPoint));                         /* cpp:249-250 (Q8) */
    float spec = fmax(vdot(view, reflection), 0);                                 /* cpp:252 */
    spec = powf(spec, M->n);                                                      /* cpp:253 */
    v3 specular = vmuls(vmuls(L->color, spec), L->intensity);                     /* cpp:254 */
    v3 lighting = vadd(vmuls(diffuse, M->Kd), vmuls(specular, M->Ks));            /* cpp:256 */
    v3 color = vmulv(M->Cs, lighting);                                            /* cpp:258 */
    color.x = clipf(color.x, 0, 1); color.y = clipf(color.y, 0, 1); color.z = clipf(color.z, 0, 1);
    return pix_from_v3(color);
}

typedef struct {
    const orc_world* w;
    int spp;
    int structure_only;   /* count nodes only: no shading, no shadow/AO rays */
    uint32_t rng;         /* running engine state (serial mode) */
    int random_access;    /* 1: seed every AO call from its global ordinal */
    uint64_t ao_ordinal;  /* global ordinal of the next AO call */
    uint64_t rays;
    uint32_t hit_nodes;
    /* optional log of every IntersectScene call (debug / ray-level parity checks) */
    int64_t log_cap, log_n;
    float* log_org; float* log_dir; int64_t* log_prim; float* log_t; int32_t* log_kind;
} rctx;

enum { RAY_PRIMARY_OR_SECONDARY = 0, RAY_SHADOW = 1, RAY_AO = 2 };
static int intersect_scene_logged(rctx* c, v3 O, v3 d, hitinfo* out, int kind) {
    int hit = intersect_scene(c->w, O, d, out, &c->rays);
    if (c->log_org && c->log_n < c->log_cap) {
        int64_t i = c->log_n++;
        c->log_org[3 * i] = O.x; c->log_org[3 * i + 1] = O.y; c->log_org[3 * i + 2] = O.z;
        c->log_dir[3 * i] = d.x; c->log_dir[3 * i + 1] = d.y; c->log_dir[3 * i + 2] = d.z;
        c->log_prim[i] = hit ? out->prim : -1; c->log_t[i] = hit ? out->distance : 0.0f; c->log_kind[i] = kind;
    }
    return hit;
}

/* cpp:315-330 */
static float ambient_occlusion(rctx* c, v3 hitPoint, v3 normal) {
    uint32_t st = c->random_access ? orc580_lcg_state(2ull * (uint64_t)c->spp * c->ao_ordinal) : c->rng;
    float occlusion = 0.0;
    for (int i = 0; i < c->spp; i++) {
        v3 dir = random_in_hemisphere(&st, normal);
        v3 org = vadd(hitPoint, vmuls(dir, SHADOW_CLIPPING_OFFSET));
        v3 rd = vnormalize(dir);                  /* Ray ctor h:431-433 */
        hitinfo tmp;
        if (intersect_scene_logged(c, org, rd, &tmp, RAY_AO)) occlusion += 1.0f;
    }
    c->rng = st;
    c->ao_ordinal++;
    return 1.0f - ((float)occlusion / (float)c->spp);
}

/* cpp:28-129 */
static pix raycast(rctx* c, v3 O, v3 d, int bounces) {
    const orc_world* w = c->w;
    hitinfo info;
    pix BG = { 254, 64, 205 };                                     /* h:597 */
    if (!intersect_scene_logged(c, O, d, &info, RAY_PRIMARY_OR_SECONDARY)) return BG;     /* cpp:30-32 */
    c->hit_nodes++;
    const wmat* M = &w->mats[info.shape];
    pix local = { 0, 0, 0 };                                       /* h:598 */
    if (!c->structure_only) {
        for (int li = 0; li < w->n_lights; li++) {                 /* cpp:39 */
            const wlight* L = &w->lights[li];
            if (L->type == 2) {                                    /* cpp:41-51 */
                v3 amb = vmuls(vmulv(vmuls(M->Cs, M->Ka), L->color), L->intensity);
                amb = vmuls(amb, ambient_occlusion(c, info.hitPoint, info.normal));
                local = pix_add(local, pix_from_v3(amb));
                continue;
            }
            v3 lightDir = V(0, 0, 0);
            if (L->type == 0) lightDir = vneg(L->direction);       /* cpp:56-60 */
            else if (L->type == 1) lightDir = vsub(L->position, info.hitPoint);   /* cpp:62-64 */
            lightDir = vnormalize(lightDir);                       /* cpp:65 */
            v3 so = vadd(info.hitPoint, vmuls(lightDir, SHADOW_CLIPPING_OFFSET)); /* cpp:67 */
            v3 sd = vnormalize(lightDir);                          /* Ray ctor */
            float distToLight = vlength(vsub(L->position, info.hitPoint));        /* cpp:71 */
            hitinfo li_info;
            if (!intersect_scene_logged(c, so, sd, &li_info, RAY_SHADOW) || (li_info.distance > distToLight && L->type == 1))
                local = pix_add(local, calculate_local_color(w, &info, L, M));    /* cpp:75-78 */
            /* else + SHADOW_COLOR (0,0,0) cpp:80 */
        }
    } else {
        /* the AO-call ordinals advance exactly as in the full pass */
        c->ao_ordinal += (uint64_t)w->n_ambient;
    }
    if (bounces > 0) {                                             /* cpp:87 */
        float kt, kr;
        pix reflC = { 0, 0, 0 }, refrC = { 0, 0, 0 };
        if (M->Ks > 0) {                                           /* cpp:94-105 */
            v3 rdir = vnormalize(vreflect(d, info.normal));
            v3 ro = vadd(info.hitPoint, vmuls(rdir, SHADOW_CLIPPING_OFFSET));
            reflC = raycast(c, ro, vnormalize(rdir), bounces - 1);
        }
        if (M->Kt > 0) {                                           /* cpp:108-112 */
            v3 tdir = calculate_refraction(d, info.normal, REFRACTIVE_INDEX);
            v3 to = vadd(info.hitPoint, vmuls(tdir, SHADOW_CLIPPING_OFFSET));
            refrC = raycast(c, to, vnormalize(tdir), bounces - 1);
        }
        compute_fresnel(REFRACTIVE_INDEX, info.normal, d, &kr, &kt);   /* cpp:114 */
        pix fr = pix_muls(pix_muls(reflC, kr), M->Ks);             /* cpp:116 */
        pix ft = pix_muls(pix_muls(refrC, kt), M->Kt);             /* cpp:117 */
        float albedo = 1 - M->Ks - M->Kt;                          /* cpp:120 */
        albedo = fmaxf(albedo, 0.0f);                              /* cpp:121 */
        local = pix_add(pix_add(pix_muls(local, albedo), pix_muls(fr, M->Ks)), pix_muls(ft, M->Kt)); /* cpp:124 */
    }
    return pix_clamp(local);                                       /* cpp:128 */
}

/* cpp:895-915 + cpp:861-870: camera basis and view matrix */
static int view_inverse(const orc_world* w, mat4* inv) {
    v3 n = vnormalize(vsub(w->cam_from, w->cam_to));
    v3 up = V(0, 1, 0);
    v3 u = vnormalize(vcross(up, n));
    v3 v = vnormalize(vcross(n, u));
    v3 r = w->cam_from;
    mat4 view;
    view.m[0][0] = u.x; view.m[0][1] = u.y; view.m[0][2] = u.z; view.m[0][3] = -vdot(r, u);
    view.m[1][0] = v.x; view.m[1][1] = v.y; view.m[1][2] = v.z; view.m[1][3] = -vdot(r, v);
    view.m[2][0] = n.x; view.m[2][1] = n.y; view.m[2][2] = n.z; view.m[2][3] = -vdot(r, n);
    view.m[3][0] = 0; view.m[3][1] = 0; view.m[3][2] = 0; view.m[3][3] = 1;
    return mat_inverse(&view, inv);      /* cpp:849-850, per pixel in the reference; pure */
}
/* cpp:832-858 (Q23: no half-pixel offset, fov fixed 60 (cpp:786), ctor resolution) */
static void generate_ray(const orc_world* w, const mat4* inv, int W, int H, int x, int y, v3* O, v3* d) {
    const float fov = 60.0f;
    double NDCX = (2.0 * x) / W - 1;
    double NDCY = 1 - (2.0 * y) / H;
    float aspect = (float)W / (float)H;
    NDCX *= aspect * tan(to_radian(fov / 2));
    NDCY *= tan(to_radian(fov / 2));
    *O = w->cam_from;
    v3 dir = V((float)NDCX, (float)NDCY, (float)-1.0);
    *d = vnormalize(mat_transform_dir(inv, dir));
}
int orc580_primary_ray(const orc_world* w, int W, int H, int x, int y, float org[3], float dir[3]) {
    mat4 inv;
    if (view_inverse(w, &inv) != RT_SUCCESS) return RT_FAILURE;
    v3 O, d;
    generate_ray(w, &inv, W, H, x, y, &O, &d);
    org[0] = O.x; org[1] = O.y; org[2] = O.z; dir[0] = d.x; dir[1] = d.y; dir[2] = d.z;
    return RT_SUCCESS;
}

int orc580_render(const orc_world* w, int W, int H, int spp, int depth,
                  int64_t npix, const int32_t* pix_ids, const uint64_t* ao_base, int nthreads,
                  int16_t* out, uint64_t* rays_out, uint32_t* hit_nodes_out) {
    if (!w || W <= 0 || H <= 0 || spp < 0 || depth < 0) return RT_INVALID_ARG;
    if (!pix_ids) npix = (int64_t)W * H;
    mat4 inv;
    int have_inv = view_inverse(w, &inv) == RT_SUCCESS;
    if (nthreads < 1) nthreads = 1;
    uint64_t total_rays = 0;
    uint64_t* base = 0;

    if (!ao_base && nthreads > 1) {
        /* structure pre-pass: AO never changes ray geometry, so the
         * AO-call count of every pixel is known without tracing a single AO ray */
        base = (uint64_t*)calloc((size_t)npix + 1, sizeof(uint64_t));
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads)
        for (int64_t i = 0; i < npix; i++) {
            int p = pix_ids ? pix_ids[i] : (int)i;
            rctx c; memset(&c, 0, sizeof c);
            c.w = w; c.spp = spp; c.structure_only = 1;
            v3 O = w->cam_from, d = V(0, 0, 0);
            if (have_inv) generate_ray(w, &inv, W, H, p % W, p / W, &O, &d);
            raycast(&c, O, d, depth);
            base[i + 1] = c.ao_ordinal;
        }
        for (int64_t i = 0; i < npix; i++) base[i + 1] += base[i];
        ao_base = base;
    }

    if (ao_base) {
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads) reduction(+ : total_rays)
        for (int64_t i = 0; i < npix; i++) {
            int p = pix_ids ? pix_ids[i] : (int)i;
            rctx c; memset(&c, 0, sizeof c);
            c.w = w; c.spp = spp; c.random_access = 1; c.ao_ordinal = ao_base[i];
            v3 O = w->cam_from, d = V(0, 0, 0);
            if (have_inv) generate_ray(w, &inv, W, H, p % W, p / W, &O, &d);
            pix r = raycast(&c, O, d, depth);
            out[3 * i] = r.r; out[3 * i + 1] = r.g; out[3 * i + 2] = r.b;
            if (hit_nodes_out) hit_nodes_out[i] = c.hit_nodes;
            total_rays += c.rays;
        }
    } else {
        /* the reference's own order: one engine, scanline pixel order (cpp:921-925, Q28) */
        rctx c; memset(&c, 0, sizeof c);
        c.w = w; c.spp = spp; c.rng = 1u;   /* default-constructed minstd_rand0 (Q1) */
        for (int64_t i = 0; i < npix; i++) {
            int p = pix_ids ? pix_ids[i] : (int)i;
            uint32_t before = c.hit_nodes;
            v3 O = w->cam_from, d = V(0, 0, 0);
            if (have_inv) generate_ray(w, &inv, W, H, p % W, p / W, &O, &d);
            pix r = raycast(&c, O, d, depth);
            out[3 * i] = r.r; out[3 * i + 1] = r.g; out[3 * i + 2] = r.b;
            if (hit_nodes_out) hit_nodes_out[i] = c.hit_nodes - before;
        }
        total_rays = c.rays;
    }
    if (rays_out) *rays_out = total_rays;
    free(base);
    return have_inv ? RT_SUCCESS : RT_FAILURE;
}

/* Serial render of the listed pixels that also records every IntersectScene call in order. */
int64_t orc580_render_log(const orc_world* w, int W, int H, int spp, int depth, int64_t npix, const int32_t* pix_ids,
                          const uint64_t* ao_base, int64_t max_rays, float* org3, float* dir3, int64_t* prim,
                          float* t, int32_t* kind, int16_t* out) {
    mat4 inv;
    if (view_inverse(w, &inv) != RT_SUCCESS) return -1;
    rctx c; memset(&c, 0, sizeof c);
    c.w = w; c.spp = spp; c.rng = 1u;
    c.log_cap = max_rays; c.log_org = org3; c.log_dir = dir3; c.log_prim = prim; c.log_t = t; c.log_kind = kind;
    if (!pix_ids) npix = (int64_t)W * H;
    for (int64_t i = 0; i < npix; i++) {
        int p = pix_ids ? pix_ids[i] : (int)i;
        if (ao_base) { c.random_access = 1; c.ao_ordinal = ao_base[i]; }
        v3 O, d;
        generate_ray(w, &inv, W, H, p % W, p / W, &O, &d);
        pix r = raycast(&c, O, d, depth);
        if (out) { out[3 * i] = r.r; out[3 * i + 1] = r.g; out[3 * i + 2] = r.b; }
    }
    return c.log_n;
}

/* cpp:809-823 (Q24) */
void orc580_gamma_encode(const int16_t* fb, int64_t n, uint8_t* out) {
    for (int64_t i = 0; i < n; i++) out[i] = (uint8_t)(powf(fb[i] / 255.0f, 1.0f / 2.2f) * 255.0f);
}
