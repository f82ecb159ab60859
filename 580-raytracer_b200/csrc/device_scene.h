// Device-resident scene: what rt580_upload_scene leaves in HBM.
//
// Layout (64 B per triangle record, 64 B per 2-child node):
//   PrimRec[n_leaf]   Morton-sorted primitive records, 4 x float4 each, so a leaf visit is
//                     four coalescable 16-byte loads.  Zero-area triangles are dropped
//                     (N == 0 => |N.d| < EPSILON always, Raytracer.cpp:365-373).
//   BvhNode[n_leaf-1] LBVH inner nodes, both child boxes inline (Aila-Laine layout).
//   vn / prim_material / materials / lights : shading data addressed by the reference's
//                     primitive ORDER index (shape order, triangle order), which is also
//                     the tie-break key for equal distances (cpp:494, cpp:513).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rt580 {

#define RT_PRIM_TRIANGLE 0
#define RT_PRIM_SPHERE   1
#define RT_PRIM_SLOWPATH 2   // flag bit: triangle whose totalArea is outside [1e-30,1e30] -> exact division path

// triangle: a = (v0.xyz, D)  b = (v1.xyz, totalArea)  c = (v2.xyz, bits(prim order))  d = (N.xyz, bits(type|flags))
//           v* world space (cpp:353-355), N = normalize(cross(E1,E2)) (cpp:362-365),
//           D = -dot(N, v0) (cpp:377), totalArea (cpp:389)
// sphere  : a = (centre.xyz, radius) b = (r*r,0,0,0) c = (0,0,0, bits(prim order)) d = (0,0,0, bits(type))
struct __align__(16) PrimRec { float4 a, b, c, d; };

// kids.x / kids.y: >= 0 inner node index, < 0 leaf (~index into PrimRec)
struct __align__(16) BvhNode {
    float4 xy0;   // child0: lo.x hi.x lo.y hi.y
    float4 xy1;   // child1: lo.x hi.x lo.y hi.y
    float4 z01;   // child0 lo.z hi.z, child1 lo.z hi.z
    int4 kids;    // child0, child1, unused, unused
};

struct DeviceScene {
    const PrimRec* prims;
    const BvhNode* nodes;
    const float4* far;         // [n_leaf] far-field filter records (N.xyz, thr), see bvh_build.cu
    float far_tmin;            // no far-field acceptance is possible at t < far_tmin (sliver list aside)
    const int32_t* leaf_of_prim; // [n_prims] index into prims of a primitive order index (-1: dropped)
    const int32_t* always_idx; // [n_always] sliver triangles whose far field starts inside the scene:
    int32_t n_always;          //            tested exactly for every ray
    float extent;              // E: the boxes and far-field records are valid for ray origins with |coordinate| <= E
    unsigned int* diag;        // [2] device counters: far-field scans, linear fallbacks (rare events)
    int32_t farfield;          // 1: replay the reference's far-field acceptances (exact), 0: skip
    int32_t n_leaf;            // primitives in the BVH: prims[0, n_leaf)
    int32_t n_big;             // very large primitives kept out of the tree: prims[n_leaf, n_leaf + n_big), tested
                               // first for every ray (they would bloat every ancestor box of an LBVH, and as
                               // the likeliest occluders they end most any-hit rays at once)
    // the distinct planes of the large triangles (a wall is two coplanar triangles): the plane part of the
    // reference's triangle test (cpp:367-382) is the same arithmetic for every triangle of a plane, so
    // it is done once per plane and rejects all its triangles at once
    const float4* big_planes;               // [n_big_planes] (N.xyz, D), bit-identical in all member records
    const unsigned long long* big_masks;    // [n_big_planes] members: bit k = prims[n_leaf + k]
    const int32_t* big_plane_newn;          // [n_big_planes] 1: another normal than the plane before (planes are sorted
                                            //                by normal: parallel walls share N.d and N.O in the ordering pass)
    unsigned long long big_sphere_mask;     // large spheres (no plane): bit k = prims[n_leaf + k]
    int32_t n_big_planes;
    // A box no large primitive reaches into (the inside of a room, the space above a floor), found at upload.  A
    // shadow ray that starts inside it and runs to a point light inside it stays inside it (convex): the large
    // primitives cannot stop it and their plane tests are skipped.  big_free_light: per light, the light and the
    // 0.2 units by which the reference's ray overshoots it lie inside.  big_free_on = 0: no such box.
    float big_free_lo[3], big_free_hi[3];
    const int32_t* big_free_light;          // [n_lights]
    int32_t big_free_on;
    int32_t nan_leaf, nan_prim; // a ray with a NaN in it: every compare of cpp:371 / 382 / 396 is false, the reference's loop keeps the FIRST
                               // triangle of the scene (cpp:487-499), at t = NaN.  prims[nan_leaf] is its record (outside tree and lists), -1: no triangle
    int32_t n_all;             // n_leaf + n_big: what the linear loops and the far-field scan walk
    int32_t n_prims;           // primitives in reference order (incl. dropped ones)
    const float4* vn;          // [n_prims][3] object-space vertex normals (Q10); unused for spheres
    const int32_t* prim_material;   // [n_prims]
    const float* materials;    // [n_materials][8] Cs.rgb Ka Kd Ks Kt n
    int32_t n_materials;
    const int32_t* light_type; // [n_lights]
    const float* light_f;      // [n_lights][10]
    int32_t n_lights;
    int32_t n_ambient;
    int32_t n_nonambient;
    // Point-light clearance maps (smap.cuh): per point light a cube map of a LOWER bound of the distance
    // from the light to the nearest padded leaf box in each direction.  A shadow ray whose far end
    // (its origin, seen from the light) is nearer than that bound cannot meet a tree primitive: it
    // skips the traversal.  Pure culling, like the boxes of the tree itself.
    // Far-field direction grid (fargrid.cuh): per cell of a cube map of RAY DIRECTIONS, the triangles a ray of that
    // direction that leaves the scene can still "hit" 10^5..10^8 units away.  fg_K == 0: no grid (the O(n) scan runs).
    const float4* fg_A;                     // [n_all] (N.xyz, D)   (-: spheres carry zeros)
    const float2* fg_B;                     // [n_all] (T: lower bound of the ray parameter of a far-field hit, < 0: none; unused)
    const unsigned long long* fg_start;     // [6 K K + 1]
    const uint32_t* fg_entries;             // (k6 << 26) | index into prims: T for the cell = T * 2^(k6/4)
    const unsigned int* fg_cell_tmin;       // [6 K K] float bits: the smallest T among the cell's entries (all ones: empty cell)
    const uint32_t* fg_wide;                // [fg_n_wide] triangles whose far field begins too near for a direction index:
    int32_t fg_n_wide;                      //             every ray that leaves the scene filters them
    int32_t fg_K;
    float fg_dmax;                          // bound on |N.O + D| over in-scene ray origins
    const uint32_t* fg_sph;                 // [fg_n_sph] the spheres among prims (rays from outside the scene test them all when aimed at it)
    int32_t fg_n_sph;
    float fg_rmax;                          // largest sphere radius
    float fg_center[3];                     // a point of the origin box: directions of far hit points are taken from here
    float fg_tmin;                          // smallest T of the grid's triangles
    // in-scene ray origins (fargrid.cuh in_scene): box around every primitive + padding + 0.25, and the camera
    float ob_lo[3], ob_hi[3], ob_cam[3];
    const float* smap;             // [n_smap][6][smap_res][smap_res]
    int32_t smap_res;
    const int32_t* smap_of_light;  // [n_lights] map index, -1: none (not a point light, or a primitive too close to it)
    int32_t n_smap;
};

}  // namespace rt580
