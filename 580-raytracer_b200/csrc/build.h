// Host-side interface of the device scene build (bvh_build.cu).
#pragma once
#include "device_scene.h"
#include <cstddef>

namespace rt580 {

struct BuildInput {            // device pointers (already uploaded)
    const float4* tri_v0; const float4* tri_v1; const float4* tri_v2;
    const int32_t* tri_prim; int64_t n_tris;
    const float4* sph; const int32_t* sph_prim; int64_t n_spheres;
    float origin_hint[3];      // camera position: bounds |ray origin| for the box padding
    int64_t n_prims;           // primitives in reference order
};

struct BuildOutput {
    PrimRec* prims;            // [n_leaf] Morton order (cudaMalloc'ed, caller frees)
    BvhNode* nodes;            // [max(n_leaf-1,1)]
    float4* far;               // [n_leaf] far-field filter records
    float far_tmin;            // min over triangles of T_far
    int n_always;              // triangles that are far-field candidates for every ray (slivers)
    int32_t* always_idx;       // [n_always] their indices into prims
    int32_t* leaf_of_prim;     // [n_prims] inverse of the Morton permutation (-1 for dropped triangles)
    int n_leaf;
    int n_dropped;             // zero-area triangles
    unsigned int max_depth;
    float pad, extent;
    float bounds_lo[3], bounds_hi[3];
    int launches;
};

bool build_bvh(const BuildInput& in, BuildOutput* out, cudaStream_t stream, char* err, size_t errlen);

}  // namespace rt580
