// Host-side interface of the device scene build (bvh_build.cu).
#pragma once
#include "device_scene.h"
#include <cstddef>

namespace rt580 {

// Bump allocator over one device block that only ever grows.  The scene build used to cudaMalloc /
// cudaFree ~25 buffers per upload; inside a process that also hosts another CUDA allocator
// (PyTorch in bench.py) those calls took 0.2 - 3 s per upload.  Two arenas owned by the context
// (temporaries, scene) make a re-upload allocation-free.
struct DevArena {
    char* base = nullptr;
    size_t cap = 0, off = 0;
    void reset() { off = 0; }
    template <typename T> T* take(size_t n) {
        const size_t bytes = ((n ? n : 1) * sizeof(T) + 255) & ~(size_t)255;
        if (off + bytes > cap) return nullptr;
        T* p = reinterpret_cast<T*>(base + off);
        off += bytes;
        return p;
    }
};
// make sure the arena holds at least `bytes` (contents are lost when it has to grow) and reset it
bool arena_reserve(DevArena& a, size_t bytes, char* err, size_t errlen);
void arena_release(DevArena& a);
// upper bounds of what build_bvh takes from the two arenas for n_in input primitives
size_t build_tmp_bytes(int64_t n_in);
size_t build_out_bytes(int64_t n_in, int64_t n_prims);

// grow-only device buffer
template <typename T> struct DBuf {
    T* p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t n, size_t keep, cudaStream_t s) {
        if (n <= cap) return cudaSuccess;
        size_t ncap = cap ? cap : 1024;
        while (ncap < n) ncap = ncap + ncap / 2 + 1024;
        T* q = nullptr;
        cudaError_t e = cudaMalloc((void**)&q, ncap * sizeof(T));
        if (e != cudaSuccess) return e;
        if (keep && p) { e = cudaMemcpyAsync(q, p, keep * sizeof(T), cudaMemcpyDeviceToDevice, s); if (e != cudaSuccess) return e; }
        if (p) { cudaStreamSynchronize(s); cudaFree(p); }
        p = q; cap = ncap;
        return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct BuildInput {            // device pointers (already uploaded)
    const float4* tri_v0; const float4* tri_v1; const float4* tri_v2;
    const int32_t* tri_prim; int64_t n_tris;
    const float4* sph; const int32_t* sph_prim; int64_t n_spheres;
    float origin_hint[3];      // camera position: bounds |ray origin| for the box padding
    int64_t n_prims;           // primitives in reference order
    int64_t first_tri;         // input index of the triangle that comes first in reference order (-1: no triangle): a ray with a NaN in
                               // it "hits" that one (cpp:371, 382, 396 all compare false), whatever its area
};

struct BuildOutput {
    PrimRec* prims;            // [n_leaf] Morton order (all outputs live in the `out` arena)
    BvhNode* nodes;            // [max(n_leaf-1,1)]
    float4* far;               // [n_leaf] far-field filter records
    float far_tmin;            // min over triangles of T_far
    int n_always;              // triangles that are far-field candidates for every ray (slivers)
    int32_t* always_idx;       // [n_always] their indices into prims
    int32_t* leaf_of_prim;     // [n_prims] inverse of the Morton permutation (-1 for dropped triangles)
    int n_leaf;                // primitives in the tree
    int n_big;                 // large primitives after them in `prims` / `far`
    int nan_leaf;              // prims[nan_leaf]: the record of that first triangle, kept outside the tree and the lists (-1: none)
    int n_dropped;             // zero-area triangles
    unsigned int max_depth;
    float pad, extent;
    float pad_max;             // largest padding of a leaf box (pad + L_near): every near-field hit lies within it of its primitive
    float bounds_lo[3], bounds_hi[3];
    int launches;
};

bool build_bvh(const BuildInput& in, BuildOutput* out, DevArena& tmp, DevArena& outa, cudaStream_t stream, char* err,
               size_t errlen);

// ---- far-field direction grid (fargrid.cuh, fargrid_build.cu) ----
struct FgBuildInput {
    const PrimRec* prims; const float4* far_old; int n_all;
    int K;                      // cells per cube-face edge, 0: per-triangle constants only (no grid)
    float extent;
    float ob_lo[3], ob_hi[3], cam[3];
    float4* fgA; float2* fgB; uint32_t* wide; uint32_t* sph;   // [n_all] each (scene arena)
    unsigned int* counters;                         // [4] scratch
    DBuf<unsigned int>* counts; DBuf<unsigned long long>* start; DBuf<unsigned long long>* bsum; DBuf<uint32_t>* entries;
    DBuf<unsigned int>* cell_tmin;                  // [6 K K] smallest T per cell (float bits)
};
struct FgBuildOutput {
    int K; int n_wide; int n_sph; unsigned long long n_entries; float t_min; float diag;
};
int fg_default_K(long long n_tris);
bool fg_build(const FgBuildInput& in, FgBuildOutput* out, cudaStream_t stream, char* err, size_t errlen);

}  // namespace rt580
