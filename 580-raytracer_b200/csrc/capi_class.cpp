// C ABI mirror of the Raytracer class (include/rt580.h, "class mirror"): lets a foreign
// runtime (ctypes in tests/ and bench.py, cgo/JNI elsewhere) drive the same object the C++
// main() drives.  Every entry point maps 1:1 onto a reference method (Raytracer.h:557-588).
#include "raytracer.h"
#include <cstring>
#include <exception>

struct rt580_raytracer { Raytracer* rt; };

extern "C" {

rt580_raytracer* rt580_raytracer_new(int width, int height) {
    try { return new rt580_raytracer{ new Raytracer(width, height) }; } catch (...) { return nullptr; }
}
void rt580_raytracer_delete(rt580_raytracer* h) {
    if (!h) return;
    delete h->rt;
    delete h;
}
int rt580_raytracer_set_assets_path(rt580_raytracer* h, const char* dir) {
    if (!h || !dir) return RT_INVALID_ARG;
    std::string d(dir);
    if (!d.empty() && d.back() != '/') d += '/';
    h->rt->SetAssetsPath(d);
    return RT_SUCCESS;
}
int rt580_raytracer_set_options(rt580_raytracer* h, int depth, int ao_spp, int rng_mode, int traversal, int device,
                                int farfield) {
    if (!h) return RT_INVALID_ARG;
    h->rt->SetBounces(depth); h->rt->SetAmbientOcclusionSamples(ao_spp); h->rt->SetRngMode(rng_mode);
    h->rt->SetTraversal(traversal); h->rt->SetDevice(device); h->rt->SetFarField(farfield);
    return RT_SUCCESS;
}
int rt580_raytracer_set_quiet(rt580_raytracer* h, int quiet) {
    if (!h) return RT_INVALID_ARG;
    h->rt->SetQuiet(quiet != 0);
    return RT_SUCCESS;
}
int rt580_raytracer_set_gpus(rt580_raytracer* h, int n_gpus) {
    if (!h || n_gpus < 1) return RT_INVALID_ARG;
    h->rt->SetGpus(n_gpus);
    return RT_SUCCESS;
}
int rt580_raytracer_set_mesh_cache(rt580_raytracer* h, const char* dir) {
    if (!h) return RT_INVALID_ARG;
    h->rt->SetMeshCacheDir(dir ? dir : "");
    return RT_SUCCESS;
}
int rt580_raytracer_set_device_flatten(rt580_raytracer* h, int on) {
    if (!h) return RT_INVALID_ARG;
    h->rt->SetDeviceFlatten(on != 0);
    return RT_SUCCESS;
}
int rt580_raytracer_instanced_scene(rt580_raytracer* h, rt580_instanced_scene* out) {
    if (!h || !out) return RT_INVALID_ARG;
    return h->rt->GetInstancedScene(out);
}
int rt580_raytracer_mesh_cache_hits(rt580_raytracer* h) { return h ? h->rt->MeshCacheHits() : -1; }
int rt580_raytracer_load_scene_json(rt580_raytracer* h, const char* scene) {
    if (!h || !scene) return RT_INVALID_ARG;
    try { return h->rt->LoadSceneJSON(scene); } catch (...) { return RT_FAILURE; }
}
int rt580_raytracer_render(rt580_raytracer* h, const char* output_ppm) {
    if (!h) return RT_INVALID_ARG;
    try {
        if (!output_ppm || !output_ppm[0]) return h->rt->RenderToFrameBuffer();
        return h->rt->Render(output_ppm);
    } catch (...) { return RT_FAILURE; }
}
int rt580_raytracer_flush_ppm(rt580_raytracer* h, const char* output_ppm) {
    if (!h || !output_ppm) return RT_INVALID_ARG;
    try { return h->rt->FlushFrameBufferToPPM(output_ppm); } catch (...) { return RT_FAILURE; }
}
const int16_t* rt580_raytracer_framebuffer(rt580_raytracer* h) {
    return h ? reinterpret_cast<const int16_t*>(h->rt->FrameBuffer()) : nullptr;
}
int rt580_raytracer_stats(rt580_raytracer* h, rt580_stats* stats) {
    if (!h || !stats) return RT_INVALID_ARG;
    *stats = h->rt->Stats();
    return RT_SUCCESS;
}
int rt580_raytracer_flat_scene(rt580_raytracer* h, rt580_flat_scene* out) {
    if (!h || !out) return RT_INVALID_ARG;
    return h->rt->GetFlatScene(out);
}
int rt580_raytracer_render_params(rt580_raytracer* h, rt580_render_params* out) {
    if (!h || !out) return RT_INVALID_ARG;
    return h->rt->GetRenderParams(out);
}

}  // extern "C"
