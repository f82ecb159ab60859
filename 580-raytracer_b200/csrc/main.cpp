// Command-line driver.  With no arguments it is the reference's main() (Raytracer.cpp:944-953):
// 500x500, simpleSphereScene.json from ./Assets/, output.ppm.
//   rt580_main [scene.json] [width] [height] [output.ppm] [assets_dir] [ao_spp] [depth]
//              [--gpus N] [--bench FRAMES] [--farfield exact|off] [--device D] [--mesh-cache DIR] [--no-ppm] [--device-flatten]
// --gpus N    rows interleaved over GPUs D .. D+N-1 of this process (Raytracer::SetGpus)
// --bench K   render K frames (after one warm-up frame) and print SURVEY 8d's table for this configuration: rays by kind,
//             ms per frame, Mrays/s (one ray = one IntersectScene call of the reference), and where the device time goes
#include "raytracer.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

int main(int argc, char** argv) {
    std::vector<std::string> pos;
    int gpus = 1, bench = 0, device = 0, farfield = RT580_FARFIELD_EXACT;
    bool ppm = true;
    std::string mesh_cache;
    bool device_flatten = false;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto next = [&]() -> const char* { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "--gpus") gpus = atoi(next());
        else if (a == "--bench") bench = atoi(next());
        else if (a == "--device") device = atoi(next());
        else if (a == "--farfield") farfield = strcmp(next(), "off") == 0 ? RT580_FARFIELD_OFF : RT580_FARFIELD_EXACT;
        else if (a == "--mesh-cache") mesh_cache = next();
        else if (a == "--no-ppm") ppm = false;
        else if (a == "--device-flatten") device_flatten = true;
        else if (a == "--help" || a == "-h") {
            fprintf(stderr, "usage: rt580_main [scene.json] [width] [height] [output.ppm] [assets_dir] [ao_spp] [depth] "
                            "[--gpus N] [--bench FRAMES] [--farfield exact|off] [--device D] [--mesh-cache DIR] [--no-ppm] [--device-flatten]\n");
            return 0;
        } else pos.push_back(a);
    }
    const std::string scene = pos.size() > 0 ? pos[0] : "simpleSphereScene.json";
    const int w = pos.size() > 1 ? atoi(pos[1].c_str()) : 500, h = pos.size() > 2 ? atoi(pos[2].c_str()) : 500;
    const std::string out = pos.size() > 3 ? pos[3] : "output.ppm";
    Raytracer rt(w, h);
    if (pos.size() > 4) { std::string d = pos[4]; if (!d.empty() && d.back() != '/') d += '/'; rt.SetAssetsPath(d); }
    if (pos.size() > 5) rt.SetAmbientOcclusionSamples(atoi(pos[5].c_str()));
    if (pos.size() > 6) rt.SetBounces(atoi(pos[6].c_str()));
    rt.SetGpus(gpus); rt.SetDevice(device); rt.SetFarField(farfield); rt.SetDeviceFlatten(device_flatten);
    if (!mesh_cache.empty()) rt.SetMeshCacheDir(mesh_cache);
    if (bench > 0) rt.SetQuiet(true);
    const auto t_load = std::chrono::steady_clock::now();
    int st = rt.LoadSceneJSON(scene);
    if (st != RT_SUCCESS) return st;
    const double load_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_load).count();
    auto rays_of = [](const rt580_stats& s) { return (double)(s.rays_primary + s.rays_secondary + s.rays_shadow + s.rays_ao); };
    if (bench > 0) {
        st = rt.RenderToFrameBuffer();                        // warm-up: upload, builds, buffer growth
        if (st != RT_SUCCESS) return st;
        printf("scene %s  %dx%d  ao_spp %s  depth %s  gpus %d  far field %s  (load + flatten %.2f s, mesh cache hits %d)\n", scene.c_str(), w, h,
               pos.size() > 5 ? pos[5].c_str() : "128", pos.size() > 6 ? pos[6].c_str() : "4", gpus, farfield == RT580_FARFIELD_OFF ? "off" : "exact", load_s,
               rt.MeshCacheHits());
        printf("%-6s %12s %12s %12s %14s %14s %10s %10s %12s\n", "frame", "primary", "secondary", "shadow", "ao", "rays", "ms", "wall ms", "Mrays/s");
        double sum_ms = 0, sum_rays = 0;
        for (int f = 0; f < bench; f++) {
            const auto t0 = std::chrono::steady_clock::now();
            st = rt.RenderToFrameBuffer();
            if (st != RT_SUCCESS) return st;
            const double wall = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            const rt580_stats& s = rt.Stats();
            const double rays = rays_of(s);
            printf("%-6d %12llu %12llu %12llu %14llu %14.0f %10.3f %10.3f %12.1f\n", f, (unsigned long long)s.rays_primary, (unsigned long long)s.rays_secondary,
                   (unsigned long long)s.rays_shadow, (unsigned long long)s.rays_ao, rays, s.ms_total, wall, s.ms_total > 0 ? rays / s.ms_total / 1e3 : 0.0);
            sum_ms += s.ms_total; sum_rays += rays;
        }
        printf("mean   %66.0f %10.3f %23.1f\n", sum_rays / bench, sum_ms / bench, sum_ms > 0 ? sum_rays / sum_ms / 1e3 : 0.0);
        rt580_profile pr;
        if (rt.Context() && rt580_frame_profile(rt.Context(), &pr) == RT580_SUCCESS) {
            static const char* names[RT580_N_CLASSES] = { "primary", "closest", "shadow_gen", "shadow_tree", "ao_gen", "ao_tree", "far_any", "far_closest", "order", "resolve" };
            printf("device time by class of the last frame (GPU %d):\n", device);
            for (int k = 0; k < RT580_N_CLASSES; k++)
                printf("  %-12s %10.3f ms %14llu rays %6u launches\n", names[k], pr.ms[k], (unsigned long long)pr.rays[k], pr.launches[k]);
        }
        if (ppm) st = rt.FlushFrameBufferToPPM(out);
        return st;
    }
    st = ppm ? rt.Render(out) : rt.RenderToFrameBuffer();
    const rt580_stats& s = rt.Stats();
    const double rays = rays_of(s);
    fprintf(stderr, "rays %.0f  %.3f ms  %.1f Mrays/s\n", rays, s.ms_total, s.ms_total > 0 ? rays / s.ms_total / 1e3 : 0.0);
    return st;
}
