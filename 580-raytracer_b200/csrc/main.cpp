// Command-line driver.  With no arguments it is the reference's main() (Raytracer.cpp:944-953):
// 500x500, simpleSphereScene.json from ./Assets/, output.ppm.
//   rt580_main [scene.json] [width] [height] [output.ppm] [assets_dir] [ao_spp] [depth]
#include "raytracer.h"
#include <cstdio>
#include <cstdlib>
#include <string>

int main(int argc, char** argv) {
    const std::string scene = argc > 1 ? argv[1] : "simpleSphereScene.json";
    const int w = argc > 2 ? atoi(argv[2]) : 500, h = argc > 3 ? atoi(argv[3]) : 500;
    const std::string out = argc > 4 ? argv[4] : "output.ppm";
    Raytracer rt(w, h);
    if (argc > 5) { std::string d = argv[5]; if (!d.empty() && d.back() != '/') d += '/'; rt.SetAssetsPath(d); }
    if (argc > 6) rt.SetAmbientOcclusionSamples(atoi(argv[6]));
    if (argc > 7) rt.SetBounces(atoi(argv[7]));
    int st = rt.LoadSceneJSON(scene);
    if (st != RT_SUCCESS) return st;
    st = rt.Render(out);
    const rt580_stats& s = rt.Stats();
    const double rays = (double)(s.rays_primary + s.rays_secondary + s.rays_shadow + s.rays_ao);
    fprintf(stderr, "rays %.0f  %.3f ms  %.1f Mrays/s\n", rays, s.ms_total, s.ms_total > 0 ? rays / s.ms_total / 1e3 : 0.0);
    return st;
}
