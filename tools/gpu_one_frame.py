"""Upload a benchmark scene and render frames of it (no torch): the process to put under ncu.
usage: python tools/gpu_one_frame.py <workload> <W> <H> [frames] [farfield exact|off]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import __graft_entry__ as ge  # noqa: E402


def main():
    name, W, H = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    frames = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    far = sys.argv[5] if len(sys.argv) > 5 else "exact"
    pkg = ge.load_package()
    d = bench.scene_dir(name)
    rt = pkg.Raytracer(W, H)
    rt.SetAssetsPath(d)
    rt.SetOptions(depth=4, ao_spp=16, farfield=pkg.FARFIELD_OFF if far == "off" else pkg.FARFIELD_EXACT)
    assert rt.LoadSceneJSON(name + ".json") == pkg.RT_SUCCESS
    ctx = pkg.Context(0)
    t0 = time.time()
    ctx.upload_scene(rt.flat_scene())
    print("upload %.1f ms" % ((time.time() - t0) * 1e3), ctx.scene_info().as_dict())
    p = rt.render_params()
    for _ in range(frames):
        t0 = time.time()
        fb, st = ctx.render(p)
        print("frame %.1f ms wall; device %.2f (structure %.2f order %.2f ao %.2f resolve %.2f); rays %d (ao %d traversed %d) shadow traversed %d far_scans %d linear %d launches %d" % (
            (time.time() - t0) * 1e3, st.ms_total, st.ms_structure, st.ms_order, st.ms_ao, st.ms_resolve, st.rays, st.rays_ao, st.ao_rays_traversed, st.shadow_rays_traversed,
            st.far_scans, st.linear_fallbacks, st.kernel_launches), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
