"""A/B of the far-field direction grid against the O(n) filter scan it replaces (RT580_FAR_GRID=0): same frame,
bit for bit, on an open synthetic scene; prints the timings of both.
usage: python tools/gpu_far_ab.py [n_teapots] [W] [H] [spp] [K]"""
import os
import shutil
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402


def main():
    n_teapots = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 640
    H = int(sys.argv[3]) if len(sys.argv) > 3 else 360
    spp = int(sys.argv[4]) if len(sys.argv) > 4 else 4
    K = sys.argv[5] if len(sys.argv) > 5 else ""
    pkg = ge.load_package()
    import importlib.util
    spec = importlib.util.spec_from_file_location("scenegen", os.path.join(ge.PKG_DIR, "scenegen.py"))
    sg = importlib.util.module_from_spec(spec); spec.loader.exec_module(sg)
    d = "/tmp/rt580_far_ab"
    os.makedirs(d, exist_ok=True)
    shutil.copy(os.path.join(ge.ASSETS, "teapot.json"), d)
    sg.write_synthetic_scene(d, "ab", n_teapots=n_teapots, n_spheres=max(4, n_teapots // 2), seed=5)
    rt = pkg.Raytracer(W, H)
    rt.SetAssetsPath(d)
    rt.SetOptions(depth=4, ao_spp=spp)
    assert rt.LoadSceneJSON("ab.json") == pkg.RT_SUCCESS
    out = {}
    for tag, env in (("grid", K), ("scan", "0")):
        if env == "":
            os.environ.pop("RT580_FAR_GRID", None)
        else:
            os.environ["RT580_FAR_GRID"] = env
        ctx = pkg.Context(0)
        t0 = time.time()
        ctx.upload_scene(rt.flat_scene())
        t1 = time.time()
        fb, st = ctx.render(rt.render_params())
        t2 = time.time()
        fb, st = ctx.render(rt.render_params())
        t3 = time.time()
        out[tag] = (fb, st)
        print("%s: upload %.1f ms, frame %.1f ms (device %.2f: structure %.2f ao %.2f), rays %d, far_scans %d, linear %d, info %s" % (
            tag, (t1 - t0) * 1e3, (t3 - t2) * 1e3, st.ms_total, st.ms_structure, st.ms_ao, st.rays, st.far_scans, st.linear_fallbacks,
            ctx.scene_info().as_dict()), flush=True)
        ctx.close()
    a, b = out["grid"][0], out["scan"][0]
    nd = int((a != b).any(axis=-1).sum())
    print("pixels that differ: %d of %d; rays %d vs %d" % (nd, W * H, out["grid"][1].rays, out["scan"][1].rays))
    if nd:
        ys, xs = np.nonzero((a != b).any(axis=-1))
        print("first differing pixels:", list(zip(ys[:10].tolist(), xs[:10].tolist())))
        sys.exit(1)


if __name__ == "__main__":
    main()
