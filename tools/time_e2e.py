"""GPU box: where does the e2e step (upload + build + render with host buffers) spend its time?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
import bench
pkg = load_package()
d = bench.scene_dir("c4_room")
t = time.time(); rt = pkg.Raytracer(3840, 2160); rt.SetAssetsPath(d); rt.SetOptions(depth=4, ao_spp=16); assert rt.LoadSceneJSON("c4_room.json") == 0
print("LoadSceneJSON + flatten %.1f ms" % ((time.time() - t) * 1e3))
flat = rt.flat_scene(); p = rt.render_params()
ctx = pkg.Context(0)
for i in range(4):
    t0 = time.time(); ctx.upload_scene(flat); t1 = time.time(); fb, st = ctx.render(p); t2 = time.time()
    print("upload+build %.1f ms (device build %.1f ms)   render %.1f ms (device %.1f ms)" % ((t1 - t0) * 1e3, ctx.build_ms(), (t2 - t1) * 1e3, st.ms_total))
