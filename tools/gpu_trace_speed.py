"""Debug helper (GPU box): raw traversal speed of the checker entry points on the bench scene."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
import bench
pkg = load_package()
name = sys.argv[1] if len(sys.argv) > 1 else "c4_room"
d = bench.scene_dir(name)
rt = pkg.Raytracer(64, 64); rt.SetAssetsPath(d); assert rt.LoadSceneJSON(name + ".json") == 0
ctx = pkg.Context(0); ctx.upload_scene(rt.flat_scene()); print(ctx.scene_info().as_dict())
rng = np.random.default_rng(0)
n = 2_000_000
def run(label, org, dirs, tmax=None):
    for rep in range(2):
        t0 = time.time()
        if tmax is None:
            p, t = ctx.trace_closest(org, dirs, pkg.TRAVERSAL_BVH); hits = (p >= 0).mean()
        else:
            h = ctx.trace_any(org, dirs, tmax, pkg.TRAVERSAL_BVH); hits = h.mean()
        dt = time.time() - t0
    print("%-40s %8.1f Mrays/s (incl. copies)  hit fraction %.3f" % (label, n / dt / 1e6, hits))
# floor points, upward hemisphere
org = np.stack([rng.uniform(-60, 60, n), np.full(n, -0.2), rng.uniform(-60, 60, n)], 1).astype(np.float32)
dirs = rng.normal(size=(n, 3)); dirs[:, 1] = np.abs(dirs[:, 1]); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True); dirs = dirs.astype(np.float32)
run("closest: floor -> up hemisphere", org, dirs)
run("any inf: floor -> up hemisphere", org, dirs, np.full(n, np.inf, np.float32))
run("any t<=20: floor -> up hemisphere", org, dirs, np.full(n, 20, np.float32))
# camera-like rays
cam = np.array([0, 33.75, 85.5], np.float32)
tgt = np.stack([rng.uniform(-70, 70, n), np.full(n, 0.0), rng.uniform(-70, 70, n)], 1)
dd = tgt - cam; dd /= np.linalg.norm(dd, axis=1, keepdims=True)
run("closest: camera -> floor points", np.tile(cam, (n, 1)), dd.astype(np.float32))
# mid-air random
org2 = np.stack([rng.uniform(-60, 60, n), rng.uniform(3, 30, n), rng.uniform(-60, 60, n)], 1).astype(np.float32)
d2 = rng.normal(size=(n, 3)); d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
run("closest: mid-air random", org2, d2.astype(np.float32))
m = 200000
for label, o, dd_, tm in [("closest floor->up", org[:m], dirs[:m], None), ("any inf floor->up", org[:m], dirs[:m], np.full(m, np.inf, np.float32)),
                          ("closest camera->floor", np.tile(cam, (m, 1)), dd[:m].astype(np.float32), None), ("closest mid-air", org2[:m], d2[:m].astype(np.float32), None)]:
    c = ctx.trace_profile(o, dd_, tm)
    print("%-24s node visits mean %.0f p99 %.0f max %d | leaf tests mean %.1f max %d | far scans %d linear %d" % (
        label, c[:, 0].mean(), np.percentile(c[:, 0], 99), c[:, 0].max(), c[:, 1].mean(), c[:, 1].max(), c[:, 2].sum(), c[:, 3].sum()))
