"""Debug helper (GPU box): every ray the oracle traces for a frame, re-traced on the GPU (BVH and linear)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package, ASSETS
import oracle
pkg = load_package()
scene, W, H, spp, depth = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
orc = oracle.Oracle(oracle.load_scene_json(ASSETS, scene))
log = orc.render_log(W, H, spp, depth)
rt = pkg.Raytracer(W, H); rt.SetAssetsPath(ASSETS); assert rt.LoadSceneJSON(scene) == 0
fs = rt.flat_scene(); arr = pkg.flat_scene_arrays(fs)
ctx = pkg.Context(0); ctx.upload_scene(fs)
print(ctx.scene_info().as_dict(), "rays logged", len(log["t"]))
for trav, name in ((pkg.TRAVERSAL_BVH, "bvh"), (pkg.TRAVERSAL_BRUTE_FORCE, "lin")):
    p, t = ctx.trace_closest(log["org"], log["dir"], trav)
    hit = log["prim"] >= 0
    bad = np.nonzero((p.astype(np.int64) != log["prim"]) | (hit & (t.view(np.uint32) != log["t"].view(np.uint32))))[0]
    print(name, "mismatches vs oracle:", len(bad), "by kind", np.bincount(log["kind"][bad], minlength=3))
    for i in bad[:6]:
        print("  ray", i, "kind", log["kind"][i], "o", log["org"][i], "d", log["dir"][i], "oracle prim", log["prim"][i], "t", log["t"][i], "| gpu prim", p[i], "t", t[i])
        k = np.nonzero(arr["tri_prim"] == log["prim"][i])[0]
        if len(k):
            k = k[0]; v = np.stack([arr["tri_v0"][k, :3], arr["tri_v1"][k, :3], arr["tri_v2"][k, :3]]).astype(np.float64)
            o = log["org"][i].astype(np.float64); dd = log["dir"][i].astype(np.float64)
            P = o + dd * float(log["t"][i])
            e1 = v[1] - v[0]; e2 = v[2] - v[0]; N = np.cross(e1, e2); area2 = np.linalg.norm(N); N /= area2
            lo, hi = v.min(0), v.max(0)
            print("     P", P, "dist to unpadded box", np.maximum(np.maximum(lo - P, P - hi), 0), "N.d", N @ dd, "alt", area2 / np.linalg.norm(e2), area2 / np.linalg.norm(e1),
                  "diam", max(np.linalg.norm(e1), np.linalg.norm(e2), np.linalg.norm(v[2] - v[1])), "verts", v.tolist())
