"""Debug helper (GPU box): primary rays of a frame, BVH vs GPU linear loop; details of mismatches."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package, ASSETS
import oracle
pkg = load_package()
scene, W, H = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
orc = oracle.Oracle(oracle.load_scene_json(ASSETS, scene))
rt = pkg.Raytracer(W, H); rt.SetAssetsPath(ASSETS); assert rt.LoadSceneJSON(scene) == 0
fs = rt.flat_scene(); arr = pkg.flat_scene_arrays(fs)
ctx = pkg.Context(0); ctx.upload_scene(fs)
print(ctx.scene_info().as_dict())
org = np.zeros((W * H, 3), np.float32); d = np.zeros((W * H, 3), np.float32)
for i in range(W * H):
    _, o, dd = orc.primary_ray(W, H, i % W, i // W); org[i] = o; d[i] = dd
pb, tb = ctx.trace_closest(org, d, pkg.TRAVERSAL_BVH)
pl, tl = ctx.trace_closest(org, d, pkg.TRAVERSAL_BRUTE_FORCE)
bad = np.nonzero((pb != pl) | (tb.view(np.uint32) != tl.view(np.uint32)))[0]
print("primary rays", W * H, "hits", int((pl >= 0).sum()), "mismatches", len(bad))
for i in bad[:8]:
    p = pl[i]
    print("pixel", i % W, i // W, "o", org[i], "d", d[i], "lin prim", p, "t", tl[i], "| bvh prim", pb[i], "t", tb[i])
    k = np.nonzero(arr["tri_prim"] == p)[0]
    if len(k):
        k = k[0]; v = np.stack([arr["tri_v0"][k, :3], arr["tri_v1"][k, :3], arr["tri_v2"][k, :3]]).astype(np.float64)
        o = org[i].astype(np.float64); dd = d[i].astype(np.float64)
        P = o + dd * float(tl[i])
        e1 = v[1] - v[0]; e2 = v[2] - v[0]; N = np.cross(e1, e2); area2 = np.linalg.norm(N); N /= area2
        hb = area2 / np.linalg.norm(e2); hc = area2 / np.linalg.norm(e1)
        lo, hi = v.min(0), v.max(0)
        dist_box = np.maximum(np.maximum(lo - P, P - hi), 0)
        print("    verts", v.tolist())
        print("    P", P, "dist to unpadded box per axis", dist_box, "N.d", N @ dd, "altitudes hb hc", hb, hc, "diam", max(np.linalg.norm(e1), np.linalg.norm(e2), np.linalg.norm(v[2]-v[1])))
