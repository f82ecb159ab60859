"""Debug helper (GPU box): characterise BVH vs linear-loop mismatches on random rays."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package, ASSETS
pkg = load_package()
scene = sys.argv[1] if len(sys.argv) > 1 else "mix_small.json"
rt = pkg.Raytracer(8, 8); rt.SetAssetsPath(ASSETS); assert rt.LoadSceneJSON(scene) == 0
fs = rt.flat_scene(); arr = pkg.flat_scene_arrays(fs)
ctx = pkg.Context(0); ctx.upload_scene(fs)
rng = np.random.default_rng(580)
n = 400000
org = rng.uniform(-14, 14, (n, 3)).astype(np.float32); org[:, 1] = rng.uniform(-0.3, 9, n).astype(np.float32)
d = rng.normal(size=(n, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
pb, tb = ctx.trace_closest(org, d, pkg.TRAVERSAL_BVH)
pl, tl = ctx.trace_closest(org, d, pkg.TRAVERSAL_BRUTE_FORCE)
bad = np.nonzero((pb != pl) | (tb.view(np.uint32) != tl.view(np.uint32)))[0]
print("scene info", ctx.scene_info().as_dict()); print("rays", n, "hits(lin)", int((pl >= 0).sum()), "mismatches", len(bad))
nt = fs.n_tris
for i in bad[:15]:
    p = pl[i]
    kind = "tri" if p in set(arr["tri_prim"][:0]) else ""
    print("ray", i, "o", org[i], "d", d[i], "lin prim", p, "t", tl[i], "| bvh prim", pb[i], "t", tb[i])
    # locate primitive
    k = np.nonzero(arr["tri_prim"] == p)[0]
    if len(k):
        k = k[0]; v = np.stack([arr["tri_v0"][k, :3], arr["tri_v1"][k, :3], arr["tri_v2"][k, :3]]).astype(np.float64)
        lo, hi = v.min(0), v.max(0)
        o = org[i].astype(np.float64); dd = d[i].astype(np.float64)
        with np.errstate(divide="ignore", invalid="ignore"):
            t0 = (lo - o) / dd; t1 = (hi - o) / dd
        tn = np.nanmax(np.minimum(t0, t1)); tf = np.nanmin(np.maximum(t0, t1))
        print("    triangle", k, "verts", v.tolist(), "unpadded slab tn %.9g tf %.9g  (tn<=tf: %s) hit t %.9g" % (tn, tf, tn <= tf, tl[i]))
    else:
        k = np.nonzero(arr["sph_prim"] == p)[0]
        print("    sphere", arr["sph_center_r"][k[0]] if len(k) else None)
missed_prims = pl[bad]
print("missed prim histogram (top):", np.unique(missed_prims, return_counts=True))
