"""Debug helper (GPU box): locate the first pixel whose ray-tree structure differs from the oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package, ASSETS
import oracle
pkg = load_package()
scene, W, H, spp, depth = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
orc = oracle.Oracle(oracle.load_scene_json(ASSETS, scene))
ref, rays, hits = orc.render(W, H, spp, depth, nthreads=os.cpu_count())
for trav in (1, 2):
    rt = pkg.Raytracer(W, H); rt.SetAssetsPath(ASSETS); rt.SetOptions(depth=depth, ao_spp=spp, traversal=trav)
    assert rt.LoadSceneJSON(scene) == 0
    ctx = pkg.Context(0); ctx.upload_scene(rt.flat_scene())
    fb, st = ctx.render(rt.render_params())
    base = ctx.last_frame_ao_base(W * H)
    n_amb = 1
    gh = np.diff(np.concatenate([base, [st.hit_nodes * n_amb]])) // n_amb
    bad = np.nonzero(gh != hits)[0]
    print("trav", trav, "rays", st.rays, rays, "hit nodes", st.hit_nodes, hits.sum(), "pixels with different node count:", len(bad), bad[:10], [(int(b % W), int(b // W), int(gh[b]), int(hits[b])) for b in bad[:10]])
    print("   diff pixels", int((fb != ref).any(axis=-1).sum()))
