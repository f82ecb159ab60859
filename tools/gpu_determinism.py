"""Debug helper (GPU box): render the benchmark scene N times (whole frame and a 1/8 row slice, two streams in
flight in the structure pass) and check that every frame is bit-identical to the first - a data race between the
streams would show up as a frame that differs."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
W, H = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1920, 1080)
pkg = load_package()
d = bench.scene_dir("c4_room")
rt = pkg.Raytracer(W, H)
rt.SetAssetsPath(d)
rt.SetOptions(depth=4, ao_spp=16)
assert rt.LoadSceneJSON("c4_room.json") == 0
ctx = pkg.Context(0)
ctx.upload_scene(rt.flat_scene())
for world in (1, 8):
    p = rt.render_params().copy()
    p.row_first, p.row_step, p.n_rows = pkg.rows_for_rank(H, 0, world)
    ref = None
    for i in range(n):
        fb, st = ctx.render(p)
        h = hashlib.md5(fb.tobytes()).hexdigest()
        if ref is None:
            ref = (h, st.rays)
        assert (h, st.rays) == ref, "frame %d of the 1/%d slice differs: %s %d vs %s %d" % (i, world, h, st.rays, ref[0], ref[1])
    print("1/%d of the rows at %dx%d: %d identical frames, md5 %s, %d rays" % (world, W, H, n, ref[0], ref[1]))
