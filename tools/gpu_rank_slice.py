"""Debug helper (GPU box): what ONE rank of an N-GPU run does, timed on one GPU.

    python tools/gpu_rank_slice.py [--world 8] [--rank 0] [--frames 5] [--workload c4_room]

Renders the rows rank `rank` of `world` owns (interleaved, as bench.py partitions them) and
prints wall ms per frame next to the device-phase times, so that the fixed per-frame cost that
limits strong scaling (host round trips, small launches, slow-ray scans) can be measured without
an 8-GPU box.  world=1 is the whole frame.
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--rank", type=int, default=0)
    ap.add_argument("--frames", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--workload", default="c4_room")
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--spp", type=int, default=16)
    ap.add_argument("--contiguous", action="store_true", help="one band of H/world consecutive rows instead of interleaved rows")
    args = ap.parse_args()
    pkg = load_package()
    d = bench.scene_dir(args.workload)
    rt = pkg.Raytracer(args.width, args.height)
    rt.SetAssetsPath(d)
    rt.SetOptions(depth=4, ao_spp=args.spp)
    assert rt.LoadSceneJSON(args.workload + ".json") == 0
    ctx = pkg.Context(0)
    ctx.upload_scene(rt.flat_scene())
    p = rt.render_params().copy()
    p.row_first, p.row_step, p.n_rows = pkg.rows_for_rank(args.height, args.rank, args.world)
    if args.contiguous:
        n = args.height // args.world
        p.row_first, p.row_step, p.n_rows = args.rank * n, 1, n
    import torch
    band = torch.empty((max(p.n_rows, 1), args.width, 3), dtype=torch.int16, device="cuda")
    walls, sts = [], []
    for i in range(args.warmup + args.frames):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        counts = ctx.render_begin(p)
        t1 = time.perf_counter()
        bases = np.concatenate([[0], np.cumsum(counts[:-1])]).astype(np.uint64) * np.uint64(1)
        _, st = ctx.render_finish(p, bases, device_ptr=band.data_ptr())
        t2 = time.perf_counter()
        if i >= args.warmup:
            walls.append(((t2 - t0) * 1e3, (t1 - t0) * 1e3, (t2 - t1) * 1e3))
            sts.append(st)
    w = np.array(walls)
    st = sts[-1]
    dev = np.array([[s.ms_structure, s.ms_order, s.ms_ao, s.ms_resolve, s.ms_ao_kernel] for s in sts]).mean(0)
    print("world %d rank %d: %d rows, %d rays, %d launches" % (args.world, args.rank, p.n_rows, st.rays, st.kernel_launches))
    print("  wall ms/frame: total %.3f (min %.3f)  begin %.3f  finish %.3f" % (w[:, 0].mean(), w[:, 0].min(), w[:, 1].mean(), w[:, 2].mean()))
    print("  device ms:     structure %.3f  order %.3f  ao %.3f (kernels %.3f)  resolve %.3f  sum %.3f" % (
        dev[0], dev[1], dev[2], dev[4], dev[3], dev[:4].sum()))
    print("  Mrays/s of this slice: %.1f   x world = %.1f" % (st.rays / w[:, 0].mean() / 1e3, st.rays / w[:, 0].mean() / 1e3 * args.world))
    print("  far scans %d, linear fallbacks %d; through the tree: %d of %d shadow rays, %d of %d AO rays" % (
        st.far_scans, st.linear_fallbacks, st.shadow_rays_traversed, st.rays_shadow, st.ao_rays_traversed, st.rays_ao))


if __name__ == "__main__":
    main()
