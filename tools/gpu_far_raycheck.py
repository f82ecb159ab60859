"""Ray-level check of the deferred far-field machinery (direction grid, arc walk, inflated tree) against the oracle's
linear loop: the rays the oracle logs for sampled pixels of an open synthetic scene are re-traced on the GPU.
usage: python tools/gpu_far_raycheck.py [n_teapots] [W] [H] [n_pixels] [seed]"""
import os
import shutil
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402


def main():
    n_teapots = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 960
    H = int(sys.argv[3]) if len(sys.argv) > 3 else 540
    n_pix = int(sys.argv[4]) if len(sys.argv) > 4 else 300
    seed = int(sys.argv[5]) if len(sys.argv) > 5 else 1
    pkg = ge.load_package()
    oracle = ge.load_oracle()
    import importlib.util
    spec = importlib.util.spec_from_file_location("scenegen", os.path.join(ge.PKG_DIR, "scenegen.py"))
    sg = importlib.util.module_from_spec(spec); spec.loader.exec_module(sg)
    d = "/tmp/rt580_far_ab"
    os.makedirs(d, exist_ok=True)
    shutil.copy(os.path.join(ge.ASSETS, "teapot.json"), d)
    sg.write_synthetic_scene(d, "ab", n_teapots=n_teapots, n_spheres=max(4, n_teapots // 2), seed=5)
    rt = pkg.Raytracer(W, H)
    rt.SetAssetsPath(d)
    rt.SetOptions(depth=4, ao_spp=8)
    assert rt.LoadSceneJSON("ab.json") == pkg.RT_SUCCESS
    ctx = pkg.Context(0)
    ctx.upload_scene(rt.flat_scene())
    orc = oracle.Oracle(oracle.load_scene_json(d, "ab.json"))
    rng = np.random.default_rng(seed)
    pix = np.sort(rng.choice(W * H, n_pix, replace=False)).astype(np.int32)
    log = orc.render_log(W, H, 8, 4, pix=pix, ao_base=np.zeros(n_pix, np.uint64), max_rays=1 << 22)
    org, dirs, prim, t, kind = log["org"], log["dir"], log["prim"], log["t"], log["kind"]
    far = np.abs(org).max(axis=1) > 1e3
    print("rays %d (closest %d shadow %d ao %d), from outside the scene %d" % (len(t), (kind == 0).sum(), (kind == 1).sum(), (kind == 2).sum(), far.sum()))
    p, tt = ctx.trace_closest(org, dirs, pkg.TRAVERSAL_BVH)
    bad = (p.astype(np.int64) != prim) | ((prim >= 0) & (tt.view(np.uint32) != t.view(np.uint32)))
    print("closest: %d mismatches (%d among rays from outside)" % (bad.sum(), (bad & far).sum()))
    for i in np.flatnonzero(bad)[:12]:
        print("  ray %d kind %d |O| %.4g O %s d %s oracle (%d, %.9g) gpu (%d, %.9g)" % (i, kind[i], np.abs(org[i]).max(), org[i], dirs[i], prim[i], t[i], p[i], tt[i]))
    hit = ctx.trace_any(org, dirs, np.full(len(t), np.inf, np.float32), pkg.TRAVERSAL_BVH)
    bad2 = hit.astype(bool) != (prim >= 0)
    print("any (unbounded): %d mismatches (%d among rays from outside)" % (bad2.sum(), (bad2 & far).sum()))
    for i in np.flatnonzero(bad2)[:12]:
        print("  ray %d kind %d |O| %.4g O %s d %s oracle (%d, %.9g) gpu hit %d" % (i, kind[i], np.abs(org[i]).max(), org[i], dirs[i], prim[i], t[i], hit[i]))
    # bounded any-hit: tmax just above / below the oracle's t
    h = prim >= 0
    for scale in (1.0, 0.999):
        tm = np.where(h, t * np.float32(scale), np.float32(1e30)).astype(np.float32)
        hit = ctx.trace_any(org, dirs, tm, pkg.TRAVERSAL_BVH)
        expect = h & (t <= tm)
        bad3 = hit.astype(bool) != expect
        print("any (tmax = %.3f t): %d mismatches (%d among rays from outside)" % (scale, bad3.sum(), (bad3 & far).sum()))
    # synthetic rays, many: the GPU's own linear loop (BRUTE_FORCE) is the checker
    n = 300000
    info = ctx.scene_info()
    E = info.extent
    u = rng.normal(size=(n, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
    R = 10.0 ** rng.uniform(3.5, 8.3, n)
    O = (u * R[:, None]).astype(np.float32)
    dd = rng.normal(size=(n, 3)); dd /= np.linalg.norm(dd, axis=1, keepdims=True)
    cls = rng.integers(0, 5, n)
    # 1: radially outwards (+ small deviation), 2: aimed at a point of the scene, 3: sideways, 4: in-scene origin, 0: random
    dev = 10.0 ** rng.uniform(-6, -1, n)
    dd[cls == 1] = (u + dev[:, None] * dd)[cls == 1]
    target = rng.uniform(-0.6 * E, 0.6 * E, (n, 3)); target[:, 1] = rng.uniform(0, 8, n)
    dd[cls == 2] = (target - O)[cls == 2]
    side = np.cross(u, dd); dd[cls == 3] = side[cls == 3]
    O[cls == 4] = target[cls == 4].astype(np.float32)
    dd /= np.linalg.norm(dd, axis=1, keepdims=True)
    dd = dd.astype(np.float32)
    p1, t1 = ctx.trace_closest(O, dd, pkg.TRAVERSAL_BVH)
    p2, t2 = ctx.trace_closest(O, dd, pkg.TRAVERSAL_BRUTE_FORCE)
    badc = (p1 != p2) | ((p2 >= 0) & (t1.view(np.uint32) != t2.view(np.uint32)))
    print("synthetic closest: %d mismatches of %d; hits %d; by class %s" % (badc.sum(), n, (p2 >= 0).sum(), [int((badc & (cls == k)).sum()) for k in range(5)]))
    for i in np.flatnonzero(badc)[:10]:
        print("  ray %d class %d R %.4g O %s d %s linear (%d, %.9g) deferred (%d, %.9g)" % (i, cls[i], R[i], O[i], dd[i], p2[i], t2[i], p1[i], t1[i]))
    for i in np.flatnonzero(badc)[:3]:
        print("single ray", i, ctx.trace_closest(O[i:i + 1], dd[i:i + 1], pkg.TRAVERSAL_BVH), flush=True)
    tm = np.where(rng.random(n) < 0.5, np.float32(np.inf), (R * rng.uniform(0.5, 1.5, n)).astype(np.float32)).astype(np.float32)
    h1 = ctx.trace_any(O, dd, tm, pkg.TRAVERSAL_BVH)
    h2 = ctx.trace_any(O, dd, tm, pkg.TRAVERSAL_BRUTE_FORCE)
    bada = h1 != h2
    print("synthetic any: %d mismatches of %d; hits %d; by class %s" % (bada.sum(), n, h2.sum(), [int((bada & (cls == k)).sum()) for k in range(5)]))
    for i in np.flatnonzero(bada)[:10]:
        print("  ray %d class %d R %.4g O %s d %s tmax %.6g linear %d deferred %d (closest linear: %d, %.9g)" % (i, cls[i], R[i], O[i], dd[i], tm[i], h2[i], h1[i], p2[i], t2[i]))
    ctx.close()
    sys.exit(1 if (bad.sum() or bad2.sum() or badc.sum() or bada.sum()) else 0)


if __name__ == "__main__":
    main()
