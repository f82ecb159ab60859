"""GPU parity tests proper: the CUDA path through the C ABI (librt580.so) against
  - the committed reference-generated goldens (tests/golden/*.npz; T0 = the reference itself),
  - the T1 oracle on the same inputs at sizes it finishes in seconds,
  - size-independent properties (BVH == linear loop, row partition == whole frame).
Bar: bit-exact int16 frame buffer and identical IntersectScene call counts; the PPM tolerance
of north_star (<=1 LSB on >=99.9 % of pixels, max <=2 LSB) is asserted on top of that."""
import ctypes
import os

import numpy as np
import pytest

from conftest import ASSETS, load_golden

pytestmark = pytest.mark.gpu
NT = os.cpu_count() or 1

GOLDEN_TAGS = ["c1_500_spp128", "c1_500_spp16", "c1_200_spp64", "c1_200_spp8_d0", "c1_200_spp8_d2", "ao_500_spp128",
               "tri_500_spp128", "teapots_160x90_spp16", "teapots_96_spp1", "teapots_point_96x64_spp4",
               "mix_small_128x72_spp4", "wide_37x23_spp3"]


def make_rt(pkg, scene, W, H, spp, depth, traversal=None, rng=None):
    rt = pkg.Raytracer(W, H)
    rt.SetAssetsPath(ASSETS)
    rt.SetOptions(depth=depth, ao_spp=spp, traversal=pkg.TRAVERSAL_AUTO if traversal is None else traversal,
                  rng_mode=pkg.RNG_REFERENCE_LCG if rng is None else rng)
    assert rt.LoadSceneJSON(scene) == pkg.RT_SUCCESS
    return rt


def render(pkg, scene, W, H, spp, depth, traversal=None, rng=None):
    rt = make_rt(pkg, scene, W, H, spp, depth, traversal, rng)
    st = rt.Render("")
    assert st == pkg.RT_SUCCESS, pkg.lib().rt580_last_error().decode()
    return rt.frame_buffer(), rt.stats()


def assert_ppm_tolerance(oracle, fb, ref):
    a = oracle.gamma_encode(fb).astype(np.int32)
    b = oracle.gamma_encode(ref).astype(np.int32)
    err = np.abs(a - b).max(axis=-1)
    assert err.max() <= 2, "max PPM channel error %d LSB" % err.max()
    assert (err <= 1).mean() >= 0.999


@pytest.mark.parametrize("tag", GOLDEN_TAGS)
@pytest.mark.parametrize("traversal", ["auto", "bvh"])
def test_frame_matches_reference_golden(pkg, oracle, tag, traversal):
    g = load_golden(tag)
    trav = {"auto": pkg.TRAVERSAL_AUTO, "bvh": pkg.TRAVERSAL_BVH}[traversal]
    fb, st = render(pkg, g["scene"], g["W"], g["H"], g["spp"], g["depth"], traversal=trav)
    assert st.rays == g["rays"], "IntersectScene calls: %d vs reference %d" % (st.rays, g["rays"])
    ndiff = int((fb != g["fb"]).any(axis=-1).sum())
    assert ndiff == 0, "%d pixels differ from the reference frame buffer" % ndiff
    assert_ppm_tolerance(oracle, fb, g["fb"])


def test_render_writes_the_reference_ppm(pkg, tmp_path):
    """Raytracer::Render(outputName) end to end: the PPM file has the reference's md5."""
    import hashlib
    g = load_golden("c1_500_spp16")
    rt = make_rt(pkg, g["scene"], g["W"], g["H"], g["spp"], g["depth"])
    out = str(tmp_path / "output.ppm")
    assert rt.Render(out) == pkg.RT_SUCCESS
    with open(out, "rb") as f:
        assert hashlib.md5(f.read()).hexdigest() == g["ppm_md5"]


@pytest.mark.parametrize("scene,W,H,spp,depth", [
    ("scene.json", 320, 180, 4, 4),
    ("scene_point.json", 256, 144, 2, 3),
    ("mix_small.json", 256, 144, 8, 4),
    ("simpleSphereSceneAO.json", 333, 111, 7, 1),
    ("simpleScene.json", 64, 64, 1, 4),
])
def test_frame_matches_oracle(pkg, oracle, oracle_scene, scene, W, H, spp, depth):
    ref, rays, hits = oracle_scene(scene).render(W, H, spp, depth, nthreads=NT)
    fb, st = render(pkg, scene, W, H, spp, depth)
    assert st.rays == rays
    assert st.hit_nodes == int(hits.sum())
    assert np.array_equal(fb, ref), "%d pixels differ" % int((fb != ref).any(axis=-1).sum())


def test_bvh_equals_linear_loop_on_random_rays(pkg, oracle, oracle_scene):
    """Closest hit by LBVH == by the GPU linear loop == by the oracle's linear loop:
    same primitive, same t bits."""
    scene = "mix_small.json"
    rt = make_rt(pkg, scene, 8, 8, 1, 0)
    ctx = pkg.Context(0)
    ctx.upload_scene(rt.flat_scene())
    rng = np.random.default_rng(580)
    n = 200000
    org = rng.uniform(-14, 14, (n, 3)).astype(np.float32)
    org[:, 1] = rng.uniform(-0.3, 9, n).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    d[:50] = 0.0                                 # zero-direction rays (TIR, Q20)
    d[50:100, 0] = 0.0                           # axis-parallel components
    d[100:150] = np.array([0, -1, 0], np.float32)
    p_bvh, t_bvh = ctx.trace_closest(org, d, pkg.TRAVERSAL_BVH)
    p_lin, t_lin = ctx.trace_closest(org, d, pkg.TRAVERSAL_BRUTE_FORCE)
    assert np.array_equal(p_bvh, p_lin)
    assert np.array_equal(t_bvh.view(np.uint32), t_lin.view(np.uint32))
    m = 20000                                    # oracle is O(N) per ray
    p_orc, t_orc = oracle_scene(scene).intersect(org[:m], d[:m], nthreads=NT)
    assert np.array_equal(p_bvh[:m].astype(np.int64), p_orc)
    hit = p_orc >= 0
    assert np.array_equal(t_bvh[:m][hit].view(np.uint32), t_orc[hit].view(np.uint32))
    assert hit.sum() > 1000
    # any-hit with a distance bound == "closest t <= tmax"
    tmax = rng.uniform(0.5, 30, n).astype(np.float32)
    any_bvh = ctx.trace_any(org, d, tmax, pkg.TRAVERSAL_BVH)
    any_lin = ctx.trace_any(org, d, tmax, pkg.TRAVERSAL_BRUTE_FORCE)
    expect = ((p_lin >= 0) & (t_lin <= tmax)).astype(np.uint8)
    assert np.array_equal(any_bvh, expect)
    assert np.array_equal(any_lin, expect)
    ctx.close()


def test_ao_stream_matches_oracle(pkg, oracle):
    ctx = pkg.Context(0)
    for normal, step in [((0, 1, 0), 0), ((0.6, -0.48, 0.64), 2 * 128 * 12345), ((-1, 0, 0), 4_000_000_123)]:
        a = ctx.hemisphere_stream(normal, step, 512)
        b = oracle.hemisphere_stream(normal, step, 512)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    ctx.close()


def test_device_powf_matches_libm(pkg):
    """cpp:253 truncates right after powf, so the device must return libm's float."""
    ctx = pkg.Context(0)
    rng = np.random.default_rng(1)
    n = 2_000_000
    x = np.concatenate([rng.random(n // 2, dtype=np.float32),
                        (1.0 - rng.random(n // 2, dtype=np.float32) * np.float32(1e-3)).astype(np.float32)])
    x[:8] = [0.0, 1.0, 1e-38, 1e-45, 0.5, 0.99999994, 3e-39, 0.25]
    y = rng.choice(np.array([2, 5, 10, 32, 700, 900, 1, 0.5, 0, 1 / 2.2], np.float32), n)
    libm = ctypes.CDLL("libm.so.6")
    libm.powf.restype = ctypes.c_float
    libm.powf.argtypes = [ctypes.c_float, ctypes.c_float]
    got = ctx.powf(x, y)
    idx = rng.choice(n, 200000, replace=False)
    idx[:8] = np.arange(8)
    want = np.array([libm.powf(float(x[i]), float(y[i])) for i in idx], np.float32)
    assert np.array_equal(got[idx].view(np.uint32), want.view(np.uint32))
    ctx.close()


def test_row_partition_equals_whole_frame(pkg):
    """Two contexts on one GPU play two ranks: interleaved rows, row counts exchanged on the
    host, gathered frame == single-context frame bit for bit (SURVEY 8e gate)."""
    scene, W, H, spp, depth = "mix_small.json", 160, 90, 4, 4
    whole, st_whole = render(pkg, scene, W, H, spp, depth)
    rt = make_rt(pkg, scene, W, H, spp, depth)
    base_params = rt.render_params()
    for world in (2, 3):
        ctxs, params, counts = [], [], []
        for r in range(world):
            c = pkg.Context(0)
            c.upload_scene(rt.flat_scene())
            p = base_params.copy()
            p.row_first, p.row_step, p.n_rows = pkg.rows_for_rank(H, r, world)
            ctxs.append(c); params.append(p)
            counts.append(c.render_begin(p))
        bases = pkg.row_bases_from_counts(H, world, counts)
        bands, rays = [], 0
        for r in range(world):
            fb, st = ctxs[r].render_finish(params[r], bases[r])
            bands.append(fb); rays += st.rays
        got = pkg.interleave_rows(H, W, world, bands)
        assert rays == st_whole.rays
        assert np.array_equal(got, whole)
        for c in ctxs:
            c.close()


def test_counter_rng_mode_is_partition_invariant(pkg):
    """RT580_RNG_COUNTER needs no exchange: any row split gives the same image; only the AO
    noise differs from the reference stream."""
    scene, W, H, spp, depth = "simpleSphereScene.json", 120, 80, 8, 4
    whole, _ = render(pkg, scene, W, H, spp, depth, rng=pkg.RNG_COUNTER)
    ref, _ = render(pkg, scene, W, H, spp, depth)
    rt = make_rt(pkg, scene, W, H, spp, depth, rng=pkg.RNG_COUNTER)
    p = rt.render_params()
    ctx = pkg.Context(0)
    ctx.upload_scene(rt.flat_scene())
    bands = []
    for r in range(2):
        q = p.copy()
        q.row_first, q.row_step, q.n_rows = pkg.rows_for_rank(H, r, 2)
        ctx.render_begin(q)
        fb, _ = ctx.render_finish(q, None)
        bands.append(fb)
    assert np.array_equal(pkg.interleave_rows(H, W, 2, bands), whole)
    # same geometry, different AO noise: background / unshaded pixels agree, mean close
    assert np.array_equal(whole[0], ref[0])
    assert abs(float(whole.mean()) - float(ref.mean())) < 1.0
    ctx.close()


def test_sampled_pixels_of_a_larger_scene(pkg, oracle, oracle_scene, tmp_path):
    """The C4/C5 checking mode at a size the oracle still reaches: a generated scene with ~16k
    triangles at 640x360; the oracle renders a few hundred sampled pixels, seeded with the AO
    ordinals the GPU structure pass reports."""
    import importlib.util
    import __graft_entry__ as ge
    spec = importlib.util.spec_from_file_location("scenegen", os.path.join(ge.PKG_DIR, "scenegen.py"))
    sg = importlib.util.module_from_spec(spec); spec.loader.exec_module(sg)
    d = str(tmp_path)
    import shutil
    shutil.copy(os.path.join(ASSETS, "teapot.json"), d)
    sg.write_synthetic_scene(d, "mid", n_teapots=16, n_spheres=24, seed=7)
    W, H, spp, depth = 640, 360, 4, 4
    rt = pkg.Raytracer(W, H)
    rt.SetAssetsPath(d)
    rt.SetOptions(depth=depth, ao_spp=spp)
    assert rt.LoadSceneJSON("mid.json") == pkg.RT_SUCCESS
    ctx = pkg.Context(0)
    ctx.upload_scene(rt.flat_scene())
    fb, st = ctx.render(rt.render_params())
    base = ctx.last_frame_ao_base(W * H)
    rng = np.random.default_rng(3)
    pix = np.sort(rng.choice(W * H, 400, replace=False)).astype(np.int32)
    orc = oracle.Oracle(oracle.load_scene_json(d, "mid.json"))
    ref, _, _ = orc.render(W, H, spp, depth, pix=pix, ao_base=base[pix], nthreads=NT)
    got = fb.reshape(-1, 3)[pix]
    assert np.array_equal(got, ref), "%d sampled pixels differ" % int((got != ref).any(axis=-1).sum())
    assert (got != np.array([254, 64, 205], np.int16)).any(axis=-1).sum() > 50   # the sample is not all background
    ctx.close()


def test_error_behaviour(pkg, tmp_path):
    rt = pkg.Raytracer(32, 32)
    rt.SetAssetsPath(ASSETS)
    assert rt.LoadSceneJSON("does_not_exist.json") == pkg.RT_FAILURE          # cpp:650-653
    assert rt.Render("") == pkg.RT_FAILURE
    (tmp_path / "broken.json").write_text("{ \"scene\": ")
    rt.SetAssetsPath(str(tmp_path))
    assert rt.LoadSceneJSON("broken.json") == pkg.RT_FAILURE                  # cpp:657-663
    ctx = pkg.Context(0)
    with pytest.raises(pkg.Rt580Error):
        ctx.render(pkg.RenderParams())                                         # no scene uploaded
    with pytest.raises(pkg.Rt580Error):
        pkg.Context(9999)
    ctx.close()


def test_empty_scene_and_single_primitive(pkg, oracle, tmp_path):
    import json
    cam = {"from": [0, 0, 5], "to": [0, 0, 0], "bounds": [0.1, 10, 1, -1, 1, -1], "resolution": [8, 8]}
    (tmp_path / "empty.json").write_text(json.dumps({"scene": {"shapes": [], "lights": [], "camera": cam}}))
    rt = pkg.Raytracer(16, 8)
    rt.SetAssetsPath(str(tmp_path))
    assert rt.LoadSceneJSON("empty.json") == pkg.RT_SUCCESS
    assert rt.Render("") == pkg.RT_SUCCESS
    assert (rt.frame_buffer().reshape(-1, 3) == np.array([254, 64, 205], np.int16)).all()
    # one sphere, forced through the one-node BVH
    import shutil
    shutil.copy(os.path.join(ASSETS, "1sphere.json"), str(tmp_path))
    one = {"scene": {"shapes": [{"id": "s", "geometry": "1sphere",
                                 "material": {"Cs": [1, 0.5, 0.2], "Ka": 0.3, "Kd": 0.8, "Ks": 0.5, "Kt": 0.4, "n": 20},
                                 "transforms": [{"T": [0, 0, 0]}]}],
                     "lights": [{"type": "ambient", "color": [1, 1, 1], "intensity": 0.3},
                                {"type": "point", "color": [1, 1, 1], "intensity": 1.0, "position": [3, 3, 3]}],
                     "camera": cam}}
    (tmp_path / "one.json").write_text(json.dumps(one))
    orc = oracle.Oracle(oracle.load_scene_json(str(tmp_path), "one.json"))
    ref, rays, _ = orc.render(64, 48, 4, 4, nthreads=NT)
    for trav in (pkg.TRAVERSAL_BVH, pkg.TRAVERSAL_BRUTE_FORCE):
        rt = pkg.Raytracer(64, 48)
        rt.SetAssetsPath(str(tmp_path))
        rt.SetOptions(depth=4, ao_spp=4, traversal=trav)
        assert rt.LoadSceneJSON("one.json") == pkg.RT_SUCCESS
        assert rt.Render("") == pkg.RT_SUCCESS
        assert rt.stats().rays == rays
        assert np.array_equal(rt.frame_buffer(), ref)


def _scenegen():
    import importlib.util
    import __graft_entry__ as ge
    spec = importlib.util.spec_from_file_location("scenegen", os.path.join(ge.PKG_DIR, "scenegen.py"))
    sg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sg)
    return sg


def test_closed_room_scene(pkg, oracle, tmp_path):
    """The benchmark's scene family at a size the oracle renders whole: teapots + spheres in the
    closed double-walled room.  Exercises the large-primitive list (the 24 wall triangles stay out
    of the tree), the persistent any-hit kernel and a frame without a single escaping ray."""
    import shutil
    d = str(tmp_path)
    shutil.copy(os.path.join(ASSETS, "teapot.json"), d)
    _scenegen().write_synthetic_scene(d, "room", n_teapots=3, n_spheres=6, seed=11, room=True)
    W, H, spp, depth = 200, 112, 4, 4
    orc = oracle.Oracle(oracle.load_scene_json(d, "room.json"))
    ref, rays, hits = orc.render(W, H, spp, depth, nthreads=NT)
    for trav in (pkg.TRAVERSAL_AUTO, pkg.TRAVERSAL_BRUTE_FORCE):
        rt = pkg.Raytracer(W, H)
        rt.SetAssetsPath(d)
        rt.SetOptions(depth=depth, ao_spp=spp, traversal=trav)
        assert rt.LoadSceneJSON("room.json") == pkg.RT_SUCCESS
        assert rt.Render("") == pkg.RT_SUCCESS
        st = rt.stats()
        assert st.rays == rays and st.hit_nodes == int(hits.sum())
        assert np.array_equal(rt.frame_buffer(), ref)
    assert (hits > 0).all()          # closed room: every primary ray hits something


def test_every_ray_of_a_frame_matches_the_oracle(pkg, oracle, oracle_scene):
    """Ray-level parity: every IntersectScene call the oracle makes for a frame (primary, secondary,
    shadow, AO - including the rays that start at far-field hits 10^6 units away) is re-traced
    through the LBVH path: same primitive, same t bits."""
    scene, W, H, spp, depth = "scene.json", 192, 108, 2, 4
    log = oracle_scene(scene).render_log(W, H, spp, depth)
    rt = make_rt(pkg, scene, W, H, spp, depth)
    ctx = pkg.Context(0)
    ctx.upload_scene(rt.flat_scene())
    p, t = ctx.trace_closest(log["org"], log["dir"], pkg.TRAVERSAL_BVH)
    assert np.array_equal(p.astype(np.int64), log["prim"])
    hit = log["prim"] >= 0
    assert np.array_equal(t[hit].view(np.uint32), log["t"][hit].view(np.uint32))
    assert len(log["t"]) > 20000 and set(np.unique(log["kind"])) == {0, 1, 2}
    far = np.abs(log["org"]).max(axis=1) > 1e4
    assert far.sum() > 0, "this frame is known to contain rays that start at far-field hits"
    ctx.close()


def test_device_side_exchange_equals_host_exchange(pkg):
    """The multi-GPU path of bench.py on one GPU: contexts play the ranks, the per-row counts stay on
    the device (rt580_row_counts_to_device -> [world][max_rows], what the NCCL all-gather produces),
    rt580_render_finish_interleaved derives every rank's stream offsets there, and each context
    stores its rows into a whole frame (rt580_frame_export; the other ranks of a real run map that
    allocation over CUDA IPC).  Result == the single-context frame, bit for bit."""
    import torch
    scene, W, H, spp, depth = "mix_small.json", 160, 90, 4, 4
    whole, st_whole = render(pkg, scene, W, H, spp, depth)
    rt = make_rt(pkg, scene, W, H, spp, depth)
    base_params = rt.render_params()
    for world in (1, 2, 4):
        max_rows = (H + world - 1) // world
        all_d = torch.zeros((world, max_rows), dtype=torch.int64, device="cuda")
        ctxs, params = [], []
        for r in range(world):
            c = pkg.Context(0)
            c.upload_scene(rt.flat_scene())
            c.frame_export(W, H)                      # every context its own whole frame in this single-process test
            p = base_params.copy()
            p.row_first, p.row_step, p.n_rows = pkg.rows_for_rank(H, r, world)
            assert c.render_begin(p, want_counts=False) is None
            c.row_counts_to_device(all_d[r].data_ptr(), max_rows)
            ctxs.append(c); params.append(p)
        torch.cuda.synchronize()
        got = np.zeros((H, W, 3), np.int16)
        rays = 0
        for r in range(world):
            st = ctxs[r].render_finish_interleaved(all_d.data_ptr(), world, r, max_rows)
            rays += st.rays
            full = ctxs[r].frame_read(W, H)
            first, step, n = pkg.rows_for_rank(H, r, world)
            got[first:first + n * step:step] = full[first:first + n * step:step]
        assert rays == st_whole.rays
        assert np.array_equal(got, whole)
        # a finish that does not match the partition is refused
        ctxs[0].render_begin(params[0], want_counts=False)
        with pytest.raises(pkg.Rt580Error):
            ctxs[0].render_finish_interleaved(all_d.data_ptr(), world + 1, 0, max_rows)
        for c in ctxs:
            c.close()


@pytest.mark.parametrize("tag", ["teapots_96_spp1", "mix_small_128x72_spp4", "teapots_point_96x64_spp4"])
def test_deferred_queue_overflow_repeats_the_pass(pkg, tag, monkeypatch):
    """A frame that "hardly leaks" (few slow rays so far) gets a small deferred any-hit queue; if it
    overflows after all, the pass is repeated with a queue that takes every ray.  RT580_SLOW_ANY_CAP=1
    forces that on open scenes at a low resolution: still the reference's frame, bit for bit."""
    g = load_golden(tag)
    monkeypatch.setenv("RT580_SLOW_ANY_CAP", "1")
    fb, st = render(pkg, g["scene"], g["W"], g["H"], g["spp"], g["depth"], traversal=pkg.TRAVERSAL_BVH)
    assert st.far_scans + st.linear_fallbacks > 1, "the frame has no slow rays: nothing overflowed"
    assert st.rays == g["rays"]
    assert np.array_equal(fb, g["fb"])


def test_page_locked_host_arrays(pkg):
    """rt580_host_alloc memory as the destination of rt580_render (what the host class uses)."""
    scene, W, H, spp, depth = "simpleSphereScene.json", 96, 64, 4, 4
    rt = make_rt(pkg, scene, W, H, spp, depth)
    ctx = pkg.Context(0)
    ctx.upload_scene(rt.flat_scene())
    ha = pkg.HostArray((H, W, 3), np.int16)
    fb, _ = ctx.render(rt.render_params(), out=ha.array)
    ref, _ = ctx.render(rt.render_params())
    assert fb is ha.array and np.array_equal(fb, ref)
    ha.close(); ctx.close()


def test_ppm_body_from_the_device(pkg, oracle):
    """rt580_frame_rgb8 (FlushFrameBufferToPPM's gamma + truncation on the device): the bytes
    equal the host restatement of cpp:809-823 applied to the int16 frame, for a whole frame and for a
    band of rows whose element count is not a multiple of four."""
    scene, W, H, spp, depth = "simpleSphereScene.json", 123, 77, 4, 4
    rt = make_rt(pkg, scene, W, H, spp, depth)
    ctx = pkg.Context(0)
    ctx.upload_scene(rt.flat_scene())
    lut = oracle.gamma_encode(np.arange(256, dtype=np.int16))
    p = rt.render_params()
    fb, _ = ctx.render(p)
    assert np.array_equal(ctx.frame_rgb8(lut, H, W), oracle.gamma_encode(fb))
    q = p.copy()
    q.row_first, q.row_step, q.n_rows = 1, 3, 25
    ctx.render_begin(q)
    band, _ = ctx.render_finish(q, np.zeros(25, np.uint64))
    assert np.array_equal(ctx.frame_rgb8(lut, 25, W), oracle.gamma_encode(band))
    with pytest.raises(pkg.Rt580Error):
        ctx.render_begin(q)
        ctx.frame_rgb8(lut, 25, W)          # no finished frame
    ctx.close()


def test_benchmark_scene_at_full_size(pkg, oracle):
    """BASELINE config 4 as bench.py runs it - c4_room: 1,000,448 triangles + 1000 spheres, 3840x2160,
    depth 4, 16 AO samples, reference stream - checked at full size:
      * 48 sampled pixels against the T1 oracle (the reference's linear loop over a million primitives),
        seeded with the AO ordinals of the GPU's structure pass;
      * ray accounting: rays = primary + secondary + nodes * (shadow lights + spp * ambient lights);
      * the frame is reproducible (second render bit-identical) and equals the frame assembled from
        three interleaved row sets with the stream offsets derived on the device."""
    import bench
    import torch
    name, W, H, spp, depth = "c4_room", 3840, 2160, 16, 4
    d = bench.scene_dir(name)
    rt = pkg.Raytracer(W, H)
    rt.SetAssetsPath(d)
    rt.SetOptions(depth=depth, ao_spp=spp)
    assert rt.LoadSceneJSON(name + ".json") == pkg.RT_SUCCESS
    ctx = pkg.Context(0)
    ctx.upload_scene(rt.flat_scene())
    p = rt.render_params()
    fb, st = ctx.render(p)
    base = ctx.last_frame_ao_base(W * H)
    assert st.rays_primary == W * H
    assert st.rays_shadow == st.hit_nodes * 3 and st.rays_ao == st.hit_nodes * spp
    assert st.rays == st.rays_primary + st.rays_secondary + st.rays_shadow + st.rays_ao
    assert st.shadow_rays_traversed < st.rays_shadow      # the clearance maps of the three point lights are in use
    rng = np.random.default_rng(580)
    pix = np.sort(rng.choice(W * H, 48, replace=False)).astype(np.int32)
    orc = oracle.Oracle(oracle.load_scene_json(d, name + ".json"))
    ref, _, _ = orc.render(W, H, spp, depth, pix=pix, ao_base=base[pix], nthreads=NT)
    got = fb.reshape(-1, 3)[pix]
    assert np.array_equal(got, ref), "%d of 48 sampled pixels differ from the oracle" % int((got != ref).any(axis=-1).sum())
    fb2, st2 = ctx.render(p)
    assert st2.rays == st.rays and np.array_equal(fb, fb2)
    # three "ranks" on this one GPU, device-side exchange (the same context, one row set after the other)
    world = 3
    max_rows = (H + world - 1) // world
    all_d = torch.zeros((world, max_rows), dtype=torch.int64, device="cuda")
    ctxs = [ctx] + [pkg.Context(0) for _ in range(world - 1)]
    for c in ctxs[1:]:
        c.upload_scene(rt.flat_scene())
    qs = []
    for r in range(world):
        q = p.copy()
        q.row_first, q.row_step, q.n_rows = pkg.rows_for_rank(H, r, world)
        ctxs[r].render_begin(q, want_counts=False)
        ctxs[r].row_counts_to_device(all_d[r].data_ptr(), max_rows)
        qs.append(q)
    torch.cuda.synchronize()
    got3 = np.empty_like(fb)
    for r in range(world):
        band = torch.empty((qs[r].n_rows, W, 3), dtype=torch.int16, device="cuda")
        ctxs[r].render_finish_interleaved(all_d.data_ptr(), world, r, max_rows, device_ptr=band.data_ptr())
        first, step, n = pkg.rows_for_rank(H, r, world)
        got3[first:first + n * step:step] = band.cpu().numpy()
    assert np.array_equal(got3, fb)
    for c in ctxs:
        c.close()


@pytest.mark.parametrize("scene,spp", [("scene.json", 16), ("scene.json", 64), ("scene_point.json", 16)])
def test_mesh_scene_at_1080p_sampled(pkg, oracle, oracle_scene, scene, spp):
    """BASELINE configs 2 and 3 at their named size: the 4-teapot mesh scene (and its point-light variant) at
    1920x1080, depth 4, 16 / 64 AO samples.  An open scene: rays escape, so the far-field replay, the linear
    fallback and the "leaky" form of the deferred queues all run.  300 sampled pixels (half of them drawn from
    the pixels that hit something) against the T1 oracle, seeded with the GPU's AO ordinals."""
    W, H, depth = 1920, 1080, 4
    rt = make_rt(pkg, scene, W, H, spp, depth)
    ctx = pkg.Context(0)
    ctx.upload_scene(rt.flat_scene())
    fb, st = ctx.render(rt.render_params())
    base = ctx.last_frame_ao_base(W * H)
    flat = fb.reshape(-1, 3)
    covered = np.flatnonzero((flat != np.array([254, 64, 205], np.int16)).any(axis=-1))
    assert covered.size > 1000
    rng = np.random.default_rng(2)
    pix = np.unique(np.concatenate([rng.choice(W * H, 150, replace=False), rng.choice(covered, 150, replace=False)])).astype(np.int32)
    ref, _, _ = oracle_scene(scene).render(W, H, spp, depth, pix=pix, ao_base=base[pix], nthreads=NT)
    got = flat[pix]
    assert np.array_equal(got, ref), "%d sampled pixels differ" % int((got != ref).any(axis=-1).sum())
    assert st.far_scans + st.linear_fallbacks > 0
    ctx.close()


def test_nan_rays_hit_the_first_triangle(pkg, oracle, oracle_scene):
    """A ray with a NaN in it: every compare of cpp:371 / 382 / 396 is false, so the reference's loop keeps the FIRST triangle
    of the scene at t = NaN (cpp:487-499) and spheres reject it (h:558-560).  Same answer from the LBVH path, the GPU's linear
    loop and the shared-memory path of a tiny scene."""
    rng = np.random.default_rng(4)
    n = 64
    org = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    for i in range(n):                                    # a NaN in one component of the origin or of the direction; a few plain rays
        if i % 8 == 7:
            continue
        (org if i % 2 else d)[i, i % 3] = np.nan
    for scene in ("mix_small.json", "simpleSphereScene.json", "scene.json"):
        rt = make_rt(pkg, scene, 8, 8, 1, 0)
        ctx = pkg.Context(0)
        ctx.upload_scene(rt.flat_scene())
        p_orc, t_orc = oracle_scene(scene).intersect(org, d, nthreads=1)
        nan = np.isnan(org).any(axis=1) | np.isnan(d).any(axis=1)
        assert (p_orc[nan] == p_orc[nan][0]).all() and p_orc[nan][0] >= 0 and np.isnan(t_orc[nan]).all()
        for trav in (pkg.TRAVERSAL_BVH, pkg.TRAVERSAL_BRUTE_FORCE, pkg.TRAVERSAL_AUTO):
            p, t = ctx.trace_closest(org, d, trav)
            assert np.array_equal(p.astype(np.int64), p_orc)
            assert np.isnan(t[nan]).all()
            hit = p_orc >= 0
            assert np.array_equal(t[hit & ~nan].view(np.uint32), t_orc[hit & ~nan].view(np.uint32))
            anyh = ctx.trace_any(org, d, np.full(n, 5.0, np.float32), trav)
            assert anyh[nan].all()                        # cpp:75: "distance > distToLight" is false for NaN too -> occluded
        ctx.close()


def _synthetic_scene(pkg, tmp_dir, name, **kw):
    import shutil
    shutil.copy(os.path.join(ASSETS, "teapot.json"), tmp_dir)
    _scenegen().write_synthetic_scene(tmp_dir, name, **kw)
    rt = pkg.Raytracer(64, 36)
    rt.SetAssetsPath(tmp_dir)
    rt.SetOptions(depth=4, ao_spp=4)
    assert rt.LoadSceneJSON(name + ".json") == pkg.RT_SUCCESS
    return rt


def _rays_from_outside(rng, n, extent):
    """Origins 10^3.5 .. 10^8.3 units out (the children of far-field hits start there), directions of five kinds:
    random, radially outwards (what a reflection off a grazed triangle gives), aimed at the scene (shadow rays to a point
    light), sideways, and in-scene origins that escape."""
    u = rng.normal(size=(n, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
    R = 10.0 ** rng.uniform(3.5, 8.3, n)
    O = (u * R[:, None]).astype(np.float32)
    dd = rng.normal(size=(n, 3)); dd /= np.linalg.norm(dd, axis=1, keepdims=True)
    cls = rng.integers(0, 5, n)
    dev = 10.0 ** rng.uniform(-6, -1, n)
    dd[cls == 1] = (u + dev[:, None] * dd)[cls == 1]
    target = rng.uniform(-0.6 * extent, 0.6 * extent, (n, 3)); target[:, 1] = rng.uniform(0, 8, n)
    dd[cls == 2] = (target - O)[cls == 2]
    dd[cls == 3] = np.cross(u, dd)[cls == 3]
    O[cls == 4] = target[cls == 4].astype(np.float32)
    dd /= np.linalg.norm(dd, axis=1, keepdims=True)
    return O, dd.astype(np.float32), cls, R


@pytest.mark.parametrize("n_teapots,grid", [(40, ""), (200, ""), (40, "32"), (40, "overflow")])
def test_far_field_machinery_equals_the_linear_loop(pkg, tmp_path, monkeypatch, n_teapots, grid):
    """The rays the tree cannot answer alone - the ones that leave the scene (far-field direction grid, fargrid.cuh) and the ones
    that start 10^4..10^8 units outside it (arc walk over the grid, tree with inflated boxes) - against the GPU's own linear
    loop over every primitive (the reference's loop, cpp:476-521): same primitive, same t bits, same any-hit answer, on
    300,000 synthetic rays of five kinds over an open scene of 41K / 205K triangles."""
    if grid == "overflow":
        monkeypatch.setenv("RT580_ARC_MAX_CELLS", "8")       # long arcs are given up early: the filtered scan of every record (k_far_linear)
    elif grid:
        monkeypatch.setenv("RT580_FAR_GRID", grid)           # a coarse grid: long lists, many rays per cell
    rt = _synthetic_scene(pkg, str(tmp_path), "far", n_teapots=n_teapots, n_spheres=max(4, n_teapots // 2), seed=5)
    ctx = pkg.Context(0)
    ctx.upload_scene(rt.flat_scene())
    rng = np.random.default_rng(17)
    n = 300000
    O, dd, cls, R = _rays_from_outside(rng, n, ctx.scene_info().extent)
    p1, t1 = ctx.trace_closest(O, dd, pkg.TRAVERSAL_BVH)
    p2, t2 = ctx.trace_closest(O, dd, pkg.TRAVERSAL_BRUTE_FORCE)
    assert np.array_equal(p1, p2)
    assert np.array_equal(t1[p2 >= 0].view(np.uint32), t2[p2 >= 0].view(np.uint32))
    for k in range(5):
        assert ((p2 >= 0) & (cls == k)).sum() > 1000, "class %d hardly ever hits: the test would be vacuous" % k
    assert ((p2 >= 0) & (t2 > 1e4)).sum() > 50000            # far-field "hits" 10^4 .. 10^9 units away
    tm = np.where(rng.random(n) < 0.5, np.float32(np.inf), (R * rng.uniform(0.5, 1.5, n)).astype(np.float32)).astype(np.float32)
    h1 = ctx.trace_any(O, dd, tm, pkg.TRAVERSAL_BVH)
    h2 = ctx.trace_any(O, dd, tm, pkg.TRAVERSAL_BRUTE_FORCE)
    assert np.array_equal(h1, h2)
    assert np.array_equal(h2.astype(bool), (p2 >= 0) & (t2 <= tm))
    ctx.close()


def test_far_field_grid_frame_equals_the_scan_it_replaces(pkg, tmp_path, monkeypatch):
    """A whole frame of an open scene with the far-field direction grid and with the O(n) filter scan of round 1
    (RT580_FAR_GRID=0): the same int16 frame, the same ray count."""
    rt = _synthetic_scene(pkg, str(tmp_path), "ab", n_teapots=24, n_spheres=12, seed=5)
    out = {}
    for tag, env in (("grid", None), ("scan", "0")):
        if env is None:
            monkeypatch.delenv("RT580_FAR_GRID", raising=False)
        else:
            monkeypatch.setenv("RT580_FAR_GRID", env)
        ctx = pkg.Context(0)
        ctx.upload_scene(rt.flat_scene())
        p = rt.render_params()
        p.width, p.height, p.ao_spp = 480, 270, 4
        out[tag] = ctx.render(p)
        ctx.close()
    assert out["grid"][1].rays == out["scan"][1].rays
    assert out["grid"][1].far_scans > 10000 and out["grid"][1].linear_fallbacks > 100
    assert np.array_equal(out["grid"][0], out["scan"][0])


def _sampled_pixels_against_the_oracle(pkg, oracle, d, name, W, H, spp, depth, n_pix, seed):
    rt = pkg.Raytracer(W, H)
    rt.SetAssetsPath(d)
    rt.SetOptions(depth=depth, ao_spp=spp)
    assert rt.LoadSceneJSON(name + ".json") == pkg.RT_SUCCESS
    ctx = pkg.Context(0)
    ctx.upload_scene(rt.flat_scene())
    fb, st = ctx.render(rt.render_params())
    base = ctx.last_frame_ao_base(W * H)
    flat = fb.reshape(-1, 3)
    covered = np.flatnonzero((flat != np.array([254, 64, 205], np.int16)).any(axis=-1))
    rng = np.random.default_rng(seed)
    pix = np.unique(np.concatenate([rng.choice(W * H, n_pix // 2, replace=False), rng.choice(covered, n_pix // 2, replace=False)])).astype(np.int32)
    orc = oracle.Oracle(oracle.load_scene_json(d, name + ".json"))
    ref, rays_ref, hits_ref = orc.render(W, H, spp, depth, pix=pix, ao_base=base[pix], nthreads=NT)
    got = flat[pix]
    return ctx, rt, fb, st, pix, got, ref, hits_ref


def test_open_benchmark_scene_at_full_size(pkg, oracle):
    """BASELINE config 4 as SURVEY 8d specifies it and as bench.py runs it - c4_open: 977 teapot instances (1,000,448 triangles)
    + 1000 spheres over an open floor, ambient + directional + point light, 3840x2160, depth 4, 16 AO samples, reference
    stream, far field exact.  Rays escape: the AO rays walk the LBVH, a fifth of those that leave the scene are "hit" by a
    triangle 10^5..10^8 units away and their children start out there.
      * 2,000 sampled pixels (half of them drawn from the pixels that hit something) against the T1 oracle - the reference's
        linear loop over a million primitives per ray - seeded with the AO ordinals of the GPU's structure pass;
      * the per-pixel hit-node counts of those pixels (the structure of the ray trees) equal the oracle's;
      * ray accounting, and the AO pass is not vacuous: most AO rays go through the tree, some pixels are partly occluded."""
    import bench
    name, W, H, spp, depth = "c4_open", 3840, 2160, 16, 4
    d = bench.scene_dir(name)
    ctx, rt, fb, st, pix, got, ref, hits_ref = _sampled_pixels_against_the_oracle(pkg, oracle, d, name, W, H, spp, depth, 2000, 580)
    assert np.array_equal(got, ref), "%d of %d sampled pixels differ from the oracle" % (int((got != ref).any(axis=-1).sum()), len(pix))
    base = ctx.last_frame_ao_base(W * H + 1) if False else None
    assert st.rays_primary == W * H
    assert st.rays_shadow == st.hit_nodes * 2 and st.rays_ao == st.hit_nodes * spp
    assert st.rays == st.rays_primary + st.rays_secondary + st.rays_shadow + st.rays_ao
    assert st.ao_rays_traversed > 0.5 * st.rays_ao                 # the AO rays really walk the tree
    assert st.far_scans > 10_000_000 and st.linear_fallbacks > 10_000_000
    assert (hits_ref > 0).sum() > 800 and (hits_ref == 0).sum() > 200
    # node counts of the sampled pixels from the GPU's AO ordinals: base[p + 1] - base[p] for pixels of one row
    fb2, st2 = ctx.render(rt.render_params())
    assert st2.rays == st.rays and np.array_equal(fb, fb2)
    ctx.close()


def test_bvh_equals_brute_force_at_a_million_primitives(pkg):
    """SURVEY 8c check (i) at C4 scale: closest hit through the LBVH (+ large-primitive list + far-field machinery) against the
    GPU's brute-force loop over all 1,001,450 primitives of c4_open - same primitive, same t bits - on 100,000 rays: camera
    rays, rays between random points of the scene, rays that escape, rays from outside."""
    import bench
    name = "c4_open"
    d = bench.scene_dir(name)
    rt = pkg.Raytracer(3840, 2160)
    rt.SetAssetsPath(d)
    assert rt.LoadSceneJSON(name + ".json") == pkg.RT_SUCCESS
    ctx = pkg.Context(0)
    ctx.upload_scene(rt.flat_scene())
    E = ctx.scene_info().extent
    rng = np.random.default_rng(9)
    n = 100000
    a = rng.uniform(-0.65 * E, 0.65 * E, (n, 3)); a[:, 1] = rng.uniform(-0.3, 9, n)
    b = rng.uniform(-0.65 * E, 0.65 * E, (n, 3)); b[:, 1] = rng.uniform(-0.3, 9, n)
    O = a.astype(np.float32)
    dd = b - a
    up = rng.random(n) < 0.3                                  # a third of them leave the scene upwards
    dd[up] = rng.normal(size=(int(up.sum()), 3)); dd[up, 1] = np.abs(dd[up, 1])
    cam = np.array(rt.render_params().camera_from[:], np.float32)
    O[: n // 5] = cam
    dd[: n // 5] = b[: n // 5] - cam
    Of, df, _, _ = _rays_from_outside(rng, n // 5, E)
    O[-(n // 5):] = Of; dd[-(n // 5):] = df
    dd /= np.linalg.norm(dd, axis=1, keepdims=True)
    dd = dd.astype(np.float32)
    p1, t1 = ctx.trace_closest(O, dd, pkg.TRAVERSAL_BVH)
    p2, t2 = ctx.trace_closest(O, dd, pkg.TRAVERSAL_BRUTE_FORCE)
    assert np.array_equal(p1, p2), "%d rays hit another primitive" % int((p1 != p2).sum())
    assert np.array_equal(t1[p2 >= 0].view(np.uint32), t2[p2 >= 0].view(np.uint32))
    assert (p2 >= 0).sum() > 50000 and ((p2 >= 0) & (t2 > 1e4)).sum() > 1000 and (p2 < 0).sum() > 5000
    ctx.close()


def test_ten_million_triangles_sampled(pkg, oracle):
    """BASELINE config 5's scene - c5_open: 9766 teapot instances = 10,000,384 triangles over an open floor - far field exact, at a
    reduced resolution (the exact frame at 7680x4320 is dominated by rays that start 10^5..10^8 units outside the scene:
    nine of ten rays that leave a scene of 10^7 triangles are "hit" by float noise, cpp:392): 96 sampled pixels against the
    T1 oracle's linear loop over all ten million triangles, and the same frame from two interleaved row sets."""
    import bench
    name, W, H, spp, depth = "c5_open", 480, 270, 16, 4
    d = bench.scene_dir(name)
    ctx, rt, fb, st, pix, got, ref, hits_ref = _sampled_pixels_against_the_oracle(pkg, oracle, d, name, W, H, spp, depth, 96, 5)
    assert np.array_equal(got, ref), "%d of %d sampled pixels differ from the oracle" % (int((got != ref).any(axis=-1).sum()), len(pix))
    assert st.linear_fallbacks > st.rays // 3              # the rays from outside the scene are the bulk of this frame
    p = rt.render_params()
    counts, ps = [], []
    for r in range(2):
        q = p.copy()
        q.row_first, q.row_step, q.n_rows = pkg.rows_for_rank(H, r, 2)
        ps.append(q)
    bands = []
    c0 = ctx.render_begin(ps[0])
    ctx2 = pkg.Context(0)
    ctx2.upload_scene(rt.flat_scene())
    c1 = ctx2.render_begin(ps[1])
    bases = pkg.row_bases_from_counts(H, 2, [c0, c1])
    bands.append(ctx.render_finish(ps[0], bases[0])[0])
    bands.append(ctx2.render_finish(ps[1], bases[1])[0])
    assert np.array_equal(pkg.interleave_rows(H, W, 2, bands), fb)
    ctx.close(); ctx2.close()


def test_bvh_equals_brute_force_at_ten_million_primitives(pkg):
    """SURVEY 8c check (i) at C5 scale: LBVH (+ far-field machinery) against the GPU's brute-force loop over all 10,000,386
    primitives of c5_open on 30,000 rays: same primitive, same t bits."""
    import bench
    name = "c5_open"
    d = bench.scene_dir(name)
    rt = pkg.Raytracer(7680, 4320)
    rt.SetAssetsPath(d)
    assert rt.LoadSceneJSON(name + ".json") == pkg.RT_SUCCESS
    ctx = pkg.Context(0)
    ctx.upload_scene(rt.flat_scene())
    E = ctx.scene_info().extent
    rng = np.random.default_rng(10)
    n = 30000
    a = rng.uniform(-0.65 * E, 0.65 * E, (n, 3)); a[:, 1] = rng.uniform(-0.3, 9, n)
    b = rng.uniform(-0.65 * E, 0.65 * E, (n, 3)); b[:, 1] = rng.uniform(-0.3, 9, n)
    O = a.astype(np.float32)
    dd = b - a
    up = rng.random(n) < 0.3
    dd[up] = rng.normal(size=(int(up.sum()), 3)); dd[up, 1] = np.abs(dd[up, 1])
    Of, df, _, _ = _rays_from_outside(rng, n // 4, E)
    O[-(n // 4):] = Of; dd[-(n // 4):] = df
    dd /= np.linalg.norm(dd, axis=1, keepdims=True)
    dd = dd.astype(np.float32)
    p1, t1 = ctx.trace_closest(O, dd, pkg.TRAVERSAL_BVH)
    p2, t2 = ctx.trace_closest(O, dd, pkg.TRAVERSAL_BRUTE_FORCE)
    assert np.array_equal(p1, p2), "%d rays hit another primitive" % int((p1 != p2).sum())
    assert np.array_equal(t1[p2 >= 0].view(np.uint32), t2[p2 >= 0].view(np.uint32))
    assert (p2 >= 0).sum() > 15000 and ((p2 >= 0) & (t2 > 1e4)).sum() > 1000
    ctx.close()


def test_deep_tree_of_clustered_centroids(pkg, oracle, tmp_path):
    """An LBVH over centroids clustered at every scale is a chain: triangles at 2^20, 2^19, ... 2^-2 along x, then along y, then
    along z (each halving peels one leaf off: ~66 levels), and 1500 coincident triangles at the far end (equal Morton keys:
    ordered by index, another ~11 levels).  Round 1 refused such scenes (depth > 64); the traversal stack now holds any tree
    the build can produce.  Same frame as the oracle, bit for bit."""
    import json
    import shutil
    d = str(tmp_path)
    shutil.copy(os.path.join(ASSETS, "1triangle.json"), d)
    shapes = []
    mat = {"Cs": [0.8, 0.5, 0.3], "Ka": 0.3, "Kd": 0.7, "Ks": 0.4, "Kt": 0.0, "n": 10}

    def tri(pos, k):
        shapes.append({"id": "t%d" % len(shapes), "geometry": "1triangle", "material": dict(mat, Cs=[0.3 + 0.1 * (k % 7), 0.5, 0.9 - 0.1 * (k % 5)]),
                       "transforms": [{"S": [1, 1, 1]}, {"T": pos}]})
    for axis in range(3):
        for k in range(23):
            pos = [0.0, 0.0, 0.0]
            pos[axis] = float(2.0 ** (20 - k))
            tri(pos, k)
    for k in range(1500):
        tri([0.0, 0.0, 0.0], k)
    scene = {"scene": {"shapes": shapes,
                       "lights": [{"type": "ambient", "color": [1, 1, 1], "intensity": 0.3},
                                  {"type": "directional", "color": [1, 1, 1], "intensity": 0.9, "from": [2, 5, 9], "to": [0, 0, 0]}],
                       "camera": {"from": [3, 2, 12], "to": [0, 0, 0], "bounds": [0.1, 10, 1, -1, 1, -1], "resolution": [8, 8]}}}
    with open(os.path.join(d, "deep.json"), "w") as f:
        json.dump(scene, f)
    W, H, spp, depth = 96, 64, 2, 2
    rt = pkg.Raytracer(W, H)
    rt.SetAssetsPath(d)
    rt.SetOptions(depth=depth, ao_spp=spp, traversal=pkg.TRAVERSAL_BVH)
    assert rt.LoadSceneJSON("deep.json") == pkg.RT_SUCCESS
    assert rt.Render("") == pkg.RT_SUCCESS, pkg.lib().rt580_last_error().decode()
    st = rt.stats()
    assert st.bvh_max_depth > 64, "the scene did not produce a deep tree (depth %d)" % st.bvh_max_depth
    orc = oracle.Oracle(oracle.load_scene_json(d, "deep.json"))
    ref, rays, _ = orc.render(W, H, spp, depth, nthreads=NT)
    assert st.rays == rays
    assert np.array_equal(rt.frame_buffer(), ref)


def test_more_ranks_than_rows(pkg):
    """rows_for_rank hands a rank beyond the frame's height n_rows = -1 ("no row"; 0 would mean the whole frame, rt580.h): it
    takes part in the exchange with empty hands, and the assembled frame is still the single-context frame."""
    import torch
    scene, W, H, spp, depth = "simpleSphereScene.json", 64, 3, 4, 2
    whole, st_whole = render(pkg, scene, W, H, spp, depth)
    rt = make_rt(pkg, scene, W, H, spp, depth)
    world = 5
    assert pkg.rows_for_rank(H, 4, world)[2] == -1
    max_rows = (H + world - 1) // world
    all_d = torch.zeros((world, max_rows), dtype=torch.int64, device="cuda")
    ctxs, got, rays = [], np.zeros((H, W, 3), np.int16), 0
    for r in range(world):
        c = pkg.Context(0)
        c.upload_scene(rt.flat_scene())
        c.frame_export(W, H)
        p = rt.render_params().copy()
        p.row_first, p.row_step, p.n_rows = pkg.rows_for_rank(H, r, world)
        c.render_begin(p, want_counts=False)
        c.row_counts_to_device(all_d[r].data_ptr(), max_rows)
        ctxs.append(c)
    torch.cuda.synchronize()
    for r in range(world):
        st = ctxs[r].render_finish_interleaved(all_d.data_ptr(), world, r, max_rows)
        rays += st.rays
        first, step, n = pkg.rows_for_rank(H, r, world)
        if n > 0:
            got[first:first + n * step:step] = ctxs[r].frame_read(W, H)[first:first + n * step:step]
    assert rays == st_whole.rays and np.array_equal(got, whole)
    for c in ctxs:
        c.close()


def test_ao_directions_match_libm_on_two_million_samples(pkg, oracle):
    """cpp:277-278 computes (float)(r * cos a), (float)(r * sin a) through the C library's DOUBLE cos / sin; neither glibc's nor
    CUDA's is correctly rounded, so equality of the float results is a property to measure, not one that holds by construction:
    2,097,152 consecutive samples of the reference stream (and the same number from far down the stream), device vs host."""
    ctx = pkg.Context(0)
    for normal, step in [((0.0, 1.0, 0.0), 0), ((0.3, -0.2, 0.93), 3_000_000_001)]:
        n = 1 << 21
        a = ctx.hemisphere_stream(normal, step, n)
        b = oracle.hemisphere_stream(normal, step, n)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), "%d directions differ" % int((a.view(np.uint32) != b.view(np.uint32)).any(axis=1).sum())
    ctx.close()


def test_ppm_body_into_a_misaligned_device_buffer(pkg, oracle):
    """rt580_frame_rgb8 with out_on_device and a destination that is not 4-byte aligned (advisor finding, round 1)."""
    import torch
    scene, W, H, spp, depth = "simpleSphereScene.json", 61, 37, 2, 1
    rt = make_rt(pkg, scene, W, H, spp, depth)
    ctx = pkg.Context(0)
    ctx.upload_scene(rt.flat_scene())
    fb, _ = ctx.render(rt.render_params())
    lut = np.ascontiguousarray(oracle.gamma_encode(np.arange(256, dtype=np.int16)), np.uint8)
    buf = torch.zeros(W * H * 3 + 8, dtype=torch.uint8, device="cuda")
    for off in (0, 1, 3):
        buf.zero_()
        st = pkg.lib().rt580_frame_rgb8(ctx._h, lut.ctypes.data, buf.data_ptr() + off, 1)
        assert st == pkg.RT_SUCCESS, pkg.lib().rt580_last_error().decode()
        torch.cuda.synchronize()
        got = buf[off:off + W * H * 3].cpu().numpy().reshape(H, W, 3)
        assert np.array_equal(got, oracle.gamma_encode(fb))
    ctx.close()


def _main_binary():
    import conftest
    p = os.path.join(os.path.dirname(os.path.abspath(conftest.__file__)), "..", "580-raytracer_b200", "rt580_main")
    p = os.path.abspath(p)
    if not os.path.exists(p):
        pytest.fail("580-raytracer_b200/rt580_main is not built (python -c 'import __graft_entry__ as g; g.build()')")
    return p


@pytest.mark.gpu
def test_main_binary_writes_the_reference_image(tmp_path):
    """SURVEY 8f-4 / Raytracer.cpp:944-953: the command-line driver with the reference's main() configuration
    (simpleSphereScene.json, 500x500, spp 128, depth 4) writes the PPM the unmodified reference writes (Appendix B md5)."""
    import hashlib
    import subprocess
    out = str(tmp_path / "output.ppm")
    r = subprocess.run([_main_binary(), "simpleSphereScene.json", "500", "500", out, ASSETS], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    with open(out, "rb") as f:
        assert hashlib.md5(f.read()).hexdigest() == "a00a8b5cb0a0e7dd3bd85f675a43e94a"
    # --bench: the same frame after K timed frames, and the 8d table on stdout
    out2 = str(tmp_path / "bench.ppm")
    r = subprocess.run([_main_binary(), "simpleSphereScene.json", "500", "500", out2, ASSETS, "--bench", "3", "--mesh-cache", str(tmp_path)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Mrays/s" in r.stdout and "38812073" in r.stdout.replace(",", ""), r.stdout
    with open(out2, "rb") as f:
        assert hashlib.md5(f.read()).hexdigest() == "a00a8b5cb0a0e7dd3bd85f675a43e94a"


@pytest.mark.gpu
def test_main_binary_on_two_gpus_writes_the_same_image(tmp_path):
    """--gpus 2: rows interleaved over two contexts of one process, the same bytes."""
    import hashlib
    import subprocess
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU on this box")
    out = str(tmp_path / "output.ppm")
    r = subprocess.run([_main_binary(), "simpleSphereScene.json", "500", "500", out, ASSETS, "--gpus", "2"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    with open(out, "rb") as f:
        assert hashlib.md5(f.read()).hexdigest() == "a00a8b5cb0a0e7dd3bd85f675a43e94a"


@pytest.mark.gpu
def test_class_on_several_contexts_of_one_gpu_equals_one_context(pkg):
    """Raytracer::SetGpus drives N contexts from one process; with one GPU on the box the contexts cannot be on different
    devices, so this covers the exchange through the explicit C ABI instead: 3 contexts on device 0, rows interleaved."""
    scene, W, H, spp, depth = "simpleScene.json", 75, 41, 3, 3
    whole, st = render(pkg, scene, W, H, spp, depth)
    rt = make_rt(pkg, scene, W, H, spp, depth)
    fs, rp = rt.flat_scene(), rt.render_params()
    world = 3
    ctxs = [pkg.Context(0) for _ in range(world)]
    for c in ctxs:
        c.upload_scene(fs)
    import copy
    counts, params = [], []
    for r, c in enumerate(ctxs):
        p = copy.copy(rp)
        p.row_first, p.row_step, p.n_rows = pkg.rows_for_rank(H, r, world)
        params.append(p)
        counts.append(c.render_begin(p))
    bases = pkg.row_bases_from_counts(H, world, counts)
    frame = np.zeros((H, W, 3), np.int16)
    for r, c in enumerate(ctxs):
        band, _ = c.render_finish(params[r], bases[r])
        frame[r::world] = band
    assert (frame == whole).all()


@pytest.mark.gpu
@pytest.mark.parametrize("scene", ["simpleScene.json", "scene.json", "mix_small.json", "simpleSphereScene.json"])
def test_device_flatten_equals_host_flatten(pkg, scene):
    """FlattenScene's TransformPoint calls (cpp:353-355, h:234-248) on the device.  The arrays the device produces
    from meshes + model matrices are the host's byte for byte (rotations, scales and translations in the transforms), and a frame
    rendered from them is the same frame."""
    W, H, spp, depth = 64, 40, 2, 3
    rt = make_rt(pkg, scene, W, H, spp, depth)
    host = pkg.flat_scene_arrays(rt.flat_scene())
    inst = rt.instanced_scene()
    ctx = pkg.Context(0)
    dev = ctx.flatten_instanced(inst, int(rt.flat_scene().n_tris), int(rt.flat_scene().n_spheres))
    for k, v in dev.items():
        assert v.tobytes() == np.ascontiguousarray(host[k]).tobytes(), k
    ctx.upload_instanced_scene(inst)
    fb_dev, st_dev = ctx.render(rt.render_params())
    ctx.upload_scene(rt.flat_scene())
    fb_host, st_host = ctx.render(rt.render_params())
    assert (fb_dev == fb_host).all() and st_dev.rays == st_host.rays
    ctx.close()
    # and through the class
    rt2 = make_rt(pkg, scene, W, H, spp, depth)
    rt2.SetDeviceFlatten(True)
    assert rt2.Render("") == pkg.RT_SUCCESS
    assert (rt2.frame_buffer() == fb_host).all()
