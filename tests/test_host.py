"""Host side of the drop-in class (no GPU needed): LoadSceneJSON / LoadMesh / FlattenScene /
InitializeRenderer through the C ABI class mirror, against the oracle's restatement and, where
the reference library is present, against the reference's own loader and ComputeModelMatrix."""
import json
import os

import numpy as np
import pytest

from conftest import ASSETS

SCENES = ["simpleSphereScene.json", "simpleSphereSceneAO.json", "simpleScene.json", "scene.json", "scene_point.json",
          "mix_small.json"]


def transform_point(M, p):
    """Matrix::TransformPoint (h:234-248) in float32 with the reference's operation order."""
    f = np.float32
    out = []
    for r in range(3):
        out.append(f(f(f(M[r, 0] * p[0]) + f(M[r, 1] * p[1])) + f(M[r, 2] * p[2])) + M[r, 3])
    return np.array(out, np.float32)


@pytest.mark.parametrize("scene", SCENES)
def test_flatten_matches_oracle_model(pkg, oracle, scene):
    rt = pkg.Raytracer(8, 8)
    rt.SetAssetsPath(ASSETS)
    assert rt.LoadSceneJSON(scene) == pkg.RT_SUCCESS
    fs = rt.flat_scene()
    arr = pkg.flat_scene_arrays(fs)
    sa = oracle.load_scene_json(ASSETS, scene)
    assert fs.n_prims == sa.n_prims
    assert fs.n_materials == len(sa.shape_mesh) and fs.n_lights == len(sa.light_type)
    assert np.array_equal(arr["materials"], sa.shape_material)
    assert np.array_equal(arr["light_type"], sa.light_type)
    assert np.array_equal(arr["light_f"].view(np.uint32), sa.light_f.view(np.uint32))
    prim, ti, si = 0, 0, 0
    rng = np.random.default_rng(0)
    for s, m in enumerate(sa.shape_mesh):
        M = oracle.model_matrix(sa.shape_srt[s])
        if sa.mesh_type[m] == 0:
            b, e = sa.mesh_tri_begin[m], sa.mesh_tri_begin[m + 1]
            n = e - b
            assert np.array_equal(arr["tri_prim"][ti:ti + n], np.arange(prim, prim + n))
            assert (arr["tri_material"][ti:ti + n] == s).all()
            assert np.array_equal(arr["tri_n0"][ti:ti + n, :3], sa.tri_nrm[b:e, 0:3])      # object space (Q10)
            assert np.array_equal(arr["tri_n2"][ti:ti + n, :3], sa.tri_nrm[b:e, 6:9])
            for k in rng.choice(n, min(n, 40), replace=False):
                for v, key in enumerate(["tri_v0", "tri_v1", "tri_v2"]):
                    want = transform_point(M, sa.tri_pos[b + k, 3 * v:3 * v + 3])
                    assert np.array_equal(arr[key][ti + k, :3].view(np.uint32), want.view(np.uint32))
            ti += n
            prim += n
        else:
            assert arr["sph_prim"][si] == prim and arr["sph_material"][si] == s
            assert np.array_equal(arr["sph_center_r"][si, :3].view(np.uint32), M[:3, 3].copy().view(np.uint32))   # Q12
            assert arr["sph_center_r"][si, 3] == sa.mesh_radius[m]
            si += 1
            prim += 1
    assert list(fs.origin_hint) == [float(x) for x in sa.cam_from]


@pytest.mark.parametrize("scene", ["simpleSphereScene.json", "scene.json", "mix_small.json"])
def test_camera_matches_oracle_primary_rays(pkg, oracle, scene):
    """InitializeRenderer + the hoisted inverse view matrix: directions rebuilt from the render
    params equal the oracle's GenerateRay (cpp:832-858) bit for bit."""
    W, H = 37, 23
    rt = pkg.Raytracer(W, H)
    rt.SetAssetsPath(ASSETS)
    assert rt.LoadSceneJSON(scene) == pkg.RT_SUCCESS
    rp = rt.render_params()
    assert (rp.width, rp.height, rp.depth, rp.ao_spp) == (W, H, 4, 128) and rp.fov_degrees == 60.0
    inv = np.array(list(rp.inv_view3x3), np.float32).reshape(3, 3)
    orc = oracle.Oracle(oracle.load_scene_json(ASSETS, scene))
    f = np.float32
    import math
    rad = f(f(30.0) * (3.14159265 / 180))
    for (x, y) in [(0, 0), (W - 1, H - 1), (5, 17), (20, 3)]:
        ndcx = f(((2.0 * x) / W - 1) * (float(f(W) / f(H)) * math.tan(float(rad))))
        ndcy = f((1 - (2.0 * y) / H) * math.tan(float(rad)))
        d = np.array([ndcx, ndcy, f(-1.0)], np.float32)
        w = np.array([f(f(f(inv[r, 0] * d[0]) + f(inv[r, 1] * d[1])) + f(inv[r, 2] * d[2])) for r in range(3)], np.float32)
        ln = np.sqrt(f(f(f(w[0] * w[0]) + f(w[1] * w[1])) + f(w[2] * w[2])))
        w = (w / ln).astype(np.float32)
        st, o, dd = orc.primary_ray(W, H, x, y)
        assert st == 0
        assert np.array_equal(w.view(np.uint32), dd.view(np.uint32))
        assert np.array_equal(np.array(list(rp.camera_from), np.float32), o)


def test_loader_error_behaviour(pkg, tmp_path):
    rt = pkg.Raytracer(4, 4)
    rt.SetAssetsPath(str(tmp_path))
    assert rt.LoadSceneJSON("missing.json") == pkg.RT_FAILURE                      # cpp:650-653
    (tmp_path / "bad.json").write_text("{ not json")
    assert rt.LoadSceneJSON("bad.json") == pkg.RT_FAILURE                          # cpp:657-663
    (tmp_path / "nokey.json").write_text(json.dumps({"scene": {"shapes": [{"id": "a"}]}}))
    assert rt.LoadSceneJSON("nokey.json") == pkg.RT_FAILURE                        # cpp:775-778
    cam = {"from": [0, 0, 5], "to": [0, 0, 0], "bounds": [0.1, 10, 1, -1, 1, -1], "resolution": [8, 8]}
    sc = {"scene": {"shapes": [{"id": "s", "geometry": "absentmesh", "material": {"Cs": [1, 1, 1], "Ka": 1, "Kd": 1, "Ks": 0, "Kt": 0, "n": 1},
                                "transforms": []}], "lights": [], "camera": cam}}
    (tmp_path / "nomesh.json").write_text(json.dumps(sc))
    assert rt.LoadSceneJSON("nomesh.json") == pkg.RT_FAILURE                       # cpp:597-600 via cpp:719
    # "last value wins" transforms, integer-typed numbers, optional notes (cpp:673, 688-716)
    import shutil
    shutil.copy(os.path.join(ASSETS, "1sphere.json"), str(tmp_path))
    sc["scene"]["shapes"][0].update({"geometry": "1sphere", "notes": "x",
                                     "transforms": [{"T": [1, 2, 3]}, {"S": [2, 2, 2]}, {"T": [4, 5, 6]}, {"Ry": 90}]})
    (tmp_path / "ok.json").write_text(json.dumps(sc))
    assert rt.LoadSceneJSON("ok.json") == pkg.RT_SUCCESS
    arr = pkg.flat_scene_arrays(rt.flat_scene())
    # S * R * T (Q11): translation (4,5,6) rotated by Ry(90) and scaled by 2
    assert np.allclose(arr["sph_center_r"][0], [12.0, 10.0, -8.0, 1.0], atol=1e-5)


def test_flush_ppm_matches_oracle_gamma(pkg, oracle, tmp_path):
    """FlushFrameBufferToPPM (cpp:796-830) on a frame buffer that was never rendered (all zeros)
    and the gamma table itself against the oracle's restatement."""
    rt = pkg.Raytracer(5, 3)
    out = str(tmp_path / "z.ppm")
    assert rt.FlushFrameBufferToPPM(out) == pkg.RT_SUCCESS
    data = open(out, "rb").read()
    assert data.startswith(b"P6\n5 3\n255\n") and data[len(b"P6\n5 3\n255\n"):] == bytes(45)
    assert rt.FlushFrameBufferToPPM(str(tmp_path / "nodir" / "z.ppm")) == pkg.RT_FAILURE   # cpp:802-806
    lut = oracle.gamma_encode(np.arange(256, dtype=np.int16))
    assert lut[0] == 0 and lut[255] == 255 and (np.diff(lut.astype(int)) >= 0).all()


def test_row_partition_helpers(pkg):
    H = 11
    for world in (1, 2, 3, 4, 8, 16):
        seen = []
        for r in range(world):
            first, step, n = pkg.rows_for_rank(H, r, world)
            seen += list(range(first, first + n * step, step))
        assert sorted(seen) == list(range(H))
    counts = [np.array([1, 0, 5, 2, 0, 1], np.uint64), np.array([3, 3, 0, 0, 7], np.uint64)]
    bases = pkg.row_bases_from_counts(H, 2, counts)
    per_row = np.zeros(H, np.uint64)
    per_row[0::2] = counts[0]
    per_row[1::2] = counts[1]
    excl = np.concatenate([[0], np.cumsum(per_row)[:-1]]).astype(np.uint64)
    assert np.array_equal(bases[0], excl[0::2]) and np.array_equal(bases[1], excl[1::2])


def test_mesh_cache_is_bit_identical_and_self_healing(pkg, tmp_path):
    """the binary mesh cache next to LoadMesh (cpp:568-643).  The flattened scene with the cache, cold and warm, is
    the scene without it, byte for byte; a cache whose JSON changed, or that is truncated / corrupted, is ignored and rewritten;
    a cache directory that cannot be written is not an error."""
    scene = "simpleScene.json"                                       # teapot + spheres: polygon and sphere meshes

    def load(cache_dir):
        rt = pkg.Raytracer(16, 16)
        rt.SetAssetsPath(ASSETS)
        if cache_dir is not None:
            rt.SetMeshCacheDir(cache_dir)
        assert rt.LoadSceneJSON(scene) == pkg.RT_SUCCESS
        arr = pkg.flat_scene_arrays(rt.flat_scene())
        return {k: np.array(v, copy=True) for k, v in arr.items() if isinstance(v, np.ndarray)}, rt.MeshCacheHits()

    def same(a, b):
        assert a.keys() == b.keys()
        for k in a:
            assert a[k].tobytes() == b[k].tobytes(), k

    plain, hits = load(None)
    assert hits == 0
    cdir = str(tmp_path / "cache")
    os.makedirs(cdir)
    cold, hits = load(cdir)
    assert hits == 0
    files = sorted(os.listdir(cdir))
    assert files and all(f.endswith(".rt580mesh") for f in files)
    same(plain, cold)
    warm, hits = load(cdir)
    assert hits == len(files)
    same(plain, warm)
    # truncated and bit-flipped cache files: fall back to the JSON, rewrite
    victim = os.path.join(cdir, files[0])
    good = open(victim, "rb").read()
    open(victim, "wb").write(good[:len(good) // 2])
    again, hits = load(cdir)
    assert hits == len(files) - 1
    same(plain, again)
    assert open(victim, "rb").read() == good
    open(victim, "wb").write(good[:16] + bytes([good[16] ^ 1]) + good[17:])    # the recorded JSON size no longer matches
    again, hits = load(cdir)
    assert hits == len(files) - 1
    same(plain, again)
    # the JSON changes (other assets directory, same mesh name, different bytes): the old cache entry must not be used
    import shutil
    adir = tmp_path / "assets"
    shutil.copytree(ASSETS, str(adir))
    cam = {"from": [0, 0, 5], "to": [0, 0, 0], "bounds": [0.1, 10, 1, -1, 1, -1], "resolution": [8, 8]}
    sc = {"scene": {"shapes": [{"id": "s", "geometry": "1sphere", "material": {"Cs": [1, 1, 1], "Ka": 1, "Kd": 1, "Ks": 0, "Kt": 0, "n": 1},
                                "transforms": []}], "lights": [], "camera": cam}}
    (adir / "one.json").write_text(json.dumps(sc))

    def radius():
        rt = pkg.Raytracer(8, 8)
        rt.SetAssetsPath(str(adir))
        rt.SetMeshCacheDir(cdir)
        assert rt.LoadSceneJSON("one.json") == pkg.RT_SUCCESS
        return float(pkg.flat_scene_arrays(rt.flat_scene())["sph_center_r"][0][3]), rt.MeshCacheHits()

    r0, h0 = radius()
    r1, h1 = radius()
    assert (h0, h1) == (0, 1) and r0 == r1
    text = (adir / "1sphere.json").read_text()
    (adir / "1sphere.json").write_text(text + "\n")                          # same mesh, other bytes: the key changes
    r2, h2 = radius()
    assert h2 == 0 and r2 == r0
    # a directory that does not exist: loads fine, writes nothing
    nowhere, hits = load(str(tmp_path / "no" / "such" / "dir"))
    assert hits == 0
    same(plain, nowhere)


def test_far_field_bound_holds_against_float_arithmetic():
    """fargrid.cuh's necessary condition for a far-field accept, against the reference's float evaluation of cpp:392
    (tools/far_bound_check.py, 4e5 adversarial samples per distance factor here; 2.4e7 in DESIGN.md 2.1)."""
    import subprocess
    import sys
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "far_bound_check.py"), "400000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "accepts below the bound: 0" in r.stdout
