"""Pins the T1 oracle (oracle/oracle580.c) to the reference itself (T0 = oracle/_ref/libref580.so,
the reference's unmodified Raytracer.cpp/.h compiled by oracle/build_ref.sh) function by function
and frame by frame.  Skipped where the reference library has not been built."""
import os

import numpy as np
import pytest

from conftest import ASSETS

NT = os.cpu_count() or 1


@pytest.fixture(scope="module")
def t0(oracle):
    if not oracle.t0_available():
        pytest.skip("oracle/_ref/libref580.so not built (needs /root/reference)")
    oracle.t0_lib()
    return oracle


def test_model_matrix(t0):
    rng = np.random.default_rng(0)
    for _ in range(300):
        srt = np.concatenate([rng.uniform(0.3, 3, 3), rng.uniform(-360, 360, 3), rng.uniform(-20, 20, 3)]).astype(np.float32)
        if rng.random() < 0.3:
            srt[3:6] = rng.choice([0, 90, 180, 270, 45], 3)
        a, b = t0.model_matrix(srt), t0.t0_model_matrix(srt)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), srt


def test_fresnel_and_refraction(t0):
    rng = np.random.default_rng(1)
    for _ in range(2000):
        n = rng.normal(size=3); n /= np.linalg.norm(n)
        i = rng.normal(size=3); i /= np.linalg.norm(i)
        a = t0.fresnel(2.5, n.astype(np.float32), i.astype(np.float32))
        b = t0.t0_fresnel(2.5, n.astype(np.float32), i.astype(np.float32))
        assert np.float32(a[0]).view(np.uint32) == np.float32(b[0]).view(np.uint32)
        assert np.float32(a[1]).view(np.uint32) == np.float32(b[1]).view(np.uint32)
        assert np.array_equal(a[2].view(np.uint32), b[2].view(np.uint32))


def test_ao_sample_stream(t0):
    """std::default_random_engine + uniform_real_distribution<float> as libstdc++ implements them
     == the closed form the oracle and the GPU use."""
    for normal in [(0, 1, 0), (0, -1, 0), (0.6, -0.48, 0.64)]:
        a = t0.hemisphere_stream(normal, 0, 20000)
        b = t0.t0_hemisphere_stream(normal, 20000)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_loader_agrees_with_reference_loader(t0):
    for scene in ["simpleSphereScene.json", "scene.json", "scene_point.json", "mix_small.json"]:
        st, sf, ntri, rad, lf, cam = t0.t0_dump_scene(ASSETS, scene)
        assert st == 0
        sa = t0.load_scene_json(ASSETS, scene)
        assert np.array_equal(sf[:, :8], sa.shape_material)
        assert np.array_equal(sf[:, 8:], sa.shape_srt)
        for i, m in enumerate(sa.shape_mesh):
            assert ntri[i] == (sa.mesh_tri_begin[m + 1] - sa.mesh_tri_begin[m] if sa.mesh_type[m] == 0 else 0)
            assert rad[i] == sa.mesh_radius[m]
        assert np.array_equal(lf[:, 0].astype(np.int32), sa.light_type)
        assert np.array_equal(lf[:, 1:].view(np.uint32), sa.light_f.view(np.uint32))
        assert np.array_equal(cam[:3], sa.cam_from) and np.array_equal(cam[3:], sa.cam_to)


@pytest.mark.parametrize("scene,W,H,spp,depth", [
    ("simpleSphereScene.json", 90, 60, 3, 4), ("simpleSphereScene.json", 64, 64, 2, 1),
    ("simpleSphereSceneAO.json", 70, 50, 5, 4), ("simpleScene.json", 120, 80, 4, 4),
    ("scene.json", 48, 27, 2, 4), ("scene_point.json", 40, 30, 1, 3), ("mix_small.json", 40, 24, 2, 4),
])
def test_frames(t0, scene, W, H, spp, depth):
    st, ref, rays0, _ = t0.t0_render(ASSETS, scene, W, H, spp, depth)
    assert st == 0
    orc = t0.Oracle(t0.load_scene_json(ASSETS, scene))
    fb, rays1, _ = orc.render(W, H, spp, depth, nthreads=1)
    assert rays1 == rays0 and np.array_equal(fb, ref)
    fb2, rays2, _ = orc.render(W, H, spp, depth, nthreads=NT)
    assert rays2 == rays0 and np.array_equal(fb2, ref)
