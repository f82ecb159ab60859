"""The T1 oracle (oracle/oracle580.c) against the committed reference-generated goldens
(tests/golden/*.npz, produced by the reference's own code, see tests/golden/make_golden.py).
Bit-exact: integer frame buffer, IntersectScene call count, and the gamma-encoded PPM bytes."""
import hashlib
import os

import numpy as np
import pytest

from conftest import load_golden

NT = os.cpu_count() or 1

TAGS = ["c1_500_spp16", "c1_200_spp64", "c1_200_spp8_d0", "c1_200_spp8_d2", "tri_500_spp128",
        "teapots_160x90_spp16", "teapots_96_spp1", "teapots_point_96x64_spp4", "mix_small_128x72_spp4",
        "wide_37x23_spp3"]


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_matches_reference_golden(oracle, oracle_scene, tag):
    g = load_golden(tag)
    orc = oracle_scene(g["scene"])
    fb, rays, _ = orc.render(g["W"], g["H"], g["spp"], g["depth"], nthreads=NT)
    assert rays == g["rays"]
    assert np.array_equal(fb, g["fb"]), "%d channels differ" % int((fb != g["fb"]).sum())
    # the compared artefact of north_star is the PPM: header + gamma-encoded bytes (cpp:796-830)
    ppm = b"P6\n%d %d\n255\n" % (g["W"], g["H"]) + oracle.gamma_encode(fb).tobytes()
    assert hashlib.md5(ppm).hexdigest() == g["ppm_md5"]


def test_oracle_serial_stream_equals_random_access(oracle, oracle_scene):
    """One running engine (the reference's order, cpp:921-925) == per-call modular
    exponentiation."""
    orc = oracle_scene("simpleSphereScene.json")
    a, ra, ha = orc.render(96, 64, 5, 4, nthreads=1)
    b, rb, hb = orc.render(96, 64, 5, 4, nthreads=NT)
    assert ra == rb and np.array_equal(a, b) and np.array_equal(ha, hb)


def test_oracle_pixel_subset_with_bases(oracle, oracle_scene):
    """Sampled pixels + their AO ordinals reproduce the full frame (the C4/C5 checking mode)."""
    orc = oracle_scene("simpleSphereScene.json")
    W, H, spp, depth = 80, 60, 4, 4
    full, _, hits = orc.render(W, H, spp, depth, nthreads=NT)
    n_amb = 1
    base = np.concatenate([[0], np.cumsum(hits.astype(np.uint64))[:-1]]) * n_amb
    pix = np.array([0, 17, 1234, 2400, 2401, 3999, 4799], np.int32)
    sub, _, _ = orc.render(W, H, spp, depth, pix=pix, ao_base=base[pix], nthreads=2)
    assert np.array_equal(sub, full.reshape(-1, 3)[pix])


def test_lcg_closed_form(oracle):
    x = 1
    for n in range(1, 2000):
        x = (x * 16807) % 2147483647
        assert oracle.lcg_state(n) == x
    # exponent reduction mod (M-1)
    assert oracle.lcg_state(2147483646) == 1
    assert oracle.lcg_state(2147483646 + 5) == oracle.lcg_state(5)


def test_empty_and_degenerate_scenes(oracle, tmp_path):
    import json
    # empty scene: every pixel is background (cpp:30-32)
    sc = {"scene": {"shapes": [], "lights": [], "camera": {"from": [0, 0, 5], "to": [0, 0, 0],
                                                         "bounds": [0.1, 10, 1, -1, 1, -1], "resolution": [8, 8]}}}
    (tmp_path / "empty.json").write_text(json.dumps(sc))
    orc = oracle.Oracle(oracle.load_scene_json(str(tmp_path), "empty.json"))
    fb, rays, hits = orc.render(8, 6, 4, 4)
    assert rays == 48 and hits.sum() == 0
    assert (fb.reshape(-1, 3) == np.array([254, 64, 205], np.int16)).all()
