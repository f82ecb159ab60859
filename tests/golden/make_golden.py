#!/usr/bin/env python
"""Regenerates tests/golden/ from the reference itself (run in the build container, where
/root/reference exists; the GPU box only ever reads the committed results).

  assets/*.json   the reference's own scene + mesh inputs (Assets/ of the reference repo,
                  re-serialised compactly: same keys, same numbers) and synthetic scenes in
                  the same schema written by 580-raytracer_b200/scenegen.py
  *.npz           raw int16 frame buffers + IntersectScene call counts + PPM md5 produced by
                  the T0 oracle = the reference's unmodified Raytracer.cpp/.h compiled by
                  oracle/build_ref.sh

The md5 values of SURVEY.md Appendix B are asserted here, so a toolchain whose libm /
libstdc++ changes the reference's image is noticed at generation time.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402

ASSETS = os.path.join(HERE, "assets")
REF_ASSETS = oracle.REFERENCE_ASSETS

SURVEY_MD5 = {
    ("simpleSphereScene.json", 500, 500, 128, 4): "a00a8b5cb0a0e7dd3bd85f675a43e94a",
    ("simpleSphereScene.json", 500, 500, 16, 4): "49a95d7cfe9b3d5d8a1209328cb29002",
    ("simpleSphereSceneAO.json", 500, 500, 128, 4): "b8c3a767334244507fe75184fcb901a5",
    ("simpleScene.json", 500, 500, 128, 4): "d04a4c066694a363b0e6964c9472cf31",
    ("scene.json", 160, 90, 16, 4): "079389e8cb45207c1d00b84e4a147094",
    ("scene.json", 96, 96, 1, 4): "075fae364a2282fb7fb3cb411706a913",
}

# (tag, scene, W, H, spp, depth)
GOLDENS = [
    ("c1_500_spp128", "simpleSphereScene.json", 500, 500, 128, 4),      # BASELINE configs[0]: the reference's main()
    ("c1_500_spp16", "simpleSphereScene.json", 500, 500, 16, 4),
    ("c1_200_spp64", "simpleSphereScene.json", 200, 200, 64, 4),
    ("c1_200_spp8_d0", "simpleSphereScene.json", 200, 200, 8, 0),
    ("c1_200_spp8_d2", "simpleSphereScene.json", 200, 200, 8, 2),
    ("ao_500_spp128", "simpleSphereSceneAO.json", 500, 500, 128, 4),
    ("tri_500_spp128", "simpleScene.json", 500, 500, 128, 4),
    ("teapots_160x90_spp16", "scene.json", 160, 90, 16, 4),             # BASELINE configs[1..2] at oracle-feasible size
    ("teapots_96_spp1", "scene.json", 96, 96, 1, 4),
    ("teapots_point_96x64_spp4", "scene_point.json", 96, 64, 4, 4),     # directional + point + two ambient lights
    ("mix_small_128x72_spp4", "mix_small.json", 128, 72, 4, 4),         # synthetic generator, small
    ("wide_37x23_spp3", "simpleSphereScene.json", 37, 23, 3, 4),        # ragged sizes, odd spp
]


def copy_reference_assets():
    os.makedirs(ASSETS, exist_ok=True)
    for name in sorted(os.listdir(REF_ASSETS)):
        if not name.endswith(".json") or name == "1plane.json":   # 1plane: unsupported type (Q26), used by no scene
            continue
        with open(os.path.join(REF_ASSETS, name)) as f:
            doc = json.load(f)
        with open(os.path.join(ASSETS, name), "w") as f:
            json.dump(doc, f, separators=(",", ":"))


def make_point_scene():
    with open(os.path.join(ASSETS, "scene.json")) as f:
        doc = json.load(f)
    doc["scene"]["lights"].append({"id": "pointLight", "type": "point", "color": [0.9, 0.8, 1.0], "intensity": 0.7,
                                   "position": [0.5, 4.0, 3.0]})
    doc["scene"]["lights"].append({"id": "ambient2", "type": "ambient", "color": [0.2, 0.4, 0.3], "intensity": 0.5})
    with open(os.path.join(ASSETS, "scene_point.json"), "w") as f:
        json.dump(doc, f, separators=(",", ":"))


def make_synthetic():
    pkg_dir = os.path.join(ROOT, "580-raytracer_b200")
    import importlib.util
    spec = importlib.util.spec_from_file_location("scenegen", os.path.join(pkg_dir, "scenegen.py"))
    sg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sg)
    info = sg.write_synthetic_scene(ASSETS, "mix_small", n_teapots=6, n_spheres=12, seed=580)
    print("mix_small:", info)


def main():
    oracle.build(ref=True)
    copy_reference_assets()
    make_point_scene()
    make_synthetic()
    for tag, scene, W, H, spp, depth in GOLDENS:
        ppm = os.path.join("/tmp", tag + ".ppm")
        st, fb, rays, secs = oracle.t0_render(ASSETS, scene, W, H, spp, depth, ppm_out=ppm)
        assert st == 0, (tag, st)
        with open(ppm, "rb") as f:
            md5 = hashlib.md5(f.read()).hexdigest()
        want = SURVEY_MD5.get((scene, W, H, spp, depth))
        if want is not None:
            assert md5 == want, "T0 output for %s differs from SURVEY Appendix B: %s != %s" % (tag, md5, want)
        np.savez_compressed(os.path.join(HERE, tag + ".npz"), fb=fb, rays=np.uint64(rays), ppm_md5=md5,
                            scene=scene, W=W, H=H, spp=spp, depth=depth)
        print("%-28s rays=%-10d %.1fs md5=%s%s" % (tag, rays, secs, md5, " (= SURVEY anchor)" if want else ""))


if __name__ == "__main__":
    main()
