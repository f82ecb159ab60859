import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
ASSETS = os.path.join(GOLDEN, "assets")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(tag):
    z = np.load(os.path.join(GOLDEN, tag + ".npz"))
    return {"fb": z["fb"], "rays": int(z["rays"]), "ppm_md5": str(z["ppm_md5"]), "scene": str(z["scene"]),
            "W": int(z["W"]), "H": int(z["H"]), "spp": int(z["spp"]), "depth": int(z["depth"])}


@pytest.fixture(scope="session")
def oracle():
    import oracle as o
    o.build(ref=False)          # gcc build of the T1 oracle (seconds); T0 is prebuilt where available
    return o


@pytest.fixture(scope="session")
def pkg():
    from __graft_entry__ import load_package
    p = load_package()
    if not os.path.exists(p.LIB_PATH):
        p.build()
    return p


@pytest.fixture(scope="session")
def assets():
    return ASSETS


_scene_cache = {}


@pytest.fixture(scope="session")
def oracle_scene(oracle):
    def get(name):
        if name not in _scene_cache:
            _scene_cache[name] = oracle.Oracle(oracle.load_scene_json(ASSETS, name))
        return _scene_cache[name]
    return get


def ppm_channel_errors(pkg_or_oracle_gamma, fb_a, fb_b):
    a = pkg_or_oracle_gamma(fb_a).astype(np.int32)
    b = pkg_or_oracle_gamma(fb_b).astype(np.int32)
    return np.abs(a - b)
