"""The bench lines committed under profiles/ carry every key the driver's contract names (bench.py docstring,
task statement section 4): a guard against a later edit of bench.py dropping one.  CPU only: it reads the committed
JSON lines, it does not run the benchmark."""
import json
import os

import pytest

from conftest import ROOT

PROFILES = os.path.join(ROOT, "profiles")


def load(name):
    with open(os.path.join(PROFILES, name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def test_own_arm_line_has_the_contract_keys():
    d = load("r01_bench.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in d, k
    assert d["metric"] == "Mrays/s" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["config"]["workload"] == "c4_room" and "l2" in d["config"]
    assert abs(d["value"] - d["rays_per_frame"] / d["ms_per_step"] / 1e3) < 1e-6 * d["value"]
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] == 3840 * 2160 * 6
    assert e["value"] < d["value"]                       # copies and the scene build are inside the e2e region
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["unit"] == d["unit"] and c["sample"]
    assert d["gpu_launches"] > 0
    assert d["clocks"]["reasons"] == [] and d["clocks"]["sm_mhz"] >= 0.9 * d["clocks"]["sm_max_mhz"]


def test_reference_arm_line():
    d = load("r01_bench_reference.json")
    assert d["impl"] == "reference" and d["metric"] == "Mrays/s" and d["config"]["workload"] == "c4_room"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == d["value"]


@pytest.mark.parametrize("n", [2, 4, 8])
def test_multi_gpu_lines_were_verified(n):
    d = load("r01_bench_%dgpu.json" % n)
    one = load("r01_bench.json")
    assert d["n_gpus"] == n and d["scaling"] == "strong" and d["config"]["workload"] == one["config"]["workload"]
    assert d["verify"]["multi_gpu_frame_equals_single_gpu_frame"] is True
    assert d["rays_per_frame"] == one["rays_per_frame"]          # the same frame, split over the ranks
    assert d["value"] > one["value"]


# ---- round 2: the headline workload is the SPECIFIED open scene (c4_open, exact far field) -----------------------------

def test_round2_own_arm_line():
    d = load("r02_bench_1gpu.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks", "ray_classes", "records"):
        assert k in d, k
    assert d["config"]["workload"] == "c4_open" and d["config"]["farfield"] == "exact" and "open floor" in d["config"]["scene"]
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["vs_baseline"] is None
    assert abs(d["value"] - d["rays_per_frame"] / d["ms_per_step"] / 1e3) < 1e-6 * d["value"]
    e = d["e2e"]
    assert 0 < e["h2d_bytes_per_step"] < 1 << 20          # meshes + one matrix per shape: FlattenScene runs on the device
    assert e["d2h_bytes_per_step"] == 3840 * 2160 * 6 and e["value"] < d["value"] and e["resident"]["value"] <= d["value"] * 1.01
    r = d["roofline"]
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1
    assert r["traffic"] > 0 and "profiles" not in (r["traffic_source"] or "") and r["ncu"]["warp_instructions_per_ray"] > 0
    assert 0 < r["issue"]["frac"] < 1 and "frac" not in r["hbm"]
    # the ncu figures are loaded from the committed capture, not typed into bench.py
    cap = load("r02_ncu_k_anyhit_c4_open.json")
    assert r["ncu"]["warp_instructions_per_ray"] == cap["counters"]["warp_instructions_per_ray"]
    assert abs(r["traffic"] - cap["dram_bytes_per_ray"] * r["rays_per_launch"]) < 1e-6 * r["traffic"]
    # every ray class is timed; the far field is the larger part of this frame and the line says so
    cls = d["ray_classes"]
    assert set(cls) == {"primary", "closest", "shadow_gen", "shadow_tree", "ao_gen", "ao_tree", "far_any", "far_closest", "order", "resolve"}
    assert abs(sum(v["ms"] for v in cls.values()) - d["ms_per_step"]) < 0.1 * d["ms_per_step"]
    # the records: the closed room of round 1 and both 10 M-triangle scenes at 7680x4320
    rec = d["records"]
    assert set(rec) == {"c4_room", "c5_open", "c5_room"}
    assert rec["c5_open"]["width"] == 7680 and rec["c5_room"]["height"] == 4320 and rec["c5_room"]["primitives_in_tree"] > 9_000_000
    assert rec["c5_open"]["farfield"] == "off" and rec["c5_open"]["farfield_note"] and rec["c5_room"]["farfield"] == "exact"
    assert d["clocks"]["reasons"] == [] and d["gpu_launches"] > 0


def test_round2_reference_arm_line():
    d = load("r02_bench_reference.json")
    assert d["impl"] == "reference" and d["config"]["workload"] == "c4_open"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"]


@pytest.mark.parametrize("n", [2, 4, 8])
def test_round2_multi_gpu_lines_were_verified(n):
    d = load("r02_bench_%dgpu.json" % n)
    one = load("r02_bench_1gpu.json")
    assert d["n_gpus"] == n and d["scaling"] == "strong" and d["config"]["workload"] == "c4_open"
    assert d["verify"]["multi_gpu_frame_equals_single_gpu_frame"] is True          # unconditional for N > 1
    assert d["rays_per_frame"] == one["rays_per_frame"] and d["value"] > one["value"]
    for name, r in d["records"].items():
        assert r["verify"]["multi_gpu_frame_equals_single_gpu_frame"] is True, name
        assert r["rays_per_frame"] == one["records"][name]["rays_per_frame"], name
