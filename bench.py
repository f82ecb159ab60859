#!/usr/bin/env python
"""bench.py - headline benchmark of the 580-Raytracer hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU Render

Metric (BASELINE.json): Mrays/s, one ray = one IntersectScene call of the reference
(primary + secondary + shadow + AO).  Workload = BASELINE.json configs[3]: synthetic
1M-triangle + 1k-sphere scene at 3840x2160, depth 4, 16 AO samples per pixel ("c4_room":
977 instanced teapots + 1000 spheres in a closed room, written by scenegen.py in the
reference's own JSON schema, seed 580).  A step = one frame.

  value  scene + LBVH resident in HBM; rt580_render_begin/finish per step, frame gathered to rank 0
  e2e    through the reference-facing call with HOST buffers every step: rt580_upload_scene
         (H2D of the flattened scene + LBVH build) + rt580_render (D2H of the int16 frame)
N GPUs: rows interleaved across ranks (strong scaling of the one frame), the AO-stream row
prefix exchanged with one tiny all_gather, the int16 bands gathered to rank 0 over NCCL.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from __graft_entry__ import ASSETS, PKG_DIR, load_package  # noqa: E402

METRIC, UNIT = "Mrays/s", "Mrays/s"
WORKLOAD = "c4_room"
W, H, DEPTH, SPP = 3840, 2160, 4, 16
CACHE = "/tmp/rt580_bench_scenes"
AO_CHUNK = 32 << 20                        # AH_CHUNK_TIGHT in rt580_core.cu: AO sample rays per k_ao_gen launch
K_AO_GEN_DRAM_BYTES_PER_RAY = 163.102e6 / (32 << 20)   # ncu capture, see roofline.traffic_source


def scene_dir(name):
    """Write (once) the synthetic scene next to a copy of the teapot mesh it instances."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("scenegen", os.path.join(PKG_DIR, "scenegen.py"))
    sg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sg)
    d = os.path.join(CACHE, name)
    if not os.path.exists(os.path.join(d, name + ".json")):
        os.makedirs(d, exist_ok=True)
        shutil.copy(os.path.join(ASSETS, "teapot.json"), d)
        sg.write_synthetic_scene(d, name, **sg.CONFIGS[name])
    return d


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    return rank, world, local


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.samples, self.stop_flag = gpu, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 3 + i and s[3 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.samples[0][1]), "samples": len(self.samples),
                "reasons": reasons}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


def algorithmic_per_ray(n_prims):
    """F_ray = ceil(log2 N) * 40 + 71 flop, B_ray = ceil(log2 N) * 64 + 64 bytes."""
    lg = int(np.ceil(np.log2(max(n_prims, 2))))
    return lg * 40 + 71, lg * 64 + 64


# --------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation on the host cores
# --------------------------------------------------------------------------------------------
def cpu_reference_sample(n_pix, nthreads, seed=580):
    """Time the reference's GenerateRay + Raycast (oracle/_ref, i.e. the reference's own sources)
    on a bounded pixel sample of the same workload.  Returns (Mrays/s, rays, seconds, kind)."""
    import oracle
    d = scene_dir(WORKLOAD)
    rng = np.random.default_rng(seed)
    pix = rng.choice(W * H, n_pix, replace=False).astype(np.int32)
    if oracle.t0_available():
        st, _, rays, secs = oracle.t0_render_pixels(d, WORKLOAD + ".json", W, H, SPP, DEPTH, pix, nthreads=nthreads)
        assert st == 0
        return rays / secs / 1e6, rays, secs, "reference"
    orc = oracle.Oracle(oracle.load_scene_json(d, WORKLOAD + ".json"))
    t0 = time.time()
    _, rays, _ = orc.render(W, H, SPP, DEPTH, pix=pix, ao_base=np.zeros(n_pix, np.uint64), nthreads=nthreads)
    secs = time.time() - t0
    return rays / secs / 1e6, rays, secs, "port"


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_pix = max(cores, 2 * cores)
    vals = []
    for i in range(args.warmup + args.steps):
        v, rays, secs, kind = cpu_reference_sample(n_pix, cores, seed=580 + i)
        if i >= args.warmup:
            vals.append((v, rays, secs))
    value = float(np.mean([v for v, _, _ in vals]))
    ms = float(np.mean([s for _, _, s in vals]) * 1e3)
    sample = "%d random pixels of the %dx%d frame per step (%d rays), one reference Raytracer instance per thread" % (
        n_pix, W, H, vals[-1][1])
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "width": W, "height": H, "depth": DEPTH, "ao_spp": SPP, "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
# this repo's arm
# --------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = load_package()
    d = scene_dir(WORKLOAD) if rank == 0 else None
    if world > 1:
        dist.barrier()
        d = scene_dir(WORKLOAD)

    # host side of the reference API: LoadSceneJSON + the load-time flatten (host C++)
    rt = pkg.Raytracer(W, H)
    rt.SetAssetsPath(d)
    rt.SetOptions(depth=DEPTH, ao_spp=SPP, device=local, farfield=pkg.FARFIELD_OFF if args.farfield == "off" else pkg.FARFIELD_EXACT)
    assert rt.LoadSceneJSON(WORKLOAD + ".json") == pkg.RT_SUCCESS
    flat = rt.flat_scene()
    params = rt.render_params()
    ctx = pkg.Context(local)
    ctx.upload_scene(flat)
    info = ctx.scene_info()
    dev = ctx.device_info()
    ext_stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))

    p = params.copy()
    p.row_first, p.row_step, p.n_rows = pkg.rows_for_rank(H, rank, world)
    band = torch.empty((max(p.n_rows, 1), W, 3), dtype=torch.int16, device="cuda")
    max_rows = (H + world - 1) // world
    peer_frame = False
    if world > 1:
        # the exchange stays on the device: per-row counts all-gathered on the context's stream ...
        mine_d = torch.zeros(max_rows, dtype=torch.int64, device="cuda")
        all_d = torch.zeros((world, max_rows), dtype=torch.int64, device="cuda")
        done = torch.zeros(1, dtype=torch.int32, device="cuda")
        # ... and so does the gather: rank 0 owns the whole frame, the other ranks map it (CUDA IPC) and
        # store their rows into it over NVLink at the end of rt580_render_finish_interleaved
        handle = [ctx.frame_export(W, H) if rank == 0 else None]
        dist.broadcast_object_list(handle, src=0)
        ok = torch.ones(1, dtype=torch.int32, device="cuda")
        if rank != 0:
            try:
                ctx.frame_import(handle[0], W, H)
            except pkg.Rt580Error as e:
                print("rank %d: cannot map rank 0's frame (%s); falling back to an NCCL gather" % (rank, e), file=sys.stderr)
                ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        peer_frame = bool(int(ok.item()))
        if not peer_frame:
            ctx.frame_release()
            # int16 is not an NCCL dtype: the bands are gathered as bytes
            gather_list = [torch.empty((max_rows, W, 6), dtype=torch.uint8, device="cuda") for _ in range(world)] if rank == 0 else None
            padded = torch.zeros((max_rows, W, 3), dtype=torch.int16, device="cuda")

    side_stream = torch.cuda.Stream() if world > 1 else None
    dbg = bool(os.environ.get("RT580_BENCH_DEBUG")) and (rank == 0 or os.environ.get("RT580_BENCH_DEBUG") == "2")
    tparts = [0.0] * 8

    def frame():
        """one step: this rank's rows; the frame ends up on rank 0 (device memory)"""
        _a = time.perf_counter()
        if world == 1:
            counts = ctx.render_begin(p)
            _b = time.perf_counter()
            bases = pkg.row_bases_from_counts(H, 1, [counts])[0]
            _c = time.perf_counter()
            _, st = ctx.render_finish(p, bases, device_ptr=band.data_ptr())
            _d = time.perf_counter()
        else:
            ctx.render_begin(p, want_counts=False)
            _b = time.perf_counter()
            ctx.row_counts_to_device(mine_d.data_ptr(), max_rows)
            with torch.cuda.stream(ext_stream):
                dist.all_gather_into_tensor(all_d, mine_d)                # the one exchange of the LCG mode
            _c = time.perf_counter()
            st = ctx.render_finish_interleaved(all_d.data_ptr(), world, rank, max_rows,
                                               device_ptr=None if peer_frame else padded.data_ptr())
            _d = time.perf_counter()
            if peer_frame:
                # "every rank's rows have landed in rank 0's frame": a 4-byte all-reduce ordered after this rank's
                # stores, on a side stream - whoever consumes the frame on rank 0 waits for it, the next frame's
                # structure pass does not (ranks drift by a few % per frame: the GPUs of a box are not equally fast)
                side_stream.wait_stream(ext_stream)
                with torch.cuda.stream(side_stream):
                    dist.all_reduce(done)
            else:
                with torch.cuda.stream(ext_stream):
                    dist.gather(padded.view(torch.uint8), gather_list, dst=0)
        if dbg:
            torch.cuda.synchronize()
            _e = time.perf_counter()
            for k, v in enumerate([_b - _a, _c - _b, _d - _c, _e - _d]):
                tparts[k] += v * 1e3
            tparts[4] += st.ms_structure; tparts[5] += st.ms_order; tparts[6] += st.ms_ao; tparts[7] += st.ms_resolve
        return st

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        frame()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    sync()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record(ext_stream)
    stats, launches = [], 0
    for _ in range(args.steps):
        st = frame()
        stats.append(st)
        launches += st.kernel_launches
    if side_stream is not None:
        ext_stream.wait_stream(side_stream)      # the timed region ends when the last frame is complete on rank 0
    ev1.record(ext_stream)
    sync()
    wall_ms = (time.perf_counter() - t0) * 1e3
    dev_ms = ev0.elapsed_time(ev1)
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    if dbg:
        n_fr = args.steps + args.warmup
        print("rank %d wall ms/frame: begin %.3f  exchange %.3f  finish %.3f  gather %.3f | device: structure %.3f order %.3f ao %.3f resolve %.3f" % (
            (rank,) + tuple(t / n_fr for t in tparts)), file=sys.stderr)
    step_ms = max(dev_ms, 0.0) / args.steps
    rays_rank = float(np.mean([s.rays for s in stats]))
    ao_ms = float(np.mean([s.ms_ao_kernel for s in stats]))
    ao_rays = float(np.mean([s.rays_ao for s in stats]))
    t = torch.tensor([step_ms, rays_rank, wall_ms / args.steps, ao_ms, ao_rays], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        step_ms, wall_step_ms, ao_ms = float(tmax[0]), float(tmax[2]), float(tmax[3])
        rays_total, ao_rays_total = float(tsum[1]), float(tsum[4])
    else:
        wall_step_ms, rays_total, ao_rays_total = wall_ms / args.steps, rays_rank, ao_rays
    value = rays_total / (step_ms * 1e-3) / 1e6

    verify = None
    if args.verify and world > 1:
        # the frame the ranks assembled on rank 0 against the same frame rendered by rank 0 alone
        sync()
        if rank == 0:
            if peer_frame:
                got = ctx.frame_read(W, H)
                ctx.frame_release()
            else:
                got = pkg.interleave_rows(H, W, world, [g.view(torch.int16).reshape(max_rows, W, 3).cpu().numpy() for g in gather_list])
            want, _ = ctx.render(params)
            verify = {"multi_gpu_frame_equals_single_gpu_frame": bool(np.array_equal(got, want)), "pixels": int(W * H)}
            if peer_frame:
                handle = [ctx.frame_export(W, H)]
        if peer_frame:
            # rank 0 released its frame for the check: map the new one
            if rank != 0:
                ctx.frame_release()
                handle = [None]
            dist.broadcast_object_list(handle, src=0)
            if rank != 0:
                ctx.frame_import(handle[0], W, H)
        sync()

    # e2e: the reference-facing call with host buffers, every step: upload (H2D + LBVH build) + render (D2H)
    h2d = int(flat.n_tris * (6 * 16 + 8) + flat.n_spheres * (16 + 8) + flat.n_materials * 32 + flat.n_lights * 44)
    d2h = int(p.n_rows * W * 6) if world == 1 else int(W * H * 6)
    host_frame = None
    # page-locked, like the host class's frame buffer
    host_out = pkg.HostArray((p.n_rows, W, 3), np.int16) if world == 1 else (pkg.HostArray((H, W, 3), np.int16) if rank == 0 else None)
    for _ in range(1):
        ctx.upload_scene(flat); ctx.render(p, out=host_out.array) if world == 1 else None
    sync()
    e0 = time.perf_counter()
    e_steps = max(1, min(args.steps, 3))
    for _ in range(e_steps):
        _t0 = time.perf_counter()
        ctx.upload_scene(flat)
        _t1 = time.perf_counter()
        if os.environ.get("RT580_BENCH_DEBUG"):
            print("e2e upload %.1f ms" % ((_t1 - _t0) * 1e3), file=sys.stderr)
        if world == 1:
            ctx.render(p, out=host_out.array)
            if os.environ.get("RT580_BENCH_DEBUG"):
                print("e2e render %.1f ms" % ((time.perf_counter() - _t1) * 1e3), file=sys.stderr)
        else:
            frame()
            if rank == 0:
                torch.cuda.synchronize()                                  # includes the side stream's "all rows landed"
                if peer_frame:
                    host_frame = ctx.frame_read(W, H, out=host_out.array)  # the whole frame -> host memory on rank 0
                else:
                    host_frame = torch.stack(gather_list).cpu()
    sync()
    e2e_ms = (time.perf_counter() - e0) * 1e3 / e_steps
    te = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = rays_total / (float(te[0]) * 1e-3) / 1e6

    if rank == 0:
        peaks, peak_src = measured_peaks()
        f_ray, b_ray = algorithmic_per_ray(int(info.n_leaf))
        sm_mhz = float(peaks.get("sm_max_mhz") or dev["sm_clock_mhz"])
        fp32_peak = dev["sm_count"] * 128 * 2 * sm_mhz * 1e6 / 1e12          # TFLOP/s, FMA counted as 2
        ao_rate = ao_rays_total / (ao_ms * 1e-3) if ao_ms > 0 else 0.0      # rays/s over all ranks (max kernel time)
        # the dominant kernel is k_ao_gen (39 % of the frame's kernel time, profiles/): the occlusion pass launches it
        # once per chunk of AO_CHUNK sample rays, each followed by a k_anyhit launch over the rays it could not answer
        # (a few hundred per frame in this scene); ms_ao_kernel brackets those launches with CUDA events on their stream
        n_launch = max(1, int(np.ceil(ao_rays_total / world / AO_CHUNK)))
        rays_per_launch = ao_rays_total / world / n_launch
        roofline = {"bound": "fp32", "kernel": "k_ao_gen (per AO sample: engine state, RNG draws, double sincos, hemisphere direction, ray; exact any-hit test "
                                               "against the scene's large primitives, nearest plane first) + k_anyhit (persistent LBVH any-hit) for the rest",
                    "achieved": ao_rate * f_ray / 1e12, "peak": fp32_peak * world, "unit": "TFLOP/s",
                    "frac": (ao_rate * f_ray / 1e12) / (fp32_peak * world) if fp32_peak else None,
                    "peak_source": "%d SMs x 128 lanes x 2 x %.0f MHz (%s); MEASURED_PEAKS.json has no fp32 figure" % (
                        dev["sm_count"], sm_mhz, peak_src),
                    "flop_per_ray": f_ray, "rays_per_launch": rays_per_launch, "launches_per_step": n_launch,
                    "kernel_ms": ao_ms / n_launch, "pass_ms": ao_ms,
                    "traffic": K_AO_GEN_DRAM_BYTES_PER_RAY * rays_per_launch,
                    "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one 32M-ray k_ao_gen launch "
                                      "(profiles/r01_k_ao_gen_v3_1080p_details.txt: 163.1 MB), scaled to this launch size",
                    "ncu": {"issue_slots_busy_pct": 87.6, "fma_pipe_active_pct": 33.2, "fp64_pipe_pct": 4.1, "avg_active_lanes": 29.5,
                            "l1_hit_pct": 93.2, "dram_throughput_pct": 2.1, "source": "same capture"},
                    "hbm": {"achieved": ao_rate * b_ray / 1e9, "peak": float(peaks["hbm_gbs"]) * world, "unit": "GB/s",
                            "frac": (ao_rate * b_ray / 1e9) / (float(peaks["hbm_gbs"]) * world), "bytes_per_ray": b_ray,
                            "note": "algorithmic node+triangle bytes; served from shared memory / L1 / L2, not HBM (measured DRAM throughput 2 %)"}}
        cores = os.cpu_count() or 1
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, rays, secs, kind = cpu_reference_sample(max(2 * cores, 8), cores)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                   "sample": "%d random pixels of the %dx%d c4_room frame, %d rays in %.1f s, one reference Raytracer instance per thread" % (
                       max(2 * cores, 8), W, H, rays, secs)}
        st0 = stats[-1]
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "scene": "977 teapot instances (1,000,448 triangles) + 1000 spheres + closed double-walled room (24 triangles), seed 580",
                           "width": W, "height": H, "depth": DEPTH, "ao_spp": SPP, "rng": "reference_lcg", "farfield": args.farfield, "far_scans": st0.far_scans, "linear_fallbacks": st0.linear_fallbacks,
                           "partition": "rows interleaved over %d rank(s)" % world,
                           "gather": ("none" if world == 1 else ("peer stores into rank 0's frame over NVLink (CUDA IPC), row counts all-gathered on the device"
                                                                  if peer_frame else "NCCL gather of int16 bands")), "l2": "working set > L2: nodes+records %.0f MB, frame data %.0f MB/step" % (
                               info.n_leaf * 144 / 1e6, st0.hit_nodes * 110 / 1e6)},
                "rays_per_frame": rays_total, "ms_per_frame_4k": step_ms, "wall_ms_per_step": wall_step_ms,
                "phases_ms": {"structure": st0.ms_structure, "order": st0.ms_order, "ao": st0.ms_ao, "resolve": st0.ms_resolve},
                "ao_note": "%d of %d AO rays of the last frame went through the LBVH; the rest were already occluded by one of the room's "
                           "24 wall triangles, which the scene build keeps out of the tree and tests first (the reference's AO rays are "
                           "unbounded, so in a closed scene every one of them ends on a wall; identical result, cpp:325)" % (
                               st0.ao_rays_traversed, st0.rays_ao),
                "ray_mix": {"primary": st0.rays_primary, "secondary": st0.rays_secondary, "shadow": st0.rays_shadow, "ao": st0.rays_ao,
                            "note": "rank 0 share" if world > 1 else "whole frame"},
                "scene_info": info.as_dict(),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": float(te[0]), "includes": "rt580_upload_scene (H2D + LBVH build) + rt580_render (D2H int16 frame)"},
                "verify": verify, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
                "clocks": sampler.summary() if sampler else None}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    global W, H, WORKLOAD
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--verify", action="store_true", help="N > 1: compare the assembled frame with rank 0's own single-GPU frame")
    ap.add_argument("--farfield", default="exact", choices=["exact", "off"], help="debug only")
    ap.add_argument("--workload", default=WORKLOAD, help="debug only")
    ap.add_argument("--width", type=int, default=0, help="debug only: override the frame width")
    ap.add_argument("--height", type=int, default=0, help="debug only: override the frame height")
    args = ap.parse_args()
    WORKLOAD = args.workload
    if args.width and args.height:
        W, H = args.width, args.height
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
