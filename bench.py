#!/usr/bin/env python
"""bench.py - headline benchmark of the 580-Raytracer hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU Render

Metric (BASELINE.json): Mrays/s, one ray = one IntersectScene call of the reference (primary + secondary + shadow + AO).
Headline workload = BASELINE.json configs[3] as SURVEY.md 8d specifies it: "c4_open" - 977 instanced teapots (1,000,448
triangles) + 1000 spheres over an open floor, ambient + directional (1,10,1)->0 + one point light, 3840x2160, depth 4,
16 AO samples per pixel, the reference's own random stream, far field exact (scene files written by scenegen.py in the
reference's JSON schema, seed 580).  Rays escape, AO rays walk the LBVH.  A step = one frame.

  value    scene + LBVH + far-field grid resident in HBM; rt580_render_begin / finish per step, frame gathered to rank 0
  e2e      through the reference-facing calls with HOST buffers every step: rt580_upload_scene (H2D of the flattened scene
           + LBVH + far-field grid build) + rt580_render (D2H of the int16 frame); e2e.resident: the render call alone
  records  the same measurement, fewer steps, for c4_room (the closed room round 1 benchmarked) and for BASELINE
           configs[4] "c5_open" (9766 teapots = 10,000,384 triangles at 7680x4320)
N GPUs: rows interleaved across ranks (strong scaling of the one frame), the AO-stream row prefix exchanged with one tiny
all_gather on the device, the rows stored into rank 0's frame over NVLink.  N > 1 always verifies (outside the timed
region) that the assembled frame equals rank 0's own single-GPU frame bit for bit.
"""
import argparse
import ctypes
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
WORKLOAD = "c4_open"
DEPTH, SPP = 4, 16
CACHE = "/tmp/rt580_bench_scenes"
HOST_FLATTEN = False     # --host-flatten
NCU_CAPTURE = os.path.join(ROOT, "profiles", "r02_ncu_k_anyhit_c4_open.json")     # written by tools/ncu_capture_to_json.py

WORKLOADS = {
    # name: (width, height, scene description)
    "c4_open": (3840, 2160, "977 teapot instances (1,000,448 triangles) + 1000 spheres over an open floor (2 triangles); "
                            "ambient + directional (1,10,1)->0 + 1 point light; seed 580 (SURVEY 8d C4)"),
    "c4_room": (3840, 2160, "977 teapot instances (1,000,448 triangles) + 1000 spheres in a closed double-walled room (24 triangles); "
                            "ambient + 3 point lights; seed 580 (round 1's benchmark scene)"),
    "c5_open": (7680, 4320, "9766 teapot instances (10,000,384 triangles) over an open floor (2 triangles); ambient + directional + "
                            "1 point light; seed 580 (SURVEY 8d C5)"),
    "c5_room": (7680, 4320, "9766 teapot instances (10,000,384 triangles) in a closed double-walled room; seed 580"),
}


def scene_dir(name):
    """Write (once) the synthetic scene next to a copy of the teapot mesh it instances."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("scenegen", os.path.join(PKG_DIR, "scenegen.py"))
    sg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sg)
    d = os.path.join(CACHE, name)
    if not os.path.exists(os.path.join(d, name + ".json")):
        tmp = d + ".tmp%d" % os.getpid()
        os.makedirs(tmp, exist_ok=True)
        shutil.copy(os.path.join(ASSETS, "teapot.json"), tmp)
        sg.write_synthetic_scene(tmp, name, **sg.CONFIGS[name])
        try:
            os.rename(tmp, d)
        except OSError:
            shutil.rmtree(tmp, ignore_errors=True)       # another rank was faster
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


# --------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation on the host cores
# --------------------------------------------------------------------------------------------
def cpu_reference_sample(workload, n_pix, nthreads, seed=580):
    """Time the reference's GenerateRay + Raycast (oracle/_ref, i.e. the reference's own sources) on a bounded pixel sample
    of the workload.  Returns (Mrays/s, rays, seconds, kind)."""
    import oracle
    W, H, _ = WORKLOADS[workload]
    d = scene_dir(workload)
    rng = np.random.default_rng(seed)
    pix = rng.choice(W * H, n_pix, replace=False).astype(np.int32)
    if oracle.t0_available():
        st, _, rays, secs = oracle.t0_render_pixels(d, workload + ".json", W, H, SPP, DEPTH, pix, nthreads=nthreads)
        assert st == 0
        return rays / secs / 1e6, rays, secs, "reference"
    orc = oracle.Oracle(oracle.load_scene_json(d, workload + ".json"))
    t0 = time.time()
    _, rays, _ = orc.render(W, H, SPP, DEPTH, pix=pix, ao_base=np.zeros(n_pix, np.uint64), nthreads=nthreads)
    secs = time.time() - t0
    return rays / secs / 1e6, rays, secs, "port"


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    W, H, _ = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    n_pix = max(cores, 2 * cores)
    vals = []
    for i in range(args.warmup + args.steps):
        v, rays, secs, kind = cpu_reference_sample(args.workload, n_pix, cores, seed=580 + i)
        if i >= args.warmup:
            vals.append((v, rays, secs))
    value = float(np.mean([v for v, _, _ in vals]))
    ms = float(np.mean([s for _, _, s in vals]) * 1e3)
    sample = "%d random pixels of the %dx%d %s frame per step (%d rays), one reference Raytracer instance per thread, %d threads" % (
        n_pix, W, H, args.workload, vals[-1][1], cores)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "width": W, "height": H, "depth": DEPTH, "ao_spp": SPP, "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
# this repo's arm
# --------------------------------------------------------------------------------------------
class Rig:
    """One rank's context for one workload: scene loaded through the host class, uploaded, rows of this rank, the exchange."""

    def __init__(self, pkg, torch, dist, workload, farfield, rank, world, local):
        self.pkg, self.torch, self.dist = pkg, torch, dist
        self.rank, self.world, self.local = rank, world, local
        self.workload = workload
        self.host_flatten = HOST_FLATTEN
        self.W, self.H, self.scene_text = WORKLOADS[workload]
        W, H = self.W, self.H
        d = scene_dir(workload) if rank == 0 else None
        if world > 1:
            dist.barrier()
            d = scene_dir(workload)
        t0 = time.perf_counter()
        self.rt = pkg.Raytracer(W, H)                   # host side of the reference API: LoadSceneJSON + the load-time flatten (host C++)
        self.rt.SetAssetsPath(d)
        self.rt.SetOptions(depth=DEPTH, ao_spp=SPP, device=local, farfield=pkg.FARFIELD_OFF if farfield == "off" else pkg.FARFIELD_EXACT)
        assert self.rt.LoadSceneJSON(workload + ".json") == pkg.RT_SUCCESS
        self.load_s = time.perf_counter() - t0
        self.flat = self.rt.flat_scene()
        self.params = self.rt.render_params()
        self.ctx = pkg.Context(local)
        t0 = time.perf_counter()
        self.ctx.upload_scene(self.flat)
        self.upload_ms = (time.perf_counter() - t0) * 1e3
        self.info = self.ctx.scene_info()
        self.dev = self.ctx.device_info()
        self.stream = torch.cuda.ExternalStream(self.ctx.stream(), device=torch.device("cuda", local))
        p = self.params.copy()
        p.row_first, p.row_step, p.n_rows = pkg.rows_for_rank(H, rank, world)
        self.p = p
        self.band = torch.empty((max(p.n_rows, 1), W, 3), dtype=torch.int16, device="cuda")
        self.max_rows = (H + world - 1) // world
        self.peer_frame = False
        if world > 1:
            # the exchange stays on the device: per-row counts all-gathered on the context's stream ...
            self.mine_d = torch.zeros(self.max_rows, dtype=torch.int64, device="cuda")
            self.all_d = torch.zeros((world, self.max_rows), dtype=torch.int64, device="cuda")
            self.done = torch.zeros(1, dtype=torch.int32, device="cuda")
            self.side_stream = torch.cuda.Stream()
            self._map_frame()

    def _map_frame(self):
        """... and so does the gather: rank 0 owns the whole frame, the other ranks map it (CUDA IPC) and store their rows into
        it over NVLink at the end of rt580_render_finish_interleaved"""
        pkg, dist, torch = self.pkg, self.dist, self.torch
        handle = [self.ctx.frame_export(self.W, self.H) if self.rank == 0 else None]
        dist.broadcast_object_list(handle, src=0)
        ok = torch.ones(1, dtype=torch.int32, device="cuda")
        if self.rank != 0:
            try:
                self.ctx.frame_import(handle[0], self.W, self.H)
            except pkg.Rt580Error as e:
                print("rank %d: cannot map rank 0's frame (%s)" % (self.rank, e), file=sys.stderr)
                ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        self.peer_frame = bool(int(ok.item()))
        if not self.peer_frame:
            raise SystemExit("bench.py: CUDA IPC mapping of rank 0's frame failed; the multi-GPU path needs peer access")

    def release_frame(self):
        """importers release before the exporter (rt580.h)"""
        if self.world > 1 and self.peer_frame:
            if self.rank != 0:
                self.ctx.frame_release()
            self.dist.barrier()
            if self.rank == 0:
                self.ctx.frame_release()
            self.peer_frame = False

    def frame(self):
        """one step: this rank's rows; the frame ends up on rank 0 (device memory)"""
        pkg, dist, torch = self.pkg, self.dist, self.torch
        if self.world == 1:
            counts = self.ctx.render_begin(self.p)
            bases = pkg.row_bases_from_counts(self.H, 1, [counts])[0]
            _, st = self.ctx.render_finish(self.p, bases, device_ptr=self.band.data_ptr())
        else:
            self.ctx.render_begin(self.p, want_counts=False)
            self.ctx.row_counts_to_device(self.mine_d.data_ptr(), self.max_rows)
            with torch.cuda.stream(self.stream):
                dist.all_gather_into_tensor(self.all_d, self.mine_d)                # the one exchange of the LCG mode
            st = self.ctx.render_finish_interleaved(self.all_d.data_ptr(), self.world, self.rank, self.max_rows, device_ptr=None)
            # "every rank's rows have landed in rank 0's frame": a 4-byte all-reduce ordered after this rank's stores, on a
            # side stream - whoever consumes the frame on rank 0 waits for it, the next frame's structure pass does not
            self.side_stream.wait_stream(self.stream)
            with torch.cuda.stream(self.side_stream):
                dist.all_reduce(self.done)
        return st

    def sync(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def timed(self, steps, warmup, sampler=None):
        """-> dict: whole-job value over `steps` frames, max over ranks (device time, CUDA events on the context's stream)"""
        torch, dist = self.torch, self.dist
        for _ in range(warmup):
            self.frame()
        if sampler:
            sampler.start()
        self.sync()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record(self.stream)
        stats, launches, profs = [], 0, []
        for _ in range(steps):
            st = self.frame()
            stats.append(st)
            launches += st.kernel_launches
            profs.append(self.ctx.frame_profile())
        if self.world > 1:
            self.stream.wait_stream(self.side_stream)      # the timed region ends when the last frame is complete on rank 0
        ev1.record(self.stream)
        self.sync()
        wall_ms = (time.perf_counter() - t0) * 1e3
        if sampler:
            sampler.stop_flag = True
            sampler.join(timeout=2)
        step_ms = max(ev0.elapsed_time(ev1), 0.0) / steps
        rays_rank = float(np.mean([s.rays for s in stats]))
        cls_ms = np.mean([[pr.ms[k] for k in range(self.pkg.N_CLASSES)] for pr in profs], axis=0)
        cls_rays = np.mean([[float(pr.rays[k]) for k in range(self.pkg.N_CLASSES)] for pr in profs], axis=0)
        cls_launch = np.mean([[float(pr.launches[k]) for k in range(self.pkg.N_CLASSES)] for pr in profs], axis=0)
        t = torch.tensor([step_ms, wall_ms / steps] + list(cls_ms), dtype=torch.float64, device="cuda")
        r = torch.tensor([rays_rank] + list(cls_rays), dtype=torch.float64, device="cuda")
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(r, op=dist.ReduceOp.SUM)
        t, r = t.cpu().numpy(), r.cpu().numpy()
        st0 = stats[-1]
        return {"step_ms": float(t[0]), "wall_step_ms": float(t[1]), "rays_total": float(r[0]), "value": float(r[0]) / (float(t[0]) * 1e-3) / 1e6,
                "class_ms": t[2:].tolist(), "class_rays": r[1:].tolist(), "class_launches": cls_launch.tolist(), "launches": int(launches),
                "st0": st0, "stats": stats}

    def verify(self):
        """N > 1: the frame the ranks assembled on rank 0 against the same frame rendered by rank 0 alone (outside any timed region)"""
        if self.world == 1:
            return None
        self.frame()
        self.torch.cuda.synchronize()                     # includes the side stream's "all rows landed"
        self.sync()
        out = None
        if self.rank == 0:
            got = self.ctx.frame_read(self.W, self.H)
            want, _ = self.ctx.render(self.params)
            out = {"multi_gpu_frame_equals_single_gpu_frame": bool(np.array_equal(got, want)), "pixels": int(self.W * self.H)}
        self.sync()
        return out

    def e2e(self, steps):
        """the reference-facing calls with host buffers: upload (H2D + builds) + render (D2H), every step; and the render alone"""
        pkg, torch, dist = self.pkg, self.torch, self.dist
        W, H, p = self.W, self.H, self.p
        flat = self.flat
        # the class's upload with SetDeviceFlatten: meshes in object space + one model matrix per shape (rt580_upload_instanced_scene,
        # SURVEY 8f-2); --host-flatten uploads the flattened arrays instead (rt580_upload_scene)
        inst = None if self.host_flatten else self.rt.instanced_scene()
        if inst is None:
            h2d = int(flat.n_tris * (6 * 16 + 8) + flat.n_spheres * (16 + 8) + flat.n_materials * 32 + flat.n_lights * 44)
        else:
            mesh_tris = int(np.frombuffer((ctypes.c_int64 * (inst.n_meshes + 1)).from_address(inst.mesh_first), np.int64)[-1]) if inst.n_meshes else 0
            h2d = int(mesh_tris * 72 + (inst.n_meshes + 1) * 8 + inst.n_shapes * (64 + 4 + 4 + 32 + 16) + inst.n_lights * 44)
        d2h = int(p.n_rows * W * 6) if self.world == 1 else int(W * H * 6)
        host_out = pkg.HostArray((p.n_rows, W, 3), np.int16) if self.world == 1 else (pkg.HostArray((H, W, 3), np.int16) if self.rank == 0 else None)

        def one(upload):
            if upload:
                if inst is None:
                    self.ctx.upload_scene(flat)
                else:
                    self.ctx.upload_instanced_scene(inst)
            if self.world == 1:
                self.ctx.render(p, out=host_out.array)
            else:
                self.frame()
                if self.rank == 0:
                    torch.cuda.synchronize()
                    self.ctx.frame_read(W, H, out=host_out.array)

        res = {}
        for key, upload in (("cold", True), ("resident", False)):
            if upload and self.world > 1:
                # (a re-upload keeps the mapped frame: rt580_upload_scene does not touch it)
                pass
            one(upload)
            self.sync()
            t0 = time.perf_counter()
            for _ in range(steps):
                one(upload)
            self.sync()
            ms = (time.perf_counter() - t0) * 1e3 / steps
            te = torch.tensor([ms], dtype=torch.float64, device="cuda")
            if self.world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            res[key] = float(te[0])
        return res, h2d, d2h

    def close(self):
        self.release_frame()
        self.ctx.close()
        self.rt.close()


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

    rig = Rig(pkg, torch, dist, args.workload, args.farfield, rank, world, local)
    W, H = rig.W, rig.H
    sampler = ClockSampler(local) if rank == 0 else None
    m = rig.timed(args.steps, args.warmup, sampler)
    verify = rig.verify()
    # one more frame with the counting variants of the tree kernels (outside the timed region): node visits / leaf tests per ray
    rig.ctx.set_profiling(True)
    rig.frame()
    rig.sync()
    visits = rig.ctx.frame_profile()
    rig.ctx.set_profiling(False)
    vis = torch.tensor([float(visits.nodes_any), float(visits.leaves_any), float(visits.nodes_closest), float(visits.leaves_closest),
                        float(visits.rays[pkg.CLASS_NAMES.index("ao_tree")] + visits.rays[pkg.CLASS_NAMES.index("shadow_tree")]),
                        float(visits.rays[pkg.CLASS_NAMES.index("closest")])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vis, op=dist.ReduceOp.SUM)
    vis = vis.cpu().numpy()
    e2e_ms, h2d, d2h = rig.e2e(max(1, args.steps))

    records = {}
    info, dev, st0 = rig.info, rig.dev, m["st0"]
    head_scene = {"n_leaf": int(info.n_leaf), "bvh_max_depth": int(info.bvh_max_depth), "far_tmin": float(info.far_tmin)}
    rig.close()
    for spec in [w for w in args.records.split(",") if w and w != args.workload]:
        name, _, rec_far = spec.partition(":")
        rec_far = rec_far or args.farfield
        try:
            r2 = Rig(pkg, torch, dist, name, rec_far, rank, world, local)
            m2 = r2.timed(args.record_steps, 1)
            v2 = r2.verify()
            e2, _, _ = r2.e2e(1) if name.startswith("c4") else ({"cold": None, "resident": None}, 0, 0)
            s2 = m2["st0"]
            records[name] = {"value": m2["value"], "unit": UNIT, "ms_per_frame": m2["step_ms"], "rays_per_frame": m2["rays_total"],
                             "width": r2.W, "height": r2.H, "steps": args.record_steps, "warmup": 1, "n_gpus": world, "farfield": rec_far,
                             "farfield_note": None if rec_far == "exact" else (
                                 "far field OFF: with 10^7 triangles nine of ten rays that leave this open scene are 'hit' by a triangle 10^5..10^8 units "
                                 "away (float noise of cpp:392), three quarters of all rays of the exact frame then start out there; the exact frame "
                                 "(bit-identical to the reference, tests/test_gpu_parity.py at reduced resolution) costs ~40x this one, see DESIGN.md"),
                             "scene": r2.scene_text, "primitives_in_tree": int(r2.info.n_leaf),
                             "ao_rays_traversed_rank0": int(s2.ao_rays_traversed), "rays_ao_rank0": int(s2.rays_ao),
                             "far_scans_rank0": int(s2.far_scans), "linear_fallbacks_rank0": int(s2.linear_fallbacks),
                             "phases_ms_rank0": {"structure": s2.ms_structure, "order": s2.ms_order, "ao": s2.ms_ao, "resolve": s2.ms_resolve},
                             "class_ms_max_over_ranks": dict(zip(pkg.CLASS_NAMES, m2["class_ms"])),
                             "host_load_and_flatten_s": r2.load_s, "upload_and_build_ms": r2.upload_ms,
                             "e2e_ms": e2, "verify": v2}
            r2.close()
        except Exception as e:      # a record must not take the headline down with it
            records[name] = {"error": "%s: %s" % (type(e).__name__, e)}
            if world > 1:
                raise

    if rank == 0:
        peaks, peak_src = measured_peaks()
        sm_mhz = float(peaks.get("sm_max_mhz") or dev["sm_clock_mhz"])
        fp32_peak = dev["sm_count"] * 128 * 2 * sm_mhz * 1e6 / 1e12          # TFLOP/s, FMA counted as 2
        names = pkg.CLASS_NAMES
        cm = dict(zip(names, m["class_ms"])); cr = dict(zip(names, m["class_rays"])); cl = dict(zip(names, m["class_launches"]))
        # SURVEY 8d's model with the MEASURED visits: one 2-child node = 40 flop (64 B), one primitive test = 71 flop (64 B)
        any_rays = max(vis[4], 1.0); clo_rays = max(vis[5], 1.0)
        npr_any, lpr_any = vis[0] / any_rays, vis[1] / any_rays
        npr_clo, lpr_clo = vis[2] / clo_rays, vis[3] / clo_rays
        f_any, b_any = npr_any * 40 + lpr_any * 71, (npr_any + lpr_any) * 64
        f_clo = npr_clo * 40 + lpr_clo * 71
        classes = {}
        for n in names:
            classes[n] = {"rays": cr[n], "ms": cm[n], "launches": cl[n]}
        for n, f in (("ao_tree", f_any), ("shadow_tree", f_any), ("closest", f_clo)):
            if cm[n] > 0:
                classes[n]["flop_per_ray"] = f
                classes[n]["tflops"] = cr[n] * f / (cm[n] * 1e-3) / 1e12
                classes[n]["frac_fp32_peak"] = classes[n]["tflops"] / (fp32_peak * world)
        # the dominant tree kernel: k_anyhit over the AO rays (persistent LBVH any-hit traversal); CUDA events on its stream
        # around every launch of the timed region (rt580_frame_profile), max over ranks
        ao_ms, ao_rays, ao_l = cm["ao_tree"], cr["ao_tree"], max(cl["ao_tree"], 1.0)
        ncu = None
        if os.path.exists(NCU_CAPTURE):
            with open(NCU_CAPTURE) as f:
                ncu = json.load(f)
        traffic = None
        if ncu and ncu.get("dram_bytes_per_ray") is not None:
            traffic = ncu["dram_bytes_per_ray"] * ao_rays / world / ao_l
        ach = ao_rays * f_any / (ao_ms * 1e-3) / 1e12 if ao_ms > 0 else 0.0
        roofline = {"bound": "fp32", "kernel": "k_anyhit (persistent any-hit traversal of the LBVH) over the AO rays",
                    "achieved": ach, "peak": fp32_peak * world, "unit": "TFLOP/s", "frac": ach / (fp32_peak * world) if fp32_peak else None,
                    "peak_source": "%d SMs x 128 lanes x 2 x %.0f MHz (%s); MEASURED_PEAKS.json has no fp32 figure; the reference arithmetic is "
                                   "unfused (-fmad=false), so at most half of an FMA-counted peak is reachable" % (dev["sm_count"], sm_mhz, peak_src),
                    "flop_per_ray": f_any, "flop_model": "measured %.1f inner-node visits x 40 + %.2f leaf tests x 71 per ray (SURVEY 8d / App. D), "
                                                         "counted by the kernel's counting variant on one extra frame" % (npr_any, lpr_any),
                    "rays_per_launch": ao_rays / world / ao_l, "launches_per_step": ao_l, "kernel_ms": ao_ms / ao_l, "pass_ms": ao_ms,
                    "algorithmic_bytes_per_launch": b_any * ao_rays / world / ao_l,
                    "traffic": traffic,
                    "traffic_source": (ncu or {}).get("source"),
                    "ncu": (ncu or {}).get("counters"),
                    # what actually bounds the kernel: instruction issue (4 warp instructions per SM and clock), at the lane
                    # utilisation an incoherent any-hit traversal reaches; instructions per ray from the committed ncu capture
                    "issue": None if not ncu else {
                        "achieved": ncu["counters"]["warp_instructions_per_ray"] * ao_rays / (ao_ms * 1e-3) / 1e9 if ao_ms > 0 else 0.0,
                        "peak": dev["sm_count"] * 4 * sm_mhz * 1e6 / 1e9 * world, "unit": "G warp instructions/s",
                        "frac": (ncu["counters"]["warp_instructions_per_ray"] * ao_rays / (ao_ms * 1e-3) / 1e9) / (dev["sm_count"] * 4 * sm_mhz * 1e6 / 1e9 * world) if ao_ms > 0 else None,
                        "warp_instructions_per_ray": ncu["counters"]["warp_instructions_per_ray"],
                        "active_lanes_per_instruction": ncu["counters"]["active_lanes_per_instruction"]},
                    "hbm": {"achieved": ao_rays * b_any / (ao_ms * 1e-3) / 1e9 if ao_ms > 0 else 0.0, "peak": float(peaks["hbm_gbs"]) * world, "unit": "GB/s",
                            "bytes_per_ray": b_any,
                            "note": "algorithmic node + primitive bytes per ray x rays / kernel time: what the L1 / L2 hierarchy serves; "
                                    "the DRAM share is `traffic`"}}
        cores = os.cpu_count() or 1
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, rays, secs, kind = cpu_reference_sample(args.workload, max(2 * cores, 8), cores)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                   "sample": "%d random pixels of the %dx%d %s frame, %d rays in %.1f s, one reference Raytracer instance per thread (%d threads; "
                             "the reference itself is single-threaded)" % (max(2 * cores, 8), W, H, args.workload, rays, secs, cores)}
        e2e_value = m["rays_total"] / (e2e_ms["cold"] * 1e-3) / 1e6
        line = {"metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": m["step_ms"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": args.workload, "scene": rig.scene_text, "width": W, "height": H, "depth": DEPTH, "ao_spp": SPP,
                           "rng": "reference_lcg", "farfield": args.farfield, "far_scans": st0.far_scans, "linear_fallbacks": st0.linear_fallbacks,
                           "partition": "rows interleaved over %d rank(s)" % world,
                           "gather": "none" if world == 1 else "peer stores into rank 0's frame over NVLink (CUDA IPC), row counts all-gathered on the device",
                           "l2": "inputs larger than L2: nodes + records %.0f MB, far-field lists, frame data %.0f MB/step" % (
                               info.n_leaf * 144 / 1e6, st0.hit_nodes * 110 / 1e6)},
                "rays_per_frame": m["rays_total"], "ms_per_frame_4k": m["step_ms"], "wall_ms_per_step": m["wall_step_ms"],
                "phases_ms": {"structure": st0.ms_structure, "order": st0.ms_order, "ao": st0.ms_ao, "resolve": st0.ms_resolve},
                "ray_mix": {"primary": st0.rays_primary, "secondary": st0.rays_secondary, "shadow": st0.rays_shadow, "ao": st0.rays_ao,
                            "ao_rays_traversed": st0.ao_rays_traversed, "shadow_rays_traversed": st0.shadow_rays_traversed,
                            "note": "rank 0 share" if world > 1 else "whole frame"},
                "ray_classes": classes,
                "scene_info": head_scene,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms["cold"],
                        "includes": ("rt580_upload_scene (H2D of the flattened arrays" if HOST_FLATTEN else "rt580_upload_instanced_scene (H2D of meshes + model matrices, FlattenScene on the device") +
                                    " + LBVH + far-field grid build) + rt580_render (D2H int16 frame), every step",
                        "resident": {"value": m["rays_total"] / (e2e_ms["resident"] * 1e-3) / 1e6, "ms_per_step": e2e_ms["resident"],
                                     "includes": "rt580_render with the scene already uploaded (what a second Raytracer::Render call costs): D2H int16 frame"}},
                "verify": verify, "records": records, "gpu_launches": int(m["launches"]), "roofline": roofline, "cpu_baseline": cpu,
                "clocks": sampler.summary() if sampler else None}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--host-flatten", action="store_true", help="e2e uploads the host-flattened arrays (rt580_upload_scene) instead of meshes + matrices")
    ap.add_argument("--farfield", default="exact", choices=["exact", "off"], help="debug only")
    ap.add_argument("--workload", default=WORKLOAD, choices=sorted(WORKLOADS), help="headline workload")
    ap.add_argument("--records", default="c4_room,c5_open:off,c5_room", help="comma-separated workloads measured after the headline, "
                    "each optionally name:farfield ('' = none)")
    ap.add_argument("--record-steps", type=int, default=2)
    ap.add_argument("--width", type=int, default=0, help="debug only: override the frame width")
    ap.add_argument("--height", type=int, default=0, help="debug only: override the frame height")
    args = ap.parse_args()
    global HOST_FLATTEN
    HOST_FLATTEN = bool(args.host_flatten)
    if args.width and args.height:
        for k in list(WORKLOADS):
            WORKLOADS[k] = (args.width, args.height, WORKLOADS[k][2])
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
