"""580-raytracer_b200: Python host-side mirror of the reference's `Raytracer` class over the
C ABI of include/rt580.h (librt580.so: hand-written CUDA for sm_100a).

    rt = Raytracer(500, 500)                  # Raytracer.h:588
    rt.LoadSceneJSON("simpleSphereScene.json")# Raytracer.h:572
    rt.Render("output.ppm")                   # Raytracer.h:586

The directory name is not a Python identifier; load it with
`__graft_entry__.load_package()` (importlib by path).  There is no CPU path: every compute
call goes to the CUDA library and fails loudly when it or a GPU is missing.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librt580.so")

RT_SUCCESS, RT_FAILURE, RT_INVALID_ARG = 0, 1, 2
RNG_REFERENCE_LCG, RNG_COUNTER = 0, 1
TRAVERSAL_AUTO, TRAVERSAL_BVH, TRAVERSAL_BRUTE_FORCE = 0, 1, 2
FARFIELD_EXACT, FARFIELD_OFF = 0, 1
LIGHT_DIRECTIONAL, LIGHT_POINT, LIGHT_AMBIENT = 0, 1, 2

# every symbol include/rt580.h declares (tests/test_abi.py checks the header against this list)
EXPORTS = [
    "rt580_create", "rt580_destroy", "rt580_last_error", "rt580_device_info", "rt580_host_alloc", "rt580_host_free", "rt580_upload_scene", "rt580_build_ms",
    "rt580_scene_info_get", "rt580_get_stream",
    "rt580_render", "rt580_render_begin", "rt580_render_finish", "rt580_row_counts_to_device", "rt580_render_finish_interleaved",
    "rt580_frame_export", "rt580_frame_import", "rt580_frame_release", "rt580_frame_read", "rt580_frame_rgb8", "rt580_trace_closest", "rt580_trace_any", "rt580_trace_profile",
    "rt580_last_frame_ao_base", "rt580_hemisphere_stream", "rt580_powf", "rt580_set_profiling", "rt580_frame_profile",
    "rt580_raytracer_new", "rt580_raytracer_delete", "rt580_raytracer_set_assets_path", "rt580_raytracer_set_options", "rt580_raytracer_set_quiet",
    "rt580_raytracer_load_scene_json", "rt580_raytracer_render", "rt580_raytracer_flush_ppm",
    "rt580_raytracer_framebuffer", "rt580_raytracer_stats", "rt580_raytracer_flat_scene",
    "rt580_upload_instanced_scene", "rt580_flatten_instanced", "rt580_raytracer_set_device_flatten", "rt580_raytracer_instanced_scene",
    "rt580_raytracer_render_params", "rt580_raytracer_set_gpus", "rt580_raytracer_set_mesh_cache", "rt580_raytracer_mesh_cache_hits",
]


class Rt580Error(RuntimeError):
    def __init__(self, status, message):
        super().__init__("rt580 status %d: %s" % (status, message))
        self.status = status


class FlatScene(ctypes.Structure):
    _fields_ = [
        ("n_prims", ctypes.c_int64), ("n_tris", ctypes.c_int64),
        ("tri_v0", ctypes.c_void_p), ("tri_v1", ctypes.c_void_p), ("tri_v2", ctypes.c_void_p),
        ("tri_n0", ctypes.c_void_p), ("tri_n1", ctypes.c_void_p), ("tri_n2", ctypes.c_void_p),
        ("tri_prim", ctypes.c_void_p), ("tri_material", ctypes.c_void_p),
        ("n_spheres", ctypes.c_int64), ("sph_center_r", ctypes.c_void_p), ("sph_prim", ctypes.c_void_p),
        ("sph_material", ctypes.c_void_p),
        ("n_materials", ctypes.c_int32), ("materials", ctypes.c_void_p),
        ("n_lights", ctypes.c_int32), ("light_type", ctypes.c_void_p), ("light_f", ctypes.c_void_p),
        ("origin_hint", ctypes.c_float * 3),
    ]


class InstancedScene(ctypes.Structure):
    """rt580_instanced_scene: the scene before FlattenScene (meshes in object space, one model matrix per shape)."""
    _fields_ = [
        ("n_meshes", ctypes.c_int32), ("mesh_first", ctypes.c_void_p), ("mesh_tris", ctypes.c_void_p),
        ("n_shapes", ctypes.c_int32), ("shape_mesh", ctypes.c_void_p), ("shape_matrix", ctypes.c_void_p),
        ("shape_radius", ctypes.c_void_p), ("materials", ctypes.c_void_p),
        ("n_lights", ctypes.c_int32), ("light_type", ctypes.c_void_p), ("light_f", ctypes.c_void_p),
        ("origin_hint", ctypes.c_float * 3),
    ]


class RenderParams(ctypes.Structure):
    _fields_ = [
        ("width", ctypes.c_int32), ("height", ctypes.c_int32), ("fov_degrees", ctypes.c_float),
        ("camera_from", ctypes.c_float * 3), ("inv_view3x3", ctypes.c_float * 9),
        ("depth", ctypes.c_int32), ("ao_spp", ctypes.c_int32), ("rng_mode", ctypes.c_int32),
        ("traversal", ctypes.c_int32), ("row_first", ctypes.c_int32), ("row_step", ctypes.c_int32),
        ("n_rows", ctypes.c_int32), ("farfield", ctypes.c_int32),
    ]

    def copy(self):
        c = RenderParams()
        ctypes.memmove(ctypes.addressof(c), ctypes.addressof(self), ctypes.sizeof(RenderParams))
        return c


class SceneInfo(ctypes.Structure):
    _fields_ = [
        ("n_leaf", ctypes.c_int64), ("n_dropped", ctypes.c_int64), ("n_always", ctypes.c_int64),
        ("far_tmin", ctypes.c_float), ("pad", ctypes.c_float), ("extent", ctypes.c_float),
        ("build_ms", ctypes.c_float), ("bvh_max_depth", ctypes.c_uint32), ("reserved", ctypes.c_uint32),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


class Stats(ctypes.Structure):
    _fields_ = [
        ("rays_primary", ctypes.c_uint64), ("rays_secondary", ctypes.c_uint64), ("rays_shadow", ctypes.c_uint64),
        ("rays_ao", ctypes.c_uint64), ("hit_nodes", ctypes.c_uint64), ("ao_calls", ctypes.c_uint64),
        ("ms_structure", ctypes.c_float), ("ms_order", ctypes.c_float), ("ms_ao", ctypes.c_float),
        ("ms_resolve", ctypes.c_float), ("ms_total", ctypes.c_float), ("ms_ao_kernel", ctypes.c_float),
        ("kernel_launches", ctypes.c_uint32), ("bvh_max_depth", ctypes.c_uint32),
        ("far_scans", ctypes.c_uint32), ("linear_fallbacks", ctypes.c_uint32),
        ("ao_rays_traversed", ctypes.c_uint64), ("shadow_rays_traversed", ctypes.c_uint64),
    ]

    @property
    def rays(self):
        return self.rays_primary + self.rays_secondary + self.rays_shadow + self.rays_ao

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if not k.startswith("reserved")}


N_CLASSES = 10
CLASS_NAMES = ["primary", "closest", "shadow_gen", "shadow_tree", "ao_gen", "ao_tree", "far_any", "far_closest", "order", "resolve"]


class Profile(ctypes.Structure):
    _fields_ = [
        ("ms", ctypes.c_float * N_CLASSES), ("rays", ctypes.c_uint64 * N_CLASSES), ("launches", ctypes.c_uint32 * N_CLASSES),
        ("reserved", ctypes.c_uint32),
        ("nodes_any", ctypes.c_uint64), ("leaves_any", ctypes.c_uint64), ("nodes_closest", ctypes.c_uint64), ("leaves_closest", ctypes.c_uint64),
    ]

    def as_dict(self):
        d = {n: {"ms": float(self.ms[i]), "rays": int(self.rays[i]), "launches": int(self.launches[i])} for i, n in enumerate(CLASS_NAMES)}
        d["visits"] = {"nodes_any": int(self.nodes_any), "leaves_any": int(self.leaves_any),
                       "nodes_closest": int(self.nodes_closest), "leaves_closest": int(self.leaves_closest)}
        return d


def build(verbose=False):
    """Compile librt580.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", HERE, "librt580.so"], stdout=out)
    return LIB_PATH


_lib = None


def lib():
    """The CUDA library.  No fallback: a missing .so is an error, not a detour."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Rt580Error(RT_FAILURE, "librt580.so is not built (run __graft_entry__.build()); there is no CPU path")
        L = ctypes.CDLL(LIB_PATH)
        vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
        L.rt580_last_error.restype = ctypes.c_char_p
        L.rt580_create.argtypes = [i32, ctypes.POINTER(vp)]
        L.rt580_destroy.argtypes = [vp]
        L.rt580_destroy.restype = None
        L.rt580_device_info.argtypes = [vp, vp, vp, vp]
        L.rt580_get_stream.argtypes = [vp, ctypes.POINTER(vp)]
        L.rt580_host_alloc.restype = vp
        L.rt580_host_alloc.argtypes = [ctypes.c_uint64]
        L.rt580_host_free.restype = None
        L.rt580_host_free.argtypes = [vp]
        L.rt580_upload_scene.argtypes = [vp, ctypes.POINTER(FlatScene)]
        L.rt580_build_ms.argtypes = [vp, vp]
        L.rt580_scene_info_get.argtypes = [vp, ctypes.POINTER(SceneInfo)]
        L.rt580_render.argtypes = [vp, ctypes.POINTER(RenderParams), vp, ctypes.POINTER(Stats)]
        L.rt580_render_begin.argtypes = [vp, ctypes.POINTER(RenderParams), vp]
        L.rt580_render_finish.argtypes = [vp, vp, vp, i32, ctypes.POINTER(Stats)]
        L.rt580_row_counts_to_device.argtypes = [vp, vp, i32]
        L.rt580_render_finish_interleaved.argtypes = [vp, vp, i32, i32, i32, vp, i32, ctypes.POINTER(Stats)]
        L.rt580_frame_export.argtypes = [vp, i32, i32, vp]
        L.rt580_frame_import.argtypes = [vp, vp, i32, i32]
        L.rt580_frame_release.argtypes = [vp]
        L.rt580_frame_read.argtypes = [vp, vp]
        L.rt580_frame_rgb8.argtypes = [vp, vp, vp, i32]
        L.rt580_trace_closest.argtypes = [vp, i64, vp, vp, i32, vp, vp]
        L.rt580_trace_any.argtypes = [vp, i64, vp, vp, vp, i32, vp]
        L.rt580_trace_profile.argtypes = [vp, i64, vp, vp, vp, vp]
        L.rt580_last_frame_ao_base.argtypes = [vp, vp]
        L.rt580_hemisphere_stream.argtypes = [vp, vp, ctypes.c_uint64, i32, vp]
        L.rt580_powf.argtypes = [vp, i64, vp, vp, vp]
        L.rt580_set_profiling.argtypes = [vp, i32]
        L.rt580_frame_profile.argtypes = [vp, ctypes.POINTER(Profile)]
        L.rt580_raytracer_new.restype = vp
        L.rt580_raytracer_new.argtypes = [i32, i32]
        L.rt580_raytracer_delete.argtypes = [vp]
        L.rt580_raytracer_delete.restype = None
        L.rt580_raytracer_set_assets_path.argtypes = [vp, ctypes.c_char_p]
        L.rt580_raytracer_set_options.argtypes = [vp, i32, i32, i32, i32, i32, i32]
        L.rt580_raytracer_set_quiet.argtypes = [vp, i32]
        L.rt580_raytracer_load_scene_json.argtypes = [vp, ctypes.c_char_p]
        L.rt580_raytracer_set_gpus.argtypes = [vp, i32]
        L.rt580_raytracer_set_device_flatten.argtypes = [vp, i32]
        L.rt580_raytracer_instanced_scene.argtypes = [vp, ctypes.POINTER(InstancedScene)]
        L.rt580_upload_instanced_scene.argtypes = [vp, ctypes.POINTER(InstancedScene)]
        L.rt580_flatten_instanced.argtypes = [vp, ctypes.POINTER(InstancedScene)] + [vp] * 11
        L.rt580_raytracer_set_mesh_cache.argtypes = [vp, ctypes.c_char_p]
        L.rt580_raytracer_mesh_cache_hits.argtypes = [vp]
        L.rt580_raytracer_render.argtypes = [vp, ctypes.c_char_p]
        L.rt580_raytracer_flush_ppm.argtypes = [vp, ctypes.c_char_p]
        L.rt580_raytracer_framebuffer.restype = vp
        L.rt580_raytracer_framebuffer.argtypes = [vp]
        L.rt580_raytracer_stats.argtypes = [vp, ctypes.POINTER(Stats)]
        L.rt580_raytracer_flat_scene.argtypes = [vp, ctypes.POINTER(FlatScene)]
        L.rt580_raytracer_render_params.argtypes = [vp, ctypes.POINTER(RenderParams)]
        _lib = L
    return _lib


def _check(status):
    if status != RT_SUCCESS:
        raise Rt580Error(status, lib().rt580_last_error().decode(errors="replace"))


class Context:
    """rt580_context: one GPU, one uploaded scene, frames rendered through the C ABI."""

    def __init__(self, device=0):
        h = ctypes.c_void_p()
        _check(lib().rt580_create(device, ctypes.byref(h)))
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            lib().rt580_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def device_info(self):
        sm, mhz, mem = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_uint64()
        _check(lib().rt580_device_info(self._h, ctypes.addressof(sm), ctypes.addressof(mhz), ctypes.addressof(mem)))
        return {"sm_count": sm.value, "sm_clock_mhz": mhz.value, "hbm_bytes": mem.value}

    def stream(self):
        s = ctypes.c_void_p()
        _check(lib().rt580_get_stream(self._h, ctypes.byref(s)))
        return s.value or 0

    def upload_scene(self, flat: FlatScene):
        _check(lib().rt580_upload_scene(self._h, ctypes.byref(flat)))

    def upload_instanced_scene(self, inst: InstancedScene):
        """The same scene from meshes + one model matrix per shape: FlattenScene runs on the device (rt580.h)."""
        _check(lib().rt580_upload_instanced_scene(self._h, ctypes.byref(inst)))

    def flatten_instanced(self, inst: InstancedScene, n_tris, n_spheres):
        """The device flatten alone, copied back: arrays named as in flat_scene_arrays (a test hook)."""
        f4 = lambda n: np.zeros((n, 4), np.float32)
        i1 = lambda n: np.zeros(n, np.int32)
        out = {"tri_v0": f4(n_tris), "tri_v1": f4(n_tris), "tri_v2": f4(n_tris), "tri_n0": f4(n_tris), "tri_n1": f4(n_tris), "tri_n2": f4(n_tris),
               "tri_prim": i1(n_tris), "tri_material": i1(n_tris), "sph_center_r": f4(n_spheres), "sph_prim": i1(n_spheres), "sph_material": i1(n_spheres)}
        order = ["tri_v0", "tri_v1", "tri_v2", "tri_n0", "tri_n1", "tri_n2", "tri_prim", "tri_material", "sph_center_r", "sph_prim", "sph_material"]
        _check(lib().rt580_flatten_instanced(self._h, ctypes.byref(inst), *[out[k].ctypes.data if out[k].size else None for k in order]))
        return out

    def scene_info(self):
        info = SceneInfo()
        _check(lib().rt580_scene_info_get(self._h, ctypes.byref(info)))
        return info

    def build_ms(self):
        ms = ctypes.c_float()
        _check(lib().rt580_build_ms(self._h, ctypes.addressof(ms)))
        return ms.value

    def render(self, params: RenderParams, out=None):
        """Whole frame (or the rows in params) -> host int16 [n_rows][width][3]; `out`: a preallocated
        array to fill (e.g. over page-locked memory, host_array())."""
        n_rows = max(params.n_rows, 0) or (params.height if params.n_rows == 0 else 0)
        fb = np.empty((n_rows, params.width, 3), np.int16) if out is None else out
        assert fb.dtype == np.int16 and fb.size == n_rows * params.width * 3 and fb.flags["C_CONTIGUOUS"]
        st = Stats()
        _check(lib().rt580_render(self._h, ctypes.byref(params), fb.ctypes.data, ctypes.byref(st)))
        return fb, st

    def render_begin(self, params: RenderParams, want_counts=True):
        """Structure pass.  Returns the per-row hit-node counts, or None with want_counts=False
        (they then stay on the device for row_counts_to_device)."""
        n_rows = max(params.n_rows, 0) or (params.height if params.n_rows == 0 else 0)
        if not want_counts:
            _check(lib().rt580_render_begin(self._h, ctypes.byref(params), None))
            return None
        rows = np.zeros(max(n_rows, 1), np.uint64)
        _check(lib().rt580_render_begin(self._h, ctypes.byref(params), rows.ctypes.data))
        return rows[:n_rows]

    def row_counts_to_device(self, dst_device_ptr, max_rows):
        _check(lib().rt580_row_counts_to_device(self._h, dst_device_ptr, max_rows))

    def render_finish_interleaved(self, all_counts_device_ptr, world, rank, max_rows, device_ptr=None):
        """Finish with the all-gathered row counts on the device (no host round trip)."""
        st = Stats()
        _check(lib().rt580_render_finish_interleaved(self._h, all_counts_device_ptr, world, rank, max_rows, device_ptr,
                                                      1 if device_ptr else 0, ctypes.byref(st)))
        return st

    def frame_rgb8(self, lut256, n_rows, width):
        """The last finished frame as the PPM's 8-bit RGB body (gamma table applied on the device)."""
        lut = np.ascontiguousarray(lut256, np.uint8)
        assert lut.size == 256
        out = np.empty((n_rows, width, 3), np.uint8)
        _check(lib().rt580_frame_rgb8(self._h, lut.ctypes.data, out.ctypes.data, 0))
        return out

    def frame_export(self, width, height):
        """Rank 0: allocate the whole frame; returns the 64-byte CUDA IPC handle for the other ranks."""
        h = ctypes.create_string_buffer(64)
        _check(lib().rt580_frame_export(self._h, width, height, h))
        return h.raw

    def frame_import(self, handle, width, height):
        buf = ctypes.create_string_buffer(bytes(handle), 64)
        _check(lib().rt580_frame_import(self._h, buf, width, height))

    def frame_release(self):
        _check(lib().rt580_frame_release(self._h))

    def frame_read(self, width, height, out=None):
        fb = np.empty((height, width, 3), np.int16) if out is None else out
        assert fb.dtype == np.int16 and fb.size == height * width * 3 and fb.flags["C_CONTIGUOUS"]
        _check(lib().rt580_frame_read(self._h, fb.ctypes.data))
        return fb

    def render_finish(self, params: RenderParams, row_ao_base=None, out=None, device_ptr=None):
        """out: host int16 array, or device_ptr: raw CUDA pointer (e.g. torch tensor data_ptr())."""
        n_rows = max(params.n_rows, 0) or (params.height if params.n_rows == 0 else 0)
        st = Stats()
        base_ptr = None
        if row_ao_base is not None:
            row_ao_base = np.ascontiguousarray(row_ao_base, np.uint64)
            assert row_ao_base.size == n_rows
            base_ptr = row_ao_base.ctypes.data
        if device_ptr is not None:
            _check(lib().rt580_render_finish(self._h, base_ptr, device_ptr, 1, ctypes.byref(st)))
            return None, st
        if out is None:
            out = np.empty((n_rows, params.width, 3), np.int16)
        _check(lib().rt580_render_finish(self._h, base_ptr, out.ctypes.data, 0, ctypes.byref(st)))
        return out, st

    def trace_closest(self, org, dirs, traversal=TRAVERSAL_AUTO):
        org = np.ascontiguousarray(org, np.float32).reshape(-1, 3)
        dirs = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        n = org.shape[0]
        prim = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32)
        _check(lib().rt580_trace_closest(self._h, n, org.ctypes.data, dirs.ctypes.data, traversal, prim.ctypes.data, t.ctypes.data))
        return prim, t

    def trace_any(self, org, dirs, tmax, traversal=TRAVERSAL_AUTO):
        org = np.ascontiguousarray(org, np.float32).reshape(-1, 3)
        dirs = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        tmax = np.ascontiguousarray(tmax, np.float32)
        n = org.shape[0]
        hit = np.zeros(n, np.uint8)
        _check(lib().rt580_trace_any(self._h, n, org.ctypes.data, dirs.ctypes.data, tmax.ctypes.data, traversal, hit.ctypes.data))
        return hit

    def trace_profile(self, org, dirs, tmax=None):
        org = np.ascontiguousarray(org, np.float32).reshape(-1, 3)
        dirs = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        n = org.shape[0]
        cnt = np.zeros((n, 4), np.uint32)
        tptr = None
        if tmax is not None:
            tmax = np.ascontiguousarray(tmax, np.float32)
            tptr = tmax.ctypes.data
        _check(lib().rt580_trace_profile(self._h, n, org.ctypes.data, dirs.ctypes.data, tptr, cnt.ctypes.data))
        return cnt

    def set_profiling(self, count_visits):
        _check(lib().rt580_set_profiling(self._h, 1 if count_visits else 0))

    def frame_profile(self):
        """Per-class device times and ray counts of the last finished frame (rt580.h RT580_CLASS_*)."""
        pr = Profile()
        _check(lib().rt580_frame_profile(self._h, ctypes.byref(pr)))
        return pr

    def last_frame_ao_base(self, n_pixels):
        out = np.zeros(n_pixels, np.uint64)
        _check(lib().rt580_last_frame_ao_base(self._h, out.ctypes.data))
        return out

    def hemisphere_stream(self, normal, step, n):
        normal = np.ascontiguousarray(normal, np.float32)
        out = np.zeros((n, 3), np.float32)
        _check(lib().rt580_hemisphere_stream(self._h, normal.ctypes.data, step, n, out.ctypes.data))
        return out

    def powf(self, x, y):
        x = np.ascontiguousarray(x, np.float32)
        y = np.ascontiguousarray(y, np.float32)
        out = np.zeros(x.shape, np.float32)
        _check(lib().rt580_powf(self._h, x.size, x.ctypes.data, y.ctypes.data, out.ctypes.data))
        return out


class Raytracer:
    """The reference's class, method for method (Raytracer.h:557-588), over the C ABI."""

    def __init__(self, width, height, quiet=True):
        self._h = lib().rt580_raytracer_new(width, height)
        if not self._h:
            raise Rt580Error(RT_FAILURE, "rt580_raytracer_new failed")
        lib().rt580_raytracer_set_quiet(self._h, 1 if quiet else 0)
        self.width, self.height = width, height
        self._opts = dict(depth=4, ao_spp=128, rng_mode=RNG_REFERENCE_LCG, traversal=TRAVERSAL_AUTO, device=0,
                          farfield=FARFIELD_EXACT)

    def close(self):
        if getattr(self, "_h", None):
            lib().rt580_raytracer_delete(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def SetAssetsPath(self, directory):
        return lib().rt580_raytracer_set_assets_path(self._h, directory.encode())

    def SetOptions(self, **kw):
        self._opts.update(kw)
        o = self._opts
        return lib().rt580_raytracer_set_options(self._h, o["depth"], o["ao_spp"], o["rng_mode"], o["traversal"], o["device"], o["farfield"])

    def SetGpus(self, n_gpus):
        return lib().rt580_raytracer_set_gpus(self._h, n_gpus)

    def SetDeviceFlatten(self, on=True):
        return lib().rt580_raytracer_set_device_flatten(self._h, 1 if on else 0)

    def instanced_scene(self):
        inst = InstancedScene()
        st = lib().rt580_raytracer_instanced_scene(self._h, ctypes.byref(inst))
        if st != RT_SUCCESS:
            raise Rt580Error(st, "no scene loaded")
        return inst   # borrows the Raytracer's buffers: keep `self` alive while it is used

    def SetMeshCacheDir(self, directory):
        return lib().rt580_raytracer_set_mesh_cache(self._h, (directory or "").encode())

    def MeshCacheHits(self):
        return lib().rt580_raytracer_mesh_cache_hits(self._h)

    def LoadSceneJSON(self, scene_path):
        return lib().rt580_raytracer_load_scene_json(self._h, scene_path.encode())

    def Render(self, output_name=""):
        return lib().rt580_raytracer_render(self._h, (output_name or "").encode())

    def FlushFrameBufferToPPM(self, output_name):
        return lib().rt580_raytracer_flush_ppm(self._h, output_name.encode())

    def frame_buffer(self):
        p = lib().rt580_raytracer_framebuffer(self._h)
        buf = (ctypes.c_int16 * (self.width * self.height * 3)).from_address(p)
        return np.frombuffer(buf, np.int16).reshape(self.height, self.width, 3).copy()

    def stats(self):
        st = Stats()
        lib().rt580_raytracer_stats(self._h, ctypes.byref(st))
        return st

    def flat_scene(self):
        fs = FlatScene()
        st = lib().rt580_raytracer_flat_scene(self._h, ctypes.byref(fs))
        if st != RT_SUCCESS:
            raise Rt580Error(st, "no scene loaded")
        return fs     # borrows the Raytracer's buffers: keep `self` alive while it is used

    def render_params(self):
        rp = RenderParams()
        st = lib().rt580_raytracer_render_params(self._h, ctypes.byref(rp))
        if st != RT_SUCCESS:
            raise Rt580Error(st, "no scene loaded")
        return rp


def flat_scene_arrays(fs: FlatScene):
    """numpy views of a FlatScene (for tests)."""
    def arr(ptr, n, width, dt):
        if not ptr or n == 0:
            return np.zeros((0, width) if width > 1 else (0,), dt)
        ct = {np.float32: ctypes.c_float, np.int32: ctypes.c_int32}[dt]
        buf = (ct * (n * width)).from_address(ptr)
        a = np.frombuffer(buf, dt)
        return a.reshape(n, width) if width > 1 else a
    nt, ns = fs.n_tris, fs.n_spheres
    return {
        "tri_v0": arr(fs.tri_v0, nt, 4, np.float32), "tri_v1": arr(fs.tri_v1, nt, 4, np.float32),
        "tri_v2": arr(fs.tri_v2, nt, 4, np.float32), "tri_n0": arr(fs.tri_n0, nt, 4, np.float32),
        "tri_n1": arr(fs.tri_n1, nt, 4, np.float32), "tri_n2": arr(fs.tri_n2, nt, 4, np.float32),
        "tri_prim": arr(fs.tri_prim, nt, 1, np.int32), "tri_material": arr(fs.tri_material, nt, 1, np.int32),
        "sph_center_r": arr(fs.sph_center_r, ns, 4, np.float32), "sph_prim": arr(fs.sph_prim, ns, 1, np.int32),
        "sph_material": arr(fs.sph_material, ns, 1, np.int32),
        "materials": arr(fs.materials, fs.n_materials, 8, np.float32),
        "light_type": arr(fs.light_type, fs.n_lights, 1, np.int32),
        "light_f": arr(fs.light_f, fs.n_lights, 10, np.float32),
    }


class HostArray:
    """numpy array over rt580_host_alloc memory (page-locked when a GPU is present)."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * self.dtype.itemsize
        self._p = lib().rt580_host_alloc(max(n, 1))
        if not self._p:
            raise MemoryError("rt580_host_alloc(%d)" % n)
        buf = (ctypes.c_char * max(n, 1)).from_address(self._p)
        self.array = np.frombuffer(buf, self.dtype, int(np.prod(shape))).reshape(shape)

    def close(self):
        if getattr(self, "_p", None):
            self.array = None
            lib().rt580_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- multi-GPU host logic (pure index arithmetic; the collectives live in bench.py / callers) ----
def rows_for_rank(height, rank, world):
    """Interleaved row partition: rank r renders rows r, r+world, ... (load balance: in the
    mesh scenes most rows are background).  -> (row_first, row_step, n_rows)"""
    if rank >= height:
        return (0, world, -1)            # n_rows < 0: this rank renders no row (n_rows == 0 would mean the whole frame, rt580.h)
    return (rank, world, (height - rank + world - 1) // world)


def row_bases_from_counts(height, world, counts_per_rank):
    """counts_per_rank[r][k] = hit nodes of rank r's k-th row.  Returns, per rank, the number
    of hit nodes in all rows that precede each of its rows in scanline order (the AO-stream
    prefix of SURVEY Appendix C) - the `row_ao_base` argument of rt580_render_finish."""
    per_row = np.zeros(height, np.uint64)
    for r in range(world):
        first, step, n = rows_for_rank(height, r, world)
        n = max(n, 0)
        per_row[first:first + n * step:step] = np.asarray(counts_per_rank[r][:n], np.uint64)
    excl = np.concatenate([[np.uint64(0)], np.cumsum(per_row, dtype=np.uint64)[:-1]]) if height else per_row
    out = []
    for r in range(world):
        first, step, n = rows_for_rank(height, r, world)
        n = max(n, 0)
        out.append(np.ascontiguousarray(excl[first:first + n * step:step], np.uint64))
    return out


def interleave_rows(height, width, world, bands):
    """bands[r]: [n_rows_r, width, 3] -> full [height, width, 3] frame."""
    fb = np.empty((height, width, 3), np.int16)
    for r in range(world):
        first, step, n = rows_for_rank(height, r, world)
        if n > 0:
            fb[first:first + n * step:step] = bands[r][:n]
    return fb
