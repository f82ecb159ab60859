"""TEST INFRASTRUCTURE ONLY - Python bindings of the two oracle tiers.

  T0  oracle/_ref/libref580.so   the reference's own Raytracer.cpp/.h compiled from
                                 /root/reference by oracle/build_ref.sh (ground truth)
  T1  oracle/liboracle580.so     this repo's CPU restatement (oracle580.c), pinned
                                 bit-for-bit against T0 by tests/test_oracle_vs_ref.py

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.  The product (580-raytracer_b200/) never does.
"""
import ctypes
import json
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
T1_LIB = os.path.join(HERE, "liboracle580.so")
T0_LIB = os.path.join(HERE, "_ref", "libref580.so")
REFERENCE_ASSETS = "/root/reference/580 Raytracer/Assets/"

RT_SUCCESS, RT_FAILURE, RT_INVALID_ARG = 0, 1, 2
LIGHT_DIRECTIONAL, LIGHT_POINT, LIGHT_AMBIENT = 0, 1, 2   # h:520-524


def build(ref=True):
    """Compile the T1 oracle and (when /root/reference exists) the T0 reference library."""
    subprocess.check_call(["make", "-s", "-C", HERE, "liboracle580.so"])
    if ref:
        subprocess.check_call(["bash", os.path.join(HERE, "build_ref.sh")])


# --------------------------------------------------------------------------------------
# Scene description in the reference loader's in-memory form (cpp:589-779), parsed with
# Python's json module: numbers go double -> float32, the same two roundings nlohmann
# json + the float assignment perform.
# --------------------------------------------------------------------------------------
class SceneArrays:
    def __init__(self):
        self.shape_mesh = []
        self.shape_material = []
        self.shape_srt = []
        self.mesh_names = []
        self.mesh_type = []
        self.mesh_tri_begin = [0]
        self.mesh_radius = []
        self.tri_pos = []
        self.tri_nrm = []
        self.light_type = []
        self.light_f = []
        self.cam_from = [0.0, 0.0, 0.0]
        self.cam_to = [0.0, 0.0, 0.0]

    def finalize(self):
        f32, i32 = np.float32, np.int32
        self.shape_mesh = np.ascontiguousarray(self.shape_mesh, i32)
        self.shape_material = np.ascontiguousarray(self.shape_material, f32).reshape(-1, 8)
        self.shape_srt = np.ascontiguousarray(self.shape_srt, f32).reshape(-1, 9)
        self.mesh_type = np.ascontiguousarray(self.mesh_type, i32)
        self.mesh_tri_begin = np.ascontiguousarray(self.mesh_tri_begin, i32)
        self.mesh_radius = np.ascontiguousarray(self.mesh_radius, f32)
        self.tri_pos = (np.concatenate(self.tri_pos) if self.tri_pos else np.zeros((0, 9))).astype(f32).reshape(-1, 9)
        self.tri_nrm = (np.concatenate(self.tri_nrm) if self.tri_nrm else np.zeros((0, 9))).astype(f32).reshape(-1, 9)
        self.tri_pos = np.ascontiguousarray(self.tri_pos)
        self.tri_nrm = np.ascontiguousarray(self.tri_nrm)
        self.light_type = np.ascontiguousarray(self.light_type, i32)
        self.light_f = np.ascontiguousarray(self.light_f, f32).reshape(-1, 10)
        self.cam_from = np.ascontiguousarray(self.cam_from, f32)
        self.cam_to = np.ascontiguousarray(self.cam_to, f32)
        return self

    @property
    def n_prims(self):
        n = 0
        for m in self.shape_mesh:
            n += (self.mesh_tri_begin[m + 1] - self.mesh_tri_begin[m]) if self.mesh_type[m] == 0 else 1
        return int(n)


def _load_mesh(sa, assets_dir, name, cache):
    """cpp:589-643: mesh type comes from data[0].type only (Q26); cached by name."""
    if name in cache:
        return cache[name]
    with open(os.path.join(assets_dir, name + ".json")) as f:
        data = json.load(f)["data"]
    kind = data[0]["type"]
    idx = len(sa.mesh_type)
    if kind == "polygon":
        pos = np.array([[it["v%d" % k]["v"] for k in range(3)] for it in data], dtype=np.float64).reshape(-1, 9)
        nrm = np.array([[it["v%d" % k]["n"] for k in range(3)] for it in data], dtype=np.float64).reshape(-1, 9)
        sa.mesh_type.append(0)
        sa.mesh_radius.append(0.0)
        sa.tri_pos.append(pos)
        sa.tri_nrm.append(nrm)
        sa.mesh_tri_begin.append(sa.mesh_tri_begin[-1] + len(data))
    elif kind == "sphere":
        sa.mesh_type.append(1)
        sa.mesh_radius.append(float(data[-1]["radius"]))   # loop keeps the last item (cpp:631-638)
        sa.mesh_tri_begin.append(sa.mesh_tri_begin[-1])
    else:
        raise ValueError("unsupported mesh type %r (Q26)" % kind)
    sa.mesh_names.append(name)
    cache[name] = idx
    return idx


def load_scene_json(assets_dir, scene_name):
    """cpp:645-779 restated: transforms are 'last value wins' regardless of order."""
    with open(os.path.join(assets_dir, scene_name)) as f:
        sc = json.load(f)["scene"]
    sa = SceneArrays()
    cache = {}
    f32 = np.float32
    for sh in sc.get("shapes", []):
        m = sh["material"]
        S, R, T = [1.0, 1.0, 1.0], [0.0, 0.0, 0.0], [0.0, 0.0, 0.0]
        for tr in sh["transforms"]:
            if "Rx" in tr: R[0] = tr["Rx"]
            if "Ry" in tr: R[1] = tr["Ry"]
            if "Rz" in tr: R[2] = tr["Rz"]
            if isinstance(tr.get("S"), list): S = list(tr["S"][:3])
            if isinstance(tr.get("T"), list): T = list(tr["T"][:3])
        sa.shape_material.append(list(m["Cs"][:3]) + [m["Ka"], m["Kd"], m["Ks"], m["Kt"], m["n"]])
        sa.shape_srt.append(S + R + T)
        sa.shape_mesh.append(_load_mesh(sa, assets_dir, sh["geometry"], cache))
    if "camera" in sc:
        sa.cam_from = list(sc["camera"]["from"][:3])
        sa.cam_to = list(sc["camera"]["to"][:3])
    for li in sc.get("lights", []):
        color = list(li["color"][:3])
        pos, dirv = [0.0, 0.0, 0.0], [0.0, 0.0, 0.0]
        kind = li["type"]
        if kind == "directional":
            d = np.array(li["to"][:3], f32) - np.array(li["from"][:3], f32)      # cpp:757
            length = np.sqrt(f32(f32(d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]))  # h:109-116
            if length > 0:
                d = d / length
            dirv = [float(x) for x in d]
            t = LIGHT_DIRECTIONAL
        elif kind == "ambient":
            t = LIGHT_AMBIENT
        elif kind == "point":
            t = LIGHT_POINT
            pos = list(li["position"][:3])
        else:
            raise ValueError("unknown light type %r" % kind)
        sa.light_type.append(t)
        sa.light_f.append(color + [li["intensity"]] + pos + dirv)
    return sa.finalize()


# --------------------------------------------------------------------------------------
# T1
# --------------------------------------------------------------------------------------
class _OrcScene(ctypes.Structure):
    _fields_ = [
        ("n_shapes", ctypes.c_int32), ("shape_mesh", ctypes.c_void_p), ("shape_material", ctypes.c_void_p),
        ("shape_srt", ctypes.c_void_p), ("n_meshes", ctypes.c_int32), ("mesh_type", ctypes.c_void_p),
        ("mesh_tri_begin", ctypes.c_void_p), ("mesh_radius", ctypes.c_void_p), ("tri_pos", ctypes.c_void_p),
        ("tri_nrm", ctypes.c_void_p), ("n_lights", ctypes.c_int32), ("light_type", ctypes.c_void_p),
        ("light_f", ctypes.c_void_p), ("cam_from", ctypes.c_float * 3), ("cam_to", ctypes.c_float * 3),
    ]


_t1 = None


def t1_lib():
    global _t1
    if _t1 is None:
        if not os.path.exists(T1_LIB):
            build(ref=False)
        lib = ctypes.CDLL(T1_LIB)
        lib.orc580_prepare.restype = ctypes.c_void_p
        lib.orc580_prepare.argtypes = [ctypes.POINTER(_OrcScene)]
        lib.orc580_free.argtypes = [ctypes.c_void_p]
        lib.orc580_num_prims.restype = ctypes.c_int64
        lib.orc580_num_prims.argtypes = [ctypes.c_void_p]
        lib.orc580_model_matrix.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        lib.orc580_fresnel.argtypes = [ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_void_p]
        lib.orc580_hemisphere_stream.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p]
        lib.orc580_lcg_state.restype = ctypes.c_uint32
        lib.orc580_lcg_state.argtypes = [ctypes.c_uint64]
        lib.orc580_primary_ray.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 4 + [ctypes.c_void_p] * 2
        lib.orc580_render.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 4 + [
            ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_void_p]
        lib.orc580_intersect_batch.argtypes = [ctypes.c_void_p, ctypes.c_int64] + [ctypes.c_void_p] * 4 + [ctypes.c_int]
        lib.orc580_render_log.restype = ctypes.c_int64
        lib.orc580_render_log.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 4 + [ctypes.c_int64, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_int64] + [ctypes.c_void_p] * 6
        lib.orc580_gamma_encode.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
        _t1 = lib
    return _t1


def _ptr(a):
    return a.ctypes.data if a is not None and a.size else None


class Oracle:
    """T1 oracle over one scene."""

    def __init__(self, scene: SceneArrays):
        self.scene = scene
        s = _OrcScene()
        s.n_shapes = len(scene.shape_mesh)
        s.shape_mesh, s.shape_material, s.shape_srt = _ptr(scene.shape_mesh), _ptr(scene.shape_material), _ptr(scene.shape_srt)
        s.n_meshes = len(scene.mesh_type)
        s.mesh_type, s.mesh_tri_begin, s.mesh_radius = _ptr(scene.mesh_type), _ptr(scene.mesh_tri_begin), _ptr(scene.mesh_radius)
        s.tri_pos, s.tri_nrm = _ptr(scene.tri_pos), _ptr(scene.tri_nrm)
        s.n_lights = len(scene.light_type)
        s.light_type, s.light_f = _ptr(scene.light_type), _ptr(scene.light_f)
        s.cam_from[:] = [float(x) for x in scene.cam_from]
        s.cam_to[:] = [float(x) for x in scene.cam_to]
        self._world = t1_lib().orc580_prepare(ctypes.byref(s))

    def __del__(self):
        try:
            if self._world:
                t1_lib().orc580_free(self._world)
                self._world = None
        except Exception:
            pass

    @property
    def n_prims(self):
        return t1_lib().orc580_num_prims(self._world)

    def render(self, W, H, spp, depth, pix=None, ao_base=None, nthreads=1):
        """-> (fb int16 [npix,3] (or [H,W,3] for a full frame), rays, hit_nodes[npix])"""
        if pix is not None:
            pix = np.ascontiguousarray(pix, np.int32)
            n = pix.size
        else:
            n = W * H
        if ao_base is not None:
            ao_base = np.ascontiguousarray(ao_base, np.uint64)
            assert ao_base.size == n
        out = np.zeros((n, 3), np.int16)
        hits = np.zeros(n, np.uint32)
        rays = ctypes.c_uint64(0)
        st = t1_lib().orc580_render(self._world, W, H, spp, depth, n, _ptr(pix) if pix is not None else None,
                                    _ptr(ao_base) if ao_base is not None else None, nthreads,
                                    out.ctypes.data, ctypes.addressof(rays), hits.ctypes.data)
        if st != RT_SUCCESS:
            raise RuntimeError("orc580_render status %d" % st)
        if pix is None:
            out = out.reshape(H, W, 3)
        return out, rays.value, hits

    def render_log(self, W, H, spp, depth, pix=None, ao_base=None, max_rays=1 << 22):
        """Serial render that logs every IntersectScene call: -> dict(org, dir, prim, t, kind, fb)"""
        if pix is not None:
            pix = np.ascontiguousarray(pix, np.int32)
            n = pix.size
        else:
            n = W * H
        if ao_base is not None:
            ao_base = np.ascontiguousarray(ao_base, np.uint64)
        org = np.zeros((max_rays, 3), np.float32); dirs = np.zeros((max_rays, 3), np.float32)
        prim = np.zeros(max_rays, np.int64); t = np.zeros(max_rays, np.float32); kind = np.zeros(max_rays, np.int32)
        out = np.zeros((n, 3), np.int16)
        m = t1_lib().orc580_render_log(self._world, W, H, spp, depth, n, _ptr(pix) if pix is not None else None,
                                       _ptr(ao_base) if ao_base is not None else None, max_rays, org.ctypes.data,
                                       dirs.ctypes.data, prim.ctypes.data, t.ctypes.data, kind.ctypes.data, out.ctypes.data)
        if m < 0:
            raise RuntimeError("orc580_render_log failed")
        return {"org": org[:m], "dir": dirs[:m], "prim": prim[:m], "t": t[:m], "kind": kind[:m], "fb": out}

    def intersect(self, org, dirs, nthreads=1):
        org = np.ascontiguousarray(org, np.float32).reshape(-1, 3)
        dirs = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        n = org.shape[0]
        prim = np.zeros(n, np.int64)
        t = np.zeros(n, np.float32)
        t1_lib().orc580_intersect_batch(self._world, n, org.ctypes.data, dirs.ctypes.data, prim.ctypes.data,
                                        t.ctypes.data, nthreads)
        return prim, t

    def primary_ray(self, W, H, x, y):
        o = np.zeros(3, np.float32)
        d = np.zeros(3, np.float32)
        st = t1_lib().orc580_primary_ray(self._world, W, H, x, y, o.ctypes.data, d.ctypes.data)
        return st, o, d


def model_matrix(srt):
    srt = np.ascontiguousarray(srt, np.float32)
    m = np.zeros(16, np.float32)
    t1_lib().orc580_model_matrix(srt.ctypes.data, m.ctypes.data)
    return m.reshape(4, 4)


def fresnel(ior, n, i):
    n = np.ascontiguousarray(n, np.float32)
    i = np.ascontiguousarray(i, np.float32)
    kr, kt = ctypes.c_float(), ctypes.c_float()
    refr = np.zeros(3, np.float32)
    t1_lib().orc580_fresnel(ior, n.ctypes.data, i.ctypes.data, ctypes.addressof(kr), ctypes.addressof(kt), refr.ctypes.data)
    return kr.value, kt.value, refr


def hemisphere_stream(normal, step, n):
    normal = np.ascontiguousarray(normal, np.float32)
    out = np.zeros((n, 3), np.float32)
    t1_lib().orc580_hemisphere_stream(normal.ctypes.data, step, n, out.ctypes.data)
    return out


def lcg_state(steps):
    return t1_lib().orc580_lcg_state(steps)


def gamma_encode(fb):
    fb = np.ascontiguousarray(fb, np.int16)
    out = np.zeros(fb.shape, np.uint8)
    t1_lib().orc580_gamma_encode(fb.ctypes.data, fb.size, out.ctypes.data)
    return out


# --------------------------------------------------------------------------------------
# T0
# --------------------------------------------------------------------------------------
_t0 = None


def t0_available():
    return os.path.exists(T0_LIB)


def t0_lib():
    global _t0
    if _t0 is None:
        if not os.path.exists(T0_LIB):
            raise RuntimeError("oracle/_ref/libref580.so missing: run oracle/build_ref.sh where /root/reference exists")
        lib = ctypes.CDLL(T0_LIB)
        lib.ref580_render.argtypes = [ctypes.c_char_p, ctypes.c_char_p] + [ctypes.c_int] * 4 + [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_char_p]
        lib.ref580_render_pixels.argtypes = [ctypes.c_char_p, ctypes.c_char_p] + [ctypes.c_int] * 4 + [
            ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        lib.ref580_dump_scene.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int] + [ctypes.c_void_p] * 4 + [
            ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        lib.ref580_model_matrix.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        lib.ref580_fresnel.argtypes = [ctypes.c_float] + [ctypes.c_void_p] * 5
        lib.ref580_hemisphere_stream.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        _t0 = lib
    return _t0


def _dir(assets_dir):
    d = assets_dir if assets_dir.endswith("/") else assets_dir + "/"
    return d.encode()


def t0_render(assets_dir, scene, W, H, spp, depth, ppm_out=None):
    """-> (status, fb int16 [H,W,3], rays, seconds)"""
    fb = np.zeros((H, W, 3), np.int16)
    rays, secs = ctypes.c_uint64(0), ctypes.c_double(0)
    st = t0_lib().ref580_render(_dir(assets_dir), scene.encode(), W, H, spp, depth, fb.ctypes.data,
                                ctypes.addressof(rays), ctypes.addressof(secs), (ppm_out or "").encode())
    return st, fb, rays.value, secs.value


def t0_render_pixels(assets_dir, scene, W, H, spp, depth, pix, nthreads=1):
    pix = np.ascontiguousarray(pix, np.int32)
    out = np.zeros((pix.size, 3), np.int16)
    rays, secs = ctypes.c_uint64(0), ctypes.c_double(0)
    st = t0_lib().ref580_render_pixels(_dir(assets_dir), scene.encode(), W, H, spp, depth, pix.size, pix.ctypes.data,
                                       nthreads, out.ctypes.data, ctypes.addressof(rays), ctypes.addressof(secs))
    return st, out, rays.value, secs.value


def t0_dump_scene(assets_dir, scene, max_shapes=65536, max_lights=64):
    sf = np.zeros((max_shapes, 17), np.float32)
    ntri = np.zeros(max_shapes, np.int32)
    rad = np.zeros(max_shapes, np.float32)
    lf = np.zeros((max_lights, 11), np.float32)
    ns, nl = ctypes.c_int(0), ctypes.c_int(0)
    cam = np.zeros(6, np.float32)
    st = t0_lib().ref580_dump_scene(_dir(assets_dir), scene.encode(), max_shapes, sf.ctypes.data, ntri.ctypes.data,
                                    rad.ctypes.data, ctypes.addressof(ns), max_lights, lf.ctypes.data,
                                    ctypes.addressof(nl), cam.ctypes.data)
    return st, sf[:ns.value], ntri[:ns.value], rad[:ns.value], lf[:nl.value], cam


def t0_model_matrix(srt):
    srt = np.ascontiguousarray(srt, np.float32)
    m = np.zeros(16, np.float32)
    t0_lib().ref580_model_matrix(srt.ctypes.data, m.ctypes.data)
    return m.reshape(4, 4)


def t0_fresnel(ior, n, i):
    n = np.ascontiguousarray(n, np.float32)
    i = np.ascontiguousarray(i, np.float32)
    kr, kt = ctypes.c_float(), ctypes.c_float()
    refr = np.zeros(3, np.float32)
    t0_lib().ref580_fresnel(ior, n.ctypes.data, i.ctypes.data, ctypes.addressof(kr), ctypes.addressof(kt), refr.ctypes.data)
    return kr.value, kt.value, refr


def t0_hemisphere_stream(normal, n):
    normal = np.ascontiguousarray(normal, np.float32)
    out = np.zeros((n, 3), np.float32)
    t0_lib().ref580_hemisphere_stream(normal.ctypes.data, n, out.ctypes.data)
    return out
