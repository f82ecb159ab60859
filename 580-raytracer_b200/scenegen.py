"""Synthetic scene generator for the benchmark configurations.

Writes scene + mesh JSON files in the reference's own schema (Raytracer.cpp:589-779), so the
unmodified reference loader, the oracle and this repo's loader all read the same input:
instanced `teapot` shapes on a jittered grid over a floor, plus sphere shapes, per-shape
materials, ambient + directional + point light.  Deterministic for a given seed (580).

Instancing keeps the JSON small: a shape names its mesh by `geometry`, and the loader caches
meshes by name (cpp:590-594), so 1 M triangles are 977 shape records + one teapot file.
"""
import json
import math
import os
import random

TEAPOT_TRIS = 1024


def _mesh_floor(half, y):
    """two upward-facing triangles (the reference's floor.json winds the other way, Q7)"""
    def vert(x, z, u, v):
        return {"v": [x, y, z], "n": [0.0, 1.0, 0.0], "t": [u, v]}
    return {"data": [
        {"type": "polygon", "v0": vert(-half, -half, 0.0, 0.0), "v1": vert(half, half, 1.0, 1.0), "v2": vert(half, -half, 1.0, 0.0)},
        {"type": "polygon", "v0": vert(-half, -half, 0.0, 0.0), "v1": vert(-half, half, 0.0, 1.0), "v2": vert(half, half, 1.0, 1.0)},
    ]}


def _translation_for(p, scale, ry_deg):
    """The reference's model matrix is S * R * T (cpp:584, Q11): the translation is rotated and
    scaled.  Return T such that the shape lands at world position p: T = R^-1 S^-1 p."""
    a = math.radians(ry_deg)
    c, s = math.cos(a), math.sin(a)
    q = [p[0] / scale[0], p[1] / scale[1], p[2] / scale[2]]
    # Ry = [[c,0,s],[0,1,0],[-s,0,c]]  ->  Ry^-1 = Ry^T
    return [c * q[0] - s * q[2], q[1], s * q[0] + c * q[2]]


def _mesh_room(half, y0, y1, shell=1.0):
    """Closed room with thick walls: an inner box (floor, ceiling, four walls: 12 triangles) and an
    outer box `shell` units further out (12 more), all geometric normals facing inward.  The outer
    box catches the rays the reference starts 0.2 units BEHIND an inner wall (its shadow / AO /
    reflection origins are hit point + 0.2 * direction, cpp:67, cpp:98, cpp:322, so a hit closer
    than 0.2 to a wall tunnels through it)."""
    def tri(a, b, c, n):
        return {"type": "polygon", "v0": {"v": a, "n": n, "t": [0.0, 0.0]}, "v1": {"v": b, "n": n, "t": [1.0, 0.0]},
                "v2": {"v": c, "n": n, "t": [1.0, 1.0]}}
    data = []
    for h, ya, yb in ((half, y0, y1), (half + shell, y0 - shell, y1 + shell)) if shell > 0 else ((half, y0, y1),):
        data += _box_inward(h, ya, yb, tri)
    return {"data": data}


def _box_inward(h, y0, y1, tri):
    c = [[-h, y0, -h], [h, y0, -h], [h, y0, h], [-h, y0, h], [-h, y1, -h], [h, y1, -h], [h, y1, h], [-h, y1, h]]
    quads = [((0, 2, 1), (0, 3, 2), [0.0, 1.0, 0.0]),      # floor (normal up)
             ((4, 5, 6), (4, 6, 7), [0.0, -1.0, 0.0]),     # ceiling
             ((0, 1, 5), (0, 5, 4), [0.0, 0.0, 1.0]),      # z = -h wall
             ((3, 6, 2), (3, 7, 6), [0.0, 0.0, -1.0]),     # z = +h wall
             ((0, 4, 7), (0, 7, 3), [1.0, 0.0, 0.0]),      # x = -h wall
             ((1, 2, 6), (1, 6, 5), [-1.0, 0.0, 0.0])]     # x = +h wall
    def inward(a, b, cc, n):
        """wind the triangle so that its GEOMETRIC normal cross(v1-v0, v2-v0) points into the room: the
        reference shoots AO rays around the geometric normal as is (Q7)"""
        e1 = [b[i] - a[i] for i in range(3)]
        e2 = [cc[i] - a[i] for i in range(3)]
        g = [e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]]
        return (a, b, cc) if sum(g[i] * n[i] for i in range(3)) > 0 else (a, cc, b)
    data = []
    for t0, t1, n in quads:
        for t in (t0, t1):
            a, b, cc = inward(c[t[0]], c[t[1]], c[t[2]], n)
            data.append(tri(a, b, cc, n))
    return data


def write_synthetic_scene(out_dir, name, n_teapots, n_spheres, seed=580, spacing=4.5, teapot_mesh="teapot",
                          with_floor=True, point_light=True, room=False):
    """Returns a dict describing the scene (triangle count, suggested camera...)."""
    rng = random.Random(seed)
    os.makedirs(out_dir, exist_ok=True)
    nx = max(1, int(math.ceil(math.sqrt(n_teapots))))
    nz = max(1, int(math.ceil(n_teapots / nx)))
    half_x, half_z = 0.5 * (nx - 1) * spacing, 0.5 * (nz - 1) * spacing
    half = max(half_x, half_z) + 2.0 * spacing
    shapes = []

    def material():
        ks = 0.0 if rng.random() < 0.5 else round(rng.uniform(0.2, 0.9), 3)
        kt = 0.0 if rng.random() < 0.75 else round(rng.uniform(0.1, 0.8), 3)
        return {"Cs": [round(rng.random(), 3), round(rng.random(), 3), round(rng.random(), 3)],
                "Ka": round(rng.uniform(0.1, 0.5), 3), "Kd": round(rng.uniform(0.3, 0.9), 3), "Ks": ks, "Kt": kt,
                "n": rng.choice([2, 5, 10, 32, 700])}

    room_height = round(0.8 * half + 12.0, 3)
    if room:
        # Closed room instead of an open floor: no ray leaves the scene, so the reference's far-field
        # acceptances (float noise on rays that escape) never decide a pixel.
        room_name = "room_%s" % name
        with open(os.path.join(out_dir, room_name + ".json"), "w") as f:
            json.dump(_mesh_room(round(half + 2.0 * spacing, 3), -0.4, room_height), f)
        shapes.append({"id": "room", "geometry": room_name,
                       "material": {"Cs": [0.6, 0.6, 0.6], "Ka": 0.2, "Kd": 0.7, "Ks": 0.3, "Kt": 0.0, "n": 32},
                       "transforms": [{"S": [1, 1, 1]}, {"T": [0, 0, 0]}]})
    elif with_floor:
        floor_name = "floor_%s" % name
        with open(os.path.join(out_dir, floor_name + ".json"), "w") as f:
            json.dump(_mesh_floor(round(half, 3), -0.4), f)
        shapes.append({"id": "floor", "geometry": floor_name,
                       "material": {"Cs": [0.5, 0.5, 0.5], "Ka": 0.2, "Kd": 0.7, "Ks": 0.3, "Kt": 0.0, "n": 32},
                       "transforms": [{"S": [1, 1, 1]}, {"T": [0, 0, 0]}]})
    for i in range(n_teapots):
        gx, gz = i % nx, i // nx
        # y >= 0.4: the teapot spans y in [-0.351, 0.569] * S, S <= 2, so it stays above the floor at -0.4
        p = [gx * spacing - half_x + rng.uniform(-0.8, 0.8), rng.uniform(0.4, 1.0), gz * spacing - half_z + rng.uniform(-0.8, 0.8)]
        ry = round(rng.uniform(0.0, 360.0), 2)
        sc = [round(rng.uniform(0.5, 2.0), 3) for _ in range(3)]
        t = [round(v, 4) for v in _translation_for(p, sc, ry)]
        shapes.append({"id": "teapot%d" % i, "geometry": teapot_mesh, "material": material(),
                       "transforms": [{"Ry": ry}, {"S": sc}, {"T": t}]})
    radii = [0.5, 1.0, 1.5]
    for r in radii:
        with open(os.path.join(out_dir, "sphere_r%03d.json" % int(r * 100)), "w") as f:
            json.dump({"data": [{"type": "sphere", "radius": r}]}, f)
    for i in range(n_spheres):
        r = rng.choice(radii)
        p = [rng.uniform(-half_x, half_x) if half_x > 0 else rng.uniform(-2, 2), rng.uniform(2.5, 7.0),
             rng.uniform(-half_z, half_z) if half_z > 0 else rng.uniform(-2, 2)]
        shapes.append({"id": "sphere%d" % i, "geometry": "sphere_r%03d" % int(r * 100), "material": material(),
                       "transforms": [{"S": [1, 1, 1]}, {"T": [round(v, 4) for v in p]}]})
    lights = [{"id": "ambientLight", "type": "ambient", "color": [1, 1, 1], "intensity": 0.2}]
    if not room:
        lights.append({"id": "directionalLight", "type": "directional", "color": [1, 1, 1], "intensity": 1.0,
                       "from": [1, 10, 1], "to": [0, 0, 0]})
    if point_light:
        lights.append({"id": "pointLight", "type": "point", "color": [1, 0.9, 0.8], "intensity": 0.8,
                       "position": [0, round(min(0.6 * half + 8.0, room_height - 2.0), 3), 0]})
    if room:
        # a closed room has no use for a directional light (the ceiling blocks it everywhere, and its
        # shadow rays would start outside the walls it grazes): two more point lights inside instead
        lights.append({"id": "pointLight2", "type": "point", "color": [0.8, 0.9, 1.0], "intensity": 0.6,
                       "position": [round(0.7 * half, 3), round(0.5 * room_height, 3), round(0.7 * half, 3)]})
        lights.append({"id": "pointLight3", "type": "point", "color": [1.0, 1.0, 1.0], "intensity": 0.5,
                       "position": [round(-0.6 * half, 3), round(0.3 * room_height, 3), round(-0.5 * half, 3)]})
        cam_from = [0, round(0.45 * room_height, 3), round(half + 1.5 * spacing, 3)]      # inside the room
    else:
        cam_from = [0, round(0.55 * half + 3.0, 3), round(1.35 * half + 6.0, 3)]
    scene = {"scene": {"shapes": shapes, "lights": lights,
                       "camera": {"from": cam_from, "to": [0, 0, 0], "bounds": [0.1, 1000, 0, 5, 5, 0],
                                  "resolution": [3840, 2160]}}}
    with open(os.path.join(out_dir, name + ".json"), "w") as f:
        json.dump(scene, f)
    return {"scene": name + ".json", "n_triangles": n_teapots * TEAPOT_TRIS + (24 if room else (2 if with_floor else 0)),
            "n_spheres": n_spheres, "n_shapes": len(shapes), "camera_from": cam_from, "half_extent": half}


# the named benchmark configurations (BASELINE.json configs[3], configs[4])
CONFIGS = {
    "c4_room": dict(n_teapots=977, n_spheres=1000, room=True),    # 1,000,448 teapot triangles + 1k spheres in a closed room
    "c4_open": dict(n_teapots=977, n_spheres=1000, room=False),   # same on an open floor (rays escape)
    "c5_room": dict(n_teapots=9766, n_spheres=0, room=True),      # 10,000,384 triangles
    "c5_open": dict(n_teapots=9766, n_spheres=0, room=False),     # same on an open floor (BASELINE configs[4] as SURVEY 8d specifies it)
}
