"""xrtg_scene_desc assembled in pure Python / numpy — no libxrthost.so, no libxrtgpu.so.

Used by `bench.py --impl reference` (and the cpu_baseline leg) so that the process which times the reference's CPU renderer
maps nothing but oracle/ libraries: the benchmark scenes are restated here from the same tables (scenes.CORNELL_SHAPES, the
displaced sphere, the procedural grid) and listed in INSERTION order with `insert_seq`; the CPU checkers re-insert them in
that order into the reference's own `Scene` (oracle/ref_harness.cpp: xrtref_scene_create), so the reference's
std::unordered_map decides the iteration order exactly as it does for the host C++ scene.

NOT a product path: libxrtgpu.so wants objects[] in the reference's iteration order, which only the host library knows
(same container, same keys). tests/test_host.py checks that both builders describe the same scene.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import capi


def _flat_normal(v0, v1, v2):
    n = np.cross(v1 - v0, v2 - v0).astype(np.float32)
    ln = np.float32(np.sqrt(np.float32(n[0] * n[0] + n[1] * n[1] + n[2] * n[2])))
    return (n / ln).astype(np.float32)


class FlatScene:
    """Insertion-ordered scene description; `desc()` returns a POINTER(SceneDesc) whose arrays this object keeps alive."""

    def __init__(self):
        self.objects = []    # (name, kind, first, count, material, area_light, medium)
        self.tris = []       # (18,) float32 rows
        self.spheres, self.boxes, self.materials, self.area_lights, self.delta_lights, self.media, self.grids = [], [], [], [], [], [], []
        self._keep = []

    # ---- geometry ----
    def _material(self, albedo):
        if albedo is None:
            return -1
        self.materials.append(tuple(float(x) for x in albedo))
        return len(self.materials) - 1

    def add_mesh(self, name, tris, albedo, area_light=-1):
        tris = np.ascontiguousarray(tris, dtype=np.float32).reshape(-1, 18)
        first = sum(len(t) for t in self.tris)
        self.tris.append(tris)
        self.objects.append((name, capi.OBJ_MESH, first, len(tris), self._material(albedo) if area_light < 0 else -1, area_light, -1))

    def add_quads(self, name, quads, albedo):
        """OBJ-style quads, fan-triangulated (0,1,2), (0,2,3) with flat normals from the winding (scene.cpp:118-125)."""
        rows = []
        for q in quads:
            v = [np.asarray(p, np.float32) for p in q]
            for a, b, c in ((0, 1, 2), (0, 2, 3)):
                n = _flat_normal(v[a], v[b], v[c])
                rows.append(np.concatenate([v[a], v[b], v[c], n, n, n]))
        self.add_mesh(name, np.array(rows, np.float32), albedo)

    def add_quad_light(self, name, v0, v1, v2, Le):
        """QuadLight + its proxy mesh (v0,v1,v2), (v1,v3,v2) (light.cpp:70-82)."""
        v0, v1, v2 = (np.asarray(v, np.float32) for v in (v0, v1, v2))
        v3 = (v0 + (v1 - v0) + (v2 - v0)).astype(np.float32)
        n = _flat_normal(v0, v1, v2)
        li = len(self.area_lights)
        self.area_lights.append((capi.LIGHT_QUAD, v0, v1, v2, 0.0, tuple(float(x) for x in Le)))
        self.add_mesh(name, np.array([np.concatenate([v0, v1, v2, n, n, n]), np.concatenate([v1, v3, v2, n, n, n])], np.float32), None, area_light=li)

    def add_triangle_light(self, name, v0, v1, v2, Le):
        """TriangleLight + its one-triangle proxy mesh (light.cpp:32-47)."""
        v0, v1, v2 = (np.asarray(v, np.float32) for v in (v0, v1, v2))
        n = _flat_normal(v0, v1, v2)
        li = len(self.area_lights)
        self.area_lights.append((capi.LIGHT_TRIANGLE, v0, v1, v2, 0.0, tuple(float(x) for x in Le)))
        self.add_mesh(name, np.array([np.concatenate([v0, v1, v2, n, n, n])], np.float32), None, area_light=li)

    def add_heterogeneous_medium(self, name, g, voxels, origin, voxel_size, abs_color, scat_color, mul=1.0):
        """HeterogeneousMedium over a dense grid + its BoxMesh proxy over the active-voxel bounds (grid.h:58-69, medium.h)."""
        v = np.ascontiguousarray(voxels, dtype=np.float32)
        nz, ny, nx = v.shape
        self._keep.append(v)
        nzr = np.nonzero(v != 0.0)
        if len(nzr[0]):
            lo = [int(nzr[2].min()), int(nzr[1].min()), int(nzr[0].min())]
            hi = [int(nzr[2].max()), int(nzr[1].max()), int(nzr[0].max())]
        else:
            lo, hi = [0, 0, 0], [nx - 1, ny - 1, nz - 1]
        vs = np.float32(voxel_size)
        org = np.asarray(origin, np.float32)
        self.grids.append((nx, ny, nz, v, org, float(vs), 0.0, lo, hi, float(max(v.max(), 0.0))))
        self.media.append((capi.MEDIUM_HETEROGENEOUS, float(g), tuple(map(float, abs_color)), tuple(map(float, scat_color)), float(mul), len(self.grids) - 1))
        pmin = [float(np.float32(org[a] + vs * np.float32(lo[a]))) for a in range(3)]
        pmax = [float(np.float32(org[a] + vs * np.float32(hi[a] + 1))) for a in range(3)]
        self.boxes.append((pmin, pmax))
        self.objects.append((name, capi.OBJ_BOX, len(self.boxes) - 1, 1, -1, -1, len(self.media) - 1))

    # ---- the C structure ----
    def desc(self):
        d = capi.SceneDesc()
        d.abi_version = capi.ABI_VERSION
        tris = np.concatenate(self.tris) if self.tris else np.zeros((0, 18), np.float32)
        tris = np.ascontiguousarray(tris, dtype=np.float32)
        self._keep.append(tris)
        objs = (capi.Object * max(len(self.objects), 1))()
        for i, (name, kind, first, count, mat, light, med) in enumerate(self.objects):
            o = objs[i]
            o.kind, o.first, o.count, o.material, o.area_light, o.medium, o.insert_seq = kind, first, count, mat, light, med, i
            o.name = name.encode()
        mats = (capi.Material * max(len(self.materials), 1))()
        for i, a in enumerate(self.materials):
            mats[i].kind = 0
            mats[i].albedo[:] = a
        lights = (capi.AreaLight * max(len(self.area_lights), 1))()
        for i, (kind, v0, v1, v2, r, Le) in enumerate(self.area_lights):
            L = lights[i]
            L.kind, L.radius = kind, r
            L.v0[:], L.v1[:], L.v2[:], L.Le[:] = [float(x) for x in v0], [float(x) for x in v1], [float(x) for x in v2], Le
        media = (capi.Medium * max(len(self.media), 1))()
        for i, (kind, g, sa, ss, mul, grid) in enumerate(self.media):
            m = media[i]
            m.kind, m.g, m.density_mul, m.grid = kind, g, mul, grid
            m.sigma_a[:], m.sigma_s[:] = sa, ss
        grids = (capi.Grid * max(len(self.grids), 1))()
        for i, (nx, ny, nz, v, org, vs, bg, lo, hi, mx) in enumerate(self.grids):
            g = grids[i]
            g.nx, g.ny, g.nz, g.voxel_size, g.background, g.max_density = nx, ny, nz, vs, bg, mx
            g.data = v.ctypes.data_as(C.POINTER(C.c_float))
            g.origin[:] = [float(x) for x in org]
            g.active_min[:], g.active_max[:] = lo, hi
        boxes = (capi.Box * max(len(self.boxes), 1))()
        for i, (pmin, pmax) in enumerate(self.boxes):
            boxes[i].pmin[:], boxes[i].pmax[:] = pmin, pmax
        d.n_objects, d.n_triangles, d.n_spheres, d.n_boxes = len(self.objects), len(tris), 0, len(self.boxes)
        d.n_materials, d.n_area_lights, d.n_delta_lights, d.n_media, d.n_grids = len(self.materials), len(self.area_lights), 0, len(self.media), len(self.grids)
        d.objects = objs
        d.triangles = C.cast(tris.ctypes.data, C.POINTER(capi.Triangle))
        d.boxes, d.materials, d.area_lights, d.media, d.grids = boxes, mats, lights, media, grids
        self._keep += [objs, mats, lights, media, grids, boxes, d]
        return C.pointer(d)


def make_camera(width, height, c2w, fov) -> capi.Camera:
    """PinholeCamera (camera.h:41-47): scale = tan(0.5 * deg2rad(FOV)) in fp32, aspect = W / H."""
    cam = capi.Camera()
    cam.c2w[:] = [float(x) for x in c2w]
    cam.scale = float(np.tan(np.float32(0.5) * np.float32(np.float32(fov) * np.float32(math.pi) / np.float32(180.0)), dtype=np.float32))
    cam.aspect = float(np.float32(width) / np.float32(height))
    return cam


def cornell_box(light: str = "quad") -> FlatScene:
    """scenes.cornell_box(light) without the host library: shapes in OBJ order (face-less shapes dropped), then the light
    ("quad", or "triangle" = half of the quad)."""
    from . import scenes
    s = FlatScene()
    for shape, mat, quads in scenes.CORNELL_SHAPES:
        if quads:
            s.add_quads(shape, quads, scenes.CORNELL_MATERIALS[mat])
    q = scenes.CORNELL_QUAD_LIGHT
    if light == "triangle":
        s.add_triangle_light("TriangleLight", q["v0"], q["v1"], q["v2"], q["Le"])
    else:
        s.add_quad_light("QuadLight", q["v0"], q["v1"], q["v2"], q["Le"])
    return s


def cornell_mesh_scene(ntheta=707, nphi=707) -> FlatScene:
    from . import scenes
    s = cornell_box()
    s.add_mesh("tess_sphere", scenes.displaced_sphere_tris((278.0, 200.0, 280.0), 150.0, ntheta, nphi), (0.75, 0.75, 0.75))
    return s


def volume_scene(n=256, abs_color=(0.01, 0.01, 0.01), scat_color=(0.05, 0.05, 0.05)) -> FlatScene:
    from . import scenes
    s = FlatScene()
    vox = scenes.procedural_density(n)
    size = 300.0
    s.add_heterogeneous_medium("medium", 0.0, vox, (278.0 - size / 2, 274.0 - size / 2, 280.0 - size / 2), size / vox.shape[0], abs_color, scat_color, 1.0)
    q = scenes.CORNELL_QUAD_LIGHT
    s.add_quad_light("QuadLight", q["v0"], q["v1"], q["v2"], q["Le"])
    return s
