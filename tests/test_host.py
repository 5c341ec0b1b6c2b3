"""Host-side logic: OBJ ingest, the host C++ API mirror and the flattening boundary (no GPU needed)."""
import ctypes as C

import numpy as np
import pytest

from xraytracer_b200 import api, capi, scenes

# std::unordered_map<std::string,...> iteration order of the Cornell scene under libstdc++ (the order
# Scene::intersect walks objects, scene.cpp:193); pinned against the compiled reference in test_oracle_vs_reference.py
CORNELL_ORDER = ["tall_block", "short_block", "QuadLight", "green_wall", "red_wall", "back_wall", "ceiling", "floor"]


def test_cornell_flatten(cornell):
    host, desc = cornell
    d = desc.contents
    assert host.object_names() == CORNELL_ORDER
    # 34 OBJ triangles (17 quads fan-triangulated) + 2 light proxy triangles; face-less shapes dropped
    assert d.n_triangles == 36 and d.n_objects == 8 and d.n_spheres == 0 and d.n_boxes == 0
    assert d.n_area_lights == 1 and d.n_materials == 3  # white, green, red (one Lambert per used MTL entry)
    objs = [d.objects[i] for i in range(d.n_objects)]
    assert sorted(o.insert_seq for o in objs) == list(range(8))
    light = [o for o in objs if o.area_light >= 0]
    assert len(light) == 1 and light[0].material == -1 and light[0].count == 2
    # ranges tile the triangle array exactly once
    covered = sorted((o.first, o.first + o.count) for o in objs)
    assert covered[0][0] == 0 and covered[-1][1] == 36
    for a, b in zip(covered, covered[1:]):
        assert a[1] == b[0]


def test_obj_fan_triangulation_and_flat_normals(cornell):
    _, desc = cornell
    d = desc.contents
    floor = [d.objects[i] for i in range(d.n_objects) if d.objects[i].name == b"floor"][0]
    assert floor.count == 6
    t0, t1 = d.triangles[floor.first], d.triangles[floor.first + 1]
    q = scenes.CORNELL_SHAPES[0][2][0]
    assert tuple(t0.v0) == pytest.approx(q[0]) and tuple(t0.v1) == pytest.approx(q[1]) and tuple(t0.v2) == pytest.approx(q[2])
    assert tuple(t1.v0) == pytest.approx(q[0]) and tuple(t1.v1) == pytest.approx(q[2]) and tuple(t1.v2) == pytest.approx(q[3])
    # no vn in the file -> flat normal from the winding (scene.cpp:118-125); the floor faces +y
    assert tuple(t0.n0) == pytest.approx((0, 1, 0), abs=1e-6) and tuple(t0.n0) == tuple(t0.n1) == tuple(t0.n2)


def test_quad_light_description(cornell):
    _, desc = cornell
    L = desc.contents.area_lights[0]
    q = scenes.CORNELL_QUAD_LIGHT
    assert L.kind == capi.LIGHT_QUAD and tuple(L.v0) == q["v0"] and tuple(L.v1) == q["v1"] and tuple(L.v2) == q["v2"]
    assert tuple(L.Le) == q["Le"]


def test_camera_scale_matches_reference_formula():
    cam = scenes.make_camera(1920, 1080)
    # camera.h:44: scale = tan(0.5f * deg2rad(FOV)), deg2rad = deg / 180.0f * PI in fp32
    pi = np.float32(3.14159265359)
    rad = np.float32(np.float32(np.float32(60.0) / np.float32(180.0)) * pi)
    assert cam.scale == pytest.approx(float(np.tan(np.float32(0.5) * rad)), rel=1e-6)
    assert cam.aspect == pytest.approx(1920 / 1080)
    assert list(cam.c2w) == [float(np.float32(x)) for x in scenes.CORNELL_C2W]


def test_media_spheres_and_delta_lights_flatten():
    s = scenes.HostScene()
    vox = np.zeros((4, 4, 4), np.float32)
    vox[1:3, 1:3, 1:4] = 0.5
    s.add_heterogeneous_medium("medium", 0.25, vox, (0, 0, 0), 2.0, (0.1, 0.2, 0.3), (0.4, 0.5, 0.6), 2.0)
    s.add_sphere_light("SphereLight", (0, 10, 0), 1.5, (3, 3, 3), l2w=[1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 2, 0, 1])
    s.add_sphere("ball", (1, 2, 3), 0.5, (0.2, 0.3, 0.4))
    s.add_point_light("p", [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 5, 5, -1, 1], (0.5, 0.25, 1.0), 50.0)
    s.add_distant_light("d", None, (1, 1, 1), 2.0)
    d = s.flatten().contents
    assert d.n_boxes == 1 and d.n_spheres == 2 and d.n_media == 1 and d.n_grids == 1 and d.n_delta_lights == 2
    g = d.grids[0]
    assert (g.nx, g.ny, g.nz) == (4, 4, 4) and g.max_density == 0.5
    assert list(g.active_min) == [1, 1, 1] and list(g.active_max) == [3, 2, 2]
    b = d.boxes[0]  # indexToWorld(active_min), indexToWorld(active_max + 1) (grid.h:58-69)
    assert tuple(b.pmin) == (2, 2, 2) and tuple(b.pmax) == (8, 6, 6)
    m = d.media[0]
    assert m.kind == capi.MEDIUM_HETEROGENEOUS and m.density_mul == 2.0 and m.grid == 0 and m.g == 0.25
    L = d.area_lights[0]
    assert L.kind == capi.LIGHT_SPHERE and tuple(L.v0) == (0, 12, 0) and L.radius == 1.5  # centre through lightToWorld
    assert tuple(d.delta_lights[0].pos_or_dir) == (5, 5, -1) and tuple(d.delta_lights[0].radiance) == (25.0, 12.5, 50.0)
    assert tuple(d.delta_lights[1].pos_or_dir) == (0, 0, -1) and tuple(d.delta_lights[1].radiance) == (2, 2, 2)


def test_sphere_mesh_tessellation_counts():
    s = scenes.HostScene()
    s.add_sphere_mesh("sm", (0, 0, 0), 1.0, 6, 8, (1, 1, 1))
    d = s.flatten().contents
    assert d.n_triangles == 2 * 6 * 8  # primitive.cpp:187-204: two triangles per (theta, phi) cell
    v = np.array([[list(d.triangles[i].v0), list(d.triangles[i].v1), list(d.triangles[i].v2)] for i in range(d.n_triangles)])
    assert np.allclose(np.linalg.norm(v, axis=-1), 1.0, atol=1e-6)


def test_empty_scene_flattens():
    s = scenes.HostScene()
    d = s.flatten().contents
    assert d.n_objects == 0 and d.n_triangles == 0


def test_missing_obj_raises():
    s = scenes.HostScene()
    with pytest.raises(RuntimeError, match="failed to load"):
        s.load_obj("/nonexistent/file.obj")


def test_displaced_sphere_generator_is_deterministic():
    a = scenes.displaced_sphere_tris((0, 0, 0), 1.0, 12, 16)
    b = scenes.displaced_sphere_tris((0, 0, 0), 1.0, 12, 16)
    assert a.shape == (2 * 12 * 16, 18) and np.array_equal(a, b) and np.isfinite(a).all()


def test_pure_python_scene_description_equals_host_scene():
    """xraytracer_b200/flatdesc.py restates the benchmark scenes without libxrthost.so (the CPU reference arm of bench.py must not
    map product libraries). Through the compiled reference both descriptions must render the very same image: the checker
    re-inserts the objects in insertion order into the reference's own map, so ids, tie-breaks and sample streams agree."""
    import numpy as np
    from xraytracer_b200 import api, capi, flatdesc, scenes
    if not capi.have_reference():
        pytest.skip("oracle/_ref/libxrtref.so not built")
    W, H = 48, 36
    cam_h = scenes.make_camera(W, H)
    cam_f = flatdesc.make_camera(W, H, scenes.CORNELL_C2W, scenes.CORNELL_FOV)
    assert list(cam_h.c2w) == list(cam_f.c2w) and cam_h.scale == cam_f.scale and cam_h.aspect == cam_f.aspect
    host = scenes.cornell_box("quad")
    flat = flatdesc.cornell_box()
    a, _, _ = api.ReferenceScene(host.flatten()).render(cam_h, W, H, 4, capi.INT_GI, 3)
    b, _, _ = api.ReferenceScene(flat.desc()).render(cam_f, W, H, 4, capi.INT_GI, 3)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    ht, ft = scenes.cornell_box("triangle"), flatdesc.cornell_box("triangle")   # bench.py workload c2
    a, _, _ = api.ReferenceScene(ht.flatten()).render(cam_h, W, H, 4, capi.INT_DIRECT, 1)
    b, _, _ = api.ReferenceScene(ft.desc()).render(cam_f, W, H, 4, capi.INT_DIRECT, 1)
    assert a.max() > 0 and np.array_equal(a.view(np.uint32), b.view(np.uint32))
    hv = scenes.volume_scene(n=24)
    fv = flatdesc.volume_scene(n=24)
    a, _, _ = api.ReferenceScene(hv.flatten()).render(cam_h, W, H, 4, capi.INT_VOLUME, 8)
    b, _, _ = api.ReferenceScene(fv.desc()).render(cam_f, W, H, 4, capi.INT_VOLUME, 8)
    assert a.max() > 0 and np.array_equal(a.view(np.uint32), b.view(np.uint32))
    hm = scenes.cornell_mesh_scene(12, 12)
    fm = flatdesc.cornell_mesh_scene(12, 12)
    a, _, _ = api.ReferenceScene(hm.flatten()).render(cam_h, W, H, 2, capi.INT_GI, 3)
    b, _, _ = api.ReferenceScene(fm.desc()).render(cam_f, W, H, 2, capi.INT_GI, 3)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_malformed_obj_indices_are_load_errors(tmp_path):
    """Face tokens that are 0, missing, out of range or reach before the first element used to become out-of-bounds reads in
    Scene::loadObj; they are load errors now (like the tinyobj-based reference, scene.cpp:52-64). Normals on only some vertices
    of a face fall back to the flat normal."""
    base = "v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nnewmtl m\n"
    (tmp_path / "m.mtl").write_text("newmtl m\nKd 1 1 1\n")
    head = "mtllib m.mtl\no a\nusemtl m\nv 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\n"
    for k, face in enumerate(["f 1 2 4", "f 0 1 2", "f 1 2 -4", "f 1//2 2//1 3//1", "f 1/5 2 3"]):
        path = tmp_path / f"bad{k}.obj"
        path.write_text(head + face + "\n")
        s = scenes.HostScene()
        with pytest.raises(RuntimeError, match="face references"):
            s.load_obj(path)
    good = tmp_path / "partial.obj"
    good.write_text(head + "f 1//1 2 3\n")     # a normal on one vertex only -> flat normal for the face
    s = scenes.HostScene()
    s.load_obj(good)
    d = s.flatten().contents
    assert d.n_triangles == 1
    t = d.triangles[0]
    assert list(t.n0) == list(t.n1) == list(t.n2) == [0.0, 0.0, 1.0]
