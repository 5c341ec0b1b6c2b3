"""Randomised differential tests on small scenes (<= 64 triangles): quads (coplanar pairs), lone triangles, coplanar duplicates,
a degenerate triangle, analytic spheres, one or two area lights. Every scene is rendered by
  * the fused per-bounce kernel (plane-paired records in the throughput build) and the three-kernel pipeline, same seed;
  * the exact instantiation and the CPU oracle (libxrtoracle.so), bit-level comparison of the sample stream's image."""
import numpy as np
import pytest

from conftest import require_gpu
from xraytracer_b200 import api, capi, scenes

pytestmark = pytest.mark.gpu


def _tri(v0, v1, v2):
    v0, v1, v2 = (np.asarray(v, np.float32) for v in (v0, v1, v2))
    n = np.cross(v1 - v0, v2 - v0).astype(np.float32)
    ln = float(np.linalg.norm(n))
    n = n / ln if ln > 0 else np.array([0, 1, 0], np.float32)
    return np.concatenate([v0, v1, v2, n, n, n]).astype(np.float32)


def _quad(p, e1, e2):
    p, e1, e2 = (np.asarray(v, np.float32) for v in (p, e1, e2))
    return [_tri(p, p + e1, p + e1 + e2), _tri(p, p + e1 + e2, p + e2)]


def random_scene(seed):
    rng = np.random.default_rng(seed)
    s = scenes.HostScene()
    # an open room of quads around [0,100]^3 seen from -z, random albedos
    room = [((0, 0, 0), (100, 0, 0), (0, 0, 100)), ((0, 100, 0), (0, 0, 100), (100, 0, 0)), ((0, 0, 100), (100, 0, 0), (0, 100, 0)),
            ((0, 0, 0), (0, 0, 100), (0, 100, 0)), ((100, 0, 0), (0, 100, 0), (0, 0, 100))]
    for k, (p, e1, e2) in enumerate(room):
        s.add_mesh(f"wall{k}", np.array(_quad(p, e1, e2)), rng.uniform(0.2, 0.9, 3))
    # random tilted quads, lone triangles, a duplicate of one quad (coplanar, overlapping) and a zero-area triangle
    extra = []
    for k in range(int(rng.integers(2, 6))):
        p = rng.uniform(15, 70, 3); e1 = rng.uniform(-25, 25, 3); e2 = np.cross(e1, rng.uniform(-1, 1, 3)); e2 *= 20 / max(np.linalg.norm(e2), 1e-3)
        extra.append(_quad(p, e1, e2))
        s.add_mesh(f"quad{k}", np.array(extra[-1]), rng.uniform(0.2, 0.9, 3))
    for k in range(int(rng.integers(1, 4))):
        s.add_mesh(f"tri{k}", np.array([_tri(rng.uniform(10, 90, 3), rng.uniform(10, 90, 3), rng.uniform(10, 90, 3))]), rng.uniform(0.2, 0.9, 3))
    s.add_mesh("dup", np.array(extra[0]), (0.9, 0.1, 0.1))
    s.add_mesh("degenerate", np.array([_tri((50, 50, 50), (50, 50, 50), (60, 50, 50))]), (0.5, 0.5, 0.5))
    if seed % 2:
        s.add_sphere("ball", rng.uniform(30, 70, 3), 8.0, (0.3, 0.8, 0.3))
    s.add_quad_light("QuadLight", (35, 99.5, 35), (35, 99.5, 65), (65, 99.5, 35), (40.0, 40.0, 40.0))
    if seed % 3 == 0:
        s.add_sphere_light("SphereLight", (20, 60, 50), 4.0, (60.0, 40.0, 20.0))
    cam = scenes.make_camera(160, 120, [-1, 0, 0, 0, 0, 1, 0, 0, 0, 0, -1, 0, 50.0, 50.0, -140.0, 1], 50.0)
    return s, cam


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_random_small_scene_fused_vs_three_kernel_and_oracle(seed):
    require_gpu()
    host, cam = random_scene(seed)
    desc = host.flatten()
    gpu, orc = api.GpuScene(desc, 0), api.OracleScene(desc)
    assert 0 < gpu.info()["n_triangles"] <= 64
    W, H = 160, 120
    # primary hits: BVH == oracle, bit for bit
    assert np.array_equal(gpu.trace_primary(cam, W, H, 1), orc.trace_primary(cam, W, H, 1))
    for integ, depth in ((capi.INT_GI, 3), (capi.INT_DIRECT, 1), (capi.INT_INDIRECT, 2)):
        # exact instantiation (fused kernel over the per-triangle list) against the CPU oracle
        # (odd seeds have a Lambert SPHERE: Sphere::intersect leaves dpdu/dpdv stale in the reference — SURVEY §9-T4 — and the exact
        #  instantiation reproduces which earlier mesh hit they come from, so even these bounce identically)
        a, sa = gpu.render(cam, W, H, 2, integ, depth, flags=capi.FLAG_EXACT)
        b, _, sb = orc.render(cam, W, H, 2, integ, depth)
        assert sa["closest_rays"] == sb["closest_rays"] and sa["shadow_rays"] == sb["shadow_rays"], (seed, integ)
        assert np.abs(a - b).max() <= 2e-5 * max(1.0, float(np.abs(b).max())), (seed, integ)
        # throughput instantiation: plane-paired fused kernel against the three-kernel pipeline, same seed
        gpu.set_tuning(fused_bounce=1)
        f, sf = gpu.render(cam, W, H, 16, integ, depth, seed=seed)
        gpu.set_tuning(fused_bounce=0)
        u, su = gpu.render(cam, W, H, 16, integ, depth, seed=seed)
        gpu.set_tuning()
        assert abs(sf["closest_rays"] - su["closest_rays"]) <= 2e-4 * su["closest_rays"] + 2, (seed, integ)
        assert abs(sf["shadow_rays"] - su["shadow_rays"]) <= 2e-4 * su["shadow_rays"] + 2, (seed, integ)
        assert abs(float(f.mean()) - float(u.mean())) <= 1e-3 * float(u.mean()) + 1e-6, (seed, integ)
        differing = np.abs(f - u).max(axis=-1) > 1e-3 * (1.0 + np.abs(u).max(axis=-1))
        assert differing.mean() < 0.01, (seed, integ)


def _grid_scene(n_quads, n_lights=1):
    """n_quads floor tiles (2 triangles each) under `n_lights` small quad lights."""
    s = scenes.HostScene()
    side = int(np.ceil(np.sqrt(n_quads)))
    tris = []
    for k in range(n_quads):
        x, z = (k % side) * 10.0, (k // side) * 10.0
        tris += _quad((x, 0.0, z), (0, 0, 9.5), (9.5, 0, 0))
    s.add_mesh("floor", np.array(tris), (0.7, 0.7, 0.7))
    s.add_mesh("wall", np.array(_quad((0, 0, side * 10.0), (side * 10.0, 0, 0), (0, 40, 0))), (0.8, 0.3, 0.3))
    for k in range(n_lights):
        x = 2.0 + 5.5 * k
        s.add_quad_light(f"QuadLight{k}", (x + 4.0, 30.0, 5.0), (x + 4.0, 30.0, 9.0), (x, 30.0, 5.0), (50.0, 50.0, 50.0))   # faces down
    half = side * 5.0
    cam = scenes.make_camera(128, 96, [-1, 0, 0, 0, 0, 0.8, 0.6, 0, 0, 0.6, -0.8, 0, half, 60.0, -40.0, 1], 60.0)   # looks forward and down
    return s, cam


@pytest.mark.parametrize("n_quads,n_lights", [(30, 1), (31, 1), (32, 1), (12, 17), (12, 16), (12, 2)])
def test_pipeline_selection_boundaries(n_quads, n_lights):
    """Scenes just below / above the limits that select the fused small-scene kernel (64 triangles, 16 lights): whichever pipeline
    runs, the exact instantiation reproduces the oracle and the throughput one converges to the same mean."""
    require_gpu()
    host, cam = _grid_scene(n_quads, n_lights)
    desc = host.flatten()
    gpu, orc = api.GpuScene(desc, 0), api.OracleScene(desc)
    W, H = 128, 96
    assert np.array_equal(gpu.trace_primary(cam, W, H, 1), orc.trace_primary(cam, W, H, 1))
    a, sa = gpu.render(cam, W, H, 2, capi.INT_GI, 3, flags=capi.FLAG_EXACT)
    b, _, sb = orc.render(cam, W, H, 2, capi.INT_GI, 3)
    assert sa["closest_rays"] == sb["closest_rays"] and sa["shadow_rays"] == sb["shadow_rays"]
    assert np.abs(a - b).max() <= 2e-5 * max(1.0, float(np.abs(b).max()))
    f, sf = gpu.render(cam, W, H, 4096, capi.INT_GI, 3, seed=3)
    r, _, _ = orc.render(cam, W, H, 512, capi.INT_GI, 3)
    assert float(r.mean()) > 1e-3 and abs(float(f.mean()) - float(r.mean())) < 0.005 * float(r.mean())
    fused = sf["bounce_launches"] > 0
    assert fused == (gpu.info()["n_triangles"] <= 64 and n_lights <= 16)


def test_degenerate_views():
    """1x1 image, a camera that looks away from the scene (every pixel outside the scissor), one sample per wave."""
    require_gpu()
    host = scenes.cornell_box("quad")
    gpu = api.GpuScene(host.flatten(), 0)
    img, st = gpu.render(scenes.make_camera(1, 1), 1, 1, 8, capi.INT_GI, 3, seed=1)
    assert img.shape == (1, 1, 3) and np.isfinite(img).all() and st["samples"] == 8
    away = scenes.make_camera(64, 48, [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 278.0, 274.4, -750.0, 1], 60.0)   # looks down -z: away from the box
    for integ, colour in ((capi.INT_GI, 0.0), (capi.INT_DIRECT, 0.18)):
        img, st = gpu.render(away, 64, 48, 4, integ, 3, seed=1)
        assert np.allclose(img, colour) and st["primary_hits"] == 0
    a, _ = gpu.render(scenes.make_camera(96, 54), 96, 54, 6, capi.INT_GI, 3, seed=5, samples_per_wave=1)
    b, _ = gpu.render(scenes.make_camera(96, 54), 96, 54, 6, capi.INT_GI, 3, seed=5, samples_per_wave=4)
    assert np.allclose(a, b, rtol=1e-5, atol=1e-6)


def test_hundreds_of_area_lights_render_in_budgeted_waves():
    """An emissive mesh = one TriangleLight per triangle. The three-kernel pipeline queues one shadow ray per (path, light): the
    wave size follows a byte budget (and shrinks to ranges of pixels when even one sample of every pixel does not fit) instead
    of the allocation failing. Same image whatever the budget; exact mode still reproduces the oracle."""
    require_gpu()
    s = scenes.HostScene()
    s.add_mesh("floor", np.array(_quad((0, 0, 0), (0, 0, 100), (100, 0, 0))), (0.7, 0.7, 0.7))
    s.add_mesh("slab", np.array(_quad((30, 20, 30), (0, 0, 30), (30, 0, 0))), (0.3, 0.6, 0.9))
    n_lights = 0
    for k in range(12):
        for j in range(12):
            x, z = 8.0 * k + 2.0, 8.0 * j + 2.0
            s.add_triangle_light(f"L{k}_{j}", (x + 4.0, 60.0, z), (x, 60.0, z + 4.0), (x, 60.0, z), (30.0, 30.0, 30.0))   # faces down
            n_lights += 1
    desc = s.flatten()
    gpu, orc = api.GpuScene(desc, 0), api.OracleScene(desc)
    W, H = 96, 72
    cam = scenes.make_camera(W, H, [-1, 0, 0, 0, 0, 0.8, 0.6, 0, 0, 0.6, -0.8, 0, 50.0, 70.0, -60.0, 1], 60.0)
    a, sa = gpu.render(cam, W, H, 2, capi.INT_GI, 2, flags=capi.FLAG_EXACT)
    b, _, sb = orc.render(cam, W, H, 2, capi.INT_GI, 2)
    assert sa["shadow_rays"] == sb["shadow_rays"] > W * H * n_lights // 8
    assert np.abs(a - b).max() <= 5e-5 * max(1.0, float(np.abs(b).max()))
    base, s0 = gpu.render(cam, W, H, 8, capi.INT_GI, 2, seed=4)
    for mb in (64, 4, 1):   # 4 MB / 1 MB: not even one sample of every pixel fits -> pixel-tiled waves
        gpu.set_tuning(workspace_mb=mb)
        alt, s1 = gpu.render(cam, W, H, 8, capi.INT_GI, 2, seed=4)
        assert s1["shadow_rays"] == s0["shadow_rays"] and s1["closest_rays"] == s0["closest_rays"]
        assert np.allclose(alt, base, rtol=2e-5, atol=1e-5), mb     # (shadow contributions are added with float atomics)
        if mb <= 4:
            assert s1["kernel_launches"] > s0["kernel_launches"]
    gpu.set_tuning()
    e, _ = gpu.render(cam, W, H, 2, capi.INT_DIRECT, 1, flags=capi.FLAG_EXACT)
    gpu.set_tuning(workspace_mb=1)
    e2, _ = gpu.render(cam, W, H, 2, capi.INT_DIRECT, 1, flags=capi.FLAG_EXACT)   # exact mode, pixel-tiled
    assert np.allclose(e, e2, rtol=1e-5, atol=1e-6)


def test_render_u8_keeps_the_float_image_on_the_device(cornell_scene=None):
    require_gpu()
    host = scenes.cornell_box("quad")
    gpu = api.GpuScene(host.flatten(), 0)
    cam = scenes.make_camera(96, 64)
    img, _ = gpu.render(cam, 96, 64, 16, capi.INT_GI, 3, seed=1)
    for gamma, bgr in ((0.0, False), (1.2, True)):
        u8, st = gpu.render_u8(cam, 96, 64, 16, capi.INT_GI, 3, gamma=gamma, bgr=bgr, seed=1)
        assert np.array_equal(u8, api.image_to_u8(img, gamma, bgr)) and st["samples"] == 96 * 64 * 16
