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
def test_random_small_scene_fused_vs_three_kernel_and_oracle(seed, monkeypatch):
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
        monkeypatch.setenv("XRT_FUSED_BOUNCE", "1")
        f, sf = gpu.render(cam, W, H, 16, integ, depth, seed=seed)
        monkeypatch.setenv("XRT_FUSED_BOUNCE", "0")
        u, su = gpu.render(cam, W, H, 16, integ, depth, seed=seed)
        monkeypatch.delenv("XRT_FUSED_BOUNCE")
        assert abs(sf["closest_rays"] - su["closest_rays"]) <= 2e-4 * su["closest_rays"] + 2, (seed, integ)
        assert abs(sf["shadow_rays"] - su["shadow_rays"]) <= 2e-4 * su["shadow_rays"] + 2, (seed, integ)
        assert abs(float(f.mean()) - float(u.mean())) <= 1e-3 * float(u.mean()) + 1e-6, (seed, integ)
        differing = np.abs(f - u).max(axis=-1) > 1e-3 * (1.0 + np.abs(u).max(axis=-1))
        assert differing.mean() < 0.01, (seed, integ)
