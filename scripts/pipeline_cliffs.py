"""Throughput on both sides of every limit that switches the render pipeline (api.cu: choosePipeline):
   64 / 66 mesh triangles   fused k_bounce_small  ->  k_primary + shade + connect + extend (simple kernels, shared-memory triangle loops off)
   16 / 17 area lights      fused k_bounce_small  ->  three-kernel pipeline (one shadow-queue entry per light)
   <= 512 / > 512 BVH nodes simple run-to-completion kernels -> raygen + k_trace8 (resumable traversal of the eight-child tree)
GI depth 3, 1920x1080, 16 spp, throughput instantiation. usage: python scripts/pipeline_cliffs.py  (prints a markdown table)"""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from xraytracer_b200 import api, capi, scenes


def tri(v0, v1, v2):
    v0, v1, v2 = (np.asarray(v, np.float32) for v in (v0, v1, v2))
    n = np.cross(v1 - v0, v2 - v0).astype(np.float32)
    n /= max(float(np.linalg.norm(n)), 1e-20)
    return np.concatenate([v0, v1, v2, n, n, n]).astype(np.float32)


def quad(p, e1, e2):
    p, e1, e2 = (np.asarray(v, np.float32) for v in (p, e1, e2))
    return [tri(p, p + e1, p + e1 + e2), tri(p, p + e1 + e2, p + e2)]


def room(n_extra_quads, n_lights):
    """Closed-ish room [0,100]^3 (5 quads) + n_extra_quads small floating quads + n_lights ceiling lights."""
    s = scenes.HostScene()
    walls = [((0, 0, 0), (0, 0, 100), (100, 0, 0)), ((0, 100, 0), (100, 0, 0), (0, 0, 100)), ((0, 0, 100), (0, 100, 0), (100, 0, 0)),
             ((0, 0, 0), (0, 100, 0), (0, 0, 100)), ((100, 0, 0), (0, 0, 100), (0, 100, 0))]
    t = []
    for p, e1, e2 in walls:
        t += quad(p, e1, e2)
    rng = np.random.RandomState(1)
    for k in range(n_extra_quads):
        c = rng.uniform(15, 85, 3)
        t += quad(c, (8, 0, 2), (0, 1, 8))
    s.add_mesh("geo", np.array(t), (0.7, 0.7, 0.7))
    for k in range(n_lights):
        x = 4.0 + 5.4 * k
        s.add_quad_light(f"L{k}", (x + 4, 99.5, 40), (x + 4, 99.5, 60), (x, 99.5, 40), (30.0, 30.0, 30.0))   # faces down
    return s


def sphere_room(nt):
    extra = lambda h: h.add_mesh("tess", scenes.displaced_sphere_tris((278, 200, 280), 150, nt, nt), (0.75, 0.75, 0.75))
    return scenes.cornell_box("quad", extra=extra)


W, H, SPP = 1920, 1080, 16
rows = []
cases = [("27 quads + 1 light (64 triangles incl. light proxy)", room(26, 1), scenes.make_camera(W, H, [-1, 0, 0, 0, 0, 1, 0, 0, 0, 0, -1, 0, 50.0, 50.0, -140.0, 1], 50.0)),
         ("28 quads + 1 light (66 triangles)", room(27, 1), None), ("12 quads + 8 lights", room(7, 8), None), ("12 quads + 16 lights", room(7, 16), None), ("12 quads + 17 lights", room(7, 17), None)]
cam_room = cases[0][2]
for nt in (19, 20, 21, 22):
    cases.append((f"Cornell box + {2 * nt * nt}-triangle sphere", sphere_room(nt), scenes.make_camera(W, H)))
for name, host, cam in cases:
    cam = cam or cam_room
    g = api.GpuScene(host.flatten(), 0)
    info = g.info()
    best, st = 1e9, None
    for _ in range(3):
        _, st = g.render(cam, W, H, SPP, capi.INT_GI, 3, seed=1, flags=capi.FLAG_STAGE_TIMES)
        best = min(best, st["render_ms"])
    pipe = "fused k_bounce_small" if st["bounce_launches"] else ("k_primary + simple kernels" if st["primary_hits"] else f"raygen + k_trace ({info['wide_arity']}-child tree)")
    rows.append((name, info["n_triangles"], info["n_bvh_nodes"], pipe, W * H * SPP / best / 1e3, (st["closest_rays"] + st["shadow_rays"]) / best / 1e3))
print("| scene | triangles | BVH2 nodes | pipeline | Msamples/s | Mrays/s |\n|---|---|---|---|---|---|")
for r in rows:
    print(f"| {r[0]} | {r[1]} | {r[2]} | {r[3]} | {r[4]:.0f} | {r[5]:.0f} |")
