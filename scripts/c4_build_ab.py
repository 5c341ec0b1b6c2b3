"""c4 (999,698 triangles): host SAH build vs GPU LBVH build, each traversed as the eight- / four-child tree.
usage: python scripts/c4_build_ab.py [spp]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
from xraytracer_b200 import api, capi, scenes

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 16
wl = bench.WORKLOADS["c4"]
host = bench.build_scene(wl["scene"])
desc = host.flatten()
cam = scenes.make_camera(wl["width"], wl["height"])
for name, flags in (("host SAH", 0), ("GPU LBVH", capi.BUILD_LBVH_GPU)):
    t0 = time.perf_counter()
    scene = api.GpuScene(desc, 0, build_flags=flags)
    t1 = time.perf_counter()
    info = scene.info()
    for arity in (8, 4):
        scene.set_tuning(wide_bvh=arity)
        best = 1e9
        for it in range(3):
            img, st = scene.render(cam, wl["width"], wl["height"], spp, capi.INT_GI, 3, seed=1234, flags=capi.FLAG_STAGE_TIMES)
            best = min(best, st["render_ms"])
        cst = scene.render(cam, wl["width"], wl["height"], 2, capi.INT_GI, 3, seed=1234, flags=capi.FLAG_COUNTERS)[1]
        print(f"{name}: create {1e3 * (t1 - t0):.0f} ms (ingest {info['build_ms']:.0f}, bvh {info['bvh_build_ms']:.0f}, upload {info['upload_ms']:.0f}), arity {arity}: "
              f"{best:.2f} ms / {spp} spp = {wl['width'] * wl['height'] * spp / best / 1e3:.0f} Msamples/s, nodes/ray {cst['nodes_visited'] / cst['closest_rays']:.2f}, "
              f"tris/ray {cst['tris_tested'] / cst['closest_rays']:.2f}, shadow nodes/ray {cst['nodes_visited_shadow'] / max(cst['shadow_rays'], 1):.2f}, wide nodes {info['n_wide_nodes']}, depth {info['bvh_depth']}")
