"""Build time and traversal speed of the GPU LBVH vs the host SAH tree on workload c4 (not a test)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
from xraytracer_b200 import api, capi, scenes
host = bench.build_scene("mesh1m"); desc = host.flatten()
cam = scenes.make_camera(1920, 1080)
for name, flags in (("host SAH", 0), ("GPU LBVH", capi.BUILD_LBVH_GPU)):
    t = time.time(); g = api.GpuScene(desc, 0, build_flags=flags); wall = time.time() - t
    t = time.time(); g2 = api.GpuScene(desc, 0, build_flags=flags); wall2 = time.time() - t
    info = g2.info()
    for _ in range(2):
        img, st = g2.render(cam, 1920, 1080, 8, capi.INT_GI, 3, seed=1, flags=capi.FLAG_STAGE_TIMES)
    _, cst = g2.render(cam, 1920, 1080, 2, capi.INT_GI, 3, seed=1, flags=capi.FLAG_COUNTERS)
    print(f"{name}: scene_create {wall2*1e3:.0f} ms wall (first {wall*1e3:.0f}), ingest_ms {info['build_ms']:.1f}, bvh_build_ms {info['bvh_build_ms']:.1f}, nodes {info['n_bvh_nodes']}, depth {info['bvh_depth']}; "
          f"render 8spp {st['render_ms']:.2f} ms (extend {st['extend_ms']:.2f} connect {st['connect_ms']:.2f}); "
          f"nodes/ray {cst['nodes_visited']/cst['closest_rays']:.1f} tris/ray {cst['tris_tested']/cst['closest_rays']:.1f}")
    del g, g2
