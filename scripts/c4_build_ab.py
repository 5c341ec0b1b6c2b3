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
import os
variants = [("host SAH", capi.BUILD_HOST, ""), ("GPU LBVH", capi.BUILD_LBVH_GPU, "")]
for r, ct, top, w in ((16, 16, 1, 1), (16, 16, 4, 1), (16, 16, 8, 1), (16, 16, 16, 1), (16, 16, 32, 1), (16, 16, 128, 1), (16, 16, 8, 0), (16, 16, 16, 0), (16, 16, 32, 0),
                      (16, 16, 1024, 1), (24, 16, 16, 1), (12, 16, 16, 1)):
    variants.append((f"GPU PLOC r={r} ct={ct / 16:g} top={top} w={w}", capi.BUILD_GPU, f"ploc_radius={r},ploc_ct_x16={ct},ploc_top={top},ploc_weight={w}"))
only = os.environ.get("AB_ONLY")
for name, flags, tuning in variants:
    if only and only not in name:
        continue
    base = os.environ.get("XRT_TUNING_BASE", "")
    os.environ["XRT_TUNING"] = ",".join(x for x in (base, tuning) if x)
    t0 = time.perf_counter()
    scene = api.GpuScene(desc, 0, build_flags=flags)
    t1 = time.perf_counter()
    info = scene.info()
    for arity in ((8,) if flags == capi.BUILD_GPU else (8, 4)):
        scene.set_tuning(wide_bvh=arity)
        best = 1e9
        for it in range(6):
            img, st = scene.render(cam, wl["width"], wl["height"], spp, capi.INT_GI, 3, seed=1234, flags=capi.FLAG_STAGE_TIMES)
            best = min(best, st["render_ms"])
        cst = scene.render(cam, wl["width"], wl["height"], 2, capi.INT_GI, 3, seed=1234, flags=capi.FLAG_COUNTERS)[1]
        print(f"{name}: create {1e3 * (t1 - t0):.0f} ms (ingest {info['build_ms']:.0f}, bvh {info['bvh_build_ms']:.0f}, upload {info['upload_ms']:.0f}), arity {arity}: "
              f"{best:.2f} ms / {spp} spp = {wl['width'] * wl['height'] * spp / best / 1e3:.0f} Msamples/s, nodes/ray {cst['nodes_visited'] / cst['closest_rays']:.2f}, "
              f"tris/ray {cst['tris_tested'] / cst['closest_rays']:.2f}, shadow nodes/ray {cst['nodes_visited_shadow'] / max(cst['shadow_rays'], 1):.2f}, wide nodes {info['n_wide_nodes']}, depth {info['bvh_depth']}, SAH {info['bvh_sah_cost']:.1f}", flush=True)
