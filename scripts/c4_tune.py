"""c4 (999,698 triangles, device-built tree): sweep of the k_trace8 refill thresholds / steps per vote / leaf threshold.
usage: python scripts/c4_tune.py [spp]"""
import itertools, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
from xraytracer_b200 import api, capi, scenes

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 16
wl = bench.WORKLOADS["c4"]
host = bench.build_scene(wl["scene"])
scene = api.GpuScene(host.flatten(), 0)
cam = scenes.make_camera(wl["width"], wl["height"])
W, H = wl["width"], wl["height"]


def run(**kw):
    scene.set_tuning(**kw)
    best, bst = 1e9, None
    for it in range(5):
        _, st = scene.render(cam, W, H, spp, capi.INT_GI, 3, seed=1234, flags=capi.FLAG_STAGE_TIMES)
        if st["render_ms"] < best:
            best, bst = st["render_ms"], st
    print(f"{kw}: {best:.2f} ms = {W * H * spp / best / 1e3:.0f} Msamples/s (extend {bst['extend_ms']:.2f} shade {bst['shade_ms']:.2f} connect {bst['connect_ms']:.2f})", flush=True)
    return best


run()
for e0, e, c in ((1, 24, 24), (8, 24, 24), (16, 24, 24), (24, 24, 24), (1, 16, 16), (1, 20, 20), (1, 28, 28), (1, 24, 16), (1, 24, 28), (1, 16, 24), (12, 20, 24)):
    run(thr_ext0=e0, thr_ext=e, thr_con=c)
for spv in (1, 2, 3, 4):
    run(steps_per_vote=spv)
for lt in (1, 2, 4, 6, 8, 12):
    run(leaf_threshold=lt)
for spv, lt in ((1, 2), (1, 8), (3, 8), (4, 8)):
    run(steps_per_vote=spv, leaf_threshold=lt)
