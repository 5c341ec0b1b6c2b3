"""c5 (256^3 volume): sweep of the k_volume_paths lockstep threshold / steps per vote. usage: python scripts/c5_tune.py [spp]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
from xraytracer_b200 import api, capi, scenes

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 64
wl = bench.WORKLOADS["c5"]
host = bench.build_scene(wl["scene"])
scene = api.GpuScene(host.flatten(), 0)
cam = scenes.make_camera(wl["width"], wl["height"])
W, H = wl["width"], wl["height"]
for rounds, thr, spv in ((1, 16, 3), (2, 16, 3), (3, 16, 3), (4, 16, 3), (1, 16, 2), (2, 16, 2), (2, 12, 3), (2, 20, 3), (2, 24, 3), (2, 20, 4), (3, 20, 3), (3, 24, 4), (1, 16, 3), (2, 16, 3)):
    if True:
        scene.set_tuning(thr_vol=thr, spv_vol=spv | (rounds << 8))
        best = 1e9
        for it in range(4):
            _, st = scene.render(cam, W, H, spp, capi.INT_VOLUME, wl["max_depth"], seed=1234, flags=capi.FLAG_STAGE_TIMES)
            best = min(best, st["render_ms"])
        print(f"rounds={rounds} thr_vol={thr} spv_vol={spv}: {best:.2f} ms = {W * H * spp / best / 1e3:.0f} Msamples/s", flush=True)
