"""c4: any-hit / closest-hit overlap on two streams, A/B at several wave sizes. usage: python scripts/c4_overlap_ab.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
from xraytracer_b200 import api, capi, scenes

wl = bench.WORKLOADS["c4"]
host = bench.build_scene(wl["scene"])
scene = api.GpuScene(host.flatten(), 0)
cam = scenes.make_camera(wl["width"], wl["height"])
W, H = wl["width"], wl["height"]
for S, spp in ((32, 64), (8, 32), (4, 16)):
    for ov in (0, 1, 0, 1):
        scene.set_tuning(overlap_connect=ov)
        best = 1e9
        for it in range(5):
            img, st = scene.render(cam, W, H, spp, capi.INT_GI, 3, seed=1234, samples_per_wave=S)
            best = min(best, st["render_ms"])
        print(f"S={S} spp={spp} overlap={ov}: {best:.2f} ms = {W * H * spp / best / 1e3:.0f} Msamples/s mean {float(img.mean()):.6f}", flush=True)
