"""Samples per wave A/B (c3 / c4 / c5). usage: python scripts/wave_size_ab.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
from xraytracer_b200 import api, capi, scenes

import os
BIG = os.environ.get("WAVE_AB_BIG")
for name, spp in ((("c3", 256), ("c4", 128), ("c5", 256)) if BIG else (("c3", 64), ("c4", 32), ("c5", 64))):
    wl = bench.WORKLOADS[name]
    host = bench.build_scene(wl["scene"])
    scene = api.GpuScene(host.flatten(), 0)
    cam = scenes.make_camera(wl["width"], wl["height"])
    W, H = wl["width"], wl["height"]
    integ = capi.INTEGRATOR_NAMES.index(wl["integrator"])
    for S in ((16, 32, 64, 128) if BIG else (1, 2, 4, 8, 16, 32)):
        best = 1e9
        for it in range(4):
            _, st = scene.render(cam, W, H, spp, integ, wl["max_depth"], seed=1234, samples_per_wave=S)
            best = min(best, st["render_ms"])
        print(f"{name} S={S}: {best:.2f} ms = {W * H * spp / best / 1e3:.0f} Msamples/s", flush=True)
    del scene
