"""Short, torch-free driver for ncu captures: a few waves of one workload through the C ABI.
usage: python scripts/profile_step.py c3|c4|c5 [spp]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
from xraytracer_b200 import api, capi, scenes

wl = dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3"])
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 8
host = bench.build_scene(wl["scene"])
scene = api.GpuScene(host.flatten(), 0)
cam = scenes.make_camera(wl["width"], wl["height"])
integ = capi.INTEGRATOR_NAMES.index(wl["integrator"])
for it in range(3):
    img, st = scene.render(cam, wl["width"], wl["height"], spp, integ, wl["max_depth"], seed=1234, flags=capi.FLAG_STAGE_TIMES)
    print(f"iter {it}: {st['render_ms']:.2f} ms, {wl['width'] * wl['height'] * spp / st['render_ms'] / 1e3:.1f} Msamples/s, extend {st['extend_ms']:.2f} "
          f"shade {st['shade_ms']:.2f} connect {st['connect_ms']:.2f} launches {st['kernel_launches']}")
