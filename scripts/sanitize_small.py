"""Small run of every kernel (both instantiations, all integrators, parity hooks) for compute-sanitizer."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import numpy as np
from golden_cases import CASES, build_case
from xraytracer_b200 import api, capi, scenes

for name, case in CASES.items():
    host, cam = build_case(case)
    gpu = api.GpuScene(host.flatten(), 0)
    for integ, depth, spp in case["renders"]:
        for flags in (capi.FLAG_EXACT, 0, capi.FLAG_COUNTERS, capi.FLAG_BRUTE_FORCE):
            img, st = gpu.render(cam, case["w"], case["h"], 2, integ, depth, flags=flags, seed=1)
            assert np.isfinite(img).all()
    gpu.trace_primary(cam, case["w"], case["h"], 2)
    print("ok", name)
s = scenes.cornell_box("quad", extra=lambda h: h.add_mesh("tess", scenes.displaced_sphere_tris((278, 200, 280), 150, 40, 40), (0.7, 0.7, 0.7)))
gpu = api.GpuScene(s.flatten(), 0)
cam = scenes.make_camera(48, 32)
for flags in (capi.FLAG_EXACT, 0):
    gpu.render(cam, 48, 32, 2, capi.INT_GI, 3, flags=flags)
rng = np.random.RandomState(1)
o = rng.uniform(50, 500, (2000, 3)).astype(np.float32); d = rng.normal(size=(2000, 3)).astype(np.float32)
gpu.trace_rays(o, d); gpu.trace_rays(o, d, np.full(2000, 300, np.float32), any_hit=True)
print("ok deep")
