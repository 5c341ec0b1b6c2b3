import sys, time
sys.path.insert(0, '/root/repo')
import bench
from xraytracer_b200 import api, capi, scenes
wl = bench.WORKLOADS["c4"]
host = bench.build_scene(wl["scene"]); desc = host.flatten()
cam = scenes.make_camera(1920, 1080)
# mimic the bench: a c3 scene with a big workspace first
h3 = bench.build_scene("cornell"); s3 = api.GpuScene(h3.flatten(), 0)
s3.render(cam, 1920, 1080, 64, capi.INT_GI, 3, seed=1)
del s3
for k in range(6):
    t0 = time.perf_counter(); s = api.GpuScene(desc, 0); t1 = time.perf_counter()
    i = s.info(); print(f"create {1e3*(t1-t0):.1f} ms, build_ms {i['build_ms']:.1f}, bvh {i['bvh_build_ms']:.1f}", flush=True)
    if k % 2 == 0:
        s.render(cam, 1920, 1080, 32, capi.INT_GI, 3, seed=1)
    del s
