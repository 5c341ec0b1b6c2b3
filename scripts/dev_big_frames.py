import sys; sys.path.insert(0, '/root/repo')
import numpy as np, bench
from xraytracer_b200 import api, capi, scenes
for name, (W, H, spp) in (("c3", (3840, 2160, 40)), ("c4", (3840, 2160, 12)), ("c5", (2560, 1440, 48)), ("c3", (1921, 1081, 33))):
    wl = bench.WORKLOADS[name]
    h = bench.build_scene(wl["scene"]); s = api.GpuScene(h.flatten(), 0); cam = scenes.make_camera(W, H)
    integ = capi.INTEGRATOR_NAMES.index(wl["integrator"])
    img, st = s.render(cam, W, H, spp, integ, wl["max_depth"], seed=5)
    g = s.check_guards()
    # reference: the same render in waves of 4 samples (the round-1 wave size): identical sample set, fp32 re-association only
    img2, st2 = s.render(cam, W, H, spp, integ, wl["max_depth"], seed=5, samples_per_wave=4)
    print(name, W, H, spp, f"{W*H*spp/st['render_ms']/1e3:.0f} Msamples/s", "guards", g, "mean", float(img.mean()), "max|diff| vs S=4", float(np.abs(img-img2).max()),
          "rays equal", st["closest_rays"] == st2["closest_rays"] and st["shadow_rays"] == st2["shadow_rays"], "finite", bool(np.isfinite(img).all()), flush=True)
    del s
