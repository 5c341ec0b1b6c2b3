import sys; sys.path.insert(0,'.')
import numpy as np
from xraytracer_b200 import api, capi, scenes
for n, light, W, H in [(32,'quad',64,64),(32,'sphere',64,64),(32,'quad',640,360),(256,'quad',640,360),(256,'quad',1920,1080)]:
    s = scenes.volume_scene(n=n, light=light); d = s.flatten()
    g = api.GpuScene(d,0)
    cam = scenes.make_camera(W,H)
    a, st = g.render(cam, W,H, 2, capi.INT_VOLUME, 16, seed=1, flags=capi.FLAG_STAGE_TIMES)
    print(n, light, W, H, 'fast: closest', st['closest_rays'], 'steps', st['tracking_steps'], 'launches', st['kernel_launches'], 'mean', a.mean(), 'ms', st['render_ms'])
    if W <= 64:
        o = api.OracleScene(d)
        b, _, ost = o.render(cam, W,H, 2, capi.INT_VOLUME, 16)
        a, st = g.render(cam, W,H, 2, capi.INT_VOLUME, 16, flags=capi.FLAG_EXACT)
        print('   exact: closest', st['closest_rays'], ost['closest_rays'], 'steps', st['tracking_steps'], ost['tracking_steps'], 'maxabs', np.abs(a-b).max())
