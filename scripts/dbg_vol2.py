import sys; sys.path.insert(0,'.')
import numpy as np
from xraytracer_b200 import api, capi, scenes
s = scenes.volume_scene(n=32, light='quad'); d = s.flatten()
g = api.GpuScene(d,0)
for W,H,spp,spw,fl in [(1280,720,2,0,0),(1600,900,2,0,0),(1920,1080,1,0,0),(1920,1080,2,1,0),(1920,1080,2,0,0),(1920,1080,2,0,capi.FLAG_EXACT),(2048,1024,1,0,0),(2048,1025,1,0,0)]:
    cam = scenes.make_camera(W,H)
    a, st = g.render(cam, W,H, spp, capi.INT_VOLUME, 16, seed=1, flags=fl, samples_per_wave=spw)
    print(W,H,spp,spw,fl, 'paths/wave', W*H*(spw or spp), 'closest', st['closest_rays'], 'steps', st['tracking_steps'], 'launches', st['kernel_launches'], 'mean %.5f'%a.mean())
hits = g.trace_primary(scenes.make_camera(1920,1080),1920,1080,1)
print('primary hit frac at 1080p', (hits['prim']>=0).mean())
