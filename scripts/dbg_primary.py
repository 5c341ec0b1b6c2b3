import sys; sys.path.insert(0,'.')
import numpy as np
from xraytracer_b200 import api, capi, scenes
s = scenes.cornell_box("quad"); desc = s.flatten()
gpu = api.GpuScene(desc,0); orc = api.OracleScene(desc)
for (W,H,spp) in [(128,96,4),(256,256,8),(512,512,4),(512,512,16)]:
    cam = scenes.make_camera(W,H)
    a = gpu.trace_primary(cam,W,H,spp); b = orc.trace_primary(cam,W,H,spp)
    bad = np.argwhere(a['prim']!=b['prim'])
    badt = np.argwhere(a['t'].view(np.uint32)!=b['t'].view(np.uint32))
    print(W,H,spp,'prim mism',len(bad),'t mism',len(badt), 'of', a.size)
    for idx in bad[:6]:
        print('   ', idx, a[tuple(idx)], b[tuple(idx)])
    if len(bad):
        print('   rows', np.unique(bad[:,0])[:20], 'samples', np.unique(bad[:,2]))
    jit = np.random.RandomState(3).random_sample((H*W*spp,2)).astype(np.float32)
    a = gpu.trace_primary(cam,W,H,spp,jitter=jit); b = orc.trace_primary(cam,W,H,spp,jitter=jit)
    print('   supplied jitter: prim mism', (a['prim']!=b['prim']).sum(), 't mism', (a['t'].view(np.uint32)!=b['t'].view(np.uint32)).sum())
