import sys; sys.path.insert(0,'.')
import numpy as np, ctypes as C
from xraytracer_b200 import api, capi, scenes
s = scenes.cornell_box("quad"); desc = s.flatten()
gpu = api.GpuScene(desc,0); orc = api.OracleScene(desc)
W=H=512; spp=16
cam = scenes.make_camera(W,H)
a = gpu.trace_primary(cam,W,H,spp); b = orc.trace_primary(cam,W,H,spp)
c = gpu.trace_primary(cam,W,H,spp, flags=capi.FLAG_BRUTE_FORCE)
bad = np.argwhere(a['prim']!=b['prim'])
print('bvh mism', len(bad), 'brute mism', (c['prim']!=b['prim']).sum())
# reconstruct the ray with the oracle's sampler + camera kat
o = capi.oracle()
for (i,j,k) in bad:
    seq = api.kat(o,'xrto_','sampler', int(j+W*i), 2*spp, n_out=2*spp)
    r0, r1 = seq[2*k], seq[2*k+1]
    u = np.float32(np.float32(j)+r0)/np.float32(W); v = np.float32(np.float32(i)+r1)/np.float32(H)
    ray = api.kat(o,'xrto_','camera', C.byref(cam), C.c_float(u), C.c_float(v), n_out=6)
    print('pixel',i,j,k,'xi',r0.hex(),r1.hex(),'u,v',u,v,'ray',[float(x).hex() for x in ray], ray)
    org = ray[:3][None]; d = ray[3:][None]
    print('  trace_rays bvh', gpu.trace_rays(org,d), 'brute', gpu.trace_rays(org,d,flags=capi.FLAG_BRUTE_FORCE), 'oracle', orc.trace_rays(org,d))
