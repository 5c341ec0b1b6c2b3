"""Verbose development check run on the GPU box (not a test): parity of every integrator vs the oracle."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from xraytracer_b200 import api, capi, scenes

def cmp(name, a, b):
    eq = np.array_equal(a.view(np.uint32), b.view(np.uint32))
    d = np.abs(a - b)
    rel = np.sqrt(((a - b) ** 2).mean()) / max(float(np.sqrt((b ** 2).mean())), 1e-20)
    print(f"  {name}: bit-equal={eq} maxabs={d.max():.3e} relRMSE={rel:.3e} frac(>1e-4)={(d > 1e-4).mean():.2e} mean gpu={a.mean():.6f} ref={b.mean():.6f}")

W, H = 128, 96
host = scenes.cornell_box("quad")
desc = host.flatten()
cam = scenes.make_camera(W, H)
t = time.time(); gpu = api.GpuScene(desc, 0); print("scene create %.3fs" % (time.time() - t), gpu.info())
orc = api.OracleScene(desc)

ha = gpu.trace_primary(cam, W, H, 4); hb = orc.trace_primary(cam, W, H, 4)
print("primary (mt jitter): equal", np.array_equal(ha, hb), "prim equal", np.array_equal(ha['prim'], hb['prim']), "t equal", np.array_equal(ha['t'], hb['t']))
hc = gpu.trace_primary(cam, W, H, 4, flags=capi.FLAG_BRUTE_FORCE)
print("primary brute: equal", np.array_equal(hc, hb))
if not np.array_equal(ha, hb):
    bad = np.argwhere(ha['prim'] != hb['prim'])
    print("  mismatches:", len(bad), bad[:5], ha[tuple(bad[0])] if len(bad) else None, hb[tuple(bad[0])] if len(bad) else None)

for integ, depth in [(0, 1), (1, 1), (2, 1), (3, 3), (4, 3), (5, 3)]:
    a, st = gpu.render(cam, W, H, 8, integ, depth, flags=capi.FLAG_EXACT | capi.FLAG_COUNTERS)
    b, sec, ost = orc.render(cam, W, H, 8, integ, depth)
    print(capi.INTEGRATOR_NAMES[integ], "gpu rays", st['closest_rays'], st['shadow_rays'], "drop", st['dropped_samples'], "| oracle", ost['closest_rays'], ost['shadow_rays'], ost['dropped_samples'], "| ms %.2f" % st['render_ms'])
    cmp("exact", a, b)
    a2, st2 = gpu.render(cam, W, H, 64, integ, depth, seed=1)
    b2, _, _ = orc.render(cam, W, H, 64, integ, depth)
    cmp("fast64", a2, b2)

# volume
for kind in range(3):
    s = scenes.vpt_scene(kind); d = s.flatten()
    g = api.GpuScene(d, 0); o = api.OracleScene(d)
    c = scenes.make_camera(64, 64, scenes.VPT_C2W, scenes.VPT_FOV)
    for integ in (6, 7):
        a, st = g.render(c, 64, 64, 8, integ, 10, flags=capi.FLAG_EXACT)
        b, _, ost = o.render(c, 64, 64, 8, integ, 10)
        print("homog", kind, capi.INTEGRATOR_NAMES[integ], st['closest_rays'], ost['closest_rays'], st['tracking_steps'], ost['tracking_steps'])
        cmp("exact", a, b)
s = scenes.volume_scene(n=32, abs_color=(0.01, 0.02, 0.03), scat_color=(0.05, 0.04, 0.03), g=0.3); d = s.flatten()
g = api.GpuScene(d, 0); o = api.OracleScene(d)
c = scenes.make_camera(64, 64)
for integ in (6, 7):
    a, st = g.render(c, 64, 64, 8, integ, 16, flags=capi.FLAG_EXACT)
    b, _, ost = o.render(c, 64, 64, 8, integ, 16)
    print("hetero", capi.INTEGRATOR_NAMES[integ], st['closest_rays'], ost['closest_rays'], st['tracking_steps'], ost['tracking_steps'])
    cmp("exact", a, b)
    a, st = g.render(c, 64, 64, 256, integ, 16, seed=3)
    b, _, ost = o.render(c, 64, 64, 256, integ, 16)
    cmp("fast256", a, b)

# throughput taste
W, H = 1920, 1080
cam = scenes.make_camera(W, H)
for spp in (16, 64):
    a, st = gpu.render(cam, W, H, spp, capi.INT_GI, 3, seed=1, flags=capi.FLAG_COUNTERS)
    a, st = gpu.render(cam, W, H, spp, capi.INT_GI, 3, seed=1, flags=capi.FLAG_COUNTERS)
    ms = st['render_ms']
    print(f"GI 1080p {spp}spp: {ms:.1f} ms  {W*H*spp/ms/1e3:.1f} Msamples/s  {(st['closest_rays']+st['shadow_rays'])/ms/1e3:.1f} Mrays/s  extend {st['extend_ms']:.1f} shade {st['shade_ms']:.1f} connect {st['connect_ms']:.1f} other {st['other_ms']:.1f} nodes/ray {st['nodes_visited']/(st['closest_rays']+st['shadow_rays']):.1f} tris/ray {st['tris_tested']/(st['closest_rays']+st['shadow_rays']):.1f}")
a, st = gpu.render(cam, W, H, 64, capi.INT_GI, 3, seed=1)
print(f"GI 1080p 64spp no counters: {st['render_ms']:.1f} ms {W*H*64/st['render_ms']/1e3:.1f} Msamples/s")
