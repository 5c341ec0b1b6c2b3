"""Prints a hash of a few throughput-mode renders (development aid: two builds that claim to draw the same samples must print the
same lines). usage: python scripts/dev_image_hash.py"""
import hashlib, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from xraytracer_b200 import api, capi, scenes

host = scenes.volume_scene(n=48, light="quad")
gpu = api.GpuScene(host.flatten(), 0)
cam = scenes.make_camera(320, 180)
for integ in (capi.INT_VOLUME, capi.INT_VOLUME_NEE):
    img, st = gpu.render(cam, 320, 180, 16, integ, 12, seed=7)
    print("volume", integ, hashlib.sha1(img.tobytes()).hexdigest()[:16], st["tracking_steps"], st["closest_rays"], float(img.mean()))
box = scenes.cornell_box("quad")
g2 = api.GpuScene(box.flatten(), 0)
img, st = g2.render(cam, 320, 180, 16, capi.INT_GI, 3, seed=7)
print("cornell gi", hashlib.sha1(img.tobytes()).hexdigest()[:16], st["closest_rays"], st["shadow_rays"], float(img.mean()))
