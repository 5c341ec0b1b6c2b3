"""Generates tests/golden/reference_vectors.npz from the COMPILED REFERENCE (oracle/_ref/libxrtref.so, built in place
from /root/reference/Src by oracle/Makefile). The reference ships no tests or golden vectors of its own,
so these outputs of the reference itself are what pins the oracle port (tests/test_golden.py) wherever
/root/reference is absent (the GPU box).

    python tests/golden/make_golden.py        # run in the build container (needs /root/reference)
"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from xraytracer_b200 import api, capi, scenes  # noqa: E402

sys.path.insert(0, str(ROOT / "tests"))
from golden_cases import CASES, KAT_SEEDS, build_case  # noqa: E402


def main():
    ref_lib = capi.reference()
    out = {}
    for name, case in CASES.items():
        host, cam = build_case(case)
        desc = host.flatten()
        ref = api.ReferenceScene(desc)
        out[f"{name}/order"] = np.array(ref.object_order(), dtype=np.int32)
        for integ, depth, spp in case["renders"]:
            img, _, _ = ref.render(cam, case["w"], case["h"], spp, integ, depth)
            out[f"{name}/img/{capi.INTEGRATOR_NAMES[integ]}"] = img
        if case.get("primary"):
            out[f"{name}/primary"] = ref.trace_primary(cam, case["w"], case["h"], case["primary"])
        for li in range(desc.contents.n_area_lights):
            out[f"{name}/light{li}"] = np.stack([ref.kat_light_sample(li, (100.0, 200.0, 300.0), s) for s in KAT_SEEDS])
    f3 = lambda v: (C.c_float * 3)(*v)
    out["kat/sampler"] = api.kat(ref_lib, "xrtref_", "sampler", 1234, 2000, n_out=2000)
    out["kat/sampler_seed0"] = api.kat(ref_lib, "xrtref_", "sampler", 0, 700, n_out=700)
    cam = scenes.make_camera(1920, 1080)
    out["kat/camera"] = np.stack([api.kat(ref_lib, "xrtref_", "camera", C.byref(cam), C.c_float(u), C.c_float(v), n_out=6)
                                  for u, v in [(0.0, 0.0), (0.5, 0.5), (0.999, 0.001), (0.25, 0.75)]])
    normals = [(0, 0, 1), (0, 0, -1), (0.6, 0.0, 0.8), (0.3, -0.9, -0.31622776), (1, 0, 0), (0, 1, 0), (0.1, 0.2, 0.3)]
    out["kat/onb"] = np.stack([api.kat(ref_lib, "xrtref_", "onb", f3(n), n_out=6) for n in normals])
    out["kat/lambert"] = np.stack([api.kat(ref_lib, "xrtref_", "lambert_sample", f3((0, 1, 0)), f3((0.1, 0.9, 0.2)), s, n_out=4)
                                   for s in KAT_SEEDS])
    out["kat/hg"] = np.stack([api.kat(ref_lib, "xrtref_", "hg_sample", C.c_float(g), f3((0.3, 0.5, 0.81)), s, n_out=4)
                              for g in (0.0, 0.5, -0.7) for s in KAT_SEEDS])
    path = Path(__file__).with_name("reference_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, path.stat().st_size, "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
