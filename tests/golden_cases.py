"""Scene/camera cases shared by tests/golden/make_golden.py (which runs the compiled reference) and the tests that
compare the oracle port and the GPU against those vectors."""
import numpy as np

from xraytracer_b200 import capi, scenes

KAT_SEEDS = [0, 1, 2, 3, 7, 12345]

CASES = {
    # BASELINE configs 1-3 at fixture size: every surface integrator on the Cornell box
    "cornell_quad": dict(scene="cornell", light="quad", w=48, h=36, primary=2,
                         renders=[(capi.INT_NORMAL, 1, 4), (capi.INT_FURNACE, 1, 4), (capi.INT_DIRECT, 1, 4),
                                  (capi.INT_INDIRECT, 3, 4), (capi.INT_GI, 3, 4), (capi.INT_WHITTED, 3, 2)]),
    "cornell_triangle": dict(scene="cornell", light="triangle", w=32, h=24, primary=0,
                             renders=[(capi.INT_DIRECT, 1, 4), (capi.INT_GI, 3, 4)]),
    "cornell_sphere": dict(scene="cornell", light="sphere", w=32, h=24, primary=1,
                           renders=[(capi.INT_DIRECT, 1, 4), (capi.INT_GI, 3, 4)]),
    "cornell_two_lights": dict(scene="cornell", light="quad+sphere", w=32, h=24, primary=0,
                               renders=[(capi.INT_DIRECT, 1, 2), (capi.INT_GI, 3, 2)]),
    "cornell_delta": dict(scene="cornell_delta", w=32, h=24, primary=0, renders=[(capi.INT_WHITTED, 3, 2)]),
    # examples/vpt.cpp: homogeneous media
    "vpt_mis": dict(scene="vpt", kind=capi.MEDIUM_HOMOGENEOUS_MIS, w=32, h=32, primary=1,
                    renders=[(capi.INT_VOLUME, 10, 4), (capi.INT_VOLUME_NEE, 10, 4)]),
    "vpt_achromatic": dict(scene="vpt", kind=capi.MEDIUM_HOMOGENEOUS_ACHROMATIC, w=32, h=32, primary=0,
                           renders=[(capi.INT_VOLUME, 10, 4)]),
    "vpt_nomis": dict(scene="vpt", kind=capi.MEDIUM_HOMOGENEOUS_NOMIS, w=32, h=32, primary=0, renders=[(capi.INT_VOLUME, 10, 4)]),
    # BASELINE config 5 at fixture size: heterogeneous medium + delta tracking
    "hetero": dict(scene="hetero", w=32, h=32, primary=1, renders=[(capi.INT_VOLUME, 16, 4), (capi.INT_VOLUME_NEE, 16, 4)]),
}


def small_grid():
    """16^3 chromatic test grid (deterministic)."""
    return scenes.procedural_density(16, seed=7, blobs=4)


def build_case(case):
    kind = case["scene"]
    if kind == "cornell":
        return scenes.cornell_box(case["light"]), scenes.make_camera(case["w"], case["h"])
    if kind == "cornell_delta":
        def extra(s):
            s.add_point_light("PointLight", [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 278.0, 400.0, 279.5, 1], (0.63, 0.33, 0.03), 50000.0)
            s.add_distant_light("DistantLight", [0.95292, 0.289503, 0.0901785, 0, -0.0960954, 0.5704, -0.815727, 0,
                                                 -0.287593, 0.768656, 0.571365, 0, 0, 0, 0, 1], (1.0, 1.0, 1.0), 1.0)
        return scenes.cornell_box("quad", extra=extra), scenes.make_camera(case["w"], case["h"])
    if kind == "vpt":
        return scenes.vpt_scene(case["kind"]), scenes.make_camera(case["w"], case["h"], scenes.VPT_C2W, scenes.VPT_FOV)
    if kind == "hetero":
        s = scenes.volume_scene(abs_color=(0.01, 0.02, 0.03), scat_color=(0.05, 0.04, 0.03), mul=1.0, g=0.3, light="sphere",
                                voxels=small_grid())
        return s, scenes.make_camera(case["w"], case["h"])
    raise ValueError(kind)
