"""Pins the oracle port (oracle/port/xrt_oracle.cpp) against golden vectors that were produced by the reference
itself (tests/golden/make_golden.py ran the compiled reference, oracle/_ref). Bit-exact: the port restates the same
fp32 arithmetic in the same order with the same mt19937 stream. Runs anywhere (no /root/reference, no GPU)."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

from golden_cases import CASES, KAT_SEEDS, build_case
from xraytracer_b200 import api, capi, scenes

GOLD = np.load(Path(__file__).parent / "golden" / "reference_vectors.npz")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("name", sorted(CASES))
def test_port_images_match_reference_bit_for_bit(name):
    case = CASES[name]
    host, cam = build_case(case)
    desc = host.flatten()
    # the host Scene walks objects in the reference's unordered_map order
    assert [desc.contents.objects[i].insert_seq for i in range(desc.contents.n_objects)] == GOLD[f"{name}/order"].tolist()
    orc = api.OracleScene(desc)
    for integ, depth, spp in case["renders"]:
        img, _, _ = orc.render(cam, case["w"], case["h"], spp, integ, depth)
        gold = GOLD[f"{name}/img/{capi.INTEGRATOR_NAMES[integ]}"]
        assert np.array_equal(bits(img), bits(gold)), f"{name} {capi.INTEGRATOR_NAMES[integ]}: max abs {np.abs(img - gold).max()}"
    if case.get("primary"):
        hits = orc.trace_primary(cam, case["w"], case["h"], case["primary"])
        assert np.array_equal(hits, GOLD[f"{name}/primary"])
    for li in range(desc.contents.n_area_lights):
        got = np.stack([orc.kat_light_sample(li, (100.0, 200.0, 300.0), s) for s in KAT_SEEDS])
        assert np.array_equal(bits(got), bits(GOLD[f"{name}/light{li}"]))


def test_sampler_stream_is_libstdcxx_mt19937():
    lib = capi.oracle()
    assert np.array_equal(bits(api.kat(lib, "xrto_", "sampler", 1234, 2000, n_out=2000)), bits(GOLD["kat/sampler"]))
    assert np.array_equal(bits(api.kat(lib, "xrto_", "sampler", 0, 700, n_out=700)), bits(GOLD["kat/sampler_seed0"]))
    # first raw outputs of std::mt19937(5489) are 3499211612, 581869302 (the C++ standard's known answer is the
    # 10000th = 4123659995); seed 1234 first float must be float(raw)/2^32
    g = GOLD["kat/sampler"]
    assert (g >= 0).all() and (g < 1).all()


def test_function_kats():
    lib = capi.oracle()
    f3 = lambda v: (C.c_float * 3)(*v)
    cam = scenes.make_camera(1920, 1080)
    got = np.stack([api.kat(lib, "xrto_", "camera", C.byref(cam), C.c_float(u), C.c_float(v), n_out=6)
                    for u, v in [(0.0, 0.0), (0.5, 0.5), (0.999, 0.001), (0.25, 0.75)]])
    assert np.array_equal(bits(got), bits(GOLD["kat/camera"]))
    normals = [(0, 0, 1), (0, 0, -1), (0.6, 0.0, 0.8), (0.3, -0.9, -0.31622776), (1, 0, 0), (0, 1, 0), (0.1, 0.2, 0.3)]
    got = np.stack([api.kat(lib, "xrto_", "onb", f3(n), n_out=6) for n in normals])
    assert np.array_equal(bits(got), bits(GOLD["kat/onb"]))
    got = np.stack([api.kat(lib, "xrto_", "lambert_sample", f3((0, 1, 0)), f3((0.1, 0.9, 0.2)), s, n_out=4) for s in KAT_SEEDS])
    assert np.array_equal(bits(got), bits(GOLD["kat/lambert"]))
    got = np.stack([api.kat(lib, "xrto_", "hg_sample", C.c_float(g), f3((0.3, 0.5, 0.81)), s, n_out=4)
                    for g in (0.0, 0.5, -0.7) for s in KAT_SEEDS])
    assert np.array_equal(bits(got), bits(GOLD["kat/hg"]))


def test_normal_integrator_drops_negative_samples():
    """0.5*(ns+1) is slightly negative where a normal component is -1.0000001; those samples are dropped
    but still counted in the divisor. The golden image contains such pixels and the port reports them."""
    host, cam = build_case(CASES["cornell_quad"])
    orc = api.OracleScene(host.flatten())
    _, _, st = orc.render(cam, 48, 36, 4, capi.INT_NORMAL, 1)
    assert st["dropped_samples"] > 0 and st["closest_rays"] == 48 * 36 * 4


def test_furnace_converges_to_albedo():
    """Furnace estimator (integrator.h:59-66): E[f*cos/pdf] = albedo wherever a Lambert surface is seen and 0 on
    emitter proxies / misses, so the image mean must equal the mean albedo of the primary hits."""
    host, cam = build_case(CASES["cornell_quad"])
    desc = host.flatten()
    orc = api.OracleScene(desc)
    W, H = 32, 24
    img, _, _ = orc.render(cam, W, H, 256, capi.INT_FURNACE, 1)
    d = desc.contents
    albedo = []
    for i in range(d.n_objects):
        o = d.objects[i]
        a = list(d.materials[o.material].albedo) if o.material >= 0 else [0.0, 0.0, 0.0]
        albedo += [a] * (o.count if o.kind == capi.OBJ_MESH else 1)
    albedo = np.array(albedo + [[0.0, 0.0, 0.0]], dtype=np.float32)  # last row = miss (prim -1)
    prim = orc.trace_primary(cam, W, H, 64)["prim"]
    expect = albedo[prim].mean(axis=(0, 1, 2))
    assert np.abs(img.mean(axis=(0, 1)) - expect).max() < 0.01
