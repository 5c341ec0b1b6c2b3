"""Drop-in proof with the REAL reference classes (pytest -m gpu): one reference `Scene` object (built by the reference's own
addObj / addAreaLight) is rendered (a) by the reference's CPU renderer and (b) by RefGpuRenderer — a `Renderer` subclass
compiled against the reference's OWN headers that flattens that Scene and calls libxrtgpu.so through the C ABI, invoked
polymorphically through `Renderer*` as examples/cornellbox.cpp:61-63 does. Needs oracle/_ref (travels prebuilt)."""
import numpy as np
import pytest

from conftest import require_gpu
from golden_cases import CASES, build_case
from xraytracer_b200 import api, capi, scenes

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not capi.REF_GPU_LIB.exists(), reason="oracle/_ref/libxrtrefgpu.so not built")]


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_gpu_renderer_beside_the_reference_cpu_renderer_on_one_reference_scene():
    require_gpu()
    host = scenes.cornell_box("quad")
    ref = api.ReferenceGpuScene(host.flatten())          # a reference Scene; from here on only reference objects are used
    W, H = 160, 120
    cam = scenes.make_camera(W, H)
    for integ, depth, exact_bits in ((capi.INT_NORMAL, 1, True), (capi.INT_DIRECT, 1, True), (capi.INT_GI, 3, False),
                                     (capi.INT_INDIRECT, 3, False), (capi.INT_WHITTED, 3, True), (capi.INT_FURNACE, 1, False)):
        cpu, _, _ = ref.render(cam, W, H, 4, integ, depth)                       # NormalRenderer::doRender on the host cores
        gpu = ref.render_gpu(cam, W, H, 4, integ, depth, flags=capi.FLAG_EXACT)   # RefGpuRenderer -> libxrtgpu.so
        if exact_bits:
            assert np.array_equal(bits(cpu), bits(gpu)), capi.INTEGRATOR_NAMES[integ]
        else:
            assert np.abs(cpu - gpu).max() < 2e-5, capi.INTEGRATOR_NAMES[integ]
    fast = ref.render_gpu(cam, W, H, 8192, capi.INT_GI, 3, seed=3)               # throughput path, same Scene object
    conv, _, _ = ref.render(cam, W, H, 1024, capi.INT_GI, 3)
    assert abs(float(fast.mean()) - float(conv.mean())) < 0.005 * float(conv.mean())


def test_gpu_renderer_on_reference_volume_scenes():
    require_gpu()
    for name in ("vpt_mis", "hetero"):
        case = CASES[name]
        host, cam = build_case(case)
        ref = api.ReferenceGpuScene(host.flatten())
        for integ in (capi.INT_VOLUME, capi.INT_VOLUME_NEE):
            cpu, _, _ = ref.render(cam, case["w"], case["h"], 4, integ, 12)
            gpu = ref.render_gpu(cam, case["w"], case["h"], 4, integ, 12, flags=capi.FLAG_EXACT)
            assert np.abs(cpu - gpu).max() < 5e-5, (name, integ)
