"""Multi-GPU behind the C ABI (pytest -m gpu): xrtg_scene_create_multi + xrtg_render split the samples of ONE call across the
devices of the handle and finish with the fused peer-memory reduce + `image /= n_samples` kernel (csrc/multi.cu; the role of
ParallelRenderer::render, renderer.cpp:83-99, and of renderer.cpp:98). On a one-GPU box the same code path runs with one
device listed twice (two replicas, two host threads, two streams); with >= 2 GPUs it runs over NVLink peer memory."""
import ctypes as C

import numpy as np
import pytest

from conftest import require_gpu
from xraytracer_b200 import api, capi, scenes

pytestmark = pytest.mark.gpu


def device_sets():
    n = capi.gpu().xrtg_device_count()
    sets = [[0, 0], [0, 0, 0]]
    if n >= 2:
        sets += [list(range(min(n, k))) for k in (2, 4, 8) if k <= n or k == 2]
    return sets


def test_multi_gpu_render_equals_single_device_render(cornell):
    require_gpu()
    host, desc = cornell
    W, H, spp = 160, 90, 24
    cam = scenes.make_camera(W, H)
    single = api.GpuScene(desc, 0)
    ref, rst = single.render(cam, W, H, spp, capi.INT_GI, 3, seed=5)
    for devs in device_sets():
        multi = api.GpuScene(desc, devices=devs)
        assert multi.device_count() == len(devs) and multi.info()["n_devices"] == len(devs)
        img, st = multi.render(cam, W, H, spp, capi.INT_GI, 3, seed=5)
        # same sample set (counter RNG keyed by sample index); only the fp32 association of the per-pixel sum differs
        assert np.allclose(img, ref, rtol=3e-6, atol=1e-6), devs
        assert st["n_devices"] == len(devs) and st["samples"] == W * H * spp
        assert (st["closest_rays"], st["shadow_rays"], st["dropped_samples"]) == (rst["closest_rays"], rst["shadow_rays"], rst["dropped_samples"])
        # SUM_ONLY through the multi handle = the per-pixel sum
        s, _ = multi.render(cam, W, H, spp, capi.INT_GI, 3, seed=5, flags=capi.FLAG_SUM_ONLY)
        assert np.allclose(s / spp, ref, rtol=3e-6, atol=1e-6)
        # a sample range of a larger render (resumable accumulation) splits again inside the handle
        part, _ = multi.render(cam, W, H, 8, capi.INT_GI, 3, seed=5, sample_offset=8, spp_total=spp, flags=capi.FLAG_SUM_ONLY)
        want, _ = single.render(cam, W, H, 8, capi.INT_GI, 3, seed=5, sample_offset=8, spp_total=spp, flags=capi.FLAG_SUM_ONLY)
        assert np.allclose(part, want, rtol=3e-6, atol=1e-6)
        # exact mode replays one mt19937 stream per pixel: not split, rendered on device 0, still bit-exact
        e, _ = multi.render(cam, W, H, 2, capi.INT_NORMAL, 1, flags=capi.FLAG_EXACT)
        e1, _ = single.render(cam, W, H, 2, capi.INT_NORMAL, 1, flags=capi.FLAG_EXACT)
        assert np.array_equal(e.view(np.uint32), e1.view(np.uint32))
        multi.upload()   # re-upload of every replica leaves the result unchanged
        again, _ = multi.render(cam, W, H, spp, capi.INT_GI, 3, seed=5)
        assert np.array_equal(again.view(np.uint32), img.view(np.uint32))


def test_multi_gpu_volume_and_deep_scenes():
    require_gpu()
    devs = device_sets()[-1]
    W, H = 128, 72
    cam = scenes.make_camera(W, H)
    vol = scenes.volume_scene(n=24)           # grid descriptors carry per-device voxel pointers
    d = vol.flatten()
    a, sa = api.GpuScene(d, 0).render(cam, W, H, 16, capi.INT_VOLUME, 8, seed=2)
    mv = api.GpuScene(d, devices=devs)
    b, sb = mv.render(cam, W, H, 16, capi.INT_VOLUME, 8, seed=2)
    assert np.allclose(a, b, rtol=3e-6, atol=1e-6) and sa["tracking_steps"] == sb["tracking_steps"] > 0
    # re-upload: host -> device 0 once, then the broadcast tree of peer copies (voxels, their 3-D texture copies, every array)
    mv.upload()
    b2, sb2 = mv.render(cam, W, H, 16, capi.INT_VOLUME, 8, seed=2)
    assert np.array_equal(b2, b) and sb2["tracking_steps"] == sb["tracking_steps"]
    extra = lambda h: h.add_mesh("tess", scenes.displaced_sphere_tris((278, 200, 280), 150, 40, 40), (0.75, 0.75, 0.75))
    deep = scenes.cornell_box("quad", extra=extra)
    d = deep.flatten()
    a, sa = api.GpuScene(d, 0).render(cam, W, H, 8, capi.INT_GI, 3, seed=2)
    md = api.GpuScene(d, devices=devs)
    b, sb = md.render(cam, W, H, 8, capi.INT_GI, 3, seed=2)
    assert np.allclose(a, b, rtol=3e-5, atol=1e-5) and sa["closest_rays"] == sb["closest_rays"]
    md.upload()
    b2, _ = md.render(cam, W, H, 8, capi.INT_GI, 3, seed=2)
    assert np.array_equal(b2, b)


def test_reduce_finalize_kernel_and_exchange_buffers(cornell):
    """xrtg_reduce_finalize on local buffers: out[i] = (a[i] + b[i] + c[i]) / divisor over arbitrary (unaligned) slices."""
    require_gpu()
    import torch
    host, desc = cornell
    g = api.GpuScene(desc, 0)
    n = 1000 * 3 + 1
    rng = np.random.RandomState(0)
    parts = [rng.uniform(0, 10, n).astype(np.float32) for _ in range(3)]
    dev = [torch.from_numpy(p).cuda() for p in parts]
    out = torch.full((n,), -1.0, dtype=torch.float32, device="cuda")
    for first, count, div in ((0, n, 7.0), (5, 1001, 0.0), (1, 2, 3.0), (n - 3, 3, 2.0), (8, 0, 1.0)):
        out.fill_(-1.0)
        g.reduce_finalize([t.data_ptr() for t in dev], out.data_ptr(), first, count, div, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        want = np.full(n, -1.0, np.float32)
        acc = (parts[0] + parts[1]) + parts[2]
        want[first:first + count] = (acc / np.float32(div) if div > 0 else acc)[first:first + count]
        assert np.array_equal(out.cpu().numpy(), want), (first, count, div)
    # exchange buffers are plain, exportable device allocations
    p0 = g.exchange_buffer(0, 4096)
    p1 = g.exchange_buffer(1, 4096)
    assert p0 and p1 and p0 != p1 and g.exchange_buffer(0, 1024) == p0
    assert len(g.ipc_export(p0)) == 64
    with pytest.raises(RuntimeError):
        g.exchange_buffer(7, 16)


def test_errors_of_the_multi_gpu_handle(cornell):
    require_gpu()
    host, desc = cornell
    with pytest.raises(RuntimeError, match="out of range"):
        api.GpuScene(desc, devices=[0, 99])
    with pytest.raises(RuntimeError):
        api.GpuScene(desc, devices=[])
    m = api.GpuScene(desc, devices=[0, 0])
    import torch
    buf = torch.empty((8, 8, 3), dtype=torch.float32, device="cuda")
    with pytest.raises(RuntimeError, match="multi-GPU"):
        m.render_device(scenes.make_camera(8, 8), 8, 8, 4, capi.INT_GI, 3, buf.data_ptr())
