"""The drop-in boundary: libxrtgpu.so loads, exports every symbol include/xrtgpu.h declares, and fails loudly —
never falls back — when there is no CUDA device."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from xraytracer_b200 import api, capi, scenes

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "xrtgpu.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(xrtg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported():
    lib = capi.gpu()
    syms = declared_symbols()
    assert len(syms) >= 11
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in xrtgpu.h but not exported"
    assert sorted(capi.GPU_SYMBOLS) == syms


def test_abi_version_and_struct_sizes():
    lib = capi.gpu()
    assert lib.xrtg_abi_version() == capi.ABI_VERSION
    assert C.sizeof(capi.Triangle) == 72 and C.sizeof(capi.Hit) == 16 and C.sizeof(capi.Sphere) == 16
    assert C.sizeof(capi.Camera) == 72 and C.sizeof(capi.RenderParams) == 40


def test_only_c_abi_is_exported():
    """hidden visibility: nothing but xrtg_* leaves the CUDA library, xrth_* the host library."""
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", str(capi.GPU_LIB)], capture_output=True, text=True).stdout
    names = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert names and all(n.startswith("xrtg_") for n in names), names
    # the host library exports the scripting surface plus the C++ drop-in API (Scene, GpuRenderer) and nothing else
    out = subprocess.run(["nm", "-D", "-C", "--defined-only", str(capi.HOST_LIB)], capture_output=True, text=True).stdout
    names = [l.split(" T ", 1)[1] for l in out.splitlines() if " T " in l]
    assert names and all(n.startswith(("xrth_", "Scene::", "GpuRenderer::")) for n in names), names


def test_invalid_arguments_are_rejected(cornell):
    lib = capi.gpu()
    h = C.c_void_p()
    assert lib.xrtg_scene_create(None, 0, C.byref(h)) == -1
    assert b"NULL" in lib.xrtg_last_error()
    bad = capi.SceneDesc()
    bad.abi_version = 99
    assert lib.xrtg_scene_create(C.byref(bad), 0, C.byref(h)) == -1
    assert b"abi_version" in lib.xrtg_last_error()


def test_no_device_fails_loudly_without_fallback(cornell):
    """On a machine without a GPU scene creation must return XRTG_ERR_NO_DEVICE — there is no CPU render path."""
    lib = capi.gpu()
    if lib.xrtg_device_count() > 0:
        pytest.skip("a CUDA device is present")
    _, desc = cornell
    h = C.c_void_p()
    rc = lib.xrtg_scene_create(desc, 0, C.byref(h))
    assert rc == -2 and not h.value
    assert b"no CUDA device" in lib.xrtg_last_error()
    with pytest.raises(RuntimeError, match="no CUDA device"):
        api.GpuScene(desc, 0)
    # the C++ GpuRenderer (through the scripting surface) must throw too, not render on the CPU
    host, _ = cornell
    rgb = np.zeros((8, 8, 3), np.float32)
    rc = capi.host().xrth_render(host.h, 1.0, (C.c_float * 16)(*scenes.CORNELL_C2W), 60.0, capi.INT_NORMAL, 1, 1, 8, 8, 0, 0,
                                 rgb.ctypes.data, None)
    assert rc != 0 and b"no CUDA device" in capi.host().xrth_last_error()
    assert not rgb.any()
