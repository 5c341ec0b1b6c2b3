"""Python-side handles with one interface over the three implementations that consume an xrtg_scene_desc:

  GpuScene        libxrtgpu.so       the product: sm_100a wavefront kernels behind the C ABI (include/xrtgpu.h)
  OracleScene     libxrtoracle.so    CPU restatement (oracle/port)            — tests / cpu_baseline only
  ReferenceScene  libxrtref.so       the compiled reference (oracle/_ref)     — tests / cpu_baseline only

Only GpuScene is product code; it raises when libxrtgpu.so is missing or no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi

HIT_DTYPE = np.dtype([("t", np.float32), ("u", np.float32), ("v", np.float32), ("prim", np.int32)])


def _params(width, height, spp, integrator, max_depth, flags=0, seed=0, sample_offset=0, spp_total=0, samples_per_wave=0):
    p = capi.RenderParams()
    p.width, p.height, p.spp, p.sample_offset, p.spp_total = width, height, spp, sample_offset, spp_total
    p.integrator, p.max_depth, p.seed, p.flags, p.samples_per_wave = integrator, max_depth, seed, flags, samples_per_wave
    return p


def _rays(org, dir, tmax):
    org = np.ascontiguousarray(org, dtype=np.float32).reshape(-1, 3)
    dir = np.ascontiguousarray(dir, dtype=np.float32).reshape(-1, 3)
    assert org.shape == dir.shape
    if tmax is not None:
        tmax = np.ascontiguousarray(tmax, dtype=np.float32).reshape(-1)
        assert len(tmax) == len(org)
    return org, dir, tmax


class GpuScene:
    """Device scene (SAH BVH + flattened arrays resident in HBM) created through xrtg_scene_create."""

    def __init__(self, desc, device: int = 0, build_flags: int = 0, devices=None):
        """devices: None = one device (`device`); an int N = devices 0..N-1 behind ONE handle (xrtg_scene_create_multi);
        a list = exactly those devices."""
        self.lib = capi.gpu()
        h = C.c_void_p()
        if devices is None:
            rc = self.lib.xrtg_scene_create2(desc, device, build_flags, C.byref(h))
        else:
            ids = list(range(devices)) if isinstance(devices, int) else [int(d) for d in devices]
            arr = (C.c_int * len(ids))(*ids)
            rc = self.lib.xrtg_scene_create_multi(desc, len(ids), arr, build_flags, C.byref(h))
            device = ids[0] if ids else 0
        if rc != 0:
            raise RuntimeError(f"xrtg_scene_create failed ({rc}): {self.lib.xrtg_last_error().decode()}")
        self.h = h
        self.device = device

    def set_tuning(self, **kw):
        """Development / test switches of the pipeline selection (xrtg_tuning); omitted keys return to their defaults."""
        t = capi.Tuning(**kw)
        self._chk(self.lib.xrtg_scene_set_tuning(self.h, C.byref(t)), "xrtg_scene_set_tuning")

    def check_guards(self) -> int:
        """Guard bytes around the workspace buffers that a kernel overwrote (0 = no out-of-bounds write)."""
        n = C.c_int()
        self._chk(self.lib.xrtg_scene_check_guards(self.h, C.byref(n)), "xrtg_scene_check_guards")
        return n.value

    def selfcheck(self) -> int:
        """Structural check of the device-resident trees (xrtg_scene_selfcheck); returns the number of violations."""
        n = C.c_int()
        self._chk(self.lib.xrtg_scene_selfcheck(self.h, C.byref(n)), "xrtg_scene_selfcheck")
        if n.value:
            self.last_selfcheck_error = self.lib.xrtg_last_error().decode()
        return n.value

    def device_count(self) -> int:
        return self.lib.xrtg_scene_device_count(self.h)

    def exchange_buffer(self, slot: int, nbytes: int) -> int:
        """Device pointer of an exportable scene-owned buffer (one process per GPU + CUDA IPC): slot 0 = per-pixel SUM,
        slot 1 = the final image on the root rank."""
        p = C.c_void_p()
        self._chk(self.lib.xrtg_exchange_buffer(self.h, slot, nbytes, C.byref(p)), "xrtg_exchange_buffer")
        return p.value

    def ipc_export(self, device_ptr: int) -> bytes:
        buf = (C.c_ubyte * 64)()
        self._chk(self.lib.xrtg_ipc_export(C.c_void_p(device_ptr), buf), "xrtg_ipc_export")
        return bytes(buf)

    def ipc_open(self, handle: bytes) -> int:
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        p = C.c_void_p()
        self._chk(self.lib.xrtg_ipc_open(self.h, buf, C.byref(p)), "xrtg_ipc_open")
        return p.value

    def ipc_close(self, device_ptr: int):
        self._chk(self.lib.xrtg_ipc_close(self.h, C.c_void_p(device_ptr)), "xrtg_ipc_close")

    def reduce_finalize(self, parts, out_ptr: int, first: int, count: int, divisor: float, stream: int = 0):
        """out[i] = sum_k parts[k][i] / divisor over [first, first+count), asynchronously on `stream` (fused reduce + finalize;
        parts / out may be peer-device or IPC-mapped pointers)."""
        arr = (C.c_void_p * len(parts))(*[C.c_void_p(p) for p in parts])
        self._chk(self.lib.xrtg_reduce_finalize(self.h, arr, len(parts), C.c_void_p(out_ptr), first, count, divisor, C.c_void_p(stream)),
                  "xrtg_reduce_finalize")

    def render_u8(self, cam, width, height, spp, integrator, max_depth=1, gamma=0.0, bgr=False, flags=0, seed=0):
        """Render + `image /= spp` + gammaCorrection + 8-bit quantisation on the device; only the bytes come back."""
        p = _params(width, height, spp, integrator, max_depth, flags, seed)
        out = np.empty((height, width, 3), dtype=np.uint8)
        st = capi.Stats()
        self._chk(self.lib.xrtg_render_u8(self.h, C.byref(cam), C.byref(p), gamma, int(bgr), out.ctypes.data, C.byref(st)), "xrtg_render_u8")
        return out, st.as_dict()

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.xrtg_scene_destroy(self.h)
            self.h = None

    def _chk(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): {self.lib.xrtg_last_error().decode()}")

    def info(self) -> dict:
        i = capi.SceneInfo()
        self._chk(self.lib.xrtg_scene_get_info(self.h, C.byref(i)), "xrtg_scene_get_info")
        return i.as_dict()

    def upload(self):
        self._chk(self.lib.xrtg_scene_upload(self.h), "xrtg_scene_upload")

    def render(self, cam, width, height, spp, integrator, max_depth=1, flags=0, seed=0, sample_offset=0, spp_total=0,
               samples_per_wave=0, out=None):
        """Host-buffer render through xrtg_render. Returns (rgb[H,W,3] float32, stats dict)."""
        p = _params(width, height, spp, integrator, max_depth, flags, seed, sample_offset, spp_total, samples_per_wave)
        rgb = out if out is not None else np.empty((height, width, 3), dtype=np.float32)
        st = capi.Stats()
        self._chk(self.lib.xrtg_render(self.h, C.byref(cam), C.byref(p), rgb.ctypes.data, C.byref(st)), "xrtg_render")
        return rgb, st.as_dict()

    def render_device(self, cam, width, height, spp, integrator, max_depth, device_ptr: int, stream: int = 0, flags=0, seed=0,
                      sample_offset=0, spp_total=0, samples_per_wave=0, want_stats=True):
        """Asynchronous render into a device buffer (e.g. a torch tensor's data_ptr()) on `stream`."""
        p = _params(width, height, spp, integrator, max_depth, flags, seed, sample_offset, spp_total, samples_per_wave)
        st = capi.Stats()
        self._chk(self.lib.xrtg_render_device(self.h, C.byref(cam), C.byref(p), C.c_void_p(device_ptr), C.c_void_p(stream),
                                               C.byref(st) if want_stats else None), "xrtg_render_device")
        return st.as_dict() if want_stats else None

    def trace_primary(self, cam, width, height, spp, jitter=None, flags=0):
        hits = np.empty(width * height * spp, dtype=HIT_DTYPE)
        jp = None
        if jitter is not None:
            jitter = np.ascontiguousarray(jitter, dtype=np.float32)
            assert jitter.size == width * height * spp * 2
            jp = jitter.ctypes.data
        self._chk(self.lib.xrtg_trace_primary(self.h, C.byref(cam), width, height, spp, jp, flags, hits.ctypes.data),
                  "xrtg_trace_primary")
        return hits.reshape(height, width, spp)

    def trace_rays(self, org, dir, tmax=None, any_hit=False, flags=0, src_prim=None):
        """src_prim (any_hit with FLAG_FAST_HOOK): per ray, the primitive the shadow ray starts on (see XRTG_FLAG_HOOK_SRC_PRIM)."""
        org, dir, tmax = _rays(org, dir, tmax)
        hits = np.empty(len(org), dtype=HIT_DTYPE)
        if src_prim is not None:
            hits["prim"] = np.asarray(src_prim, dtype=np.int32)
            flags |= capi.FLAG_HOOK_SRC_PRIM
        self._chk(self.lib.xrtg_trace_rays(self.h, len(org), org.ctypes.data, dir.ctypes.data,
                                            None if tmax is None else tmax.ctypes.data, int(any_hit), flags, hits.ctypes.data),
                  "xrtg_trace_rays")
        return hits


def image_to_u8(rgb: np.ndarray, gamma: float = 0.0, bgr: bool = False, device: int = 0) -> np.ndarray:
    """Image::gammaCorrection + the 8-bit quantisation of writePPM / writeMat (image.h:80-136) on the device."""
    rgb = np.ascontiguousarray(rgb, dtype=np.float32)
    h, w, _ = rgb.shape
    out = np.empty((h, w, 3), dtype=np.uint8)
    lib = capi.gpu()
    rc = lib.xrtg_image_to_u8(device, rgb.ctypes.data, w, h, gamma, int(bgr), out.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"xrtg_image_to_u8 failed ({rc}): {lib.xrtg_last_error().decode()}")
    return out


class _CpuScene:
    """Shared wrapper of the two CPU checkers (same entry-point shapes, different prefix)."""
    pfx = ""
    has_stats = False

    def __init__(self, desc):
        self.lib = self._lib()
        h = C.c_void_p()
        rc = self.fn("scene_create")(desc, C.byref(h))
        if rc != 0:
            raise RuntimeError(f"{self.pfx}scene_create failed: {self.fn('last_error')().decode()}")
        self.h = h

    def fn(self, name):
        return getattr(self.lib, self.pfx + name)

    def __del__(self):
        if getattr(self, "h", None):
            self.fn("scene_destroy")(self.h)
            self.h = None

    def max_threads(self):
        return self.fn("max_threads")()

    def render(self, cam, width, height, spp, integrator, max_depth=1, flags=0, nthreads=0, pixel_stride=1, spp_total=0):
        """Returns (rgb[H,W,3], seconds of the pixel loop, stats dict or None)."""
        p = _params(width, height, spp, integrator, max_depth, flags, 0, 0, spp_total)
        rgb = np.zeros((height, width, 3), dtype=np.float32)
        sec = C.c_double()
        if self.has_stats:
            st = capi.Stats()
            rc = self.fn("render")(self.h, C.byref(cam), C.byref(p), nthreads, pixel_stride, rgb.ctypes.data, C.byref(sec), C.byref(st))
            stats = st.as_dict()
        else:
            rc = self.fn("render")(self.h, C.byref(cam), C.byref(p), nthreads, pixel_stride, rgb.ctypes.data, C.byref(sec))
            stats = None
        if rc != 0:
            raise RuntimeError(f"{self.pfx}render failed: {self.fn('last_error')().decode()}")
        return rgb, sec.value, stats

    def trace_primary(self, cam, width, height, spp, jitter=None):
        hits = np.empty(width * height * spp, dtype=HIT_DTYPE)
        jp = None
        if jitter is not None:
            jitter = np.ascontiguousarray(jitter, dtype=np.float32)
            jp = jitter.ctypes.data
        self.fn("trace_primary")(self.h, C.byref(cam), width, height, spp, jp, hits.ctypes.data)
        return hits.reshape(height, width, spp)

    def trace_rays(self, org, dir, tmax=None, any_hit=False):
        org, dir, tmax = _rays(org, dir, tmax)
        hits = np.empty(len(org), dtype=HIT_DTYPE)
        self.fn("trace_rays")(self.h, len(org), org.ctypes.data, dir.ctypes.data, None if tmax is None else tmax.ctypes.data,
                              int(any_hit), hits.ctypes.data)
        return hits

    # known-answer hooks
    def kat_light_sample(self, li, pos, seed):
        out = (C.c_float * 8)()
        self.fn("kat_light_sample")(self.h, li, (C.c_float * 3)(*pos), seed, out)
        return np.array(out[:], dtype=np.float32)


class OracleScene(_CpuScene):
    pfx = "xrto_"
    has_stats = True

    @staticmethod
    def _lib():
        return capi.oracle()


class ReferenceScene(_CpuScene):
    pfx = "xrtref_"
    has_stats = False

    @staticmethod
    def _lib():
        return capi.reference()

    def render_gpu(self, cam, width, height, spp, integrator, max_depth=1, flags=0, seed=0):
        """The SAME reference Scene object rendered by RefGpuRenderer (oracle/ref_harness.cpp): a `Renderer` subclass built
        against the reference's own headers that calls libxrtgpu.so — the binding INTEGRATION.md describes."""
        p = _params(width, height, spp, integrator, max_depth, flags, seed)
        rgb = np.zeros((height, width, 3), dtype=np.float32)
        rc = self.lib.xrtref_render_gpu(self.h, C.byref(cam), C.byref(p), rgb.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"xrtref_render_gpu failed: {self.lib.xrtref_last_error().decode()}")
        return rgb

    def object_order(self):
        buf = (C.c_int32 * 4096)()
        n = self.lib.xrtref_object_order(self.h, buf, 4096)
        return list(buf[:n])


class ReferenceGpuScene(ReferenceScene):
    """The same through libxrtrefgpu.so, which adds RefGpuRenderer (and with it a link to libxrtgpu.so): drop-in test only."""

    @staticmethod
    def _lib():
        return capi.reference(with_gpu=True)


def kat(lib, pfx, name, *args, n_out):
    out = (C.c_float * n_out)()
    getattr(lib, pfx + "kat_" + name)(*args, out)
    return np.array(out[:], dtype=np.float32)
