"""ctypes bindings of the C ABI in include/xrtgpu.h (libxrtgpu.so), of the host scripting surface
(libxrthost.so, xraytracer_b200/host/capi.cpp) and — for tests/bench baselines only — of the two
oracle libraries (oracle/libxrtoracle.so = CPU restatement, oracle/_ref/libxrtref.so = the compiled
reference). The product libraries are loaded eagerly by :func:`gpu` / :func:`host` and raise if
missing: there is no Python or CPU fallback for the render path.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
PKG = Path(__file__).resolve().parent

F3 = C.c_float * 3
I3 = C.c_int32 * 3


class Triangle(C.Structure):
    _fields_ = [("v0", F3), ("v1", F3), ("v2", F3), ("n0", F3), ("n1", F3), ("n2", F3)]


class Sphere(C.Structure):
    _fields_ = [("center", F3), ("radius", C.c_float)]


class Box(C.Structure):
    _fields_ = [("pmin", F3), ("pmax", F3)]


class Object(C.Structure):
    _fields_ = [("kind", C.c_int32), ("first", C.c_int32), ("count", C.c_int32), ("material", C.c_int32),
                ("area_light", C.c_int32), ("medium", C.c_int32), ("insert_seq", C.c_int32), ("_pad", C.c_int32),
                ("name", C.c_char_p)]


class Material(C.Structure):
    _fields_ = [("kind", C.c_int32), ("albedo", F3)]


class AreaLight(C.Structure):
    _fields_ = [("kind", C.c_int32), ("v0", F3), ("v1", F3), ("v2", F3), ("radius", C.c_float), ("Le", F3)]


class DeltaLight(C.Structure):
    _fields_ = [("kind", C.c_int32), ("pos_or_dir", F3), ("radiance", F3)]


class Medium(C.Structure):
    _fields_ = [("kind", C.c_int32), ("g", C.c_float), ("sigma_a", F3), ("sigma_s", F3), ("density_mul", C.c_float),
                ("grid", C.c_int32)]


class Grid(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("data", C.POINTER(C.c_float)),
                ("origin", F3), ("voxel_size", C.c_float), ("background", C.c_float), ("active_min", I3),
                ("active_max", I3), ("max_density", C.c_float)]


class SceneDesc(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("n_objects", C.c_int32), ("n_triangles", C.c_int32),
                ("n_spheres", C.c_int32), ("n_boxes", C.c_int32), ("n_materials", C.c_int32),
                ("n_area_lights", C.c_int32), ("n_delta_lights", C.c_int32), ("n_media", C.c_int32),
                ("n_grids", C.c_int32),
                ("objects", C.POINTER(Object)), ("triangles", C.POINTER(Triangle)), ("spheres", C.POINTER(Sphere)),
                ("boxes", C.POINTER(Box)), ("materials", C.POINTER(Material)), ("area_lights", C.POINTER(AreaLight)),
                ("delta_lights", C.POINTER(DeltaLight)), ("media", C.POINTER(Medium)), ("grids", C.POINTER(Grid))]


class Camera(C.Structure):
    _fields_ = [("c2w", C.c_float * 16), ("scale", C.c_float), ("aspect", C.c_float)]


class RenderParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("spp", C.c_int32), ("sample_offset", C.c_int32),
                ("spp_total", C.c_int32), ("integrator", C.c_int32), ("max_depth", C.c_int32), ("seed", C.c_uint32),
                ("flags", C.c_uint32), ("samples_per_wave", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("closest_rays", C.c_uint64), ("shadow_rays", C.c_uint64),
                ("dropped_samples", C.c_uint64), ("nodes_visited", C.c_uint64), ("tris_tested", C.c_uint64),
                ("nodes_visited_shadow", C.c_uint64), ("tris_tested_shadow", C.c_uint64),
                ("tracking_steps", C.c_uint64), ("kernel_launches", C.c_uint64), ("extend_launches", C.c_uint64),
                ("shade_launches", C.c_uint64), ("connect_launches", C.c_uint64), ("render_ms", C.c_float),
                ("extend_ms", C.c_float), ("connect_ms", C.c_float), ("shade_ms", C.c_float), ("other_ms", C.c_float),
                ("h2d_ms", C.c_float), ("d2h_ms", C.c_float), ("primary_hits", C.c_uint64),
                ("bounce_entries", C.c_uint64), ("bounce_launches", C.c_uint64), ("rays_traced", C.c_uint64),
                ("truncated_paths", C.c_uint64), ("reduce_ms", C.c_float), ("n_devices", C.c_int32),
                ("untraced_closest", C.c_uint64), ("untraced_shadow", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Hit(C.Structure):
    _fields_ = [("t", C.c_float), ("u", C.c_float), ("v", C.c_float), ("prim", C.c_int32)]


class SceneInfo(C.Structure):
    _fields_ = [("n_prims", C.c_int32), ("n_triangles", C.c_int32), ("n_bvh_nodes", C.c_int32),
                ("bvh_depth", C.c_int32), ("bvh_sah_cost", C.c_float), ("build_ms", C.c_float),
                ("upload_ms", C.c_float), ("device_bytes", C.c_uint64), ("upload_bytes", C.c_uint64),
                ("bvh_build_ms", C.c_float), ("bvh_builder", C.c_int32),
                ("small_records_all", C.c_int32), ("small_records_occ", C.c_int32), ("small_flagged", C.c_int32),
                ("n_wide_nodes", C.c_int32), ("wide_arity", C.c_int32), ("n_devices", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


TUNING_KEYS = ["fused_bounce", "volume_paths", "scissor", "brute_secondary", "brute_shadow", "thr_ext0", "thr_ext", "thr_con",
               "steps_per_vote", "leaf_threshold", "thr_vol", "spv_vol", "wide_bvh", "max_leaf", "workspace_mb", "stage_dump", "primary_masks",
               "gpu_build", "ploc_radius", "ploc_ct_x16", "ploc_top", "grid_texture", "overlap_connect", "ploc_weight"]


class Tuning(C.Structure):
    """xrtg_tuning: development / test switches of the pipeline selection; -1 = the measured default."""
    _fields_ = [(k, C.c_int32) for k in TUNING_KEYS]

    def __init__(self, **kw):
        super().__init__()
        for k in TUNING_KEYS:
            setattr(self, k, int(kw.pop(k, -1)))
        if kw:
            raise TypeError(f"unknown tuning keys: {sorted(kw)}")


# enums of xrtgpu.h
INT_NORMAL, INT_FURNACE, INT_DIRECT, INT_INDIRECT, INT_GI, INT_WHITTED, INT_VOLUME, INT_VOLUME_NEE = range(8)
INTEGRATOR_NAMES = ["normal", "furnace", "direct", "indirect", "gi", "whitted", "volume", "volume_nee"]
FLAG_EXACT, FLAG_COUNTERS, FLAG_BRUTE_FORCE, FLAG_SUM_ONLY, FLAG_STAGE_TIMES, FLAG_FAST_HOOK, FLAG_HOOK_SRC_PRIM = 1, 2, 4, 8, 16, 32, 64
OBJ_MESH, OBJ_SPHERE, OBJ_BOX = 0, 1, 2
LIGHT_QUAD, LIGHT_TRIANGLE, LIGHT_SPHERE = 0, 1, 2
MEDIUM_HOMOGENEOUS_MIS, MEDIUM_HOMOGENEOUS_ACHROMATIC, MEDIUM_HOMOGENEOUS_NOMIS, MEDIUM_HETEROGENEOUS = range(4)
ABI_VERSION = 2
BUILD_LBVH_GPU, BUILD_GPU, BUILD_HOST = 1, 2, 4

# every symbol include/xrtgpu.h declares (checked by tests/test_abi.py)
GPU_SYMBOLS = ["xrtg_abi_version", "xrtg_device_count", "xrtg_last_error", "xrtg_scene_create", "xrtg_scene_create2", "xrtg_scene_upload",
               "xrtg_scene_get_info", "xrtg_scene_destroy", "xrtg_render", "xrtg_render_device", "xrtg_trace_primary",
               "xrtg_trace_rays", "xrtg_image_to_u8", "xrtg_bvh_selftest", "xrtg_small_scene_selftest", "xrtg_scene_create_multi",
               "xrtg_scene_device_count", "xrtg_scene_set_tuning", "xrtg_exchange_buffer", "xrtg_ipc_export", "xrtg_ipc_open",
               "xrtg_ipc_close", "xrtg_reduce_finalize", "xrtg_render_u8", "xrtg_scene_check_guards", "xrtg_scene_selfcheck", "xrtg_top_sah_selftest"]

GPU_LIB = Path(os.environ["XRT_GPU_LIB"]) if os.environ.get("XRT_GPU_LIB") else PKG / "csrc" / "libxrtgpu.so"   # (override: A/B builds)
HOST_LIB = PKG / "host" / "libxrthost.so"
ORACLE_LIB = ROOT / "oracle" / "libxrtoracle.so"
REF_LIB = ROOT / "oracle" / "_ref" / "libxrtref.so"          # CPU checker / reference arm: no product library behind it
REF_GPU_LIB = ROOT / "oracle" / "_ref" / "libxrtrefgpu.so"   # the same + RefGpuRenderer -> libxrtgpu.so (drop-in test only)

_cache = {}
P = C.POINTER
VP = C.c_void_p


def _load(path: Path, what: str):
    key = str(path)
    if key not in _cache:
        if not path.exists():
            raise RuntimeError(f"{what} not built: {path} is missing — run `python -c 'import __graft_entry__ as g; "
                               f"g.build()'` (there is no fallback for it)")
        _cache[key] = C.CDLL(key)  # RTLD_LOCAL: host and reference libraries both define `Scene` etc.
    return _cache[key]


def gpu():
    """libxrtgpu.so with argtypes set. Raises if the CUDA library has not been built."""
    lib = _load(GPU_LIB, "CUDA library libxrtgpu.so")
    if getattr(lib, "_typed", False):
        return lib
    lib.xrtg_abi_version.restype = C.c_int
    lib.xrtg_device_count.restype = C.c_int
    lib.xrtg_last_error.restype = C.c_char_p
    lib.xrtg_scene_create.argtypes = [P(SceneDesc), C.c_int, P(VP)]
    lib.xrtg_scene_create2.argtypes = [P(SceneDesc), C.c_int, C.c_uint32, P(VP)]
    lib.xrtg_scene_upload.argtypes = [VP]
    lib.xrtg_scene_get_info.argtypes = [VP, P(SceneInfo)]
    lib.xrtg_scene_destroy.argtypes = [VP]
    lib.xrtg_scene_destroy.restype = None
    lib.xrtg_render.argtypes = [VP, P(Camera), P(RenderParams), VP, P(Stats)]
    lib.xrtg_render_device.argtypes = [VP, P(Camera), P(RenderParams), VP, VP, P(Stats)]
    lib.xrtg_trace_primary.argtypes = [VP, P(Camera), C.c_int, C.c_int, C.c_int, VP, C.c_uint32, VP]
    lib.xrtg_trace_rays.argtypes = [VP, C.c_int64, VP, VP, VP, C.c_int, C.c_uint32, VP]
    lib.xrtg_image_to_u8.argtypes = [C.c_int, VP, C.c_int, C.c_int, C.c_float, C.c_int, VP]
    lib.xrtg_bvh_selftest.argtypes = [VP, C.c_int, C.c_int, P(C.c_int), P(C.c_int), P(C.c_float)]
    lib.xrtg_small_scene_selftest.argtypes = [VP, VP, C.c_int, P(C.c_int), P(C.c_int), P(C.c_int)]
    lib.xrtg_scene_create_multi.argtypes = [P(SceneDesc), C.c_int, P(C.c_int), C.c_uint32, P(VP)]
    lib.xrtg_scene_device_count.argtypes = [VP]
    lib.xrtg_scene_set_tuning.argtypes = [VP, P(Tuning)]
    lib.xrtg_scene_check_guards.argtypes = [VP, P(C.c_int)]
    lib.xrtg_scene_selfcheck.argtypes = [VP, P(C.c_int)]
    lib.xrtg_top_sah_selftest.argtypes = [VP, VP, VP, C.c_int, C.c_int, P(C.c_int)]
    lib.xrtg_exchange_buffer.argtypes = [VP, C.c_int, C.c_size_t, P(VP)]
    lib.xrtg_ipc_export.argtypes = [VP, VP]
    lib.xrtg_ipc_open.argtypes = [VP, VP, P(VP)]
    lib.xrtg_ipc_close.argtypes = [VP, VP]
    lib.xrtg_reduce_finalize.argtypes = [VP, P(VP), C.c_int, VP, C.c_size_t, C.c_size_t, C.c_float, VP]
    lib.xrtg_render_u8.argtypes = [VP, P(Camera), P(RenderParams), C.c_float, C.c_int, VP, P(Stats)]
    lib._typed = True
    return lib


def host():
    """libxrthost.so (host C++ API + scripting surface); pulls in libxrtgpu.so."""
    gpu()
    lib = _load(HOST_LIB, "host library libxrthost.so")
    if getattr(lib, "_typed", False):
        return lib
    fp = P(C.c_float)
    lib.xrth_last_error.restype = C.c_char_p
    lib.xrth_scene_new.restype = VP
    lib.xrth_scene_free.argtypes = [VP]
    lib.xrth_scene_free.restype = None
    lib.xrth_scene_load_obj.argtypes = [VP, C.c_char_p]
    lib.xrth_scene_add_mesh.argtypes = [VP, C.c_char_p, VP, C.c_int, fp]
    lib.xrth_scene_add_sphere.argtypes = [VP, C.c_char_p, fp, C.c_float, fp]
    lib.xrth_scene_add_sphere_mesh.argtypes = [VP, C.c_char_p, fp, C.c_float, C.c_int, C.c_int, fp]
    lib.xrth_scene_add_quad_light.argtypes = [VP, C.c_char_p, fp, fp, fp, fp, fp]
    lib.xrth_scene_add_triangle_light.argtypes = [VP, C.c_char_p, fp, fp, fp, fp, fp]
    lib.xrth_scene_add_sphere_light.argtypes = [VP, C.c_char_p, fp, C.c_float, fp, fp]
    lib.xrth_scene_add_point_light.argtypes = [VP, C.c_char_p, fp, fp, C.c_float]
    lib.xrth_scene_add_distant_light.argtypes = [VP, C.c_char_p, fp, fp, C.c_float]
    lib.xrth_scene_add_homogeneous_medium.argtypes = [VP, C.c_char_p, C.c_int, C.c_float, fp, fp, fp, fp]
    lib.xrth_scene_add_heterogeneous_medium.argtypes = [VP, C.c_char_p, C.c_float, C.c_int, C.c_int, C.c_int, VP, fp,
                                                        C.c_float, fp, fp, C.c_float]
    lib.xrth_scene_flatten.argtypes = [VP]
    lib.xrth_scene_flatten.restype = P(SceneDesc)
    lib.xrth_camera_make.argtypes = [C.c_float, fp, C.c_float, P(Camera)]
    lib.xrth_render.argtypes = [VP, C.c_float, fp, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32,
                                C.c_uint32, VP, P(Stats)]
    lib._typed = True
    return lib


def _type_oracle(lib, pfx, render_has_stats):
    lib_fn = lambda n: getattr(lib, pfx + n)
    lib_fn("last_error").restype = C.c_char_p
    lib_fn("max_threads").restype = C.c_int
    lib_fn("scene_create").argtypes = [P(SceneDesc), P(VP)]
    lib_fn("scene_destroy").argtypes = [VP]
    lib_fn("scene_destroy").restype = None
    if render_has_stats:
        lib_fn("render").argtypes = [VP, P(Camera), P(RenderParams), C.c_int, C.c_int, VP, P(C.c_double), P(Stats)]
    else:
        lib_fn("render").argtypes = [VP, P(Camera), P(RenderParams), C.c_int, C.c_int, VP, P(C.c_double)]
    lib_fn("trace_primary").argtypes = [VP, P(Camera), C.c_int, C.c_int, C.c_int, VP, VP]
    lib_fn("trace_rays").argtypes = [VP, C.c_int64, VP, VP, VP, C.c_int, VP]
    fp = P(C.c_float)
    lib_fn("kat_sampler").argtypes = [C.c_uint32, C.c_int, fp]
    lib_fn("kat_camera").argtypes = [P(Camera), C.c_float, C.c_float, fp]
    lib_fn("kat_onb").argtypes = [fp, fp]
    lib_fn("kat_light_sample").argtypes = [VP, C.c_int, fp, C.c_uint32, fp]
    lib_fn("kat_lambert_sample").argtypes = [fp, fp, C.c_uint32, fp]
    lib_fn("kat_hg_sample").argtypes = [C.c_float, fp, C.c_uint32, fp]
    for n in ("kat_sampler", "kat_camera", "kat_onb", "kat_light_sample", "kat_lambert_sample", "kat_hg_sample"):
        lib_fn(n).restype = None


def oracle():
    """oracle/libxrtoracle.so — CPU restatement. TEST / BASELINE USE ONLY."""
    lib = _load(ORACLE_LIB, "oracle port libxrtoracle.so")
    if not getattr(lib, "_typed", False):
        _type_oracle(lib, "xrto_", True)
        lib._typed = True
    return lib


def have_reference() -> bool:
    return REF_LIB.exists()


def reference(with_gpu: bool = False):
    """oracle/_ref/libxrtref.so — the compiled reference. TEST / BASELINE USE ONLY. with_gpu: libxrtrefgpu.so, the variant
    that also carries RefGpuRenderer and therefore links libxrtgpu.so (tests/test_gpu_dropin_reference.py)."""
    lib = _load(REF_GPU_LIB, "compiled reference + GPU binding libxrtrefgpu.so") if with_gpu else _load(REF_LIB, "compiled reference libxrtref.so")
    if not getattr(lib, "_typed", False):
        _type_oracle(lib, "xrtref_", False)
        lib.xrtref_object_order.argtypes = [VP, P(C.c_int32), C.c_int]
        lib.xrtref_render_pstl.argtypes = [VP, P(Camera), P(RenderParams), VP, P(C.c_double)]
        if with_gpu:
            lib.xrtref_render_gpu.argtypes = [VP, P(Camera), P(RenderParams), VP]
        lib._typed = True
    return lib
