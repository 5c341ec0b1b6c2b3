"""The host SAH BVH builder (xraytracer_b200/csrc/bvh.cpp) is new functionality relative to the reference (Scene::build() is
an empty hook) and pure host code: xrtg_bvh_selftest builds a tree and checks its structure without a CUDA device."""
import ctypes as C

import numpy as np
import pytest

from xraytracer_b200 import capi, scenes


def selftest(tris9, max_leaf=4):
    tris9 = np.ascontiguousarray(tris9, dtype=np.float32).reshape(-1, 9)
    n_nodes, depth, sah = C.c_int(), C.c_int(), C.c_float()
    lib = capi.gpu()
    rc = lib.xrtg_bvh_selftest(tris9.ctypes.data if len(tris9) else None, len(tris9), max_leaf, C.byref(n_nodes), C.byref(depth), C.byref(sah))
    assert rc == 0, lib.xrtg_last_error().decode()
    return n_nodes.value, depth.value, sah.value


def test_cornell_box_tree(cornell):
    _, desc = cornell
    d = desc.contents
    tris = np.array([list(d.triangles[i].v0) + list(d.triangles[i].v1) + list(d.triangles[i].v2) for i in range(d.n_triangles)], np.float32)
    n_nodes, depth, sah = selftest(tris)
    assert 8 <= n_nodes <= 72 and 3 <= depth <= 16 and sah > 0


@pytest.mark.parametrize("max_leaf", [1, 2, 4])
def test_displaced_sphere_tree(max_leaf):
    tris = scenes.displaced_sphere_tris((278, 200, 280), 150, 96, 96)[:, :9]
    n_nodes, depth, sah = selftest(tris, max_leaf)
    assert depth <= 56 and n_nodes >= len(tris) // (2 * max_leaf)


def test_degenerate_inputs():
    assert selftest(np.zeros((0, 9), np.float32))[0] == 1                      # empty scene: one empty root
    one = np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0]], np.float32)
    assert selftest(one)[1] == 1                                               # single leaf
    same = np.tile(one, (1000, 1))                                             # 1000 identical triangles: centroids coincide
    n_nodes, depth, _ = selftest(same)
    assert depth <= 56
    sliver = np.array([[i, 0, 0, i + 1e-7, 0, 0, i, 1e-7, 0] for i in range(300)], np.float32)   # zero-extent axes
    assert selftest(sliver)[1] <= 56
    rng = np.random.RandomState(0)
    soup = rng.uniform(-1e4, 1e4, (20000, 9)).astype(np.float32)               # huge overlapping triangles
    assert selftest(soup)[1] <= 56


def test_selftest_rejects_bad_arguments():
    lib = capi.gpu()
    assert lib.xrtg_bvh_selftest(None, 3, 4, None, None, None) == -1
    assert lib.xrtg_bvh_selftest(None, 0, 9, None, None, None) == -1


def top_sah(lo, hi, counts=None, by_clusters=1):
    lo = np.ascontiguousarray(lo, dtype=np.float32).reshape(-1, 3)
    hi = np.ascontiguousarray(hi, dtype=np.float32).reshape(-1, 3)
    cnt = None if counts is None else np.ascontiguousarray(counts, dtype=np.uint32)
    depth = C.c_int()
    lib = capi.gpu()
    rc = lib.xrtg_top_sah_selftest(lo.ctypes.data, hi.ctypes.data, None if cnt is None else cnt.ctypes.data, len(lo), by_clusters, C.byref(depth))
    assert rc == 0, lib.xrtg_last_error().decode()
    return depth.value


@pytest.mark.parametrize("by_clusters", [0, 1])
def test_top_level_sweep_sah_builder(by_clusters):
    """The top levels of a device-built tree are split on the HOST (csrc/gpu_build.cu: TopBuilder) from the few thousand clusters
    the PLOC rounds leave standing: random boxes of very different sizes (a few scene-sized 'walls' among small clusters),
    a regular row of equal boxes and 2 000 coincident boxes (equal-cost splits must take the most balanced one, or the depth
    explodes) — every cluster once, node boxes = union of the children's, counts add up (checked inside the library)."""
    rng = np.random.RandomState(5)
    c = rng.uniform(0, 500, (1500, 3))
    r = rng.uniform(0.5, 6.0, (1500, 1))
    lo, hi = c - r, c + r
    lo[:36], hi[:36] = np.array([0, 0, 0]) + rng.uniform(0, 1, (36, 3)), np.array([550, 550, 560]) - rng.uniform(0, 1, (36, 3))   # "walls"
    counts = rng.randint(1, 400, 1500)
    counts[:36] = 1
    d = top_sah(lo, hi, counts, by_clusters)
    assert 11 <= d <= 60, d
    x = np.arange(1024, dtype=np.float32)[:, None] * np.array([[10.0, 0, 0]], np.float32)
    assert top_sah(x, x + 9.0, None, by_clusters) <= 14          # a row of equal boxes: balanced
    same = np.tile(np.array([[100.0, 100, 100]], np.float32), (2000, 1))
    assert top_sah(same, same + 50.0, None, by_clusters) <= 13   # coincident boxes: log2(2000) + 1, not a chain of 2000
    assert top_sah(same[:1], same[:1] + 1.0) == 1
