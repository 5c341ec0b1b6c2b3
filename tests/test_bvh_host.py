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
