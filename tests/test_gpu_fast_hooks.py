"""Index parity of the TIMED kernels (pytest -m gpu).

The bit-exact tests in test_gpu_parity.py run the exact instantiation (no FMA, Moeller-Trumbore, k_raygen + k_trace). What
bench.py times is the throughput instantiation: k_primary with its screen-space scissor, the shared-memory small-scene tracer
of k_bounce_small (plane-paired records, hull-pruned occluders), the simple kernels on mid-size scenes and k_trace on the
wide BVH with plane-equation (Havel-Herout) triangle records. XRTG_FLAG_FAST_HOOK routes the parity hooks through exactly
those entry points, and these tests hold their primitive ids / occlusion flags against the oracle
(Scene::intersect scene.cpp:190-200, Scene::occluded :202-211, Mesh::rayTriangleIntersect primitive.cpp:140-168).

Bar: ids identical, except for rays whose reference hit lies within EDGE_EPS (barycentric units) of a triangle edge or whose two
candidate hits are closer than T_EPS (relative) — coplanar duplicates and shared edges, where the plane-equation and the
Moeller-Trumbore arithmetic may legitimately land on either side. Every mismatch is checked individually; the overall
mismatch rate must stay below RATE."""
import ctypes as C

import numpy as np
import pytest

from conftest import require_gpu
from xraytracer_b200 import api, capi, scenes
from test_gpu_random_scenes import random_scene

pytestmark = pytest.mark.gpu

EDGE_EPS = 2e-4   # barycentric distance to the nearest edge below which a different (adjacent) primitive is accepted
T_EPS = 1e-4      # relative difference in t below which two hits count as the same point
RATE = 2e-4       # accepted fraction of rays with differing ids (all of them individually justified)
FAST = capi.FLAG_FAST_HOOK


def edge_distance(h):
    return np.minimum(np.minimum(h["u"], h["v"]), 1.0 - h["u"] - h["v"])


def _inside_triangle(P, tri, tol=2e-3):
    """P lies in the plane of `tri` (3x3 vertices) and inside it, up to `tol` in barycentric units / relative plane distance."""
    e1, e2 = tri[1] - tri[0], tri[2] - tri[0]
    n = np.cross(e1, e2)
    nn = float(np.dot(n, n))
    if not nn > 0:
        return False
    w = P - tri[0]
    if abs(float(np.dot(n, w))) / np.sqrt(nn) > tol * max(1.0, float(np.abs(P).max())):
        return False
    u = float(np.dot(np.cross(w, e2), n)) / nn
    v = float(np.dot(np.cross(e1, w), n)) / nn
    return min(u, v, 1.0 - u - v) >= -tol


def check_closest(fast, ref, what, rate=RATE, sphere_prims=(), rays=None, verts=None, uv_tol=2e-4, dup_lower_id=False):
    """ids equal except at edges / coplanar duplicates; t, u, v of agreeing hits equal to the arithmetic's precision.
    rays = (org, dir) and verts (prim_table) allow the stricter geometric checks: the t error measured ACROSS the surface, and
    coplanar duplicates (a point inside two overlapping triangles of one plane) recognised as such — there the
    reference's pick depends on the last bit of two Moeller-Trumbore evaluations, the plane-paired records share one plane
    and always return the lower id; those are reported, not limited."""
    fast, ref = fast.ravel(), ref.ravel()
    edge_eps = max(EDGE_EPS, uv_tol)   # an edge is as sharp as the barycentrics are
    same = fast["prim"] == ref["prim"]
    both = same & (ref["prim"] >= 0)
    if rays is not None and verts is not None:
        both = both & ~np.isin(ref["prim"], list(sphere_prims))   # (a sphere's limb has no plane; spheres are compared by id)
    dt = np.abs(fast["t"][both] - ref["t"][both])
    if rays is not None and verts is not None:
        # an error dt along the ray moves the hit point by dt * |cos| across the surface: that is what the arithmetic controls
        tri = verts[ref["prim"][both]]
        n = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
        with np.errstate(invalid="ignore", divide="ignore"):
            cosang = np.abs(np.einsum("ij,ij->i", n, rays[1][both])) / np.linalg.norm(n, axis=1)
        cosang = np.where(np.isfinite(cosang), cosang, 1.0)
        # ... to a few ulps of the coordinates involved (origin up to 750 units from the world origin, hits up to ~1500 away)
        scale = np.maximum(np.maximum(np.abs(rays[0][both]).max(axis=1), ref["t"][both]), 1.0)
        err = dt * cosang / scale
        worst = int(np.argmax(err)) if len(err) else 0
        assert np.quantile(err, 0.9999) < 4e-6 and err.max(initial=0.0) < 5e-5, \
            f"{what}: hit points of agreeing hits differ by {err.max()} (across the surface, relative to the coordinates): prim {ref['prim'][both][worst]}, " \
            f"t {fast['t'][both][worst]} vs {ref['t'][both][worst]}, cos {cosang[worst]}"
    else:
        rel_t = dt / np.maximum(ref["t"][both], 1e-3)
        assert np.quantile(rel_t, 0.9999) < 2e-5 and rel_t.max(initial=0.0) < 2e-3, f"{what}: t of agreeing hits differs by {rel_t.max()}"
    tri = both & ~np.isin(ref["prim"], list(sphere_prims))
    du = np.maximum(np.abs(fast["u"][tri] - ref["u"][tri]), np.abs(fast["v"][tri] - ref["v"][tri]))
    assert du.max(initial=0.0) < 50 * uv_tol and (du > uv_tol).mean() < 1e-4, f"{what}: barycentrics differ by {du.max()}"
    n_dup = n_edge = 0
    for i in np.nonzero(~same)[0]:
        f, r = fast[i], ref[i]
        if f["prim"] >= 0 and r["prim"] >= 0:
            # two different primitives. Either they are hit at (numerically) the same distance — a shared edge or coplanar
            # duplicates — or one arithmetic grazes a SILHOUETTE edge of a nearer primitive that the other one just misses: then
            # the nearer of the two hits must lie within EDGE_EPS of an edge of its triangle
            if abs(f["t"] - r["t"]) <= T_EPS * max(r["t"], 1e-3):
                if rays is not None and verts is not None and int(f["prim"]) not in sphere_prims and int(r["prim"]) not in sphere_prims:
                    P = rays[0][i].astype(np.float64) + float(r["t"]) * rays[1][i].astype(np.float64)
                    if _inside_triangle(P, verts[f["prim"]].astype(np.float64)) and _inside_triangle(P, verts[r["prim"]].astype(np.float64)) \
                            and edge_distance(np.array([r], dtype=r.dtype))[0] > edge_eps:
                        n_dup += 1      # coplanar duplicates overlapping at this point
                        # (only the plane-paired block shares ONE plane between duplicates and therefore always returns the lower id;
                        #  per-triangle records differ in the last bits of t exactly like the reference's own evaluations do)
                        assert not dup_lower_id or f["prim"] < r["prim"], f"{what}: ray {i}: coplanar duplicates {f['prim']} / {r['prim']}: the block must return the lower id"
                        continue
                n_edge += 1
                continue
            near = f if f["t"] < r["t"] else r
            n_edge += 1
            if int(near["prim"]) in sphere_prims:
                continue
            assert edge_distance(np.array([near], dtype=near.dtype))[0] < edge_eps, \
                f"{what}: ray {i} hits prim {f['prim']} at {f['t']} vs reference {r['prim']} at {r['t']}, and the nearer hit {near} is not on an edge"
        else:
            # hit vs miss: only on a silhouette edge (or a sphere's limb, where the discriminant changes sign)
            n_edge += 1
            h = f if f["prim"] >= 0 else r
            if int(h["prim"]) in sphere_prims:
                continue
            assert edge_distance(np.array([h], dtype=h.dtype))[0] < edge_eps, f"{what}: ray {i} hit/miss disagreement away from any edge: {f} vs {r}"
    assert n_edge <= rate * len(ref) + 2, f"{what}: {n_edge} of {len(ref)} ids differ at edges"
    return n_edge, n_dup


def check_anyhit(gpu, orc, org, d, tmax, what, src_prim=None, rate=5e-4):
    a = gpu.trace_rays(org, d, tmax, any_hit=True, flags=FAST, src_prim=src_prim)["prim"]
    b = orc.trace_rays(org, d, tmax, any_hit=True)["prim"]
    mis = np.nonzero(a != b)[0]
    assert len(mis) <= rate * len(b) + 2, f"{what}: {len(mis)} of {len(b)} occlusion flags differ"
    # every disagreement must be a borderline ray: the reference's own answer flips when the segment is shortened / lengthened by
    # T_EPS or shifted sideways by 1e-3 of its length (an occluder edge or the segment's end point within eps)
    for i in mis[:200]:
        o, dd, t = org[i], d[i], tmax[i]
        side = np.cross(dd, [0.3, 0.5, 0.81])
        side = side / max(np.linalg.norm(side), 1e-9)
        up = np.cross(dd, side)
        shifts = [(o, t * (1 - 3 * T_EPS)), (o, t * (1 + 3 * T_EPS))] + [(o + s * 2e-3 * max(t, 1.0) * v, t) for s in (-1, 1) for v in (side, up)]
        oo = np.array([s[0] for s in shifts], dtype=np.float32)
        tt = np.array([s[1] for s in shifts], dtype=np.float32)
        flips = orc.trace_rays(oo, np.tile(dd, (len(shifts), 1)).astype(np.float32), tt, any_hit=True)["prim"]
        assert len(set(flips.tolist()) | {int(b[i])}) > 1, f"{what}: ray {i} (tmax {t}) flag {a[i]} vs reference {b[i]}, not a borderline ray"
    return len(mis)


def prim_table(desc):
    """Per global primitive id (assigned by walking objects[] in order, xrtgpu.h): triangle vertices and the geometric normal
    normalize((v1-v0) x (v2-v0)) of primitive.cpp:105 (NaN rows for spheres / boxes)."""
    d = desc.contents if hasattr(desc, "contents") else desc
    n_prims = sum(d.objects[i].count if d.objects[i].kind == capi.OBJ_MESH else 1 for i in range(d.n_objects))
    verts = np.full((n_prims, 3, 3), np.nan, np.float32)
    pid = 0
    tris = np.ctypeslib.as_array(C.cast(d.triangles, C.POINTER(C.c_float)), shape=(max(d.n_triangles, 1), 18))
    for i in range(d.n_objects):
        o = d.objects[i]
        if o.kind == capi.OBJ_MESH:
            verts[pid:pid + o.count] = tris[o.first:o.first + o.count, :9].reshape(-1, 3, 3)
            pid += o.count
        else:
            pid += 1
    with np.errstate(invalid="ignore", divide="ignore"):
        n = np.cross(verts[:, 1] - verts[:, 0], verts[:, 2] - verts[:, 0])
        n = n / np.linalg.norm(n, axis=1, keepdims=True)
    return verts, n.astype(np.float32)


def random_rays(n, lo, hi, seed):
    rng = np.random.RandomState(seed)
    org = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    tmax = rng.uniform(5, 700, n).astype(np.float32)
    return org, d, tmax


def shadow_rays_from_hits(org, d, hits, ng, targets, bias=0.01):
    """NEE-style rays: from every mesh hit, offset by bias * ng like integrator.h:100 / :260, towards random target points."""
    ok = (hits["prim"] >= 0) & np.isfinite(ng[np.maximum(hits["prim"], 0)]).all(axis=1)
    p = org[ok] + hits["t"][ok, None] * d[ok]
    src = hits["prim"][ok]
    o2 = (p + bias * ng[src]).astype(np.float32)
    tgt = targets[np.arange(len(o2)) % len(targets)]
    v = tgt - o2
    dist = np.linalg.norm(v, axis=1)
    keep = dist > 1.0
    return o2[keep], (v[keep] / dist[keep, None]).astype(np.float32), (dist[keep] - bias).astype(np.float32), src[keep].astype(np.int32)


# ---- BASELINE config 1 geometry: the Cornell box through k_primary (scissor on) and the SmallTracer ------------------------

def test_c1_fast_primary_ids_vs_oracle(cornell):
    require_gpu()
    host, desc = cornell
    gpu, orc = api.GpuScene(desc, 0), api.OracleScene(desc)
    W = H = 512
    cam = scenes.make_camera(W, H)
    fast = gpu.trace_primary(cam, W, H, 16, flags=FAST)   # k_primary<.., JITTER>, mt19937 jitter of renderer.cpp:44-47, scissor on
    ref = orc.trace_primary(cam, W, H, 16)
    n, _ = check_closest(fast, ref, "C1 primary")
    print(f"C1 512x512x16 = {ref.size} primary rays through k_primary: {n} id mismatches (all on edges)")
    # 16:9 frame (the scissor now removes 36 % of the columns): pixels outside it must be misses in the reference too
    W, H = 640, 360
    cam = scenes.make_camera(W, H)
    fast = gpu.trace_primary(cam, W, H, 4, flags=FAST)
    ref = orc.trace_primary(cam, W, H, 4)
    check_closest(fast, ref, "C1 16:9 primary")
    # supplied jitter, box seen off-centre and from inside
    for c2w, fov in (([-1, 0, 0, 0, 0, 1, 0, 0, 0, 0, -1, 0, 1400.0, 900.0, -2000.0, 1], 50.0), ([-1, 0, 0, 0, 0, 1, 0, 0, 0, 0, -1, 0, 278.0, 274.4, 200.0, 1], 60.0)):
        cam = scenes.make_camera(320, 180, c2w, fov)
        jit = np.random.RandomState(5).random_sample((320 * 180 * 2, 2)).astype(np.float32)
        check_closest(gpu.trace_primary(cam, 320, 180, 2, jitter=jit, flags=FAST), orc.trace_primary(cam, 320, 180, 2, jitter=jit), "C1 moved camera")


def test_c1_small_scene_tracer_ids_and_occlusion_vs_oracle(cornell):
    """Secondary rays of the Cornell box go through SmallTracer (groupedClosest / groupedAnyHit on the plane-paired block, the
    occluder section without the six hull-pruned walls): closest-hit ids and NEE occlusion flags against the oracle."""
    require_gpu()
    host, desc = cornell
    gpu, orc = api.GpuScene(desc, 0), api.OracleScene(desc)
    info = gpu.info()
    assert info["small_records_all"] > 0 and info["small_records_occ"] < info["small_records_all"]   # plane-paired block with pruned occluders in use
    org, d, tmax = random_rays(400000, 5, 545, 3)
    ref = orc.trace_rays(org, d)
    fast = gpu.trace_rays(org, d, flags=FAST)
    verts, ng = prim_table(desc)
    n, dup = check_closest(fast, ref, "Cornell secondary closest", rays=(org, d), verts=verts, dup_lower_id=True)
    # NEE rays exactly as the integrators build them: hit + 0.01 * ng towards points on the quad light
    rng = np.random.RandomState(1)
    light = np.stack([rng.uniform(213, 343, 4096), np.full(4096, 548.0), rng.uniform(227, 332, 4096)], 1).astype(np.float32)
    o2, d2, t2, src = shadow_rays_from_hits(org, d, ref, ng, light)
    m = check_anyhit(gpu, orc, o2, d2, t2, "Cornell NEE shadow rays", src_prim=src)
    # and unrelated random segments inside the box (both ends inside the hull)
    m2 = check_anyhit(gpu, orc, org, d, np.minimum(tmax, np.where(ref["prim"] >= 0, ref["t"] * 0.999, tmax)).astype(np.float32), "Cornell random segments")
    print(f"Cornell SmallTracer: {n} id mismatches at edges + {dup} coplanar-duplicate picks (floor vs block footprints, rays from inside the blocks) of "
          f"{len(org)}, {m} + {m2} occlusion mismatches of {len(o2)} + {len(org)} (all borderline)")


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_random_small_scenes_fast_ids_vs_oracle(seed):
    """Random small scenes (random winding, coplanar duplicates, a degenerate triangle, spheres on odd seeds): primary rays
    through k_primary, incoherent rays and NEE rays through SmallTracer."""
    require_gpu()
    host, cam = random_scene(seed)
    desc = host.flatten()
    gpu, orc = api.GpuScene(desc, 0), api.OracleScene(desc)
    d_ = desc.contents if hasattr(desc, "contents") else desc
    verts, ng = prim_table(desc)
    spheres = [i for i in range(len(ng)) if not np.isfinite(ng[i]).all()]
    W, H = 160, 120
    check_closest(gpu.trace_primary(cam, W, H, 4, flags=FAST), orc.trace_primary(cam, W, H, 4), f"random scene {seed} primary", rate=1e-3, sphere_prims=spheres)
    org, d, tmax = random_rays(100000, 2, 98, seed)
    ref = orc.trace_rays(org, d)
    check_closest(gpu.trace_rays(org, d, flags=FAST), ref, f"random scene {seed} closest", rate=1e-3, sphere_prims=spheres, rays=(org, d), verts=verts)
    rng = np.random.RandomState(seed)
    light = np.stack([rng.uniform(35, 65, 512), np.full(512, 99.5), rng.uniform(35, 65, 512)], 1).astype(np.float32)
    o2, d2, t2, src = shadow_rays_from_hits(org, d, ref, ng, light)
    check_anyhit(gpu, orc, o2, d2, t2, f"random scene {seed} NEE", src_prim=src, rate=2e-3)


def test_outward_wound_room_self_shadowing():
    """A closed room whose walls are wound so that their geometric normals point OUT of the room. The reference never flips ng
    towards the ray: a shadow ray from such a wall starts 0.01 BEHIND it and the wall shadows itself. The
    hull-pruned occluder section must not lose that: flagged primitives use the unpruned section."""
    require_gpu()
    def quad(a, e1, e2, flip):
        a, e1, e2 = (np.array(x, np.float32) for x in (a, e1, e2))
        t = [[a, a + e1, a + e2], [a + e1, a + e1 + e2, a + e2]]
        if flip:
            t = [[x[0], x[2], x[1]] for x in t]
        return [np.concatenate([np.concatenate(x), np.zeros(9, np.float32)]) for x in t]
    results = {}
    for flip in (False, True):
        s = scenes.HostScene()
        tris = []
        # normals of the un-flipped quads point INTO the room [0,100]^3
        tris += quad((0, 0, 0), (0, 0, 100), (100, 0, 0), flip)        # floor, +y
        tris += quad((0, 100, 0), (100, 0, 0), (0, 0, 100), flip)      # ceiling, -y
        tris += quad((0, 0, 100), (0, 100, 0), (100, 0, 0), flip)      # back, -z
        tris += quad((0, 0, 0), (0, 100, 0), (0, 0, 100), flip)        # left, +x
        tris += quad((100, 0, 0), (0, 0, 100), (0, 100, 0), flip)      # right, -x
        arr = np.array(tris, np.float32)
        for t in arr:   # shading normals = geometric normals
            n = np.cross(t[3:6] - t[0:3], t[6:9] - t[0:3]); n /= np.linalg.norm(n)
            t[9:12] = t[12:15] = t[15:18] = n
        s.add_mesh("room", arr, (0.7, 0.7, 0.7))
        blk = []
        blk += quad((30, 0, 30), (0, 0, 30), (30, 0, 0), False)   # a slab above the floor
        for t in blk:
            t[1] = t[4] = t[7] = 20.0
            t[9:18] = np.tile([0, 1, 0], 3)
        s.add_mesh("slab", np.array(blk, np.float32), (0.3, 0.6, 0.9))
        s.add_quad_light("QuadLight", (65, 99.5, 35), (65, 99.5, 65), (35, 99.5, 35), (40.0, 40.0, 40.0))   # faces down
        desc = s.flatten()
        gpu, orc = api.GpuScene(desc, 0), api.OracleScene(desc)
        info = gpu.info()
        assert info["small_records_all"] > 0
        assert (info["small_flagged"] > 0) == flip, info
        cam = scenes.make_camera(96, 72, [-1, 0, 0, 0, 0, 1, 0, 0, 0, 0, -1, 0, 50.0, 50.0, -140.0, 1], 50.0)
        a, _ = gpu.render(cam, 96, 72, 2048, capi.INT_DIRECT, 1, seed=3)
        b, _, _ = orc.render(cam, 96, 72, 512, capi.INT_DIRECT, 1)
        results[flip] = float(b.mean())
        assert abs(float(a.mean()) - float(b.mean())) < 0.005 * float(b.mean()), (flip, a.mean(), b.mean())
        verts, ng = prim_table(desc)
        org, d, _ = random_rays(60000, 2, 98, 7)
        ref = orc.trace_rays(org, d)
        rng = np.random.RandomState(2)
        light = np.stack([rng.uniform(35, 65, 512), np.full(512, 99.5), rng.uniform(35, 65, 512)], 1).astype(np.float32)
        o2, d2, t2, src = shadow_rays_from_hits(org, d, ref, ng, light)
        check_anyhit(gpu, orc, o2, d2, t2, f"room flip={flip}", src_prim=src)
    assert results[True] < 0.8 * results[False]   # outward-wound walls receive no direct light in the reference (cos clamps to 0)


# ---- mid-size scene: shallow BVH with more than 64 triangles -> the simple run-to-completion kernels --------------------------

def test_mid_size_scene_simple_kernels_fast_ids_vs_oracle():
    require_gpu()
    extra = lambda h: h.add_mesh("tess", scenes.displaced_sphere_tris((278, 200, 280), 150, 7, 7), (0.75, 0.75, 0.75))
    s = scenes.cornell_box("quad", extra=extra)
    desc = s.flatten()
    gpu, orc = api.GpuScene(desc, 0), api.OracleScene(desc)
    info = gpu.info()
    assert info["n_triangles"] > 64 and info["n_bvh_nodes"] <= 512
    cam = scenes.make_camera(320, 180)
    check_closest(gpu.trace_primary(cam, 320, 180, 4, flags=FAST), orc.trace_primary(cam, 320, 180, 4), "mid-size primary")
    org, d, tmax = random_rays(100000, 20, 530, 11)
    ref = orc.trace_rays(org, d)
    verts, _ = prim_table(desc)
    check_closest(gpu.trace_rays(org, d, flags=FAST), ref, "mid-size closest", rays=(org, d), verts=verts)
    check_anyhit(gpu, orc, org, d, tmax, "mid-size any-hit")


# ---- deep BVHs: k_trace on the wide tree + plane-equation triangle records ------------------------------------------------------

def test_deep_bvh_fast_ids_vs_oracle():
    """3.2 k-triangle scene (> 512 nodes): raygen + k_trace over the wide tree and the plane-equation triangles, against the oracle."""
    require_gpu()
    extra = lambda h: (h.add_mesh("tess", scenes.displaced_sphere_tris((278, 200, 280), 150, 40, 40), (0.75, 0.75, 0.75)),
                       h.add_sphere("ball", (120.0, 80.0, 400.0), 60.0, (0.5, 0.5, 0.5)))
    s = scenes.cornell_box("quad", extra=extra)
    desc = s.flatten()
    gpu, orc = api.GpuScene(desc, 0), api.OracleScene(desc)
    info = gpu.info()
    assert info["n_bvh_nodes"] > 512 and info["wide_arity"] >= 4
    verts, ng = prim_table(desc)
    spheres = [i for i in range(len(ng)) if not np.isfinite(ng[i]).all()]
    cam = scenes.make_camera(320, 180)
    n0, _ = check_closest(gpu.trace_primary(cam, 320, 180, 4, flags=FAST), orc.trace_primary(cam, 320, 180, 4), "deep primary", sphere_prims=spheres, uv_tol=1e-3)
    org, d, tmax = random_rays(60000, 20, 530, 9)
    ref = orc.trace_rays(org, d)
    n1, _ = check_closest(gpu.trace_rays(org, d, flags=FAST), ref, "deep closest", sphere_prims=spheres, rays=(org, d), verts=verts, uv_tol=1e-3)
    m = check_anyhit(gpu, orc, org, d, tmax, "deep any-hit")
    rng = np.random.RandomState(1)
    light = np.stack([rng.uniform(213, 343, 4096), np.full(4096, 548.0), rng.uniform(227, 332, 4096)], 1).astype(np.float32)
    o2, d2, t2, src = shadow_rays_from_hits(org, d, ref, ng, light)
    m2 = check_anyhit(gpu, orc, o2, d2, t2, "deep NEE")
    print(f"3.2k-triangle scene through k_trace (wide tree, plane-equation triangles): {n0} + {n1} id mismatches, {m} + {m2} occlusion mismatches")


def test_c4_full_size_fast_ids_vs_exact_traversal():
    """BASELINE config 4 geometry (999,698 triangles; edges ~0.7 units): the fast instantiation on the wide tree against the exact
    instantiation (which equals brute force and the oracle, test_gpu_parity.py) — primary rays of the 1080p frame (every 4th
    pixel in x and y), incoherent rays, occlusion; plus a handful of rays against the oracle's own brute force."""
    require_gpu()
    s = scenes.cornell_mesh_scene(707, 707)
    desc = s.flatten()
    gpu = api.GpuScene(desc, 0)
    W, H = 480, 270
    cam = scenes.make_camera(W, H)
    jit = np.random.RandomState(4).random_sample((W * H * 2, 2)).astype(np.float32)
    ref = gpu.trace_primary(cam, W, H, 2, jitter=jit)
    # (triangle edges are ~0.7 units at distances of ~1000: barycentrics carry ~1e-3 of absolute noise in either arithmetic)
    n0, _ = check_closest(gpu.trace_primary(cam, W, H, 2, jitter=jit, flags=FAST), ref, "c4 primary", rate=2e-3, uv_tol=5e-3)
    org, d, tmax = random_rays(200000, 30, 520, 21)
    ref2 = gpu.trace_rays(org, d)
    verts, _ = prim_table(desc)
    n1, _ = check_closest(gpu.trace_rays(org, d, flags=FAST), ref2, "c4 closest", rate=2e-3, rays=(org, d), verts=verts, uv_tol=5e-3)
    a = gpu.trace_rays(org, d, tmax, any_hit=True, flags=FAST)["prim"]
    b = gpu.trace_rays(org, d, tmax, any_hit=True)["prim"]
    assert (a != b).sum() <= 1e-3 * len(b)
    orc = api.OracleScene(desc)
    check_closest(gpu.trace_rays(org[:64], d[:64], flags=FAST), orc.trace_rays(org[:64], d[:64]), "c4 vs oracle brute force", rate=0.05, uv_tol=5e-3)
    print(f"c4 999,698 triangles: {n0} of {ref.size} primary and {n1} of {len(org)} incoherent ids differ, {(a != b).sum()} occlusion flags")
