// small_scene.h — plane-grouped triangle block for scenes of at most 64 triangles (host side, see small_scene.cpp).
//
// Coplanar triangles (the two halves of every OBJ quad, the Cornell floor with both block footprints) are tested in PAIRS that
// share ONE ray/plane intersection; each triangle then costs only its two barycentric plane equations. The block is copied
// verbatim into shared memory by k_bounce_small (wavefront.cuh) — layout in float4 units:
//   [0]            int4 (offAll, offOcc, totalF4, offOccFull)
//   section:       int4 (nRecords, nPairs, offRecords, offIds)          offsets from the start of the block
//                  one record = a supporting plane and two triangles in it, 20 floats:
//                      N.xyz | d ,  A: n1.xyz | d1 , n2.xyz | d2 ,  B: n1 | d1 , n2 | d2
//                  (an unpaired triangle gets a B that no point is inside of: 0 | -1 , 0 | -1)
//                  10 float4 per PAIR of records, component-interleaved — float 2c + j is component c of record j — so that
//                  the kernel computes two records per packed fp32x2 instruction; an odd count is completed by a dummy
//                  record whose plane no ray can hit (N = 0)
//                  1 int4 per pair: primitive ids (A, B) of record 0, (A, B) of record 1; -1 for padding
// Records of one plane are consecutive, carry bit-identical plane words and are in primitive-id order, so "strictly smaller t
// wins" between records and "A before B" inside one reproduce the reference's first-wins rule for coplanar duplicates.
// `All` holds every mesh triangle (closest hit), `Occ` only the triangles that are not emitter proxies (Scene::occluded skips
// those, scene.cpp:206) and whose plane does not have the whole scene on one side (hull pruning, below), `OccFull` every
// triangle that is not an emitter proxy: the integrators start shadow rays at hit + bias * ng with ng never flipped towards the
// ray, so a hit on a triangle whose normal points out of the hull starts its shadow ray BEHIND a hull plane and
// the reference lets that plane shadow it. Such primitives are flagged (kMetaShadowOutside, api.cu) and a warp with a
// flagged lane tests the full section; `OccFull` = `Occ` when nothing was pruned.
#pragma once
#include <cstdint>
#include <vector>

namespace xrt {

constexpr int kSmallBlockMaxF4 = 704; // shared-memory budget of the kernel (11 KB)

struct SmallBlockInfo {
    int nRecordsAll = 0, nRecordsOcc = 0, nPlanesAll = 0, nPruned = 0;
    int nCovered = 0; // triangles left out of the closest-hit section because earlier coplanar triangles cover them
};

// Plane-equation record of one triangle (16 floats: N|d, n1|d1, n2|d2, id|flags|0|0 — the last two as int bits), computed in
// double: N = e1 x e2, d = N.v0 give t = (d - N.o)/(N.dir); u = n1.P + d1 with n1 = (e2 x N)/|N|^2, v = n2.P + d2 with
// n2 = (N x e1)/|N|^2 (Havel & Herout 2010). flags bit 0 = emitter proxy.
void makePlaneRecord(const float v0[3], const float v1[3], const float v2[3], int id, int flags, float rec[16]);

// Host-only consistency check of buildSmallBlock (no CUDA device needed), see xrtg_small_scene_selftest in xrtgpu.h.
int smallBlockSelftest(const float* tris9, const int* emitterFlags, int n, SmallBlockInfo* info);

// ftrisId: 4 floats x 4 per triangle (N|d, n1|d1, n2|d2, id|flags) in primitive-id order. Returns false (block left empty) when
// grouping does not pay (fewer than 4 triangles saved) or the block would not fit.
//
// hullPoints (3 floats each, may be null): every point a shadow ray can start or end at — all triangle vertices, the corners of
// the spheres' bounding boxes, point-light positions. A triangle whose plane has ALL of them on one closed side (a wall of a
// closed room: the scene lies inside its half-space) can never be crossed by a segment between two such points, so it is
// left out of the occluder section. Pass null when a light is infinitely far away (DistantLight) — its shadow rays leave the hull.
bool buildSmallBlock(const float* ftrisId, int nTris, std::vector<float>& block, SmallBlockInfo* info, const float* hullPoints = nullptr,
                     int nHullPoints = 0, std::vector<int>* pruned = nullptr, const float* tris9 = nullptr);
// tris9 (9 floats per triangle, same order as ftrisId; may be null): enables the removal of coplanar duplicates that lie inside
// the union of earlier triangles of the same plane — see coveredCoplanarDuplicates() in small_scene.cpp.
// the triangles of `tris` (ascending indices into ftrisId / tris9) that lie inside the union of EARLIER coplanar triangles of the list
std::vector<int> coveredCoplanarDuplicates(const float* ftrisId, const float* tris9, const std::vector<int>& tris);
// signed distance of `point` to the plane of record `rec` (16 floats), positive on the side the triangle's normal e1 x e2 points to
double planeSignedDistance(const float* rec, const float* point);
// true if all points lie on one closed side of the plane of record `rec` (16 floats), within 1e-5 of the points' extent
bool planeBoundsPoints(const float* rec, const float* points, int nPoints);

} // namespace xrt
