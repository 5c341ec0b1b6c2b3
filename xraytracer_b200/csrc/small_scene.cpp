// small_scene.cpp — groups the triangles of a small scene by supporting plane (see small_scene.h).
#include "small_scene.h"
#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>

namespace xrt {
namespace {

struct Plane {
    double n[3], delta; // unit normal with a canonical sign, n.x = delta
    bool valid;
};

Plane planeOf(const float* rec)
{
    Plane p{};
    const double N[3] = {rec[0], rec[1], rec[2]};
    const double len = std::sqrt(N[0] * N[0] + N[1] * N[1] + N[2] * N[2]);
    p.valid = std::isfinite(len) && len > 0.0 && std::isfinite(double(rec[3]));
    if (!p.valid) return p;
    double sgn = 1.0;
    for (int a = 0; a < 3; ++a)
        if (std::fabs(N[a]) > 1e-9 * len) { sgn = N[a] < 0 ? -1.0 : 1.0; break; }
    for (int a = 0; a < 3; ++a) p.n[a] = sgn * N[a] / len;
    p.delta = sgn * double(rec[3]) / len;
    return p;
}

bool samePlane(const Plane& a, const Plane& b)
{
    if (!a.valid || !b.valid) return false;
    for (int k = 0; k < 3; ++k)
        if (std::fabs(a.n[k] - b.n[k]) > 1e-6) return false;
    return std::fabs(a.delta - b.delta) <= 2e-6 * std::fmax(1.0, std::fabs(a.delta));
}

void put(std::vector<float>& v, size_t f4, float a, float b, float c, float d)
{
    v[4 * f4] = a; v[4 * f4 + 1] = b; v[4 * f4 + 2] = c; v[4 * f4 + 3] = d;
}
float asFloat(int i) { float f; std::memcpy(&f, &i, 4); return f; }

// One logical record: 20 floats (N|d, A: n1|d1, n2|d2, B: n1|d1, n2|d2) + the two primitive ids.
struct Rec {
    float f[20];
    int id[2];
};
Rec dummyRecord() // a plane no ray can hit (N = 0 -> det = 0) holding two triangles no point is inside of
{
    Rec r{};
    r.f[3] = 1.f;
    r.f[7] = r.f[11] = r.f[15] = r.f[19] = -1.f;
    r.id[0] = r.id[1] = -1;
    return r;
}
// Records are stored in PAIRS, component-interleaved, so that the kernel handles two records per packed fp32x2 instruction
// (FFMA2 / FADD2 / FMUL2 of sm_100): float 2c + j of a pair's 40 floats is component c of its record j.
void writeRecord(std::vector<float>& out, size_t offRecs, size_t offIds, size_t r, const Rec& rec)
{
    const size_t pair = r / 2, j = r % 2;
    for (int c = 0; c < 20; ++c) out[4 * (offRecs + 10 * pair) + 2 * size_t(c) + j] = rec.f[c];
    out[4 * (offIds + pair) + 2 * j] = asFloat(rec.id[0]);
    out[4 * (offIds + pair) + 2 * j + 1] = asFloat(rec.id[1]);
}
Rec readRecord(const std::vector<float>& blk, size_t offRecs, size_t offIds, size_t r)
{
    Rec rec{};
    const size_t pair = r / 2, j = r % 2;
    for (int c = 0; c < 20; ++c) rec.f[c] = blk[4 * (offRecs + 10 * pair) + 2 * size_t(c) + j];
    for (int k = 0; k < 2; ++k) std::memcpy(&rec.id[k], &blk[4 * (offIds + pair) + 2 * j + size_t(k)], 4);
    return rec;
}

// one section; returns its size in float4
size_t emitSection(const float* rec, const std::vector<int>& tris, const std::vector<Plane>& planes, size_t at, std::vector<float>& out, int& nRecords,
                   int& nPlanes)
{
    std::vector<std::vector<int>> groups;
    for (int t : tris) {
        bool placed = false;
        for (auto& g : groups)
            if (samePlane(planes[size_t(g[0])], planes[size_t(t)])) { g.push_back(t); placed = true; break; }
        if (!placed) groups.push_back({t});
    }
    nPlanes = int(groups.size());
    nRecords = 0;
    for (auto& g : groups) nRecords += int((g.size() + 1) / 2);
    const size_t nPairs = (size_t(nRecords) + 1) / 2;
    const size_t offRecs = at + 1, offIds = offRecs + 10 * nPairs;
    const size_t end = offIds + nPairs;
    out.resize(4 * end, 0.f);
    put(out, at, asFloat(nRecords), asFloat(int(nPairs)), asFloat(int(offRecs)), asFloat(int(offIds)));
    size_t r = 0;
    for (const auto& g : groups) {
        const float* r0 = rec + 16 * size_t(g[0]); // the plane words of the group's first triangle serve every record of the plane
        for (size_t k = 0; k < g.size(); k += 2, ++r) {
            Rec R = dummyRecord();
            std::memcpy(R.f, r0, 16);
            const float* a = rec + 16 * size_t(g[k]);
            std::memcpy(R.f + 4, a + 4, 32);
            std::memcpy(&R.id[0], a + 12, 4);
            if (k + 1 < g.size()) {
                const float* b = rec + 16 * size_t(g[k + 1]);
                std::memcpy(R.f + 12, b + 4, 32);
                std::memcpy(&R.id[1], b + 12, 4);
            }
            writeRecord(out, offRecs, offIds, r, R);
        }
    }
    if (r % 2) writeRecord(out, offRecs, offIds, r, dummyRecord()); // odd count: the last pair's second half
    return end - at;
}

// ---- coplanar duplicates that can never matter (the Cornell floor shape carries both block footprints) ----
// 2-D convex polygon clipping in the plane's dominant projection, double precision.
struct P2 { double x, y; };
double polyArea(const std::vector<P2>& p)
{
    double a = 0.0;
    for (size_t i = 0; i < p.size(); ++i) { const P2& u = p[i]; const P2& v = p[(i + 1) % p.size()]; a += u.x * v.y - v.x * u.y; }
    return 0.5 * a;
}
// splits convex polygon `poly` by the line through a -> b: `in` = the part on the left (inside for a CCW clip polygon), `out` = the rest
void splitByEdge(const std::vector<P2>& poly, P2 a, P2 b, std::vector<P2>& in, std::vector<P2>& out)
{
    in.clear(); out.clear();
    auto side = [&](const P2& p) { return (b.x - a.x) * (p.y - a.y) - (b.y - a.y) * (p.x - a.x); };
    for (size_t i = 0; i < poly.size(); ++i) {
        const P2& p = poly[i]; const P2& q = poly[(i + 1) % poly.size()];
        const double sp = side(p), sq = side(q);
        if (sp >= 0) in.push_back(p);
        if (sp <= 0) out.push_back(p);
        if ((sp > 0 && sq < 0) || (sp < 0 && sq > 0)) {
            const double t = sp / (sp - sq);
            const P2 x{p.x + t * (q.x - p.x), p.y + t * (q.y - p.y)};
            in.push_back(x); out.push_back(x);
        }
    }
}
// true if triangle T lies inside the union of the triangles `cover` (all in one plane), up to 1e-9 of T's area
bool coveredBy(const P2 T[3], const std::vector<std::array<P2, 3>>& cover)
{
    std::vector<std::vector<P2>> pieces{{T[0], T[1], T[2]}};
    if (polyArea(pieces[0]) < 0) std::reverse(pieces[0].begin(), pieces[0].end());
    const double areaT = std::fabs(polyArea(pieces[0]));
    if (!(areaT > 0.0)) return false; // degenerate triangles are left alone (their records reject every ray anyway)
    for (const auto& L0 : cover) {
        std::array<P2, 3> L = L0;
        if (polyArea({L[0], L[1], L[2]}) < 0) std::swap(L[1], L[2]); // CCW
        std::vector<std::vector<P2>> next;
        for (const auto& piece : pieces) {
            std::vector<P2> cur = piece, in, out;
            for (int e = 0; e < 3 && cur.size() >= 3; ++e) { // what leaves through edge e stays a piece; the rest goes on
                splitByEdge(cur, L[size_t(e)], L[size_t((e + 1) % 3)], in, out);
                if (out.size() >= 3 && std::fabs(polyArea(out)) > 1e-12 * areaT) next.push_back(out);
                cur = in;
            }
            // `cur` is now the part of the piece inside L: covered, dropped
        }
        pieces.swap(next);
        if (pieces.empty()) return true;
    }
    double left = 0.0;
    for (const auto& piece : pieces) left += std::fabs(polyArea(piece));
    return left <= 1e-9 * areaT;
}

} // namespace

// Triangles of `tris` (indices into the record / vertex arrays, ascending ids) that can be DROPPED from a section without
// changing any result: a triangle that lies inside the union of coplanar triangles of the same section which precede it.
// Closest hit: the records of one plane share bit-identical plane words, so a point inside both is always credited to the
// earlier (lower-id) triangle — the reference's own first-wins rule (scene.cpp:193-197) — and the later one never wins. Any hit:
// a segment that crosses the dropped triangle crosses one of the covering triangles at the same point.
std::vector<int> coveredCoplanarDuplicates(const float* ftrisId, const float* tris9, const std::vector<int>& tris)
{
    std::vector<int> dropped;
    if (!tris9) return dropped;
    std::vector<Plane> planes;
    for (int t : tris) planes.push_back(planeOf(ftrisId + 16 * size_t(t)));
    for (size_t k = 1; k < tris.size(); ++k) {
        if (!planes[k].valid) continue;
        // dominant projection axis of the plane normal
        int ax = 0;
        for (int a = 1; a < 3; ++a) if (std::fabs(planes[k].n[a]) > std::fabs(planes[k].n[ax])) ax = a;
        const int ux = (ax + 1) % 3, uy = (ax + 2) % 3;
        auto proj = [&](int t, P2 out[3]) { for (int v = 0; v < 3; ++v) out[v] = P2{tris9[9 * size_t(t) + 3 * v + ux], tris9[9 * size_t(t) + 3 * v + uy]}; };
        std::vector<std::array<P2, 3>> cover;
        for (size_t j = 0; j < k; ++j) {
            if (!samePlane(planes[j], planes[k]) || std::find(dropped.begin(), dropped.end(), tris[j]) != dropped.end()) continue;
            P2 L[3];
            proj(tris[j], L);
            cover.push_back({L[0], L[1], L[2]});
        }
        if (cover.empty()) continue;
        P2 T[3];
        proj(tris[k], T);
        if (coveredBy(T, cover)) dropped.push_back(tris[k]);
    }
    return dropped;
}

void makePlaneRecord(const float v0[3], const float v1[3], const float v2[3], int id, int flags, float rec[16])
{
    const double E1[3] = {double(v1[0]) - v0[0], double(v1[1]) - v0[1], double(v1[2]) - v0[2]};
    const double E2[3] = {double(v2[0]) - v0[0], double(v2[1]) - v0[1], double(v2[2]) - v0[2]};
    const double N[3] = {E1[1] * E2[2] - E1[2] * E2[1], E1[2] * E2[0] - E1[0] * E2[2], E1[0] * E2[1] - E1[1] * E2[0]};
    const double nn = N[0] * N[0] + N[1] * N[1] + N[2] * N[2];
    const double n1[3] = {(E2[1] * N[2] - E2[2] * N[1]) / nn, (E2[2] * N[0] - E2[0] * N[2]) / nn, (E2[0] * N[1] - E2[1] * N[0]) / nn};
    const double n2[3] = {(N[1] * E1[2] - N[2] * E1[1]) / nn, (N[2] * E1[0] - N[0] * E1[2]) / nn, (N[0] * E1[1] - N[1] * E1[0]) / nn};
    const double dN = N[0] * v0[0] + N[1] * v0[1] + N[2] * v0[2];
    const double d1 = -(n1[0] * v0[0] + n1[1] * v0[1] + n1[2] * v0[2]);
    const double d2 = -(n2[0] * v0[0] + n2[1] * v0[1] + n2[2] * v0[2]);
    const float out[16] = {float(N[0]),  float(N[1]),  float(N[2]),  float(dN), float(n1[0]), float(n1[1]), float(n1[2]), float(d1),
                           float(n2[0]), float(n2[1]), float(n2[2]), float(d2), asFloat(id),  asFloat(flags), 0.f,        0.f};
    std::memcpy(rec, out, sizeof(out));
}

// Checks a block against the triangles it was built from: every triangle exactly once in `All`, every non-emitter exactly once
// in `Occ`, padding slots reject every point, each triangle's centroid lies on its record's plane and inside its own barycentric
// equations, records of one plane carry identical plane words. Returns 0 (valid block), 1 (valid, but pairing does not pay so
// no block is emitted) or -1 (inconsistent).
int smallBlockSelftest(const float* tris9, const int* emitterFlags, int n, SmallBlockInfo* info)
{
    std::vector<float> recs(size_t(std::max(n, 1)) * 16);
    for (int t = 0; t < n; ++t)
        makePlaneRecord(tris9 + 9 * size_t(t), tris9 + 9 * size_t(t) + 3, tris9 + 9 * size_t(t) + 6, t, emitterFlags ? emitterFlags[t] : 0, &recs[16 * size_t(t)]);
    std::vector<float> block;
    SmallBlockInfo bi;
    const bool ok = buildSmallBlock(recs.data(), n, block, &bi, tris9, 3 * n, nullptr, tris9); // hull = the triangles' own vertices
    if (info) *info = bi;
    if (!ok) return 1;
    // coplanar duplicates the builder may leave out: inside the union of earlier coplanar triangles of the same list. Checked
    // independently of the clipping code: a few interior sample points of each dropped triangle must lie inside an earlier,
    // kept, coplanar triangle.
    std::vector<int> allList, occList;
    for (int t = 0; t < n; ++t) { allList.push_back(t); if (!(emitterFlags && (emitterFlags[t] & 1))) occList.push_back(t); }
    const std::vector<int> covAll = coveredCoplanarDuplicates(recs.data(), tris9, allList), covOcc = coveredCoplanarDuplicates(recs.data(), tris9, occList);
    for (int pass = 0; pass < 2; ++pass) {
        const std::vector<int>& cov = pass ? covOcc : covAll;
        const std::vector<int>& list = pass ? occList : allList;
        for (int t : cov) {
            const float* v = tris9 + 9 * size_t(t);
            const double w3[4][3] = {{1 / 3.0, 1 / 3.0, 1 / 3.0}, {0.8, 0.1, 0.1}, {0.1, 0.8, 0.1}, {0.1, 0.1, 0.8}};
            for (const auto& w : w3) {
                const double P[3] = {w[0] * v[0] + w[1] * v[3] + w[2] * v[6], w[0] * v[1] + w[1] * v[4] + w[2] * v[7], w[0] * v[2] + w[1] * v[5] + w[2] * v[8]};
                bool inside = false;
                for (int j : list) {
                    if (j >= t) break;
                    if (std::find(cov.begin(), cov.end(), j) != cov.end()) continue;
                    const float* a = &recs[16 * size_t(j)];
                    const double u = a[4] * P[0] + a[5] * P[1] + a[6] * P[2] + a[7], vv = a[8] * P[0] + a[9] * P[1] + a[10] * P[2] + a[11];
                    const double nl = std::sqrt(double(a[0]) * a[0] + double(a[1]) * a[1] + double(a[2]) * a[2]);
                    const double dist = nl > 0 ? std::fabs(a[0] * P[0] + a[1] * P[1] + a[2] * P[2] - a[3]) / nl : 1e30;
                    if (u >= -1e-6 && vv >= -1e-6 && u + vv <= 1.0 + 1e-6 && dist <= 1e-3 * std::fmax(1.0, std::fabs(P[0]) + std::fabs(P[1]) + std::fabs(P[2]))) { inside = true; break; }
                }
                if (!inside) return -1;
            }
        }
    }
    if (bi.nCovered != int(covAll.size())) return -1;
    auto asInt = [](float f) { int i; std::memcpy(&i, &f, 4); return i; };
    const int total = asInt(block[2]);
    if (size_t(total) * 4 != block.size() || total > kSmallBlockMaxF4) return -1;
    for (int sec = 0; sec < 2; ++sec) {
        const int off = asInt(block[size_t(sec)]);
        const int nRec = asInt(block[4 * size_t(off)]), nPairs = asInt(block[4 * size_t(off) + 1]), offRecs = asInt(block[4 * size_t(off) + 2]),
                  offIds = asInt(block[4 * size_t(off) + 3]);
        if (nPairs != (nRec + 1) / 2) return -1;
        std::vector<int> seen(size_t(n), 0);
        for (int r = 0; r < 2 * nPairs; ++r) {
            const Rec rec = readRecord(block, size_t(offRecs), size_t(offIds), size_t(r));
            const float* R = rec.f;
            if (r >= nRec) { // the dummy half of the last pair: a plane with N = 0
                if (R[0] != 0.f || R[1] != 0.f || R[2] != 0.f || rec.id[0] != -1 || rec.id[1] != -1) return -1;
                continue;
            }
            for (int k = 0; k < 2; ++k) {
                const int id = rec.id[k];
                const float* a = R + 4 + 8 * k;
                if (id < 0) { // padding: n1 = n2 = 0, d1 = d2 = -1
                    if (k == 0 || a[0] != 0.f || a[1] != 0.f || a[2] != 0.f || a[3] != -1.f || a[7] != -1.f) return -1;
                    continue;
                }
                if (id >= n || seen[size_t(id)]++) return -1;
                if (sec == 1 && emitterFlags && (emitterFlags[id] & 1)) return -1;
                const float* v = tris9 + 9 * size_t(id);
                const double c[3] = {(double(v[0]) + v[3] + v[6]) / 3, (double(v[1]) + v[4] + v[7]) / 3, (double(v[2]) + v[5] + v[8]) / 3};
                const double nlen = std::sqrt(double(R[0]) * R[0] + double(R[1]) * R[1] + double(R[2]) * R[2]);
                if (!(nlen > 0.0) || !std::isfinite(nlen)) continue; // degenerate triangle: its record rejects every ray
                const double dist = (R[0] * c[0] + R[1] * c[1] + R[2] * c[2] - R[3]) / nlen; // centroid to the record's plane
                double scale = 1.0;
                for (int q = 0; q < 3; ++q) scale = std::fmax(scale, std::fabs(c[q]));
                if (!(std::fabs(dist) <= 1e-4 * scale)) return -1;
                const double u = a[0] * c[0] + a[1] * c[1] + a[2] * c[2] + a[3], w = a[4] * c[0] + a[5] * c[1] + a[6] * c[2] + a[7];
                if (!(std::fabs(u - 1.0 / 3) < 1e-3 && std::fabs(w - 1.0 / 3) < 1e-3)) return -1;
            }
        }
        for (int t = 0; t < n; ++t) {
            const bool emitter = emitterFlags && (emitterFlags[t] & 1);
            // a non-emitter may be missing from the occluder section only if the whole scene lies on one side of its plane
            const bool pruned = sec == 1 && !emitter && seen[size_t(t)] == 0 && planeBoundsPoints(&recs[16 * size_t(t)], tris9, 3 * n);
            const bool expect = sec == 0 || !emitter;
            const std::vector<int>& cov = sec == 0 ? covAll : covOcc;
            const bool covered = seen[size_t(t)] == 0 && std::find(cov.begin(), cov.end(), t) != cov.end();
            if (seen[size_t(t)] != (expect ? 1 : 0) && !pruned && !covered) return -1;
        }
    }
    return 0;
}

double planeSignedDistance(const float* rec, const float* point)
{
    const double N[3] = {rec[0], rec[1], rec[2]};
    const double len = std::sqrt(N[0] * N[0] + N[1] * N[1] + N[2] * N[2]);
    if (!(len > 0.0)) return 0.0;
    return (N[0] * point[0] + N[1] * point[1] + N[2] * point[2] - double(rec[3])) / len;
}

bool planeBoundsPoints(const float* rec, const float* points, int nPoints)
{
    const Plane p = planeOf(rec);
    if (!p.valid || nPoints <= 0) return false;
    double ext = 0.0;
    for (int k = 0; k < 3 * nPoints; ++k) ext = std::fmax(ext, std::fabs(double(points[k])));
    const double tol = 1e-5 * std::fmax(ext, 1e-6);
    bool pos = false, neg = false;
    for (int k = 0; k < nPoints; ++k) {
        const double sd = p.n[0] * points[3 * k] + p.n[1] * points[3 * k + 1] + p.n[2] * points[3 * k + 2] - p.delta;
        if (!(std::fabs(sd) < 1e300)) return false; // NaN / inf coordinates: keep the triangle
        pos = pos || sd > tol;
        neg = neg || sd < -tol;
    }
    return !(pos && neg);
}

bool buildSmallBlock(const float* ftrisId, int nTris, std::vector<float>& block, SmallBlockInfo* info, const float* hullPoints, int nHullPoints,
                     std::vector<int>* pruned, const float* tris9)
{
    block.clear();
    if (nTris <= 0 || nTris > 64) return false;
    std::vector<Plane> planes{};
    planes.resize(size_t(nTris));
    std::vector<int> all, occ, occFull;
    SmallBlockInfo bi;
    if (pruned) pruned->clear();
    for (int t = 0; t < nTris; ++t) {
        planes[size_t(t)] = planeOf(ftrisId + 16 * size_t(t));
        all.push_back(t);
        int flags;
        std::memcpy(&flags, ftrisId + 16 * size_t(t) + 13, 4);
        if ((flags & 1) != 0) continue; // emitter proxy: never an occluder (scene.cpp:206)
        occFull.push_back(t);
        if (hullPoints && planeBoundsPoints(ftrisId + 16 * size_t(t), hullPoints, nHullPoints)) {
            ++bi.nPruned;
            if (pruned) pruned->push_back(t);
            continue;
        }
        occ.push_back(t);
    }
    // coplanar duplicates covered by earlier triangles of the same section can never change a result: leave them out
    auto dropCovered = [&](std::vector<int>& list) {
        const std::vector<int> gone = coveredCoplanarDuplicates(ftrisId, tris9, list);
        for (int t : gone) list.erase(std::find(list.begin(), list.end(), t));
        return int(gone.size());
    };
    bi.nCovered = dropCovered(all);
    dropCovered(occ);
    dropCovered(occFull);
    block.assign(4, 0.f);
    const size_t offAll = 1;
    int planesOcc = 0, recordsFull = 0;
    const size_t nAll = emitSection(ftrisId, all, planes, offAll, block, bi.nRecordsAll, bi.nPlanesAll);
    const size_t offOcc = offAll + nAll;
    const size_t nOcc = emitSection(ftrisId, occ, planes, offOcc, block, bi.nRecordsOcc, planesOcc);
    // third section: EVERY occluder, for shadow rays that start outside the hull (see small_scene.h); the pruned section itself
    // when nothing was pruned
    size_t offFull = offOcc, total = offOcc + nOcc;
    if (bi.nPruned > 0) {
        offFull = total;
        total += emitSection(ftrisId, occFull, planes, offFull, block, recordsFull, planesOcc);
    }
    put(block, 0, asFloat(int(offAll)), asFloat(int(offOcc)), asFloat(int(total)), asFloat(int(offFull)));
    if (info) *info = bi;
    // pays only if most triangles find a coplanar partner: ~31 instructions per record (packed, two records at a time) against
    // ~33 per triangle of the plain loop
    const bool pays = 31.0 * bi.nRecordsAll < 0.8 * 33.0 * nTris;
    if (!pays || total > size_t(kSmallBlockMaxF4)) { block.clear(); return false; }
    return true;
}

} // namespace xrt
