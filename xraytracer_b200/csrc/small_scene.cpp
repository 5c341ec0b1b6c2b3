// small_scene.cpp — groups the triangles of a small scene by supporting plane (see small_scene.h).
#include "small_scene.h"
#include <algorithm>
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
    const size_t offRecs = at + 1, offIds = offRecs + 5 * size_t(nRecords);
    const size_t end = offIds + (2 * size_t(nRecords) + 3) / 4;
    out.resize(4 * end, 0.f);
    put(out, at, asFloat(nRecords), 0.f, asFloat(int(offRecs)), asFloat(int(offIds)));
    size_t r = 0;
    for (const auto& g : groups) {
        const float* r0 = rec + 16 * size_t(g[0]); // the plane words of the group's first triangle serve every record of the plane
        for (size_t k = 0; k < g.size(); k += 2, ++r) {
            const float* a = rec + 16 * size_t(g[k]);
            put(out, offRecs + 5 * r, r0[0], r0[1], r0[2], r0[3]);
            put(out, offRecs + 5 * r + 1, a[4], a[5], a[6], a[7]);
            put(out, offRecs + 5 * r + 2, a[8], a[9], a[10], a[11]);
            out[4 * offIds + 2 * r] = a[12];
            if (k + 1 < g.size()) {
                const float* b = rec + 16 * size_t(g[k + 1]);
                put(out, offRecs + 5 * r + 3, b[4], b[5], b[6], b[7]);
                put(out, offRecs + 5 * r + 4, b[8], b[9], b[10], b[11]);
                out[4 * offIds + 2 * r + 1] = b[12];
            }
            else {
                put(out, offRecs + 5 * r + 3, 0.f, 0.f, 0.f, -1.f);
                put(out, offRecs + 5 * r + 4, 0.f, 0.f, 0.f, -1.f);
                out[4 * offIds + 2 * r + 1] = asFloat(-1);
            }
        }
    }
    return end - at;
}

} // namespace

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
    const bool ok = buildSmallBlock(recs.data(), n, block, &bi, tris9, 3 * n); // hull = the triangles' own vertices
    if (info) *info = bi;
    if (!ok) return 1;
    auto asInt = [](float f) { int i; std::memcpy(&i, &f, 4); return i; };
    const int total = asInt(block[2]);
    if (size_t(total) * 4 != block.size() || total > kSmallBlockMaxF4) return -1;
    for (int sec = 0; sec < 2; ++sec) {
        const int off = asInt(block[size_t(sec)]);
        const int nRec = asInt(block[4 * size_t(off)]), offRecs = asInt(block[4 * size_t(off) + 2]), offIds = asInt(block[4 * size_t(off) + 3]);
        std::vector<int> seen(size_t(n), 0);
        for (int r = 0; r < nRec; ++r) {
            const float* R = &block[4 * (size_t(offRecs) + 5 * size_t(r))];
            for (int k = 0; k < 2; ++k) {
                const int id = asInt(block[4 * size_t(offIds) + 2 * size_t(r) + size_t(k)]);
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
            if (seen[size_t(t)] != (expect ? 1 : 0) && !pruned) return -1;
        }
    }
    return 0;
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

bool buildSmallBlock(const float* ftrisId, int nTris, std::vector<float>& block, SmallBlockInfo* info, const float* hullPoints, int nHullPoints)
{
    block.clear();
    if (nTris <= 0 || nTris > 64) return false;
    std::vector<Plane> planes{};
    planes.resize(size_t(nTris));
    std::vector<int> all, occ;
    SmallBlockInfo bi;
    for (int t = 0; t < nTris; ++t) {
        planes[size_t(t)] = planeOf(ftrisId + 16 * size_t(t));
        all.push_back(t);
        int flags;
        std::memcpy(&flags, ftrisId + 16 * size_t(t) + 13, 4);
        if ((flags & 1) != 0) continue; // emitter proxy: never an occluder (scene.cpp:206)
        if (hullPoints && planeBoundsPoints(ftrisId + 16 * size_t(t), hullPoints, nHullPoints)) { ++bi.nPruned; continue; }
        occ.push_back(t);
    }
    block.assign(4, 0.f);
    const size_t offAll = 1;
    int planesOcc = 0;
    const size_t nAll = emitSection(ftrisId, all, planes, offAll, block, bi.nRecordsAll, bi.nPlanesAll);
    const size_t offOcc = offAll + nAll;
    const size_t nOcc = emitSection(ftrisId, occ, planes, offOcc, block, bi.nRecordsOcc, planesOcc);
    const size_t total = offOcc + nOcc;
    put(block, 0, asFloat(int(offAll)), asFloat(int(offOcc)), asFloat(int(total)), 0.f);
    if (info) *info = bi;
    // pays only if most triangles find a coplanar partner: ~48 instructions per record against ~33 per triangle of the plain loop
    const bool pays = 48.0 * bi.nRecordsAll < 0.9 * 33.0 * nTris;
    if (!pays || total > size_t(kSmallBlockMaxF4)) { block.clear(); return false; }
    return true;
}

} // namespace xrt
