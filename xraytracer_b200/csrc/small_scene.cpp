// small_scene.cpp — groups the triangles of a small scene by supporting plane (see small_scene.h).
#include "small_scene.h"
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

bool buildSmallBlock(const float* ftrisId, int nTris, std::vector<float>& block, SmallBlockInfo* info)
{
    block.clear();
    if (nTris <= 0 || nTris > 64) return false;
    std::vector<Plane> planes{};
    planes.resize(size_t(nTris));
    std::vector<int> all, occ;
    for (int t = 0; t < nTris; ++t) {
        planes[size_t(t)] = planeOf(ftrisId + 16 * size_t(t));
        all.push_back(t);
        int flags;
        std::memcpy(&flags, ftrisId + 16 * size_t(t) + 13, 4);
        if ((flags & 1) == 0) occ.push_back(t);
    }
    SmallBlockInfo bi;
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
