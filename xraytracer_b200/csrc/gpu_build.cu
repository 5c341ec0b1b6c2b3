// gpu_build.cu — device-side scene ingest, PLOC BVH build and eight-child collapse (gpu_build.h).
//
// Why: for the 1 M-triangle workload the host ingest (pinned staging, per-triangle records, binned-SAH build, two collapses) is
// 0.5-1.3 s against a 73 ms render step. Here the raw triangles go up once and every derived array is produced in HBM:
//
//   k_ingest          raw xrtg_triangle -> trisId / ftrisId / prims (bit-identical to the host loop in api.cu)
//   k_boxes, k_morton, cub radix sort (a sorting primitive outside the render path)
//   PLOC rounds       k_ploc_nn (nearest neighbour by merged surface area within +-radius in Morton order, boxes staged in shared
//                     memory) -> k_ploc_flag (mutual pairs) -> inclusive scan -> k_ploc_emit (merged nodes in scan order:
//                     the tree is deterministic, no atomics); every merge also decides, bottom-up, whether the new subtree is
//                     cheaper as ONE leaf of <= 4 triangles than as an inner node (SAH with the children's final costs)
//   k_leafpos         leaf-order position of every triangle = sum of the left siblings' counts on the way to the root
//   k_emit_nodes      BvhNode records (children's padded boxes in the parent, root = record 0) for the two-child traversal
//   k_collapse8       level-synchronous collapse into eight-child quantised nodes, one thread per wide node
//
// Like every BVH here it is new functionality relative to the reference (Scene::build() is an empty hook, scene.h:22-24) and is
// held to "same answer as the brute-force loops" (tests/test_gpu_build.py).
#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cuda_runtime.h>
#include "gpu_build.h"

namespace xrt {
namespace {

constexpr int kB = 256;
constexpr uint32_t kLeafFlag = 0x80000000u;
constexpr int kMaxTreeDepth = 120; // k_trace: 24 shared + 104 local stack entries, one push per level

// ------------------------------------------------------------------------------------------------------------------ ingest

__device__ __forceinline__ float4 mk4(const float* p, float w) { return make_float4(p[0], p[1], p[2], w); }

__global__ void k_ingest(const xrtg_triangle* __restrict__ raw, const MeshRange* __restrict__ ranges, int nRanges, int nTris,
                         float4* __restrict__ trisId, float4* __restrict__ ftrisId, float4* __restrict__ prims)
{
    const int tk = blockIdx.x * blockDim.x + threadIdx.x;
    if (tk >= nTris) return;
    int lo = 0, hi = nRanges - 1; // last range with triStart <= tk
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (ranges[mid].triStart <= tk) lo = mid; else hi = mid - 1;
    }
    const MeshRange r = ranges[lo];
    const int k = tk - r.triStart;
    const xrtg_triangle t = raw[size_t(r.srcFirst) + size_t(k)];
    const int id = r.id0 + k;
    // e1 = v1 - v0, e2 = v2 - v0 (primitive.cpp:142-143), ng = normalize(e1 x e2) (primitive.cpp:105): the reference's fp32
    // operation order, every operation rounded separately (no FMA contraction)
    float e1[3], e2[3], c[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) { e1[a] = __fsub_rn(t.v1[a], t.v0[a]); e2[a] = __fsub_rn(t.v2[a], t.v0[a]); }
    c[0] = __fsub_rn(__fmul_rn(e1[1], e2[2]), __fmul_rn(e1[2], e2[1]));
    c[1] = __fsub_rn(__fmul_rn(e1[2], e2[0]), __fmul_rn(e1[0], e2[2]));
    c[2] = __fsub_rn(__fmul_rn(e1[0], e2[1]), __fmul_rn(e1[1], e2[0]));
    const float len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(c[0], c[0]), __fmul_rn(c[1], c[1])), __fmul_rn(c[2], c[2])));
    const float ng[3] = {__fdiv_rn(c[0], len), __fdiv_rn(c[1], len), __fdiv_rn(c[2], len)};
    trisId[3 * size_t(tk)] = mk4(t.v0, __int_as_float(id));
    trisId[3 * size_t(tk) + 1] = mk4(e1, __int_as_float(r.emitter));
    trisId[3 * size_t(tk) + 2] = mk4(e2, 0.f);
    // plane-equation record (small_scene.cpp: makePlaneRecord), double arithmetic with every operation rounded separately
    {
        const double v0[3] = {double(t.v0[0]), double(t.v0[1]), double(t.v0[2])};
        const double E1[3] = {__dsub_rn(double(t.v1[0]), v0[0]), __dsub_rn(double(t.v1[1]), v0[1]), __dsub_rn(double(t.v1[2]), v0[2])};
        const double E2[3] = {__dsub_rn(double(t.v2[0]), v0[0]), __dsub_rn(double(t.v2[1]), v0[1]), __dsub_rn(double(t.v2[2]), v0[2])};
        auto cross = [](const double* a, const double* b, double* o) {
            o[0] = __dsub_rn(__dmul_rn(a[1], b[2]), __dmul_rn(a[2], b[1]));
            o[1] = __dsub_rn(__dmul_rn(a[2], b[0]), __dmul_rn(a[0], b[2]));
            o[2] = __dsub_rn(__dmul_rn(a[0], b[1]), __dmul_rn(a[1], b[0]));
        };
        auto dot = [](const double* a, const double* b) { return __dadd_rn(__dadd_rn(__dmul_rn(a[0], b[0]), __dmul_rn(a[1], b[1])), __dmul_rn(a[2], b[2])); };
        double N[3], x1[3], x2[3];
        cross(E1, E2, N);
        const double nn = dot(N, N);
        cross(E2, N, x1);
        cross(N, E1, x2);
        const double n1[3] = {__ddiv_rn(x1[0], nn), __ddiv_rn(x1[1], nn), __ddiv_rn(x1[2], nn)};
        const double n2[3] = {__ddiv_rn(x2[0], nn), __ddiv_rn(x2[1], nn), __ddiv_rn(x2[2], nn)};
        const double dN = dot(N, v0), d1 = -dot(n1, v0), d2 = -dot(n2, v0);
        float4* f = ftrisId + 4 * size_t(tk);
        f[0] = make_float4(float(N[0]), float(N[1]), float(N[2]), float(dN));
        f[1] = make_float4(float(n1[0]), float(n1[1]), float(n1[2]), float(d1));
        f[2] = make_float4(float(n2[0]), float(n2[1]), float(n2[2]), float(d2));
        f[3] = make_float4(__int_as_float(id), __int_as_float(r.emitter), 0.f, 0.f);
    }
    float4* p = prims + 4 * size_t(id);
    p[0] = mk4(t.n0, ng[0]);
    p[1] = mk4(t.n1, ng[1]);
    p[2] = mk4(t.n2, ng[2]);
    p[3] = make_float4(r.albedo[0], r.albedo[1], r.albedo[2], __uint_as_float(r.meta));
}

__global__ void k_scatter_prims(const float4* __restrict__ recs, const int* __restrict__ ids, int count, float4* __restrict__ prims)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const size_t id = size_t(ids[k]);
#pragma unroll
    for (int j = 0; j < 4; ++j) prims[4 * id + j] = recs[4 * size_t(k) + j];
}

// --------------------------------------------------------------------------------------------------- boxes, Morton codes

__device__ __forceinline__ float3 f3min(float3 a, float3 b) { return make_float3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
__device__ __forceinline__ float3 f3max(float3 a, float3 b) { return make_float3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
// order-preserving float <-> uint so that atomicMin/atomicMax work on floats of either sign
__device__ __forceinline__ uint32_t fenc(float f) { const uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__host__ __device__ inline float fdec(uint32_t u)
{
    const uint32_t v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(v);
#else
    float f;
    std::memcpy(&f, &v, 4);
    return f;
#endif
}

__global__ void k_boxes(const float4* __restrict__ trisId, uint32_t n, float4* __restrict__ lo, float4* __restrict__ hi, uint32_t* bounds)
{
    __shared__ uint32_t s[6];
    if (threadIdx.x < 6) s[threadIdx.x] = threadIdx.x < 3 ? 0xffffffffu : 0u;
    __syncthreads();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float4 a = trisId[3 * size_t(i)], e1 = trisId[3 * size_t(i) + 1], e2 = trisId[3 * size_t(i) + 2];
        const float3 v0 = make_float3(a.x, a.y, a.z);
        // v0 + e1 may differ from the original vertex by an ulp; the conservative padding (2^-15 of the scene magnitude) is five
        // orders of magnitude larger
        const float3 v1 = make_float3(a.x + e1.x, a.y + e1.y, a.z + e1.z), v2 = make_float3(a.x + e2.x, a.y + e2.y, a.z + e2.z);
        const float3 l = f3min(v0, f3min(v1, v2)), h = f3max(v0, f3max(v1, v2));
        lo[i] = make_float4(l.x, l.y, l.z, 0.f);
        hi[i] = make_float4(h.x, h.y, h.z, 0.f);
        atomicMin(&s[0], fenc(l.x)); atomicMin(&s[1], fenc(l.y)); atomicMin(&s[2], fenc(l.z));
        atomicMax(&s[3], fenc(h.x)); atomicMax(&s[4], fenc(h.y)); atomicMax(&s[5], fenc(h.z));
    }
    __syncthreads();
    if (threadIdx.x < 3) atomicMin(&bounds[threadIdx.x], s[threadIdx.x]);
    else if (threadIdx.x < 6) atomicMax(&bounds[threadIdx.x], s[threadIdx.x]);
}

// 21 bits per axis -> 63-bit Morton code
__device__ __forceinline__ unsigned long long expand21(unsigned long long v)
{
    v &= 0x1fffffull;
    v = (v | (v << 32)) & 0x1f00000000ffffull;
    v = (v | (v << 16)) & 0x1f0000ff0000ffull;
    v = (v | (v << 8)) & 0x100f00f00f00f00full;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}

__global__ void k_morton(const float4* __restrict__ lo, const float4* __restrict__ hi, uint32_t n, const uint32_t* __restrict__ bounds,
                         unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float3 smin = make_float3(fdec(bounds[0]), fdec(bounds[1]), fdec(bounds[2]));
    const float3 smax = make_float3(fdec(bounds[3]), fdec(bounds[4]), fdec(bounds[5]));
    const float4 l = lo[i], h = hi[i];
    const float cx = 0.5f * (l.x + h.x), cy = 0.5f * (l.y + h.y), cz = 0.5f * (l.z + h.z);
    const float ex = fmaxf(smax.x - smin.x, 1e-30f), ey = fmaxf(smax.y - smin.y, 1e-30f), ez = fmaxf(smax.z - smin.z, 1e-30f);
    const float kMax = 2097151.f;
    const unsigned long long x = (unsigned long long)(fminf(fmaxf((cx - smin.x) / ex, 0.f) * 2097152.f, kMax));
    const unsigned long long y = (unsigned long long)(fminf(fmaxf((cy - smin.y) / ey, 0.f) * 2097152.f, kMax));
    const unsigned long long z = (unsigned long long)(fminf(fmaxf((cz - smin.z) / ez, 0.f) * 2097152.f, kMax));
    keys[i] = (expand21(x) << 2) | (expand21(y) << 1) | expand21(z);
    vals[i] = i;
}

// -------------------------------------------------------------------------------------------------------------------- PLOC
// Node ids: 0..n-1 = the triangles in Morton order, n..2n-2 = merged nodes in creation order (the last one is the root).
// lo[id] = box.lo | SAH cost of the subtree (absolute, in half-areas); hi[id] = box.hi | triangle count | kLeafFlag if the
// subtree is one leaf.

__device__ __forceinline__ float halfArea(float lx, float ly, float lz, float hx, float hy, float hz)
{
    const float dx = hx - lx, dy = hy - ly, dz = hz - lz;
    return dx * dy + dy * dz + dz * dx;
}

__global__ void k_ploc_init(const uint32_t* __restrict__ order, const float4* __restrict__ triLo, const float4* __restrict__ triHi, int n,
                            float4* __restrict__ lo, float4* __restrict__ hi, int2* __restrict__ child, int* __restrict__ cid)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t t = order[i];
    const float4 l = triLo[t], h = triHi[t];
    lo[i] = make_float4(l.x, l.y, l.z, halfArea(l.x, l.y, l.z, h.x, h.y, h.z)); // one triangle test
    hi[i] = make_float4(h.x, h.y, h.z, __uint_as_float(1u | kLeafFlag));
    child[i] = make_int2(-1, -1);
    cid[i] = i;
}

// Total order on candidate PAIRS with equal merged area: a hash of the unordered pair, then the pair itself. Both ends of a pair
// compute the same key, so the globally smallest pair under (area, key) is always mutual — every round merges at least once — and
// on tie-heavy input (a regular grid of equal triangles: "ties to the lower index" would merge ONE pair per round there) the
// hash picks a random matching instead of a chain.
__device__ __forceinline__ unsigned long long pairKey(int a, int b)
{
    const uint32_t lo = uint32_t(min(a, b)), hi = uint32_t(max(a, b));
    uint32_t h = lo * 0x9E3779B1u ^ hi * 0x85EBCA77u;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return ((unsigned long long)h << 32) | (unsigned long long)(lo * 128u + (hi - lo));
}

// The loop state lives in device memory so that several rounds can be enqueued without a host round trip: state[r & 1] =
// (clusters m, next node id) read by round r, state[(r + 1) & 1] written by it. A round that finds m <= top does nothing but
// carry the state (and the cluster list) forward, so enqueueing a few rounds too many is harmless.
struct PlocState {
    int m, nodeBase, rounds, pad;
};

// nearest neighbour of cluster i among clusters i-R..i+R: smallest surface area of the merged box, ties by pairKey
__global__ void k_ploc_nn(const int* __restrict__ cid, const PlocState* __restrict__ stIn, int top, const float4* __restrict__ lo,
                          const float4* __restrict__ hi, int R, int* __restrict__ nn)
{
    extern __shared__ float sbox[]; // 6 x (kB + 2R)
    const int m = stIn->m;
    if (m <= top || int(blockIdx.x) * kB >= m) return;
    const int W = kB + 2 * R;
    const int base = blockIdx.x * kB - R;
    for (int t = threadIdx.x; t < W; t += kB) {
        const int g = base + t;
        float4 l = make_float4(0.f, 0.f, 0.f, 0.f), h = l;
        if (g >= 0 && g < m) { const int id = cid[g]; l = lo[id]; h = hi[id]; }
        sbox[t] = l.x; sbox[W + t] = l.y; sbox[2 * W + t] = l.z;
        sbox[3 * W + t] = h.x; sbox[4 * W + t] = h.y; sbox[5 * W + t] = h.z;
    }
    __syncthreads();
    const int i = blockIdx.x * kB + threadIdx.x;
    if (i >= m) return;
    const int me = threadIdx.x + R;
    const float lx = sbox[me], ly = sbox[W + me], lz = sbox[2 * W + me], hx = sbox[3 * W + me], hy = sbox[4 * W + me], hz = sbox[5 * W + me];
    float best = FLT_MAX;
    int bestJ = -1;
    unsigned long long bestKey = ~0ull;
    bool haveKey = false;
    const int o0 = max(-R, -i), o1 = min(R, m - 1 - i);
    for (int o = o0; o <= o1; ++o) {
        if (o == 0) continue;
        const int t = me + o;
        float d = halfArea(fminf(lx, sbox[t]), fminf(ly, sbox[W + t]), fminf(lz, sbox[2 * W + t]), fmaxf(hx, sbox[3 * W + t]),
                           fmaxf(hy, sbox[4 * W + t]), fmaxf(hz, sbox[5 * W + t]));
        d = (d == d) ? fminf(d, FLT_MAX) : FLT_MAX; // NaN / inf boxes: still a valid (worst) candidate
        if (d < best || bestJ < 0) { best = d; bestJ = i + o; haveKey = false; }
        else if (d == best) {
            if (!haveKey) { bestKey = pairKey(i, bestJ); haveKey = true; }
            const unsigned long long k = pairKey(i, i + o);
            if (k < bestKey) { bestKey = k; bestJ = i + o; }
        }
    }
    nn[i] = bestJ;
}

// packed[i] = (cluster i survives the round) | (cluster i is the lower half of a mutual pair, i.e. creates a node) << 32;
// 0 beyond the live clusters (the scan runs over the host's upper bound of m)
__global__ void k_ploc_flag(const int* __restrict__ nn, const PlocState* __restrict__ stIn, int top, int mUpper, unsigned long long* __restrict__ packed)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= mUpper) return;
    const int m = stIn->m;
    if (m <= top) return; // (nothing reads packed / incl in a carried-forward round)
    if (i >= m) { packed[i] = 0ull; return; }
    const int j = nn[i];
    const bool mutual = nn[j] == i;
    const unsigned long long valid = (mutual && i > j) ? 0ull : 1ull, merge = (mutual && i < j) ? 1ull : 0ull;
    packed[i] = valid | (merge << 32);
}

__global__ void k_ploc_emit(const int* __restrict__ cid, const int* __restrict__ nn, const unsigned long long* __restrict__ incl,
                            const PlocState* __restrict__ stIn, PlocState* __restrict__ stOut, int top, int mUpper, float ct, int maxLeaf,
                            float4* __restrict__ lo, float4* __restrict__ hi, int2* __restrict__ child, int* __restrict__ parent, int* __restrict__ cidOut)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= mUpper) return;
    const int m = stIn->m, nodeBase = stIn->nodeBase;
    if (m <= top) { // converged: carry the state and the cluster list forward
        if (i < m) cidOut[i] = cid[i];
        if (i == 0) *stOut = *stIn;
        return;
    }
    if (i >= m) return;
    const unsigned long long s = incl[i];
    if (i == m - 1) { // the totals of the round = the inclusive sums at the last live cluster
        PlocState o;
        o.m = int(uint32_t(s)); o.nodeBase = nodeBase + int(uint32_t(s >> 32)); o.rounds = stIn->rounds + 1; o.pad = 0;
        *stOut = o;
    }
    const int j = nn[i];
    const bool mutual = nn[j] == i;
    if (mutual && i > j) return; // absorbed by its partner
    const int pos = int(uint32_t(s)) - 1;
    if (!mutual) { cidOut[pos] = cid[i]; return; }
    const int id = nodeBase + int(uint32_t(s >> 32)) - 1;
    const int a = cid[i], b = cid[j];
    const float4 la = lo[a], ha = hi[a], lb = lo[b], hb = hi[b];
    const float lx = fminf(la.x, lb.x), ly = fminf(la.y, lb.y), lz = fminf(la.z, lb.z);
    const float hx = fmaxf(ha.x, hb.x), hy = fmaxf(ha.y, hb.y), hz = fmaxf(ha.z, hb.z);
    const float A = halfArea(lx, ly, lz, hx, hy, hz);
    const uint32_t c = (__float_as_uint(ha.w) & ~kLeafFlag) + (__float_as_uint(hb.w) & ~kLeafFlag);
    const float costInner = ct * A + la.w + lb.w, costLeaf = A * float(c);
    const bool leaf = c <= uint32_t(maxLeaf) && costLeaf <= costInner;
    lo[id] = make_float4(lx, ly, lz, leaf ? costLeaf : costInner);
    hi[id] = make_float4(hx, hy, hz, __uint_as_float(c | (leaf ? kLeafFlag : 0u)));
    child[id] = make_int2(a, b);
    parent[a] = id;
    parent[b] = id;
    cidOut[pos] = id;
}

// top of the tree: the surviving clusters' boxes go to the host, the nodes built there come back
__global__ void k_gather_clusters(const int* __restrict__ cid, int m, const float4* __restrict__ lo, const float4* __restrict__ hi, float4* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    out[2 * i] = lo[cid[i]];
    out[2 * i + 1] = hi[cid[i]];
}
__global__ void k_scatter_parents(const int2* __restrict__ idParent, int count, int* __restrict__ parent)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) parent[idParent[i].x] = idParent[i].y;
}

// Leaf-order position of every triangle (depth-first order of the tree: a subtree's triangles are contiguous, so a collapsed leaf is
// a range), first[] of every node whose leftmost triangle this is, the depth of the tree above the collapsed leaves and their number.
__global__ void k_leafpos(int n, int root, const int* __restrict__ parent, const int2* __restrict__ child, const float4* __restrict__ hi,
                          int* __restrict__ leafPos, int* __restrict__ first, int* __restrict__ stats /*0: depth, 1: leaves*/)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int node = i, off = 0, level = 0, topLevel = 0, offAtTop = 0; // a single triangle is a leaf: the topmost leaf-flagged ancestor so far is itself
    while (node != root) {
        const int p = parent[node];
        const int2 c = child[p];
        if (c.y == node) off += int(__float_as_uint(hi[c.x].w) & ~kLeafFlag);
        node = p;
        ++level;
        if (__float_as_uint(hi[p].w) & kLeafFlag) { topLevel = level; offAtTop = off; }
    }
    leafPos[i] = off;
    atomicMax(&stats[0], level - topLevel + 1);
    if (offAtTop == 0) atomicAdd(&stats[1], 1);
    node = i;
    first[node] = off;
    while (node != root) {
        const int p = parent[node];
        if (child[p].x != node) break;
        node = p;
        first[node] = off;
    }
}

// BvhNode record of merged node id -> index (2n-2) - id, so that the root is record 0. A child that is a (collapsed) leaf is
// referenced by its triangle range, an inner child by its record index. Records of nodes below a collapsed leaf are written too
// (harmless, never referenced).
__global__ void k_emit_nodes(int n, const float4* __restrict__ lo, const float4* __restrict__ hi, const int2* __restrict__ child,
                             const int* __restrict__ first, float pad, BvhNode* __restrict__ out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n - 1) return;
    const int id = n + k;
    const int2 c = child[id];
    BvhNode nd;
    const int cs[2] = {c.x, c.y};
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const float4 l = lo[cs[s]], h = hi[cs[s]];
        float* blo = s ? nd.lo1 : nd.lo0;
        float* bhi = s ? nd.hi1 : nd.hi0;
        blo[0] = l.x - pad; blo[1] = l.y - pad; blo[2] = l.z - pad;
        bhi[0] = h.x + pad; bhi[1] = h.y + pad; bhi[2] = h.z + pad;
        const uint32_t w = __float_as_uint(h.w);
        const bool leaf = (w & kLeafFlag) != 0;
        (s ? nd.child1 : nd.child0) = leaf ? first[cs[s]] : (2 * n - 2) - cs[s];
        (s ? nd.count1 : nd.count0) = leaf ? int(w & ~kLeafFlag) : 0;
    }
    out[(2 * n - 2) - id] = nd;
}

__global__ void k_scatter_tris(const float4* __restrict__ src, const uint32_t* __restrict__ order, const int* __restrict__ leafPos, uint32_t n,
                               float4* __restrict__ dst, int f4PerTri)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = order[i];
    const size_t d = size_t(leafPos[i]);
    for (int k = 0; k < f4PerTri; ++k) dst[size_t(f4PerTri) * d + k] = src[size_t(f4PerTri) * s + k];
}

// ---------------------------------------------------------------------------------------------------- eight-child collapse

struct Slot8 {
    float lo[3], hi[3];
    int32_t child, count;
};

__device__ __forceinline__ void slotsOf(const BvhNode* __restrict__ nodes, int n, Slot8 s[2])
{
    const float4* p = reinterpret_cast<const float4*>(nodes + n);
    const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    const int4 d = __ldg(reinterpret_cast<const int4*>(p + 3));
    s[0].lo[0] = a.x; s[0].lo[1] = a.y; s[0].lo[2] = a.z; s[0].hi[0] = a.w; s[0].hi[1] = b.x; s[0].hi[2] = b.y;
    s[1].lo[0] = b.z; s[1].lo[1] = b.w; s[1].lo[2] = c.x; s[1].hi[0] = c.y; s[1].hi[1] = c.z; s[1].hi[2] = c.w;
    s[0].child = d.x; s[1].child = d.y; s[0].count = d.z; s[1].count = d.w;
}

__device__ __forceinline__ float slotArea(const Slot8& s)
{
    const float dx = s.hi[0] - s.lo[0], dy = s.hi[1] - s.lo[1], dz = s.hi[2] - s.lo[2];
    return dx * dy + dy * dz + dz * dx;
}

// counters: 0 = wide nodes allocated, 1 = triangles placed, 2 = entries of the next level's queue, 3 = depth
__global__ void k_collapse8(const BvhNode* __restrict__ nodes, const int2* __restrict__ qIn, int nIn, int level, int2* __restrict__ qOut,
                            uint32_t* __restrict__ counters, Bvh8Node* __restrict__ out, uint32_t* __restrict__ order8)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nIn) return;
    const int2 todo = qIn[e]; // x = two-child node, y = wide node
    // ---- gather up to eight children: open the largest inner child while there is room (bvh.cpp: collapseBvh8) ----
    Slot8 ch[8];
    int n = 0;
    Slot8 two[2];
    slotsOf(nodes, todo.x, two);
    for (int k = 0; k < 2; ++k) if (two[k].count >= 0) ch[n++] = two[k];
    while (n < 8) {
        int best = -1;
        float bestArea = -1.f;
        for (int k = 0; k < n; ++k)
            if (ch[k].count == 0) { const float a = slotArea(ch[k]); if (a > bestArea) { bestArea = a; best = k; } }
        if (best < 0) break;
        slotsOf(nodes, ch[best].child, two);
        int added = 0;
        Slot8 repl[2];
        for (int k = 0; k < 2; ++k) if (two[k].count >= 0) repl[added++] = two[k];
        if (added == 0) { ch[best] = ch[--n]; continue; }
        if (n - 1 + added > 8) break;
        ch[best] = repl[0];
        if (added == 2) ch[n++] = repl[1];
    }
    // ---- node frame: origin one step below the children's lower corner, one power-of-two step per axis, 253 steps of extent ----
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int k = 0; k < n; ++k)
        for (int a = 0; a < 3; ++a) { lo[a] = fminf(lo[a], ch[k].lo[a]); hi[a] = fmaxf(hi[a], ch[k].hi[a]); }
    Bvh8Node w;
    memset(&w, 0, sizeof(w));
    float scale[3];
    for (int a = 0; a < 3; ++a) {
        const float ext = fmaxf(hi[a] - lo[a], 1e-30f);
        int ex = int(ceil(log2(double(ext) / 253.0)));
        while (ldexp(253.0, ex) < double(hi[a]) - double(lo[a])) ++ex;
        ex = min(max(ex, -100), 100);
        w.e[a] = uint8_t(ex + 127);
        scale[a] = ldexpf(1.0f, ex);
        w.p[a] = lo[a] - scale[a];
        while (!((double(lo[a]) - double(w.p[a])) / double(scale[a]) >= 1.0)) w.p[a] = nextafterf(w.p[a], -FLT_MAX);
    }
    // ---- slot assignment: child k prefers the slot whose octant direction its centre is displaced to (greedy) ----
    float cen[3];
    for (int a = 0; a < 3; ++a) cen[a] = 0.5f * (lo[a] + hi[a]);
    float cost[8][8];
    for (int k = 0; k < n; ++k)
        for (int s8 = 0; s8 < 8; ++s8) {
            float c = 0.f;
            for (int a = 0; a < 3; ++a) c += (((s8 >> a) & 1) ? 1.f : -1.f) * (0.5f * (ch[k].lo[a] + ch[k].hi[a]) - cen[a]);
            cost[k][s8] = c;
        }
    int slotOf[8], childAt[8];
    for (int k = 0; k < 8; ++k) { slotOf[k] = -1; childAt[k] = -1; }
    for (int round = 0; round < n; ++round) {
        int bk = -1, bs = -1;
        float bc = -FLT_MAX;
        for (int k = 0; k < n; ++k) {
            if (slotOf[k] >= 0) continue;
            for (int s8 = 0; s8 < 8; ++s8)
                if (childAt[s8] < 0 && (bk < 0 || cost[k][s8] > bc)) { bc = cost[k][s8]; bk = k; bs = s8; }
        }
        slotOf[bk] = bs;
        childAt[bs] = bk;
    }
    // ---- emit ----
    int nInner = 0, nTri = 0;
    for (int k = 0; k < n; ++k) { if (ch[k].count == 0) ++nInner; else nTri += ch[k].count; }
    const uint32_t childBase = nInner ? atomicAdd(&counters[0], uint32_t(nInner)) : 0u;
    const uint32_t triBase = nTri ? atomicAdd(&counters[1], uint32_t(nTri)) : 0u;
    const uint32_t qBase = nInner ? atomicAdd(&counters[2], uint32_t(nInner)) : 0u;
    if (nInner == 0) atomicMax(&counters[3], uint32_t(level)); // (a node with inner children is not the deepest)
    w.childBase = childBase;
    w.triBase = triBase;
    uint32_t nextChild = 0, nextTri = 0;
    for (int s8 = 0; s8 < 8; ++s8) {
        for (int a = 0; a < 3; ++a) { w.qlo[a][s8] = 255; w.qhi[a][s8] = 0; } // empty: inverted box
        const int k = childAt[s8];
        if (k < 0) continue;
        const Slot8& c = ch[k];
        for (int a = 0; a < 3; ++a) {
            const double ql = floor((double(c.lo[a]) - double(w.p[a])) / double(scale[a])) - 1.0; // one step of margin (wf_trace8.cuh: byteMagic)
            const double qh = ceil((double(c.hi[a]) - double(w.p[a])) / double(scale[a])) + 1.0;
            w.qlo[a][s8] = uint8_t(fmin(fmax(ql, 0.0), 255.0));
            w.qhi[a][s8] = uint8_t(fmin(fmax(qh, 0.0), 255.0));
        }
        if (c.count == 0) {
            w.imask |= uint8_t(1u << s8);
            qOut[qBase + nextChild] = make_int2(c.child, int(childBase + nextChild));
            ++nextChild;
        }
        else {
            w.validTri |= ((1u << c.count) - 1u) << (4 * s8);
            for (int i = 0; i < c.count; ++i) order8[triBase + nextTri++] = uint32_t(c.child + i);
        }
    }
    uint4* dst = reinterpret_cast<uint4*>(out + todo.y);
    const uint4* srcw = reinterpret_cast<const uint4*>(&w);
#pragma unroll
    for (int k = 0; k < 5; ++k) dst[k] = srcw[k];
}

__global__ void k_gather4(const float4* __restrict__ src, const uint32_t* __restrict__ order, uint32_t n, float4* __restrict__ dst)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t s = size_t(order[i]);
#pragma unroll
    for (int k = 0; k < 4; ++k) dst[4 * size_t(i) + k] = src[4 * s + k];
}

// ---- the last merges on the host: full-sweep SAH over the clusters PLOC leaves standing ---------------------------------------
// PLOC runs until `topClusters` clusters remain (default 8: essentially the whole tree is agglomerated on the device) and the
// remaining ones — typically the handful of scene-sized boxes, e.g. the walls of a room around a detailed mesh — are split
// top-down with an exact sweep SAH (cost = area x weight on both sides, all three axes) in a few microseconds. Handing over
// EARLIER was measured and is worse: with 128..16384 clusters left to the sweep the 1 M-triangle scene traverses 4-5 % slower
// (3.25-3.36 node fetches per ray against 2.87) — the agglomerated top follows the geometry, the sweep only sees cluster boxes.
// Agglomerating to the root costs more rounds (137 instead of 55) and a deeper two-child tree (57-62 levels instead of 45), which
// is why k_trace's stack holds 128 entries.
struct TopCluster {
    float lo[3], hi[3], cost;
    uint32_t count;
    bool leaf;
    int id;
};
struct TopBuilder {
    std::vector<TopCluster> nodes; // [0, m) = the clusters, then the new inner nodes
    std::vector<int2> children;    // of the new inner nodes, as indices into `nodes`
    std::vector<int> order;        // working permutation of the cluster indices
    float ct = 1.f;
    int maxLeaf = 4;
    bool byClusters = false; // weight the two sides of a split by their number of clusters instead of their number of triangles
    static float harea(const float* lo, const float* hi)
    {
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return dx * dy + dy * dz + dz * dx;
    }
    // builds the subtree over order[first, first + count), returns its index in `nodes`
    int build(int first, int count)
    {
        if (count == 1) return order[size_t(first)];
        int bestAxis = -1, bestSplit = -1, bestBal = INT32_MAX;
        float bestCost = FLT_MAX;
        std::vector<float> rightArea(static_cast<size_t>(count));
        std::vector<uint32_t> rightCnt(static_cast<size_t>(count));
        for (int a = 0; a < 3; ++a) {
            std::sort(order.begin() + first, order.begin() + first + count, [&](int x, int y) {
                const float cx = nodes[size_t(x)].lo[a] + nodes[size_t(x)].hi[a], cy = nodes[size_t(y)].lo[a] + nodes[size_t(y)].hi[a];
                return cx < cy || (cx == cy && x < y);
            });
            float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
            uint32_t c = 0;
            for (int k = count - 1; k > 0; --k) {
                const TopCluster& t = nodes[size_t(order[size_t(first + k)])];
                for (int b = 0; b < 3; ++b) { lo[b] = std::min(lo[b], t.lo[b]); hi[b] = std::max(hi[b], t.hi[b]); }
                c += t.count;
                rightArea[size_t(k)] = harea(lo, hi);
                rightCnt[size_t(k)] = c;
            }
            for (int b = 0; b < 3; ++b) { lo[b] = FLT_MAX; hi[b] = -FLT_MAX; }
            c = 0;
            for (int k = 0; k + 1 < count; ++k) {
                const TopCluster& t = nodes[size_t(order[size_t(first + k)])];
                for (int b = 0; b < 3; ++b) { lo[b] = std::min(lo[b], t.lo[b]); hi[b] = std::max(hi[b], t.hi[b]); }
                c += t.count;
                const float wl = byClusters ? float(k + 1) : float(c), wr = byClusters ? float(count - k - 1) : float(rightCnt[size_t(k + 1)]);
                const float cost = harea(lo, hi) * wl + rightArea[size_t(k + 1)] * wr;
                // equal costs (coincident boxes): the most balanced split, or thousands of duplicates become a chain
                const int bal = std::abs(2 * (k + 1) - count);
                if (cost < bestCost || (cost == bestCost && bal < bestBal)) { bestCost = cost; bestAxis = a; bestSplit = k + 1; bestBal = bal; }
            }
        }
        if (bestAxis < 0) { bestAxis = 0; bestSplit = count / 2; } // (NaN boxes: halve the range)
        if (bestAxis != 2)
            std::sort(order.begin() + first, order.begin() + first + count, [&](int x, int y) {
                const int a = bestAxis;
                const float cx = nodes[size_t(x)].lo[a] + nodes[size_t(x)].hi[a], cy = nodes[size_t(y)].lo[a] + nodes[size_t(y)].hi[a];
                return cx < cy || (cx == cy && x < y);
            });
        const int l = build(first, bestSplit), r = build(first + bestSplit, count - bestSplit);
        TopCluster n{};
        const TopCluster &L = nodes[size_t(l)], &R = nodes[size_t(r)];
        for (int b = 0; b < 3; ++b) { n.lo[b] = std::min(L.lo[b], R.lo[b]); n.hi[b] = std::max(L.hi[b], R.hi[b]); }
        n.count = L.count + R.count;
        const float A = harea(n.lo, n.hi), costInner = ct * A + L.cost + R.cost, costLeaf = A * float(n.count);
        n.leaf = n.count <= uint32_t(maxLeaf) && costLeaf <= costInner; // the same leaf rule as k_ploc_emit
        n.cost = n.leaf ? costLeaf : costInner;
        n.id = -1;
        nodes.push_back(n);
        children.push_back(make_int2(l, r));
        return int(nodes.size()) - 1; // children are always created before their parent: the root is the last node
    }
};

// One cudaMalloc / cudaFree per build: on this driver every large cudaMalloc + cudaFree pair costs milliseconds (twenty separate
// scratch arrays were 120 ms of a 125 ms build); requests are registered first, then carved out of a single allocation.
struct Arena {
    struct Req { void** p; size_t bytes; };
    std::vector<Req> reqs;
    void* base = nullptr;
    ~Arena() { if (base) cudaFree(base); }
    template <typename T> void want(T** p, size_t count) { reqs.push_back({reinterpret_cast<void**>(p), (sizeof(T) * (count ? count : 1) + 255) & ~size_t(255)}); }
    cudaError_t commit()
    {
        size_t total = 0;
        for (const Req& r : reqs) total += r.bytes;
        const cudaError_t e = cudaMalloc(&base, total);
        if (e != cudaSuccess) return e;
        size_t off = 0;
        for (const Req& r : reqs) { *r.p = static_cast<char*>(base) + off; off += r.bytes; }
        return cudaSuccess;
    }
};

#define GB(call)                               \
    do {                                       \
        cudaError_t e__ = (call);              \
        if (e__ != cudaSuccess) return e__;    \
    } while (0)

} // namespace

void launchIngest(const xrtg_triangle* dRaw, const MeshRange* dRanges, int nRanges, int nTris, float4* trisId, float4* ftrisId, float4* prims,
                  cudaStream_t st)
{
    if (nTris <= 0) return;
    k_ingest<<<(nTris + 127) / 128, 128, 0, st>>>(dRaw, dRanges, nRanges, nTris, trisId, ftrisId, prims);
}

// Host-only check of the top-level builder: splits `n` clusters (boxes lo/hi, triangle counts) and verifies that every cluster is
// referenced exactly once, that every node's box is the union of its children's, that counts add up and that the depth stays
// logarithmic on coincident boxes. Returns the depth of the tree, or -1 on an inconsistency.
int topSahSelftest(const float* lo3, const float* hi3, const uint32_t* counts, int n, int byClusters)
{
    if (n < 1) return -1;
    TopBuilder T;
    T.byClusters = byClusters != 0;
    T.nodes.resize(static_cast<size_t>(n));
    T.order.resize(static_cast<size_t>(n));
    for (int k = 0; k < n; ++k) {
        TopCluster& c = T.nodes[size_t(k)];
        for (int a = 0; a < 3; ++a) { c.lo[a] = lo3[3 * k + a]; c.hi[a] = hi3[3 * k + a]; }
        c.count = counts ? counts[k] : 1u;
        c.cost = TopBuilder::harea(c.lo, c.hi) * float(c.count);
        c.leaf = c.count <= 4;
        c.id = k;
        T.order[size_t(k)] = k;
    }
    T.nodes.reserve(2 * size_t(n));
    const int root = T.build(0, n);
    if (int(T.children.size()) != n - 1 || root != int(T.nodes.size()) - 1) return -1;
    std::vector<int> seen(static_cast<size_t>(n), 0);
    int maxDepth = 0;
    bool ok = true;
    struct Item { int node, depth; };
    std::vector<Item> stack{{root, 1}};
    while (!stack.empty()) {
        const Item it = stack.back();
        stack.pop_back();
        maxDepth = std::max(maxDepth, it.depth);
        if (it.node < n) { seen[size_t(it.node)]++; continue; }
        const int2 ch = T.children[size_t(it.node - n)];
        const TopCluster &N = T.nodes[size_t(it.node)], &L = T.nodes[size_t(ch.x)], &R = T.nodes[size_t(ch.y)];
        if (ch.x >= it.node || ch.y >= it.node) ok = false; // children are created before their parent
        if (N.count != L.count + R.count) ok = false;
        for (int a = 0; a < 3; ++a)
            if (N.lo[a] != std::min(L.lo[a], R.lo[a]) || N.hi[a] != std::max(L.hi[a], R.hi[a])) ok = false;
        stack.push_back({ch.x, it.depth + 1});
        stack.push_back({ch.y, it.depth + 1});
    }
    for (int k = 0; k < n; ++k)
        if (seen[size_t(k)] != 1) ok = false;
    return ok ? maxDepth : -1;
}

void launchScatterPrims(const float4* dRecs, const int* dIds, int count, float4* prims, cudaStream_t st)
{
    if (count > 0) k_scatter_prims<<<(count + 127) / 128, 128, 0, st>>>(dRecs, dIds, count, prims);
}

cudaError_t buildPlocDevice(const float4* dTrisId, uint32_t n, float4* dTrisLeafOrder, BvhNode* dNodes, cudaStream_t st, GpuBuildInfo* info,
                            const float4* dFastId, float4* dFastLeafOrder, const PlocParams& prm)
{
    if (n < 2) return cudaErrorInvalidValue;
    const auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!prm.verbose) return;
        cudaStreamSynchronize(st);
        std::fprintf(stderr, "ploc: %-34s at %8.2f ms\n", what, std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count());
    };
    Arena S;
    float4 *triLo, *triHi, *lo, *hi, *dTop;
    uint32_t *bounds, *vals, *vals2;
    unsigned long long *keys = nullptr, *keys2 = nullptr, *packed = nullptr, *incl = nullptr;
    int2 *child, *dIdParent;
    int *parent, *cidA, *cidB, *nn, *leafPos, *first, *stats;
    PlocState* state;
    unsigned char* tmp;
    const size_t nNodes = 2 * size_t(n) - 1;
    const int topClusters = prm.topClusters < 1 ? 1 : prm.topClusters;
    const size_t topCap = size_t(topClusters) < size_t(n) ? size_t(topClusters) : size_t(n);
    size_t sortBytes = 0, scanBytes = 0;
    GB(cub::DeviceRadixSort::SortPairs(nullptr, sortBytes, keys, keys2, (uint32_t*)nullptr, (uint32_t*)nullptr, int(n), 0, 63, st));
    GB(cub::DeviceScan::InclusiveSum(nullptr, scanBytes, packed, incl, int(n), st));
    S.want(&triLo, n); S.want(&triHi, n); S.want(&lo, nNodes); S.want(&hi, nNodes);
    S.want(&bounds, 8); S.want(&vals, n); S.want(&vals2, n); S.want(&keys, n); S.want(&keys2, n);
    S.want(&packed, n); S.want(&incl, n); S.want(&child, nNodes); S.want(&parent, nNodes);
    S.want(&cidA, n); S.want(&cidB, n); S.want(&nn, n); S.want(&leafPos, n); S.want(&first, nNodes); S.want(&stats, 4);
    S.want(&tmp, sortBytes > scanBytes ? sortBytes : scanBytes);
    S.want(&dTop, 2 * topCap); S.want(&dIdParent, 2 * topCap); S.want(&state, 2);
    GB(S.commit());
    const uint32_t initB[8] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u, 0u, 0u};
    GB(cudaMemcpyAsync(bounds, initB, sizeof(initB), cudaMemcpyHostToDevice, st));
    GB(cudaMemsetAsync(stats, 0, sizeof(int) * 4, st));
    const int grid = int((n + kB - 1) / kB);
    lap("scratch allocation");
    k_boxes<<<grid, kB, 0, st>>>(dTrisId, n, triLo, triHi, bounds);
    k_morton<<<grid, kB, 0, st>>>(triLo, triHi, n, bounds, keys, vals);
    GB(cub::DeviceRadixSort::SortPairs(tmp, sortBytes, keys, keys2, vals, vals2, int(n), 0, 63, st));
    k_ploc_init<<<grid, kB, 0, st>>>(vals2, triLo, triHi, int(n), lo, hi, child, cidA);
    lap("boxes, Morton codes, sort");
    // ---- merge rounds ----
    const int R = prm.radius < 1 ? 1 : (prm.radius > 64 ? 64 : prm.radius);
    const int maxLeaf = prm.maxLeaf < 1 ? 1 : (prm.maxLeaf > 4 ? 4 : prm.maxLeaf);
    const size_t nnSmem = sizeof(float) * 6 * size_t(kB + 2 * R);
    // Rounds are enqueued in batches of kRoundsPerSync with ONE read-back of the loop state per batch (a read-back per round made the
    // build hostage to host scheduling: 137 synchronisations, 20 ms on a quiet box and 300 ms on a busy one). Grids and the scan are
    // sized for the cluster count known at the last read-back; rounds past convergence only carry the state forward.
    constexpr int kRoundsPerSync = 8;
    int m = int(n), nodeBase = int(n), iterations = 0;
    int *cid = cidA, *cidNext = cidB;
    PlocState* stHost = nullptr;
    GB(cudaMallocHost(&stHost, sizeof(PlocState)));
    struct Unpin { PlocState* p; ~Unpin() { cudaFreeHost(p); } } unpin{stHost};
    *stHost = PlocState{m, nodeBase, 0, 0};
    GB(cudaMemcpyAsync(state, stHost, sizeof(PlocState), cudaMemcpyHostToDevice, st));
    GB(cudaStreamSynchronize(st));
    int round = 0;
    while (m > topClusters) {
        const int mUpper = m, g = (mUpper + kB - 1) / kB;
        for (int k = 0; k < kRoundsPerSync; ++k, ++round) {
            const PlocState* in = state + (round & 1);
            PlocState* out = state + ((round + 1) & 1);
            k_ploc_nn<<<g, kB, nnSmem, st>>>(cid, in, topClusters, lo, hi, R, nn);
            k_ploc_flag<<<g, kB, 0, st>>>(nn, in, topClusters, mUpper, packed);
            GB(cub::DeviceScan::InclusiveSum(tmp, scanBytes, packed, incl, mUpper, st));
            k_ploc_emit<<<g, kB, 0, st>>>(cid, nn, incl, in, out, topClusters, mUpper, prm.traversalCost, maxLeaf, lo, hi, child, parent, cidNext);
            int* t = cid; cid = cidNext; cidNext = t;
        }
        GB(cudaMemcpyAsync(stHost, state + (round & 1), sizeof(PlocState), cudaMemcpyDeviceToHost, st));
        GB(cudaStreamSynchronize(st));
        if (stHost->m >= m && stHost->m > topClusters) return cudaErrorUnknown; // (cannot happen: the globally closest pair is always mutual)
        if (prm.verbose) std::fprintf(stderr, "ploc: after %d rounds: %d clusters\n", stHost->rounds, stHost->m);
        m = stHost->m;
        nodeBase = stHost->nodeBase;
        iterations = stHost->rounds;
        if (iterations > 4096) return cudaErrorNotSupported; // degenerate input (one merge per round): leave it to the host builder
    }
    lap("clustering rounds");
    if (m > 1) {
        // ---- the top of the tree: sweep SAH over the m surviving clusters on the host (TopBuilder) ----
        k_gather_clusters<<<(m + kB - 1) / kB, kB, 0, st>>>(cid, m, lo, hi, dTop);
        std::vector<float4> top(2 * size_t(m));
        std::vector<int> ids(static_cast<size_t>(m));
        GB(cudaMemcpyAsync(top.data(), dTop, sizeof(float4) * top.size(), cudaMemcpyDeviceToHost, st));
        GB(cudaMemcpyAsync(ids.data(), cid, sizeof(int) * size_t(m), cudaMemcpyDeviceToHost, st));
        GB(cudaStreamSynchronize(st));
        TopBuilder T;
        T.ct = prm.traversalCost;
        T.maxLeaf = maxLeaf;
        T.byClusters = prm.topByClusters;
        T.nodes.resize(size_t(m));
        T.order.resize(size_t(m));
        for (int k = 0; k < m; ++k) {
            TopCluster& c = T.nodes[size_t(k)];
            const float4 l = top[2 * size_t(k)], h = top[2 * size_t(k) + 1];
            c.lo[0] = l.x; c.lo[1] = l.y; c.lo[2] = l.z; c.cost = l.w;
            c.hi[0] = h.x; c.hi[1] = h.y; c.hi[2] = h.z;
            uint32_t w;
            std::memcpy(&w, &h.w, 4);
            c.count = w & ~kLeafFlag;
            c.leaf = (w & kLeafFlag) != 0;
            c.id = ids[size_t(k)];
            T.order[size_t(k)] = k;
        }
        T.nodes.reserve(2 * size_t(m));
        T.build(0, m);
        const int nNew = m - 1; // new node k (creation order) gets id nodeBase + k: the root, created last, is 2n - 2
        if (int(T.children.size()) != nNew) return cudaErrorUnknown;
        for (int k = 0; k < nNew; ++k) T.nodes[size_t(m + k)].id = nodeBase + k;
        std::vector<float4> newLo(static_cast<size_t>(nNew)), newHi(static_cast<size_t>(nNew));
        std::vector<int2> newChild(static_cast<size_t>(nNew)), idParent;
        for (int k = 0; k < nNew; ++k) {
            const TopCluster& c = T.nodes[size_t(m + k)];
            const uint32_t w = c.count | (c.leaf ? kLeafFlag : 0u);
            float wf;
            std::memcpy(&wf, &w, 4);
            newLo[size_t(k)] = make_float4(c.lo[0], c.lo[1], c.lo[2], c.cost);
            newHi[size_t(k)] = make_float4(c.hi[0], c.hi[1], c.hi[2], wf);
            const int2 ch = T.children[size_t(k)];
            newChild[size_t(k)] = make_int2(T.nodes[size_t(ch.x)].id, T.nodes[size_t(ch.y)].id);
            idParent.push_back(make_int2(newChild[size_t(k)].x, nodeBase + k));
            idParent.push_back(make_int2(newChild[size_t(k)].y, nodeBase + k));
        }
        GB(cudaMemcpyAsync(lo + nodeBase, newLo.data(), sizeof(float4) * size_t(nNew), cudaMemcpyHostToDevice, st));
        GB(cudaMemcpyAsync(hi + nodeBase, newHi.data(), sizeof(float4) * size_t(nNew), cudaMemcpyHostToDevice, st));
        GB(cudaMemcpyAsync(child + nodeBase, newChild.data(), sizeof(int2) * size_t(nNew), cudaMemcpyHostToDevice, st));
        GB(cudaMemcpyAsync(dIdParent, idParent.data(), sizeof(int2) * idParent.size(), cudaMemcpyHostToDevice, st));
        k_scatter_parents<<<(int(idParent.size()) + kB - 1) / kB, kB, 0, st>>>(dIdParent, int(idParent.size()), parent);
        GB(cudaStreamSynchronize(st)); // (the staging vectors go out of scope)
        nodeBase += nNew;
    }
    lap("top levels (host sweep SAH)");
    const int root = int(nNodes) - 1;
    if (nodeBase != int(nNodes)) return cudaErrorUnknown;
    // ---- leaf order, node records ----
    uint32_t hb[8];
    GB(cudaMemcpyAsync(hb, bounds, sizeof(hb), cudaMemcpyDeviceToHost, st));
    k_leafpos<<<grid, kB, 0, st>>>(int(n), root, parent, child, hi, leafPos, first, stats);
    int hstats[4];
    float4 rootLo;
    GB(cudaMemcpyAsync(hstats, stats, sizeof(hstats), cudaMemcpyDeviceToHost, st));
    GB(cudaMemcpyAsync(&rootLo, lo + root, sizeof(float4), cudaMemcpyDeviceToHost, st));
    float4 rootHi;
    GB(cudaMemcpyAsync(&rootHi, hi + root, sizeof(float4), cudaMemcpyDeviceToHost, st));
    GB(cudaStreamSynchronize(st));
    float mag = 0.f;
    for (int a = 0; a < 6; ++a) mag = fmaxf(mag, fabsf(fdec(hb[a])));
    const float pad = fmaxf(mag * (1.f / 32768.f), 1e-30f); // 2^-15 of the largest absolute scene coordinate, as bvh.cpp
    k_emit_nodes<<<grid, kB, 0, st>>>(int(n), lo, hi, child, first, pad, dNodes);
    k_scatter_tris<<<grid, kB, 0, st>>>(dTrisId, vals2, leafPos, n, dTrisLeafOrder, 3);
    if (dFastId && dFastLeafOrder) k_scatter_tris<<<grid, kB, 0, st>>>(dFastId, vals2, leafPos, n, dFastLeafOrder, 4);
    GB(cudaStreamSynchronize(st));
    GB(cudaGetLastError());
    lap("leaf order, node records, gathers");
    if (info) {
        info->depth = hstats[0];
        info->nNodes = hstats[1] - 1;
        info->nAllocated = int(n) - 1;
        info->iterations = iterations;
        info->pad = pad;
        const float rootArea = (rootHi.x - rootLo.x) * (rootHi.y - rootLo.y) + (rootHi.y - rootLo.y) * (rootHi.z - rootLo.z) + (rootHi.z - rootLo.z) * (rootHi.x - rootLo.x);
        info->sahCost = rootLo.w / fmaxf(rootArea, 1e-30f);
        for (int a = 0; a < 3; ++a) { info->lo[a] = fdec(hb[a]); info->hi[a] = fdec(hb[3 + a]); }
    }
    if (prm.verbose) std::fprintf(stderr, "ploc: %d rounds, depth %d, %d leaves, SAH %.2f\n", iterations, hstats[0], hstats[1], info ? info->sahCost : 0.f);
    if (hstats[0] > kMaxTreeDepth) return cudaErrorNotSupported; // deeper than k_trace's stack (wavefront.cuh: kStackLocalDeep): host builder
    return cudaSuccess;
}

cudaError_t collapseBvh8Device(const BvhNode* dNodes, uint32_t nTris, const float4* dFastLeafOrder, Bvh8Node** dNodes8, uint32_t* nNodes8,
                               float4* dFtris8, int* depth8, cudaStream_t st)
{
    *dNodes8 = nullptr;
    *nNodes8 = 0;
    if (nTris < 2) return cudaErrorInvalidValue;
    Arena S;
    Bvh8Node* wide;
    int2 *qA, *qB;
    uint32_t *order8, *counters;
    // every wide node absorbs at least one two-child node, and there are at most nTris - 1 of those
    S.want(&wide, nTris); S.want(&qA, nTris); S.want(&qB, nTris); S.want(&order8, nTris); S.want(&counters, 4);
    GB(S.commit());
    const uint32_t init[4] = {1u, 0u, 0u, 0u};
    const int2 rootEntry = make_int2(0, 0);
    GB(cudaMemcpyAsync(counters, init, sizeof(init), cudaMemcpyHostToDevice, st));
    GB(cudaMemcpyAsync(qA, &rootEntry, sizeof(rootEntry), cudaMemcpyHostToDevice, st));
    uint32_t* countersHost = nullptr;
    GB(cudaMallocHost(&countersHost, sizeof(uint32_t) * 4));
    struct Unpin { uint32_t* p; ~Unpin() { cudaFreeHost(p); } } unpin{countersHost};
    int nIn = 1, level = 1;
    while (nIn > 0) {
        k_collapse8<<<(nIn + 63) / 64, 64, 0, st>>>(dNodes, qA, nIn, level, qB, counters, wide, order8);
        GB(cudaMemcpyAsync(countersHost, counters, sizeof(uint32_t) * 4, cudaMemcpyDeviceToHost, st));
        GB(cudaMemsetAsync(counters + 2, 0, sizeof(uint32_t), st));
        GB(cudaStreamSynchronize(st));
        nIn = int(countersHost[2]);
        if (countersHost[0] > nTris) return cudaErrorUnknown;
        int2* t = qA; qA = qB; qB = t;
        if (++level > 128) return cudaErrorNotSupported;
    }
    GB(cudaGetLastError());
    const uint32_t nWide = countersHost[0], nPlaced = countersHost[1];
    if (nPlaced != nTris) return cudaErrorUnknown; // every triangle sits in exactly one leaf child
    GB(cudaMalloc(reinterpret_cast<void**>(dNodes8), sizeof(Bvh8Node) * size_t(nWide)));
    GB(cudaMemcpyAsync(*dNodes8, wide, sizeof(Bvh8Node) * size_t(nWide), cudaMemcpyDeviceToDevice, st));
    k_gather4<<<(nTris + kB - 1) / kB, kB, 0, st>>>(dFastLeafOrder, order8, nTris, dFtris8);
    GB(cudaStreamSynchronize(st));
    GB(cudaGetLastError());
    *nNodes8 = nWide;
    *depth8 = int(countersHost[3]);
    return cudaSuccess;
}

} // namespace xrt
