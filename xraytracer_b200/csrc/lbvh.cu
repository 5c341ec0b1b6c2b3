// lbvh.cu — GPU-side BVH build: a linear BVH after Karras, "Maximizing Parallelism in the
// Construction of BVHs, Octrees, and k-d Trees" (HPG 2012), emitted directly in the traversal kernels' node format
// (bvh.h: BvhNode, both children's padded boxes in the parent, one triangle per leaf).
//
//   triangle AABBs + scene bounds  ->  30-bit Morton code of each centroid  ->  radix sort (cub::DeviceRadixSort, a
//   sorting primitive outside the render path)  ->  radix tree (one thread per internal node, index tie-break so equal
//   codes still split)  ->  bottom-up box fit with per-node arrival counters  ->  BvhNode records + leaf-ordered triangles
//
// Like the host SAH builder this is NEW functionality relative to the reference (Scene::build() is an empty hook,
// scene.h:22-24); any valid BVH must return exactly what brute force returns, which tests/test_gpu_parity.py checks for
// this builder with the same BVH == brute force == oracle comparisons.
#include <cfloat>
#include <cstdint>
#include <cstring>
#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>
#include "bvh.h"
#include "lbvh.h"

namespace xrt {
namespace {

constexpr int kB = 256;

__device__ __forceinline__ float3 f3min(float3 a, float3 b) { return make_float3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
__device__ __forceinline__ float3 f3max(float3 a, float3 b) { return make_float3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }

// order-preserving float <-> uint so that atomicMin/atomicMax work on floats of either sign
__device__ __forceinline__ uint32_t fenc(float f) { const uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float fdec(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

// per-triangle box (lo, hi as float4 pairs) and the scene bounds (6 encoded uints: lo xyz, hi xyz)
__global__ void k_boxes(const float4* __restrict__ trisId, uint32_t n, float4* __restrict__ lo, float4* __restrict__ hi, uint32_t* bounds)
{
    __shared__ uint32_t s[6];
    if (threadIdx.x < 6) s[threadIdx.x] = threadIdx.x < 3 ? 0xffffffffu : 0u;
    __syncthreads();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float4 a = trisId[3 * size_t(i)], e1 = trisId[3 * size_t(i) + 1], e2 = trisId[3 * size_t(i) + 2];
        const float3 v0 = make_float3(a.x, a.y, a.z);
        // v1 = v0 + e1 may differ from the original vertex by an ulp; the conservative padding (2^-15 of the scene
        // magnitude) is five orders of magnitude larger
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

__device__ __forceinline__ uint32_t expandBits(uint32_t v)
{
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__global__ void k_morton(const float4* __restrict__ lo, const float4* __restrict__ hi, uint32_t n, const uint32_t* __restrict__ bounds,
                         uint32_t* __restrict__ keys, uint32_t* __restrict__ vals)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float3 smin = make_float3(fdec(bounds[0]), fdec(bounds[1]), fdec(bounds[2]));
    const float3 smax = make_float3(fdec(bounds[3]), fdec(bounds[4]), fdec(bounds[5]));
    const float4 l = lo[i], h = hi[i];
    const float cx = 0.5f * (l.x + h.x), cy = 0.5f * (l.y + h.y), cz = 0.5f * (l.z + h.z);
    const float ex = fmaxf(smax.x - smin.x, 1e-30f), ey = fmaxf(smax.y - smin.y, 1e-30f), ez = fmaxf(smax.z - smin.z, 1e-30f);
    const uint32_t x = min(1023u, uint32_t(fmaxf((cx - smin.x) / ex, 0.f) * 1024.f));
    const uint32_t y = min(1023u, uint32_t(fmaxf((cy - smin.y) / ey, 0.f) * 1024.f));
    const uint32_t z = min(1023u, uint32_t(fmaxf((cz - smin.z) / ez, 0.f) * 1024.f));
    keys[i] = (expandBits(x) << 2) | (expandBits(y) << 1) | expandBits(z);
    vals[i] = i;
}

// common-prefix length of sorted keys i and j; equal keys fall back to the index so the tree stays balanced
__device__ __forceinline__ int delta(const uint32_t* __restrict__ keys, int n, int i, int j)
{
    if (j < 0 || j >= n) return -1;
    const uint32_t a = keys[i], b = keys[j];
    if (a == b) return 32 + __clz(uint32_t(i) ^ uint32_t(j));
    return __clz(a ^ b);
}

// Karras' radix tree: internal node i covers a range of sorted leaves; child index >= 0 = internal, < 0 = ~leaf
__global__ void k_hierarchy(const uint32_t* __restrict__ keys, int n, int2* __restrict__ children, int* __restrict__ parentInner,
                            int* __restrict__ parentLeaf)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) / 2;; t = (t + 1) / 2) {
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const int left = (lo == gamma) ? ~gamma : gamma;
    const int right = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    children[i] = make_int2(left, right);
    if (left >= 0) parentInner[left] = i; else parentLeaf[~left] = i;
    if (right >= 0) parentInner[right] = i; else parentLeaf[~right] = i;
}

// bottom-up: the second thread to arrive at a node merges its children's boxes
__global__ void k_fit(int n, const uint32_t* __restrict__ order, const float4* __restrict__ triLo, const float4* __restrict__ triHi,
                      const int2* __restrict__ children, const int* __restrict__ parentInner, const int* __restrict__ parentLeaf,
                      float4* __restrict__ nodeLo, float4* __restrict__ nodeHi, int* __restrict__ arrive)
{
    const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= n) return;
    int node = parentLeaf[leaf];
    while (true) {
        if (atomicAdd(&arrive[node], 1) == 0) return; // first arrival: the sibling will finish this node
        __threadfence();
        const int2 c = children[node];
        float4 l0, h0, l1, h1;
        // children boxes were written by other SMs: read them through L2 (__ldcg), never from a possibly stale L1 line
        if (c.x >= 0) { l0 = __ldcg(nodeLo + c.x); h0 = __ldcg(nodeHi + c.x); } else { l0 = triLo[order[~c.x]]; h0 = triHi[order[~c.x]]; }
        if (c.y >= 0) { l1 = __ldcg(nodeLo + c.y); h1 = __ldcg(nodeHi + c.y); } else { l1 = triLo[order[~c.y]]; h1 = triHi[order[~c.y]]; }
        nodeLo[node] = make_float4(fminf(l0.x, l1.x), fminf(l0.y, l1.y), fminf(l0.z, l1.z), 0.f);
        nodeHi[node] = make_float4(fmaxf(h0.x, h1.x), fmaxf(h0.y, h1.y), fmaxf(h0.z, h1.z), 0.f);
        __threadfence();
        if (node == 0) return;
        node = parentInner[node];
    }
}

// exact tree depth: every leaf counts its ancestors (the traversal stacks hold 64 entries)
__global__ void k_depth(int n, const int* __restrict__ parentInner, const int* __restrict__ parentLeaf, int* __restrict__ depthOut)
{
    const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= n) return;
    int d = 1, node = parentLeaf[leaf];
    while (node != 0) { node = parentInner[node]; ++d; }
    atomicMax(depthOut, d + 1);
}

// BvhNode records (children's padded boxes in the parent) and the leaf-ordered triangle array
__global__ void k_emit(int n, const uint32_t* __restrict__ order, const float4* __restrict__ triLo, const float4* __restrict__ triHi,
                       const int2* __restrict__ children, const float4* __restrict__ nodeLo, const float4* __restrict__ nodeHi, float pad,
                       BvhNode* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int2 c = children[i];
    BvhNode nd;
    const int cs[2] = {c.x, c.y};
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        float4 l, h;
        if (cs[s] >= 0) { l = nodeLo[cs[s]]; h = nodeHi[cs[s]]; } else { l = triLo[order[~cs[s]]]; h = triHi[order[~cs[s]]]; }
        float* lo = s ? nd.lo1 : nd.lo0;
        float* hi = s ? nd.hi1 : nd.hi0;
        lo[0] = l.x - pad; lo[1] = l.y - pad; lo[2] = l.z - pad;
        hi[0] = h.x + pad; hi[1] = h.y + pad; hi[2] = h.z + pad;
        (s ? nd.child1 : nd.child0) = cs[s] >= 0 ? cs[s] : ~cs[s]; // leaf: index into the leaf-ordered triangle array
        (s ? nd.count1 : nd.count0) = cs[s] >= 0 ? 0 : 1;
    }
    out[i] = nd;
}

__global__ void k_gather_tris(const float4* __restrict__ trisId, const uint32_t* __restrict__ order, uint32_t n, float4* __restrict__ tris,
                              int f4PerTri)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t src = order[i];
    for (int k = 0; k < f4PerTri; ++k) tris[size_t(f4PerTri) * i + k] = trisId[size_t(f4PerTri) * src + k];
}

#define LB(call)                                       \
    do {                                               \
        cudaError_t e__ = (call);                      \
        if (e__ != cudaSuccess) { cleanup(); return e__; } \
    } while (0)

} // namespace

cudaError_t buildLbvhDevice(const float4* dTrisId, uint32_t n, float4* dTrisLeafOrder, BvhNode* dNodes, cudaStream_t st, LbvhInfo* info,
                            const float4* dFastId, float4* dFastLeafOrder)
{
    // n >= 2 (the caller handles 0 and 1 triangles with the host builder)
    void* bufs[16] = {};
    int nb = 0;
    auto cleanup = [&]() { for (int i = 0; i < nb; ++i) cudaFree(bufs[i]); };
    auto alloc = [&](void** p, size_t bytes) { cudaError_t e = cudaMalloc(p, bytes); if (e == cudaSuccess) bufs[nb++] = *p; return e; };
    float4 *triLo, *triHi, *nodeLo, *nodeHi;
    uint32_t *bounds, *keys, *vals, *keys2, *vals2;
    int2* children;
    int *parentInner, *parentLeaf, *arrive, *depth;
    LB(alloc((void**)&triLo, sizeof(float4) * n)); LB(alloc((void**)&triHi, sizeof(float4) * n));
    LB(alloc((void**)&nodeLo, sizeof(float4) * n)); LB(alloc((void**)&nodeHi, sizeof(float4) * n));
    LB(alloc((void**)&bounds, sizeof(uint32_t) * 8));
    LB(alloc((void**)&keys, sizeof(uint32_t) * n)); LB(alloc((void**)&vals, sizeof(uint32_t) * n));
    LB(alloc((void**)&keys2, sizeof(uint32_t) * n)); LB(alloc((void**)&vals2, sizeof(uint32_t) * n));
    LB(alloc((void**)&children, sizeof(int2) * n));
    LB(alloc((void**)&parentInner, sizeof(int) * n)); LB(alloc((void**)&parentLeaf, sizeof(int) * n));
    LB(alloc((void**)&arrive, sizeof(int) * n)); LB(alloc((void**)&depth, sizeof(int)));
    const uint32_t initB[8] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u, 0u, 0u};
    LB(cudaMemcpyAsync(bounds, initB, sizeof(initB), cudaMemcpyHostToDevice, st));
    LB(cudaMemsetAsync(arrive, 0, sizeof(int) * n, st));
    LB(cudaMemsetAsync(depth, 0, sizeof(int), st));
    const int grid = int((n + kB - 1) / kB);
    k_boxes<<<grid, kB, 0, st>>>(dTrisId, n, triLo, triHi, bounds);
    k_morton<<<grid, kB, 0, st>>>(triLo, triHi, n, bounds, keys, vals);
    size_t tmpBytes = 0;
    LB(cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, keys, keys2, vals, vals2, int(n), 0, 30, st));
    void* tmp = nullptr;
    LB(alloc(&tmp, tmpBytes));
    LB(cub::DeviceRadixSort::SortPairs(tmp, tmpBytes, keys, keys2, vals, vals2, int(n), 0, 30, st));
    k_hierarchy<<<grid, kB, 0, st>>>(keys2, int(n), children, parentInner, parentLeaf);
    k_fit<<<grid, kB, 0, st>>>(int(n), vals2, triLo, triHi, children, parentInner, parentLeaf, nodeLo, nodeHi, arrive);
    k_depth<<<grid, kB, 0, st>>>(int(n), parentInner, parentLeaf, depth);
    // padding = 2^-15 of the largest absolute scene coordinate, as bvh.cpp
    uint32_t hb[8];
    LB(cudaMemcpyAsync(hb, bounds, sizeof(hb), cudaMemcpyDeviceToHost, st));
    LB(cudaStreamSynchronize(st));
    auto dec = [](uint32_t u) { uint32_t v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u; float f; std::memcpy(&f, &v, 4); return f; };
    float mag = 0.f;
    for (int a = 0; a < 6; ++a) mag = fmaxf(mag, fabsf(dec(hb[a])));
    const float pad = fmaxf(mag * (1.f / 32768.f), 1e-30f);
    k_emit<<<grid, kB, 0, st>>>(int(n), vals2, triLo, triHi, children, nodeLo, nodeHi, pad, dNodes);
    k_gather_tris<<<grid, kB, 0, st>>>(dTrisId, vals2, n, dTrisLeafOrder, 3);
    if (dFastId && dFastLeafOrder) k_gather_tris<<<grid, kB, 0, st>>>(dFastId, vals2, n, dFastLeafOrder, 4);
    int hdepth = 0;
    LB(cudaMemcpyAsync(&hdepth, depth, sizeof(int), cudaMemcpyDeviceToHost, st));
    LB(cudaStreamSynchronize(st));
    LB(cudaGetLastError());
    if (info) { info->depth = hdepth; info->pad = pad; info->nNodes = int(n) - 1; }
    cleanup();
    return cudaSuccess;
}

} // namespace xrt
