// gpu_build.h — scene ingest and BVH construction entirely on the device (gpu_build.cu): per-triangle records from the raw
// xrtg_triangle array, a PLOC bounding-volume hierarchy (parallel locally-ordered clustering, Meister & Bittner 2018) with
// SAH-decided leaves of up to four triangles, and the collapse of any two-child tree into the eight-child quantised nodes the
// throughput traversal kernel walks (bvh.h: Bvh8Node). SURVEY §8(f) rank 2; the hook it fills is the reference's empty
// Scene::build() (scene.h:22-24), the ingest it replaces is the per-triangle loop of scene.cpp:46-154 / primitive.cpp:105,142-143.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <xrtgpu.h>
#include "bvh.h"

namespace xrt {

// One mesh object of the scene description: triangles [srcFirst, srcFirst + count) of the raw array become the mesh triangles
// [triStart, triStart + count) with global primitive ids id0 + k (object order = the reference's iteration order).
struct MeshRange {
    int32_t srcFirst, triStart, id0, emitter;
    float albedo[3];
    uint32_t meta;
};

// dRaw: the xrtg_triangle array on the device; ranges (device, sorted by triStart), nRanges >= 1; nTris = mesh triangles in total.
// Writes trisId (3 float4 per triangle: v0|id, e1|emitter, e2|0 with the reference's fp32 subtraction), ftrisId (4 float4: the
// plane-equation record, double arithmetic rounded once, bit-identical to small_scene.cpp: makePlaneRecord) and the shading
// records prims[4 * id ..] (n0|ng.x n1|ng.y n2|ng.z albedo|meta; ng = normalize(e1 x e2) in the reference's operation order).
void launchIngest(const xrtg_triangle* dRaw, const MeshRange* dRanges, int nRanges, int nTris, float4* trisId, float4* ftrisId, float4* prims,
                  cudaStream_t st);

// prims[4 * ids[k] .. +4) = recs[4 * k .. +4): shading records of the analytic spheres / boxes (built on the host, there are few).
void launchScatterPrims(const float4* dRecs, const int* dIds, int count, float4* prims, cudaStream_t st);

// Host-only structural check of the top-level sweep-SAH builder (no CUDA device needed); returns the depth of the tree or -1.
int topSahSelftest(const float* lo3, const float* hi3, const uint32_t* counts, int n, int byClusters);

struct PlocParams {
    int radius = 16;          // neighbour-search window on each side of a cluster in Morton order
    float traversalCost = 1.f; // SAH: cost of one node step relative to one triangle test (leaf decision)
    int maxLeaf = 4;          // triangles per leaf, 1..4 (the traversal kernels hold count - 1 in two bits)
    int topClusters = 8;      // clustering stops here; the last few merges are an exact sweep SAH on the host (gpu_build.cu: TopBuilder).
                              // Measured on the 1 M-triangle scene: handing over at 1..16 clusters traverses 4-5 % faster than at 128..16384
    bool topByClusters = true;  // sweep SAH weights: clusters per side instead of triangles per side
    bool verbose = false;     // per-round cluster counts on stderr (XRT_TUNING stage_dump)
};

struct GpuBuildInfo {
    int depth = 0;       // levels of the two-child tree incl. the leaf level
    int nNodes = 0;      // two-child nodes in use (leaves - 1)
    int nAllocated = 0;  // BvhNode records written (n - 1; the ones below a collapsed leaf are unreferenced)
    int iterations = 0;  // PLOC merge rounds
    float pad = 0.f;
    float sahCost = 0.f; // relative to the root area, Ct = traversalCost, Ci = 1
    float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0}; // bounds of all triangles
};

// dTrisId: n >= 2 triangles in primitive-id order on the device. Writes n - 1 BvhNode records (root = record 0) and the leaf-ordered
// triangle arrays (3-float4 records from dTrisId, 4-float4 plane records from dFastId). Returns cudaErrorNotSupported when the
// clustering does not converge or the tree is deeper than the traversal stacks allow (the caller falls back to the host builder).
cudaError_t buildPlocDevice(const float4* dTrisId, uint32_t n, float4* dTrisLeafOrder, BvhNode* dNodes, cudaStream_t st, GpuBuildInfo* info,
                            const float4* dFastId, float4* dFastLeafOrder, const PlocParams& params);

// Collapses the two-child tree dNodes (root = record 0, any builder) into eight-child quantised nodes on the device — the same
// greedy "open the largest inner child" and octant slot assignment as bvh.cpp: collapseBvh8, one thread per wide node, one launch
// per tree level. *dNodes8 is cudaMalloc'ed to the exact size (caller frees); dFtris8 (nTris * 4 float4, caller-allocated) receives
// the plane records in node order, gathered from the leaf-ordered dFastLeafOrder.
cudaError_t collapseBvh8Device(const BvhNode* dNodes, uint32_t nTris, const float4* dFastLeafOrder, Bvh8Node** dNodes8, uint32_t* nNodes8,
                               float4* dFtris8, int* depth8, cudaStream_t st);

} // namespace xrt
