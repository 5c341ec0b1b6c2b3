// device_types.h — PODs shared by the host side of libxrtgpu.so (api.cpp) and the sm_100a kernels
// (wavefront.cuh). Layouts are chosen for 128-bit coalesced loads: every array is an array of float4.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace xrt {

// meta word of a primitive's shading record (prims[4*id+3].w)
enum : uint32_t {
    kMetaKindMask = 3u,      // xrtg_object_kind
    kMetaHasMaterial = 4u,
    kMetaShadowOutside = 8u, // small scenes: shadow rays from this primitive can start behind a hull-pruned plane (api.cu)
    kMetaLightShift = 8,     // (area light index + 1) in bits 8..19
    kMetaMediumShift = 20,   // (medium index + 1) in bits 20..31
};

struct DLight {      // area light, world space (light.cpp:6-14, 49-57, 84-90)
    float4 v0_kind;  // v0 | kind as int bits   (sphere: centre)
    float4 e1_r;     // e1 = v1 - v0 | radius
    float4 e2;       // e2 = v2 - v0
    float4 Ng;       // e1 x e2 (un-normalised)
    float4 Le;
    float4 v1, v2;   // triangle sampling uses A,B,C directly
};

struct DDelta {      // delta light (light.cpp:115-142)
    float4 p_kind;   // position or direction | kind
    float4 L;        // color * intensity
};

struct DMedium {
    int kind;
    float g;
    float sigma_a[3], sigma_s[3], sigma_t[3];
    float densityMul, majorant, invMajorant;
    int grid;
    int grey; // sigma_a and sigma_s are the same in all three channels (every medium of the reference's examples: nee.cpp:54, volume.cpp:49)
};

struct DGrid {
    const float* data; // nx*ny*nz, x fastest
    int nx, ny, nz;
    float origin[3];
    float voxel, background;
    float invVoxel; // 1 / voxel (throughput instantiation; the exact one divides like grid.h:71-77)
    unsigned long long tex; // cudaTextureObject_t over a 3-D cudaArray copy of `data` (linear filtering, border = 0), or 0: the
                            // throughput instantiation's lookups go through the texture unit when background == 0
};

// Everything the kernels need to know about the scene; passed by value (fits the 4 KB param space).
struct DScene {
    const float4* nodes;   // 4 float4 per BVH node (bvh.h: BvhNode)
    const float4* nodes4;  // deep trees only: the same tree collapsed to four children per node, 8 float4 each (bvh.h: Bvh4Node); or nullptr
    const uint4* nodes8;   // deep trees: eight-child nodes with 8-bit quantised child boxes, 80 B = 5 uint4 each (bvh.h: Bvh8Node); or nullptr
    const float4* tris;    // 3 float4 per triangle in LEAF order: v0|prim id, e1|flags(bit0 emitter), e2|0
    const float4* tris_id; // same triangles in PRIMITIVE-ID order (brute-force parity path), mesh triangles only
    // throughput instantiation only: 4 float4 per triangle = plane-equation form (N|d, n1|d1, n2|d2, id|flags|0|0), see
    // triangleRecord() in wavefront.cuh; leaf order / primitive-id order like tris / tris_id
    const float4* ftris;
    const float4* ftris_id;
    const float4* ftris8;  // the same records in the order the eight-child tree stores them (triBase + offset), or nullptr
    const float4* smallBlock; // plane-paired triangles of a small scene (small_scene.h) or nullptr
    int smallBlockF4;         // its size in float4 (0 = not available)
    const float4* prims;   // 4 float4 per primitive id: shading record
                           //   tri:    n0|ng.x  n1|ng.y  n2|ng.z  albedo|meta
                           //   sphere: centre|radius  -  -  albedo|meta          box: -  -  -  0|meta
    const float4* spheres; // 2 float4 per sphere: centre|radius , (prim id, emitter flag, 0, 0) as int bits
    const float4* boxes;   // 2 float4 per box: pmin|prim id bits , pmax|0
    const DLight* lights;
    const DDelta* dlights;
    const DMedium* media;
    const DGrid* grids;
    int nTris, nSpheres, nBoxes, nLights, nDelta, nPrims, nBruteTris, nMedia, nGrids;
};

// pixel bounding boxes of the (at most 64) mesh triangles of a small scene, passed to k_primary_masks by value
struct TriBoxes {
    int4 b[64];
};

struct DCamera {
    float c2w[16];
    float scale, aspect;
};

// One wave of paths. Ray queue entry i (SoA over three float4 arrays):
//   q0 = origin.xyz | throughput.x     q1 = direction.xyz | throughput.y
//   q2 = throughput.z | path id | depth | rng counter            (last three as int bits)
// Shadow queue entry: s0 = origin.xyz | tmax   s1 = direction.xyz | path id   s2 = contribution.rgb | 0
struct DQueues {
    float4 *q0[2], *q1[2], *q2[2]; // ping-pong ray queues
    float4* hits;                  // t,u,v | prim id, indexed like the ray queue being extended
    float4 *s0, *s1, *s2;          // shadow queue
    float4* radiance;              // per path id: rgb | unused
    uint32_t* ctrl;                // per bounce b: ctrl[8b+0]=#rays  +1=#shadow  +2..+4 = work-fetch cursors
    unsigned long long* stats;     // device-side statistics (kStat*), for kernels that take no separate pointer
};

// k_bounce_small appends survivors per warp from warp-private chunks of kAppendChunk output slots; the slots a warp has not used at
// the end of a launch are marked with kDeadPath in the path word. The queues therefore hold up to (kAppendChunk - 1) x resident
// warps entries more than there are paths: 148 SMs x 16 CTAs x 4 warps x 63 < kAppendSlack.
constexpr uint32_t kAppendChunk = 64;
constexpr uint32_t kDeadPath = 0xffffffffu;
constexpr uint32_t kAppendSlack = 1u << 20;
enum { kCtrlStride = 8, kCtrlRays = 0, kCtrlShadow = 1, kCtrlFetchExtend = 2, kCtrlFetchShade = 3, kCtrlFetchConnect = 4 };

// device-side statistics (uint64 each)
enum { kStatClosest = 0, kStatShadow, kStatDropped, kStatNodes, kStatTris, kStatNodesAny, kStatTrisAny, kStatSteps, kStatPrimaryHits, kStatBounceEntries, kStatScissored, kStatTruncated, kStatUntracedClosest, kStatUntracedShadow, kStatCount };

// Exact unsigned division by a run-time constant (Granlund & Montgomery): q = (t + ((n - t) >> sh1)) >> sh2 with t = umulhi(m, n).
// Path ids are split into (sample, pixel) and pixels into (row, column) by every kernel that opens a path's RNG; the compiler's
// 32-bit division is ~20 instructions, this is 4 (k_primary spent 18 % of its instructions on three divisions per path).
struct FastDiv {
    uint32_t m, sh1, sh2, d;
#ifdef __CUDACC__
    __device__ __forceinline__ uint32_t div(uint32_t n) const { const uint32_t t = uint32_t((uint64_t(m) * uint64_t(n)) >> 32); return (t + ((n - t) >> sh1)) >> sh2; }
    __device__ __forceinline__ void divmod(uint32_t n, uint32_t& q, uint32_t& r) const { q = div(n); r = n - q * d; }
#endif
};
#ifdef __CUDACC__
__host__ __device__
#endif
inline FastDiv makeFastDiv(uint32_t d)
{
    FastDiv f{};
    f.d = d ? d : 1u;
    uint32_t l = 0;
    while ((1ull << l) < f.d) ++l; // ceil(log2 d)
    f.m = uint32_t(((1ull << 32) * ((1ull << l) - f.d)) / f.d + 1ull);
    f.sh1 = l < 1u ? l : 1u;
    f.sh2 = l > 0u ? l - 1u : 0u;
    return f;
}

constexpr uint32_t kWaveScissorSkip = 1u;
struct DWave {
    int width, height;
    uint32_t nPixels;       // pixels of the whole image (stride of the per-pixel mt19937 state)
    // A wave covers the pixel range [pixelBase, pixelBase + wavePixels) x samplesThisWave samples: the whole image unless the
    // per-path workspace is so large (hundreds of area lights -> shadow-queue entries per path) that one sample of every pixel
    // would not fit the workspace budget. Path id = s * wavePixels + (pixel - pixelBase).
    uint32_t pixelBase, wavePixels;
    FastDiv byWavePixels, byWidth; // exact division by wavePixels / width (kept in step with them by setWave())
    uint32_t nPaths;        // wavePixels * samplesThisWave
    uint32_t sampleBase;    // index of the first sample of this wave (sample_offset + done so far)
    uint32_t samplesThisWave;
    int integrator, maxDepth;
    uint32_t seed;
    uint32_t flags;         // kWaveScissorSkip: out-of-scissor pixels carry no per-sample radiance (k_primary skips the write, k_accumulate the read)
    // Screen-space scissor [sx0, sx1) x [sy0, sy1): pixels outside it cannot see the scene's bounding box (every sample of such a
    // pixel is a miss), so the primary kernel resolves them without generating a ray. Full image when unknown.
    int sx0, sy0, sx1, sy1;
    // Small scenes (<= 64 mesh triangles): per 32 consecutive pixels (linear index >> 5) a 64-bit mask of the triangles whose
    // screen-space bounding box touches them (k_primary_masks, rebuilt per render from the camera). The primary kernel tests only
    // those candidates instead of walking the BVH: a superset of what any ray through these pixels can hit, so the closest hit is
    // unchanged. nullptr = not available (more triangles, or a vertex beside / behind the camera: those set every bit instead).
    const unsigned long long* primMask;
    // exact mode: per-pixel mt19937 state, word-major [624][nPixels], and the per-pixel cursor
    uint32_t* mt;
    uint32_t* mti;
};

} // namespace xrt
