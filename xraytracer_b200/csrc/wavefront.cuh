// wavefront.cuh — the sm_100a wavefront path-tracing kernels of libxrtgpu.so.
//
// Included by two translation units:
//   kernels_exact.cu  XRT_EXACT=1, compiled with -fmad=false : reproduces the reference's arithmetic (no FMA
//                     contraction, IEEE div/sqrt, same operation order) and its per-pixel std::mt19937 sample
//                     stream (renderer.cpp:35-36, sampler.h:48-49). One sample per pixel per wave.
//   kernels_fast.cu   XRT_EXACT=0, -use_fast_math : counter-based Philox4x32-7 keyed (seed,pixel)/(sample,block), several
//                     samples per pixel per wave, plane-equation triangle records instead of Moeller-Trumbore.
//
// Pipeline per wave (queue sizes live in device memory, so a wave is enqueued without a host round trip), chosen per scene:
//   small scene (<= 64 triangles):  primary (raygen fused with the bounce-0 closest hit, compact hit-only queue)
//                                   -> bounce_small per bounce (shade + NEE shadow rays + next closest hit + next RR, fused)
//   shallow BVH, volume integrator: primary -> volume_paths (every path run to completion, lockstep delta-tracking walk)
//   shallow BVH otherwise:          primary -> shade -> connect (any hit) -> extend (closest hit) ...      (simple kernels)
//   deep BVH:                       raygen -> trace<closest> -> shade -> trace<any> ...   (resumable traversal, warp refill)
//   then accumulate (sample-ordered sum, invalid samples dropped) and, after the last wave, finalize.
//
// Reference citations (paths relative to /root/reference/Src) are on each device function. The source is split by topic:
//   wf_math_rng.cuh   vector algebra, RNGs           wf_intersect.cuh  primitive tests, traversal, small-scene loops
//   wf_trace.cuh      raygen + traversal kernels     wf_shade.cuh      surface shading, fused bounce kernel
//   wf_volume.cuh     media + volume kernels         wf_launch.cuh     accumulate / finalize, launchers
#pragma once
#include <cfloat>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "device_types.h"
#include <xrtgpu.h>

#ifndef XRT_EXACT
#error "define XRT_EXACT and XRT_NS before including wavefront.cuh"
#endif

namespace xrt {
#ifndef XRT_TMA_STAGE
// k_bounce_small: queue tiles staged per thread with cp.async (0, default) or by bulk copies of the TMA unit completing on an
// mbarrier (1; SASS: UBLKCP + SYNCS.ARRIVE.TRANS64). Measured on c3 (profiles/r02_notes.md): 9 241 vs 9 053 Msamples/s — with
// cp.async every warp waits for its OWN 4 x 16 B, with the bulk copy all warps of the CTA wake on the same barrier, and the
// tile is only 8 KB, too small for the TMA unit's per-copy cost to amortise. Both variants pass the same parity tests.
#define XRT_TMA_STAGE 0
#endif

// k_bounce_small: survivors appended per warp from warp-private chunks of output slots, no CTA barrier in the tile loop (1), or
// per CTA with one atomic and three barriers per tile (0).
#ifndef XRT_WARP_APPEND
#define XRT_WARP_APPEND 1
#endif
// The same for k_primary's compact bounce-0 queue (then every consumer of that queue skips the dead slots). Measured slightly
// SLOWER on c3 (9 942 vs 10 017 Msamples/s): k_primary already pays one barrier round per 256 paths only, and the warp-private
// chunks scatter neighbouring pixels' hits over the queue. Off.
#ifndef XRT_WARP_APPEND_PRIMARY
#define XRT_WARP_APPEND_PRIMARY 0
#endif

// Deep scenes: queue entries are read once and written once per launch — loaded / stored with the streaming (evict-first) cache
// hint so that the GBs of queue traffic do not push the tree and the triangle records (78 MB on c4) out of the 126 MB L2.
#ifndef XRT_STREAM_HINTS
#define XRT_STREAM_HINTS 1
#endif
#ifndef XRT_VOL_MINB
#define XRT_VOL_MINB 4 // resident CTAs per SM k_volume_paths is compiled for (register budget 65536 / (128 x this))
#endif

namespace XRT_NS {
__device__ __forceinline__ float4 qload(const float4* p) { return XRT_STREAM_HINTS ? __ldcs(p) : *p; }
__device__ __forceinline__ void qstore(float4* p, float4 v) { if (XRT_STREAM_HINTS) __stcs(p, v); else *p = v; }

constexpr bool kExact = (XRT_EXACT != 0);
constexpr int kBlock = 128;          // threads per CTA of the traversal kernels (the surface shade kernel uses kShadeBlock)
constexpr int kStackSmem = 24;       // traversal stack entries kept in shared memory per thread
constexpr int kStackLocal = 40;      // overflow entries (local memory) of the run-to-completion walks; host builder depth limit is 56
constexpr int kVolMinBlocks = XRT_VOL_MINB;
constexpr int kStackLocalDeep = 104; // ... of k_trace, the kernel of deep trees: two-child trees of up to 120 levels (device-built
                                     // PLOC trees are deeper than top-down SAH ones: 41-53 levels on the 1 M-triangle scene)
constexpr float kPI = 3.14159265359; // geometry.h:10
constexpr float kRayEps = 1e-3f;     // geometry.h:23

#include "wf_math_rng.cuh"
#include "wf_intersect.cuh"
#include "wf_trace.cuh"
#include "wf_trace8.cuh"
#include "wf_shade.cuh"
#include "wf_volume.cuh"
#include "wf_launch.cuh"

} // namespace XRT_NS
} // namespace xrt
