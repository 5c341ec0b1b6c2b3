// lbvh.h — GPU-side linear BVH build (lbvh.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "bvh.h"

namespace xrt {

struct LbvhInfo {
    int depth = 0;
    int nNodes = 0;
    float pad = 0.f;
};

// dTrisId: n triangles (3 float4 each: v0|id, e1|flags, e2|0) in primitive-id order, resident on the device.
// Writes n-1 BvhNode records (root = node 0) and the leaf-ordered triangle array. Requires n >= 2.
// dFastId / dFastLeafOrder (optional): the 4-float4 plane-equation records, gathered into leaf order the same way.
cudaError_t buildLbvhDevice(const float4* dTrisId, uint32_t n, float4* dTrisLeafOrder, BvhNode* dNodes, cudaStream_t st, LbvhInfo* info,
                            const float4* dFastId = nullptr, float4* dFastLeafOrder = nullptr);

} // namespace xrt
