// kernels.h — launch table exported by each instantiation of wavefront.cuh (exact / fast).
#pragma once
#include <cuda_runtime.h>
#include "device_types.h"

namespace xrt {

struct KernelTable {
    void (*seedMt)(cudaStream_t, const DWave&);
    void (*raygen)(cudaStream_t, const DCamera&, const DQueues&, const DWave&, const float* jitter);
    // jitter: nullptr (production: the RNG draws it) or the parity hook's supplied samples
    void (*primary)(cudaStream_t, const DScene&, const DCamera&, const DQueues&, const DWave&, bool brute, int missMode, bool count,
                    unsigned long long* stats, const float* jitter);
    // small scenes: per-32-pixel candidate masks of the primary rays (k_primary_masks)
    void (*primaryMasks)(cudaStream_t, const TriBoxes&, int nTris, int width, uint32_t nPixels, unsigned long long* masks);
    void (*extend)(cudaStream_t, const DScene&, const DQueues&, int src, int bounce, int brute /*0 BVH, 1 brute force, 2 small-scene smem*/, bool count, unsigned long long* stats,
                   int refillThreshold, int stepsPerVote, int leafThreshold);
    void (*connect)(cudaStream_t, const DScene&, const DQueues&, int bounce, int brute, bool count, unsigned long long* stats,
                    int refillThreshold, int stepsPerVote, int leafThreshold);
    void (*shadeSurface)(cudaStream_t, const DScene&, const DQueues&, const DWave&, int src, int bounce);
    // small scenes: shade + connect + extend of one bounce in one kernel (k_bounce_small)
    void (*bounceSmall)(cudaStream_t, const DScene&, const DQueues&, const DWave&, int src, int bounce, unsigned long long* stats);
    void (*shadeVolume)(cudaStream_t, const DScene&, const DQueues&, const DWave&, int src, int bounce, bool brute, bool count,
                        unsigned long long* stats);
    // shallow BVHs: every volume path run to completion in one launch (k_volume_paths)
    void (*volumePaths)(cudaStream_t, const DScene&, const DQueues&, const DWave&, bool brute, int maxIter, int refillThreshold, int stepsPerVote, bool count,
                        unsigned long long* stats);
    void (*accumulate)(cudaStream_t, const DQueues&, const DWave&, float* accum, unsigned long long* stats);
    void (*finalize)(cudaStream_t, const float* accum, float* out, size_t n, float divisor);
    // mode: 0 k_trace, 1 simple kernels, 2 small-scene tracer (see launchTraceRays)
    void (*traceRays)(cudaStream_t, const DScene&, const DQueues&, const float* org, const float* dir, const float* tmax, long long n,
                      bool anyhit, bool brute, float4* out, unsigned long long* stats, int mode);
    // parity hook: the compact hit queue k_primary wrote -> one record per path id
    void (*scatterPrimaryHits)(cudaStream_t, const DQueues&, float4* out, uint32_t nPaths);
    void (*genJitter)(cudaStream_t, const DWave&, int spp, float* jitter);
};

const KernelTable& exactKernels(); // kernels_exact.cu : -fmad=false, mt19937 sample stream
const KernelTable& fastKernels();  // kernels_fast.cu  : FMA, Philox counter RNG

} // namespace xrt
