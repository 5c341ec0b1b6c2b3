// wf_launch.cuh — part of wavefront.cuh (included inside namespace xrt::XRT_NS, in this order): accumulate / finalize kernels and the host-side launchers behind kernels.h.
// ---------------------------------------------------------------------------------------------------------
// accumulate: validate each sample like renderer.cpp:57-73 (NaN / inf / any negative channel -> dropped, the
// divisor is unchanged) and add the wave's samples to the pixel sum IN SAMPLE ORDER (bit-exact vs
// Image::addPixel order). One thread per pixel, no atomics.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_accumulate(DQueues q, DWave w, float* __restrict__ accum, unsigned long long* stats)
{
    uint32_t dropped = 0;
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < w.nPixels; p += gridDim.x * blockDim.x) {
        float ax = accum[3 * size_t(p)], ay = accum[3 * size_t(p) + 1], az = accum[3 * size_t(p) + 2];
        for (uint32_t s = 0; s < w.samplesThisWave; ++s) {
            const float4 r = q.radiance[size_t(s) * w.nPixels + p];
            if (isnan(r.x) || isnan(r.y) || isnan(r.z)) { ++dropped; continue; }
            else if (isinf(r.x) || isinf(r.y) || isinf(r.z)) { ++dropped; continue; }
            else if (r.x < 0 || r.y < 0 || r.z < 0) { ++dropped; continue; }
            ax += r.x; ay += r.y; az += r.z;
        }
        accum[3 * size_t(p)] = ax; accum[3 * size_t(p) + 1] = ay; accum[3 * size_t(p) + 2] = az;
    }
    if (dropped) atomicAdd(stats + kStatDropped, (unsigned long long)dropped);
}

// image /= Vec3f(n_samples) (renderer.cpp:98) — IEEE division like the reference; divisor 0 = leave the sum
__global__ void __launch_bounds__(kBlock) k_finalize(const float* __restrict__ accum, float* __restrict__ out, size_t n, float divisor)
{
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
        out[i] = divisor > 0.f ? accum[i] / divisor : accum[i];
}

// ---------------------------------------------------------------------------------------------------------
// host-side launchers (called from api.cpp through the table in kernels.h)
// ---------------------------------------------------------------------------------------------------------
// persistent grids: (resident CTAs per SM for this kernel) x (number of SMs) — 148 on B200
inline int gridFor(const void* fn, int block = kBlock)
{
    static thread_local int cachedDev = -1;
    static thread_local int sms = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != cachedDev) {
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cachedDev = dev;
    }
    int perSm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, fn, block, 0);
    return sms * (perSm < 1 ? 1 : perSm);
}

inline void launchSeedMt(cudaStream_t st, const DWave& w)
{
    const int grid = int((w.nPixels + kBlock - 1) / kBlock);
    k_seed_mt<<<grid, kBlock, 0, st>>>(w.mt, w.mti, w.nPixels);
}
inline void launchGenJitter(cudaStream_t st, const DWave& w, int spp, float* jitter)
{
    const int grid = int((w.nPixels + kBlock - 1) / kBlock);
    k_gen_jitter<<<grid, kBlock, 0, st>>>(w, spp, jitter);
}
inline void launchRaygen(cudaStream_t st, const DCamera& cam, const DQueues& q, const DWave& w, const float* jitter)
{
    static thread_local int grid = 0;
    if (!grid) grid = gridFor((const void*)k_raygen);
    k_raygen<<<grid, kBlock, 0, st>>>(cam, q, w, jitter);
}
inline void launchPrimary(cudaStream_t st, const DScene& sc, const DCamera& cam, const DQueues& q, const DWave& w, bool brute, int missMode,
                          bool count, unsigned long long* stats)
{
    static thread_local int g0 = 0, g1 = 0;
    if (!g0) { g0 = gridFor((const void*)k_primary<false>); g1 = gridFor((const void*)k_primary<true>); }
    if (count) k_primary<true><<<g1, kBlock, 0, st>>>(sc, cam, q, w, brute, missMode, stats);
    else k_primary<false><<<g0, kBlock, 0, st>>>(sc, cam, q, w, brute, missMode, stats);
}
// k_trace instantiation for (closest / any hit, counters, two- / four-child tree)
template <bool ANY>
inline void launchTrace(cudaStream_t st, const DScene& sc, const DQueues& q, int src, int bounce, int brute, bool count, unsigned long long* stats,
                        float4* anyOut, int thr, int spv, int leafThr)
{
    static thread_local int g[4] = {0, 0, 0, 0};
    if (!g[0]) {
        g[0] = gridFor((const void*)k_trace<ANY, false, false>); g[1] = gridFor((const void*)k_trace<ANY, true, false>);
        g[2] = gridFor((const void*)k_trace<ANY, false, true>); g[3] = gridFor((const void*)k_trace<ANY, true, true>);
    }
    const bool wide = sc.nodes4 != nullptr;
    if (wide) {
        if (count) k_trace<ANY, true, true><<<g[3], kBlock, 0, st>>>(sc, q, src, bounce, brute, stats, anyOut, thr, spv, leafThr);
        else k_trace<ANY, false, true><<<g[2], kBlock, 0, st>>>(sc, q, src, bounce, brute, stats, anyOut, thr, spv, leafThr);
    }
    else {
        if (count) k_trace<ANY, true, false><<<g[1], kBlock, 0, st>>>(sc, q, src, bounce, brute, stats, anyOut, thr, spv, leafThr);
        else k_trace<ANY, false, false><<<g[0], kBlock, 0, st>>>(sc, q, src, bounce, brute, stats, anyOut, thr, spv, leafThr);
    }
}
inline void launchExtend(cudaStream_t st, const DScene& sc, const DQueues& q, int src, int bounce, int brute, bool count, unsigned long long* stats,
                         int thr, int spv, int leafThr)
{
    static thread_local int h0 = 0, h1 = 0;
    if (!h0) { h0 = gridFor((const void*)k_extend_simple<false>); h1 = gridFor((const void*)k_extend_simple<true>); }
    if (thr <= 0) { // shallow BVH: simple run-to-completion kernel
        if (count) k_extend_simple<true><<<h1, kBlock, 0, st>>>(sc, q, src, bounce, brute, stats);
        else k_extend_simple<false><<<h0, kBlock, 0, st>>>(sc, q, src, bounce, brute, stats);
        return;
    }
    launchTrace<false>(st, sc, q, src, bounce, brute, count, stats, nullptr, thr, spv, leafThr);
}
inline void launchConnect(cudaStream_t st, const DScene& sc, const DQueues& q, int bounce, int brute, bool count, unsigned long long* stats,
                          int thr, int spv, int leafThr)
{
    static thread_local int h0 = 0, h1 = 0;
    if (!h0) { h0 = gridFor((const void*)k_connect_simple<false>); h1 = gridFor((const void*)k_connect_simple<true>); }
    if (thr <= 0) {
        if (count) k_connect_simple<true><<<h1, kBlock, 0, st>>>(sc, q, bounce, brute, stats);
        else k_connect_simple<false><<<h0, kBlock, 0, st>>>(sc, q, bounce, brute, stats);
        return;
    }
    launchTrace<true>(st, sc, q, 0, bounce, brute, count, stats, nullptr, thr, spv, leafThr);
}
inline void launchShadeSurface(cudaStream_t st, const DScene& sc, const DQueues& q, const DWave& w, int src, int bounce)
{
    static thread_local int grid = 0;
    if (!grid) grid = gridFor((const void*)k_shade_surface, kShadeBlock);
    k_shade_surface<<<grid, kShadeBlock, 0, st>>>(sc, q, w, src, bounce);
}
inline void launchBounceSmall(cudaStream_t st, const DScene& sc, const DQueues& q, const DWave& w, int src, int bounce, unsigned long long* stats)
{
    static thread_local int g0 = 0, g1 = 0;
    if (!g0) { g0 = gridFor((const void*)k_bounce_small<false>); g1 = gridFor((const void*)k_bounce_small<true>); }
    const bool grouped = !kExact && sc.smallBlockF4 > 0 && sc.nBoxes == 0;
    if (grouped) k_bounce_small<true><<<g1, kBlock, 0, st>>>(sc, q, w, src, bounce, stats);
    else k_bounce_small<false><<<g0, kBlock, 0, st>>>(sc, q, w, src, bounce, stats);
}
inline void launchShadeVolume(cudaStream_t st, const DScene& sc, const DQueues& q, const DWave& w, int src, int bounce, bool brute, bool count,
                              unsigned long long* stats)
{
    static thread_local int g0 = 0, g1 = 0;
    if (!g0) { g0 = gridFor((const void*)k_shade_volume<false>); g1 = gridFor((const void*)k_shade_volume<true>); }
    if (count) k_shade_volume<true><<<g1, kBlock, 0, st>>>(sc, q, w, src, bounce, brute, stats);
    else k_shade_volume<false><<<g0, kBlock, 0, st>>>(sc, q, w, src, bounce, brute, stats);
}
inline void launchVolumePaths(cudaStream_t st, const DScene& sc, const DQueues& q, const DWave& w, bool brute, int maxIter, int threshold, int stepsPerVote, bool count,
                              unsigned long long* stats)
{
    static thread_local int g0 = 0, g1 = 0, g2 = 0, sel = 0;
    if (!g0) {
        g0 = gridFor((const void*)k_volume_paths<false, 5>); g1 = gridFor((const void*)k_volume_paths<true, 4>);
        g2 = gridFor((const void*)k_volume_paths<false, 4>);
        const char* e = std::getenv("XRT_VOLUME_MINB");
        sel = e ? std::atoi(e) : 4;
    }
    if (count) k_volume_paths<true, 4><<<g1, kBlock, 0, st>>>(sc, q, w, brute, maxIter, threshold, stepsPerVote, stats);
    else if (sel == 4) k_volume_paths<false, 4><<<g2, kBlock, 0, st>>>(sc, q, w, brute, maxIter, threshold, stepsPerVote, stats);
    else k_volume_paths<false, 5><<<g0, kBlock, 0, st>>>(sc, q, w, brute, maxIter, threshold, stepsPerVote, stats);
}
inline void launchAccumulate(cudaStream_t st, const DQueues& q, const DWave& w, float* accum, unsigned long long* stats)
{
    const int grid = int((w.nPixels + kBlock - 1) / kBlock);
    k_accumulate<<<grid, kBlock, 0, st>>>(q, w, accum, stats);
}
inline void launchFinalize(cudaStream_t st, const float* accum, float* out, size_t n, float divisor)
{
    const int grid = int(std::min<size_t>((n + kBlock - 1) / kBlock, 148 * 16));
    k_finalize<<<grid, kBlock, 0, st>>>(accum, out, n, divisor);
}
// parity hook: rays go through the SAME persistent traversal kernel the renderer uses. `out` = n float4 (device):
// closest -> the hit queue itself is returned by the caller; any hit -> occlusion flags are written to `out`.
inline void launchTraceRays(cudaStream_t st, const DScene& sc, const DQueues& q, const float* org, const float* dir, const float* tmax,
                            long long n, bool anyhit, bool brute, float4* out, unsigned long long* stats)
{
    const int grid = int(std::min<long long>((n + kBlock - 1) / kBlock, 148 * 16));
    k_pack_rays<<<grid > 0 ? grid : 1, kBlock, 0, st>>>(q, org, dir, tmax, uint32_t(n), anyhit ? 1 : 0);
    if (anyhit) launchTrace<true>(st, sc, q, 0, 0, brute, false, stats, out, 16, 1, 8);
    else launchTrace<false>(st, sc, q, 0, 0, brute, false, stats, nullptr, 16, 1, 8);
}
