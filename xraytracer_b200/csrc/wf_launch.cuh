// wf_launch.cuh — part of wavefront.cuh (included inside namespace xrt::XRT_NS, in this order): accumulate / finalize kernels and the host-side launchers behind kernels.h.
// ---------------------------------------------------------------------------------------------------------
// accumulate: validate each sample like renderer.cpp:57-73 (NaN / inf / any negative channel -> dropped, the
// divisor is unchanged) and add the wave's samples to the pixel sum IN SAMPLE ORDER (bit-exact vs
// Image::addPixel order). One thread per pixel, no atomics.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_accumulate(DQueues q, DWave w, float* __restrict__ accum, unsigned long long* stats)
{
    uint32_t dropped = 0;
    // pixels outside the primary kernel's screen-space scissor hold no per-sample radiance (k_primary does not write it, this kernel
    // does not read it: 86 % of the volume frame, 36 % of the Cornell frame): every sample of such a pixel is the miss colour
    const bool scissored = (w.flags & kWaveScissorSkip) != 0;
    const float mx = w.integrator == XRTG_INT_DIRECT ? float(0.18) : (w.integrator == XRTG_INT_WHITTED ? float(0.235294) : 0.f);
    const float my = w.integrator == XRTG_INT_DIRECT ? float(0.18) : (w.integrator == XRTG_INT_WHITTED ? float(0.67451) : 0.f);
    const float mz = w.integrator == XRTG_INT_DIRECT ? float(0.18) : (w.integrator == XRTG_INT_WHITTED ? float(0.843137) : 0.f);
    for (uint32_t lp = blockIdx.x * blockDim.x + threadIdx.x; lp < w.wavePixels; lp += gridDim.x * blockDim.x) {
        const uint32_t p = w.pixelBase + lp;
        float ax = accum[3 * size_t(p)], ay = accum[3 * size_t(p) + 1], az = accum[3 * size_t(p) + 2];
        if (scissored) {
            uint32_t i, j;
            w.byWidth.divmod(p, i, j);
            if (!(int(j) >= w.sx0 && int(j) < w.sx1 && int(i) >= w.sy0 && int(i) < w.sy1)) {
                if (mx != 0.f || my != 0.f || mz != 0.f)
                    for (uint32_t s = 0; s < w.samplesThisWave; ++s) { ax += mx; ay += my; az += mz; } // (the same additions, in the same order)
                accum[3 * size_t(p)] = ax; accum[3 * size_t(p) + 1] = ay; accum[3 * size_t(p) + 2] = az;
                continue;
            }
        }
        for (uint32_t s = 0; s < w.samplesThisWave; ++s) {
            const float4 r = q.radiance[size_t(s) * w.wavePixels + lp];
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
// The result depends on the kernel AND on the current device (a multi-GPU scene launches the same kernels on every device from
// one process), so it is cached per (function, device).
inline int gridFor(const void* fn, int block = kBlock)
{
    struct Entry { const void* fn; int dev, block, grid; };
    static thread_local std::vector<Entry> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    for (const Entry& e : cache)
        if (e.fn == fn && e.dev == dev && e.block == block) return e.grid;
    int sms = 0, perSm = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, fn, block, 0);
    const int grid = sms * (perSm < 1 ? 1 : perSm);
    cache.push_back(Entry{fn, dev, block, grid});
    return grid;
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
    k_raygen<<<gridFor((const void*)k_raygen), kBlock, 0, st>>>(cam, q, w, jitter);
}
inline void launchPrimary(cudaStream_t st, const DScene& sc, const DCamera& cam, const DQueues& q, const DWave& w, bool brute, int missMode,
                          bool count, unsigned long long* stats, const float* jitter)
{
    if (jitter) k_primary<false, true><<<gridFor((const void*)k_primary<false, true>), kBlock, 0, st>>>(sc, cam, q, w, brute, missMode, stats, jitter);
    else if (count) k_primary<true, false><<<gridFor((const void*)k_primary<true, false>), kBlock, 0, st>>>(sc, cam, q, w, brute, missMode, stats, nullptr);
    else k_primary<false, false><<<gridFor((const void*)k_primary<false, false>), kBlock, 0, st>>>(sc, cam, q, w, brute, missMode, stats, nullptr);
}
inline void launchPrimaryMasks(cudaStream_t st, const TriBoxes& boxes, int nTris, int width, uint32_t nPixels, unsigned long long* masks)
{
    const uint32_t nSeg = (nPixels + 31u) / 32u;
    k_primary_masks<<<int(std::min<uint32_t>((nSeg + kBlock - 1) / kBlock, 148u * 8u)), kBlock, 0, st>>>(boxes, nTris, width, nPixels, masks);
}
// k_trace instantiation for (closest / any hit, counters, two- / four-child tree)
template <bool ANY>
inline void launchTrace(cudaStream_t st, const DScene& sc, const DQueues& q, int src, int bounce, int brute, bool count, unsigned long long* stats,
                        float4* anyOut, int thr, int spv, int leafThr)
{
    if constexpr (!kExact) {
        // throughput instantiation, deep scene: the eight-child quantised tree (k_trace8); brute force stays with k_trace
        if (sc.nodes8 != nullptr && sc.ftris8 != nullptr && brute == 0) {
            // (leafThr carries the development switch "6 CTAs per SM / 80 registers" in bit 8)
            const bool six = (leafThr & 256) != 0;
            leafThr &= 255;
            if (count) k_trace8<ANY, true, 8><<<gridFor((const void*)k_trace8<ANY, true, 8>), kBlock, 0, st>>>(sc, q, src, bounce, stats, anyOut, thr, spv, leafThr);
            else if (six) k_trace8<ANY, false, 6><<<gridFor((const void*)k_trace8<ANY, false, 6>), kBlock, 0, st>>>(sc, q, src, bounce, stats, anyOut, thr, spv, leafThr);
            else k_trace8<ANY, false, 8><<<gridFor((const void*)k_trace8<ANY, false, 8>), kBlock, 0, st>>>(sc, q, src, bounce, stats, anyOut, thr, spv, leafThr);
            return;
        }
    }
    const int g[4] = {gridFor((const void*)k_trace<ANY, false, false>), gridFor((const void*)k_trace<ANY, true, false>),
                      gridFor((const void*)k_trace<ANY, false, true>), gridFor((const void*)k_trace<ANY, true, true>)};
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
    const int h0 = gridFor((const void*)k_extend_simple<false>), h1 = gridFor((const void*)k_extend_simple<true>);
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
    const int h0 = gridFor((const void*)k_connect_simple<false>), h1 = gridFor((const void*)k_connect_simple<true>);
    if (thr <= 0) {
        if (count) k_connect_simple<true><<<h1, kBlock, 0, st>>>(sc, q, bounce, brute, stats, nullptr);
        else k_connect_simple<false><<<h0, kBlock, 0, st>>>(sc, q, bounce, brute, stats, nullptr);
        return;
    }
    launchTrace<true>(st, sc, q, 0, bounce, brute, count, stats, nullptr, thr, spv, leafThr);
}
inline void launchShadeSurface(cudaStream_t st, const DScene& sc, const DQueues& q, const DWave& w, int src, int bounce)
{
    k_shade_surface<<<gridFor((const void*)k_shade_surface, kShadeBlock), kShadeBlock, 0, st>>>(sc, q, w, src, bounce);
}
inline void launchBounceSmall(cudaStream_t st, const DScene& sc, const DQueues& q, const DWave& w, int src, int bounce, unsigned long long* stats)
{
    const int g0 = gridFor((const void*)k_bounce_small<false>), g1 = gridFor((const void*)k_bounce_small<true>);
    const bool grouped = !kExact && sc.smallBlockF4 > 0 && sc.nBoxes == 0;
    if (grouped) k_bounce_small<true><<<g1, kBlock, 0, st>>>(sc, q, w, src, bounce, stats);
    else k_bounce_small<false><<<g0, kBlock, 0, st>>>(sc, q, w, src, bounce, stats);
}
inline void launchShadeVolume(cudaStream_t st, const DScene& sc, const DQueues& q, const DWave& w, int src, int bounce, bool brute, bool count,
                              unsigned long long* stats)
{
    const int g0 = gridFor((const void*)k_shade_volume<false>), g1 = gridFor((const void*)k_shade_volume<true>);
    if (count) k_shade_volume<true><<<g1, kBlock, 0, st>>>(sc, q, w, src, bounce, brute, stats);
    else k_shade_volume<false><<<g0, kBlock, 0, st>>>(sc, q, w, src, bounce, brute, stats);
}
inline void launchVolumePaths(cudaStream_t st, const DScene& sc, const DQueues& q, const DWave& w, bool brute, int maxIter, int threshold, int stepsPerVote, bool count,
                              unsigned long long* stats)
{
    if (count) k_volume_paths<true, kVolMinBlocks><<<gridFor((const void*)k_volume_paths<true, kVolMinBlocks>), kBlock, 0, st>>>(sc, q, w, brute, maxIter, threshold, stepsPerVote, stats);
    else k_volume_paths<false, kVolMinBlocks><<<gridFor((const void*)k_volume_paths<false, kVolMinBlocks>), kBlock, 0, st>>>(sc, q, w, brute, maxIter, threshold, stepsPerVote, stats);
}
inline void launchAccumulate(cudaStream_t st, const DQueues& q, const DWave& w, float* accum, unsigned long long* stats)
{
    const int grid = int((w.wavePixels + kBlock - 1) / kBlock);
    k_accumulate<<<grid, kBlock, 0, st>>>(q, w, accum, stats);
}
inline void launchFinalize(cudaStream_t st, const float* accum, float* out, size_t n, float divisor)
{
    const int grid = int(std::min<size_t>((n + kBlock - 1) / kBlock, 148 * 16));
    k_finalize<<<grid, kBlock, 0, st>>>(accum, out, n, divisor);
}
// parity hook: rays go through the SAME persistent traversal kernel the renderer uses. `out` = n float4 (device):
// closest -> the hit queue itself is returned by the caller; any hit -> occlusion flags are written to `out`.
// mode (the production pipeline the hook should exercise, chosen by the caller exactly like renderOnStream chooses it):
//   0 = k_trace (the resumable traversal of deep BVHs; also the classic exact-instantiation hook)
//   1 = k_extend_simple / k_connect_simple (shallow BVHs with more than 64 triangles)
//   2 = k_hook_small over SmallTracer (small scenes: the closest / any-hit code of k_bounce_small)
inline void launchTraceRays(cudaStream_t st, const DScene& sc, const DQueues& q, const float* org, const float* dir, const float* tmax,
                            long long n, bool anyhit, bool brute, float4* out, unsigned long long* stats, int mode)
{
    const int grid = int(std::min<long long>((n + kBlock - 1) / kBlock, 148 * 16));
    k_pack_rays<<<grid > 0 ? grid : 1, kBlock, 0, st>>>(q, org, dir, tmax, uint32_t(n), anyhit ? 1 : 0);
    if (mode == 2) {
        const bool grouped = !kExact && sc.smallBlockF4 > 0 && sc.nBoxes == 0;
        if (grouped) k_hook_small<true><<<gridFor((const void*)k_hook_small<true>), kBlock, 0, st>>>(sc, q, uint32_t(n), anyhit ? 1 : 0, out);
        else k_hook_small<false><<<gridFor((const void*)k_hook_small<false>), kBlock, 0, st>>>(sc, q, uint32_t(n), anyhit ? 1 : 0, out);
    }
    else if (mode == 1) {
        if (anyhit) k_connect_simple<false><<<gridFor((const void*)k_connect_simple<false>), kBlock, 0, st>>>(sc, q, 0, brute ? 1 : 0, stats, out);
        else k_extend_simple<false><<<gridFor((const void*)k_extend_simple<false>), kBlock, 0, st>>>(sc, q, 0, 0, brute ? 1 : 0, stats);
    }
    else if (anyhit) launchTrace<true>(st, sc, q, 0, 0, brute, false, stats, out, 16, 1, 8);
    else launchTrace<false>(st, sc, q, 0, 0, brute, false, stats, nullptr, 16, 1, 8);
}
// parity hook: compact bounce-0 queue of k_primary -> one hit record per path id (misses: prim = -1)
__global__ void __launch_bounds__(kBlock) k_hook_fill_miss(float4* __restrict__ out, uint32_t n)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = make_float4(FLT_MAX, 0.f, 0.f, __int_as_float(-1));
}
__global__ void __launch_bounds__(kBlock) k_hook_scatter_hits(DQueues q, float4* __restrict__ out)
{
    const uint32_t n = q.ctrl[kCtrlRays];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (!deadEntry(q.q2[0][i])) out[__float_as_int(q.q2[0][i].y)] = q.hits[i];
}
inline void launchScatterPrimaryHits(cudaStream_t st, const DQueues& q, float4* out, uint32_t nPaths)
{
    const int grid = int(std::min<uint32_t>((nPaths + kBlock - 1) / kBlock, 148 * 16));
    k_hook_fill_miss<<<grid > 0 ? grid : 1, kBlock, 0, st>>>(out, nPaths);
    k_hook_scatter_hits<<<grid > 0 ? grid : 1, kBlock, 0, st>>>(q, out);
}
