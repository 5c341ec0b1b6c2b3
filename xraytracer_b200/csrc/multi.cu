// multi.cu — several GPUs behind ONE xrtg_scene handle (include/xrtgpu.h: xrtg_scene_create_multi), single process.
//
// The path shards by independent samples: device g of G renders sample indices [g*spp/G, (g+1)*spp/G) of every
// pixel against its own replica of the scene arrays — the role ParallelRenderer::render (renderer.cpp:83-99) gives to the
// host's cores. The only exchange step is the sum of the G per-pixel SUM buffers followed by the reference's
// `image /= n_samples` (renderer.cpp:98). Both happen in ONE kernel per device over NVLink peer memory: device g pulls slice g
// of every device's buffer (P2P loads), adds the G values in device order, divides, and stores the slice straight into
// device 0's image (P2P stores). Each NVLink port carries 1/G of the data in each direction; no NCCL, no staging copy, no
// separate scale pass. The same kernel serves one-process-per-GPU deployments through CUDA IPC (xrtg_reduce_finalize).
#include <algorithm>
#include <thread>
#include "scene_impl.h"

using namespace xrt;

namespace {

constexpr int kMaxParts = 16;
struct PartList {
    const float* p[kMaxParts];
};

// out[i] = (parts[0][i] + parts[1][i] + ...) / divisor, i in [first, first + count). 128-bit loads / stores where the slice
// allows it. IEEE division like k_finalize (the reference divides, renderer.cpp:98); divisor <= 0 leaves the sum.
__global__ void __launch_bounds__(256) k_reduce_finalize(PartList parts, int nParts, float* __restrict__ out, size_t first, size_t count, float divisor)
{
    const size_t tid = size_t(blockIdx.x) * blockDim.x + threadIdx.x, stride = size_t(gridDim.x) * blockDim.x;
    size_t head = (4 - (first & 3)) & 3;
    if (head > count) head = count; // scalars up to the first 16-byte boundary
    const size_t nVec = (count - head) / 4;
    for (size_t v = tid; v < nVec; v += stride) {
        const size_t i = first + head + 4 * v;
        float4 acc = *reinterpret_cast<const float4*>(parts.p[0] + i);
        for (int k = 1; k < nParts; ++k) {
            const float4 x = *reinterpret_cast<const float4*>(parts.p[k] + i);
            acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
        }
        if (divisor > 0.f) { acc.x = acc.x / divisor; acc.y = acc.y / divisor; acc.z = acc.z / divisor; acc.w = acc.w / divisor; }
        *reinterpret_cast<float4*>(out + i) = acc;
    }
    const size_t nScalar = head + (count - head - 4 * nVec);
    for (size_t k = tid; k < nScalar; k += stride) {
        const size_t i = k < head ? first + k : first + head + 4 * nVec + (k - head);
        float acc = parts.p[0][i];
        for (int j = 1; j < nParts; ++j) acc += parts.p[j][i];
        out[i] = divisor > 0.f ? acc / divisor : acc;
    }
}

// counts the guard bytes (scene_impl.h: DevBuf) that no longer hold the pattern
__global__ void __launch_bounds__(256) k_check_guard(const unsigned char* __restrict__ lo, const unsigned char* __restrict__ hi, int* bad)
{
    const int i = threadIdx.x;
    if (lo[i] != kGuardPattern) atomicAdd(bad, 1);
    if (hi[i] != kGuardPattern) atomicAdd(bad, 1);
}

void rebindPointers(xrtg_scene* s)
{
    DScene& ds = s->ds;
    ds.nodes = static_cast<const float4*>(s->nodes.d);
    ds.nodes4 = static_cast<const float4*>(s->nodes4.d);
    ds.nodes8 = static_cast<const uint4*>(s->nodes8.d);
    ds.tris = static_cast<const float4*>(s->tris.d);
    ds.tris_id = static_cast<const float4*>(s->trisId.d);
    ds.ftris = static_cast<const float4*>(s->ftris.d);
    ds.ftris_id = static_cast<const float4*>(s->ftrisId.d);
    ds.ftris8 = static_cast<const float4*>(s->ftris8.d);
    ds.smallBlock = static_cast<const float4*>(s->smallBlock.d);
    ds.prims = static_cast<const float4*>(s->prims.d);
    ds.spheres = static_cast<const float4*>(s->spheres.d);
    ds.boxes = static_cast<const float4*>(s->boxes.d);
    ds.lights = static_cast<const DLight*>(s->lights.d);
    ds.dlights = static_cast<const DDelta*>(s->dlights.d);
    ds.media = static_cast<const DMedium*>(s->media.d);
    ds.grids = static_cast<const DGrid*>(s->grids.d);
}

} // namespace

namespace xrt {

void launchReduceFinalize(cudaStream_t st, const float* const* parts, int nParts, float* out, size_t first, size_t count, float divisor)
{
    if (count == 0) return;
    PartList pl{};
    for (int k = 0; k < nParts; ++k) pl.p[k] = parts[k];
    const int grid = int(std::min<size_t>((count / 4 + 255) / 256 + 1, 148 * 8));
    k_reduce_finalize<<<grid, 256, 0, st>>>(pl, nParts, out, first, count, divisor);
}

int checkGuards(xrtg_scene* s, int* violations)
{
    static_assert(kGuardBytes == 256, "k_check_guard runs one thread per guard byte");
    *violations = 0;
    const std::vector<xrtg_scene*> all = s->replicas.empty() ? std::vector<xrtg_scene*>{s} : s->replicas;
    for (xrtg_scene* r : all) {
        CU(cudaSetDevice(r->device));
        DevBuf* bufs[] = {&r->q0[0], &r->q0[1], &r->q1[0], &r->q1[1], &r->q2[0], &r->q2[1], &r->hits, &r->s0, &r->s1, &r->s2, &r->radiance, &r->ctrl, &r->accum,
                          &r->outDev, &r->mt, &r->mti, &r->stats, &r->jitter, &r->rayTmp[0], &r->rayTmp[1], &r->rayTmp[2], &r->rayTmp[3], &r->primMask, &r->multiOut};
        int* bad = nullptr;
        CU(cudaMalloc(&bad, sizeof(int)));
        CU(cudaMemsetAsync(bad, 0, sizeof(int), r->stream));
        for (DevBuf* b : bufs)
            if (b->raw) k_check_guard<<<1, 256, 0, r->stream>>>(static_cast<const unsigned char*>(b->raw), static_cast<const unsigned char*>(b->p) + b->bytes, bad);
        int h = 0;
        const cudaError_t e = cudaMemcpyAsync(&h, bad, sizeof(int), cudaMemcpyDeviceToHost, r->stream);
        const cudaError_t e2 = cudaStreamSynchronize(r->stream);
        cudaFree(bad);
        if (e != cudaSuccess || e2 != cudaSuccess) return fail(XRTG_ERR_CUDA, std::string("guard check: ") + cudaGetErrorString(e != cudaSuccess ? e : e2));
        *violations += h;
    }
    return 0;
}

// A replica of `primary` on `device`: shares the primary's pinned host arrays (scene data is built ONCE), owns its device
// copies, its stream and its wave workspace.
int createReplica(const xrtg_scene* primary, int device, xrtg_scene** out)
{
    *out = nullptr;
    CU(cudaSetDevice(device));
    auto r = std::make_unique<xrtg_scene>();
    r->device = device;
    CU(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking));
    for (auto& e : r->ev) CU(cudaEventCreate(&e));
    CU(cudaMallocHost(&r->statsHost, sizeof(unsigned long long) * kStatCount));
    CU(cudaMallocHost(&r->ctrlHost, sizeof(uint32_t) * 16));
    // Arrays with a pinned host copy are uploaded from it (shared with the primary); arrays a device-side build produced exist in
    // the primary's HBM only and are copied device to device — over NVLink when the devices are peers — with no host staging.
    Mirror *src[kSceneArrays], *dst[kSceneArrays];
    sceneArrays(const_cast<xrtg_scene*>(primary), src);
    sceneArrays(r.get(), dst);
    for (int k = 0; k < kSceneArrays; ++k) {
        if (dst[k] == &r->grids) continue; // (its own table, below)
        if (src[k]->h || src[k]->bytes == 0) { if (int rc = dst[k]->mirrorOf(*src[k])) return rc; }
        else {
            if (int rc = dst[k]->allocDevice(src[k]->bytes)) return rc;
            CU(cudaMemcpyPeerAsync(dst[k]->d, device, src[k]->d, primary->device, src[k]->bytes, r->stream));
        }
    }
    for (const auto& g : primary->gridData) {
        auto m = std::make_unique<Mirror>();
        if (int rc = m->mirrorOf(*g)) return rc;
        r->gridData.push_back(std::move(m));
    }
    // the grid descriptors hold DEVICE pointers to the voxel arrays: this replica needs its own copy of that (small) table
    if (int rc = r->grids.alloc(primary->grids.bytes)) return rc;
    if (primary->grids.bytes) {
        std::memcpy(r->grids.h, primary->grids.h, primary->grids.bytes);
        DGrid* G = static_cast<DGrid*>(r->grids.h);
        for (size_t i = 0; i < r->gridData.size(); ++i) {
            G[i].data = static_cast<const float*>(r->gridData[i]->d);
            r->gridTex.emplace_back();
            if (G[i].tex) { // the primary has a 3-D texture copy of this grid: so does the replica, on its own device
                auto t = std::make_unique<GridTexture>();
                if (int rc = t->create(G[i].nx, G[i].ny, G[i].nz)) return rc;
                G[i].tex = (unsigned long long)t->tex;
                r->gridTex.back() = std::move(t);
            }
        }
    }
    r->ds = primary->ds;
    rebindPointers(r.get());
    r->info = primary->info;
    r->tuning = primary->tuning;
    r->maxShadowPerPath = primary->maxShadowPerPath;
    r->smallTriVerts = primary->smallTriVerts;
    std::memcpy(r->boundsLo, primary->boundsLo, sizeof(r->boundsLo));
    std::memcpy(r->boundsHi, primary->boundsHi, sizeof(r->boundsHi));
    r->hasBounds = primary->hasBounds;
    if (int rc = uploadAll(r.get(), false)) return rc;
    CU(cudaStreamSynchronize(r->stream));
    *out = r.release();
    return 0;
}

// Peer access between every pair of the scene's devices (idempotent). Returns false if some pair cannot map each other's memory.
static bool enablePeerAccess(xrtg_scene* s)
{
    bool all = true;
    for (xrtg_scene* a : s->replicas)
        for (xrtg_scene* b : s->replicas) {
            if (a == b || a->device == b->device) continue; // (one device listed twice: plain local memory)
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, a->device, b->device) != cudaSuccess || !can) { all = false; continue; }
            cudaSetDevice(a->device);
            const cudaError_t e = cudaDeviceEnablePeerAccess(b->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) all = false;
            cudaGetLastError();
        }
    return all;
}

// xrtg_scene_upload on a multi-device handle: the scene arrays cross PCIe ONCE (pinned host -> device 0) and reach the other
// replicas over NVLink in a binomial tree of peer copies (replica k pulls from replica k minus its highest set bit, as soon as
// that one's copy has landed: log2(G) rounds, every device sends and receives at most one array set per round) instead of G
// concurrent host-to-device uploads of the same bytes. Only the small table of grid descriptors (device pointers) comes from each
// replica's own host copy. *done = false (nothing enqueued) when some pair of devices cannot map each other's memory.
int broadcastUpload(xrtg_scene* s, bool* done)
{
    *done = false;
    const int G = int(s->replicas.size());
    if (G < 2) return 0;
    if (!s->peerChecked) { s->peerAll = enablePeerAccess(s); s->peerChecked = true; }
    if (!s->peerAll) return 0;
    for (xrtg_scene* r : s->replicas) {
        CU(cudaSetDevice(r->device));
        if (!r->uploadEvent) CU(cudaEventCreateWithFlags(&r->uploadEvent, cudaEventDisableTiming));
    }
    xrtg_scene* root = s->replicas[0];
    CU(cudaSetDevice(root->device));
    if (int rc = uploadAll(root, false)) return rc;
    CU(cudaEventRecord(root->uploadEvent, root->stream));
    size_t h2d = root->info.upload_bytes;
    for (int k = 1; k < G; ++k) {
        int top = 1;
        while (top * 2 <= k) top *= 2;
        xrtg_scene* parent = s->replicas[size_t(k - top)];
        xrtg_scene* r = s->replicas[size_t(k)];
        CU(cudaSetDevice(r->device));
        CU(cudaStreamWaitEvent(r->stream, parent->uploadEvent, 0));
        Mirror *src[kSceneArrays], *dst[kSceneArrays];
        sceneArrays(parent, src);
        sceneArrays(r, dst);
        for (int a = 0; a < kSceneArrays; ++a) {
            if (dst[a] == &r->grids) { // this replica's own table of grid descriptors
                if (r->grids.bytes) CU(cudaMemcpyAsync(r->grids.d, r->grids.h, r->grids.bytes, cudaMemcpyHostToDevice, r->stream));
                h2d += r->grids.bytes;
                continue;
            }
            if (dst[a]->bytes && dst[a]->bytes == src[a]->bytes)
                CU(cudaMemcpyPeerAsync(dst[a]->d, r->device, src[a]->d, parent->device, dst[a]->bytes, r->stream));
        }
        for (size_t g = 0; g < r->gridData.size() && g < parent->gridData.size(); ++g) {
            CU(cudaMemcpyPeerAsync(r->gridData[g]->d, r->device, parent->gridData[g]->d, parent->device, r->gridData[g]->bytes, r->stream));
            if (g < r->gridTex.size() && r->gridTex[g])
                if (int rc = r->gridTex[g]->fill(r->gridData[g]->d, cudaMemcpyDeviceToDevice, r->stream)) return rc;
        }
        CU(cudaEventRecord(r->uploadEvent, r->stream));
        r->info.upload_bytes = 0; // (nothing of this replica's arrays crossed PCIe)
    }
    root->info.upload_bytes = h2d;
    *done = true;
    return 0;
}

int renderMulti(xrtg_scene* s, const xrtg_camera* cam, const xrtg_render_params* p, float* rgbHost, xrtg_stats* stats)
{
    const int G = int(s->replicas.size());
    if (G > kMaxParts) return fail(XRTG_ERR_UNSUPPORTED, "too many devices behind one scene handle");
    const size_t n = size_t(p->width) * size_t(p->height) * 3, bytes = n * sizeof(float);
    if (!s->peerChecked) { s->peerAll = enablePeerAccess(s); s->peerChecked = true; }
    for (xrtg_scene* r : s->replicas) {
        CU(cudaSetDevice(r->device));
        if (int rc = r->outDev.ensure(bytes)) return rc; // this device's per-pixel SUM
        if (!r->doneEvent) CU(cudaEventCreateWithFlags(&r->doneEvent, cudaEventDisableTiming));
        if (!r->pullEvent) CU(cudaEventCreateWithFlags(&r->pullEvent, cudaEventDisableTiming));
    }
    CU(cudaSetDevice(s->device));
    if (int rc = s->multiOut.ensure(bytes)) return rc; // the final image, on device 0

    // ---- one host thread per device: its share of the samples, SUM only ----
    const size_t nG = size_t(G);
    std::vector<int> rcs(nG, 0);
    std::vector<std::string> errs(nG);
    std::vector<xrtg_stats> sts(nG);
    std::vector<std::thread> threads;
    Timer wall;
    for (int g = 0; g < G; ++g)
        threads.emplace_back([&, g]() {
            xrtg_scene* r = s->replicas[size_t(g)];
            xrtg_render_params pg = *p;
            const int lo = int(int64_t(g) * p->spp / G), hi = int(int64_t(g + 1) * p->spp / G);
            pg.spp = hi - lo;
            pg.sample_offset = p->sample_offset + lo;
            pg.spp_total = p->spp_total > 0 ? p->spp_total : p->spp;
            pg.flags |= XRTG_FLAG_SUM_ONLY;
            int rc = 0;
            if (cudaSetDevice(r->device) != cudaSuccess) rc = fail(XRTG_ERR_CUDA, "cudaSetDevice failed");
            else if (pg.spp > 0) rc = renderOnStream(r, cam, &pg, static_cast<float*>(r->outDev.p), r->stream, stats ? &sts[size_t(g)] : nullptr);
            else if (cudaMemsetAsync(r->outDev.p, 0, bytes, r->stream) != cudaSuccess) rc = fail(XRTG_ERR_CUDA, "cudaMemsetAsync failed");
            if (rc == 0 && cudaEventRecord(r->doneEvent, r->stream) != cudaSuccess) rc = fail(XRTG_ERR_CUDA, "cudaEventRecord failed");
            rcs[size_t(g)] = rc;
            if (rc) errs[size_t(g)] = xrtg_last_error(); // the error slot is thread-local: carry it to the caller's thread
        });
    for (auto& t : threads) t.join();
    for (int g = 0; g < G; ++g)
        if (rcs[size_t(g)]) return fail(rcs[size_t(g)], "device " + std::to_string(s->replicas[size_t(g)]->device) + ": " + errs[size_t(g)]);

    NvtxRange nvtx("multi-GPU: fused peer-memory reduce + finalize");
    // ---- fused reduce + finalize over peer memory ----
    const int divisor = (p->flags & XRTG_FLAG_SUM_ONLY) ? 0 : (p->spp_total > 0 ? p->spp_total : p->spp);
    float* finalImg = static_cast<float*>(s->multiOut.p);
    const float* parts[kMaxParts];
    if (s->peerAll) {
        for (int g = 0; g < G; ++g) parts[g] = static_cast<const float*>(s->replicas[size_t(g)]->outDev.p);
        const size_t per = ((n + size_t(G) - 1) / size_t(G) + 3) & ~size_t(3); // slice boundaries on 16-byte boundaries
        for (int g = 0; g < G; ++g) {
            xrtg_scene* r = s->replicas[size_t(g)];
            CU(cudaSetDevice(r->device));
            for (int k = 0; k < G; ++k)
                if (k != g) CU(cudaStreamWaitEvent(r->stream, s->replicas[size_t(k)]->doneEvent, 0));
            if (g == 0) CU(cudaEventRecord(s->ev[2], s->stream));
            const size_t first = std::min(n, per * size_t(g)), count = std::min(n, per * size_t(g + 1)) - first;
            launchReduceFinalize(r->stream, parts, G, finalImg, first, count, float(divisor));
            CU(cudaGetLastError());
            if (g == 0) CU(cudaEventRecord(s->ev[3], s->stream));
            else CU(cudaEventRecord(r->pullEvent, r->stream));
        }
        CU(cudaSetDevice(s->device));
        for (int g = 1; g < G; ++g) CU(cudaStreamWaitEvent(s->stream, s->replicas[size_t(g)]->pullEvent, 0));
    }
    else {
        // no peer mapping between some pair of devices (PCIe topologies): device 0 fetches the other SUM buffers with
        // cudaMemcpyPeerAsync (staged by the driver) and runs the same kernel locally
        CU(cudaSetDevice(s->device));
        while (s->peerStage.size() < size_t(G)) s->peerStage.push_back(std::make_unique<DevBuf>());
        parts[0] = static_cast<const float*>(s->outDev.p);
        for (int g = 1; g < G; ++g) {
            if (int rc = s->peerStage[size_t(g)]->ensure(bytes)) return rc;
            CU(cudaStreamWaitEvent(s->stream, s->replicas[size_t(g)]->doneEvent, 0));
            CU(cudaMemcpyPeerAsync(s->peerStage[size_t(g)]->p, s->device, s->replicas[size_t(g)]->outDev.p, s->replicas[size_t(g)]->device, bytes, s->stream));
            parts[g] = static_cast<const float*>(s->peerStage[size_t(g)]->p);
        }
        CU(cudaEventRecord(s->ev[2], s->stream));
        launchReduceFinalize(s->stream, parts, G, finalImg, 0, n, float(divisor));
        CU(cudaGetLastError());
        CU(cudaEventRecord(s->ev[3], s->stream));
    }
    const float deviceMs = wall.ms();
    CU(cudaMemcpyAsync(rgbHost, finalImg, bytes, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    if (stats) {
        xrtg_stats t{};
        for (int g = 0; g < G; ++g) {
            const xrtg_stats& a = sts[size_t(g)];
            t.samples += a.samples; t.closest_rays += a.closest_rays; t.shadow_rays += a.shadow_rays; t.dropped_samples += a.dropped_samples;
            t.nodes_visited += a.nodes_visited; t.tris_tested += a.tris_tested; t.nodes_visited_shadow += a.nodes_visited_shadow;
            t.tris_tested_shadow += a.tris_tested_shadow; t.tracking_steps += a.tracking_steps; t.kernel_launches += a.kernel_launches;
            t.extend_launches += a.extend_launches; t.shade_launches += a.shade_launches; t.connect_launches += a.connect_launches;
            t.primary_hits += a.primary_hits; t.bounce_entries += a.bounce_entries; t.bounce_launches += a.bounce_launches;
            t.rays_traced += a.rays_traced; t.truncated_paths += a.truncated_paths;
            t.untraced_closest += a.untraced_closest; t.untraced_shadow += a.untraced_shadow;
            t.render_ms = std::max(t.render_ms, a.render_ms); // devices run concurrently: the slowest one sets the time
            t.extend_ms = std::max(t.extend_ms, a.extend_ms); t.connect_ms = std::max(t.connect_ms, a.connect_ms);
            t.shade_ms = std::max(t.shade_ms, a.shade_ms); t.other_ms = std::max(t.other_ms, a.other_ms);
        }
        t.kernel_launches += uint64_t(s->peerAll ? G : 1);
        t.n_devices = G;
        CU(cudaEventElapsedTime(&t.reduce_ms, s->ev[2], s->ev[3]));
        t.h2d_ms = deviceMs; // host wall clock from the first launch to the last reduce launch (all devices)
        *stats = t;
    }
    return 0;
}

} // namespace xrt

extern "C" {

int xrtg_scene_create_multi(const xrtg_scene_desc* desc, int ngpus, const int* devices, uint32_t build_flags, xrtg_scene** out)
{
    if (!out) return fail(XRTG_ERR_INVALID, "out is NULL");
    *out = nullptr;
    const int have = xrtg_device_count();
    if (have <= 0) return fail(XRTG_ERR_NO_DEVICE, "no CUDA device (libxrtgpu has no CPU fallback)");
    if (ngpus < 1 || ngpus > kMaxParts) return fail(XRTG_ERR_INVALID, "ngpus out of range");
    std::vector<int> devs;
    for (int g = 0; g < ngpus; ++g) {
        const int d = devices ? devices[g] : g;
        if (d < 0 || d >= have) return fail(XRTG_ERR_INVALID, "device index " + std::to_string(d) + " out of range (" + std::to_string(have) + " visible)");
        // (a device may be listed more than once: two replicas then share it — useful on a one-GPU box to exercise the very same
        //  split + fused reduce a multi-GPU box runs)
        devs.push_back(d);
    }
    xrtg_scene* primary = nullptr;
    if (int rc = xrtg_scene_create2(desc, devs[0], build_flags, &primary)) return rc;
    primary->replicas.push_back(primary);
    for (int g = 1; g < ngpus; ++g) {
        xrtg_scene* r = nullptr;
        if (int rc = createReplica(primary, devs[size_t(g)], &r)) { xrtg_scene_destroy(primary); return rc; }
        primary->replicas.push_back(r);
    }
    primary->info.n_devices = ngpus;
    primary->info.device_bytes *= uint64_t(ngpus);
    *out = primary;
    return 0;
}

int xrtg_scene_check_guards(xrtg_scene* s, int* violations)
{
    if (!s || !violations) return fail(XRTG_ERR_INVALID, "NULL argument");
    return checkGuards(s, violations);
}

int xrtg_exchange_buffer(xrtg_scene* s, int slot, size_t bytes, void** device_ptr)
{
    if (!s || !device_ptr || slot < 0 || slot >= 4 || bytes == 0) return fail(XRTG_ERR_INVALID, "bad argument");
    CU(cudaSetDevice(s->device));
    if (int rc = s->exchange[slot].ensure(bytes)) return rc;
    *device_ptr = s->exchange[slot].p;
    return 0;
}

int xrtg_ipc_export(const void* device_ptr, unsigned char handle[XRTG_IPC_HANDLE_BYTES])
{
    static_assert(sizeof(cudaIpcMemHandle_t) == XRTG_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
    if (!device_ptr || !handle) return fail(XRTG_ERR_INVALID, "NULL argument");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, const_cast<void*>(device_ptr)));
    std::memcpy(handle, &h, sizeof(h));
    return 0;
}

int xrtg_ipc_open(xrtg_scene* s, const unsigned char handle[XRTG_IPC_HANDLE_BYTES], void** device_ptr)
{
    if (!s || !handle || !device_ptr) return fail(XRTG_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(s->device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof(h));
    CU(cudaIpcOpenMemHandle(device_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

int xrtg_ipc_close(xrtg_scene* s, void* device_ptr)
{
    if (!s || !device_ptr) return fail(XRTG_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(s->device));
    CU(cudaIpcCloseMemHandle(device_ptr));
    return 0;
}

int xrtg_reduce_finalize(xrtg_scene* s, const float* const* parts, int nparts, float* out, size_t first, size_t count, float divisor, void* cuda_stream)
{
    if (!s || !parts || !out || nparts < 1 || nparts > kMaxParts) return fail(XRTG_ERR_INVALID, "bad argument");
    for (int k = 0; k < nparts; ++k)
        if (!parts[k]) return fail(XRTG_ERR_INVALID, "NULL part");
    CU(cudaSetDevice(s->device));
    launchReduceFinalize(static_cast<cudaStream_t>(cuda_stream), parts, nparts, out, first, count, divisor);
    CU(cudaGetLastError());
    return 0;
}

} // extern "C"
