// scene_impl.h — internals of the xrtg_scene handle shared by api.cu (ingest, single-device render, parity hooks) and multi.cu
// (scene replicas on several devices, the fused peer-memory reduce + finalize).
#pragma once
#include <chrono>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h> // header-only NVTX v3: ranges cost a function-pointer test unless a profiler is attached
#include <xrtgpu.h>
#include "device_types.h"

namespace xrt {

// thread-local last-error slot behind xrtg_last_error()
int fail(int code, const std::string& msg);

#define CU(call)                                                                                                    \
    do {                                                                                                            \
        cudaError_t e__ = (call);                                                                                   \
        if (e__ != cudaSuccess)                                                                                     \
            return ::xrt::fail(e__ == cudaErrorMemoryAllocation ? XRTG_ERR_OOM : XRTG_ERR_CUDA,                     \
                               std::string(#call) + ": " + cudaGetErrorString(e__));                                \
    } while (0)

// pinned host array + device mirror. A replica of a scene on another device SHARES the pinned host copy of the primary
// (ownsHost = false) and owns only its device copy.
struct Mirror {
    void* h = nullptr;
    void* d = nullptr;
    size_t bytes = 0;
    bool ownsHost = true;
    int alloc(size_t n)
    {
        release();
        bytes = n;
        if (n == 0) return 0;
        CU(cudaMallocHost(&h, n));
        CU(cudaMalloc(&d, n));
        return 0;
    }
    int allocDevice(size_t n) // device array only: produced by kernels (gpu_build.cu); the pinned copy is made on demand
    {
        release();
        bytes = n;
        if (n == 0) return 0;
        CU(cudaMalloc(&d, n));
        return 0;
    }
    void adoptDevice(void* dev, size_t n) // takes ownership of a cudaMalloc'ed array
    {
        release();
        d = dev;
        bytes = n;
    }
    int ensureHost() // pinned copy of a device-produced array (re-uploads, replicas on other devices)
    {
        if (h || bytes == 0) return 0;
        CU(cudaMallocHost(&h, bytes));
        CU(cudaMemcpy(h, d, bytes, cudaMemcpyDeviceToHost));
        return 0;
    }
    void shareHost(const Mirror& src) // a replica adopts the primary's (lazily made) pinned copy
    {
        if (!h && src.h && src.bytes == bytes) { h = src.h; ownsHost = false; }
    }
    int mirrorOf(const Mirror& src) // device copy on the CURRENT device of src's host array
    {
        release();
        bytes = src.bytes;
        h = src.h;
        ownsHost = false;
        if (bytes == 0) return 0;
        CU(cudaMalloc(&d, bytes));
        return 0;
    }
    void release()
    {
        if (h && ownsHost) cudaFreeHost(h);
        if (d) cudaFree(d);
        h = d = nullptr;
        bytes = 0;
        ownsHost = true;
    }
    ~Mirror() { release(); }
};

// 3-D texture copy of a density grid (throughput instantiation): cudaArray (the driver's tiled 3-D layout: the eight voxels of a
// trilinear lookup share one or two cache lines instead of four rows 1 KB / 256 KB apart) + a texture object with linear
// filtering, unnormalised coordinates and border addressing (outside = 0 = the grid's background).
struct GridTexture {
    cudaArray_t array = nullptr;
    cudaTextureObject_t tex = 0;
    int nx = 0, ny = 0, nz = 0;
    int create(int nx_, int ny_, int nz_)
    {
        nx = nx_; ny = ny_; nz = nz_;
        const cudaChannelFormatDesc fmt = cudaCreateChannelDesc<float>();
        CU(cudaMalloc3DArray(&array, &fmt, make_cudaExtent(size_t(nx), size_t(ny), size_t(nz))));
        cudaResourceDesc res{};
        res.resType = cudaResourceTypeArray;
        res.res.array.array = array;
        cudaTextureDesc td{};
        td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeBorder;
        td.filterMode = cudaFilterModeLinear;
        td.readMode = cudaReadModeElementType;
        td.normalizedCoords = 0;
        CU(cudaCreateTextureObject(&tex, &res, &td, nullptr));
        return 0;
    }
    int fill(const void* src, cudaMemcpyKind kind, cudaStream_t st) // src = nx*ny*nz floats, x fastest (host or device)
    {
        cudaMemcpy3DParms p{};
        p.srcPtr = make_cudaPitchedPtr(const_cast<void*>(src), size_t(nx) * sizeof(float), size_t(nx), size_t(ny));
        p.dstArray = array;
        p.extent = make_cudaExtent(size_t(nx), size_t(ny), size_t(nz));
        p.kind = kind;
        CU(cudaMemcpy3DAsync(&p, st));
        return 0;
    }
    ~GridTexture()
    {
        if (tex) cudaDestroyTextureObject(tex);
        if (array) cudaFreeArray(array);
    }
};

// Device workspace buffer with GUARD BANDS: kGuardBytes of a known pattern before and after the usable range, verified on demand
// (xrtg_scene_check_guards; the GPU tests check them after every render). compute-sanitizer is closed on this pool, so an
// out-of-bounds queue write is caught by its footprint instead: every queue / counter / radiance buffer of the wave scheduler
// is one of these, and neighbouring cudaMalloc blocks are not adjacent, so a stray store lands in a guard or faults.
constexpr size_t kGuardBytes = 256;
constexpr unsigned char kGuardPattern = 0xA5;
struct DevBuf {
    void* p = nullptr;   // usable range
    void* raw = nullptr; // allocation = guard | usable | guard
    size_t bytes = 0;
    int ensure(size_t n)
    {
        if (n <= bytes) return 0;
        if (raw) cudaFree(raw);
        p = raw = nullptr;
        bytes = 0;
        const size_t padded = (n + 255) & ~size_t(255);
        CU(cudaMalloc(&raw, padded + 2 * kGuardBytes));
        CU(cudaMemset(raw, kGuardPattern, kGuardBytes));
        CU(cudaMemset(static_cast<char*>(raw) + kGuardBytes + padded, kGuardPattern, kGuardBytes));
        p = static_cast<char*>(raw) + kGuardBytes;
        bytes = padded;
        return 0;
    }
    ~DevBuf() { if (raw) cudaFree(raw); }
};

// plain allocation (no guard bands): exportable through CUDA IPC, where the handle names the allocation's BASE address
struct PlainBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int ensure(size_t n)
    {
        if (n <= bytes) return 0;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        CU(cudaMalloc(&p, n));
        bytes = n;
        return 0;
    }
    ~PlainBuf() { if (p) cudaFree(p); }
};

// NVTX range over a scope (nsys / ncu timelines: scene ingest, every wave of a render, the multi-GPU reduce)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

struct Timer {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    float ms() const { return std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

// Development / test switches of the pipeline selection (xrtg_tuning; -1 = the measured default). None of them changes what is
// computed, only which kernels compute it. Set with xrtg_scene_set_tuning, or once at scene creation from the XRT_TUNING
// environment variable ("key=value,key=value"); the render path itself reads no environment variable.
struct Tuning {
    xrtg_tuning t;
    bool stageDump = false;
    Tuning() { std::memset(&t, 0xff, sizeof(t)); } // every field -1
};

} // namespace xrt

struct __attribute__((visibility("hidden"))) xrtg_scene { // (the C header forward-declares it inside its default-visibility region)
    int device = 0;
    cudaStream_t stream = nullptr; // uploads + host-buffer renders
    // scene arrays (pinned host copy + device copy)
    xrt::Mirror nodes, nodes4, nodes8, tris, trisId, ftris, ftrisId, ftris8, smallBlock, prims, spheres, boxes, lights, dlights, media, grids;
    std::vector<std::unique_ptr<xrt::Mirror>> gridData;
    std::vector<std::unique_ptr<xrt::GridTexture>> gridTex; // one per grid (entries may be null: no texture for that grid)
    xrt::DScene ds{};
    xrtg_scene_info info{};
    xrt::Tuning tuning;
    int maxShadowPerPath = 1;
    std::vector<float> smallTriVerts; // small scenes (<= 64 mesh triangles): 9 floats per triangle, primitive-id order (primary-ray masks)
    float boundsLo[3] = {0, 0, 0}, boundsHi[3] = {0, 0, 0}; // world bounds of every primitive (valid if hasBounds)
    bool hasBounds = false;
    // workspace
    xrt::DevBuf q0[2], q1[2], q2[2], hits, s0, s1, s2, radiance, ctrl, accum, outDev, mt, mti, stats, jitter, rayTmp[4], primMask;
    unsigned long long* statsHost = nullptr; // pinned
    uint32_t* ctrlHost = nullptr;            // pinned (volume queue polling)
    cudaEvent_t ev[4] = {};
    // three-kernel pipeline: the shadow rays of bounce b (any hit) are traced on a side stream while the main stream already traces
    // the extension rays of bounce b + 1 — two independent persistent kernels, so the tail of one overlaps the other
    cudaStream_t sideStream = nullptr;
    cudaEvent_t evShaded = nullptr, evConnected = nullptr;
    std::vector<cudaEvent_t> stageEvents; // pairs, with COUNTERS
    std::vector<int> stageKinds;
    // multi-GPU (multi.cu): replicas of this scene on further devices; empty for a single-device scene. Replica 0 is `this`.
    std::vector<xrtg_scene*> replicas;
    cudaEvent_t doneEvent = nullptr; // recorded on `stream` when this device's share of a multi-GPU render is complete
    cudaEvent_t uploadEvent = nullptr; // ... when this replica holds the scene arrays of the current xrtg_scene_upload (NVLink broadcast tree)
    cudaEvent_t pullEvent = nullptr; // ... when its slice of the fused reduce + finalize has been stored into device 0's image
    bool peerChecked = false, peerAll = false;
    xrt::DevBuf multiOut;            // device 0: the final image of a multi-GPU render
    xrt::PlainBuf exchange[4];         // exportable buffers (xrtg_exchange_buffer; one process per GPU + CUDA IPC)
    std::vector<std::unique_ptr<xrt::DevBuf>> peerStage; // device 0, topologies without peer mapping: staged copies of the other sums

    ~xrtg_scene();
};

namespace xrt {
// every scene array of a handle, in one fixed order (uploads, replicas, lazily made pinned copies)
constexpr int kSceneArrays = 16;
inline void sceneArrays(xrtg_scene* s, Mirror* out[kSceneArrays])
{
    Mirror* all[kSceneArrays] = {&s->nodes, &s->nodes4, &s->nodes8, &s->tris, &s->trisId, &s->ftris, &s->ftrisId, &s->ftris8,
                                 &s->smallBlock, &s->prims, &s->spheres, &s->boxes, &s->lights, &s->dlights, &s->media, &s->grids};
    for (int k = 0; k < kSceneArrays; ++k) out[k] = all[k];
}
// api.cu
int uploadAll(xrtg_scene* s, bool materialize = false);
int materializeHost(xrtg_scene* s); // pinned copies of the arrays a device-side build produced
int renderOnStream(xrtg_scene* s, const xrtg_camera* cam, const xrtg_render_params* p, float* out, cudaStream_t st, xrtg_stats* stats);
int checkParams(const xrtg_scene* s, const xrtg_camera* cam, const xrtg_render_params* p);
// multi.cu
int checkGuards(xrtg_scene* s, int* violations);
int createReplica(const xrtg_scene* primary, int device, xrtg_scene** out);
int renderMulti(xrtg_scene* s, const xrtg_camera* cam, const xrtg_render_params* p, float* rgbHost, xrtg_stats* stats);
int broadcastUpload(xrtg_scene* s, bool* done); // multi-device handles: host -> device 0 once, then a binomial tree of peer copies
void launchReduceFinalize(cudaStream_t st, const float* const* parts, int nParts, float* out, size_t n0, size_t n1, float divisor);
} // namespace xrt
