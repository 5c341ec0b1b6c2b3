// api.cu — C ABI of libxrtgpu.so (include/xrtgpu.h): scene ingest + SAH BVH build + HBM upload, the wave
// scheduler that drives the wavefront kernels, and the parity hooks. No CPU rendering path exists here: every
// compute entry point fails with XRTG_ERR_NO_DEVICE when no CUDA device is present.
#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include <xrtgpu.h>
#include "bvh.h"
#include "small_scene.h"
#include "device_types.h"
#include "kernels.h"
#include "lbvh.h"
#include "gpu_build.h"
#include "scene_impl.h"

using namespace xrt;

namespace {
thread_local std::string g_err;
float4 f4(const float* p, float w) { return make_float4(p[0], p[1], p[2], w); }
float asF(int v) { float f; std::memcpy(&f, &v, 4); return f; }
float asF(uint32_t v) { float f; std::memcpy(&f, &v, 4); return f; }
} // namespace

namespace xrt {
int fail(int code, const std::string& msg)
{
    g_err = msg;
    return code;
}
} // namespace xrt

xrtg_scene::~xrtg_scene()
{
    for (size_t k = 1; k < replicas.size(); ++k) { // replica 0 is this scene itself
        cudaSetDevice(replicas[k]->device);
        delete replicas[k];
    }
    cudaSetDevice(device);
    if (statsHost) cudaFreeHost(statsHost);
    if (ctrlHost) cudaFreeHost(ctrlHost);
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    for (auto& e : stageEvents) cudaEventDestroy(e);
    if (doneEvent) cudaEventDestroy(doneEvent);
    if (uploadEvent) cudaEventDestroy(uploadEvent);
    if (evShaded) cudaEventDestroy(evShaded);
    if (evConnected) cudaEventDestroy(evConnected);
    if (sideStream) cudaStreamDestroy(sideStream);
    if (pullEvent) cudaEventDestroy(pullEvent);
    if (stream) cudaStreamDestroy(stream);
}

namespace xrt {

int materializeHost(xrtg_scene* s)
{
    Mirror* all[kSceneArrays];
    sceneArrays(s, all);
    CU(cudaSetDevice(s->device));
    for (Mirror* m : all)
        if (int rc = m->ensureHost()) return rc;
    return 0;
}

int uploadAll(xrtg_scene* s, bool materialize)
{
    Mirror* all[kSceneArrays];
    sceneArrays(s, all);
    size_t total = 0;
    if (materialize)
        if (int rc = materializeHost(s)) return rc;
    for (Mirror* m : all) {
        // (arrays produced on the device have no host copy until someone asks for a re-upload)
        if (m->bytes && m->h) CU(cudaMemcpyAsync(m->d, m->h, m->bytes, cudaMemcpyHostToDevice, s->stream));
        total += m->bytes;
    }
    for (size_t i = 0; i < s->gridData.size(); ++i) {
        Mirror* g = s->gridData[i].get();
        CU(cudaMemcpyAsync(g->d, g->h, g->bytes, cudaMemcpyHostToDevice, s->stream));
        total += g->bytes;
        // the 3-D texture copy is refreshed from the array that just arrived (device to device: the voxels cross PCIe once)
        if (i < s->gridTex.size() && s->gridTex[i])
            if (int rc = s->gridTex[i]->fill(g->d, cudaMemcpyDeviceToDevice, s->stream)) return rc;
    }
    s->info.upload_bytes = total;
    return 0;
}

} // namespace xrt

namespace {

int checkDesc(const xrtg_scene_desc* d)
{
    if (!d) return fail(XRTG_ERR_INVALID, "scene desc is NULL");
    if (d->abi_version != XRTG_ABI_VERSION) return fail(XRTG_ERR_INVALID, "scene desc abi_version mismatch");
    if (d->n_objects < 0 || d->n_triangles < 0 || d->n_spheres < 0 || d->n_boxes < 0) return fail(XRTG_ERR_INVALID, "negative count in scene desc");
    if (d->n_area_lights > 4000 || d->n_media > 4000) return fail(XRTG_ERR_UNSUPPORTED, "too many lights/media");
    for (int i = 0; i < d->n_objects; ++i) {
        const xrtg_object& o = d->objects[i];
        const int lim = o.kind == XRTG_OBJ_MESH ? d->n_triangles : (o.kind == XRTG_OBJ_SPHERE ? d->n_spheres : d->n_boxes);
        if (o.kind < 0 || o.kind > 2) return fail(XRTG_ERR_INVALID, "unknown object kind");
        if (o.first < 0 || o.count < 0 || o.first + o.count > lim) return fail(XRTG_ERR_INVALID, "object geometry range out of bounds");
        if (o.material >= d->n_materials || o.area_light >= d->n_area_lights || o.medium >= d->n_media)
            return fail(XRTG_ERR_INVALID, "object references a missing material/light/medium");
        if (o.kind == XRTG_OBJ_BOX && o.medium < 0) return fail(XRTG_ERR_INVALID, "box object without a medium");
        if (o.kind != XRTG_OBJ_MESH && o.count != 1) return fail(XRTG_ERR_INVALID, "sphere/box object must have count 1");
    }
    for (int i = 0; i < d->n_media; ++i)
        if (d->media[i].kind == XRTG_MEDIUM_HETEROGENEOUS && (d->media[i].grid < 0 || d->media[i].grid >= d->n_grids))
            return fail(XRTG_ERR_INVALID, "heterogeneous medium without a grid");
    return 0;
}

// XRT_TUNING="key=value,key=value": development overrides, read ONCE per scene creation (the render path reads no environment)
struct TuningKey { const char* name; int32_t xrtg_tuning::*field; };
const TuningKey kTuningKeys[] = {
    {"fused_bounce", &xrtg_tuning::fused_bounce}, {"volume_paths", &xrtg_tuning::volume_paths}, {"scissor", &xrtg_tuning::scissor},
    {"brute_secondary", &xrtg_tuning::brute_secondary}, {"brute_shadow", &xrtg_tuning::brute_shadow}, {"thr_ext0", &xrtg_tuning::thr_ext0},
    {"thr_ext", &xrtg_tuning::thr_ext}, {"thr_con", &xrtg_tuning::thr_con}, {"steps_per_vote", &xrtg_tuning::steps_per_vote},
    {"leaf_threshold", &xrtg_tuning::leaf_threshold}, {"thr_vol", &xrtg_tuning::thr_vol}, {"spv_vol", &xrtg_tuning::spv_vol},
    {"wide_bvh", &xrtg_tuning::wide_bvh}, {"primary_masks", &xrtg_tuning::primary_masks}, {"max_leaf", &xrtg_tuning::max_leaf}, {"workspace_mb", &xrtg_tuning::workspace_mb},
    {"stage_dump", &xrtg_tuning::stage_dump}, {"gpu_build", &xrtg_tuning::gpu_build}, {"ploc_radius", &xrtg_tuning::ploc_radius},
    {"ploc_ct_x16", &xrtg_tuning::ploc_ct_x16}, {"ploc_top", &xrtg_tuning::ploc_top}, {"ploc_weight", &xrtg_tuning::ploc_weight}, {"grid_texture", &xrtg_tuning::grid_texture},
    {"overlap_connect", &xrtg_tuning::overlap_connect}};
constexpr int kGpuBuildMinTris = 1024;   // below this the device build is refused: the host path is faster than its launches and synchronisations
constexpr int kGpuBuildAutoTris = 65536; // from here on the device build is the default

void tuningFromEnvironment(Tuning& tu)
{
    const char* e = std::getenv("XRT_TUNING");
    if (!e) return;
    std::string str(e);
    size_t pos = 0;
    while (pos < str.size()) {
        size_t end = str.find(',', pos);
        if (end == std::string::npos) end = str.size();
        const std::string item = str.substr(pos, end - pos);
        const size_t eq = item.find('=');
        if (eq != std::string::npos) {
            const std::string key = item.substr(0, eq);
            const int val = std::atoi(item.c_str() + eq + 1);
            bool known = false;
            for (const TuningKey& k : kTuningKeys)
                if (key == k.name) { tu.t.*(k.field) = val; known = true; }
            if (!known) std::fprintf(stderr, "libxrtgpu: XRT_TUNING: unknown key '%s' ignored\n", key.c_str());
        }
        pos = end + 1;
    }
}
int tv(int32_t v, int dflt) { return v >= 0 ? int(v) : dflt; } // tuning value or the measured default

} // namespace

extern "C" {

int xrtg_abi_version(void) { return XRTG_ABI_VERSION; }

int xrtg_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* xrtg_last_error(void) { return g_err.c_str(); }

void xrtg_scene_destroy(xrtg_scene* s)
{
    if (!s) return;
    const int device = s->device;
    cudaSetDevice(device);
    delete s;
    // Releasing a multi-GB workspace is partly deferred by the driver: the NEXT cudaMalloc / cudaFree on the device then blocks for
    // hundreds of milliseconds (measured: a 1 MB cudaFree taking 716 ms right after a scene with a 12 GB workspace was destroyed).
    // Pay that here, where it belongs, instead of in whatever the caller creates next.
    cudaSetDevice(device);
    void* p = nullptr;
    if (cudaMalloc(&p, 256) == cudaSuccess) cudaFree(p);
    cudaGetLastError();
}

int xrtg_scene_create(const xrtg_scene_desc* d, int device, xrtg_scene** out) { return xrtg_scene_create2(d, device, 0u, out); }

int xrtg_scene_create2(const xrtg_scene_desc* d, int device, uint32_t build_flags, xrtg_scene** out)
{
    NvtxRange nvtx("xrtg_scene_create: ingest + BVH + upload");
    if (!out) return fail(XRTG_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (int rc = checkDesc(d)) return rc;
    if (xrtg_device_count() <= 0) return fail(XRTG_ERR_NO_DEVICE, "no CUDA device (libxrtgpu has no CPU fallback)");
    if (device < 0 || device >= xrtg_device_count()) return fail(XRTG_ERR_INVALID, "device index out of range");
    CU(cudaSetDevice(device));
    auto s = std::make_unique<xrtg_scene>();
    s->device = device;
    tuningFromEnvironment(s->tuning);
    CU(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    for (auto& e : s->ev) CU(cudaEventCreate(&e));
    CU(cudaMallocHost(&s->statsHost, sizeof(unsigned long long) * kStatCount));
    CU(cudaMallocHost(&s->ctrlHost, sizeof(uint32_t) * 16));

    Timer tb;
    const bool dumpCreate = tv(s->tuning.t.stage_dump, 0) != 0;
    Timer tphase;
    auto lap = [&](const char* what) {
        if (dumpCreate) std::fprintf(stderr, "scene_create: %-28s %8.2f ms\n", what, tphase.ms());
        tphase = Timer();
    };
    // ---- global primitive ids in object (= reference iteration) order ----
    int nPrims = 0;
    std::vector<int> firstPrim(d->n_objects);
    for (int i = 0; i < d->n_objects; ++i) {
        firstPrim[i] = nPrims;
        nPrims += d->objects[i].kind == XRTG_OBJ_MESH ? d->objects[i].count : 1;
    }
    int nMeshTris = 0, nSph = 0, nBox = 0;
    for (int i = 0; i < d->n_objects; ++i) {
        const xrtg_object& o = d->objects[i];
        if (o.kind == XRTG_OBJ_MESH) nMeshTris += o.count;
        else if (o.kind == XRTG_OBJ_SPHERE) nSph++;
        else nBox++;
    }
    // Device-side ingest + build (gpu_build.cu) for meshes large enough to pay for it: the raw triangles go up once and every
    // derived array is produced in HBM — no pinned staging copies (they are materialised lazily, on the first xrtg_scene_upload
    // or replica), no host loops over the triangles.
    // Default: meshes of kGpuBuildAutoTris triangles or more are built on the device (faster to build AND to traverse, measured
    // on the 1 M-triangle workload); XRTG_BUILD_GPU lowers the threshold to kGpuBuildMinTris, XRTG_BUILD_HOST / _LBVH_GPU or the
    // tuning switch gpu_build=0 keep the other builders.
    const bool otherBuilder = (build_flags & (XRTG_BUILD_HOST | XRTG_BUILD_LBVH_GPU)) != 0;
    const bool askedGpu = (build_flags & XRTG_BUILD_GPU) != 0 || s->tuning.t.gpu_build > 0;
    const bool gpuBuild = !otherBuilder && ((askedGpu && nMeshTris >= kGpuBuildMinTris) || (s->tuning.t.gpu_build != 0 && nMeshTris >= kGpuBuildAutoTris));
    auto allocArr = [&](Mirror& m, size_t bytes) { return gpuBuild ? m.allocDevice(bytes) : m.alloc(bytes); };
    if (int rc = allocArr(s->prims, sizeof(float4) * 4 * size_t(std::max(nPrims, 1)))) return rc;
    if (int rc = allocArr(s->trisId, sizeof(float4) * 3 * size_t(std::max(nMeshTris, 1)))) return rc;
    if (int rc = allocArr(s->tris, sizeof(float4) * 3 * size_t(std::max(nMeshTris, 1)))) return rc;
    if (int rc = allocArr(s->ftris, sizeof(float4) * 4 * size_t(std::max(nMeshTris, 1)))) return rc;
    if (int rc = allocArr(s->ftrisId, sizeof(float4) * 4 * size_t(std::max(nMeshTris, 1)))) return rc;
    if (int rc = s->spheres.alloc(sizeof(float4) * 2 * size_t(std::max(nSph, 1)))) return rc;
    if (int rc = s->boxes.alloc(sizeof(float4) * 2 * size_t(std::max(nBox, 1)))) return rc;
    lap(gpuBuild ? "device allocation" : "pinned + device allocation");
    if (!gpuBuild) std::memset(s->prims.h, 0, s->prims.bytes);
    float4* prims = static_cast<float4*>(s->prims.h); // (host-side arrays: NULL in the device-build path)
    float4* trisId = static_cast<float4*>(s->trisId.h);
    float4* ftrisId = static_cast<float4*>(s->ftrisId.h);
    float4* sph = static_cast<float4*>(s->spheres.h);
    float4* box = static_cast<float4*>(s->boxes.h);
    std::vector<float> buildTris(gpuBuild ? 0 : size_t(nMeshTris) * 9);
    std::vector<MeshRange> meshRanges;          // device-build path: one entry per mesh object
    std::vector<int> extraIds;                   // ... shading records of the spheres / boxes, scattered after the ingest kernel
    std::vector<float4> extraRecs;
    int ti = 0, si = 0, bi = 0;
    for (int i = 0; i < d->n_objects; ++i) {
        const xrtg_object& o = d->objects[i];
        uint32_t meta = uint32_t(o.kind);
        float alb[3] = {0, 0, 0};
        if (o.material >= 0) {
            meta |= kMetaHasMaterial;
            std::memcpy(alb, d->materials[o.material].albedo, 12);
        }
        meta |= uint32_t(o.area_light + 1) << kMetaLightShift;
        meta |= uint32_t(o.medium + 1) << kMetaMediumShift;
        const float4 rec3 = make_float4(alb[0], alb[1], alb[2], asF(meta));
        if (o.kind == XRTG_OBJ_MESH && gpuBuild) {
            if (o.count > 0) {
                MeshRange mr{};
                mr.srcFirst = o.first; mr.triStart = ti; mr.id0 = firstPrim[i]; mr.emitter = o.area_light >= 0 ? 1 : 0;
                std::memcpy(mr.albedo, alb, 12);
                mr.meta = meta;
                meshRanges.push_back(mr);
            }
            ti += o.count;
        }
        else if (o.kind == XRTG_OBJ_MESH) {
            const int ti0 = ti, id0 = firstPrim[i], emitter = o.area_light >= 0 ? 1 : 0;
            const xrtg_triangle* src = d->triangles + o.first;
            // (a 1 M-triangle mesh is one object: the per-triangle work runs on all host cores)
#pragma omp parallel for schedule(static) if (o.count > 4096)
            for (int k = 0; k < o.count; ++k) {
                const xrtg_triangle& t = src[k];
                const int id = id0 + k;
                const size_t tk = size_t(ti0) + size_t(k);
                // e1 = v1 - v0, e2 = v2 - v0 (primitive.cpp:142-143) and ng = normalize(e1 x e2)
                // (primitive.cpp:105) with the reference's fp32 operation order (host code, no FMA)
                float e1[3], e2[3], c[3];
                for (int a = 0; a < 3; ++a) { e1[a] = t.v1[a] - t.v0[a]; e2[a] = t.v2[a] - t.v0[a]; }
                c[0] = e1[1] * e2[2] - e1[2] * e2[1];
                c[1] = e1[2] * e2[0] - e1[0] * e2[2];
                c[2] = e1[0] * e2[1] - e1[1] * e2[0];
                const float len = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
                const float ng[3] = {c[0] / len, c[1] / len, c[2] / len};
                trisId[3 * tk] = f4(t.v0, asF(id));
                trisId[3 * tk + 1] = f4(e1, asF(emitter));
                trisId[3 * tk + 2] = f4(e2, 0.f);
                // plane-equation record of the throughput instantiation (wavefront.cuh: triangleRecord), built in double
                makePlaneRecord(t.v0, t.v1, t.v2, id, emitter, reinterpret_cast<float*>(ftrisId + 4 * tk));
                prims[4 * size_t(id)] = f4(t.n0, ng[0]);
                prims[4 * size_t(id) + 1] = f4(t.n1, ng[1]);
                prims[4 * size_t(id) + 2] = f4(t.n2, ng[2]);
                prims[4 * size_t(id) + 3] = rec3;
                std::memcpy(&buildTris[tk * 9], t.v0, 12);
                std::memcpy(&buildTris[tk * 9 + 3], t.v1, 12);
                std::memcpy(&buildTris[tk * 9 + 6], t.v2, 12);
            }
            ti += o.count;
        }
        else if (o.kind == XRTG_OBJ_SPHERE) {
            const xrtg_sphere& sp = d->spheres[o.first];
            const int id = firstPrim[i];
            sph[2 * si] = f4(sp.center, sp.radius);
            sph[2 * si + 1] = make_float4(asF(id), asF(int(o.area_light >= 0 ? 1 : 0)), 0.f, 0.f);
            if (gpuBuild) {
                extraIds.push_back(id);
                extraRecs.insert(extraRecs.end(), {f4(sp.center, sp.radius), make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f), rec3});
            }
            else {
                prims[4 * id] = f4(sp.center, sp.radius);
                prims[4 * id + 3] = rec3;
            }
            ++si;
        }
        else {
            const xrtg_box& b = d->boxes[o.first];
            const int id = firstPrim[i];
            box[2 * bi] = f4(b.pmin, asF(id));
            box[2 * bi + 1] = f4(b.pmax, 0.f);
            if (gpuBuild) {
                extraIds.push_back(id);
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                extraRecs.insert(extraRecs.end(), {z, z, z, rec3});
            }
            else prims[4 * id + 3] = rec3;
            ++bi;
        }
    }
    lap("triangle / shading records");
    // ---- device-side ingest, PLOC build and eight-child collapse (gpu_build.cu) ----
    Bvh bvh;
    bool builtOnGpu = false, wideOnGpu = false;
    GpuBuildInfo gbi;
    if (gpuBuild) {
        Timer tbv;
        const size_t nRaw = size_t(d->n_triangles);
        xrtg_triangle* dRaw = nullptr;
        MeshRange* dRanges = nullptr;
        int* dExtraIds = nullptr;
        float4* dExtraRecs = nullptr;
        struct Temp { void** p; ~Temp() { if (*p) cudaFree(*p); } };
        Temp t0{reinterpret_cast<void**>(&dRaw)}, t1{reinterpret_cast<void**>(&dRanges)}, t2{reinterpret_cast<void**>(&dExtraIds)}, t3{reinterpret_cast<void**>(&dExtraRecs)};
        CU(cudaMalloc(reinterpret_cast<void**>(&dRaw), sizeof(xrtg_triangle) * nRaw));
        CU(cudaMalloc(reinterpret_cast<void**>(&dRanges), sizeof(MeshRange) * meshRanges.size()));
        CU(cudaMemcpyAsync(dRaw, d->triangles, sizeof(xrtg_triangle) * nRaw, cudaMemcpyHostToDevice, s->stream));
        CU(cudaMemcpyAsync(dRanges, meshRanges.data(), sizeof(MeshRange) * meshRanges.size(), cudaMemcpyHostToDevice, s->stream));
        launchIngest(dRaw, dRanges, int(meshRanges.size()), nMeshTris, static_cast<float4*>(s->trisId.d), static_cast<float4*>(s->ftrisId.d),
                     static_cast<float4*>(s->prims.d), s->stream);
        if (!extraIds.empty()) {
            CU(cudaMalloc(reinterpret_cast<void**>(&dExtraIds), sizeof(int) * extraIds.size()));
            CU(cudaMalloc(reinterpret_cast<void**>(&dExtraRecs), sizeof(float4) * extraRecs.size()));
            CU(cudaMemcpyAsync(dExtraIds, extraIds.data(), sizeof(int) * extraIds.size(), cudaMemcpyHostToDevice, s->stream));
            CU(cudaMemcpyAsync(dExtraRecs, extraRecs.data(), sizeof(float4) * extraRecs.size(), cudaMemcpyHostToDevice, s->stream));
            launchScatterPrims(dExtraRecs, dExtraIds, int(extraIds.size()), static_cast<float4*>(s->prims.d), s->stream);
        }
        CU(cudaStreamSynchronize(s->stream));
        lap("raw upload + ingest kernel");
        if (int rc = s->nodes.allocDevice(sizeof(BvhNode) * size_t(nMeshTris - 1))) return rc;
        if (int rc = s->ftris8.allocDevice(sizeof(float4) * 4 * size_t(nMeshTris))) return rc;
        PlocParams pp;
        pp.radius = tv(s->tuning.t.ploc_radius, pp.radius);
        pp.maxLeaf = std::min(4, std::max(1, tv(s->tuning.t.max_leaf, 4)));
        if (s->tuning.t.ploc_ct_x16 >= 0) pp.traversalCost = float(s->tuning.t.ploc_ct_x16) / 16.f;
        pp.topClusters = tv(s->tuning.t.ploc_top, pp.topClusters);
        pp.topByClusters = tv(s->tuning.t.ploc_weight, 1) != 0;
        pp.verbose = dumpCreate;
        cudaError_t e = buildPlocDevice(static_cast<const float4*>(s->trisId.d), uint32_t(nMeshTris), static_cast<float4*>(s->tris.d),
                                        static_cast<BvhNode*>(s->nodes.d), s->stream, &gbi, static_cast<const float4*>(s->ftrisId.d),
                                        static_cast<float4*>(s->ftris.d), pp);
        lap("sort + PLOC + node records");
        Bvh8Node* d8 = nullptr;
        uint32_t n8 = 0;
        int depth8 = 0;
        if (e == cudaSuccess)
            e = collapseBvh8Device(static_cast<const BvhNode*>(s->nodes.d), uint32_t(nMeshTris), static_cast<const float4*>(s->ftris.d), &d8, &n8,
                                   static_cast<float4*>(s->ftris8.d), &depth8, s->stream);
        if (dumpCreate) std::fprintf(stderr, "scene_create: device build: %s, %u wide nodes, wide depth %d\n", cudaGetErrorString(e), n8, depth8);
        if (e == cudaErrorNotSupported || (e == cudaSuccess && depth8 + 2 > 64)) {
            // degenerate input (clustering without progress, or a tree deeper than the traversal stacks): the host builder decides
            if (d8) cudaFree(d8);
            cudaGetLastError();
            s.reset();
            return xrtg_scene_create2(d, device, (build_flags & ~uint32_t(XRTG_BUILD_GPU)) | XRTG_BUILD_HOST, out);
        }
        if (e != cudaSuccess) { if (d8) cudaFree(d8); return fail(XRTG_ERR_CUDA, std::string("GPU BVH build: ") + cudaGetErrorString(e)); }
        s->nodes8.adoptDevice(d8, sizeof(Bvh8Node) * size_t(n8));
        s->info.n_wide_nodes = int(n8);
        s->info.wide_arity = 8;
        bvh.depth = gbi.depth; bvh.pad = gbi.pad; bvh.sahCost = gbi.sahCost;
        bvh.nodes.resize(size_t(std::max(gbi.nNodes, 1))); // size bookkeeping only
        builtOnGpu = wideOnGpu = true;
        s->info.bvh_build_ms = tbv.ms();
        s->info.bvh_builder = 2;
        lap("eight-child collapse (device)");
    }
    // ---- world bounds of everything a ray can hit (screen-space scissor of the primary kernel) ----
    {
        float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        auto grow = [&](const float* p, float r) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], p[a] - r); hi[a] = std::max(hi[a], p[a] + r); } };
        for (size_t k = 0; k < buildTris.size(); k += 3) grow(&buildTris[k], 0.f);
        if (gpuBuild) { grow(gbi.lo, 0.f); grow(gbi.hi, 0.f); }
        for (int i = 0; i < d->n_objects; ++i) {
            const xrtg_object& o = d->objects[i];
            if (o.kind == XRTG_OBJ_SPHERE) grow(d->spheres[o.first].center, std::fabs(d->spheres[o.first].radius));
            else if (o.kind == XRTG_OBJ_BOX) { grow(d->boxes[o.first].pmin, 0.f); grow(d->boxes[o.first].pmax, 0.f); }
        }
        bool finite = nPrims > 0;
        for (int a = 0; a < 3; ++a) finite = finite && std::isfinite(lo[a]) && std::isfinite(hi[a]) && lo[a] <= hi[a];
        s->hasBounds = finite;
        if (finite) { std::memcpy(s->boundsLo, lo, 12); std::memcpy(s->boundsHi, hi, 12); }
    }
    if (nMeshTris >= 1 && nMeshTris <= 64) s->smallTriVerts = buildTris; // screen-space candidate masks of the primary rays
    // ---- BVH over all mesh triangles (emitter proxies included; any-hit skips them by flag) ----
    if (!builtOnGpu && (build_flags & XRTG_BUILD_LBVH_GPU) && nMeshTris >= 2) {
        // GPU build: triangles go up first, the tree and the leaf-ordered triangles are produced on the device and
        // mirrored back into the pinned host copies (xrtg_scene_upload re-sends them)
        if (int rc = s->nodes.alloc(sizeof(BvhNode) * size_t(nMeshTris - 1))) return rc;
        CU(cudaMemcpyAsync(s->trisId.d, s->trisId.h, s->trisId.bytes, cudaMemcpyHostToDevice, s->stream));
        CU(cudaMemcpyAsync(s->ftrisId.d, s->ftrisId.h, s->ftrisId.bytes, cudaMemcpyHostToDevice, s->stream));
        LbvhInfo li;
        CU(cudaStreamSynchronize(s->stream));
        Timer tbv;
        const cudaError_t e = buildLbvhDevice(static_cast<const float4*>(s->trisId.d), uint32_t(nMeshTris), static_cast<float4*>(s->tris.d),
                                              static_cast<BvhNode*>(s->nodes.d), s->stream, &li, static_cast<const float4*>(s->ftrisId.d),
                                              static_cast<float4*>(s->ftris.d));
        if (e != cudaSuccess) return fail(XRTG_ERR_CUDA, std::string("GPU BVH build: ") + cudaGetErrorString(e));
        if (li.depth <= 60) { // deeper than the traversal stacks allow (pathological duplicates): fall back to the host SAH build
            CU(cudaMemcpyAsync(s->nodes.h, s->nodes.d, s->nodes.bytes, cudaMemcpyDeviceToHost, s->stream));
            CU(cudaMemcpyAsync(s->tris.h, s->tris.d, s->tris.bytes, cudaMemcpyDeviceToHost, s->stream));
            CU(cudaMemcpyAsync(s->ftris.h, s->ftris.d, s->ftris.bytes, cudaMemcpyDeviceToHost, s->stream));
            CU(cudaStreamSynchronize(s->stream));
            bvh.depth = li.depth; bvh.pad = li.pad; bvh.sahCost = 0.f;
            bvh.nodes.resize(size_t(li.nNodes)); // size bookkeeping only
            builtOnGpu = true;
            s->info.bvh_build_ms = tbv.ms();
            s->info.bvh_builder = 1;
        }
    }
    if (!builtOnGpu) {
        Timer tbv;
        // at most 4 triangles per leaf (the leaf code of k_trace holds count-1 in 2 bits)
        const int maxLeaf = std::min(4, std::max(1, tv(s->tuning.t.max_leaf, 4)));
        buildBvh(buildTris.data(), uint32_t(nMeshTris), maxLeaf, bvh);
        // The traversal stacks hold 24 shared + 40 local = 64 entries and a two-child walk pushes at most one entry per level:
        // the builder stops SAH splits at depth 56 and then halves index ranges, so depth <= 56 + log2(n / maxLeaf) can exceed
        // that on pathological input (hundreds of thousands of coincident centroids). Refuse instead of overflowing silently.
        if (bvh.depth > 60) return fail(XRTG_ERR_UNSUPPORTED, "BVH depth " + std::to_string(bvh.depth) + " exceeds the traversal stack (60): degenerate triangle distribution");
        s->info.bvh_build_ms = tbv.ms();
        s->info.bvh_builder = 0;
        if (int rc = s->nodes.alloc(sizeof(BvhNode) * bvh.nodes.size())) return rc;
        std::memcpy(s->nodes.h, bvh.nodes.data(), s->nodes.bytes);
        float4* tris = static_cast<float4*>(s->tris.h);
        float4* ftris = static_cast<float4*>(s->ftris.h);
        const int64_t nOrder = int64_t(bvh.triOrder.size());
#pragma omp parallel for schedule(static) if (nOrder > 4096)
        for (int64_t k = 0; k < nOrder; ++k) {
            const uint32_t src = bvh.triOrder[size_t(k)];
            tris[3 * k] = trisId[3 * src];
            tris[3 * k + 1] = trisId[3 * src + 1];
            tris[3 * k + 2] = trisId[3 * src + 2];
            std::memcpy(ftris + 4 * k, ftrisId + 4 * size_t(src), 4 * sizeof(float4));
        }
    }
    lap("bounds + BVH build + reorder");
    // ---- deep trees: four-child form for the resumable traversal kernel (bvh.h) ----
    if (!wideOnGpu && bvh.nodes.size() > 512 && tv(s->tuning.t.wide_bvh, 4) >= 4) {
        std::vector<Bvh4Node> wide;
        const int depth4 = collapseBvh4(static_cast<const BvhNode*>(s->nodes.h), bvh.nodes.size(), wide);
        if (3 * depth4 + 1 <= 64) { // a four-child node pushes up to three entries: must fit the traversal stack (24 shared + 40 local)
            if (int rc = s->nodes4.alloc(sizeof(Bvh4Node) * wide.size())) return rc;
            std::memcpy(s->nodes4.h, wide.data(), s->nodes4.bytes);
            s->info.n_wide_nodes = int(wide.size());
            s->info.wide_arity = 4;
        }
    }
    lap("four-child collapse");
    // ---- deep trees, throughput instantiation: eight-child quantised nodes + node-ordered triangle records (bvh.h, k_trace8) ----
    if (!wideOnGpu && bvh.nodes.size() > 512) {
        std::vector<Bvh8Node> wide8;
        std::vector<uint32_t> order8;
        const int depth8 = collapseBvh8(static_cast<const BvhNode*>(s->nodes.h), bvh.nodes.size(), wide8, order8);
        if (depth8 + 2 <= 64 && order8.size() == size_t(nMeshTris)) { // one group-stack entry per level (12 shared + 52 local)
            if (int rc = s->nodes8.alloc(sizeof(Bvh8Node) * wide8.size())) return rc;
            std::memcpy(s->nodes8.h, wide8.data(), s->nodes8.bytes);
            if (int rc = s->ftris8.alloc(sizeof(float4) * 4 * order8.size())) return rc;
            const float4* ftris = static_cast<const float4*>(s->ftris.h);
            float4* f8 = static_cast<float4*>(s->ftris8.h);
            const int64_t n8 = int64_t(order8.size());
#pragma omp parallel for schedule(static) if (n8 > 4096)
            for (int64_t k = 0; k < n8; ++k) std::memcpy(f8 + 4 * k, ftris + 4 * size_t(order8[size_t(k)]), 4 * sizeof(float4));
            s->info.n_wide_nodes = int(wide8.size());
            s->info.wide_arity = 8;
        }
    }
    lap("eight-child collapse + ftris8");
    // ---- small scenes: plane-grouped triangle block for k_bounce_small (small_scene.h) ----
    int smallBlockF4 = 0;
    if (nMeshTris >= 1 && nMeshTris <= 64 && nBox == 0) {
        std::vector<float> block;
        SmallBlockInfo sbi;
        // Every point a shadow ray can start or end at (hull pruning of the occluder section, small_scene.h): triangle vertices,
        // the bounding corners of the spheres (grown by the largest shadow-ray bias), point lights. A DistantLight's shadow rays
        // leave the scene, so no pruning then.
        const float kMaxBias = 0.1f; // Whitted's 0.1 (integrator.h:337) covers Direct / GI's 0.01 (integrator.h:100, 260)
        std::vector<float> hull(buildTris);
        bool distant = false;
        for (int i = 0; i < d->n_objects; ++i)
            if (d->objects[i].kind == XRTG_OBJ_SPHERE) {
                const xrtg_sphere& sp = d->spheres[d->objects[i].first];
                const float r = std::fabs(sp.radius) + kMaxBias;
                for (int c = 0; c < 8; ++c)
                    for (int a = 0; a < 3; ++a) hull.push_back(sp.center[a] + ((c >> a) & 1 ? r : -r));
            }
        for (int i = 0; i < d->n_delta_lights; ++i) {
            if (d->delta_lights[i].kind == XRTG_DLIGHT_POINT) hull.insert(hull.end(), d->delta_lights[i].pos_or_dir, d->delta_lights[i].pos_or_dir + 3);
            else distant = true;
        }
        std::vector<int> pruned;
        if (buildSmallBlock(reinterpret_cast<const float*>(ftrisId), nMeshTris, block, &sbi, distant ? nullptr : hull.data(), int(hull.size() / 3), &pruned, buildTris.data())) {
            // Shadow rays do not start ON the surfaces: the origin is hit + bias * ng with ng never flipped towards the ray
            // (SURVEY §9-T3). On a triangle whose normal points OUT of the hull the origin lies behind a hull plane and the
            // reference lets that plane shadow it. An origin p + b * ng (b <= kMaxBias) is a convex combination of the
            // triangle's vertices and the vertices pushed out by kMaxBias, so it suffices to test those: a triangle with such a
            // vertex behind a pruned plane is flagged and its shadow rays test the unpruned occluder section (wf_shade.cuh).
            float ext = 0.f;
            for (float v : hull) ext = std::max(ext, std::fabs(v));
            const double tol = 1e-5 * std::max(double(ext), 1e-6); // the tolerance planeBoundsPoints() pruned with
            const float* recs = reinterpret_cast<const float*>(ftrisId);
            // a pruned plane has the scene on ONE side; which one is the sign of the summed vertex distances
            std::vector<double> side(pruned.size(), 0.0);
            for (size_t k = 0; k < pruned.size(); ++k)
                for (size_t h = 0; h + 2 < buildTris.size(); h += 3) side[k] += planeSignedDistance(recs + 16 * size_t(pruned[k]), &buildTris[h]);
            for (int t = 0; t < nMeshTris && !pruned.empty(); ++t) {
                int id;
                std::memcpy(&id, recs + 16 * size_t(t) + 12, 4);
                const float* v = &buildTris[size_t(t) * 9];
                const float4 n0 = prims[4 * id], n1 = prims[4 * id + 1], n2 = prims[4 * id + 2];
                const float ng[3] = {n0.w, n1.w, n2.w};
                bool outside = false;
                for (int k = 0; k < 3 && !outside; ++k) {
                    const float q[3] = {v[3 * k] + kMaxBias * ng[0], v[3 * k + 1] + kMaxBias * ng[1], v[3 * k + 2] + kMaxBias * ng[2]};
                    for (size_t pk = 0; pk < pruned.size(); ++pk) {
                        const double sd = planeSignedDistance(recs + 16 * size_t(pruned[pk]), q);
                        if ((side[pk] >= 0.0 && sd < -tol) || (side[pk] < 0.0 && sd > tol)) { outside = true; break; }
                    }
                }
                if (outside) {
                    uint32_t meta;
                    std::memcpy(&meta, &prims[4 * id + 3].w, 4);
                    meta |= kMetaShadowOutside;
                    std::memcpy(&prims[4 * id + 3].w, &meta, 4);
                    ++s->info.small_flagged;
                }
            }
            s->info.small_records_all = sbi.nRecordsAll;
            s->info.small_records_occ = sbi.nRecordsOcc;
            if (int rc = s->smallBlock.alloc(block.size() * sizeof(float))) return rc;
            std::memcpy(s->smallBlock.h, block.data(), s->smallBlock.bytes);
            smallBlockF4 = int(block.size() / 4);
        }
    }
    // ---- lights, media, grids ----
    if (int rc = s->lights.alloc(sizeof(DLight) * size_t(std::max(d->n_area_lights, 1)))) return rc;
    DLight* L = static_cast<DLight*>(s->lights.h);
    for (int i = 0; i < d->n_area_lights; ++i) {
        const xrtg_area_light& a = d->area_lights[i];
        float e1[3], e2[3], ng[3];
        for (int k = 0; k < 3; ++k) { e1[k] = a.v1[k] - a.v0[k]; e2[k] = a.v2[k] - a.v0[k]; }
        ng[0] = e1[1] * e2[2] - e1[2] * e2[1];
        ng[1] = e1[2] * e2[0] - e1[0] * e2[2];
        ng[2] = e1[0] * e2[1] - e1[1] * e2[0];
        L[i].v0_kind = f4(a.v0, asF(int(a.kind)));
        L[i].e1_r = f4(e1, a.radius);
        L[i].e2 = f4(e2, 0.f);
        L[i].Ng = f4(ng, 0.f);
        L[i].Le = f4(a.Le, 0.f);
        L[i].v1 = f4(a.v1, 0.f);
        L[i].v2 = f4(a.v2, 0.f);
    }
    if (int rc = s->dlights.alloc(sizeof(DDelta) * size_t(std::max(d->n_delta_lights, 1)))) return rc;
    DDelta* DL = static_cast<DDelta*>(s->dlights.h);
    for (int i = 0; i < d->n_delta_lights; ++i) {
        DL[i].p_kind = f4(d->delta_lights[i].pos_or_dir, asF(int(d->delta_lights[i].kind)));
        DL[i].L = f4(d->delta_lights[i].radiance, 0.f);
    }
    if (int rc = s->grids.alloc(sizeof(DGrid) * size_t(std::max(d->n_grids, 1)))) return rc;
    DGrid* G = static_cast<DGrid*>(s->grids.h);
    for (int i = 0; i < d->n_grids; ++i) {
        const xrtg_grid& g = d->grids[i];
        if (g.nx <= 0 || g.ny <= 0 || g.nz <= 0 || !g.data || !(g.voxel_size > 0)) return fail(XRTG_ERR_INVALID, "malformed density grid");
        auto m = std::make_unique<Mirror>();
        if (int rc = m->alloc(sizeof(float) * size_t(g.nx) * g.ny * g.nz)) return rc;
        std::memcpy(m->h, g.data, m->bytes);
        G[i].data = static_cast<const float*>(m->d);
        G[i].nx = g.nx; G[i].ny = g.ny; G[i].nz = g.nz;
        std::memcpy(G[i].origin, g.origin, 12);
        G[i].voxel = g.voxel_size;
        G[i].invVoxel = 1.0f / g.voxel_size;
        G[i].background = g.background;
        G[i].tex = 0ull;
        s->gridData.push_back(std::move(m));
        s->gridTex.emplace_back();
        if (g.background == 0.0f && tv(s->tuning.t.grid_texture, 1) != 0) {
            auto t = std::make_unique<GridTexture>();
            if (int rc = t->create(g.nx, g.ny, g.nz)) return rc;
            G[i].tex = (unsigned long long)t->tex;
            s->gridTex.back() = std::move(t);
        }
    }
    if (int rc = s->media.alloc(sizeof(DMedium) * size_t(std::max(d->n_media, 1)))) return rc;
    DMedium* M = static_cast<DMedium*>(s->media.h);
    for (int i = 0; i < d->n_media; ++i) {
        const xrtg_medium& m = d->media[i];
        M[i] = DMedium{};
        M[i].kind = m.kind; M[i].g = m.g; M[i].grid = m.grid; M[i].densityMul = m.density_mul;
        for (int k = 0; k < 3; ++k) { M[i].sigma_a[k] = m.sigma_a[k]; M[i].sigma_s[k] = m.sigma_s[k]; M[i].sigma_t[k] = m.sigma_a[k] + m.sigma_s[k]; }
        M[i].grey = (m.sigma_a[0] == m.sigma_a[1] && m.sigma_a[1] == m.sigma_a[2] && m.sigma_s[0] == m.sigma_s[1] && m.sigma_s[1] == m.sigma_s[2]) ? 1 : 0;
        if (m.kind == XRTG_MEDIUM_HETEROGENEOUS) {
            // HeterogeneousMedium ctor, medium.cpp:5-17
            const float maxd = m.density_mul * d->grids[m.grid].max_density;
            float mm[3];
            for (int k = 0; k < 3; ++k) mm[k] = m.sigma_a[k] * maxd + m.sigma_s[k] * maxd;
            M[i].majorant = std::max(mm[0], std::max(mm[1], mm[2]));
            M[i].invMajorant = 1.0f / M[i].majorant;
        }
    }
    lap("small block, lights, media");
    s->info.build_ms = tb.ms();

    Timer tu;
    if (int rc = uploadAll(s.get(), false)) return rc;
    CU(cudaStreamSynchronize(s->stream));
    s->info.upload_ms = tu.ms();

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
    ds.smallBlockF4 = smallBlockF4;
    ds.prims = static_cast<const float4*>(s->prims.d);
    ds.spheres = static_cast<const float4*>(s->spheres.d);
    ds.boxes = static_cast<const float4*>(s->boxes.d);
    ds.lights = static_cast<const DLight*>(s->lights.d);
    ds.dlights = static_cast<const DDelta*>(s->dlights.d);
    ds.media = static_cast<const DMedium*>(s->media.d);
    ds.grids = static_cast<const DGrid*>(s->grids.d);
    ds.nTris = nMeshTris; ds.nBruteTris = nMeshTris; ds.nSpheres = nSph; ds.nBoxes = nBox;
    ds.nLights = d->n_area_lights; ds.nDelta = d->n_delta_lights; ds.nPrims = nPrims;
    ds.nMedia = d->n_media; ds.nGrids = d->n_grids;
    s->maxShadowPerPath = std::max(1, std::max(d->n_area_lights, d->n_delta_lights));

    s->info.n_prims = nPrims;
    s->info.n_triangles = nMeshTris;
    s->info.n_bvh_nodes = int(bvh.nodes.size());
    s->info.bvh_depth = bvh.depth;
    s->info.bvh_sah_cost = bvh.sahCost;
    s->info.device_bytes = s->info.upload_bytes;
    s->info.n_devices = 1;
    *out = s.release();
    return 0;
}

int xrtg_scene_upload(xrtg_scene* s)
{
    if (!s) return fail(XRTG_ERR_INVALID, "scene is NULL");
    Timer t;
    const std::vector<xrtg_scene*> all = s->replicas.empty() ? std::vector<xrtg_scene*>{s} : s->replicas;
    // a device-built scene makes its pinned copies now (once); the replicas on other devices re-upload from the same copies
    if (int rc = materializeHost(s)) return rc;
    for (xrtg_scene* r : all) {
        if (r == s) continue;
        Mirror *src[kSceneArrays], *dst[kSceneArrays];
        sceneArrays(s, src);
        sceneArrays(r, dst);
        for (int k = 0; k < kSceneArrays; ++k)
            if (dst[k] != &r->grids) dst[k]->shareHost(*src[k]); // (every replica owns its table of grid descriptors: device pointers)
    }
    bool viaPeers = false;
    if (all.size() > 1)
        if (int rc = broadcastUpload(s, &viaPeers)) return rc;
    if (!viaPeers)
        for (xrtg_scene* r : all) { // every replica's copies are enqueued before the first one is waited for
            CU(cudaSetDevice(r->device));
            if (int rc = uploadAll(r, false)) return rc;
        }
    size_t total = 0;
    for (xrtg_scene* r : all) {
        CU(cudaSetDevice(r->device));
        CU(cudaStreamSynchronize(r->stream));
        total += r->info.upload_bytes;
    }
    CU(cudaSetDevice(s->device));
    s->info.upload_bytes = total; // all devices
    s->info.upload_ms = t.ms();
    return 0;
}

int xrtg_scene_get_info(const xrtg_scene* s, xrtg_scene_info* out)
{
    if (!s || !out) return fail(XRTG_ERR_INVALID, "NULL argument");
    *out = s->info;
    return 0;
}

} // extern "C"

namespace {

enum StageKind { kStageExtend = 0, kStageConnect, kStageShade, kStageOther };

struct StageTimer {
    xrtg_scene* s;
    cudaStream_t st;
    bool on;
    size_t used = 0;
    void begin(int kind)
    {
        if (!on) return;
        if (used + 2 > s->stageEvents.size()) {
            cudaEvent_t a, b;
            cudaEventCreate(&a); cudaEventCreate(&b);
            s->stageEvents.push_back(a); s->stageEvents.push_back(b);
        }
        if (used / 2 >= s->stageKinds.size()) s->stageKinds.push_back(kind);
        else s->stageKinds[used / 2] = kind;
        cudaEventRecord(s->stageEvents[used], st);
    }
    void end()
    {
        if (!on) return;
        cudaEventRecord(s->stageEvents[used + 1], st);
        used += 2;
    }
};

int ensureWorkspace(xrtg_scene* s, uint32_t nPixels, uint32_t maxPaths, size_t maxShadow, int maxIter, bool exact)
{
    const size_t f4b = sizeof(float4);
    // (+ the unused tails of k_bounce_small's warp-private output chunks: at most kAppendChunk - 1 slots per warp that had a tile)
    const size_t nQueue = size_t(maxPaths) + std::min<size_t>(kAppendSlack, (size_t(maxPaths) / 128 + 1) * 4 * kAppendChunk);
    for (int k = 0; k < 2; ++k) {
        if (int rc = s->q0[k].ensure(f4b * nQueue)) return rc;
        if (int rc = s->q1[k].ensure(f4b * nQueue)) return rc;
        if (int rc = s->q2[k].ensure(f4b * nQueue)) return rc;
    }
    if (int rc = s->hits.ensure(f4b * nQueue)) return rc;
    // shadow queue: one entry per (path, light) in the three-kernel pipeline; s0 doubles as the second hit buffer of the fused
    // bounce kernel, so it is never smaller than the ray queue
    const size_t nShadow = std::max<size_t>(maxShadow, nQueue);
    if (int rc = s->s0.ensure(f4b * nShadow)) return rc;
    if (int rc = s->s1.ensure(f4b * std::max<size_t>(maxShadow, 1))) return rc;
    if (int rc = s->s2.ensure(f4b * std::max<size_t>(maxShadow, 1))) return rc;
    if (int rc = s->radiance.ensure(f4b * maxPaths)) return rc;
    if (int rc = s->ctrl.ensure(sizeof(uint32_t) * kCtrlStride * size_t(maxIter + 2))) return rc;
    if (int rc = s->accum.ensure(sizeof(float) * 3 * size_t(nPixels))) return rc;
    if (int rc = s->stats.ensure(sizeof(unsigned long long) * kStatCount)) return rc;
    if (exact) {
        if (int rc = s->mt.ensure(sizeof(uint32_t) * 624 * size_t(nPixels))) return rc;
        if (int rc = s->mti.ensure(sizeof(uint32_t) * size_t(nPixels))) return rc;
    }
    return 0;
}

DQueues makeQueues(xrtg_scene* s)
{
    DQueues q{};
    for (int k = 0; k < 2; ++k) {
        q.q0[k] = static_cast<float4*>(s->q0[k].p);
        q.q1[k] = static_cast<float4*>(s->q1[k].p);
        q.q2[k] = static_cast<float4*>(s->q2[k].p);
    }
    q.hits = static_cast<float4*>(s->hits.p);
    q.s0 = static_cast<float4*>(s->s0.p);
    q.s1 = static_cast<float4*>(s->s1.p);
    q.s2 = static_cast<float4*>(s->s2.p);
    q.radiance = static_cast<float4*>(s->radiance.p);
    q.ctrl = static_cast<uint32_t*>(s->ctrl.p);
    q.stats = static_cast<unsigned long long*>(s->stats.p);
    return q;
}

DCamera makeCamera(const xrtg_camera* c)
{
    DCamera d;
    std::memcpy(d.c2w, c->c2w, sizeof(d.c2w));
    d.scale = c->scale;
    d.aspect = c->aspect;
    return d;
}

bool isVolume(int integ) { return integ == XRTG_INT_VOLUME || integ == XRTG_INT_VOLUME_NEE; }

} // namespace

namespace xrt {

int checkParams(const xrtg_scene* s, const xrtg_camera* cam, const xrtg_render_params* p)
{
    if (!s || !cam || !p) return fail(XRTG_ERR_INVALID, "NULL argument");
    if (p->width <= 0 || p->height <= 0 || p->spp <= 0) return fail(XRTG_ERR_INVALID, "width/height/spp must be positive");
    if (uint64_t(p->width) * uint64_t(p->height) > (1u << 28)) return fail(XRTG_ERR_UNSUPPORTED, "image too large");
    if (p->integrator < XRTG_INT_NORMAL || p->integrator > XRTG_INT_VOLUME_NEE) return fail(XRTG_ERR_UNSUPPORTED, "unknown integrator");
    if (p->max_depth < 0) return fail(XRTG_ERR_INVALID, "max_depth must be >= 0");
    if ((p->flags & XRTG_FLAG_EXACT) && p->sample_offset != 0)
        return fail(XRTG_ERR_UNSUPPORTED, "XRTG_FLAG_EXACT replays the per-pixel mt19937 stream from its seed: sample_offset must be 0");
    if (p->integrator == XRTG_INT_VOLUME_NEE && s->ds.nLights == 0) return fail(XRTG_ERR_INVALID, "VolumePathTracingNEE needs an area light");
    return 0;
}

} // namespace xrt

namespace {

// Conservative pixel rectangle that can see the scene's bounding box through PinholeCamera::sampleRay (camera.h:49-60):
// dir_cam = ((2u-1)s, (1-2v)s/aspect, -1), dir_world = dir_cam * R (row vector, geometry.h:653-669), origin = c2w row 3. Each box
// corner is taken back to camera space with R^-1 (double); if any corner is not strictly in front of the camera the scissor is
// the full image. One pixel of margin on every side.
void computeScissor(const xrtg_scene* s, const xrtg_camera* cam, int W, int H, int out[4])
{
    out[0] = 0; out[1] = 0; out[2] = W; out[3] = H;
    if (!s->hasBounds) return;
    const float* m = cam->c2w;
    const double R[3][3] = {{m[0], m[1], m[2]}, {m[4], m[5], m[6]}, {m[8], m[9], m[10]}};
    const double det = R[0][0] * (R[1][1] * R[2][2] - R[1][2] * R[2][1]) - R[0][1] * (R[1][0] * R[2][2] - R[1][2] * R[2][0]) +
                       R[0][2] * (R[1][0] * R[2][1] - R[1][1] * R[2][0]);
    if (!(std::fabs(det) > 1e-12) || !(cam->scale > 0.f) || !(cam->aspect > 0.f)) return;
    double inv[3][3];
    inv[0][0] = (R[1][1] * R[2][2] - R[1][2] * R[2][1]) / det; inv[0][1] = (R[0][2] * R[2][1] - R[0][1] * R[2][2]) / det;
    inv[0][2] = (R[0][1] * R[1][2] - R[0][2] * R[1][1]) / det; inv[1][0] = (R[1][2] * R[2][0] - R[1][0] * R[2][2]) / det;
    inv[1][1] = (R[0][0] * R[2][2] - R[0][2] * R[2][0]) / det; inv[1][2] = (R[0][2] * R[1][0] - R[0][0] * R[1][2]) / det;
    inv[2][0] = (R[1][0] * R[2][1] - R[1][1] * R[2][0]) / det; inv[2][1] = (R[0][1] * R[2][0] - R[0][0] * R[2][1]) / det;
    inv[2][2] = (R[0][0] * R[1][1] - R[0][1] * R[1][0]) / det;
    double span = 0.0;
    for (int a = 0; a < 3; ++a) span = std::max(span, double(s->boundsHi[a]) - s->boundsLo[a]);
    const double pad = 1e-4 * std::max(span, 1e-6);
    double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
    for (int c = 0; c < 8; ++c) {
        const double P[3] = {(c & 1 ? s->boundsHi[0] + pad : s->boundsLo[0] - pad) - m[12], (c & 2 ? s->boundsHi[1] + pad : s->boundsLo[1] - pad) - m[13],
                             (c & 4 ? s->boundsHi[2] + pad : s->boundsLo[2] - pad) - m[14]};
        // row vector * R^-1
        const double qx = P[0] * inv[0][0] + P[1] * inv[1][0] + P[2] * inv[2][0];
        const double qy = P[0] * inv[0][1] + P[1] * inv[1][1] + P[2] * inv[2][1];
        const double qz = P[0] * inv[0][2] + P[1] * inv[1][2] + P[2] * inv[2][2];
        const double len = std::sqrt(qx * qx + qy * qy + qz * qz);
        if (!(qz < -1e-6 * len)) return; // a corner beside or behind the camera: no scissor
        const double lam = -qz;
        const double u = 0.5 * (qx / (lam * cam->scale) + 1.0), v = 0.5 * (1.0 - qy * cam->aspect / (lam * cam->scale));
        x0 = std::min(x0, u * W); x1 = std::max(x1, u * W);
        y0 = std::min(y0, v * H); y1 = std::max(y1, v * H);
    }
    if (!(std::isfinite(x0) && std::isfinite(x1) && std::isfinite(y0) && std::isfinite(y1))) return;
    out[0] = int(std::max(0.0, std::floor(x0) - 1.0)); out[2] = int(std::min(double(W), std::ceil(x1) + 1.0));
    out[1] = int(std::max(0.0, std::floor(y0) - 1.0)); out[3] = int(std::min(double(H), std::ceil(y1) + 1.0));
    if (out[2] < out[0]) out[2] = out[0];
    if (out[3] < out[1]) out[3] = out[1];
}

// Pixel bounding boxes of the mesh triangles of a small scene as seen through PinholeCamera::sampleRay (camera.h:49-60), for the
// candidate masks of the primary kernel. Same inverse projection as computeScissor (double precision), one pixel of margin. A
// triangle with a vertex beside or behind the camera has no bounded projection: x0 > x1 marks "every pixel". Returns false if
// the camera matrix cannot be inverted (no masks then).
bool computeTriBoxes(const xrtg_scene* s, const xrtg_camera* cam, int W, int H, TriBoxes& out)
{
    const float* m = cam->c2w;
    const double R[3][3] = {{m[0], m[1], m[2]}, {m[4], m[5], m[6]}, {m[8], m[9], m[10]}};
    const double det = R[0][0] * (R[1][1] * R[2][2] - R[1][2] * R[2][1]) - R[0][1] * (R[1][0] * R[2][2] - R[1][2] * R[2][0]) +
                       R[0][2] * (R[1][0] * R[2][1] - R[1][1] * R[2][0]);
    if (!(std::fabs(det) > 1e-12) || !(cam->scale > 0.f) || !(cam->aspect > 0.f)) return false;
    double inv[3][3];
    inv[0][0] = (R[1][1] * R[2][2] - R[1][2] * R[2][1]) / det; inv[0][1] = (R[0][2] * R[2][1] - R[0][1] * R[2][2]) / det;
    inv[0][2] = (R[0][1] * R[1][2] - R[0][2] * R[1][1]) / det; inv[1][0] = (R[1][2] * R[2][0] - R[1][0] * R[2][2]) / det;
    inv[1][1] = (R[0][0] * R[2][2] - R[0][2] * R[2][0]) / det; inv[1][2] = (R[0][2] * R[1][0] - R[0][0] * R[1][2]) / det;
    inv[2][0] = (R[1][0] * R[2][1] - R[1][1] * R[2][0]) / det; inv[2][1] = (R[0][1] * R[2][0] - R[0][0] * R[2][1]) / det;
    inv[2][2] = (R[0][0] * R[1][1] - R[0][1] * R[1][0]) / det;
    const int n = int(s->smallTriVerts.size() / 9);
    for (int t = 0; t < n; ++t) {
        double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
        bool bounded = true;
        for (int v = 0; v < 3 && bounded; ++v) {
            const float* p = &s->smallTriVerts[size_t(t) * 9 + 3 * v];
            const double P[3] = {double(p[0]) - m[12], double(p[1]) - m[13], double(p[2]) - m[14]};
            const double qx = P[0] * inv[0][0] + P[1] * inv[1][0] + P[2] * inv[2][0]; // row vector * R^-1
            const double qy = P[0] * inv[0][1] + P[1] * inv[1][1] + P[2] * inv[2][1];
            const double qz = P[0] * inv[0][2] + P[1] * inv[1][2] + P[2] * inv[2][2];
            const double len = std::sqrt(qx * qx + qy * qy + qz * qz);
            if (!(qz < -1e-4 * len)) { bounded = false; break; } // beside / behind the camera (or NaN)
            const double lam = -qz;
            const double u = 0.5 * (qx / (lam * cam->scale) + 1.0), v2 = 0.5 * (1.0 - qy * cam->aspect / (lam * cam->scale));
            x0 = std::min(x0, u * W); x1 = std::max(x1, u * W);
            y0 = std::min(y0, v2 * H); y1 = std::max(y1, v2 * H);
        }
        bounded = bounded && std::isfinite(x0) && std::isfinite(x1) && std::isfinite(y0) && std::isfinite(y1);
        if (!bounded) { out.b[t] = make_int4(1, 0, 0, 0); continue; }
        auto clampi = [](double v, int lo, int hi) { return int(std::min(double(hi), std::max(double(lo), v))); };
        // pixel j holds the samples u in [j / W, (j + 1) / W): floor / ceil, then one more pixel on every side
        out.b[t] = make_int4(clampi(std::floor(x0) - 1.0, 0, W), clampi(std::floor(y0) - 1.0, 0, H), clampi(std::ceil(x1) + 1.0, 0, W), clampi(std::ceil(y1) + 1.0, 0, H));
        if (out.b[t].x > out.b[t].z) out.b[t].z = out.b[t].x; // (off-screen: empty box, never "every pixel")
    }
    return true;
}

// Builds the candidate masks of the primary rays on `st` and returns the device pointer (nullptr = no masks for this scene / camera).
const unsigned long long* preparePrimaryMasks(xrtg_scene* s, const KernelTable& K, const xrtg_camera* cam, int W, int H, cudaStream_t st)
{
    const int n = int(s->smallTriVerts.size() / 9);
    if (n < 1 || n > 64 || n != s->ds.nBruteTris) return nullptr;
    TriBoxes tb;
    if (!computeTriBoxes(s, cam, W, H, tb)) return nullptr;
    const uint32_t nPixels = uint32_t(W) * uint32_t(H);
    if (s->primMask.ensure(sizeof(unsigned long long) * ((size_t(nPixels) + 31) / 32)) != 0) return nullptr;
    K.primaryMasks(st, tb, n, W, nPixels, static_cast<unsigned long long*>(s->primMask.p));
    return static_cast<const unsigned long long*>(s->primMask.p);
}

// Which kernels render this scene. Chosen per (scene, integrator); the parity hooks with XRTG_FLAG_FAST_HOOK ask the same
// function, so they exercise exactly the entry points a render would.
//   deep    : > 512 BVH nodes -> raygen + k_trace (refillable state machine over the wide tree) + shade + k_trace<any>
//   small   : <= 64 triangles -> incoherent rays test every triangle from shared memory (SmallTracer)
//   fusedBounce : small scene + surface integrator -> k_primary, then ONE k_bounce_small per bounce
//   volumePaths : shallow BVH + volume integrator -> k_primary, then every path to completion in k_volume_paths
// Measured defaults (profiles/r01_notes.md); every switch can be overridden through xrtg_scene_set_tuning.
struct Pipeline {
    bool deep, small, fusedPrimary, fusedBounce, volumePaths, bruteSecondary, bruteShadow, scissor;
    int thrExt0, thrExt, thrCon, spv, leafThr, thrVol, spvVol;
};
Pipeline choosePipeline(const xrtg_scene* s, int integ, int nIter, bool exact, bool brute)
{
    const xrtg_tuning& t = s->tuning.t;
    Pipeline P{};
    P.deep = s->info.n_bvh_nodes > 512;
    // refill thresholds / steps per vote of the resumable traversal kernels, measured on the 1 M-triangle scene
    // (profiles/r02_notes.md): eight-child tree 24 / 24 / 2, four-child tree 16 / 16 / 4
    const bool eight = !exact && s->ds.nodes8 != nullptr && tv(t.wide_bvh, 8) >= 8;
    P.thrExt0 = tv(t.thr_ext0, P.deep ? 1 : 0);
    P.thrExt = tv(t.thr_ext, P.deep ? (eight ? 24 : 16) : 0);
    P.thrCon = tv(t.thr_con, P.deep ? (eight ? 24 : 16) : 0);
    if (s->info.bvh_depth > 60) { // only k_trace's stack holds a tree this deep: the run-to-completion kernels cannot be forced onto it
        P.thrExt0 = std::max(P.thrExt0, 1); P.thrExt = std::max(P.thrExt, 1); P.thrCon = std::max(P.thrCon, 1);
    }
    P.spv = tv(t.steps_per_vote, P.deep ? (eight ? 2 : 4) : 1);
    P.leafThr = tv(t.leaf_threshold, 4); // lanes that must stand at a leaf before the warp runs the triangle tests (k_trace)
    P.thrVol = tv(t.thr_vol, 16);
    // (bits 8.. of spv_vol = refill + prologue rounds per outer iteration of k_volume_paths; 1 = measured best: a second round hands the
    //  lanes whose path ended in volumePre a fresh path before the walk, but runs refill + prologue at ~20 % of the lanes — c5 12.95 -> 12.16)
    P.spvVol = tv(t.spv_vol, 3);
    if ((P.spvVol >> 8) == 0) P.spvVol |= 1 << 8;
    P.small = !P.deep && s->ds.nBruteTris > 0 && s->ds.nBruteTris <= 64;
    P.bruteSecondary = tv(t.brute_secondary, P.small ? 1 : 0) != 0;
    P.bruteShadow = tv(t.brute_shadow, P.small ? 1 : 0) != 0;
    // throughput instantiation only: the exact one must advance every pixel's mt19937 stream by its two jitter draws
    P.scissor = !exact && tv(t.scissor, 1) != 0;
    // Shallow BVHs: ray generation is fused with the primary closest hit (k_primary, compact hit-only queue). Deep BVHs: separate
    // raygen + the refillable traversal kernel, which is faster there even on primary rays.
    P.fusedPrimary = !P.deep && nIter > 0;
    const bool volume = integ == XRTG_INT_VOLUME || integ == XRTG_INT_VOLUME_NEE;
    P.volumePaths = P.fusedPrimary && volume && s->ds.nMedia <= 8 && s->ds.nGrids <= 8 && tv(t.volume_paths, 1) != 0;
    P.fusedBounce = P.fusedPrimary && P.small && !brute && P.bruteSecondary && P.bruteShadow && tv(t.fused_bounce, 1) != 0 &&
                    s->ds.nPrims <= 96 && s->ds.nLights <= 16 && // what k_bounce_small stages in shared memory (kSmallPrims / kSmallLights)
                    (integ == XRTG_INT_DIRECT || integ == XRTG_INT_WHITTED || integ == XRTG_INT_INDIRECT || integ == XRTG_INT_GI);
    return P;
}

// Which form of a deep tree the traversal kernel walks: 8 (default; throughput instantiation only), 4 or 2 children per node
DScene sceneForTraversal(const xrtg_scene* s)
{
    DScene ds = s->ds;
    const int arity = tv(s->tuning.t.wide_bvh, 8);
    if (arity < 8) { ds.nodes8 = nullptr; ds.ftris8 = nullptr; }
    if (arity < 4) ds.nodes4 = nullptr;
    return ds;
}

} // namespace

namespace xrt {

// The whole render on `st`, result (mean or sum) written to device buffer `out`.
int renderOnStream(xrtg_scene* s, const xrtg_camera* cam, const xrtg_render_params* p, float* out, cudaStream_t st, xrtg_stats* stats)
{
    NvtxRange nvtx("xrtg_render: wave scheduler");
    const bool exact = (p->flags & XRTG_FLAG_EXACT) != 0;
    const bool count = (p->flags & XRTG_FLAG_COUNTERS) != 0;
    const bool brute = (p->flags & XRTG_FLAG_BRUTE_FORCE) != 0;
    const KernelTable& K = exact ? exactKernels() : fastKernels();
    const uint32_t nPixels = uint32_t(p->width) * uint32_t(p->height);
    const int integ = p->integrator;
    const bool volume = isVolume(integ);

    int nIter;
    if (integ == XRTG_INT_INDIRECT || integ == XRTG_INT_GI) nIter = p->max_depth;
    else if (volume) nIter = std::min(4 * p->max_depth + 8, 4096); // medium crossings per path; cut paths are counted (truncated_paths)
    else nIter = 1;
    const bool hasShadow = (integ == XRTG_INT_DIRECT || integ == XRTG_INT_GI || integ == XRTG_INT_WHITTED);
    const Pipeline P = choosePipeline(s, integ, nIter, exact, brute);

    // ---- wave size from a BYTE budget ----
    // A wave is `S` samples of a range of `tile` pixels. Workspace per path: 2 x 48 B ping-pong ray queue + 16 B hit + 16 B
    // radiance, plus — in the three-kernel pipeline only — 48 B per shadow-queue entry x one entry per area (or delta) light:
    // an emissive mesh of a few hundred triangle lights needs tens of KB per path, so the wave shrinks (down to a range of
    // pixels at one sample each) instead of the allocation failing. Default budget 16 GiB, 64 M paths at most (32 samples of every
    // pixel at 1080p). BIG waves pay: every launch of a persistent queue kernel ends in a tail where the refill thresholds can no
    // longer be met, and 180 GB of HBM make the queues cheap — measured against waves of 4 samples (the round-1 default, 8 M paths):
    // c4 1.99 -> 2.54, c5 13.4 -> 16.0, c3 9.28 -> 9.69 Gsamples/s (profiles/r02_notes.md).
    const size_t shadowPerPath = (hasShadow && !P.fusedBounce) ? size_t(s->maxShadowPerPath) : 0;
    const size_t bytesPerPath = 128 + 48 * shadowPerPath;
    const size_t budget = size_t(tv(s->tuning.t.workspace_mb, 16384)) << 20;
    uint64_t maxPaths = std::min<uint64_t>(64u << 20, std::max<uint64_t>(budget / bytesPerPath, 1024));
    uint32_t S = 1, tile = nPixels; // exact: one sample per wave (the mt19937 stream of a pixel is sequential across its samples)
    if (!exact) {
        if (p->samples_per_wave > 0) S = std::min<uint32_t>(uint32_t(p->samples_per_wave), uint32_t(p->spp));
        else S = uint32_t(std::max<uint64_t>(1, std::min<uint64_t>(maxPaths / nPixels, uint64_t(p->spp))));
    }
    if (S == 1 && maxPaths < nPixels) tile = uint32_t(maxPaths); // not even one sample of every pixel fits: waves over pixel ranges
    while (uint64_t(S) * tile > (1ull << 31) - 64) --S;

    if (int rc = ensureWorkspace(s, nPixels, S * tile, size_t(S) * tile * shadowPerPath, nIter, exact)) return rc;
    const DScene ds = sceneForTraversal(s);
    DQueues q = makeQueues(s);
    const DCamera dc = makeCamera(cam);
    float* accum = static_cast<float*>(s->accum.p);
    unsigned long long* dstats = static_cast<unsigned long long*>(s->stats.p);

    DWave w{};
    w.width = p->width; w.height = p->height; w.nPixels = nPixels;
    w.byWidth = makeFastDiv(uint32_t(p->width));
    w.integrator = integ; w.maxDepth = p->max_depth; w.seed = p->seed;
    w.flags = 0;
    w.sx0 = 0; w.sy0 = 0; w.sx1 = p->width; w.sy1 = p->height;
    w.mt = static_cast<uint32_t*>(s->mt.p);
    w.mti = static_cast<uint32_t*>(s->mti.p);

    StageTimer tm{s, st, count || (p->flags & XRTG_FLAG_STAGE_TIMES) != 0};
    const int missMode = integ == XRTG_INT_DIRECT ? 1 : (integ == XRTG_INT_WHITTED ? 2 : 0);
    // Three-kernel pipeline with shadow rays: any hit (bounce b) and closest hit (bounce b + 1) are independent — run them on two
    // streams. Not with per-stage timers (their events assume one stream) and not for the fused / volume pipelines (no such pair).
    const bool overlapConnect = hasShadow && !volume && !P.fusedBounce && !tm.on && nIter > 1 && tv(s->tuning.t.overlap_connect, 1) != 0;
    bool connectPending = false;
    if (overlapConnect) {
        if (!s->sideStream) CU(cudaStreamCreateWithFlags(&s->sideStream, cudaStreamNonBlocking));
        if (!s->evShaded) CU(cudaEventCreateWithFlags(&s->evShaded, cudaEventDisableTiming));
        if (!s->evConnected) CU(cudaEventCreateWithFlags(&s->evConnected, cudaEventDisableTiming));
    }
    const bool dump = tv(s->tuning.t.stage_dump, 0) != 0;
    if (P.scissor) {
        int sc4[4];
        computeScissor(s, cam, p->width, p->height, sc4);
        w.sx0 = sc4[0]; w.sy0 = sc4[1]; w.sx1 = sc4[2]; w.sy1 = sc4[3];
        if (P.fusedPrimary) w.flags |= kWaveScissorSkip; // k_primary is the only writer of the initial radiance in these pipelines
    }
    uint64_t launches = 0, nExtend = 0, nShade = 0, nConnect = 0, nBounce = 0, truncatedHost = 0;
    CU(cudaEventRecord(s->ev[0], st));
    CU(cudaMemsetAsync(accum, 0, sizeof(float) * 3 * size_t(nPixels), st));
    CU(cudaMemsetAsync(dstats, 0, sizeof(unsigned long long) * kStatCount, st));
    if (exact) { K.seedMt(st, w); ++launches; }
    if (P.fusedPrimary && tv(s->tuning.t.primary_masks, 1) != 0) { // small scenes: screen-space candidate masks of the primary rays
        w.primMask = preparePrimaryMasks(s, K, cam, p->width, p->height, st);
        if (w.primMask) ++launches;
    }

    for (uint32_t pix0 = 0; pix0 < nPixels; pix0 += tile)
    for (uint32_t done = 0; done < uint32_t(p->spp); done += S) {
        NvtxRange nvtxWave("wave: primary / bounces / accumulate");
        const uint32_t sw = std::min<uint32_t>(S, uint32_t(p->spp) - done);
        w.pixelBase = pix0;
        w.wavePixels = std::min<uint32_t>(tile, nPixels - pix0);
        w.byWavePixels = makeFastDiv(w.wavePixels);
        w.samplesThisWave = sw;
        w.nPaths = sw * w.wavePixels;
        w.sampleBase = uint32_t(p->sample_offset) + done;
        tm.begin(kStageOther);
        CU(cudaMemsetAsync(q.ctrl, 0, sizeof(uint32_t) * kCtrlStride * size_t(nIter + 2), st));
        // No bounce at all (maxDepth 0): only the per-path radiance has to be cleared.
        if (nIter == 0) CU(cudaMemsetAsync(q.radiance, 0, sizeof(float4) * size_t(w.nPaths), st));
        else if (!P.fusedPrimary) { K.raygen(st, dc, q, w, nullptr); ++launches; }
        tm.end();
        for (int b = 0; b < nIter; ++b) {
            const int src = b & 1;
            if (P.volumePaths) { // volume integrators on a shallow BVH: primary, then every path to completion in one launch
                tm.begin(kStageExtend);
                K.primary(st, ds, dc, q, w, brute, missMode, count, dstats, nullptr); ++launches; ++nExtend;
                tm.end();
                tm.begin(kStageShade);
                K.volumePaths(st, ds, q, w, brute, nIter, P.thrVol, P.spvVol, count, dstats); ++launches; ++nShade;
                tm.end();
                break;
            }
            if (P.fusedBounce) { // small scene: primary, then one fused shade + connect + extend kernel per bounce
                if (b == 0) {
                    tm.begin(kStageExtend);
                    K.primary(st, ds, dc, q, w, brute, missMode, count, dstats, nullptr); ++launches; ++nExtend;
                    tm.end();
                }
                tm.begin(kStageShade);
                K.bounceSmall(st, ds, q, w, src, b, dstats); ++launches; ++nShade; ++nBounce;
                tm.end();
                continue;
            }
            tm.begin(kStageExtend);
            if (b == 0 && P.fusedPrimary) K.primary(st, ds, dc, q, w, brute, missMode, count, dstats, nullptr);
            else K.extend(st, ds, q, src, b, brute ? 1 : ((P.bruteSecondary && b > 0) ? 2 : 0), count, dstats, b == 0 ? P.thrExt0 : P.thrExt, P.spv, P.leafThr);
            ++launches; ++nExtend;
            tm.end();
            // (the shade kernel rewrites the shadow queue: the any-hit pass of the previous bounce must be through with it)
            if (overlapConnect && connectPending) { CU(cudaStreamWaitEvent(st, s->evConnected, 0)); connectPending = false; }
            tm.begin(kStageShade);
            if (volume) K.shadeVolume(st, ds, q, w, src, b, brute, count, dstats);
            else K.shadeSurface(st, ds, q, w, src, b);
            ++launches; ++nShade;
            tm.end();
            if (hasShadow) {
                if (overlapConnect) {
                    // any hit of bounce b on the side stream, concurrently with the closest hit of bounce b + 1 on the main one
                    CU(cudaEventRecord(s->evShaded, st));
                    CU(cudaStreamWaitEvent(s->sideStream, s->evShaded, 0));
                    K.connect(s->sideStream, ds, q, b, brute ? 1 : (P.bruteShadow ? 2 : 0), count, dstats, P.thrCon, P.spv, P.leafThr); ++launches; ++nConnect;
                    CU(cudaEventRecord(s->evConnected, s->sideStream));
                    connectPending = true;
                }
                else {
                    tm.begin(kStageConnect);
                    K.connect(st, ds, q, b, brute ? 1 : (P.bruteShadow ? 2 : 0), count, dstats, P.thrCon, P.spv, P.leafThr); ++launches; ++nConnect;
                    tm.end();
                }
            }
            if (volume) {
                // the number of loop iterations of integrator.h:418 is data dependent: poll the next queue size
                CU(cudaMemcpyAsync(s->ctrlHost, q.ctrl + (b + 1) * kCtrlStride + kCtrlRays, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
                CU(cudaStreamSynchronize(st));
                if (s->ctrlHost[0] == 0) break;
                if (b + 1 == nIter) truncatedHost += s->ctrlHost[0]; // paths still alive at the iteration bound
            }
        }
        if (overlapConnect && connectPending) { CU(cudaStreamWaitEvent(st, s->evConnected, 0)); connectPending = false; } // radiance complete
        tm.begin(kStageOther);
        K.accumulate(st, q, w, accum, dstats); ++launches;
        tm.end();
    }
    const int divisor = (p->flags & XRTG_FLAG_SUM_ONLY) ? 0 : (p->spp_total > 0 ? p->spp_total : p->spp);
    K.finalize(st, accum, out, size_t(nPixels) * 3, float(divisor)); ++launches;
    CU(cudaEventRecord(s->ev[1], st));
    CU(cudaGetLastError());

    if (stats) {
        CU(cudaMemcpyAsync(s->statsHost, dstats, sizeof(unsigned long long) * kStatCount, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        std::memset(stats, 0, sizeof(*stats));
        stats->samples = uint64_t(nPixels) * uint64_t(p->spp);
        stats->closest_rays = s->statsHost[kStatClosest];
        stats->shadow_rays = s->statsHost[kStatShadow];
        stats->dropped_samples = s->statsHost[kStatDropped];
        stats->nodes_visited = s->statsHost[kStatNodes];
        stats->tris_tested = s->statsHost[kStatTris];
        stats->nodes_visited_shadow = s->statsHost[kStatNodesAny];
        stats->tris_tested_shadow = s->statsHost[kStatTrisAny];
        stats->tracking_steps = s->statsHost[kStatSteps];
        stats->primary_hits = s->statsHost[kStatPrimaryHits];
        stats->bounce_entries = s->statsHost[kStatBounceEntries];
        stats->rays_traced = stats->closest_rays + stats->shadow_rays - s->statsHost[kStatScissored];
        stats->untraced_closest = s->statsHost[kStatUntracedClosest];
        stats->untraced_shadow = s->statsHost[kStatUntracedShadow];
        stats->truncated_paths = s->statsHost[kStatTruncated] + truncatedHost;
        stats->n_devices = 1;
        stats->bounce_launches = nBounce;
        stats->kernel_launches = launches;
        stats->extend_launches = nExtend; stats->shade_launches = nShade; stats->connect_launches = nConnect;
        CU(cudaEventElapsedTime(&stats->render_ms, s->ev[0], s->ev[1]));
        if (tm.on) {
            float acc[4] = {0, 0, 0, 0};
            for (size_t k = 0; k + 1 < tm.used; k += 2) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, s->stageEvents[k], s->stageEvents[k + 1]);
                acc[s->stageKinds[k / 2]] += ms;
                if (dump) std::fprintf(stderr, "stage %zu kind %d %.3f ms\n", k / 2, s->stageKinds[k / 2], ms);
            }
            stats->extend_ms = acc[kStageExtend]; stats->connect_ms = acc[kStageConnect];
            stats->shade_ms = acc[kStageShade]; stats->other_ms = acc[kStageOther];
        }
    }
    return 0;
}

} // namespace xrt

__global__ void k_image_to_u8(const float* __restrict__ rgb, unsigned char* __restrict__ out, size_t nPixels, float invGamma, int bgr)
{
    for (size_t p = size_t(blockIdx.x) * blockDim.x + threadIdx.x; p < nPixels; p += size_t(gridDim.x) * blockDim.x) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v = rgb[3 * p + c];
            if (invGamma > 0.f) v = powf(v, invGamma);               // image.h:85-87
            const float s = 255.0f * v;                              // image.h:103-108 / 123-128
            const unsigned int q = s >= 255.0f ? 255u : (s > 0.f ? (unsigned int)s : 0u);
            out[3 * p + (bgr ? 2 - c : c)] = (unsigned char)q;
        }
    }
}

extern "C" {

int xrtg_small_scene_selftest(const float* tris9, const int* emitter_flags, int n, int* n_records_all, int* n_records_occ, int* n_planes)
{
    if (n < 0 || (n > 0 && !tris9)) return fail(XRTG_ERR_INVALID, "bad triangle array");
    SmallBlockInfo bi;
    const int rc = smallBlockSelftest(tris9, emitter_flags, n, &bi);
    if (n_records_all) *n_records_all = bi.nRecordsAll;
    if (n_records_occ) *n_records_occ = bi.nRecordsOcc;
    if (n_planes) *n_planes = bi.nPlanesAll;
    if (rc < 0) return fail(XRTG_ERR_INVALID, "small-scene block is inconsistent with its triangles");
    return rc;
}

// Structural check of the trees RESIDENT ON THE DEVICE (whichever builder made them): the arrays are copied back and walked on the
// host. Two-child tree: every leaf-ordered triangle in exactly one leaf, every child box contains the triangles below it. Eight-child
// tree: the same with the DECODED quantised boxes shrunk by the one step of margin the traversal kernel's folded FMA relies on
// (wf_trace8.cuh: byteMagic), every primitive id exactly once in the node-ordered records, leaf children of at most four triangles.
int xrtg_scene_selfcheck(xrtg_scene* s, int* n_errors)
{
    if (!s || !n_errors) return fail(XRTG_ERR_INVALID, "NULL argument");
    *n_errors = 0;
    const int n = s->info.n_triangles;
    if (n < 2 || s->nodes.bytes == 0) return 0;
    if (int rc = materializeHost(s)) return rc;
    const float4* trisId = static_cast<const float4*>(s->trisId.h);
    const float4* tris = static_cast<const float4*>(s->tris.h);
    const BvhNode* nodes = static_cast<const BvhNode*>(s->nodes.h);
    const size_t nNodes = s->nodes.bytes / sizeof(BvhNode);
    int errors = 0;
    std::string firstMsg;
    auto err = [&](const std::string& m) { if (errors++ == 0) firstMsg = m; };
    struct Bounds { float lo[3], hi[3]; };
    auto triBounds = [](const float4* rec, Bounds& b) { // v0, v0 + e1, v0 + e2
        const float v[3][3] = {{rec[0].x, rec[0].y, rec[0].z}, {rec[0].x + rec[1].x, rec[0].y + rec[1].y, rec[0].z + rec[1].z}, {rec[0].x + rec[2].x, rec[0].y + rec[2].y, rec[0].z + rec[2].z}};
        for (int a = 0; a < 3; ++a) { b.lo[a] = std::min(v[0][a], std::min(v[1][a], v[2][a])); b.hi[a] = std::max(v[0][a], std::max(v[1][a], v[2][a])); }
    };
    auto grow = [](Bounds& a, const Bounds& b) { for (int k = 0; k < 3; ++k) { a.lo[k] = std::min(a.lo[k], b.lo[k]); a.hi[k] = std::max(a.hi[k], b.hi[k]); } };
    const Bounds kEmpty{{FLT_MAX, FLT_MAX, FLT_MAX}, {-FLT_MAX, -FLT_MAX, -FLT_MAX}};
    // ---- two-child tree ----
    {
        std::vector<int> seen(size_t(n), 0), seenId(size_t(s->info.n_prims), 0);
        int maxDepth = 0;
        std::function<Bounds(int, int)> walk = [&](int node, int depth) -> Bounds {
            Bounds all = kEmpty;
            maxDepth = std::max(maxDepth, depth + 1);
            if (node < 0 || size_t(node) >= nNodes || depth > 130) { err("two-child tree: node index out of range or runaway depth"); return all; }
            const BvhNode& b = nodes[node];
            for (int c = 0; c < 2; ++c) {
                const float* lo = c ? b.lo1 : b.lo0;
                const float* hi = c ? b.hi1 : b.hi0;
                const int child = c ? b.child1 : b.child0, count = c ? b.count1 : b.count0;
                if (count < 0) continue;
                Bounds sub = kEmpty;
                if (count > 0) {
                    if (count > 4 || child < 0 || child + count > n) { err("two-child tree: bad leaf range"); continue; }
                    for (int k = 0; k < count; ++k) {
                        Bounds tb;
                        triBounds(tris + 3 * size_t(child + k), tb);
                        grow(sub, tb);
                        seen[size_t(child + k)]++;
                        int id;
                        std::memcpy(&id, &tris[3 * size_t(child + k)].w, 4);
                        if (id < 0 || id >= s->info.n_prims) err("two-child tree: primitive id out of range");
                        else seenId[size_t(id)]++;
                    }
                }
                else sub = walk(child, depth + 1);
                for (int a = 0; a < 3; ++a)
                    if (!(lo[a] <= sub.lo[a] && sub.hi[a] <= hi[a])) { err("two-child tree: child box does not contain its subtree (node " + std::to_string(node) + ")"); break; }
                grow(all, sub);
            }
            return all;
        };
        walk(0, 0);
        for (int k = 0; k < n; ++k)
            if (seen[size_t(k)] != 1) { err("two-child tree: leaf-ordered triangle " + std::to_string(k) + " referenced " + std::to_string(seen[size_t(k)]) + " times"); break; }
        for (int k = 0; k < n; ++k) {
            int id;
            std::memcpy(&id, &trisId[3 * size_t(k)].w, 4);
            if (id < 0 || id >= s->info.n_prims || seenId[size_t(id)] != 1) { err("two-child tree: primitive " + std::to_string(id) + " not exactly once in the leaf order"); break; }
        }
        if (maxDepth > 124) err("two-child tree deeper than the traversal stack of k_trace");
    }
    // ---- eight-child tree ----
    if (s->nodes8.bytes && s->ftris8.bytes) {
        const Bvh8Node* n8 = static_cast<const Bvh8Node*>(s->nodes8.h);
        const size_t nWide = s->nodes8.bytes / sizeof(Bvh8Node);
        const float4* f8 = static_cast<const float4*>(s->ftris8.h);
        std::vector<int> tkOfId(size_t(s->info.n_prims), -1), seenId(size_t(s->info.n_prims), 0), seenNode(nWide, 0);
        for (int k = 0; k < n; ++k) {
            int id;
            std::memcpy(&id, &trisId[3 * size_t(k)].w, 4);
            if (id >= 0 && id < s->info.n_prims) tkOfId[size_t(id)] = k;
        }
        int maxDepth = 0;
        std::function<Bounds(uint32_t, int)> walk = [&](uint32_t node, int depth) -> Bounds {
            Bounds all = kEmpty;
            maxDepth = std::max(maxDepth, depth + 1);
            if (size_t(node) >= nWide || depth > 70) { err("eight-child tree: node index out of range or runaway depth"); return all; }
            if (seenNode[node]++) { err("eight-child tree: node referenced twice"); return all; }
            const Bvh8Node& w = n8[node];
            uint32_t nextChild = 0, nextTri = 0;
            for (int s8 = 0; s8 < 8; ++s8) {
                const bool inner = (w.imask >> s8) & 1u;
                const uint32_t nib = (w.validTri >> (4 * s8)) & 0xFu;
                if (inner && nib) { err("eight-child tree: slot is both inner and leaf"); continue; }
                if (!inner && !nib) continue;
                Bounds sub = kEmpty;
                if (inner) sub = walk(w.childBase + nextChild++, depth + 1);
                else {
                    if (nib != 1u && nib != 3u && nib != 7u && nib != 15u) { err("eight-child tree: malformed triangle nibble"); continue; }
                    const int count = __builtin_popcount(nib);
                    for (int k = 0; k < count; ++k) {
                        const size_t t = size_t(w.triBase) + nextTri++;
                        if (t >= size_t(n)) { err("eight-child tree: triangle index out of range"); continue; }
                        int id;
                        std::memcpy(&id, &f8[4 * t + 3].x, 4);
                        if (id < 0 || id >= s->info.n_prims || tkOfId[size_t(id)] < 0) { err("eight-child tree: bad primitive id in ftris8"); continue; }
                        seenId[size_t(id)]++;
                        Bounds tb;
                        triBounds(trisId + 3 * size_t(tkOfId[size_t(id)]), tb);
                        grow(sub, tb);
                    }
                }
                for (int a = 0; a < 3; ++a) {
                    const double step = std::ldexp(1.0, int(w.e[a]) - 127);
                    const double lo = double(w.p[a]) + (double(w.qlo[a][s8]) + 1.0) * step, hi = double(w.p[a]) + (double(w.qhi[a][s8]) - 1.0) * step;
                    if (!(lo <= double(sub.lo[a]) && double(sub.hi[a]) <= hi)) { err("eight-child tree: quantised child box (minus its margin) does not contain its subtree (node " + std::to_string(node) + ", slot " + std::to_string(s8) + ")"); break; }
                }
                grow(all, sub);
            }
            return all;
        };
        walk(0u, 0);
        for (int k = 0; k < n; ++k) {
            int id;
            std::memcpy(&id, &trisId[3 * size_t(k)].w, 4);
            if (id < 0 || id >= s->info.n_prims || seenId[size_t(id)] != 1) { err("eight-child tree: primitive " + std::to_string(id) + " not exactly once in ftris8"); break; }
        }
        for (size_t k = 0; k < nWide; ++k)
            if (seenNode[k] != 1) { err("eight-child tree: node " + std::to_string(k) + " unreachable"); break; }
        if (maxDepth + 2 > 64) err("eight-child tree deeper than the group stack");
    }
    *n_errors = errors;
    if (errors) fail(XRTG_ERR_INVALID, "selfcheck: " + firstMsg);
    return 0;
}

int xrtg_top_sah_selftest(const float* lo3, const float* hi3, const uint32_t* counts, int n, int by_clusters, int* depth)
{
    if (n < 1 || !lo3 || !hi3) return fail(XRTG_ERR_INVALID, "bad selftest arguments");
    const int d = topSahSelftest(lo3, hi3, counts, n, by_clusters);
    if (depth) *depth = d;
    if (d < 0) return fail(XRTG_ERR_INVALID, "selftest: inconsistent top-level tree");
    return 0;
}

int xrtg_bvh_selftest(const float* tri, int n, int max_leaf, int* n_nodes, int* depth, float* sah_cost)
{
    if (n < 0 || (n > 0 && !tri) || max_leaf < 1 || max_leaf > 4) return fail(XRTG_ERR_INVALID, "bad selftest arguments");
    Bvh bvh;
    buildBvh(tri, uint32_t(n), max_leaf, bvh);
    if (n_nodes) *n_nodes = int(bvh.nodes.size());
    if (depth) *depth = bvh.depth;
    if (sah_cost) *sah_cost = bvh.sahCost;
    if (n == 0) return 0;
    if (bvh.triOrder.size() != size_t(n)) return fail(XRTG_ERR_INVALID, "selftest: triOrder has the wrong size");
    if (bvh.depth > 60) return fail(XRTG_ERR_INVALID, "selftest: tree deeper than the traversal stacks");
    std::vector<int> seen(size_t(n), 0);
    struct Range { float lo[3], hi[3]; };
    // recursive descent: returns the exact bounds of the subtree and checks them against the (padded) box stored in the parent
    std::string err;
    std::vector<std::pair<int, int>> stack; // (node, slot) handled iteratively through an explicit post-order
    std::function<bool(int, int, Range&)> visit = [&](int child, int count, Range& out) -> bool {
        for (int a = 0; a < 3; ++a) { out.lo[a] = FLT_MAX; out.hi[a] = -FLT_MAX; }
        if (count > 0) { // leaf
            if (count > max_leaf) { err = "selftest: leaf larger than max_leaf"; return false; }
            for (int k = 0; k < count; ++k) {
                const uint32_t t = bvh.triOrder[size_t(child) + k];
                if (t >= uint32_t(n)) { err = "selftest: triangle index out of range"; return false; }
                seen[t]++;
                for (int v = 0; v < 3; ++v)
                    for (int a = 0; a < 3; ++a) {
                        out.lo[a] = std::min(out.lo[a], tri[9 * size_t(t) + 3 * v + a]);
                        out.hi[a] = std::max(out.hi[a], tri[9 * size_t(t) + 3 * v + a]);
                    }
            }
            return true;
        }
        if (child < 0 || size_t(child) >= bvh.nodes.size()) { err = "selftest: node index out of range"; return false; }
        const BvhNode& nd = bvh.nodes[size_t(child)];
        const int cs[2] = {nd.child0, nd.child1}, ks[2] = {nd.count0, nd.count1};
        const float* los[2] = {nd.lo0, nd.lo1};
        const float* his[2] = {nd.hi0, nd.hi1};
        for (int s2 = 0; s2 < 2; ++s2) {
            if (ks[s2] < 0) continue; // empty slot
            Range r;
            if (!visit(cs[s2], ks[s2], r)) return false;
            for (int a = 0; a < 3; ++a) {
                if (!(los[s2][a] <= r.lo[a] - 0.5f * bvh.pad) || !(his[s2][a] >= r.hi[a] + 0.5f * bvh.pad)) {
                    err = "selftest: child box does not contain its subtree with the conservative padding";
                    return false;
                }
                out.lo[a] = std::min(out.lo[a], r.lo[a]);
                out.hi[a] = std::max(out.hi[a], r.hi[a]);
            }
        }
        return true;
    };
    Range all;
    if (!visit(0, 0, all)) return fail(XRTG_ERR_INVALID, err);
    for (int t = 0; t < n; ++t)
        if (seen[size_t(t)] != 1) return fail(XRTG_ERR_INVALID, "selftest: a triangle is referenced " + std::to_string(seen[size_t(t)]) + " times");
    // ---- the four-child form of the same tree: same leaves, every child box contains its subtree ----
    std::vector<Bvh4Node> wide;
    const int depth4 = collapseBvh4(bvh.nodes.data(), bvh.nodes.size(), wide);
    if (depth4 > bvh.depth) return fail(XRTG_ERR_INVALID, "selftest: the collapsed tree is deeper than the binary one");
    std::fill(seen.begin(), seen.end(), 0);
    std::function<bool(int, int, Range&)> visit4 = [&](int child, int count, Range& out) -> bool {
        for (int a = 0; a < 3; ++a) { out.lo[a] = FLT_MAX; out.hi[a] = -FLT_MAX; }
        if (count > 0) {
            if (count > max_leaf) { err = "selftest: wide leaf larger than max_leaf"; return false; }
            for (int k = 0; k < count; ++k) {
                const uint32_t t = bvh.triOrder[size_t(child) + k];
                seen[t]++;
                for (int v = 0; v < 3; ++v)
                    for (int a = 0; a < 3; ++a) {
                        out.lo[a] = std::min(out.lo[a], tri[9 * size_t(t) + 3 * v + a]);
                        out.hi[a] = std::max(out.hi[a], tri[9 * size_t(t) + 3 * v + a]);
                    }
            }
            return true;
        }
        if (child < 0 || size_t(child) >= wide.size()) { err = "selftest: wide node index out of range"; return false; }
        const Bvh4Node& nd = wide[size_t(child)];
        int used = 0;
        for (int k = 0; k < 4; ++k) {
            if (nd.count[k] < 0) continue;
            ++used;
            Range r;
            if (!visit4(nd.child[k], nd.count[k], r)) return false;
            const float lo[3] = {nd.lox[k], nd.loy[k], nd.loz[k]}, hi[3] = {nd.hix[k], nd.hiy[k], nd.hiz[k]};
            for (int a = 0; a < 3; ++a) {
                if (!(lo[a] <= r.lo[a] - 0.5f * bvh.pad) || !(hi[a] >= r.hi[a] + 0.5f * bvh.pad)) {
                    err = "selftest: wide child box does not contain its subtree with the conservative padding";
                    return false;
                }
                out.lo[a] = std::min(out.lo[a], r.lo[a]);
                out.hi[a] = std::max(out.hi[a], r.hi[a]);
            }
        }
        if (used == 0) { err = "selftest: wide node without children"; return false; }
        return true;
    };
    if (!visit4(0, 0, all)) return fail(XRTG_ERR_INVALID, err);
    for (int t = 0; t < n; ++t)
        if (seen[size_t(t)] != 1) return fail(XRTG_ERR_INVALID, "selftest: the collapsed tree references a triangle " + std::to_string(seen[size_t(t)]) + " times");
    // ---- the eight-child quantised form: same triangles exactly once, every DEQUANTISED child box (minus its one step of margin
    //      per side, which the kernel's folded slab arithmetic may eat) contains the child's subtree ----
    std::vector<Bvh8Node> w8;
    std::vector<uint32_t> order8;
    const int depth8 = collapseBvh8(bvh.nodes.data(), bvh.nodes.size(), w8, order8);
    if (depth8 > bvh.depth || order8.size() != size_t(n)) return fail(XRTG_ERR_INVALID, "selftest: eight-child tree has the wrong depth / triangle count");
    std::fill(seen.begin(), seen.end(), 0);
    std::function<bool(uint32_t, Range&)> visit8 = [&](uint32_t ni, Range& out) -> bool {
        for (int a = 0; a < 3; ++a) { out.lo[a] = FLT_MAX; out.hi[a] = -FLT_MAX; }
        if (ni >= w8.size()) { err = "selftest: eight-child node index out of range"; return false; }
        const Bvh8Node& nd = w8[ni];
        uint32_t triNext = nd.triBase, childNext = nd.childBase;
        for (int k = 0; k < 8; ++k) {
            const uint32_t nib = (nd.validTri >> (4 * k)) & 0xfu;
            const bool inner = (nd.imask >> k) & 1u;
            if (inner && nib) { err = "selftest: slot is both inner and leaf"; return false; }
            if (!inner && !nib) continue; // empty
            Range r;
            if (inner) { if (!visit8(childNext++, r)) return false; }
            else {
                for (int a = 0; a < 3; ++a) { r.lo[a] = FLT_MAX; r.hi[a] = -FLT_MAX; }
                if (nib != 1u && nib != 3u && nib != 7u && nib != 15u) { err = "selftest: malformed triangle nibble"; return false; }
                for (uint32_t b = nib; b; b >>= 1) {
                    if (triNext >= order8.size()) { err = "selftest: eight-child triangle index out of range"; return false; }
                    const uint32_t t = bvh.triOrder[order8[triNext++]];
                    seen[t]++;
                    for (int v = 0; v < 3; ++v)
                        for (int a = 0; a < 3; ++a) { r.lo[a] = std::min(r.lo[a], tri[9 * size_t(t) + 3 * v + a]); r.hi[a] = std::max(r.hi[a], tri[9 * size_t(t) + 3 * v + a]); }
                }
            }
            for (int a = 0; a < 3; ++a) {
                const double sc = std::ldexp(1.0, int(nd.e[a]) - 127);
                const double qlo = double(nd.p[a]) + (double(nd.qlo[a][k]) + 1.0) * sc, qhi = double(nd.p[a]) + (double(nd.qhi[a][k]) - 1.0) * sc;
                if (!(qlo <= double(r.lo[a]) - 0.5 * bvh.pad) || !(qhi >= double(r.hi[a]) + 0.5 * bvh.pad)) { err = "selftest: quantised child box (without its margin) does not contain its subtree"; return false; }
                out.lo[a] = std::min(out.lo[a], r.lo[a]);
                out.hi[a] = std::max(out.hi[a], r.hi[a]);
            }
        }
        return true;
    };
    if (!visit8(0, all)) return fail(XRTG_ERR_INVALID, err);
    for (int t = 0; t < n; ++t)
        if (seen[size_t(t)] != 1) return fail(XRTG_ERR_INVALID, "selftest: the eight-child tree references a triangle " + std::to_string(seen[size_t(t)]) + " times");
    return 0;
}

int xrtg_image_to_u8(int device, const float* rgb_host, int width, int height, float gamma, int bgr, uint8_t* out_host)
{
    if (!rgb_host || !out_host || width <= 0 || height <= 0) return fail(XRTG_ERR_INVALID, "bad image arguments");
    if (xrtg_device_count() <= 0) return fail(XRTG_ERR_NO_DEVICE, "no CUDA device (libxrtgpu has no CPU fallback)");
    CU(cudaSetDevice(device));
    const size_t n = size_t(width) * height;
    float* d_in = nullptr;
    unsigned char* d_out = nullptr;
    CU(cudaMalloc(&d_in, n * 3 * sizeof(float)));
    if (cudaMalloc(&d_out, n * 3) != cudaSuccess) { cudaFree(d_in); return fail(XRTG_ERR_OOM, "cudaMalloc failed"); }
    cudaError_t e = cudaMemcpy(d_in, rgb_host, n * 3 * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        k_image_to_u8<<<int(std::min<size_t>((n + 255) / 256, 148 * 8)), 256>>>(d_in, d_out, n, gamma > 0.f ? 1.0f / gamma : 0.f, bgr);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out_host, d_out, n * 3, cudaMemcpyDeviceToHost);
    cudaFree(d_in);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail(XRTG_ERR_CUDA, std::string("xrtg_image_to_u8: ") + cudaGetErrorString(e));
    return 0;
}

int xrtg_scene_set_tuning(xrtg_scene* s, const xrtg_tuning* t)
{
    if (!s || !t) return fail(XRTG_ERR_INVALID, "NULL argument");
    for (xrtg_scene* r : s->replicas.empty() ? std::vector<xrtg_scene*>{s} : s->replicas) r->tuning.t = *t;
    return 0;
}

int xrtg_scene_device_count(const xrtg_scene* s) { return s ? std::max<int>(1, int(s->replicas.size())) : 0; }

int xrtg_render_device(xrtg_scene* s, const xrtg_camera* cam, const xrtg_render_params* p, float* rgb_device, void* cuda_stream,
                       xrtg_stats* stats)
{
    if (int rc = checkParams(s, cam, p)) return rc;
    if (!rgb_device) return fail(XRTG_ERR_INVALID, "rgb_device is NULL");
    if (s->replicas.size() > 1) return fail(XRTG_ERR_UNSUPPORTED, "xrtg_render_device on a multi-GPU scene: use xrtg_render (host buffer)");
    CU(cudaSetDevice(s->device));
    return renderOnStream(s, cam, p, rgb_device, static_cast<cudaStream_t>(cuda_stream), stats);
}

int xrtg_render(xrtg_scene* s, const xrtg_camera* cam, const xrtg_render_params* p, float* rgb_host, xrtg_stats* stats)
{
    if (int rc = checkParams(s, cam, p)) return rc;
    if (!rgb_host) return fail(XRTG_ERR_INVALID, "rgb_host is NULL");
    if (s->replicas.size() > 1 && !(p->flags & XRTG_FLAG_EXACT) && p->spp >= 2) return renderMulti(s, cam, p, rgb_host, stats);
    CU(cudaSetDevice(s->device));
    const size_t bytes = sizeof(float) * 3 * size_t(p->width) * size_t(p->height);
    if (int rc = s->outDev.ensure(bytes)) return rc;
    xrtg_stats local{};
    if (int rc = renderOnStream(s, cam, p, static_cast<float*>(s->outDev.p), s->stream, stats ? &local : nullptr)) return rc;
    CU(cudaEventRecord(s->ev[2], s->stream));
    CU(cudaMemcpyAsync(rgb_host, s->outDev.p, bytes, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaEventRecord(s->ev[3], s->stream));
    CU(cudaStreamSynchronize(s->stream));
    if (stats) {
        *stats = local;
        CU(cudaEventElapsedTime(&stats->d2h_ms, s->ev[2], s->ev[3]));
    }
    return 0;
}

int xrtg_render_u8(xrtg_scene* s, const xrtg_camera* cam, const xrtg_render_params* p, float gamma, int bgr, uint8_t* out_host, xrtg_stats* stats)
{
    if (int rc = checkParams(s, cam, p)) return rc;
    if (!out_host) return fail(XRTG_ERR_INVALID, "out_host is NULL");
    if (p->flags & XRTG_FLAG_SUM_ONLY) return fail(XRTG_ERR_INVALID, "xrtg_render_u8 quantises the MEAN image: XRTG_FLAG_SUM_ONLY makes no sense here");
    if (s->replicas.size() > 1) return fail(XRTG_ERR_UNSUPPORTED, "xrtg_render_u8 on a multi-GPU scene: render with xrtg_render, then xrtg_image_to_u8");
    CU(cudaSetDevice(s->device));
    const size_t n = size_t(p->width) * size_t(p->height);
    if (int rc = s->outDev.ensure(sizeof(float) * 3 * n)) return rc;
    if (int rc = s->rayTmp[3].ensure(3 * n)) return rc;
    xrtg_stats local{};
    if (int rc = renderOnStream(s, cam, p, static_cast<float*>(s->outDev.p), s->stream, stats ? &local : nullptr)) return rc;
    // the float image never leaves HBM: gamma + quantisation read it where k_finalize wrote it, 3 bytes per pixel go to the host
    k_image_to_u8<<<int(std::min<size_t>((n + 255) / 256, 148 * 8)), 256, 0, s->stream>>>(static_cast<const float*>(s->outDev.p), static_cast<unsigned char*>(s->rayTmp[3].p), n,
                                                                                          gamma > 0.f ? 1.0f / gamma : 0.f, bgr);
    CU(cudaGetLastError());
    CU(cudaEventRecord(s->ev[2], s->stream));
    CU(cudaMemcpyAsync(out_host, s->rayTmp[3].p, 3 * n, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaEventRecord(s->ev[3], s->stream));
    CU(cudaStreamSynchronize(s->stream));
    if (stats) {
        *stats = local;
        CU(cudaEventElapsedTime(&stats->d2h_ms, s->ev[2], s->ev[3]));
    }
    return 0;
}

// The production pipeline a parity hook with XRTG_FLAG_FAST_HOOK has to exercise: what choosePipeline() picks for a GIIntegrator
// render of this scene. 0 = k_trace, 1 = simple kernels, 2 = small-scene tracer (launchTraceRays).
static int hookMode(const xrtg_scene* s, uint32_t flags)
{
    if (!(flags & XRTG_FLAG_FAST_HOOK) || (flags & XRTG_FLAG_BRUTE_FORCE)) return 0;
    const Pipeline P = choosePipeline(s, XRTG_INT_GI, 3, false, false);
    if (P.fusedBounce) return 2;
    return P.deep ? 0 : 1;
}

int xrtg_trace_primary(xrtg_scene* s, const xrtg_camera* cam, int width, int height, int spp, const float* jitter_uv, uint32_t flags,
                       xrtg_hit* out)
{
    if (!s || !cam || !out) return fail(XRTG_ERR_INVALID, "NULL argument");
    if (width <= 0 || height <= 0 || spp <= 0) return fail(XRTG_ERR_INVALID, "width/height/spp must be positive");
    CU(cudaSetDevice(s->device));
    const uint32_t nPixels = uint32_t(width) * uint32_t(height);
    const uint64_t nPaths = uint64_t(nPixels) * uint64_t(spp);
    if (nPaths > (1ull << 30)) return fail(XRTG_ERR_UNSUPPORTED, "too many primary rays for one call");
    const bool fastHook = (flags & XRTG_FLAG_FAST_HOOK) != 0;
    const bool brute = (flags & XRTG_FLAG_BRUTE_FORCE) != 0;
    const KernelTable& X = exactKernels();                    // the mt19937 jitter stream (renderer.cpp:44-47) always comes from here
    const KernelTable& K = fastHook ? fastKernels() : X;      // default: primary-hit parity uses the no-FMA instantiation
    if (int rc = ensureWorkspace(s, nPixels, uint32_t(nPaths), 0, 1, true)) return rc;
    if (int rc = s->jitter.ensure(sizeof(float) * 2 * nPaths)) return rc;
    DQueues q = makeQueues(s);
    cudaStream_t st = s->stream;
    DWave w{};
    w.width = width; w.height = height; w.nPixels = nPixels; w.pixelBase = 0; w.wavePixels = nPixels; w.nPaths = uint32_t(nPaths);
    w.byWidth = makeFastDiv(uint32_t(width)); w.byWavePixels = makeFastDiv(uint32_t(nPixels));
    w.samplesThisWave = uint32_t(spp); w.sampleBase = 0; w.integrator = XRTG_INT_NORMAL; w.maxDepth = 1;
    w.sx0 = 0; w.sy0 = 0; w.sx1 = width; w.sy1 = height;
    w.mt = static_cast<uint32_t*>(s->mt.p);
    w.mti = static_cast<uint32_t*>(s->mti.p);
    float* dj = static_cast<float*>(s->jitter.p);
    if (jitter_uv) CU(cudaMemcpyAsync(dj, jitter_uv, sizeof(float) * 2 * nPaths, cudaMemcpyHostToDevice, st));
    else {
        X.seedMt(st, w);
        X.genJitter(st, w, spp, dj);
    }
    unsigned long long* dstats = static_cast<unsigned long long*>(s->stats.p);
    CU(cudaMemsetAsync(q.ctrl, 0, sizeof(uint32_t) * kCtrlStride * 3, st));
    const float4* hitsByPath = q.hits;
    if (fastHook && !choosePipeline(s, XRTG_INT_GI, 3, false, brute).deep) {
        // shallow BVH: the production primary kernel (raygen fused with the bounce-0 closest hit, screen-space scissor, compact
        // hit-only queue), fed with the supplied jitter; its queue is scattered back to one record per path
        const Pipeline P = choosePipeline(s, XRTG_INT_GI, 3, false, brute);
        if (P.scissor) {
            int sc4[4];
            computeScissor(s, cam, width, height, sc4);
            w.sx0 = sc4[0]; w.sy0 = sc4[1]; w.sx1 = sc4[2]; w.sy1 = sc4[3];
        }
        if (int rc = s->rayTmp[3].ensure(sizeof(float4) * nPaths)) return rc;
        if (tv(s->tuning.t.primary_masks, 1) != 0) w.primMask = preparePrimaryMasks(s, K, cam, width, height, st);
        K.primary(st, sceneForTraversal(s), makeCamera(cam), q, w, brute, 0, false, dstats, dj);
        K.scatterPrimaryHits(st, q, static_cast<float4*>(s->rayTmp[3].p), uint32_t(nPaths));
        hitsByPath = static_cast<const float4*>(s->rayTmp[3].p);
    }
    else {
        K.raygen(st, makeCamera(cam), q, w, dj);
        if (fastHook) { // deep BVH: the renderer's own bounce-0 launch (k_trace on the wide tree, production refill settings)
            const Pipeline P = choosePipeline(s, XRTG_INT_GI, 3, false, brute);
            K.extend(st, sceneForTraversal(s), q, 0, 0, brute ? 1 : 0, false, dstats, P.thrExt0, P.spv, P.leafThr);
        }
        else K.extend(st, sceneForTraversal(s), q, 0, 0, brute ? 1 : 0, false, dstats, 16, 1, 8);
    }
    CU(cudaGetLastError());
    // hits are indexed by path id = s * nPixels + pixel; the ABI wants [(pixel * spp) + s]
    std::vector<float4> tmp(nPaths);
    CU(cudaMemcpyAsync(tmp.data(), hitsByPath, sizeof(float4) * nPaths, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (uint64_t pid = 0; pid < nPaths; ++pid) {
        const uint64_t pix = pid % nPixels, k = pid / nPixels;
        xrtg_hit& h = out[pix * spp + k];
        const float4 v = tmp[pid];
        std::memcpy(&h.prim, &v.w, 4);
        h.t = h.prim >= 0 ? v.x : FLT_MAX;
        h.u = v.y; h.v = v.z;
    }
    return 0;
}

int xrtg_trace_rays(xrtg_scene* s, int64_t n, const float* org, const float* dir, const float* tmax, int any_hit, uint32_t flags,
                    xrtg_hit* out_hits)
{
    if (!s || !org || !dir || !out_hits) return fail(XRTG_ERR_INVALID, "NULL argument");
    if (n < 0) return fail(XRTG_ERR_INVALID, "negative ray count");
    if (n == 0) return 0;
    if (n > (1ll << 30)) return fail(XRTG_ERR_UNSUPPORTED, "too many rays for one call");
    CU(cudaSetDevice(s->device));
    cudaStream_t st = s->stream;
    // the rays are staged into the renderer's own queues and traced by the renderer's own traversal kernels
    if (int rc = ensureWorkspace(s, 1, uint32_t(n), size_t(n), 1, false)) return rc;
    if (int rc = s->rayTmp[0].ensure(sizeof(float) * 3 * size_t(n))) return rc;
    if (int rc = s->rayTmp[1].ensure(sizeof(float) * 3 * size_t(n))) return rc;
    if (int rc = s->rayTmp[2].ensure(sizeof(float) * size_t(n))) return rc;
    if (int rc = s->rayTmp[3].ensure(sizeof(float4) * size_t(n))) return rc;
    CU(cudaMemcpyAsync(s->rayTmp[0].p, org, sizeof(float) * 3 * size_t(n), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(s->rayTmp[1].p, dir, sizeof(float) * 3 * size_t(n), cudaMemcpyHostToDevice, st));
    if (tmax) CU(cudaMemcpyAsync(s->rayTmp[2].p, tmax, sizeof(float) * size_t(n), cudaMemcpyHostToDevice, st));
    DQueues q = makeQueues(s);
    CU(cudaMemsetAsync(q.ctrl, 0, sizeof(uint32_t) * kCtrlStride * 3, st));
    // XRTG_FLAG_HOOK_SRC_PRIM: out_hits[i].prim carries, on entry, the primitive the shadow ray starts on (what the renderer
    // knows when it traces an NEE ray); -1 everywhere otherwise
    const bool srcPrim = any_hit && (flags & XRTG_FLAG_FAST_HOOK) && (flags & XRTG_FLAG_HOOK_SRC_PRIM);
    if (srcPrim) CU(cudaMemcpyAsync(s->rayTmp[3].p, out_hits, sizeof(float4) * size_t(n), cudaMemcpyHostToDevice, st));
    else CU(cudaMemsetAsync(s->rayTmp[3].p, 0xff, sizeof(float4) * size_t(n), st));
    // default: the no-FMA instantiation through k_trace; XRTG_FLAG_FAST_HOOK: the throughput instantiation through the
    // production entry points of this scene (hookMode)
    const KernelTable& K = (flags & XRTG_FLAG_FAST_HOOK) ? fastKernels() : exactKernels();
    K.traceRays(st, sceneForTraversal(s), q, static_cast<const float*>(s->rayTmp[0].p), static_cast<const float*>(s->rayTmp[1].p),
                tmax ? static_cast<const float*>(s->rayTmp[2].p) : nullptr, n, any_hit != 0, (flags & XRTG_FLAG_BRUTE_FORCE) != 0,
                static_cast<float4*>(s->rayTmp[3].p), static_cast<unsigned long long*>(s->stats.p), hookMode(s, flags));
    CU(cudaGetLastError());
    static_assert(sizeof(xrtg_hit) == sizeof(float4), "xrtg_hit must be 16 bytes");
    CU(cudaMemcpyAsync(out_hits, any_hit ? s->rayTmp[3].p : static_cast<void*>(q.hits), sizeof(float4) * size_t(n), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

} // extern "C"
