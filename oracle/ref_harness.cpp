// ref_harness.cpp — C harness around the UNMODIFIED reference translation units compiled in place from
// /root/reference/Src (see oracle/Makefile). TEST INFRASTRUCTURE ONLY: nothing under oracle/ is linked,
// imported or executed by the product path (libxrtgpu.so / libxrthost.so); only tests/, smoke() and
// bench.py's cpu_baseline / --impl reference legs may load the resulting oracle/_ref/libxrtref.so.
// Built twice (oracle/Makefile): libxrtref.so = the CPU checker, no dependency on any product library (it only reads the
// POD definitions of xrtgpu.h); libxrtrefgpu.so = the same plus RefGpuRenderer (-DXRT_REF_WITH_GPU), linked against
// libxrtgpu.so, loaded by the drop-in test alone.
//
// What it adds on top of the reference (all additive, no reference source is copied):
//   * a reference `Scene` built from the same flattened xrtg_scene_desc the GPU consumes;
//   * DenseGrid : DensityGrid  — fp32 restatement of OpenVDBGrid's lookup (grid.h:58-84; OpenVDB is
//     an absent third-party dependency -> "parity unpinned" for the grid lookup);
//   * FurnaceIntegrator        — the dead furnace block of NormalIntegrator (integrator.h:59-66);
//   * an OpenMP pixel loop over the reference's own NormalRenderer::doRender (renderer.cpp:29-81),
//     because std::execution::par_unseq runs serially on libstdc++ without TBB;
//   * primitive-id recovery for closest hits (IntersectInfo has no primitive index, ray.h:34-39).
//
// `private`/`protected` are opened for THIS translation unit only so the harness can read
// Scene::m_objects, Mesh::m_primitives, PinholeCamera::scale etc. Class layout does not depend on
// access specifiers, so the other (unmodified) TUs stay ABI-compatible.
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <execution>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <map>
#include <stdexcept>
#include <memory>
#include <queue>
#include <random>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>
#include <omp.h>
#include <spdlog/spdlog.h>

#define private public
#define protected public
#define class struct /* DistantLight::dir / PointLight::pos are default-private members (light.h:30,44) */
#include "camera.h"
#include "grid.h"
#include "image.h"
#include "integrator.h"
#include "light.h"
#include "material.h"
#include "medium.h"
#include "primitive.h"
#include "renderer.h"
#include "scene.h"
#undef class
#undef private
#undef protected

#include "xrtgpu.h"

namespace {

// ---- restated pieces ---------------------------------------------------------------------------------

// Dense fp32 restatement of OpenVDBGrid (grid.h:22-85): world->index is (p-origin)/voxel, BoxSampler
// trilinear with z interpolated first, then y, then x, each lerp a + (b-a)*w; `background` outside the
// allocated block. OpenVDB does the weight math in double; this restatement is fp32 throughout and the
// GPU kernel follows THIS order of operations.
class DenseGrid : public DensityGrid {
public:
    explicit DenseGrid(const xrtg_grid& g) : g(g), data(g.data, g.data + size_t(g.nx) * g.ny * g.nz) {}

    AABB getBounds() const override
    {
        // indexToWorld(bbox.getStart()), indexToWorld(bbox.getEnd()); CoordBBox::getEnd() = max + 1
        AABB r;
        for (int a = 0; a < 3; ++a) {
            r.pMin[a] = g.origin[a] + g.voxel_size * float(g.active_min[a]);
            r.pMax[a] = g.origin[a] + g.voxel_size * float(g.active_max[a] + 1);
        }
        return r;
    }

    float voxel(int x, int y, int z) const
    {
        if (x < 0 || y < 0 || z < 0 || x >= g.nx || y >= g.ny || z >= g.nz) return g.background;
        return data[(size_t(z) * g.ny + y) * g.nx + x];
    }

    float getDensity(const Vec3f& p) const override
    {
        const float fx = (p[0] - g.origin[0]) / g.voxel_size;
        const float fy = (p[1] - g.origin[1]) / g.voxel_size;
        const float fz = (p[2] - g.origin[2]) / g.voxel_size;
        const float bx = std::floor(fx), by = std::floor(fy), bz = std::floor(fz);
        const float wx = fx - bx, wy = fy - by, wz = fz - bz;
        const int x = int(bx), y = int(by), z = int(bz);
        auto lerp = [](float a, float b, float w) { return a + (b - a) * w; };
        const float c00 = lerp(voxel(x, y, z), voxel(x, y, z + 1), wz);
        const float c01 = lerp(voxel(x, y + 1, z), voxel(x, y + 1, z + 1), wz);
        const float c10 = lerp(voxel(x + 1, y, z), voxel(x + 1, y, z + 1), wz);
        const float c11 = lerp(voxel(x + 1, y + 1, z), voxel(x + 1, y + 1, z + 1), wz);
        return lerp(lerp(c00, c01, wy), lerp(c10, c11, wy), wx);
    }

    float getMaxDensity() const override { return g.max_density; }
    xrtg_grid desc() const { xrtg_grid d = g; d.data = data.data(); return d; }

private:
    xrtg_grid g;
    std::vector<float> data;
};

// The furnace block that sits behind the early return of NormalIntegrator (integrator.h:59-66).
class FurnaceIntegrator : public Integrator {
public:
    Vec3f integrate(const Ray& ray_in, const Scene& scene, Sampler& sampler) const override
    {
        Vec3f radiance(0);
        IntersectInfo info;
        if (scene.intersect(ray_in, info)) {
            float pdf = 1.0f;
            Vec3f nextDir(0.0f);
            auto fr = info.hitObject->sampleBxDF(ray_in.direction, info.surfaceInfo, sampler, nextDir, pdf);
            float cos = std::max(0.0f, dot(nextDir, info.surfaceInfo.ng));
            Vec3f Li(1.0f);
            radiance = fr * cos * Li / pdf;
        }
        return radiance;
    }
};

// Calls the reference's protected per-pixel routine from an OpenMP loop.
class OmpRenderer : public NormalRenderer {
public:
    OmpRenderer(uint32_t spp, Camera* cam, Integrator* inte) : NormalRenderer(spp, cam, inte) {}
    void renderOmp(const Scene& scene, Image& image, int nthreads) const
    {
        const int W = int(image.getWidth()), H = int(image.getHeight());
        const long n = long(W) * H;
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads)
        for (long p = 0; p < n; ++p) { doRender(scene, Sampler::SamplerType::Uniform, image, int(p / W), int(p % W)); }
        image /= Vec3f(float(n_samples)); // renderer.cpp:98
    }
    // Every `stride`-th pixel in x and y only (bounded CPU samples of big scenes).
    void renderOmpStrided(const Scene& scene, Image& image, int nthreads, int stride) const
    {
        const int W = int(image.getWidth()), H = int(image.getHeight());
        const int nx = (W + stride - 1) / stride, ny = (H + stride - 1) / stride;
        const long n = long(nx) * ny;
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
        for (long p = 0; p < n; ++p) {
            doRender(scene, Sampler::SamplerType::Uniform, image, int(p / nx) * stride, int(p % nx) * stride);
        }
        image /= Vec3f(float(n_samples));
    }
};

thread_local std::string g_err;

Vec3f v3(const float* p) { return Vec3f(p[0], p[1], p[2]); }

} // namespace

struct xrtref_scene {
    Scene scene;
    std::vector<std::unique_ptr<Material>> materials;
    std::vector<std::unique_ptr<DenseGrid>> grids;
    std::vector<std::unique_ptr<Medium>> media;
    std::unordered_map<const Object*, int> firstPrim; // global primitive id of an object's first primitive
    std::unordered_map<std::string, int> insertSeqByName;
    int nPrims = 0;
};


#ifdef XRT_REF_WITH_GPU // only libxrtrefgpu.so (tests/test_gpu_dropin_reference.py) carries the GPU binding and links libxrtgpu.so;
                        // libxrtref.so — the CPU checker and the bench's reference arm — contains no reference to the product
// ---------------------------------------------------------------------------------------------------------------------
// RefGpuRenderer — the binding of INTEGRATION.md, compiled against the reference's OWN headers: a third sibling of
// NormalRenderer / ParallelRenderer (renderer.h:22-47) that flattens a reference `Scene` and renders it through the C ABI
// of libxrtgpu.so. It stands in for the `gpu_renderer.cpp` a maintainer would add to the reference tree; where that file
// would use the additive accessors, this test harness reads the private members directly (access opened above).
// ---------------------------------------------------------------------------------------------------------------------
namespace {

struct RefFlat {
    std::vector<xrtg_object> objects;
    std::vector<std::string> names;
    std::vector<xrtg_triangle> tris;
    std::vector<xrtg_sphere> spheres;
    std::vector<xrtg_box> boxes;
    std::vector<xrtg_material> materials;
    std::vector<xrtg_area_light> lights;
    std::vector<xrtg_delta_light> dlights;
    std::vector<xrtg_medium> media;
    std::vector<xrtg_grid> grids;
    xrtg_scene_desc desc{};
};

void put3(float* d, const Vec3f& v) { d[0] = v[0]; d[1] = v[1]; d[2] = v[2]; }

bool flattenReferenceScene(const Scene& scene, RefFlat& f, std::string& err)
{
    std::map<const Material*, int> matIdx;
    std::map<const AreaLight*, int> lightIdx;
    std::map<const Medium*, int> medIdx;
    for (const auto& L : scene.m_areaLights) {
        xrtg_area_light d{};
        if (auto* q = dynamic_cast<const QuadLight*>(L.get())) { d.kind = XRTG_LIGHT_QUAD; put3(d.v0, q->v0_); put3(d.v1, q->v1_); put3(d.v2, q->v2_); }
        else if (auto* t = dynamic_cast<const TriangleLight*>(L.get())) { d.kind = XRTG_LIGHT_TRIANGLE; put3(d.v0, t->v0_); put3(d.v1, t->v1_); put3(d.v2, t->v2_); }
        else if (auto* sp = dynamic_cast<const SphereLight*>(L.get())) { d.kind = XRTG_LIGHT_SPHERE; put3(d.v0, sp->center_); d.radius = sp->radius_; }
        else { err = "unknown AreaLight subclass"; return false; }
        put3(d.Le, L->Le_);
        lightIdx[L.get()] = int(f.lights.size());
        f.lights.push_back(d);
    }
    for (const auto& L : scene.m_deltaLights) {
        xrtg_delta_light d{};
        if (auto* p = dynamic_cast<const PointLight*>(L.get())) { d.kind = XRTG_DLIGHT_POINT; put3(d.pos_or_dir, p->pos); }
        else if (auto* dl = dynamic_cast<const DistantLight*>(L.get())) { d.kind = XRTG_DLIGHT_DISTANT; put3(d.pos_or_dir, dl->dir); }
        else { err = "unknown DeltaLight subclass"; return false; }
        put3(d.radiance, L->color * L->intensity);
        f.dlights.push_back(d);
    }
    int seq = 0;
    f.names.reserve(scene.m_objects.size());
    for (const auto& [name, obj] : scene.m_objects) { // the order Scene::intersect walks (scene.cpp:193)
        xrtg_object rec{};
        rec.material = rec.area_light = rec.medium = -1;
        rec.insert_seq = seq++;
        if (auto* m = dynamic_cast<const Mesh*>(obj.get())) {
            rec.kind = XRTG_OBJ_MESH; rec.first = int(f.tris.size()); rec.count = int(m->m_primitives.size());
            for (const auto& p : m->m_primitives) {
                xrtg_triangle t;
                put3(t.v0, p.vertices()[0]); put3(t.v1, p.vertices()[1]); put3(t.v2, p.vertices()[2]);
                put3(t.n0, p.normals()[0]); put3(t.n1, p.normals()[1]); put3(t.n2, p.normals()[2]);
                f.tris.push_back(t);
            }
        }
        else if (auto* sp = dynamic_cast<const Sphere*>(obj.get())) {
            rec.kind = XRTG_OBJ_SPHERE; rec.first = int(f.spheres.size()); rec.count = 1;
            xrtg_sphere d{};
            put3(d.center, sp->m_center); d.radius = sp->m_raduis;
            f.spheres.push_back(d);
        }
        else if (auto* bx = dynamic_cast<const BoxMesh*>(obj.get())) {
            rec.kind = XRTG_OBJ_BOX; rec.first = int(f.boxes.size()); rec.count = 1;
            xrtg_box d{};
            put3(d.pmin, bx->box.pMin); put3(d.pmax, bx->box.pMax);
            f.boxes.push_back(d);
        }
        else { err = "unknown Object subclass"; return false; }
        if (const Material* mat = obj->m_material) {
            auto it = matIdx.find(mat);
            if (it == matIdx.end()) {
                auto* lam = dynamic_cast<const Lambert*>(mat);
                if (!lam) { err = "material without a GPU implementation"; return false; }
                xrtg_material d{};
                d.kind = XRTG_MAT_LAMBERT; put3(d.albedo, lam->m_albedo);
                it = matIdx.emplace(mat, int(f.materials.size())).first;
                f.materials.push_back(d);
            }
            rec.material = it->second;
        }
        if (obj->m_areaLight) rec.area_light = lightIdx.at(obj->m_areaLight);
        if (const Medium* med = obj->m_medium) {
            auto it = medIdx.find(med);
            if (it == medIdx.end()) {
                xrtg_medium d{};
                d.grid = -1; d.density_mul = 1.0f;
                d.g = static_cast<const HenyeyGreenstein*>(med->phaseFunction.get())->g;
                if (auto* h = dynamic_cast<const HeterogeneousMedium*>(med)) {
                    d.kind = XRTG_MEDIUM_HETEROGENEOUS;
                    put3(d.sigma_a, h->absorptionColor); put3(d.sigma_s, h->scatteringColor); d.density_mul = h->densityMultiplier;
                    auto* dg = dynamic_cast<const DenseGrid*>(h->densityGridPtr);
                    if (!dg) { err = "density grid without a GPU implementation"; return false; }
                    d.grid = int(f.grids.size());
                    f.grids.push_back(dg->desc());
                }
                else if (auto* hm = dynamic_cast<const HomogeneousMedium*>(med)) {
                    d.kind = dynamic_cast<const HomogeneousMediumMIS*>(med) ? XRTG_MEDIUM_HOMOGENEOUS_MIS
                             : (dynamic_cast<const HomogeneousMediumAchromatic*>(med) ? XRTG_MEDIUM_HOMOGENEOUS_ACHROMATIC : XRTG_MEDIUM_HOMOGENEOUS_NOMIS);
                    put3(d.sigma_a, hm->sigma_a); put3(d.sigma_s, hm->sigma_s);
                }
                else { err = "unknown Medium subclass"; return false; }
                it = medIdx.emplace(med, int(f.media.size())).first;
                f.media.push_back(d);
            }
            rec.medium = it->second;
        }
        f.names.push_back(name);
        f.objects.push_back(rec);
    }
    for (size_t i = 0; i < f.objects.size(); ++i) f.objects[i].name = f.names[i].c_str();
    xrtg_scene_desc& d = f.desc;
    d.abi_version = XRTG_ABI_VERSION;
    d.n_objects = int(f.objects.size()); d.objects = f.objects.data();
    d.n_triangles = int(f.tris.size()); d.triangles = f.tris.data();
    d.n_spheres = int(f.spheres.size()); d.spheres = f.spheres.data();
    d.n_boxes = int(f.boxes.size()); d.boxes = f.boxes.data();
    d.n_materials = int(f.materials.size()); d.materials = f.materials.data();
    d.n_area_lights = int(f.lights.size()); d.area_lights = f.lights.data();
    d.n_delta_lights = int(f.dlights.size()); d.delta_lights = f.dlights.data();
    d.n_media = int(f.media.size()); d.media = f.media.data();
    d.n_grids = int(f.grids.size()); d.grids = f.grids.data();
    return true;
}

class RefGpuRenderer : public Renderer { // the reference's plug-in point, renderer.h:8-20
public:
    RefGpuRenderer(uint32_t spp, Camera* cam, Integrator* inte, uint32_t flags, uint32_t seed)
        : Renderer(cam, inte), n_samples(spp), flags(flags), seed(seed) {}

    void render(const Scene& scene, Sampler::SamplerType, Image& image) const override
    {
        RefFlat flat;
        std::string err;
        if (!flattenReferenceScene(scene, flat, err)) throw std::runtime_error("[RefGpuRenderer] " + err);
        xrtg_scene* gs = nullptr;
        if (xrtg_scene_create(&flat.desc, 0, &gs) != XRTG_OK) throw std::runtime_error(std::string("[RefGpuRenderer] ") + xrtg_last_error());
        auto* pc = dynamic_cast<const PinholeCamera*>(camera);
        if (!pc) { xrtg_scene_destroy(gs); throw std::runtime_error("[RefGpuRenderer] camera model without a GPU implementation"); }
        xrtg_camera cam{};
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) cam.c2w[4 * r + c] = pc->camera2world[r][c];
        cam.scale = pc->scale; cam.aspect = pc->aspect_ratio;
        xrtg_render_params p{};
        p.width = int(image.getWidth()); p.height = int(image.getHeight());
        p.spp = p.spp_total = int(n_samples);
        p.flags = flags; p.seed = seed; p.max_depth = 1;
        if (dynamic_cast<const NormalIntegrator*>(integrator)) p.integrator = XRTG_INT_NORMAL;
        else if (dynamic_cast<const FurnaceIntegrator*>(integrator)) p.integrator = XRTG_INT_FURNACE;
        else if (dynamic_cast<const DirectIntegrator*>(integrator)) p.integrator = XRTG_INT_DIRECT;
        else if (auto* i1 = dynamic_cast<const IndirectIntegrator*>(integrator)) { p.integrator = XRTG_INT_INDIRECT; p.max_depth = int(i1->m_maxDepth); }
        else if (auto* i2 = dynamic_cast<const GIIntegrator*>(integrator)) { p.integrator = XRTG_INT_GI; p.max_depth = int(i2->m_maxDepth); }
        else if (auto* i3 = dynamic_cast<const WhittedIntegrator*>(integrator)) { p.integrator = XRTG_INT_WHITTED; p.max_depth = int(i3->m_maxDepth); }
        else if (auto* i4 = dynamic_cast<const VolumePathTracing*>(integrator)) { p.integrator = XRTG_INT_VOLUME; p.max_depth = int(i4->m_maxDepth); }
        else if (auto* i5 = dynamic_cast<const VolumePathTracingNEE*>(integrator)) { p.integrator = XRTG_INT_VOLUME_NEE; p.max_depth = int(i5->m_maxDepth); }
        else { xrtg_scene_destroy(gs); throw std::runtime_error("[RefGpuRenderer] integrator without a GPU implementation"); }
        static_assert(sizeof(Vec3f) == 3 * sizeof(float), "Image::pixels must be tightly packed RGB");
        const int rc = xrtg_render(gs, &cam, &p, &image.pixels[0][0], nullptr);
        const std::string msg = rc == XRTG_OK ? "" : xrtg_last_error();
        xrtg_scene_destroy(gs);
        if (rc != XRTG_OK) throw std::runtime_error("[RefGpuRenderer] " + msg);
    }

private:
    const uint32_t n_samples, flags, seed;
};

} // namespace

#endif // XRT_REF_WITH_GPU

static std::unique_ptr<AreaLight> makeAreaLight(const xrtg_area_light& L)
{
    const Matrix44f I;
    switch (L.kind) {
    case XRTG_LIGHT_QUAD: return std::make_unique<QuadLight>(v3(L.v0), v3(L.v1), v3(L.v2), I, v3(L.Le));
    case XRTG_LIGHT_TRIANGLE: return std::make_unique<TriangleLight>(v3(L.v0), v3(L.v1), v3(L.v2), I, v3(L.Le));
    case XRTG_LIGHT_SPHERE: return std::make_unique<SphereLight>(v3(L.v0), L.radius, I, v3(L.Le));
    default: return nullptr;
    }
}

#pragma GCC visibility push(default)
extern "C" {

const char* xrtref_last_error(void) { return g_err.c_str(); }

int xrtref_max_threads(void) { return omp_get_max_threads(); }

void xrtref_scene_destroy(xrtref_scene* s) { delete s; }

// Builds a reference Scene by replaying the host Scene's insertion sequence (objects sorted by
// insert_seq) through the reference's own addObj / addAreaLight / addDeltaLight.
int xrtref_scene_create(const xrtg_scene_desc* d, xrtref_scene** out)
{
    spdlog::set_level(spdlog::level::off);
    if (!d || !out || d->abi_version != XRTG_ABI_VERSION) { g_err = "bad scene desc"; return -1; }
    auto s = std::make_unique<xrtref_scene>();
    for (int i = 0; i < d->n_materials; ++i) s->materials.push_back(std::make_unique<Lambert>(v3(d->materials[i].albedo)));
    for (int i = 0; i < d->n_grids; ++i) s->grids.push_back(std::make_unique<DenseGrid>(d->grids[i]));

    // media need the box of the object that references them (HomogeneousMedium holds its AABB)
    std::vector<AABB> mediumBox(d->n_media);
    for (int i = 0; i < d->n_objects; ++i) {
        const xrtg_object& o = d->objects[i];
        if (o.kind == XRTG_OBJ_BOX && o.medium >= 0) {
            const xrtg_box& b = d->boxes[o.first];
            mediumBox[o.medium] = AABB{v3(b.pmin), v3(b.pmax)};
        }
    }
    for (int i = 0; i < d->n_media; ++i) {
        const xrtg_medium& m = d->media[i];
        switch (m.kind) {
        case XRTG_MEDIUM_HOMOGENEOUS_MIS:
            s->media.push_back(std::make_unique<HomogeneousMediumMIS>(m.g, v3(m.sigma_a), v3(m.sigma_s), mediumBox[i]));
            break;
        case XRTG_MEDIUM_HOMOGENEOUS_ACHROMATIC:
            s->media.push_back(std::make_unique<HomogeneousMediumAchromatic>(m.g, m.sigma_a[0], m.sigma_s[0], mediumBox[i]));
            break;
        case XRTG_MEDIUM_HOMOGENEOUS_NOMIS:
            s->media.push_back(std::make_unique<HomogeneousMediumNoMIS>(m.g, v3(m.sigma_a), v3(m.sigma_s), mediumBox[i]));
            break;
        case XRTG_MEDIUM_HETEROGENEOUS:
            if (m.grid < 0 || m.grid >= d->n_grids) { g_err = "medium without grid"; return -1; }
            s->media.push_back(std::make_unique<HeterogeneousMedium>(m.g, s->grids[m.grid].get(), v3(m.sigma_a),
                                                                     v3(m.sigma_s), m.density_mul));
            break;
        default: g_err = "unknown medium kind"; return -1;
        }
    }

    std::vector<int> order(d->n_objects);
    for (int i = 0; i < d->n_objects; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(),
                     [&](int a, int b) { return d->objects[a].insert_seq < d->objects[b].insert_seq; });

    // Area lights must be appended to m_areaLights in area_lights[] order (scene.cpp:166-170); the host
    // Scene guarantees proxies are inserted in that order.
    int nextLight = 0;
    for (int oi : order) {
        const xrtg_object& o = d->objects[oi];
        const std::string name = o.name ? o.name : ("obj" + std::to_string(o.insert_seq));
        s->insertSeqByName[name] = o.insert_seq;
        if (o.area_light >= 0) {
            if (o.area_light != nextLight) { g_err = "area light proxies not inserted in light order"; return -1; }
            auto L = makeAreaLight(d->area_lights[o.area_light]);
            if (!L) { g_err = "unknown area light kind"; return -1; }
            s->scene.addAreaLight(name, std::move(L));
            ++nextLight;
            continue;
        }
        Material* mat = o.material >= 0 ? s->materials[o.material].get() : nullptr;
        if (o.kind == XRTG_OBJ_MESH) {
            std::vector<Primitive> prims;
            prims.reserve(o.count);
            for (int k = 0; k < o.count; ++k) {
                const xrtg_triangle& t = d->triangles[o.first + k];
                // texcoords as scene.cpp:128-132 synthesises them
                prims.emplace_back(std::vector<Vec3f>{v3(t.v0), v3(t.v1), v3(t.v2)},
                                   std::vector<Vec3f>{v3(t.n0), v3(t.n1), v3(t.n2)},
                                   std::vector<Vec2f>{Vec2f(0, 0), Vec2f(1, 0), Vec2f(0, 1)});
            }
            s->scene.addObj(name, std::make_unique<Mesh>(std::move(prims), mat, nullptr));
        }
        else if (o.kind == XRTG_OBJ_SPHERE) {
            const xrtg_sphere& sp = d->spheres[o.first];
            s->scene.addObj(name, std::make_unique<Sphere>(v3(sp.center), sp.radius, mat, nullptr));
        }
        else if (o.kind == XRTG_OBJ_BOX) {
            if (o.medium < 0) { g_err = "box without medium"; return -1; }
            s->scene.addObj(name, s->media[o.medium]->makeObject());
        }
        else { g_err = "unknown object kind"; return -1; }
    }
    if (nextLight != d->n_area_lights) { g_err = "area light without proxy object"; return -1; }

    for (int i = 0; i < d->n_delta_lights; ++i) {
        const xrtg_delta_light& L = d->delta_lights[i];
        const Matrix44f I;
        if (L.kind == XRTG_DLIGHT_POINT) {
            auto p = std::make_unique<PointLight>(I, v3(L.radiance), 1.0f);
            p->pos = v3(L.pos_or_dir);
            s->scene.addDeltaLight("point" + std::to_string(i), std::move(p));
        }
        else {
            auto p = std::make_unique<DistantLight>(I, v3(L.radiance), 1.0f);
            p->dir = v3(L.pos_or_dir);
            s->scene.addDeltaLight("distant" + std::to_string(i), std::move(p));
        }
    }

    // global primitive ids in the map's actual iteration order (scene.cpp:193)
    int next = 0;
    for (const auto& [name, obj] : s->scene.m_objects) {
        s->firstPrim[obj.get()] = next;
        if (auto* m = dynamic_cast<const Mesh*>(obj.get())) next += int(m->m_primitives.size());
        else next += 1;
    }
    s->nPrims = next;
    *out = s.release();
    return 0;
}

// insert_seq of every object in the reference map's iteration order; returns the object count.
int xrtref_object_order(const xrtref_scene* s, int32_t* out, int cap)
{
    int n = 0;
    for (const auto& [name, obj] : s->scene.m_objects) {
        if (n < cap) out[n] = s->insertSeqByName.at(name);
        ++n;
    }
    return n;
}

static void fillHit(const xrtref_scene* s, const Ray& ray, bool hit, const IntersectInfo& info, xrtg_hit* h)
{
    h->t = info.t; h->u = 0; h->v = 0; h->prim = -1;
    if (!hit || !info.hitObject) { h->t = FLT_MAX; return; }
    const int first = s->firstPrim.at(info.hitObject);
    if (auto* m = dynamic_cast<const Mesh*>(info.hitObject)) {
        h->u = info.surfaceInfo.barycentric[0];
        h->v = info.surfaceInfo.barycentric[1];
        // first triangle of the mesh reaching info.t == the one primitive.cpp:100 kept
        for (size_t k = 0; k < m->m_primitives.size(); ++k) {
            const auto& v = m->m_primitives[k].vertices();
            float t, u, vv;
            if (m->rayTriangleIntersect(ray.origin, ray.direction, v[0], v[1], v[2], t, u, vv) && t == info.t) {
                h->prim = first + int(k);
                return;
            }
        }
        h->prim = -2; // cannot happen
    }
    else {
        h->prim = first;
        if (dynamic_cast<const BoxMesh*>(info.hitObject)) h->u = info.t1;
    }
}

static std::unique_ptr<PinholeCamera> makeCamera(const xrtg_camera* c)
{
    Matrix44f m(c->c2w[0], c->c2w[1], c->c2w[2], c->c2w[3], c->c2w[4], c->c2w[5], c->c2w[6], c->c2w[7], c->c2w[8],
                c->c2w[9], c->c2w[10], c->c2w[11], c->c2w[12], c->c2w[13], c->c2w[14], c->c2w[15]);
    auto cam = std::make_unique<PinholeCamera>(c->aspect, m, 90.0f);
    cam->scale = c->scale; // tan(FOV/2) is evaluated by the caller (camera.h:44)
    return cam;
}

static std::unique_ptr<Integrator> makeIntegrator(int kind, int maxDepth)
{
    switch (kind) {
    case XRTG_INT_NORMAL: return std::make_unique<NormalIntegrator>();
    case XRTG_INT_FURNACE: return std::make_unique<FurnaceIntegrator>();
    case XRTG_INT_DIRECT: return std::make_unique<DirectIntegrator>();
    case XRTG_INT_INDIRECT: return std::make_unique<IndirectIntegrator>(maxDepth);
    case XRTG_INT_GI: return std::make_unique<GIIntegrator>(maxDepth);
    case XRTG_INT_WHITTED: return std::make_unique<WhittedIntegrator>(uint32_t(maxDepth));
    case XRTG_INT_VOLUME: return std::make_unique<VolumePathTracing>(uint32_t(maxDepth));
    case XRTG_INT_VOLUME_NEE: return std::make_unique<VolumePathTracingNEE>(uint32_t(maxDepth));
    default: return nullptr;
    }
}

// The reference's render loop (doRender per pixel, mt19937 seeded j+W*i, divide by spp) on `nthreads`
// OpenMP threads. pixel_stride>1 renders only every stride-th pixel in x and y (others stay 0).
// rgb = W*H*3 floats. seconds_out = wall time of the pixel loop only (scene build excluded).
int xrtref_render(xrtref_scene* s, const xrtg_camera* c, const xrtg_render_params* p, int nthreads, int pixel_stride,
                  float* rgb, double* seconds_out)
{
    spdlog::set_level(spdlog::level::off);
    if (p->sample_offset != 0) { g_err = "the reference has no sample offset"; return -1; }
    auto cam = makeCamera(c);
    auto integ = makeIntegrator(p->integrator, p->max_depth);
    if (!integ) { g_err = "unknown integrator"; return -1; }
    Image image(p->width, p->height);
    OmpRenderer r(uint32_t(p->spp), cam.get(), integ.get());
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    const auto t0 = std::chrono::steady_clock::now();
    if (pixel_stride > 1) r.renderOmpStrided(s->scene, image, nthreads, pixel_stride);
    else r.renderOmp(s->scene, image, nthreads);
    const auto t1 = std::chrono::steady_clock::now();
    if (seconds_out) *seconds_out = std::chrono::duration<double>(t1 - t0).count();
    if (rgb) {
        for (int i = 0; i < p->height; ++i)
            for (int j = 0; j < p->width; ++j) {
                const Vec3f v = image.getPixel(i, j);
                float* o = rgb + (size_t(i) * p->width + j) * 3;
                o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
            }
    }
    return 0;
}

#ifdef XRT_REF_WITH_GPU
// The same reference Scene / Camera / Integrator objects rendered through the GPU sibling renderer, called polymorphically
// through the reference's `Renderer*` exactly as examples/cornellbox.cpp:61-63 would.
int xrtref_render_gpu(xrtref_scene* s, const xrtg_camera* c, const xrtg_render_params* p, float* rgb)
{
    spdlog::set_level(spdlog::level::off);
    auto cam = makeCamera(c);
    auto integ = makeIntegrator(p->integrator, p->max_depth);
    if (!integ) { g_err = "unknown integrator"; return -1; }
    Image image(p->width, p->height);
    std::unique_ptr<Renderer> renderer = std::make_unique<RefGpuRenderer>(uint32_t(p->spp), cam.get(), integ.get(), p->flags, p->seed);
    try {
        renderer->render(s->scene, Sampler::SamplerType::Uniform, image);
    }
    catch (const std::exception& e) {
        g_err = e.what();
        return -2;
    }
    for (int i = 0; i < p->height; ++i)
        for (int j = 0; j < p->width; ++j) {
            const Vec3f v = image.getPixel(i, j);
            float* o = rgb + (size_t(i) * p->width + j) * 3;
            o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
        }
    return 0;
}

#endif // XRT_REF_WITH_GPU

// ParallelRenderer::render exactly as written (PSTL; serial on libstdc++ without TBB) — for the record.
int xrtref_render_pstl(xrtref_scene* s, const xrtg_camera* c, const xrtg_render_params* p, float* rgb, double* seconds_out)
{
    spdlog::set_level(spdlog::level::off);
    auto cam = makeCamera(c);
    auto integ = makeIntegrator(p->integrator, p->max_depth);
    if (!integ) { g_err = "unknown integrator"; return -1; }
    Image image(p->width, p->height);
    ParallelRenderer r(uint32_t(p->spp), cam.get(), integ.get());
    const auto t0 = std::chrono::steady_clock::now();
    r.render(s->scene, Sampler::SamplerType::Uniform, image);
    const auto t1 = std::chrono::steady_clock::now();
    if (seconds_out) *seconds_out = std::chrono::duration<double>(t1 - t0).count();
    if (rgb) {
        for (int i = 0; i < p->height; ++i)
            for (int j = 0; j < p->width; ++j) {
                const Vec3f v = image.getPixel(i, j);
                float* o = rgb + (size_t(i) * p->width + j) * 3;
                o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
            }
    }
    return 0;
}

// Primary rays as renderer.cpp:44-52 forms them + Scene::intersect; jitter from the caller or from the
// pixel's own mt19937 stream (two draws per sample, nothing else consumes the stream here).
int xrtref_trace_primary(xrtref_scene* s, const xrtg_camera* c, int W, int H, int spp, const float* jitter,
                         xrtg_hit* out)
{
    spdlog::set_level(spdlog::level::off);
    auto cam = makeCamera(c);
#pragma omp parallel for schedule(dynamic, 16)
    for (long p = 0; p < long(W) * H; ++p) {
        const int i = int(p / W), j = int(p % W);
        UniformSampler sampler;
        sampler.setSeed(uint32_t(j + W * i));
        for (int k = 0; k < spp; ++k) {
            float r0, r1;
            if (jitter) { r0 = jitter[(p * spp + k) * 2]; r1 = jitter[(p * spp + k) * 2 + 1]; }
            else { r0 = sampler.getNext1D(); r1 = sampler.getNext1D(); }
            const float u = (j + r0) / uint32_t(W);
            const float v = (i + r1) / uint32_t(H);
            Ray ray; float pdf;
            cam->sampleRay(Vec2f(u, v), sampler, ray, pdf);
            IntersectInfo info;
            const bool hit = s->scene.intersect(ray, info);
            fillHit(s, ray, hit, info, out + p * spp + k);
        }
    }
    return 0;
}

int xrtref_trace_rays(xrtref_scene* s, int64_t n, const float* org, const float* dir, const float* tmax, int any_hit,
                      xrtg_hit* out)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t r = 0; r < n; ++r) {
        Ray ray(v3(org + 3 * r), v3(dir + 3 * r));
        if (any_hit) {
            const bool occ = s->scene.occluded(ray, tmax ? tmax[r] : FLT_MAX);
            out[r].t = 0; out[r].u = 0; out[r].v = 0; out[r].prim = occ ? 1 : 0;
        }
        else {
            IntersectInfo info;
            const bool hit = s->scene.intersect(ray, info);
            fillHit(s, ray, hit, info, out + r);
        }
    }
    return 0;
}

// ---- known-answer hooks on single reference functions ---------------------------------------------------

// n floats from UniformSampler seeded `seed` (sampler.h:37-50): the libstdc++ mt19937 -> float mapping.
void xrtref_kat_sampler(uint32_t seed, int n, float* out)
{
    UniformSampler s;
    s.setSeed(seed);
    for (int i = 0; i < n; ++i) out[i] = s.getNext1D();
}

// PinholeCamera::sampleRay (camera.h:49-60): out = origin xyz, direction xyz.
void xrtref_kat_camera(const xrtg_camera* c, float u, float v, float* out6)
{
    spdlog::set_level(spdlog::level::off);
    auto cam = makeCamera(c);
    UniformSampler s;
    Ray ray; float pdf;
    cam->sampleRay(Vec2f(u, v), s, ray, pdf);
    for (int a = 0; a < 3; ++a) { out6[a] = ray.origin[a]; out6[3 + a] = ray.direction[a]; }
}

// orthonormalBasis (geometry.cpp:23-50): out = t xyz, b xyz.
void xrtref_kat_onb(const float* n, float* out6)
{
    Vec3f t, b;
    orthonormalBasis(v3(n), t, b);
    for (int a = 0; a < 3; ++a) { out6[a] = t[a]; out6[3 + a] = b[a]; }
}

// AreaLight::sample for light `li` of the scene from `pos`, sampler seeded `seed`:
// out = L rgb, wi xyz, pdf, tmax.
void xrtref_kat_light_sample(xrtref_scene* s, int li, const float* pos, uint32_t seed, float* out8)
{
    UniformSampler sm;
    sm.setSeed(seed);
    Vec3f wi; float pdf = 0, tmax = 0;
    const Vec3f L = s->scene.getAreaLights()[li]->sample(v3(pos), wi, pdf, tmax, sm);
    for (int a = 0; a < 3; ++a) { out8[a] = L[a]; out8[3 + a] = wi[a]; }
    out8[6] = pdf; out8[7] = tmax;
}

// Lambert::sampleDir through a SurfaceInfo built from (ng, ns) like Mesh::intersect does
// (primitive.cpp:105-109): out = wi xyz, pdf.
void xrtref_kat_lambert_sample(const float* ng, const float* ns, uint32_t seed, float* out4)
{
    Lambert m(Vec3f(1.0f));
    SurfaceInfo si;
    si.ng = v3(ng); si.ns = v3(ns);
    orthonormalBasis(si.ns, si.dpdu, si.dpdv);
    UniformSampler sm;
    sm.setSeed(seed);
    float pdf;
    const Vec3f wi = m.sampleDir(si, sm, pdf);
    out4[0] = wi[0]; out4[1] = wi[1]; out4[2] = wi[2]; out4[3] = pdf;
}

// HenyeyGreenstein::sampleDirection (medium.h:38-67): out = wi xyz, phase value.
void xrtref_kat_hg_sample(float g, const float* wo, uint32_t seed, float* out4)
{
    HenyeyGreenstein hg(g);
    UniformSampler sm;
    sm.setSeed(seed);
    Vec3f wi;
    const float f = hg.sampleDirection(v3(wo), sm, wi);
    out4[0] = wi[0]; out4[1] = wi[1]; out4[2] = wi[2]; out4[3] = f;
}

// SphereMesh::Triangulate (primitive.cpp:170-205): writes 2*nt*np triangles as 18 floats each (v0 v1 v2 n0 n1 n2);
// returns the triangle count.
int xrtref_kat_sphere_mesh(const float* center, float radius, int nt, int np, float* out, int cap)
{
    SphereMesh sm(v3(center), radius, nt, np, nullptr, nullptr);
    int n = 0;
    for (const auto& p : sm.m_primitives) {
        if (n < cap) {
            float* o = out + size_t(n) * 18;
            for (int k = 0; k < 3; ++k)
                for (int a = 0; a < 3; ++a) { o[3 * k + a] = p.vertices()[k][a]; o[9 + 3 * k + a] = p.normals()[k][a]; }
        }
        ++n;
    }
    return n;
}

// Constructor arithmetic of the host-side classes the GPU path mirrors (light.cpp:6-14,49-57,84-90,115-134; camera.h:41-47):
// kind 0 quad / 1 triangle: out9 = v0_ v1_ v2_ ; 2 sphere: out[0..2] = center_ ; 3 point: pos ; 4 distant: dir.
void xrtref_kat_light_ctor(int kind, const float* a, const float* b, const float* c, const float* l2w16, float* out9)
{
    spdlog::set_level(spdlog::level::off);
    Matrix44f m(l2w16[0], l2w16[1], l2w16[2], l2w16[3], l2w16[4], l2w16[5], l2w16[6], l2w16[7], l2w16[8], l2w16[9], l2w16[10],
                l2w16[11], l2w16[12], l2w16[13], l2w16[14], l2w16[15]);
    auto put = [&](int k, const Vec3f& v) { out9[3 * k] = v[0]; out9[3 * k + 1] = v[1]; out9[3 * k + 2] = v[2]; };
    for (int k = 0; k < 9; ++k) out9[k] = 0.f;
    if (kind == 0) { QuadLight L(v3(a), v3(b), v3(c), m, Vec3f(1.0f)); put(0, L.v0_); put(1, L.v1_); put(2, L.v2_); }
    else if (kind == 1) { TriangleLight L(v3(a), v3(b), v3(c), m, Vec3f(1.0f)); put(0, L.v0_); put(1, L.v1_); put(2, L.v2_); }
    else if (kind == 2) { SphereLight L(v3(a), 1.0f, m, Vec3f(1.0f)); put(0, L.center_); }
    else if (kind == 3) { PointLight L(m, Vec3f(1.0f), 1.0f); put(0, L.pos); }
    else { DistantLight L(m, Vec3f(1.0f), 1.0f); put(0, L.dir); }
}

float xrtref_kat_camera_scale(float fov)
{
    spdlog::set_level(spdlog::level::off);
    PinholeCamera cam(1.0f, Matrix44f(), fov);
    return cam.scale;
}

} // extern "C"
#pragma GCC visibility pop
