// capi.cpp — a narrow C scripting surface over the host C++ API (Scene / lights / media / camera /
// GpuRenderer) so the Python test-suite and bench.py can assemble scenes with the SAME host code a C++
// user of the drop-in API runs (include/xrt/*.h), then hand the flattened description to the GPU C ABI
// (xrtg_*), to the oracle port (xrto_*) and to the compiled reference harness (xrtref_*).
#include <xrt/renderer.h>
#include <cstring>
#include <stdexcept>
#include <string>

namespace {
thread_local std::string g_err;
Vec3f v3(const float* p) { return Vec3f(p[0], p[1], p[2]); }
Matrix44f m44(const float* m)
{
    if (!m) return Matrix44f();
    return Matrix44f(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], m[9], m[10], m[11], m[12], m[13], m[14], m[15]);
}
} // namespace

struct xrth_scene {
    Scene scene;
    std::vector<std::unique_ptr<Material>> materials; // user-side materials (example.cpp:43 keeps them in main)
    std::vector<std::unique_ptr<Medium>> media;
    std::vector<std::unique_ptr<DensityGrid>> grids;
    xrt::FlatScene flat;
    uint64_t flatVersion = ~0ull;
};

#define XRTH_TRY(...)                                   \
    try { __VA_ARGS__; return 0; }                      \
    catch (const std::exception& e) { g_err = e.what(); return -1; }

#pragma GCC visibility push(default)
extern "C" {

const char* xrth_last_error(void) { return g_err.c_str(); }
xrth_scene* xrth_scene_new(void) { return new xrth_scene(); }
void xrth_scene_free(xrth_scene* s) { delete s; }

int xrth_scene_load_obj(xrth_scene* s, const char* path) { XRTH_TRY(s->scene.loadObj(path)) }

static Material* newLambert(xrth_scene* s, const float* albedo)
{
    if (!albedo) return nullptr;
    s->materials.push_back(std::make_unique<Lambert>(v3(albedo)));
    return s->materials.back().get();
}

// Mesh from n xrtg_triangle records with one Lambert(albedo) material.
int xrth_scene_add_mesh(xrth_scene* s, const char* name, const xrtg_triangle* tris, int n, const float* albedo)
{
    XRTH_TRY({
        std::vector<Primitive> prims;
        prims.reserve(n);
        const std::vector<Vec2f> uv{Vec2f(0, 0), Vec2f(1, 0), Vec2f(0, 1)};
        for (int i = 0; i < n; ++i) {
            const xrtg_triangle& t = tris[i];
            prims.emplace_back(std::vector<Vec3f>{v3(t.v0), v3(t.v1), v3(t.v2)}, std::vector<Vec3f>{v3(t.n0), v3(t.n1), v3(t.n2)}, uv);
        }
        s->scene.addObj(name, std::make_unique<Mesh>(std::move(prims), newLambert(s, albedo), nullptr));
    })
}

int xrth_scene_add_sphere(xrth_scene* s, const char* name, const float* center, float radius, const float* albedo)
{
    XRTH_TRY(s->scene.addObj(name, std::make_unique<Sphere>(v3(center), radius, newLambert(s, albedo), nullptr)))
}

int xrth_scene_add_sphere_mesh(xrth_scene* s, const char* name, const float* center, float radius, int ntheta, int nphi,
                               const float* albedo)
{
    XRTH_TRY(s->scene.addObj(name, std::make_unique<SphereMesh>(v3(center), radius, ntheta, nphi, newLambert(s, albedo), nullptr)))
}

int xrth_scene_add_quad_light(xrth_scene* s, const char* name, const float* v0, const float* v1, const float* v2,
                              const float* l2w, const float* Le)
{
    XRTH_TRY(s->scene.addAreaLight(name, std::make_unique<QuadLight>(v3(v0), v3(v1), v3(v2), m44(l2w), v3(Le))))
}

int xrth_scene_add_triangle_light(xrth_scene* s, const char* name, const float* v0, const float* v1, const float* v2,
                                  const float* l2w, const float* Le)
{
    XRTH_TRY(s->scene.addAreaLight(name, std::make_unique<TriangleLight>(v3(v0), v3(v1), v3(v2), m44(l2w), v3(Le))))
}

int xrth_scene_add_sphere_light(xrth_scene* s, const char* name, const float* center, float radius, const float* l2w,
                                const float* Le)
{
    XRTH_TRY(s->scene.addAreaLight(name, std::make_unique<SphereLight>(v3(center), radius, m44(l2w), v3(Le))))
}

int xrth_scene_add_point_light(xrth_scene* s, const char* name, const float* l2w, const float* color, float intensity)
{
    XRTH_TRY(s->scene.addDeltaLight(name, std::make_unique<PointLight>(m44(l2w), v3(color), intensity)))
}

int xrth_scene_add_distant_light(xrth_scene* s, const char* name, const float* l2w, const float* color, float intensity)
{
    XRTH_TRY(s->scene.addDeltaLight(name, std::make_unique<DistantLight>(m44(l2w), v3(color), intensity)))
}

// kind = xrtg_medium_kind (homogeneous variants only)
int xrth_scene_add_homogeneous_medium(xrth_scene* s, const char* name, int kind, float g, const float* sigma_a,
                                      const float* sigma_s, const float* pmin, const float* pmax)
{
    XRTH_TRY({
        const AABB box{v3(pmin), v3(pmax)};
        std::unique_ptr<Medium> m;
        if (kind == XRTG_MEDIUM_HOMOGENEOUS_MIS) m = std::make_unique<HomogeneousMediumMIS>(g, v3(sigma_a), v3(sigma_s), box);
        else if (kind == XRTG_MEDIUM_HOMOGENEOUS_ACHROMATIC) m = std::make_unique<HomogeneousMediumAchromatic>(g, sigma_a[0], sigma_s[0], box);
        else if (kind == XRTG_MEDIUM_HOMOGENEOUS_NOMIS) m = std::make_unique<HomogeneousMediumNoMIS>(g, v3(sigma_a), v3(sigma_s), box);
        else throw std::runtime_error("not a homogeneous medium kind");
        s->scene.addObj(name, m->makeObject());
        s->media.push_back(std::move(m));
    })
}

// Dense grid (copied) + HeterogeneousMedium(g, grid, absColor, scatColor, mul) + its BoxMesh proxy.
int xrth_scene_add_heterogeneous_medium(xrth_scene* s, const char* name, float g, int nx, int ny, int nz, const float* voxels,
                                        const float* origin, float voxel_size, const float* abs_color, const float* scat_color,
                                        float density_mul)
{
    XRTH_TRY({
        std::vector<float> data(voxels, voxels + size_t(nx) * ny * nz);
        s->grids.push_back(std::make_unique<DenseGrid>(nx, ny, nz, std::move(data), v3(origin), voxel_size, 0.0f));
        auto m = std::make_unique<HeterogeneousMedium>(g, s->grids.back().get(), v3(abs_color), v3(scat_color), density_mul);
        s->scene.addObj(name, m->makeObject());
        s->media.push_back(std::move(m));
    })
}

// Flattened description (cached per scene version). Pointers stay valid until the scene is mutated or freed.
const xrtg_scene_desc* xrth_scene_flatten(xrth_scene* s)
{
    try {
        if (s->flatVersion != s->scene.version()) {
            s->scene.flatten(s->flat);
            s->flatVersion = s->scene.version();
        }
        return &s->flat.desc;
    }
    catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}

// PinholeCamera(aspect, c2w, fov) -> xrtg_camera (scale = tan(FOV/2) evaluated as camera.h:44 does)
int xrth_camera_make(float aspect, const float* c2w, float fov_deg, xrtg_camera* out)
{
    XRTH_TRY({
        PinholeCamera cam(aspect, m44(c2w), fov_deg);
        cam.describe(*out);
    })
}

// The user-facing call: GpuRenderer(spp, camera, integrator).render(scene, Uniform, image).
// rgb = W*H*3 floats (host). flags: XRTG_FLAG_EXACT / XRTG_FLAG_COUNTERS.
int xrth_render(xrth_scene* s, float aspect, const float* c2w, float fov_deg, int integrator, int max_depth, int spp, int width,
                int height, uint32_t seed, uint32_t flags, float* rgb, xrtg_stats* stats)
{
    XRTH_TRY({
        PinholeCamera cam(aspect, m44(c2w), fov_deg);
        std::unique_ptr<Integrator> integ;
        switch (integrator) {
        case XRTG_INT_NORMAL: integ = std::make_unique<NormalIntegrator>(); break;
        case XRTG_INT_FURNACE: integ = std::make_unique<FurnaceIntegrator>(); break;
        case XRTG_INT_DIRECT: integ = std::make_unique<DirectIntegrator>(); break;
        case XRTG_INT_INDIRECT: integ = std::make_unique<IndirectIntegrator>(max_depth); break;
        case XRTG_INT_GI: integ = std::make_unique<GIIntegrator>(max_depth); break;
        case XRTG_INT_WHITTED: integ = std::make_unique<WhittedIntegrator>(max_depth); break;
        case XRTG_INT_VOLUME: integ = std::make_unique<VolumePathTracing>(max_depth); break;
        case XRTG_INT_VOLUME_NEE: integ = std::make_unique<VolumePathTracingNEE>(max_depth); break;
        default: throw std::runtime_error("unknown integrator");
        }
        GpuOptions opt;
        opt.seed = seed;
        opt.exact = (flags & XRTG_FLAG_EXACT) != 0;
        opt.counters = (flags & XRTG_FLAG_COUNTERS) != 0;
        GpuRenderer r(uint32_t(spp), &cam, integ.get(), opt);
        Image image(width, height);
        r.render(s->scene, Sampler::SamplerType::Uniform, image);
        std::memcpy(rgb, image.data(), sizeof(float) * 3 * size_t(width) * height);
        if (stats) *stats = r.lastStats();
    })
}

} // extern "C"
#pragma GCC visibility pop
