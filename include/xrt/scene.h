// xrt/scene.h — Scene of the drop-in API (reference scene.h:13-47, scene.cpp:46-188) plus the flattening
// boundary: Scene::flatten() emits the C-ABI description with objects in the SAME order the reference's
// Scene::intersect walks them (std::unordered_map iteration order, scene.cpp:193) so global primitive
// ids and closest-hit tie-breaks match. Closest-hit / any-hit themselves run on the GPU over a SAH BVH
// built inside xrtg_scene_create — the role the reference leaves to its empty hook Scene::build()
// (scene.h:22-24).
#pragma once
#include <filesystem>
#include <string>
#include <unordered_map>
#include <vector>
#include "light.h"
#include "material.h"
#include "medium.h"
#include "primitive.h"
#include <xrtgpu.h>

namespace xrt {
// Owns every array an xrtg_scene_desc points to.
struct FlatScene {
    FlatGeometry geo;
    std::vector<xrtg_object> objects;
    std::vector<std::string> names;
    std::vector<xrtg_material> materials;
    std::vector<xrtg_area_light> areaLights;
    std::vector<xrtg_delta_light> deltaLights;
    std::vector<xrtg_medium> media;
    std::vector<xrtg_grid> grids;
    xrtg_scene_desc desc;
};
} // namespace xrt

class Sampler;
#pragma GCC visibility push(default) // libxrthost.so is built with hidden visibility; the C++ API is exported explicitly
class Scene {
public:
    ~Scene() = default;

    // Wavefront OBJ -> one Mesh per shape, Lambert(Kd) per MTL material, flat normals if the file has
    // none (scene.cpp:46-154). Throws std::runtime_error where the reference calls exit(1).
    void loadObj(const std::filesystem::path& filepath);
    void addObj(std::string name, std::unique_ptr<Object> obj);
    void build() { ++m_version; }
    void addDeltaLight(std::string name, std::unique_ptr<DeltaLight> light);
    // also inserts the light's proxy object under `name` (scene.cpp:166-170)
    void addAreaLight(std::string name, std::unique_ptr<AreaLight> light);
    const std::vector<std::unique_ptr<DeltaLight>>& getDeltaLights() const { return m_deltaLights; }
    const std::vector<std::unique_ptr<AreaLight>>& getAreaLights() const { return m_areaLights; }

    // ---- additive: the flattening boundary ----
    // Fills `out`; pointers inside out.desc stay valid while `out` and this Scene live unchanged.
    // Throws std::runtime_error on content the GPU path cannot represent.
    void flatten(xrt::FlatScene& out) const;
    // bumped by every mutation; GpuRenderer keys its device-scene cache on (Scene*, version)
    uint64_t version() const { return m_version; }

private:
    std::vector<std::unique_ptr<DeltaLight>> m_deltaLights;
    std::vector<std::unique_ptr<AreaLight>> m_areaLights;
    std::unordered_map<std::string, std::unique_ptr<Object>> m_objects;
    std::vector<std::unique_ptr<Material>> m_material;
    std::unordered_map<std::string, int> m_insertSeq; // addObj order, replayed by the oracle harness
    int m_nextSeq = 0;
    uint64_t m_version = 0;
};
#pragma GCC visibility pop
