// scene.cpp — host Scene: OBJ ingest and the flattening boundary (reference scene.cpp:9-188).
#include <xrt/scene.h>
#include <xrt/obj_reader.h>
#include <map>
#include <stdexcept>

// scene.cpp:9-29 — every illum maps to Lambert(Kd); "no_surface" yields no material
static std::unique_ptr<Material> makeMaterial(const xrt::obj::MaterialData& m)
{
    if (m.unknown_parameter.count("no_surface") == 1) return nullptr;
    return std::make_unique<Lambert>(Vec3f(m.diffuse[0], m.diffuse[1], m.diffuse[2]));
}

void Scene::loadObj(const std::filesystem::path& filepath)
{
    xrt::obj::Result r;
    if (!xrt::obj::load(filepath.generic_string(), filepath.parent_path().string(), r))
        throw std::runtime_error("[Scene] failed to load " + filepath.generic_string() + " : " + r.error);

    const size_t matBase = m_material.size();
    for (const auto& m : r.materials) m_material.push_back(makeMaterial(m));

    const auto& A = r.attrib;
    for (const auto& shape : r.shapes) {
        int materialID = -1;
        std::vector<Primitive> prims;
        prims.reserve(shape.mesh.num_face_vertices.size());
        size_t off = 0;
        for (size_t f = 0; f < shape.mesh.num_face_vertices.size(); ++f) {
            const size_t fv = shape.mesh.num_face_vertices[f];
            std::vector<Vec3f> P, N;
            std::vector<Vec2f> T;
            for (size_t v = 0; v < fv; ++v) {
                const auto idx = shape.mesh.indices[off + v];
                P.emplace_back(A.vertices[3 * idx.vertex_index], A.vertices[3 * idx.vertex_index + 1], A.vertices[3 * idx.vertex_index + 2]);
                if (idx.normal_index >= 0)
                    N.emplace_back(A.normals[3 * idx.normal_index], A.normals[3 * idx.normal_index + 1], A.normals[3 * idx.normal_index + 2]);
                if (idx.texcoord_index >= 0) T.emplace_back(A.texcoords[2 * idx.texcoord_index], A.texcoords[2 * idx.texcoord_index + 1]);
            }
            if (N.size() != fv) { // no normals (or normals on only some vertices): flat normal from the winding (scene.cpp:118-125)
                const Vec3f n = normalize(cross(P[1] - P[0], P[2] - P[0]));
                N = {n, n, n};
            }
            if (T.size() != fv) T = {Vec2f(0, 0), Vec2f(1, 0), Vec2f(0, 1)};
            if (materialID == -1) materialID = shape.mesh.material_ids[f]; // first face decides (scene.cpp:135-144)
            prims.emplace_back(P, N, T);
            off += fv;
        }
        if (materialID < 0) throw std::runtime_error("[Scene] shape '" + shape.name + "' has no material (the reference indexes m_material[-1], scene.cpp:152)");
        addObj(shape.name, std::make_unique<Mesh>(std::move(prims), m_material[matBase + materialID].get(), nullptr));
    }
}

void Scene::addObj(std::string name, std::unique_ptr<Object> obj)
{
    if (!m_insertSeq.count(name)) m_insertSeq[name] = m_nextSeq++;
    m_objects[name] = std::move(obj); // same call shape as scene.cpp:158 -> same libstdc++ iteration order
    ++m_version;
}

void Scene::addDeltaLight(std::string, std::unique_ptr<DeltaLight> light)
{
    m_deltaLights.push_back(std::move(light));
    ++m_version;
}

void Scene::addAreaLight(std::string name, std::unique_ptr<AreaLight> light)
{
    addObj(name, light->makeObject());
    m_areaLights.push_back(std::move(light));
}

void Scene::flatten(xrt::FlatScene& out) const
{
    out = xrt::FlatScene();
    std::map<const Material*, int> matIdx;
    std::map<const AreaLight*, int> lightIdx;
    std::map<const Medium*, int> medIdx;
    std::map<const DensityGrid*, int> gridIdx;

    for (const auto& L : m_areaLights) {
        lightIdx[L.get()] = int(out.areaLights.size());
        xrtg_area_light d{};
        L->describe(d);
        out.areaLights.push_back(d);
    }
    for (const auto& L : m_deltaLights) {
        xrtg_delta_light d{};
        L->describe(d);
        out.deltaLights.push_back(d);
    }

    out.names.reserve(m_objects.size());
    for (const auto& [name, obj] : m_objects) { // THE reference iteration order (scene.cpp:193)
        xrtg_object rec{};
        obj->flatten(out.geo, rec);
        rec.material = rec.area_light = rec.medium = -1;
        if (const Material* m = obj->material()) {
            auto it = matIdx.find(m);
            if (it == matIdx.end()) {
                xrtg_material d{};
                if (!m->describe(d)) throw std::runtime_error("[Scene] material of '" + name + "' has no GPU implementation");
                it = matIdx.emplace(m, int(out.materials.size())).first;
                out.materials.push_back(d);
            }
            rec.material = it->second;
        }
        if (const AreaLight* L = obj->areaLight()) {
            auto it = lightIdx.find(L);
            if (it == lightIdx.end()) throw std::runtime_error("[Scene] object '" + name + "' carries an area light that was not added with addAreaLight");
            rec.area_light = it->second;
        }
        if (const Medium* M = obj->medium()) {
            auto it = medIdx.find(M);
            if (it == medIdx.end()) {
                xrtg_medium d{};
                M->describe(d);
                if (const DensityGrid* G = M->grid()) {
                    auto git = gridIdx.find(G);
                    if (git == gridIdx.end()) {
                        xrtg_grid gd{};
                        if (!G->describe(gd)) throw std::runtime_error("[Scene] density grid of '" + name + "' has no GPU implementation");
                        git = gridIdx.emplace(G, int(out.grids.size())).first;
                        out.grids.push_back(gd);
                    }
                    d.grid = git->second;
                }
                it = medIdx.emplace(M, int(out.media.size())).first;
                out.media.push_back(d);
            }
            rec.medium = it->second;
        }
        rec.insert_seq = m_insertSeq.at(name);
        out.names.push_back(name);
        out.objects.push_back(rec);
    }
    for (size_t i = 0; i < out.objects.size(); ++i) out.objects[i].name = out.names[i].c_str();

    xrtg_scene_desc& d = out.desc;
    d = xrtg_scene_desc{};
    d.abi_version = XRTG_ABI_VERSION;
    d.n_objects = int(out.objects.size());       d.objects = out.objects.data();
    d.n_triangles = int(out.geo.triangles.size()); d.triangles = out.geo.triangles.data();
    d.n_spheres = int(out.geo.spheres.size());   d.spheres = out.geo.spheres.data();
    d.n_boxes = int(out.geo.boxes.size());       d.boxes = out.geo.boxes.data();
    d.n_materials = int(out.materials.size());   d.materials = out.materials.data();
    d.n_area_lights = int(out.areaLights.size()); d.area_lights = out.areaLights.data();
    d.n_delta_lights = int(out.deltaLights.size()); d.delta_lights = out.deltaLights.data();
    d.n_media = int(out.media.size());           d.media = out.media.data();
    d.n_grids = int(out.grids.size());           d.grids = out.grids.data();
}
