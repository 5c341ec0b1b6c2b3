// obj_reader.h — minimal Wavefront OBJ/MTL reader for the subset Scene::loadObj consumes
// (reference: scene.cpp:46-154 uses tinyobj::ObjReader with triangulate=true; tinyobjloader is an
// un-vendored, un-pinned dependency of the reference, so its behaviour is restated here and the SAME
// reader feeds the compiled-reference oracle (through oracle/shim/tiny_obj_loader.h) and the GPU path.
//
// Restated tinyobj semantics:
//   * `o` / `g` start a new shape; a shape that ends up with no faces is dropped (otherwise the
//     face-less `light` / `front_wall` shapes of the Cornell box would index m_material[-1] at
//     scene.cpp:152);
//   * v / vn / vt with 1-based or negative (relative) indices, `f a`, `f a/b`, `f a//c`, `f a/b/c`;
//   * polygons are fan-triangulated (0,k,k+1) — "parity unpinned": tinyobj versions differ in how they
//     split quads, which only permutes triangle ids inside a planar face;
//   * `usemtl` selects the current material id (index into the MTL's newmtl order, -1 if unknown);
//   * MTL: newmtl, Ka, Kd, Ks, Ke, Ni, illum; anything else lands in unknown_parameter.
#pragma once
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

namespace xrt {
namespace obj {

struct Index {
    int vertex_index = -1;
    int normal_index = -1;
    int texcoord_index = -1;
};

struct MeshData {
    std::vector<Index> indices;                   // 3 per triangle after triangulation
    std::vector<unsigned char> num_face_vertices; // always 3
    std::vector<int> material_ids;                // per face
};

struct Shape {
    std::string name;
    MeshData mesh;
};

struct MaterialData {
    std::string name;
    float ambient[3] = {0, 0, 0};
    float diffuse[3] = {0, 0, 0};
    float specular[3] = {0, 0, 0};
    float emission[3] = {0, 0, 0};
    float ior = 1.0f;
    int illum = 0;
    std::map<std::string, std::string> unknown_parameter;
};

struct Attrib {
    std::vector<float> vertices;  // xyz
    std::vector<float> normals;   // xyz
    std::vector<float> texcoords; // uv
};

struct Result {
    Attrib attrib;
    std::vector<Shape> shapes;
    std::vector<MaterialData> materials;
    std::string warning;
    std::string error;
};

namespace detail {

inline std::string trim(const std::string& s)
{
    size_t b = s.find_first_not_of(" \t\r\n");
    if (b == std::string::npos) return "";
    size_t e = s.find_last_not_of(" \t\r\n");
    return s.substr(b, e - b + 1);
}

inline int fixIndex(int idx, int n)
{
    if (idx > 0) return idx - 1;
    if (idx < 0) return n + idx >= 0 ? n + idx : -2; // relative index before the first element: invalid (-2), not "absent" (-1)
    return -1;
}

// parses "a", "a/b", "a//c", "a/b/c"
inline Index parseTriple(const std::string& tok, int nv, int nvt, int nvn)
{
    Index out;
    const char* p = tok.c_str();
    char* end = nullptr;
    long a = std::strtol(p, &end, 10);
    out.vertex_index = fixIndex(int(a), nv);
    if (*end != '/') return out;
    p = end + 1;
    if (*p != '/') {
        long b = std::strtol(p, &end, 10);
        out.texcoord_index = fixIndex(int(b), nvt);
        if (*end != '/') return out;
        p = end + 1;
    }
    else {
        p = p + 1;
    }
    long c = std::strtol(p, &end, 10);
    if (end != p) out.normal_index = fixIndex(int(c), nvn);
    return out;
}

inline bool loadMtl(const std::string& path, std::vector<MaterialData>& mats, std::map<std::string, int>& byName,
                    std::string& warn)
{
    std::ifstream in(path);
    if (!in) {
        warn += "material file not found: " + path + "\n";
        return false;
    }
    std::string line;
    MaterialData* cur = nullptr;
    while (std::getline(in, line)) {
        line = trim(line);
        if (line.empty() || line[0] == '#') continue;
        std::istringstream ss(line);
        std::string key;
        ss >> key;
        if (key == "newmtl") {
            std::string name;
            std::getline(ss, name);
            mats.emplace_back();
            cur = &mats.back();
            cur->name = trim(name);
            byName[cur->name] = int(mats.size()) - 1;
            continue;
        }
        if (!cur) continue;
        auto read3 = [&](float* dst) {
            float x = 0, y = 0, z = 0;
            ss >> x;
            if (ss >> y) { ss >> z; }
            else { y = z = x; }
            dst[0] = x; dst[1] = y; dst[2] = z;
        };
        if (key == "Ka") read3(cur->ambient);
        else if (key == "Kd") read3(cur->diffuse);
        else if (key == "Ks") read3(cur->specular);
        else if (key == "Ke") read3(cur->emission);
        else if (key == "Ni") ss >> cur->ior;
        else if (key == "illum") ss >> cur->illum;
        else {
            std::string rest;
            std::getline(ss, rest);
            cur->unknown_parameter[key] = trim(rest);
        }
    }
    return true;
}

} // namespace detail

// Parse `path`; `mtlDir` is where `mtllib` files are searched (scene.cpp:53 passes the OBJ's parent).
inline bool load(const std::string& path, const std::string& mtlDir, Result& out)
{
    std::ifstream in(path);
    if (!in) {
        out.error = "cannot open " + path;
        return false;
    }
    std::map<std::string, int> matByName;
    Shape cur;
    int curMat = -1;
    auto flush = [&]() {
        if (!cur.mesh.num_face_vertices.empty()) out.shapes.push_back(cur);
        cur.mesh = MeshData();
    };
    std::string line;
    std::vector<Index> face;
    while (std::getline(in, line)) {
        line = detail::trim(line);
        if (line.empty() || line[0] == '#') continue;
        std::istringstream ss(line);
        std::string key;
        ss >> key;
        if (key == "v") {
            float x = 0, y = 0, z = 0;
            ss >> x >> y >> z;
            out.attrib.vertices.insert(out.attrib.vertices.end(), {x, y, z});
        }
        else if (key == "vn") {
            float x = 0, y = 0, z = 0;
            ss >> x >> y >> z;
            out.attrib.normals.insert(out.attrib.normals.end(), {x, y, z});
        }
        else if (key == "vt") {
            float u = 0, v = 0;
            ss >> u >> v;
            out.attrib.texcoords.insert(out.attrib.texcoords.end(), {u, v});
        }
        else if (key == "f") {
            face.clear();
            std::string tok;
            const int nv = int(out.attrib.vertices.size() / 3);
            const int nvt = int(out.attrib.texcoords.size() / 2);
            const int nvn = int(out.attrib.normals.size() / 3);
            while (ss >> tok) face.push_back(detail::parseTriple(tok, nv, nvt, nvn));
            if (face.size() < 3) {
                out.warning += "degenerate face ignored\n";
                continue;
            }
            // every resolved index must name an existing attribute: a position is mandatory (0, missing or out of range is a
            // load error, like tinyobj's), texcoord / normal indices are optional (-1) but, when given, must exist too
            for (const Index& ix : face) {
                if (ix.vertex_index < 0 || ix.vertex_index >= nv) {
                    out.error = "face references vertex " + std::to_string(ix.vertex_index + 1) + " but only " + std::to_string(nv) + " are defined: " + line;
                    return false;
                }
                if (ix.texcoord_index >= nvt || ix.texcoord_index < -1 || ix.normal_index >= nvn || ix.normal_index < -1) {
                    out.error = "face references a texture coordinate / normal that is not defined: " + line;
                    return false;
                }
            }
            for (size_t k = 1; k + 1 < face.size(); ++k) {
                cur.mesh.indices.push_back(face[0]);
                cur.mesh.indices.push_back(face[k]);
                cur.mesh.indices.push_back(face[k + 1]);
                cur.mesh.num_face_vertices.push_back(3);
                cur.mesh.material_ids.push_back(curMat);
            }
        }
        else if (key == "o" || key == "g") {
            flush();
            std::string name;
            std::getline(ss, name);
            cur.name = detail::trim(name);
        }
        else if (key == "usemtl") {
            std::string name;
            std::getline(ss, name);
            name = detail::trim(name);
            auto it = matByName.find(name);
            if (it == matByName.end()) {
                out.warning += "material not found: " + name + "\n";
                curMat = -1;
            }
            else {
                curMat = it->second;
            }
        }
        else if (key == "mtllib") {
            std::string name;
            std::getline(ss, name);
            name = detail::trim(name);
            std::string full = mtlDir.empty() ? name : (mtlDir + "/" + name);
            detail::loadMtl(full, out.materials, matByName, out.warning);
        }
        // s, l, p and everything else: ignored
    }
    flush();
    return true;
}

} // namespace obj
} // namespace xrt
