// xrt/primitive.h — Primitive / Object / Sphere / Mesh / SphereMesh / BoxMesh of the drop-in API
// (reference primitive.h:6-273, primitive.cpp:170-205). Host objects are descriptions: intersection
// (Moeller-Trumbore primitive.cpp:140-168, sphere quadratic primitive.h:133-177, box slabs
// primitive.h:243-264) runs on the GPU. Each object can flatten itself into the C-ABI arrays.
#pragma once
#include <memory>
#include <vector>
#include "ray.h"
#include <xrtgpu.h>

class Primitive {
public:
    Primitive(const std::vector<Vec3f>& vertices, const std::vector<Vec3f>& normals, const std::vector<Vec2f>& texcoords)
        : m_vertices(vertices), m_normals(normals), m_texcoords(texcoords) {}
    const std::vector<Vec3f>& vertices() const { return m_vertices; }
    const std::vector<Vec3f>& normals() const { return m_normals; }
    const std::vector<Vec2f>& texcoords() const { return m_texcoords; }

private:
    std::vector<Vec3f> m_vertices, m_normals;
    std::vector<Vec2f> m_texcoords;
};

namespace xrt {
// Growing arrays behind an xrtg_scene_desc.
struct FlatGeometry {
    std::vector<xrtg_triangle> triangles;
    std::vector<xrtg_sphere> spheres;
    std::vector<xrtg_box> boxes;
};
} // namespace xrt

class Material;
class AreaLight;
class Medium;
class Object {
public:
    Object(Material* material, AreaLight* light, Medium* medium) : m_material(material), m_areaLight(light), m_medium(medium) {}
    virtual ~Object() = default;
    bool hasSurface() const { return m_material != nullptr; }
    bool hasAreaLight() const { return m_areaLight != nullptr; }
    bool hasMedium() const { return m_medium != nullptr; }
    // additive accessors (the reference keeps these protected, primitive.h:92-94)
    const Material* material() const { return m_material; }
    const AreaLight* areaLight() const { return m_areaLight; }
    const Medium* medium() const { return m_medium; }
    // appends this object's geometry and fills kind/first/count of `rec`
    virtual void flatten(xrt::FlatGeometry& geo, xrtg_object& rec) const = 0;

protected:
    Material* m_material = nullptr;
    AreaLight* m_areaLight = nullptr;
    Medium* m_medium = nullptr;
};

class Sphere : public Object {
public:
    Sphere(Vec3f center, float radius, Material* material, AreaLight* light = nullptr)
        : Object(material, light, nullptr), m_center(center), m_radius(radius) {}
    const Vec3f& center() const { return m_center; }
    float radius() const { return m_radius; }
    void flatten(xrt::FlatGeometry& geo, xrtg_object& rec) const override
    {
        rec.kind = XRTG_OBJ_SPHERE; rec.first = int32_t(geo.spheres.size()); rec.count = 1;
        geo.spheres.push_back(xrtg_sphere{{m_center[0], m_center[1], m_center[2]}, m_radius});
    }

private:
    Vec3f m_center;
    float m_radius;
};

class Mesh : public Object {
public:
    Mesh(Material* material, AreaLight* light) : Object(material, light, nullptr) {}
    Mesh(const std::vector<Primitive>& primitives, Material* material, AreaLight* light = nullptr)
        : Object(material, light, nullptr), m_primitives(primitives) {}
    Mesh(std::vector<Primitive>&& primitives, Material* material, AreaLight* light = nullptr)
        : Object(material, light, nullptr), m_primitives(std::move(primitives)) {}
    const std::vector<Primitive>& primitives() const { return m_primitives; }
    void flatten(xrt::FlatGeometry& geo, xrtg_object& rec) const override
    {
        rec.kind = XRTG_OBJ_MESH; rec.first = int32_t(geo.triangles.size()); rec.count = int32_t(m_primitives.size());
        for (const auto& p : m_primitives) {
            xrtg_triangle t;
            for (int a = 0; a < 3; ++a) {
                t.v0[a] = p.vertices()[0][a]; t.v1[a] = p.vertices()[1][a]; t.v2[a] = p.vertices()[2][a];
                t.n0[a] = p.normals()[0][a]; t.n1[a] = p.normals()[1][a]; t.n2[a] = p.normals()[2][a];
            }
            geo.triangles.push_back(t);
        }
    }

protected:
    std::vector<Primitive> m_primitives;
};

// UV-sphere tessellation, two triangles per (theta,phi) cell, poles included as degenerate slivers
// (reference primitive.cpp:170-205).
class SphereMesh : public Mesh {
public:
    SphereMesh(Vec3f center, float radius, int thetaResolution, int phiResolution, Material* mt, AreaLight* light)
        : Mesh(mt, light), center_(center), radius_(radius), num_theta_(thetaResolution), num_phi_(phiResolution)
    {
        Triangulate();
    }

private:
    void Triangulate()
    {
        std::vector<Vec3f> pos, nrm;
        for (int i = 0; i <= num_theta_; ++i) {
            const float theta = PI * i / num_theta_;
            for (int j = 0; j <= num_phi_; ++j) {
                const float phi = 2 * PI * j / num_phi_;
                // the reference's unqualified sin()/cos() resolve to the C double overloads: products are formed in double
                const double st = std::sin(double(theta)), ct = std::cos(double(theta));
                const Vec3f n(float(st * std::sin(double(phi))), float(ct), float(st * std::cos(double(phi))));
                pos.push_back(center_ + radius_ * n);
                nrm.push_back(n);
            }
        }
        const std::vector<Vec2f> uv{Vec2f(0, 0), Vec2f(1, 0), Vec2f(0, 1)};
        for (int i = 0; i < num_theta_; ++i) {
            for (int j = 0; j < num_phi_; ++j) {
                const int a = i * (num_phi_ + 1) + j, b = a + num_phi_ + 1;
                m_primitives.emplace_back(std::vector<Vec3f>{pos[a], pos[b], pos[a + 1]}, std::vector<Vec3f>{nrm[a], nrm[b], nrm[a + 1]}, uv);
                m_primitives.emplace_back(std::vector<Vec3f>{pos[b], pos[b + 1], pos[a + 1]}, std::vector<Vec3f>{nrm[b], nrm[b + 1], nrm[a + 1]}, uv);
            }
        }
    }
    Vec3f center_ = Vec3f(0.0f);
    float radius_ = 1.0f;
    int num_theta_ = 10, num_phi_ = 10;
};

// Axis-aligned proxy of a participating medium (reference primitive.h:230-273).
class BoxMesh : public Object {
public:
    BoxMesh(AABB box, Medium* medium) : Object(nullptr, nullptr, medium), box(box) {}
    const AABB& bounds() const { return box; }
    void flatten(xrt::FlatGeometry& geo, xrtg_object& rec) const override
    {
        rec.kind = XRTG_OBJ_BOX; rec.first = int32_t(geo.boxes.size()); rec.count = 1;
        geo.boxes.push_back(xrtg_box{{box.pMin[0], box.pMin[1], box.pMin[2]}, {box.pMax[0], box.pMax[1], box.pMax[2]}});
    }

private:
    AABB box;
};
