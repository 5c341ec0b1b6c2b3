// xrt/light.h — DeltaLight / DistantLight / PointLight and AreaLight / TriangleLight / QuadLight /
// SphereLight of the drop-in API (reference light.h:11-209, light.cpp:6-142). Constructors perform the
// same lightToWorld transforms as the reference so the flattened world-space data is bit-identical;
// sampling (light.cpp:21-30, 59-68; light.h:158-197) runs on the GPU.
#pragma once
#include "primitive.h"

class DeltaLight {
public:
    DeltaLight(const Matrix44f& l2w, const Vec3f& c = 1, const float& i = 1) : color(c), intensity(i), lightToWorld(l2w) {}
    virtual ~DeltaLight() {}
    virtual void describe(xrtg_delta_light& out) const = 0;
    Vec3f color;
    float intensity;
    Matrix44f lightToWorld;
};

// default direction (0,0,-1) (light.cpp:130-134)
class DistantLight : public DeltaLight {
    Vec3f dir;

public:
    DistantLight(const Matrix44f& l2w, const Vec3f& c = 1, const float& i = 1) : DeltaLight(l2w, c, i)
    {
        l2w.multDirMatrix(Vec3f(0, 0, -1), dir);
        dir = normalize(dir);
    }
    Vec3f direction() const { return dir; }
    void describe(xrtg_delta_light& out) const override
    {
        out.kind = XRTG_DLIGHT_DISTANT;
        const Vec3f L = color * intensity;
        for (int a = 0; a < 3; ++a) { out.pos_or_dir[a] = dir[a]; out.radiance[a] = L[a]; }
    }
};

// default position (0,0,0) (light.cpp:115-118)
class PointLight : public DeltaLight {
    Vec3f pos;

public:
    PointLight(const Matrix44f& l2w, const Vec3f& c = 1, const float& i = 100.0) : DeltaLight(l2w, c, i)
    {
        l2w.multVecMatrix(Vec3f(0), pos);
    }
    Vec3f position() const { return pos; }
    void describe(xrtg_delta_light& out) const override
    {
        out.kind = XRTG_DLIGHT_POINT;
        const Vec3f L = color * intensity;
        for (int a = 0; a < 3; ++a) { out.pos_or_dir[a] = pos[a]; out.radiance[a] = L[a]; }
    }
};

class AreaLight {
public:
    AreaLight(const Matrix44f& l2w, const Vec3f& Le) : Le_(Le), lightToWorld(l2w) {}
    virtual ~AreaLight() = default;
    // proxy geometry with material=nullptr, light=this (light.cpp:32-41, 70-82, 92-96)
    virtual std::unique_ptr<Object> makeObject() = 0;
    virtual void describe(xrtg_area_light& out) const = 0;
    const Vec3f& radiance() const { return Le_; }

protected:
    void fill(xrtg_area_light& out, int kind, const Vec3f& a, const Vec3f& b, const Vec3f& c, float r) const
    {
        out.kind = kind; out.radius = r;
        for (int k = 0; k < 3; ++k) { out.v0[k] = a[k]; out.v1[k] = b[k]; out.v2[k] = c[k]; out.Le[k] = Le_[k]; }
    }
    Vec3f Le_;
    Matrix44f lightToWorld;
};

class TriangleLight : public AreaLight {
public:
    TriangleLight(const Vec3f& v0, const Vec3f& v1, const Vec3f& v2, const Matrix44f& l2w, const Vec3f& Le)
        : AreaLight(l2w, Le), v0_(multVecMatrix(v0, l2w)), v1_(multVecMatrix(v1, l2w)), v2_(multVecMatrix(v2, l2w)) {}
    std::unique_ptr<Object> makeObject() override
    {
        const Vec3f n = normalize(cross(v1_ - v0_, v2_ - v0_));
        std::vector<Primitive> prims{Primitive({v0_, v1_, v2_}, {n, n, n}, {Vec2f(0, 0), Vec2f(1, 0), Vec2f(0, 1)})};
        return std::make_unique<Mesh>(std::move(prims), nullptr, this);
    }
    void describe(xrtg_area_light& out) const override { fill(out, XRTG_LIGHT_TRIANGLE, v0_, v1_, v2_, 0.0f); }

private:
    Vec3f v0_, v1_, v2_;
};

// parallelogram v0 + s*(v1-v0) + t*(v2-v0); proxy = triangles (v0,v1,v2) and (v1,v3,v2) (light.cpp:70-82)
class QuadLight : public AreaLight {
public:
    QuadLight(const Vec3f& v0, const Vec3f& v1, const Vec3f& v2, const Matrix44f& l2w, const Vec3f& Le)
        : AreaLight(l2w, Le), v0_(multVecMatrix(v0, l2w)), v1_(multVecMatrix(v1, l2w)), v2_(multVecMatrix(v2, l2w)) {}
    std::unique_ptr<Object> makeObject() override
    {
        const Vec3f e1 = v1_ - v0_, e2 = v2_ - v0_;
        const Vec3f v3 = v0_ + e1 + e2;
        const Vec3f n = normalize(cross(e1, e2));
        const std::vector<Vec2f> uv{Vec2f(0, 0), Vec2f(1, 0), Vec2f(0, 1)};
        std::vector<Primitive> prims{Primitive({v0_, v1_, v2_}, {n, n, n}, uv), Primitive({v1_, v3, v2_}, {n, n, n}, uv)};
        return std::make_unique<Mesh>(std::move(prims), nullptr, this);
    }
    void describe(xrtg_area_light& out) const override { fill(out, XRTG_LIGHT_QUAD, v0_, v1_, v2_, 0.0f); }

private:
    Vec3f v0_, v1_, v2_;
};

// proxy = analytic Sphere (light.cpp:92-96)
class SphereLight : public AreaLight {
public:
    SphereLight(const Vec3f& center, float radius, const Matrix44f& l2w, const Vec3f& Le)
        : AreaLight(l2w, Le), center_(multVecMatrix(center, l2w)), radius_(radius) {}
    std::unique_ptr<Object> makeObject() override { return std::make_unique<Sphere>(center_, radius_, nullptr, this); }
    void describe(xrtg_area_light& out) const override { fill(out, XRTG_LIGHT_SPHERE, center_, Vec3f(0), Vec3f(0), radius_); }

private:
    Vec3f center_;
    float radius_;
};
