// xrt/material.h — Material / Lambert of the drop-in API (reference material.h:6-77). f = albedo/PI with
// UNIFORM hemisphere sampling (pdf 1/2PI) is evaluated on the GPU (kernels: shade).
#pragma once
#include "geometry.h"
#include <xrtgpu.h>

class Material {
public:
    Material() = default;
    virtual ~Material() = default;
    virtual MaterialType materialType() const = 0;
    virtual bool describe(xrtg_material& out) const = 0;
};

class Lambert : public Material {
public:
    Lambert(Vec3f albedo) : m_albedo(albedo) {}
    MaterialType materialType() const override { return MaterialType::Lambert; }
    const Vec3f& albedo() const { return m_albedo; }
    bool describe(xrtg_material& out) const override
    {
        out.kind = XRTG_MAT_LAMBERT;
        for (int a = 0; a < 3; ++a) out.albedo[a] = m_albedo[a];
        return true;
    }

private:
    Vec3f m_albedo = Vec3f(0.0f);
};
