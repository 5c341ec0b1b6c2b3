// xrt/medium.h — Medium hierarchy of the drop-in API (reference medium.h:71-387, medium.cpp:5-23).
// Free-flight sampling, spectral delta tracking and Henyey-Greenstein sampling run on the GPU; the host
// objects carry the coefficients and manufacture the BoxMesh proxy exactly like the reference.
#pragma once
#include "grid.h"
#include "primitive.h"

class Medium {
public:
    explicit Medium(float g) : g_(g) {}
    virtual ~Medium() = default;
    virtual std::unique_ptr<Object> makeObject() = 0;
    virtual void describe(xrtg_medium& out) const = 0;
    virtual const DensityGrid* grid() const { return nullptr; }
    float g() const { return g_; }

protected:
    void fill(xrtg_medium& m, int kind, const Vec3f& a, const Vec3f& s, float mul) const
    {
        m.kind = kind; m.g = g_; m.density_mul = mul; m.grid = -1;
        for (int k = 0; k < 3; ++k) { m.sigma_a[k] = a[k]; m.sigma_s[k] = s[k]; }
    }
    float g_;
};

class HomogeneousMedium : public Medium {
public:
    HomogeneousMedium(float g, Vec3f a, Vec3f s, AABB box) : Medium(g), sigma_a(a), sigma_s(s), sigma_t(a + s), box(box) {}
    std::unique_ptr<Object> makeObject() override { return std::make_unique<BoxMesh>(box, this); }

protected:
    const Vec3f sigma_a, sigma_s, sigma_t;
    const AABB box;
};

class HomogeneousMediumMIS : public HomogeneousMedium {
public:
    using HomogeneousMedium::HomogeneousMedium;
    void describe(xrtg_medium& m) const override { fill(m, XRTG_MEDIUM_HOMOGENEOUS_MIS, sigma_a, sigma_s, 1.0f); }
};

class HomogeneousMediumAchromatic : public HomogeneousMedium {
public:
    HomogeneousMediumAchromatic(float g, float a, float s, AABB box) : HomogeneousMedium(g, Vec3f(a), Vec3f(s), box) {}
    void describe(xrtg_medium& m) const override { fill(m, XRTG_MEDIUM_HOMOGENEOUS_ACHROMATIC, sigma_a, sigma_s, 1.0f); }
};

class HomogeneousMediumNoMIS : public HomogeneousMedium {
public:
    using HomogeneousMedium::HomogeneousMedium;
    void describe(xrtg_medium& m) const override { fill(m, XRTG_MEDIUM_HOMOGENEOUS_NOMIS, sigma_a, sigma_s, 1.0f); }
};

class HeterogeneousMedium : public Medium {
public:
    HeterogeneousMedium(float g, const DensityGrid* densityGridPtr, const Vec3f& absorptionColor, const Vec3f& scatteringColor,
                        float densityMultiplier = 1.0f)
        : Medium(g), densityGridPtr(densityGridPtr), absorptionColor(absorptionColor), scatteringColor(scatteringColor),
          densityMultiplier(densityMultiplier) {}
    std::unique_ptr<Object> makeObject() override { return std::make_unique<BoxMesh>(densityGridPtr->getBounds(), this); }
    const DensityGrid* grid() const override { return densityGridPtr; }
    void describe(xrtg_medium& m) const override { fill(m, XRTG_MEDIUM_HETEROGENEOUS, absorptionColor, scatteringColor, densityMultiplier); }

private:
    const DensityGrid* densityGridPtr;
    const Vec3f absorptionColor, scatteringColor;
    const float densityMultiplier;
};
