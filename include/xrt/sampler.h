// xrt/sampler.h — Sampler / UniformSampler of the drop-in API (reference sampler.h:8-50, sampler.cpp:3-12).
// The host keeps the type only so `render(scene, Sampler::SamplerType::Uniform, image)` reads exactly as in
// the reference; the sample stream itself is produced on the GPU (counter RNG, or the per-pixel mt19937
// stream with GpuOptions::exact).
#pragma once
#include <memory>
#include <random>
#include "geometry.h"

class Sampler {
public:
    enum class SamplerType { Uniform };
    Sampler() {}
    explicit Sampler(uint32_t seed) : gen(seed) {}
    virtual ~Sampler() = default;
    static std::unique_ptr<Sampler> makeSampler(SamplerType st);
    void setSeed(uint32_t seed) { gen.seed(seed); }
    void discard(unsigned long long n) { gen.discard(n); }
    virtual float getNext1D() = 0;
    virtual Vec2f getNext2D() = 0;

protected:
    std::mt19937 gen;
};

class UniformSampler : public Sampler {
public:
    UniformSampler() : dis(0.0f, 1.0f) {}
    explicit UniformSampler(uint32_t seed) : Sampler(seed), dis(0.0f, 1.0f) {}
    float getNext1D() override { return dis(gen); }
    Vec2f getNext2D() override { const float a = dis(gen); const float b = dis(gen); return Vec2f(a, b); }

private:
    std::uniform_real_distribution<float> dis;
};

inline std::unique_ptr<Sampler> Sampler::makeSampler(SamplerType st)
{
    if (st == SamplerType::Uniform) return std::make_unique<UniformSampler>();
    return nullptr;
}
