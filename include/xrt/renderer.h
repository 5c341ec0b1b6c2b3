// xrt/renderer.h — Renderer plug-in point of the drop-in API (reference renderer.h:8-47) and the GPU
// sibling of the reference's NormalRenderer / ParallelRenderer:
//
//     auto renderer = std::make_unique<GpuRenderer>(n_samples, camera.get(), integrator.get());
//     renderer->render(scene, Sampler::SamplerType::Uniform, image);
//
// On return `image` holds the MEAN radiance exactly like the CPU renderers leave it (sum / spp with
// dropped NaN/inf/negative samples still counted in the divisor, renderer.cpp:57-73,98).
// There is no CPU fallback: any failure throws std::runtime_error carrying xrtg_last_error().
#pragma once
#include "camera.h"
#include "image.h"
#include "integrator.h"
#include "sampler.h"
#include "scene.h"
#include <xrtgpu.h>

class Renderer {
public:
    Renderer(Camera* cam, Integrator* inte) : camera(cam), integrator(inte) {}
    virtual ~Renderer() = default;
    virtual void render(const Scene& scene, Sampler::SamplerType st, Image& image) const = 0;

protected:
    const Camera* camera;
    const Integrator* integrator;
};

struct GpuOptions {
    int device = 0;
    uint32_t seed = 0;       // counter-RNG seed
    bool exact = false;      // reproduce the reference's per-pixel mt19937 sample stream (slow, for parity)
    bool counters = false;   // collect BVH node / triangle / tracking-step counters
    int samplesPerWave = 0;  // 0 = auto
};

class GpuRenderer : public Renderer {
public:
    GpuRenderer(uint32_t spp, Camera* cam, Integrator* inte, GpuOptions opt = GpuOptions());
    ~GpuRenderer() override;
    void render(const Scene& scene, Sampler::SamplerType st, Image& image) const override;
    // statistics of the last render() (ray counts, dropped samples, device milliseconds)
    const xrtg_stats& lastStats() const { return m_stats; }

private:
    const uint32_t n_samples;
    GpuOptions m_opt;
    // device scene cached across render() calls, keyed on (Scene*, Scene::version())
    mutable xrtg_scene* m_scene = nullptr;
    mutable const Scene* m_cachedFor = nullptr;
    mutable uint64_t m_cachedVersion = 0;
    mutable xrt::FlatScene m_flat;
    mutable xrtg_stats m_stats{};
};
