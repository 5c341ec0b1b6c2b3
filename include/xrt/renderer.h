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
    int device = 0;          // first device
    int ngpus = 1;           // devices device .. device+ngpus-1 behind one handle: the samples of a render() are split across them and
                             // the per-device sums meet in one fused peer-memory reduce + `image /= spp` kernel (xrtg_scene_create_multi);
                             // 0 = every visible device. Exact renders (one mt19937 stream per pixel) stay on the first device.
    uint32_t seed = 0;       // counter-RNG seed
    bool exact = false;      // reproduce the reference's per-pixel mt19937 sample stream (slow, for parity)
    bool counters = false;   // collect BVH node / triangle / tracking-step counters
    int samplesPerWave = 0;  // 0 = auto
};

#pragma GCC visibility push(default)
class GpuRenderer : public Renderer {
public:
    GpuRenderer(uint32_t spp, Camera* cam, Integrator* inte, GpuOptions opt = GpuOptions());
    ~GpuRenderer() override;
    void render(const Scene& scene, Sampler::SamplerType st, Image& image) const override;
    // statistics of the last render() (ray counts, dropped samples, device milliseconds)
    const xrtg_stats& lastStats() const { return m_stats; }

protected:
    const uint32_t n_samples;
    GpuOptions m_opt;

private:
    // device scene cached across render() calls, keyed on (Scene*, Scene::version())
    mutable xrtg_scene* m_scene = nullptr;
    mutable const Scene* m_cachedFor = nullptr;
    mutable uint64_t m_cachedVersion = 0;
    mutable xrt::FlatScene m_flat;
    mutable xrtg_stats m_stats{};
};

#pragma GCC visibility pop

// The reference's two renderer names (renderer.h:22-47), kept so that its examples compile and run unchanged
// (`std::make_unique<NormalRenderer>(n_samples, camera.get(), integrator.get())`). The reference's NormalRenderer (serial)
// and ParallelRenderer (PSTL) produce the SAME image — every pixel owns a std::mt19937 seeded with its linear index
// (renderer.cpp:35-36). Here both run on the GPU in exact mode, i.e. they replay that very sample stream with the
// reference's un-fused fp32 arithmetic, so their output matches the CPU renderers' (bit-exact for NormalIntegrator /
// DirectIntegrator, ~1e-7 elsewhere). Use GpuRenderer for the counter-RNG throughput path. Neither has a CPU code path.
class NormalRenderer : public GpuRenderer {
public:
    NormalRenderer(uint32_t spp, Camera* cam, Integrator* inte) : GpuRenderer(spp, cam, inte, exactOptions()) {}

protected:
    static GpuOptions exactOptions() { GpuOptions o; o.exact = true; return o; }
};

class ParallelRenderer : public NormalRenderer {
public:
    ParallelRenderer(uint32_t spp, Camera* cam, Integrator* inte) : NormalRenderer(spp, cam, inte) {}
};
