// gpu_renderer.cpp — GpuRenderer: the third sibling of NormalRenderer / ParallelRenderer
// (reference renderer.h:22-47, renderer.cpp:8-99). Flattens the host Scene once per scene version,
// creates the device scene through the C ABI and renders into the caller's Image.
#include <xrt/renderer.h>
#include <stdexcept>
#include <string>
#include <vector>

GpuRenderer::GpuRenderer(uint32_t spp, Camera* cam, Integrator* inte, GpuOptions opt)
    : Renderer(cam, inte), n_samples(spp), m_opt(opt) {}

GpuRenderer::~GpuRenderer()
{
    if (m_scene) xrtg_scene_destroy(m_scene);
}

void GpuRenderer::render(const Scene& scene, Sampler::SamplerType st, Image& image) const
{
    if (st != Sampler::SamplerType::Uniform) throw std::runtime_error("[GpuRenderer] unsupported sampler type");
    if (!m_scene || m_cachedFor != &scene || m_cachedVersion != scene.version()) {
        if (m_scene) { xrtg_scene_destroy(m_scene); m_scene = nullptr; }
        scene.flatten(m_flat);
        const int n = m_opt.ngpus > 0 ? m_opt.ngpus : xrtg_device_count() - m_opt.device;
        std::vector<int> devices;
        for (int g = 0; g < n; ++g) devices.push_back(m_opt.device + g);
        const int rc = n > 1 ? xrtg_scene_create_multi(&m_flat.desc, n, devices.data(), 0u, &m_scene) : xrtg_scene_create(&m_flat.desc, m_opt.device, &m_scene);
        if (rc != XRTG_OK) throw std::runtime_error(std::string("[GpuRenderer] scene upload failed: ") + xrtg_last_error());
        m_cachedFor = &scene;
        m_cachedVersion = scene.version();
    }
    xrtg_camera cam{};
    if (!camera->describe(cam)) throw std::runtime_error("[GpuRenderer] camera model has no GPU implementation");

    xrtg_render_params p{};
    p.width = int(image.getWidth());
    p.height = int(image.getHeight());
    p.spp = int(n_samples);
    p.sample_offset = 0;
    p.spp_total = int(n_samples);
    p.integrator = integrator->kind();
    p.max_depth = int(integrator->maxDepth());
    p.seed = m_opt.seed;
    p.flags = (m_opt.exact ? XRTG_FLAG_EXACT : 0u) | (m_opt.counters ? XRTG_FLAG_COUNTERS : 0u);
    p.samples_per_wave = m_opt.samplesPerWave;
    if (xrtg_render(m_scene, &cam, &p, image.data(), &m_stats) != XRTG_OK)
        throw std::runtime_error(std::string("[GpuRenderer] render failed: ") + xrtg_last_error());
}
