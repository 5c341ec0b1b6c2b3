// cornellbox_gpu.cpp — the Cornell-box example of the reference (examples/cornellbox.cpp:19-77) on the GPU renderer.
// Same scene assembly calls, same camera, same light; only the renderer line differs. Usage:
//   cornellbox_gpu out.ppm [width height spp] [obj | -] [ngpus]
// Scene geometry comes from an OBJ file when given, else ("-" or absent) the box is built from quads in code. ngpus > 1 (0 = every
// visible device) renders on several GPUs from this one process: GpuOptions::ngpus -> xrtg_scene_create_multi, no Python, no NCCL.
#include <xrt/renderer.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <exception>
#include <memory>
#include <string>

static void addQuad(std::vector<Primitive>& prims, Vec3f a, Vec3f b, Vec3f c, Vec3f d)
{
    const Vec3f n = normalize(cross(b - a, c - a));
    const std::vector<Vec2f> uv{Vec2f(0, 0), Vec2f(1, 0), Vec2f(0, 1)};
    prims.emplace_back(std::vector<Vec3f>{a, b, c}, std::vector<Vec3f>{n, n, n}, uv);
    prims.emplace_back(std::vector<Vec3f>{a, c, d}, std::vector<Vec3f>{n, n, n}, uv);
}

int main(int argc, char** argv)
{
    const std::string out = argc > 1 ? argv[1] : "cornellbox_gpu.ppm";
    const uint32_t width = argc > 3 ? std::atoi(argv[2]) : 780;
    const uint32_t height = argc > 3 ? std::atoi(argv[3]) : 585;
    const uint32_t n_samples = argc > 4 ? std::atoi(argv[4]) : 16;
    const uint32_t max_depth = 3;

    Image image(width, height);
    const float aspect_ratio = static_cast<float>(width) / height;
    const Matrix44f c2w(-1.0, 0, 0, 0, 0, 1.0, 0, 0, 0, 0, -1.0, 0, 278, 274.4, -750.0, 1);
    const auto camera = std::make_unique<PinholeCamera>(aspect_ratio, c2w, 60.0f);

    Scene scene;
    const auto white = std::make_unique<Lambert>(Vec3f(1, 1, 1));
    const auto red = std::make_unique<Lambert>(Vec3f(1, 0, 0));
    const auto green = std::make_unique<Lambert>(Vec3f(0, 1, 0));
    try {
        if (argc > 5 && std::string(argv[5]) != "-") scene.loadObj(argv[5]);
        else {
            std::vector<Primitive> w, r, g;
            addQuad(w, Vec3f(552.8, 0, 0), Vec3f(0, 0, 0), Vec3f(0, 0, 559.2), Vec3f(549.6, 0, 559.2));            // floor
            addQuad(w, Vec3f(556, 548.8, 0), Vec3f(556, 548.8, 559.2), Vec3f(0, 548.8, 559.2), Vec3f(0, 548.8, 0)); // ceiling
            addQuad(w, Vec3f(549.6, 0, 559.2), Vec3f(0, 0, 559.2), Vec3f(0, 548.8, 559.2), Vec3f(556, 548.8, 559.2)); // back
            addQuad(g, Vec3f(0, 0, 559.2), Vec3f(0, 0, 0), Vec3f(0, 548.8, 0), Vec3f(0, 548.8, 559.2));
            addQuad(r, Vec3f(552.8, 0, 0), Vec3f(549.6, 0, 559.2), Vec3f(556, 548.8, 559.2), Vec3f(556, 548.8, 0));
            scene.addObj("walls", std::make_unique<Mesh>(w, white.get()));
            scene.addObj("green_wall", std::make_unique<Mesh>(g, green.get()));
            scene.addObj("red_wall", std::make_unique<Mesh>(r, red.get()));
            scene.addObj("ball", std::make_unique<SphereMesh>(Vec3f(278, 120, 280), 120.0f, 64, 64, white.get(), nullptr));
        }
        scene.addAreaLight("QuadLight", std::make_unique<QuadLight>(Vec3f(343.0, 548.0, 227.0), Vec3f(343.0, 548.0, 332.0),
                                                                    Vec3f(213.0, 548.0, 227.0), Matrix44f(), 25.0f * Vec3f(1.0, 1.0, 1.0)));
        scene.build();

        const auto integrator = std::make_unique<GIIntegrator>(max_depth);
        GpuOptions opt;
        opt.ngpus = argc > 6 ? std::atoi(argv[6]) : 1;
        auto renderer = std::make_unique<GpuRenderer>(n_samples, camera.get(), integrator.get(), opt);
        renderer->render(scene, UniformSampler::SamplerType::Uniform, image); // first call: scene ingest + BVH + upload + render
        const auto t0 = std::chrono::steady_clock::now();
        renderer->render(scene, UniformSampler::SamplerType::Uniform, image); // the scene is cached: this is the render alone
        const double wall = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        const xrtg_stats& st = renderer->lastStats();
        std::printf("%ux%u, %u spp on %d GPU(s): %.2f ms wall (render + reduce + image D2H), %.1f Msamples/s, %.1f Mrays/s, reduce %.3f ms, %llu dropped samples\n",
                    width, height, n_samples, st.n_devices, wall, st.samples / wall / 1e3, (st.closest_rays + st.shadow_rays) / wall / 1e3, st.reduce_ms,
                    (unsigned long long)st.dropped_samples);
    }
    catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 2;
    }
    image.gammaCorrection(1.2f);
    image.writePPM(out);
    return 0;
}
