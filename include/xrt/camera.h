// xrt/camera.h — Camera / PinholeCamera of the drop-in API (reference camera.h:7-60). The ray formula
// itself (camera.h:52-55) runs on the GPU; the host evaluates scale = tan(FOV/2) exactly as camera.h:44
// does and hands it over in xrtg_camera.
#pragma once
#include "ray.h"
#include "sampler.h"
#include <xrtgpu.h>
#if defined(__has_include)
#if __has_include(<spdlog/spdlog.h>)
#include <spdlog/spdlog.h> // the reference's camera.h:5 exposes spdlog transitively; examples/vpt.cpp:25 relies on it
#endif
#endif

class Camera {
protected:
    float aspect_ratio;
    Matrix44f camera2world;

public:
    Camera(float aspect_ratio_, const Matrix44f& c2w) : aspect_ratio(aspect_ratio_), camera2world(c2w) {}
    virtual ~Camera() = default;
    void setTransform(const Matrix44f& c2w) { camera2world = c2w; }
    // additive accessors (the reference keeps these protected, camera.h:10-11)
    float aspect() const { return aspect_ratio; }
    const Matrix44f& c2w() const { return camera2world; }
    // fills the C-ABI camera; false if this camera model has no GPU implementation
    virtual bool describe(xrtg_camera& out) const = 0;
};

class PinholeCamera : public Camera {
    float FOV;
    float scale;

public:
    PinholeCamera(float aspect_ratio_, const Matrix44f& c2w, float FOV = 90.0f) : Camera(aspect_ratio_, c2w), FOV(FOV)
    {
        scale = std::tan(0.5f * deg2rad(FOV));
    }
    float fov() const { return FOV; }
    float getScale() const { return scale; }
    bool describe(xrtg_camera& out) const override
    {
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) out.c2w[4 * r + c] = camera2world[r][c];
        out.scale = scale;
        out.aspect = aspect_ratio;
        return true;
    }
};
