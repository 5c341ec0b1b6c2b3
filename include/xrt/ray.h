// xrt/ray.h — Ray / SurfaceInfo / IntersectInfo / AABB PODs of the drop-in API (reference ray.h:5-44).
#pragma once
#include "geometry.h"

class Ray {
public:
    Vec3f origin;
    Vec3f direction;
    Vec3f throughput;
    int depth = 0;
    Ray() {}
    Ray(const Vec3f& origin, const Vec3f& direction) : origin(origin), direction(direction) {}
    Vec3f operator()(float t) const { return origin + t * direction; }
};

struct SurfaceInfo {
    Vec3f position;
    Vec3f ng; // geometric normal (winding-dependent, never face-forwarded: primitive.cpp:105)
    Vec3f ns; // shading normal (interpolated, NOT re-normalised: primitive.cpp:106)
    Vec3f dpdu, dpdv;
    Vec2f texcoords;
    Vec2f barycentric;
};

class Object;
struct IntersectInfo {
    float t1 = kInfinity; // exit distance of a medium box
    float t = kInfinity;
    SurfaceInfo surfaceInfo;
    const Object* hitObject = nullptr;
};

struct AABB {
    Vec3f pMin;
    Vec3f pMax;
};
