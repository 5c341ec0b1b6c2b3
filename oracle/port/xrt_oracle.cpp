// xrt_oracle.cpp — CPU restatement ("port") of the reference's path-tracing hot path, operating on the
// same flattened xrtg_scene_desc the GPU consumes.
//
// TEST INFRASTRUCTURE ONLY. Nothing here is linked, imported or executed by the product path
// (libxrtgpu.so / libxrthost.so). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load oracle/libxrtoracle.so, and only as the checker / timed CPU baseline.
//
// Parity status: PINNED. The reference has no tests or golden vectors of its own, so this
// restatement is pinned against outputs of the reference itself compiled here (oracle/_ref, see
// oracle/Makefile and tests/test_oracle_vs_reference.py): images of every integrator, primary hits and
// the per-function known-answer hooks agree BIT FOR BIT on the same seeds. Two third-party pieces are
// "parity unpinned" because the reference does not vendor them: OBJ polygon triangulation order
// (tinyobjloader) and the density-grid lookup (OpenVDB); both are restated (include/xrt/obj_reader.h,
// DenseGrid below / oracle/ref_harness.cpp) and shared by every implementation compared.
//
// Every function cites the reference file:line (relative to /root/reference/Src) it follows. Arithmetic
// order, float/double promotion and RNG draw order are kept exactly; x86-64 baseline code generation
// (no -march=native, no -ffast-math) means no FMA contraction on either side.
#include <algorithm>
#include <chrono>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include <omp.h>
#include "xrtgpu.h"

namespace xo {

constexpr float kPI = 3.14159265359; // geometry.h:10
constexpr float kRayEps = 1e-3f;     // geometry.h:23

// std::min / std::max exactly (NaN behaviour included): min(a,b) = (b<a)?b:a ; max(a,b) = (a<b)?b:a
inline float smin(float a, float b) { return (b < a) ? b : a; }
inline float smax(float a, float b) { return (a < b) ? b : a; }

struct V3 {
    float x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(float s) : x(s), y(s), z(s) {}
    V3(int s) : x(float(s)), y(float(s)), z(float(s)) {}
    V3(float x, float y, float z) : x(x), y(y), z(z) {}
    explicit V3(const float* p) : x(p[0]), y(p[1]), z(p[2]) {}
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    float& at(int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
// geometry.h:177-238
inline V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator*(V3 a, V3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline V3 operator/(V3 a, V3 b) { return V3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline V3 operator+(V3 a, float k) { return V3(a.x + k, a.y + k, a.z + k); }
inline V3 operator*(V3 a, float k) { return V3(a.x * k, a.y * k, a.z * k); }
inline V3 operator*(float k, V3 a) { return a * k; }
inline V3 operator/(V3 a, float k) { return V3(a.x / k, a.y / k, a.z / k); }
inline V3 operator/(float k, V3 a) { return V3(k / a.x, k / a.y, k / a.z); }
inline V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }                                  // geometry.h:251-255
inline V3 cross(V3 a, V3 b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); } // geometry.h:257-262
inline float length(V3 v) { return std::sqrt(dot(v, v)); }                                                    // geometry.cpp:3-6
inline V3 normalize(V3 v) { return v / length(v); }                                                           // geometry.cpp:13-16
inline V3 vexp(V3 v) { return V3(std::exp(v.x), std::exp(v.y), std::exp(v.z)); }                              // geometry.cpp:18-21

// geometry.cpp:44-48 (the active branch, Duff et al.)
inline void orthonormalBasis(V3 n, V3& t, V3& b)
{
    const float sign = std::copysign(1.0f, n.z);
    const float a = -1.0f / (sign + n.z);
    const float c = n.x * n.y * a;
    t = V3(1.0f + sign * n.x * n.x * a, sign * c, -sign * n.x);
    b = V3(c, sign + n.y * n.y * a, -n.y);
}
// geometry.h:693-701
inline V3 localToWorld(V3 v, V3 lx, V3 ly, V3 lz)
{
    return V3(v.x * lx.x + v.y * ly.x + v.z * lz.x, v.x * lx.y + v.y * ly.y + v.z * lz.y, v.x * lx.z + v.y * ly.z + v.z * lz.z);
}

// ---- sampler: std::mt19937 restated + libstdc++'s uniform_real_distribution<float> mapping ---------------
// sampler.h:8-50. mt19937: 32-bit Mersenne twister (n=624, m=397, r=31, a=0x9908b0df, u=11, s=7,
// b=0x9d2c5680, t=15, c=0xefc60000, l=18, f=1812433253), generated incrementally in place (identical to
// the batch twist because word i only depends on words i, i+1 and i+397 of the evolving state).
// generate_canonical<float,24>: one 32-bit draw, float(raw) / 2^32 with round-to-nearest, and values
// that round up to 1.0 are replaced by nextafter(1,0).
struct Sampler {
    uint32_t mt[624];
    int idx = 624;
    uint64_t draws = 0;
    void setSeed(uint32_t seed)
    {
        mt[0] = seed;
        for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + uint32_t(i);
        idx = 0;
    }
    uint32_t nextU32()
    {
        const int i = idx, i1 = (i + 1 == 624) ? 0 : i + 1, im = (i + 397 >= 624) ? i + 397 - 624 : i + 397;
        const uint32_t y = (mt[i] & 0x80000000u) | (mt[i1] & 0x7fffffffu);
        uint32_t v = mt[im] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        mt[i] = v;
        idx = i1;
        v ^= v >> 11;
        v ^= (v << 7) & 0x9d2c5680u;
        v ^= (v << 15) & 0xefc60000u;
        v ^= v >> 18;
        return v;
    }
    float next()
    {
        ++draws;
        const float r = float(nextU32()) / 4294967296.0f;
        return (r >= 1.0f) ? 0x1.fffffep-1f : r;
    }
};

// ---- scene ----------------------------------------------------------------------------------------------
struct Tri { V3 v0, v1, v2, n0, n1, n2; };
struct Obj {
    int kind, first, count, material, areaLight, medium;
    int firstPrim;
};
struct AreaLight {
    int kind;
    V3 v0, v1, v2, e1, e2, Ng, Le;
    float radius;
};
struct DeltaLight { int kind; V3 p, L; };
struct Grid {
    int nx, ny, nz;
    std::vector<float> data;
    V3 origin;
    float voxel, background, maxDensity;
    int lo[3], hi[3];
    // DenseGrid restatement of OpenVDBGrid::getDensity (grid.h:71-77): fp32, z then y then x lerps
    float voxelAt(int x, int y, int z) const
    {
        if (x < 0 || y < 0 || z < 0 || x >= nx || y >= ny || z >= nz) return background;
        return data[(size_t(z) * ny + y) * nx + x];
    }
    float density(V3 p) const
    {
        const float fx = (p.x - origin.x) / voxel, fy = (p.y - origin.y) / voxel, fz = (p.z - origin.z) / voxel;
        const float bx = std::floor(fx), by = std::floor(fy), bz = std::floor(fz);
        const float wx = fx - bx, wy = fy - by, wz = fz - bz;
        const int x = int(bx), y = int(by), z = int(bz);
        auto lerp = [](float a, float b, float w) { return a + (b - a) * w; };
        const float c00 = lerp(voxelAt(x, y, z), voxelAt(x, y, z + 1), wz);
        const float c01 = lerp(voxelAt(x, y + 1, z), voxelAt(x, y + 1, z + 1), wz);
        const float c10 = lerp(voxelAt(x + 1, y, z), voxelAt(x + 1, y, z + 1), wz);
        const float c11 = lerp(voxelAt(x + 1, y + 1, z), voxelAt(x + 1, y + 1, z + 1), wz);
        return lerp(lerp(c00, c01, wy), lerp(c10, c11, wy), wx);
    }
};
struct Medium {
    int kind;
    float g;
    V3 sigma_a, sigma_s, sigma_t; // homogeneous (medium.h:124-145)
    V3 absColor, scatColor;       // heterogeneous (medium.h:283-289)
    float densityMul, majorant, invMajorant;
    int grid;
};

struct Scene {
    std::vector<Obj> objs;
    std::vector<Tri> tris;
    std::vector<xrtg_sphere> spheres;
    std::vector<xrtg_box> boxes;
    std::vector<V3> albedo;
    std::vector<AreaLight> lights;
    std::vector<DeltaLight> dlights;
    std::vector<Medium> media;
    std::vector<Grid> grids;
};

struct Ray {
    V3 o, d;
    V3 at(float t) const { return o + t * d; } // ray.h:19
};

// IntersectInfo + SurfaceInfo (ray.h:23-39); fields keep their previous value unless an object overwrites
// them, exactly as in the reference (e.g. Sphere::intersect leaves dpdu/dpdv/barycentric untouched).
struct Info {
    float t1 = FLT_MAX, t = FLT_MAX;
    V3 position, ng, ns, dpdu, dpdv;
    float bu = 0, bv = 0;
    int obj = -1;
    int prim = -1; // global primitive id (not in the reference; for the parity hooks)
};

struct Counters { uint64_t closest = 0, shadow = 0, steps = 0; };

// primitive.cpp:140-168 (CULLING undefined)
inline bool rayTriangle(V3 orig, V3 dir, V3 v0, V3 v1, V3 v2, float& t, float& u, float& v)
{
    const V3 v0v1 = v1 - v0, v0v2 = v2 - v0;
    const V3 pvec = cross(dir, v0v2);
    const float det = dot(v0v1, pvec);
    if (std::fabs(det) < FLT_EPSILON) return false;
    const float invDet = 1 / det;
    const V3 tvec = orig - v0;
    u = dot(tvec, pvec) * invDet;
    if (u < 0 || u > 1) return false;
    const V3 qvec = cross(tvec, v0v1);
    v = dot(dir, qvec) * invDet;
    if (v < 0 || u + v > 1) return false;
    t = dot(v0v2, qvec) * invDet;
    return t > FLT_EPSILON;
}

// primitive.h:133-177. The literals -0.5 and the unqualified sqrt() are double in the reference.
inline bool sphereT(const xrtg_sphere& s, V3 orig, V3 dir, float& tNear)
{
    const V3 L = orig - V3(s.center);
    const float a = dot(dir, dir);
    const float b = 2 * dot(dir, L);
    const float r2 = s.radius * s.radius;
    const float c = dot(L, L) - r2;
    float t0, t1;
    const float discr = b * b - 4 * a * c;
    if (discr < 0) return false;
    else if (discr == 0) { t0 = t1 = float(-0.5 * double(b) / double(a)); }
    else {
        const float q = (b > 0) ? float(-0.5 * (double(b) + std::sqrt(double(discr)))) : float(-0.5 * (double(b) - std::sqrt(double(discr))));
        t0 = q / a;
        t1 = c / q;
    }
    if (t0 > t1) std::swap(t0, t1);
    if (t0 < 0) {
        t0 = t1;
        if (t0 < 0) return false;
    }
    tNear = t0;
    return true;
}

// primitive.h:243-264
inline bool boxSlabs(const xrtg_box& bx, const Ray& ray, float& t0, float& t1)
{
    const V3 inv = 1.0f / ray.d;
    const V3 top = inv * (V3(bx.pmax) - ray.o);
    const V3 bot = inv * (V3(bx.pmin) - ray.o);
    const V3 tmn(smin(top.x, bot.x), smin(top.y, bot.y), smin(top.z, bot.z));
    const V3 tmx(smax(top.x, bot.x), smax(top.y, bot.y), smax(top.z, bot.z));
    t0 = smax(smax(tmn.x, tmn.y), tmn.z);
    t1 = smin(smin(tmx.x, tmx.y), tmx.z);
    if (t0 > t1 || t1 <= 0.0f) return false;
    t0 = smax(t0, 0.0f);
    return true;
}

// Scene::intersect (scene.cpp:190-200) over Mesh::intersect (primitive.cpp:83-116), Sphere::intersect
// (primitive.h:106-124) and BoxMesh::intersect (primitive.h:243-264).
inline bool intersect(const Scene& sc, const Ray& ray, Info& info)
{
    bool any = false;
    for (size_t oi = 0; oi < sc.objs.size(); ++oi) {
        const Obj& o = sc.objs[oi];
        if (o.kind == XRTG_OBJ_MESH) {
            for (int k = 0; k < o.count; ++k) {
                const Tri& T = sc.tris[o.first + k];
                float t = 0, u = 0, v = 0;
                if (rayTriangle(ray.o, ray.d, T.v0, T.v1, T.v2, t, u, v)) {
                    any = true;
                    if (t < info.t) {
                        info.t = t;
                        info.position = ray.at(t);
                        info.bu = u; info.bv = v;
                        info.ng = normalize(cross(T.v1 - T.v0, T.v2 - T.v0));
                        info.ns = T.n0 * (1.0f - u - v) + T.n1 * u + T.n2 * v;
                        orthonormalBasis(info.ns, info.dpdu, info.dpdv);
                        info.obj = int(oi);
                        info.prim = o.firstPrim + k;
                    }
                }
            }
        }
        else if (o.kind == XRTG_OBJ_SPHERE) {
            float t = 0;
            if (sphereT(sc.spheres[o.first], ray.o, ray.d, t)) {
                any = true;
                if (t < info.t) {
                    info.t = t;
                    info.position = ray.at(t);
                    info.ng = normalize(ray.at(t) - V3(sc.spheres[o.first].center));
                    info.ns = info.ng;
                    info.obj = int(oi);
                    info.prim = o.firstPrim;
                }
            }
        }
        else {
            float t0, t1;
            if (boxSlabs(sc.boxes[o.first], ray, t0, t1)) {
                any = true;
                info.obj = int(oi); // unconditional overwrite (primitive.h:259-261)
                info.t = t0;
                info.t1 = t1;
                info.prim = o.firstPrim;
            }
        }
    }
    return any;
}

// Scene::occluded (scene.cpp:202-211): emitter proxies skipped; BoxMesh::occluded is always true.
inline bool occluded(const Scene& sc, const Ray& ray, float tmax)
{
    for (const Obj& o : sc.objs) {
        if (o.areaLight >= 0) continue;
        if (o.kind == XRTG_OBJ_MESH) {
            for (int k = 0; k < o.count; ++k) {
                const Tri& T = sc.tris[o.first + k];
                float t = 0, u = 0, v = 0;
                if (rayTriangle(ray.o, ray.d, T.v0, T.v1, T.v2, t, u, v) && t < tmax) return true;
            }
        }
        else if (o.kind == XRTG_OBJ_SPHERE) {
            float t = 0;
            if (sphereT(sc.spheres[o.first], ray.o, ray.d, t) && t < tmax) return true;
        }
        else return true;
    }
    return false;
}

// AreaLight::Le (light.h:62-69) through Object::Le (primitive.cpp:56-62): one-sided, tested with ns.
inline V3 emitted(const Scene& sc, const Info& info, V3 rayDir)
{
    const int li = sc.objs[info.obj].areaLight;
    if (li < 0) return V3(0.0f);
    return (dot(rayDir, info.ns) < 0) ? sc.lights[li].Le : V3(0.0f);
}

// AreaLight::sample. Quad light.cpp:59-68, triangle light.cpp:21-30 + :43-47, sphere light.h:158-197.
// Draw order: g++ evaluates the two getNext1D() operands of light.cpp:61 and the two arguments of
// light.cpp:23 right-to-left, i.e. the SECOND-written call draws first (pinned by the KAT against
// oracle/_ref). The sphere light draws cos_theta first, then phi (separate statements).
inline V3 sampleLight(const AreaLight& L, V3 position, V3& wi, float& pdf, float& tmax, Sampler& s)
{
    if (L.kind == XRTG_LIGHT_QUAD) {
        const float rb = s.next(); // multiplies e2
        const float ra = s.next(); // multiplies e1
        const V3 d = (L.v0 + L.e1 * ra + L.e2 * rb) - position;
        tmax = length(d);
        const float dn = dot(d, L.Ng);
        if (dn >= 0) return V3(0.0f);
        wi = d / tmax;
        pdf = (tmax * tmax * tmax) / std::abs(dn);
        return L.Le;
    }
    if (L.kind == XRTG_LIGHT_TRIANGLE) {
        const float v = s.next();
        const float u = s.next();
        const float su = std::sqrt(u);
        const V3 A = L.v0, B = L.v1, C = L.v2;
        const V3 p = C + (1.f - su) * (A - C) + (v * su) * (B - C);
        const V3 d = p - position;
        tmax = length(d);
        const float dn = dot(d, L.Ng);
        if (dn >= 0) return V3(0.0f);
        wi = d / tmax;
        pdf = (2.f * tmax * tmax * tmax) / std::abs(dn);
        return L.Le;
    }
    // sphere, cone sampling
    V3 dz = L.v0 - position;
    const float dz_len_2 = dot(dz, dz);
    const float dz_len = std::sqrt(dz_len_2);
    dz = dz / V3(-dz_len); // `dz /= -dz_len` goes through Vec3::operator/=(const Vec3&)
    V3 dx, dy;
    orthonormalBasis(dz, dx, dy);
    const float sin_theta_max_2 = L.radius * L.radius / dz_len_2;
    const float sin_theta_max = std::sqrt(sin_theta_max_2);
    const float cos_theta_max = std::sqrt(smax(0.f, 1.f - sin_theta_max_2));
    const float cos_theta = 1 + (cos_theta_max - 1) * s.next();
    const float sin_theta_2 = 1.f - cos_theta * cos_theta;
    const float cos_alpha = sin_theta_2 / sin_theta_max + cos_theta * std::sqrt(smax(0.0f, 1 - sin_theta_2 / sin_theta_max_2));
    const float sin_alpha = std::sqrt(smax(0.0f, 1 - cos_alpha * cos_alpha));
    const float phi = 2 * kPI * s.next();
    const V3 n = std::cos(phi) * sin_alpha * dx + std::sin(phi) * sin_alpha * dy + cos_alpha * dz;
    const V3 p = L.v0 + n * L.radius;
    const V3 d = p - position;
    tmax = length(d);
    const float d_dot_n = dot(d, n);
    if (d_dot_n >= 0) return V3(0.0f);
    pdf = 1.f / (2.f * kPI * (1.f - cos_theta_max));
    wi = d / tmax;
    return L.Le;
}

// Lambert::sampleDir (material.h:55-73): uniform hemisphere about ng with ns's tangent frame; pdf 1/2PI.
inline V3 lambertSampleDir(const Info& info, Sampler& s, float& pdf)
{
    const float r1 = s.next();
    const float r2 = s.next();
    pdf = 1 / (2 * kPI);
    const float sinTheta = sqrtf(1 - r1 * r1);
    const float phi = 2 * kPI * r2;
    const float x = sinTheta * cosf(phi);
    const float z = sinTheta * sinf(phi);
    return localToWorld(V3(x, r1, z), info.dpdu, info.ng, info.dpdv);
}

// Object::sampleBxDF / evaluateBxDF (primitive.cpp:32-46, material.h:39-53): zero without a material.
inline V3 evalBxDF(const Scene& sc, const Info& info) { const int m = sc.objs[info.obj].material; return m < 0 ? V3(0.0f) : sc.albedo[m] / kPI; }
inline V3 sampleBxDF(const Scene& sc, const Info& info, Sampler& s, V3& wi, float& pdf)
{
    if (sc.objs[info.obj].material < 0) return V3(0.0f);
    wi = lambertSampleDir(info, s, pdf);
    return evalBxDF(sc, info);
}

// HenyeyGreenstein (medium.h:21-68). getNext2D() is Vec2f(dis(gen), dis(gen)) (sampler.h:49): g++ evaluates
// the constructor arguments right-to-left, so u[1] is drawn first (pinned by the KAT).
inline float hgEval(float g, V3 wo, V3 wi)
{
    const float cosTheta = dot(wo, wi);
    const float denom = 1 + g * g - 2 * g * cosTheta;
    const float pi4inv = 1.0f / (4.0f * kPI);
    return pi4inv * (1 - g * g) / (denom * std::sqrt(denom));
}
inline float hgSample(float g, V3 wo, Sampler& s, V3& wi)
{
    const float u1 = s.next();
    const float u0 = s.next();
    float cosTheta;
    if (std::abs(g) < 1e-3) cosTheta = 2 * u0 - 1.0f;
    else {
        const float sqrTerm = (1 - g * g) / (1 - g + 2 * g * u0);
        cosTheta = (1 + g * g - sqrTerm * sqrTerm) / (2 * g);
    }
    const float sinTheta = std::sqrt(smax(1.0f - cosTheta * cosTheta, 0.0f));
    const float phi = 2 * kPI * u1;
    const V3 local(std::cos(phi) * sinTheta, cosTheta, std::sin(phi) * sinTheta);
    V3 t, b;
    orthonormalBasis(wo, t, b);
    wi = localToWorld(local, t, wo, b);
    return hgEval(g, wo, wi);
}

// Medium::sampleWavelength (medium.h:102-115) over DiscreteEmpiricalDistribution1D (sampler.h:53-97).
// std::lower_bound over the 4-entry cdf is unrolled in its actual probe order. The reference reads cdf[4]
// (out of bounds) when u exceeds cdf[3]; the restatement clamps the channel to 2 there.
inline uint32_t sampleWavelength(V3 throughput, V3 albedo, Sampler& s, V3& pmf)
{
    const V3 ta = throughput * albedo;
    float sum = 0;
    sum += ta.x; sum += ta.y; sum += ta.z;
    float cdf[4];
    cdf[0] = 0;
    cdf[1] = cdf[0] + ta.x / sum;
    cdf[2] = cdf[1] + ta.y / sum;
    cdf[3] = cdf[2] + ta.z / sum;
    pmf = V3(cdf[1] - cdf[0], cdf[2] - cdf[1], cdf[3] - cdf[2]);
    const float u = s.next();
    int x;
    if (cdf[2] < u) x = (cdf[3] < u) ? 4 : 3;
    else if (cdf[1] < u) x = 2;
    else x = (cdf[0] < u) ? 1 : 0;
    if (x == 0) x++;
    if (x > 3) x = 3;
    return uint32_t(x - 1);
}

inline V3 analyticTr(float t, V3 sigma) { return vexp(-sigma * t); } // medium.h:95-98

// HomogeneousMedium{MIS,Achromatic,NoMIS}::sampleMedium (medium.h:154-191, 202-228, 239-276)
inline bool sampleHomogeneous(const Medium& m, const Ray& ray, V3 rayThroughput, const Info& info, Sampler& s, V3& pos, V3& dir, V3& thr)
{
    const float distToSurface = info.t1 - info.t;
    if (m.kind == XRTG_MEDIUM_HOMOGENEOUS_MIS) {
        V3 pmf(1.0f);
        const uint32_t ch = sampleWavelength(rayThroughput, m.sigma_s / m.sigma_t, s, pmf);
        const float t = -std::log(smax(1.0f - s.next(), 0.0f)) / m.sigma_t[ch];
        if (t > distToSurface - kRayEps) {
            pos = ray.at(info.t1 + kRayEps);
            dir = ray.d;
            const V3 tr = analyticTr(distToSurface, m.sigma_t);
            const V3 pdf = pmf * tr;
            thr = tr / (pdf.x + pdf.y + pdf.z);
            return false;
        }
        hgSample(m.g, ray.d, s, dir);
        pos = ray.at(info.t + t);
        const V3 tr = analyticTr(t, m.sigma_t);
        const V3 pdf = pmf * (m.sigma_t * tr);
        thr = (tr * m.sigma_s) / (pdf.x + pdf.y + pdf.z);
        return true;
    }
    if (m.kind == XRTG_MEDIUM_HOMOGENEOUS_ACHROMATIC) {
        const float t = -std::log(smax(1.0f - s.next(), 0.0f)) / m.sigma_t.x;
        if (t > distToSurface - kRayEps) {
            pos = ray.at(info.t1 + kRayEps);
            dir = ray.d;
            thr = V3(1.0f);
            return false;
        }
        hgSample(m.g, ray.d, s, dir);
        pos = ray.at(info.t + t);
        thr = m.sigma_s / m.sigma_t;
        return true;
    }
    // NoMIS
    int ch = int(3 * s.next());
    if (ch == 3) ch--;
    const float pmfw = 1.0f / 3.0f;
    const float t = -std::log(smax(1.0f - s.next(), 0.0f)) / m.sigma_t[ch];
    const float pdf_distance = m.sigma_t[ch] * std::exp(-m.sigma_t[ch] * t);
    if (t > distToSurface - kRayEps) {
        pos = ray.at(info.t1 + kRayEps);
        dir = ray.d;
        const V3 tr = analyticTr(distToSurface, m.sigma_t);
        const float p_surface = std::exp(-m.sigma_t[ch] * distToSurface);
        thr = 1.0f / 3.0f * tr / (pmfw * p_surface);
        return false;
    }
    hgSample(m.g, ray.d, s, dir);
    pos = ray.at(info.t + t);
    thr = 1.0f / 3.0f * analyticTr(t, m.sigma_t) * m.sigma_s / (pmfw * pdf_distance);
    return true;
}

inline bool anyNan(V3 v) { return std::isnan(v.x) || std::isnan(v.y) || std::isnan(v.z); }

// HeterogeneousMedium::sampleMedium — spectral delta tracking (medium.cpp:45-133)
inline bool sampleHeterogeneous(const Scene& sc, const Medium& m, const Ray& ray, V3 rayThroughput, const Info& info, Sampler& s,
                                V3& pos, V3& dir, V3& thr, Counters& cnt)
{
    const Grid& G = sc.grids[m.grid];
    V3 tt(1, 1, 1);
    float t = info.t;
    float density = m.densityMul * G.density(ray.at(t));
    V3 sigma_a = m.absColor * density;
    const V3 maj(m.majorant);
    while (true) {
        ++cnt.steps;
        V3 pmf;
        const uint32_t ch = sampleWavelength(rayThroughput * tt, (maj - sigma_a) * m.invMajorant, s, pmf);
        const float sd = -std::log(smax(1.0f - s.next(), 0.0f)) * m.invMajorant;
        t += sd;
        if (t > info.t1 - kRayEps) {
            pos = ray.at(info.t1 + kRayEps);
            dir = ray.d;
            const float rest = sd - (t - (info.t1 - kRayEps));
            const V3 tr = analyticTr(rest, maj);
            const V3 pdf = pmf * tr;
            tt = tt * (tr / (pdf.x + pdf.y + pdf.z));
            thr = anyNan(tt) ? V3(0.0f) : tt;
            return false;
        }
        density = m.densityMul * G.density(ray.at(t));
        const V3 sigma_s = m.scatColor * density;
        sigma_a = m.absColor * density;
        const V3 sigma_n = maj - sigma_a - sigma_s;
        const V3 P_s = sigma_s / (sigma_s + sigma_n);
        const V3 P_n = sigma_n / (sigma_s + sigma_n);
        if (s.next() < P_s[ch]) {
            pos = ray.at(t);
            hgSample(m.g, ray.d, s, dir);
            const V3 tr = analyticTr(sd, maj);
            const V3 pdf_distance = m.majorant * tr;
            const V3 pdf = pmf * pdf_distance * P_s;
            tt = tt * ((tr * sigma_s) / (pdf.x + pdf.y + pdf.z));
            thr = anyNan(tt) ? V3(0.0f) : tt;
            return true;
        }
        const V3 tr = analyticTr(sd, maj);
        const V3 pdf_distance = m.majorant * tr;
        const V3 pdf = pmf * pdf_distance * P_n;
        tt = tt * ((tr * sigma_n) / (pdf.x + pdf.y + pdf.z));
    }
}

inline bool sampleMedium(const Scene& sc, const Ray& ray, V3 rayThroughput, const Info& info, Sampler& s, V3& pos, V3& dir, V3& thr,
                         Counters& cnt)
{
    const Medium& m = sc.media[sc.objs[info.obj].medium];
    if (m.kind == XRTG_MEDIUM_HETEROGENEOUS) return sampleHeterogeneous(sc, m, ray, rayThroughput, info, s, pos, dir, thr, cnt);
    return sampleHomogeneous(m, ray, rayThroughput, info, s, pos, dir, thr);
}

// Medium::transmittance: analytic for homogeneous (medium.h:134-139), ratio tracking for heterogeneous
// (medium.h:360-386)
inline V3 transmittance(const Scene& sc, const Medium& m, V3 p1, V3 p2, Sampler& s, Counters& cnt)
{
    if (m.kind != XRTG_MEDIUM_HETEROGENEOUS) return analyticTr(length(p1 - p2), m.sigma_t);
    const Grid& G = sc.grids[m.grid];
    const float distToEnd = length(p1 - p2);
    float t = 0;
    Ray ray{p1, normalize(p2 - p1)};
    V3 tr(1);
    while (true) {
        const float sd = -std::log(smax(1.0f - s.next(), 0.0f)) * m.invMajorant;
        t += sd;
        if (t > distToEnd) break;
        ++cnt.steps;
        const float density = m.densityMul * G.density(ray.at(t));
        const V3 sigma_n = V3(m.majorant) - m.absColor * density - m.scatColor * density;
        tr = tr * (sigma_n * m.invMajorant);
    }
    return tr;
}

// ---- integrators -------------------------------------------------------------------------------------------

// NormalIntegrator as shipped (integrator.h:29-36)
inline V3 liNormal(const Scene& sc, const Ray& r, Sampler&, Counters& c)
{
    Info info;
    ++c.closest;
    if (intersect(sc, r, info)) return 0.5f * (info.ns + 1.0f);
    return V3(0);
}

// the furnace block (integrator.h:59-66)
inline V3 liFurnace(const Scene& sc, const Ray& r, Sampler& s, Counters& c)
{
    V3 radiance(0);
    Info info;
    ++c.closest;
    if (intersect(sc, r, info)) {
        float pdf = 1.0f;
        V3 nextDir(0.0f);
        const V3 fr = sampleBxDF(sc, info, s, nextDir, pdf);
        const float cs = smax(0.0f, dot(nextDir, info.ng));
        radiance = fr * cs * V3(1.0f) / pdf;
    }
    return radiance;
}

// shared NEE loop body of Direct/GI (integrator.h:95-108, :250-267)
inline V3 neeAllLights(const Scene& sc, const Info& info, Sampler& s, Counters& c)
{
    V3 sum(0.0f);
    for (const AreaLight& L : sc.lights) {
        V3 wi;
        float tmax, pdf = 0.0f;
        const V3 Lr = sampleLight(L, info.position, wi, pdf, tmax, s);
        if (pdf == 0) continue;
        const float bias = 0.01f;
        ++c.shadow;
        const bool vis = !occluded(sc, Ray{info.position + info.ng * bias, wi}, tmax - bias);
        const float cs = smax(0.0f, dot(info.ng, wi));
        const V3 fr = evalBxDF(sc, info);
        sum = sum + (float(vis) * fr) * Lr * cs / pdf;
    }
    return sum;
}

// DirectIntegrator (integrator.h:82-119)
inline V3 liDirect(const Scene& sc, const Ray& r, Sampler& s, Counters& c)
{
    Info info;
    ++c.closest;
    if (!intersect(sc, r, info)) return V3(float(0.18));
    if (sc.objs[info.obj].areaLight >= 0) return emitted(sc, info, r.d);
    return neeAllLights(sc, info, s, c); // radiance(0) += each light, same association
}

// IndirectIntegrator (integrator.h:129-186) and GIIntegrator (integrator.h:205-287)
inline V3 liPath(const Scene& sc, const Ray& rin, Sampler& s, Counters& c, uint32_t maxDepth, bool nee)
{
    V3 radiance(0.0f);
    Ray ray = rin;
    V3 T(1, 1, 1);
    const V3 background(0.0f);
    uint32_t depth = 0;
    while (depth < maxDepth) {
        Info info;
        ++c.closest;
        if (!intersect(sc, ray, info)) {
            radiance = radiance + T * background;
            break;
        }
        if (depth > 0) {
            const float p = smin((T.x + T.y + T.z) / 3.0f, 1.0f);
            if (s.next() >= p) break;
            T = T / V3(p);
        }
        if (sc.objs[info.obj].areaLight >= 0) {
            if (!nee || depth == 0) radiance = radiance + T * emitted(sc, info, ray.d);
            break;
        }
        if (nee) {
            // directL(0) += L_light per light (integrator.h:249-269)
            const V3 directL = neeAllLights(sc, info, s, c);
            radiance = radiance + T * directL;
        }
        float pdf = 1.0;
        V3 nextDir(0.0f);
        const V3 fr = sampleBxDF(sc, info, s, nextDir, pdf);
        const float cs = smax(.0f, dot(nextDir, info.ng));
        const float bias = 0.01f;
        T = T * (fr * cs / pdf);
        ray.o = info.position + info.ng * bias;
        ray.d = nextDir;
        depth++;
    }
    return radiance;
}

// WhittedIntegrator with only Lambert materials in existence (integrator.h:302-394): the ray queue never
// grows, so it is the root ray: Lambert hit -> delta-light diffuse term (:328-343), material-less hit
// (light proxies, media) -> nothing, miss -> sky colour (:385-389).
inline V3 liWhitted(const Scene& sc, const Ray& r, Sampler&, Counters& c)
{
    Info info;
    ++c.closest;
    if (!intersect(sc, r, info)) return V3(1, 1, 1) * V3(float(0.235294), float(0.67451), float(0.843137));
    if (sc.objs[info.obj].material < 0) return V3(0);
    V3 radiance(0);
    for (const DeltaLight& L : sc.dlights) {
        V3 wi;
        float tmax, pdf;
        if (L.kind == XRTG_DLIGHT_POINT) { // light.cpp:120-128
            const V3 ld = L.p - info.position;
            const float dist = length(ld);
            wi = ld / dist;
            pdf = dist * dist;
            tmax = dist;
        }
        else { // light.cpp:136-142
            wi = -L.p;
            pdf = 1.0f;
            tmax = FLT_MAX;
        }
        ++c.shadow;
        const bool vis = !occluded(sc, Ray{info.position + info.ng * float(0.1), wi}, tmax);
        const V3 fr = evalBxDF(sc, info);
        radiance = radiance + (float(vis) * fr) * L.L * smax(0.f, dot(info.ns, wi)) / pdf;
    }
    radiance = radiance * V3(1, 1, 1);
    return V3(0) + radiance;
}

// VolumePathTracing (integrator.h:409-473) and VolumePathTracingNEE (integrator.h:489-631).
// NOTE a hit on a plain surface advances nothing in the reference (infinite loop); the
// restatement gives up after `spinGuard` iterations without progress and returns NaN so it is visible.
inline V3 liVolume(const Scene& sc, const Ray& rin, Sampler& s, Counters& c, uint32_t maxDepth, bool nee)
{
    V3 radiance(0);
    Ray ray = rin;
    V3 T(1, 1, 1);
    const V3 background(0.0f);
    uint32_t depth = 0;
    int spinGuard = 0;
    while (depth < maxDepth) {
        Info info;
        ++c.closest;
        if (!intersect(sc, ray, info)) {
            radiance = radiance + T * background * float(depth != 0);
            break;
        }
        if (depth > 0) {
            const float p = smin((T.x + T.y + T.z) / 3.0f, 1.0f);
            if (s.next() >= p) break;
            T = T / V3(p);
        }
        const Obj& ho = sc.objs[info.obj];
        if (ho.areaLight >= 0) {
            if (!nee || depth == 0) radiance = radiance + T * emitted(sc, info, ray.d);
            break;
        }
        if (ho.medium >= 0) {
            V3 pos, dir, tm;
            const bool scattered = sampleMedium(sc, ray, T, info, s, pos, dir, tm, c);
            if (nee && scattered) {
                // sampleDirectionToLight (integrator.h:583-602) + Scene::sampleAreaLight (scene.cpp:182-188)
                unsigned int li = (unsigned int)(sc.lights.size() * s.next());
                if (li == sc.lights.size()) li--;
                const float choose = 1.0f / sc.lights.size();
                V3 dl;
                float dist, lp = 0.0f;
                const V3 Le = sampleLight(sc.lights[li], pos, dl, lp, dist, s);
                const float pdf_dir = choose * lp;
                if (pdf_dir > 0.0f) {
                    // isVisible (integrator.h:604-631): one intersect, dist_to_light ignored
                    V3 trn(1.0f);
                    bool visible = true;
                    Info sh;
                    const Ray sray{pos, dl};
                    ++c.closest;
                    if (intersect(sc, sray, sh)) {
                        const Obj& so = sc.objs[sh.obj];
                        if (so.material >= 0) visible = false;
                        else if (so.medium >= 0) trn = trn * transmittance(sc, sc.media[so.medium], sray.at(sh.t), sray.at(sh.t1), s, c);
                    }
                    if (visible) {
                        const V3 f(hgEval(sc.media[ho.medium].g, ray.d, dl));
                        const V3 Ls = trn * f * Le / pdf_dir;
                        radiance = radiance + T * tm * Ls;
                    }
                }
            }
            ray.o = pos;
            ray.d = dir;
            T = T * tm;
            if (scattered) depth++;
            spinGuard = 0;
        }
        else if (++spinGuard > 4) {
            return V3(NAN);
        }
    }
    return radiance;
}

inline V3 integrate(const Scene& sc, int kind, uint32_t maxDepth, const Ray& r, Sampler& s, Counters& c)
{
    switch (kind) {
    case XRTG_INT_NORMAL: return liNormal(sc, r, s, c);
    case XRTG_INT_FURNACE: return liFurnace(sc, r, s, c);
    case XRTG_INT_DIRECT: return liDirect(sc, r, s, c);
    case XRTG_INT_INDIRECT: return liPath(sc, r, s, c, maxDepth, false);
    case XRTG_INT_GI: return liPath(sc, r, s, c, maxDepth, true);
    case XRTG_INT_WHITTED: return liWhitted(sc, r, s, c);
    case XRTG_INT_VOLUME: return liVolume(sc, r, s, c, maxDepth, false);
    case XRTG_INT_VOLUME_NEE: return liVolume(sc, r, s, c, maxDepth, true);
    default: return V3(NAN);
    }
}

// PinholeCamera::sampleRay (camera.h:49-60) with multDirMatrix (geometry.h:653-669)
inline Ray cameraRay(const xrtg_camera& c, float u, float v)
{
    const V3 dir((2 * u - 1) * c.scale, (1 - 2 * v) * c.scale / c.aspect, -1);
    const float* m = c.c2w;
    const V3 w(dir.x * m[0] + dir.y * m[4] + dir.z * m[8], dir.x * m[1] + dir.y * m[5] + dir.z * m[9], dir.x * m[2] + dir.y * m[6] + dir.z * m[10]);
    return Ray{V3(m[12], m[13], m[14]), normalize(w)};
}

thread_local std::string g_err;

} // namespace xo

struct xrto_scene { xo::Scene sc; };

#pragma GCC visibility push(default)
extern "C" {

const char* xrto_last_error(void) { return xo::g_err.c_str(); }
int xrto_max_threads(void) { return omp_get_max_threads(); }
void xrto_scene_destroy(xrto_scene* s) { delete s; }

int xrto_scene_create(const xrtg_scene_desc* d, xrto_scene** out)
{
    using namespace xo;
    if (!d || !out || d->abi_version != XRTG_ABI_VERSION) { g_err = "bad scene desc"; return -1; }
    auto s = std::make_unique<xrto_scene>();
    Scene& sc = s->sc;
    for (int i = 0; i < d->n_triangles; ++i) {
        const xrtg_triangle& t = d->triangles[i];
        sc.tris.push_back(Tri{V3(t.v0), V3(t.v1), V3(t.v2), V3(t.n0), V3(t.n1), V3(t.n2)});
    }
    sc.spheres.assign(d->spheres, d->spheres + d->n_spheres);
    sc.boxes.assign(d->boxes, d->boxes + d->n_boxes);
    for (int i = 0; i < d->n_materials; ++i) sc.albedo.push_back(V3(d->materials[i].albedo));
    for (int i = 0; i < d->n_area_lights; ++i) {
        const xrtg_area_light& L = d->area_lights[i];
        AreaLight a;
        a.kind = L.kind; a.v0 = V3(L.v0); a.v1 = V3(L.v1); a.v2 = V3(L.v2); a.Le = V3(L.Le); a.radius = L.radius;
        a.e1 = a.v1 - a.v0; a.e2 = a.v2 - a.v0; a.Ng = cross(a.e1, a.e2); // light.cpp:6-14, 49-57
        sc.lights.push_back(a);
    }
    for (int i = 0; i < d->n_delta_lights; ++i) sc.dlights.push_back(DeltaLight{d->delta_lights[i].kind, V3(d->delta_lights[i].pos_or_dir), V3(d->delta_lights[i].radiance)});
    for (int i = 0; i < d->n_grids; ++i) {
        const xrtg_grid& g = d->grids[i];
        Grid G;
        G.nx = g.nx; G.ny = g.ny; G.nz = g.nz;
        G.data.assign(g.data, g.data + size_t(g.nx) * g.ny * g.nz);
        G.origin = V3(g.origin); G.voxel = g.voxel_size; G.background = g.background; G.maxDensity = g.max_density;
        for (int a = 0; a < 3; ++a) { G.lo[a] = g.active_min[a]; G.hi[a] = g.active_max[a]; }
        sc.grids.push_back(std::move(G));
    }
    for (int i = 0; i < d->n_media; ++i) {
        const xrtg_medium& m = d->media[i];
        Medium M{};
        M.kind = m.kind; M.g = m.g; M.grid = m.grid; M.densityMul = m.density_mul;
        M.sigma_a = V3(m.sigma_a); M.sigma_s = V3(m.sigma_s); M.sigma_t = M.sigma_a + M.sigma_s;
        M.absColor = M.sigma_a; M.scatColor = M.sigma_s;
        if (m.kind == XRTG_MEDIUM_HETEROGENEOUS) {
            // medium.cpp:5-17
            const float maxd = m.density_mul * sc.grids[m.grid].maxDensity;
            const V3 mm = M.absColor * maxd + M.scatColor * maxd;
            M.majorant = smax(mm.x, smax(mm.y, mm.z));
            M.invMajorant = 1.0f / M.majorant;
        }
        sc.media.push_back(M);
    }
    int next = 0;
    for (int i = 0; i < d->n_objects; ++i) {
        const xrtg_object& o = d->objects[i];
        sc.objs.push_back(Obj{o.kind, o.first, o.count, o.material, o.area_light, o.medium, next});
        next += (o.kind == XRTG_OBJ_MESH) ? o.count : 1;
    }
    *out = s.release();
    return 0;
}

// The reference's render loop (renderer.cpp:29-81 per pixel + :98): per-pixel mt19937 seeded j+W*i.
int xrto_render(xrto_scene* s, const xrtg_camera* cam, const xrtg_render_params* p, int nthreads, int pixel_stride, float* rgb,
                double* seconds_out, xrtg_stats* stats)
{
    using namespace xo;
    if (p->sample_offset != 0) { g_err = "the mt19937 stream has no sample offset"; return -1; }
    const int W = p->width, H = p->height;
    const int stride = pixel_stride > 1 ? pixel_stride : 1;
    const int nx = (W + stride - 1) / stride, ny = (H + stride - 1) / stride;
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    if (rgb) std::memset(rgb, 0, sizeof(float) * 3 * size_t(W) * H);
    uint64_t closest = 0, shadow = 0, steps = 0, dropped = 0;
    const auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads) reduction(+ : closest, shadow, steps, dropped)
    for (long q = 0; q < long(nx) * ny; ++q) {
        const int i = int(q / nx) * stride, j = int(q % nx) * stride;
        Sampler smp;
        smp.setSeed(uint32_t(j + W * i));
        Counters c;
        V3 acc(0);
        for (int k = 0; k < p->spp; ++k) {
            const float u = (j + smp.next()) / uint32_t(W);
            const float v = (i + smp.next()) / uint32_t(H);
            const Ray ray = cameraRay(*cam, u, v);
            const V3 radiance = integrate(s->sc, p->integrator, uint32_t(p->max_depth), ray, smp, c) / 1.0f;
            if (std::isnan(radiance.x) || std::isnan(radiance.y) || std::isnan(radiance.z)) { ++dropped; continue; }
            else if (std::isinf(radiance.x) || std::isinf(radiance.y) || std::isinf(radiance.z)) { ++dropped; continue; }
            else if (radiance.x < 0 || radiance.y < 0 || radiance.z < 0) { ++dropped; continue; }
            acc = acc + radiance;
        }
        closest += c.closest; shadow += c.shadow; steps += c.steps;
        if (rgb) {
            const V3 m = (p->flags & XRTG_FLAG_SUM_ONLY) ? acc : acc / V3(float(p->spp_total > 0 ? p->spp_total : p->spp));
            float* o = rgb + (size_t(i) * W + j) * 3;
            o[0] = m.x; o[1] = m.y; o[2] = m.z;
        }
    }
    const auto t1 = std::chrono::steady_clock::now();
    if (seconds_out) *seconds_out = std::chrono::duration<double>(t1 - t0).count();
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        stats->samples = uint64_t(nx) * ny * p->spp;
        stats->closest_rays = closest; stats->shadow_rays = shadow; stats->tracking_steps = steps; stats->dropped_samples = dropped;
    }
    return 0;
}

static void fillHit(const xo::Info& info, bool hit, const xo::Scene& sc, xrtg_hit* h)
{
    h->t = hit ? info.t : FLT_MAX; h->u = 0; h->v = 0; h->prim = hit ? info.prim : -1;
    if (!hit) return;
    const int kind = sc.objs[info.obj].kind;
    if (kind == XRTG_OBJ_MESH) { h->u = info.bu; h->v = info.bv; }
    else if (kind == XRTG_OBJ_BOX) h->u = info.t1;
}

int xrto_trace_primary(xrto_scene* s, const xrtg_camera* cam, int W, int H, int spp, const float* jitter, xrtg_hit* out)
{
    using namespace xo;
#pragma omp parallel for schedule(dynamic, 16)
    for (long q = 0; q < long(W) * H; ++q) {
        const int i = int(q / W), j = int(q % W);
        Sampler smp;
        smp.setSeed(uint32_t(j + W * i));
        for (int k = 0; k < spp; ++k) {
            float r0, r1;
            if (jitter) { r0 = jitter[(q * spp + k) * 2]; r1 = jitter[(q * spp + k) * 2 + 1]; }
            else { r0 = smp.next(); r1 = smp.next(); }
            const float u = (j + r0) / uint32_t(W);
            const float v = (i + r1) / uint32_t(H);
            const Ray ray = cameraRay(*cam, u, v);
            Info info;
            const bool hit = intersect(s->sc, ray, info);
            fillHit(info, hit, s->sc, out + q * spp + k);
        }
    }
    return 0;
}

int xrto_trace_rays(xrto_scene* s, int64_t n, const float* org, const float* dir, const float* tmax, int any_hit, xrtg_hit* out)
{
    using namespace xo;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t r = 0; r < n; ++r) {
        const Ray ray{V3(org + 3 * r), V3(dir + 3 * r)};
        if (any_hit) {
            const bool occ = occluded(s->sc, ray, tmax ? tmax[r] : FLT_MAX);
            out[r].t = 0; out[r].u = 0; out[r].v = 0; out[r].prim = occ ? 1 : 0;
        }
        else {
            Info info;
            const bool hit = intersect(s->sc, ray, info);
            fillHit(info, hit, s->sc, out + r);
        }
    }
    return 0;
}

// ---- known-answer hooks, same signatures as the xrtref_kat_* of oracle/ref_harness.cpp ------------------
void xrto_kat_sampler(uint32_t seed, int n, float* out)
{
    xo::Sampler s;
    s.setSeed(seed);
    for (int i = 0; i < n; ++i) out[i] = s.next();
}
void xrto_kat_camera(const xrtg_camera* c, float u, float v, float* out6)
{
    const xo::Ray r = xo::cameraRay(*c, u, v);
    out6[0] = r.o.x; out6[1] = r.o.y; out6[2] = r.o.z; out6[3] = r.d.x; out6[4] = r.d.y; out6[5] = r.d.z;
}
void xrto_kat_onb(const float* n, float* out6)
{
    xo::V3 t, b;
    xo::orthonormalBasis(xo::V3(n), t, b);
    out6[0] = t.x; out6[1] = t.y; out6[2] = t.z; out6[3] = b.x; out6[4] = b.y; out6[5] = b.z;
}
void xrto_kat_light_sample(xrto_scene* s, int li, const float* pos, uint32_t seed, float* out8)
{
    xo::Sampler sm;
    sm.setSeed(seed);
    xo::V3 wi;
    float pdf = 0, tmax = 0;
    const xo::V3 L = xo::sampleLight(s->sc.lights[li], xo::V3(pos), wi, pdf, tmax, sm);
    out8[0] = L.x; out8[1] = L.y; out8[2] = L.z; out8[3] = wi.x; out8[4] = wi.y; out8[5] = wi.z; out8[6] = pdf; out8[7] = tmax;
}
void xrto_kat_lambert_sample(const float* ng, const float* ns, uint32_t seed, float* out4)
{
    xo::Info info;
    info.ng = xo::V3(ng); info.ns = xo::V3(ns);
    xo::orthonormalBasis(info.ns, info.dpdu, info.dpdv);
    xo::Sampler sm;
    sm.setSeed(seed);
    float pdf;
    const xo::V3 wi = xo::lambertSampleDir(info, sm, pdf);
    out4[0] = wi.x; out4[1] = wi.y; out4[2] = wi.z; out4[3] = pdf;
}
void xrto_kat_hg_sample(float g, const float* wo, uint32_t seed, float* out4)
{
    xo::Sampler sm;
    sm.setSeed(seed);
    xo::V3 wi;
    const float f = xo::hgSample(g, xo::V3(wo), sm, wi);
    out4[0] = wi.x; out4[1] = wi.y; out4[2] = wi.z; out4[3] = f;
}

} // extern "C"
#pragma GCC visibility pop
