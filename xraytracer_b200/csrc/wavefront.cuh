// wavefront.cuh — the sm_100a wavefront path-tracing kernels of libxrtgpu.so.
//
// Included by two translation units:
//   kernels_exact.cu  XRT_EXACT=1, compiled with -fmad=false : reproduces the reference's arithmetic (no FMA
//                     contraction, IEEE div/sqrt, same operation order) and its per-pixel std::mt19937 sample
//                     stream (renderer.cpp:35-36, sampler.h:48-49). One sample per pixel per wave.
//   kernels_fast.cu   XRT_EXACT=0, -use_fast_math : counter-based Philox4x32-7 keyed (seed,pixel)/(sample,block), several
//                     samples per pixel per wave, plane-equation triangle records instead of Moeller-Trumbore.
//
// Pipeline per wave (queue sizes live in device memory, so a wave is enqueued without a host round trip):
//   shallow BVH:  primary (raygen fused with the bounce-0 closest hit, compact hit-only queue)
//   deep BVH:     raygen -> trace<closest> (resumable traversal with warp refill)
//   then per bounce: shade (Le, RR, NEE sample of every light, BSDF sample; CTA-aggregated appends into the next ray queue
//   and the shadow queue) -> connect (any hit) -> extend (closest hit of the next bounce) ... -> accumulate
//
// Reference citations (paths relative to /root/reference/Src) are on each device function.
#pragma once
#include <cfloat>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include "device_types.h"
#include <xrtgpu.h>

#ifndef XRT_EXACT
#error "define XRT_EXACT and XRT_NS before including wavefront.cuh"
#endif

namespace xrt {
namespace XRT_NS {

constexpr bool kExact = (XRT_EXACT != 0);
constexpr int kBlock = 128;          // threads per CTA of the traversal kernels (the surface shade kernel uses kShadeBlock)
constexpr int kStackSmem = 24;       // traversal stack entries kept in shared memory per thread
constexpr int kStackLocal = 40;      // overflow entries (local memory); builder depth limit is 56
constexpr float kPI = 3.14159265359; // geometry.h:10
constexpr float kRayEps = 1e-3f;     // geometry.h:23

// ---------------------------------------------------------------------------------------------------------
// small vector algebra, operation order as in geometry.h:177-262
// ---------------------------------------------------------------------------------------------------------
struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 mk(float x, float y, float z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 mk(float s) { return V3{s, s, s}; }
__device__ __forceinline__ V3 xyz(const float4& v) { return V3{v.x, v.y, v.z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(V3 a, V3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ V3 operator/(V3 a, V3 b) { return mk(a.x / b.x, a.y / b.y, a.z / b.z); }
__device__ __forceinline__ V3 operator+(V3 a, float k) { return mk(a.x + k, a.y + k, a.z + k); }
__device__ __forceinline__ V3 operator*(V3 a, float k) { return mk(a.x * k, a.y * k, a.z * k); }
__device__ __forceinline__ V3 operator*(float k, V3 a) { return a * k; }
__device__ __forceinline__ V3 operator/(V3 a, float k) { return mk(a.x / k, a.y / k, a.z / k); }
__device__ __forceinline__ V3 operator/(float k, V3 a) { return mk(k / a.x, k / a.y, k / a.z); }
__device__ __forceinline__ V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ float length(V3 v) { return sqrtf(dot(v, v)); }   // geometry.cpp:3-6
__device__ __forceinline__ V3 normalize(V3 v) { return v / length(v); }       // geometry.cpp:13-16 (divide, not rsqrt)
__device__ __forceinline__ V3 vexp(V3 v) { return mk(expf(v.x), expf(v.y), expf(v.z)); }
__device__ __forceinline__ float comp(V3 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }
// std::min / std::max with their NaN behaviour
__device__ __forceinline__ float smin(float a, float b) { return (b < a) ? b : a; }
__device__ __forceinline__ float smax(float a, float b) { return (a < b) ? b : a; }
__device__ __forceinline__ bool anyNan(V3 v) { return isnan(v.x) || isnan(v.y) || isnan(v.z); }

// geometry.cpp:44-48
__device__ __forceinline__ void orthonormalBasis(V3 n, V3& t, V3& b)
{
    const float sign = copysignf(1.0f, n.z);
    const float a = -1.0f / (sign + n.z);
    const float c = n.x * n.y * a;
    t = mk(1.0f + sign * n.x * n.x * a, sign * c, -sign * n.x);
    b = mk(c, sign + n.y * n.y * a, -n.y);
}
// geometry.h:693-701
__device__ __forceinline__ V3 localToWorld(V3 v, V3 lx, V3 ly, V3 lz)
{
    return mk(v.x * lx.x + v.y * ly.x + v.z * lz.x, v.x * lx.y + v.y * ly.y + v.z * lz.y, v.x * lx.z + v.y * ly.z + v.z * lz.z);
}

// ---------------------------------------------------------------------------------------------------------
// RNG. Exact: std::mt19937 restated (state word-major in HBM: mt[word * nPixels + pixel]) with libstdc++'s
// generate_canonical<float,24> mapping (sampler.h:37-50). Fast: Philox4x32-7, key = (seed, pixel),
// counter = (sample, block): the union of samples over GPUs equals the 1-GPU sample set.
// ---------------------------------------------------------------------------------------------------------
struct Rng {
    // exact
    uint32_t* mt;
    uint32_t* mti;
    uint32_t stride, pix, idx;
    // fast
    uint32_t k0, k1, sample, ctr, bufBlock;
    uint4 buf;

    __device__ __forceinline__ void open(const DWave& w, uint32_t pid, uint32_t counter)
    {
        pix = pid % w.nPixels;
        if constexpr (kExact) {
            mt = w.mt; mti = w.mti; stride = w.nPixels;
            idx = mti[pix];
        }
        else {
            k0 = w.seed; k1 = pix;
            sample = w.sampleBase + pid / w.nPixels;
            ctr = counter;
            bufBlock = 0xffffffffu;
        }
    }
    __device__ __forceinline__ uint32_t close()
    {
        if constexpr (kExact) { mti[pix] = idx; return 0; }
        else return ctr;
    }
    // Throughput instantiation: skip to the start of the next 4-draw block, so that lanes walking in lockstep (trackStep) all
    // compute their Philox block in the same instruction instead of one quarter of the lanes at a time. No-op for mt19937.
    __device__ __forceinline__ void alignBlock()
    {
        if constexpr (!kExact) ctr = (ctr + 3u) & ~3u;
    }
    __device__ __forceinline__ void philox(uint32_t block)
    {
        uint32_t c0 = sample, c1 = block, c2 = 0x243F6A88u, c3 = 0x85A308D3u, a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 7; ++r) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            c0 = hi1 ^ c1 ^ a; c1 = lo1; c2 = hi0 ^ c3 ^ b; c3 = lo0;
            a += 0x9E3779B9u; b += 0xBB67AE85u;
        }
        buf = make_uint4(c0, c1, c2, c3);
        bufBlock = block;
    }
    __device__ __forceinline__ float next()
    {
        if constexpr (kExact) {
            const uint32_t i = idx, i1 = (i + 1 == 624) ? 0 : i + 1, im = (i + 397 >= 624) ? i + 397 - 624 : i + 397;
            const uint32_t y = (mt[size_t(i) * stride + pix] & 0x80000000u) | (mt[size_t(i1) * stride + pix] & 0x7fffffffu);
            uint32_t v = mt[size_t(im) * stride + pix] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            mt[size_t(i) * stride + pix] = v;
            idx = i1;
            v ^= v >> 11;
            v ^= (v << 7) & 0x9d2c5680u;
            v ^= (v << 15) & 0xefc60000u;
            v ^= v >> 18;
            const float r = __uint2float_rn(v) * 2.3283064365386963e-10f; // float(raw) / 2^32
            return (r >= 1.0f) ? 0x1.fffffep-1f : r;
        }
        else {
            const uint32_t blk = ctr >> 2;
            if (blk != bufBlock) philox(blk);
            const uint32_t lane = ctr & 3u;
            ++ctr;
            const uint32_t v = lane == 0 ? buf.x : (lane == 1 ? buf.y : (lane == 2 ? buf.z : buf.w));
            return float(v >> 8) * 5.9604644775390625e-8f; // [0,1), 24 bits
        }
    }
};

// mt19937 seeding (gen.seed(j + W*i), renderer.cpp:36): one thread per pixel
__global__ void __launch_bounds__(kBlock) k_seed_mt(uint32_t* mt, uint32_t* mti, uint32_t nPixels)
{
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < nPixels; p += gridDim.x * blockDim.x) {
        uint32_t s = p; // seed = j + W*i = linear pixel index
        mt[p] = s;
        for (uint32_t i = 1; i < 624; ++i) {
            s = 1812433253u * (s ^ (s >> 30)) + i;
            mt[size_t(i) * nPixels + p] = s;
        }
        mti[p] = 0;
    }
}

// jitter of the primary samples exactly as renderer.cpp:44-47 draws them when nothing else consumes the
// stream (parity hook xrtg_trace_primary with jitter_uv == NULL): thread per pixel, samples in order.
__global__ void __launch_bounds__(kBlock) k_gen_jitter(DWave w, int spp, float* __restrict__ jitter)
{
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < w.nPixels; p += gridDim.x * blockDim.x) {
        DWave w1 = w;
        w1.samplesThisWave = 1;
        for (int s = 0; s < spp; ++s) {
            Rng rng;
            w1.sampleBase = w.sampleBase + uint32_t(s);
            rng.open(w1, p, 0);
            jitter[(size_t(p) * spp + s) * 2] = rng.next();
            jitter[(size_t(p) * spp + s) * 2 + 1] = rng.next();
            rng.close();
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// intersection primitives
// ---------------------------------------------------------------------------------------------------------

// Mesh::rayTriangleIntersect (primitive.cpp:140-168, CULLING undefined); e1 = v1-v0, e2 = v2-v0 precomputed
// on the host with the same fp32 subtraction the reference performs per ray.
__device__ __forceinline__ bool rayTriangle(V3 orig, V3 dir, V3 v0, V3 e1, V3 e2, float& t, float& u, float& v)
{
    const V3 pvec = cross(dir, e2);
    const float det = dot(e1, pvec);
    if (fabsf(det) < FLT_EPSILON) return false;
    const float invDet = 1 / det;
    const V3 tvec = orig - v0;
    u = dot(tvec, pvec) * invDet;
    if (u < 0 || u > 1) return false;
    const V3 qvec = cross(tvec, e1);
    v = dot(dir, qvec) * invDet;
    if (v < 0 || u + v > 1) return false;
    t = dot(e2, qvec) * invDet;
    return t > FLT_EPSILON;
}

struct Hit {
    float t, u, v;
    int prim;
};

// Closest-hit candidate rule that reproduces "first strictly smaller t in primitive order wins"
// (scene.cpp:193-197, primitive.cpp:100) under an arbitrary visiting order: lower t, or equal t and lower id.
__device__ __forceinline__ void consider(Hit& h, float t, float u, float v, int id)
{
    if (t < h.t || (t == h.t && id < h.prim)) { h.t = t; h.u = u; h.v = v; h.prim = id; }
}

// One triangle record against one ray. ANY=false: closest-hit candidate (ids > minId only) folded into `h`; ANY=true:
// returns true if the triangle (not an emitter proxy) occludes the ray before h.t.
//   exact instantiation : 3 float4 (v0|id, e1|flags, e2|0) and the reference's Moeller-Trumbore, bit for bit.
//   fast instantiation  : 4 float4 in plane-equation form (Havel & Herout 2010): N = e1 x e2, d = N.v0 give t = (d - N.o)/(N.dir);
//                         two affine functions of the hit point give the barycentrics, u = n1.P + d1, v = n2.P + d2 with
//                         n1 = (e2 x N)/|N|^2, n2 = (N x e1)/|N|^2. ~25 instructions instead of ~42; the same acceptance rules
//                         (|det| >= FLT_EPSILON with det = N.dir = -MT's det, u,v >= 0, u+v <= 1, t > FLT_EPSILON). It is an
//                         independent Monte-Carlo path anyway; parity lives in the exact instantiation.
constexpr int kTriF4 = kExact ? 3 : 4;
__device__ __forceinline__ const float4* triArray(const DScene& sc, bool idOrder)
{
    if constexpr (kExact) return idOrder ? sc.tris_id : sc.tris;
    else return idOrder ? sc.ftris_id : sc.ftris;
}
template <bool ANY, bool LDG>
__device__ __forceinline__ bool triangleRecord(const float4* __restrict__ rec, V3 o, V3 d, Hit& h, int minId)
{
    const float4 q0 = LDG ? __ldg(rec) : rec[0], q1 = LDG ? __ldg(rec + 1) : rec[1], q2 = LDG ? __ldg(rec + 2) : rec[2];
    if constexpr (kExact) {
        const int id = __float_as_int(q0.w);
        float t, u, v;
        if (ANY) return (__float_as_int(q1.w) & 1) == 0 && rayTriangle(o, d, xyz(q0), xyz(q1), xyz(q2), t, u, v) && t < h.t;
        if (id > minId && rayTriangle(o, d, xyz(q0), xyz(q1), xyz(q2), t, u, v)) consider(h, t, u, v, id);
        return false;
    }
    else {
        const float det = dot(xyz(q0), d);
        const float t = fmaf(-o.x, q0.x, fmaf(-o.y, q0.y, fmaf(-o.z, q0.z, q0.w))) * (1.0f / det);
        const V3 P = o + t * d;
        const float u = fmaf(P.x, q1.x, fmaf(P.y, q1.y, fmaf(P.z, q1.z, q1.w)));
        const float v = fmaf(P.x, q2.x, fmaf(P.y, q2.y, fmaf(P.z, q2.z, q2.w)));
        const bool ok = fminf(fminf(u, v), 1.f - (u + v)) >= 0.f && !(fabsf(det) < FLT_EPSILON) && t > FLT_EPSILON;
        if (!ok) return false;
        const int4 m = LDG ? __ldg(reinterpret_cast<const int4*>(rec + 3)) : *reinterpret_cast<const int4*>(rec + 3);
        if (ANY) return (m.y & 1) == 0 && t < h.t;
        if (m.x > minId) consider(h, t, u, v, m.x);
        return false;
    }
}

// Sphere::doIntersect / solveQuadratic (primitive.h:133-177); the -0.5 literals and the unqualified sqrt() make
// the root computation double precision in the reference.
__device__ __forceinline__ bool sphereT(float4 cr, V3 orig, V3 dir, float& tNear)
{
    const V3 L = orig - xyz(cr);
    const float a = dot(dir, dir);
    const float b = 2 * dot(dir, L);
    const float r2 = cr.w * cr.w;
    const float c = dot(L, L) - r2;
    float t0, t1;
    const float discr = b * b - 4 * a * c;
    if (discr < 0) return false;
    else if (discr == 0) { t0 = t1 = float(-0.5 * double(b) / double(a)); }
    else {
        const float q = (b > 0) ? float(-0.5 * (double(b) + sqrt(double(discr)))) : float(-0.5 * (double(b) - sqrt(double(discr))));
        t0 = q / a;
        t1 = c / q;
    }
    if (t0 > t1) { const float s = t0; t0 = t1; t1 = s; }
    if (t0 < 0) {
        t0 = t1;
        if (t0 < 0) return false;
    }
    tNear = t0;
    return true;
}

// BoxMesh::intersect slabs (primitive.h:243-264)
__device__ __forceinline__ bool boxSlabs(V3 pmin, V3 pmax, V3 o, V3 d, float& t0, float& t1)
{
    const V3 inv = 1.0f / d;
    const V3 top = inv * (pmax - o);
    const V3 bot = inv * (pmin - o);
    const V3 tmn = mk(smin(top.x, bot.x), smin(top.y, bot.y), smin(top.z, bot.z));
    const V3 tmx = mk(smax(top.x, bot.x), smax(top.y, bot.y), smax(top.z, bot.z));
    t0 = smax(smax(tmn.x, tmn.y), tmn.z);
    t1 = smin(smin(tmx.x, tmx.y), tmx.z);
    if (t0 > t1 || t1 <= 0.0f) return false;
    t0 = smax(t0, 0.0f);
    return true;
}

struct TraceCounters {
    uint32_t nodes = 0, tris = 0;
};

// Conservative slab test against a PADDED child box (bvh.cpp pads by 2^-15 of the scene magnitude); fminf/fmaxf
// drop NaNs (0*inf), which only ever makes the test pass. tmaxRay is inclusive: a node whose entry distance
// equals the current best is still visited (tie-break by primitive id needs it).
__device__ __forceinline__ bool slab(const float lo0, const float lo1, const float lo2, const float hi0, const float hi1,
                                     const float hi2, V3 idir, V3 ood, float tmaxRay, float& tnear)
{
    const float x0 = fmaf(lo0, idir.x, -ood.x), x1 = fmaf(hi0, idir.x, -ood.x);
    const float y0 = fmaf(lo1, idir.y, -ood.y), y1 = fmaf(hi1, idir.y, -ood.y);
    const float z0 = fmaf(lo2, idir.z, -ood.z), z1 = fmaf(hi2, idir.z, -ood.z);
    const float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));
    const float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tmaxRay));
    tnear = tn;
    return tn <= tf;
}

// BVH2 traversal, while-while, per-thread stack in shared memory ([entry][thread], conflict-free) with a
// local-memory overflow. ANY=false: closest hit into `h` (h.t / h.prim pre-set by the caller = current best;
// only primitives with id > minId are considered). ANY=true: returns true at the first triangle with
// t < h.t that is not an emitter proxy (Scene::occluded, scene.cpp:202-211).
template <bool ANY, bool COUNT>
__device__ __forceinline__ bool traverse(const DScene& sc, V3 o, V3 d, Hit& h, int minId, int* sstack, TraceCounters& tc)
{
    // Box tests only: a zero (or denormal) direction component would make lo*idir - o*idir evaluate inf - inf = NaN,
    // which fminf/fmaxf then drop on the wrong side. Clamping |d| to 1e-20 keeps every slab distance finite and
    // ordered (a ray parallel to a slab and outside it still misses, inside it still spans (-huge, +huge)); the
    // triangle test below always uses the true direction.
    const float kTiny = 1e-20f;
    const V3 ds = mk(fabsf(d.x) < kTiny ? copysignf(kTiny, d.x) : d.x, fabsf(d.y) < kTiny ? copysignf(kTiny, d.y) : d.y,
                     fabsf(d.z) < kTiny ? copysignf(kTiny, d.z) : d.z);
    const V3 idir = 1.0f / ds;
    const V3 ood = mk(o.x * idir.x, o.y * idir.y, o.z * idir.z);
    int lstack[kStackLocal];
    int sp = 0;
    int node = 0; // root
    const float4* __restrict__ nodes = sc.nodes;
    const float4* __restrict__ tris = triArray(sc, false);
    while (true) {
        // ---- inner nodes ----
        const float4 n0 = __ldg(nodes + 4 * node), n1 = __ldg(nodes + 4 * node + 1), n2 = __ldg(nodes + 4 * node + 2);
        const int4 n3 = __ldg(reinterpret_cast<const int4*>(nodes + 4 * node + 3));
        if (COUNT) tc.nodes++;
        float t0n, t1n;
        const bool h0 = (n3.z >= 0) && slab(n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, idir, ood, h.t, t0n);
        const bool h1 = (n3.w >= 0) && slab(n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, idir, ood, h.t, t1n);
        // children to process now: leaves are intersected immediately, inner children pushed / descended
        int next = -1;
        int c0 = n3.x, k0 = n3.z, c1 = n3.y, k1 = n3.w;
        bool a0 = h0, a1 = h1;
        if (a0 && a1 && t1n < t0n) { // visit the nearer child first
            const int tc_ = c0; c0 = c1; c1 = tc_;
            const int tk = k0; k0 = k1; k1 = tk;
        }
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const bool act = s == 0 ? a0 : a1;
            const int c = s == 0 ? c0 : c1, k = s == 0 ? k0 : k1;
            if (!act) continue;
            if (k > 0) {
                for (int i = 0; i < k; ++i) {
                    if (COUNT) tc.tris++;
                    if (triangleRecord<ANY, true>(tris + kTriF4 * (c + i), o, d, h, minId)) return true;
                }
            }
            else if (next < 0) next = c;
            else { // push the farther inner child
                if (sp < kStackSmem) sstack[sp * kBlock] = c;
                else lstack[sp - kStackSmem] = c;
                ++sp;
            }
        }
        if (next >= 0) { node = next; continue; }
        if (sp == 0) break;
        --sp;
        node = (sp < kStackSmem) ? sstack[sp * kBlock] : lstack[sp - kStackSmem];
    }
    return false;
}

// Brute force in primitive-id order (parity/debug path; XRTG_FLAG_BRUTE_FORCE): the reference's own loops.
template <bool ANY>
__device__ __forceinline__ bool bruteTris(const DScene& sc, V3 o, V3 d, Hit& h, int minId)
{
    const float4* __restrict__ tris = triArray(sc, true);
    for (int i = 0; i < sc.nBruteTris; ++i)
        if (triangleRecord<ANY, true>(tris + kTriF4 * i, o, d, h, minId)) return true;
    return false;
}

// Small scenes (<= kSmallSceneTris triangles, e.g. every scene the reference ships): for INCOHERENT rays a warp that walks
// a BVH diverges (8-11 of 32 lanes active per instruction, ncu), whereas testing every triangle — the reference's own loop,
// primitive.cpp:83-138 — keeps all 32 lanes converged. Triangles are staged once per CTA in shared memory (48 B each,
// broadcast reads) and the Moeller-Trumbore test is evaluated branch-free: the same operations in the same order as
// rayTriangle(), the rejections of primitive.cpp:153-167 folded into one predicate (a NaN anywhere ends in `t > eps`
// being false, exactly like the reference's early returns).
constexpr int kSmallSceneTris = 64;
constexpr int kSmallBlockF4 = 704; // = kSmallBlockMaxF4 (small_scene.h)
template <bool ANY, bool OCCLUDERS_ONLY = false>
__device__ __forceinline__ bool smallSceneTris(const float4* __restrict__ st, int n, V3 o, V3 d, Hit& h, int minId)
{
    bool occluded = false;
    if constexpr (kExact) {
#pragma unroll 2
        for (int i = 0; i < n; ++i) {
            const float4 q0 = st[3 * i], q1 = st[3 * i + 1], q2 = st[3 * i + 2];
            const V3 v0 = xyz(q0), e1 = xyz(q1), e2 = xyz(q2);
            const V3 pvec = cross(d, e2);
            const float det = dot(e1, pvec);
            const float invDet = 1 / det;
            const V3 tvec = o - v0;
            const float u = dot(tvec, pvec) * invDet;
            const V3 qvec = cross(tvec, e1);
            const float v = dot(d, qvec) * invDet;
            const float t = dot(e2, qvec) * invDet;
            const bool ok = !(fabsf(det) < FLT_EPSILON) && !(u < 0 || u > 1) && !(v < 0 || u + v > 1) && (t > FLT_EPSILON);
            if (ANY) occluded = occluded || (ok && (__float_as_int(q1.w) & 1) == 0 && t < h.t);
            else {
                const int id = __float_as_int(q0.w);
                if (ok && id > minId) consider(h, t, u, v, id);
            }
        }
    }
    else {
        // Plane-equation records (see triangleRecord). fma chains seeded with the plane offsets: 18 FP ops per triangle; the three
        // barycentric conditions collapse into one FMNMX3. The list is in primitive-id order, so for the closest hit "strictly
        // smaller t wins" IS the reference's first-wins rule (scene.cpp:193-197) — only (t, index) are carried through the loop and
        // u, v, id are re-derived for the winner. Any-hit keeps the minimum valid t (branch-free) and compares once at the end.
        if (ANY) {
            float tmin = FLT_MAX;
#pragma unroll 2
            for (int i = 0; i < n; ++i) {
                const float4 q0 = st[4 * i], q1 = st[4 * i + 1], q2 = st[4 * i + 2];
                const float det = dot(xyz(q0), d);
                const float t = fmaf(-o.x, q0.x, fmaf(-o.y, q0.y, fmaf(-o.z, q0.z, q0.w))) * (1.0f / det);
                const V3 P = o + t * d;
                const float u = fmaf(P.x, q1.x, fmaf(P.y, q1.y, fmaf(P.z, q1.z, q1.w)));
                const float v = fmaf(P.x, q2.x, fmaf(P.y, q2.y, fmaf(P.z, q2.z, q2.w)));
                bool ok = fminf(fminf(u, v), 1.f - (u + v)) >= 0.f && !(fabsf(det) < FLT_EPSILON) && t > FLT_EPSILON;
                if (!OCCLUDERS_ONLY) ok = ok && (__float_as_int(st[4 * i + 3].y) & 1) == 0;
                tmin = fminf(tmin, ok ? t : FLT_MAX);
            }
            occluded = tmin < h.t;
        }
        else {
            float best = h.t;
            int bi = -1;
            if (minId < 0) {
#pragma unroll 2
                for (int i = 0; i < n; ++i) {
                    const float4 q0 = st[4 * i], q1 = st[4 * i + 1], q2 = st[4 * i + 2];
                    const float det = dot(xyz(q0), d);
                    const float t = fmaf(-o.x, q0.x, fmaf(-o.y, q0.y, fmaf(-o.z, q0.z, q0.w))) * (1.0f / det);
                    const V3 P = o + t * d;
                    const float u = fmaf(P.x, q1.x, fmaf(P.y, q1.y, fmaf(P.z, q1.z, q1.w)));
                    const float v = fmaf(P.x, q2.x, fmaf(P.y, q2.y, fmaf(P.z, q2.z, q2.w)));
                    const bool ok = fminf(fminf(u, v), 1.f - (u + v)) >= 0.f && !(fabsf(det) < FLT_EPSILON) && t > FLT_EPSILON && t < best;
                    if (ok) { best = t; bi = i; }
                }
            }
            else { // a BoxMesh hit came first: only primitives after it in object order may replace it (primitive.h:259-261)
                for (int i = 0; i < n; ++i) {
                    const float4 q0 = st[4 * i], q1 = st[4 * i + 1], q2 = st[4 * i + 2];
                    const int id = __float_as_int(st[4 * i + 3].x);
                    const float det = dot(xyz(q0), d);
                    const float t = fmaf(-o.x, q0.x, fmaf(-o.y, q0.y, fmaf(-o.z, q0.z, q0.w))) * (1.0f / det);
                    const V3 P = o + t * d;
                    const float u = fmaf(P.x, q1.x, fmaf(P.y, q1.y, fmaf(P.z, q1.z, q1.w)));
                    const float v = fmaf(P.x, q2.x, fmaf(P.y, q2.y, fmaf(P.z, q2.z, q2.w)));
                    const bool ok = fminf(fminf(u, v), 1.f - (u + v)) >= 0.f && !(fabsf(det) < FLT_EPSILON) && t > FLT_EPSILON && t < best;
                    if (ok && id > minId) { best = t; bi = i; }
                }
            }
            if (bi >= 0) {
                const float4 q1 = st[4 * bi + 1], q2 = st[4 * bi + 2];
                const V3 P = o + best * d;
                h.t = best;
                h.u = fmaf(P.x, q1.x, fmaf(P.y, q1.y, fmaf(P.z, q1.z, q1.w)));
                h.v = fmaf(P.x, q2.x, fmaf(P.y, q2.y, fmaf(P.z, q2.z, q2.w)));
                h.prim = __float_as_int(st[4 * bi + 3].x);
            }
        }
    }
    return occluded;
}

// Scene::intersect (scene.cpp:190-200) for one ray: boxes first (the LAST box hit in object order overwrites
// whatever came before it, primitive.h:259-261; objects after it win only with a strictly smaller t), then
// triangles through the BVH, then analytic spheres. Box hits return t1 in h.u.
template <bool COUNT, bool SMALL = false>
__device__ __forceinline__ void closestHit(const DScene& sc, V3 o, V3 d, bool brute, Hit& h, int* sstack, TraceCounters& tc,
                                           const float4* smallTris = nullptr)
{
    h.t = FLT_MAX; h.u = 0.f; h.v = 0.f; h.prim = 0x7fffffff;
    int minId = -1;
    for (int b = 0; b < sc.nBoxes; ++b) { // boxes[] are in object order
        const float4 bl = __ldg(sc.boxes + 2 * b), bh = __ldg(sc.boxes + 2 * b + 1);
        float t0, t1;
        if (boxSlabs(xyz(bl), xyz(bh), o, d, t0, t1)) { h.t = t0; h.u = t1; h.v = 0.f; h.prim = __float_as_int(bl.w); minId = h.prim; }
    }
    if (sc.nTris > 0) {
        if (SMALL || smallTris) smallSceneTris<false>(smallTris, sc.nBruteTris, o, d, h, minId);
        else if (brute) bruteTris<false>(sc, o, d, h, minId);
        else traverse<false, COUNT>(sc, o, d, h, minId, sstack, tc);
    }
    for (int s = 0; s < sc.nSpheres; ++s) {
        const float4 cr = __ldg(sc.spheres + 2 * s);
        const int4 meta = __ldg(reinterpret_cast<const int4*>(sc.spheres + 2 * s + 1));
        float t;
        if (meta.x > minId && sphereT(cr, o, d, t)) consider(h, t, 0.f, 0.f, meta.x);
    }
    if (h.prim == 0x7fffffff) h.prim = -1;
}

// Scene::occluded (scene.cpp:202-211): BoxMesh::occluded is always true (primitive.h:266-268)
template <bool COUNT, bool SMALL = false>
__device__ __forceinline__ bool anyHit(const DScene& sc, V3 o, V3 d, float tmax, bool brute, int* sstack, TraceCounters& tc,
                                       const float4* smallTris = nullptr)
{
    if (sc.nBoxes > 0) return true;
    Hit h;
    h.t = tmax; h.prim = 0x7fffffff; h.u = h.v = 0.f;
    if (sc.nTris > 0) {
        if (SMALL || smallTris) { if (smallSceneTris<true>(smallTris, sc.nBruteTris, o, d, h, -1)) return true; }
        else if (brute ? bruteTris<true>(sc, o, d, h, -1) : traverse<true, COUNT>(sc, o, d, h, -1, sstack, tc)) return true;
    }
    for (int s = 0; s < sc.nSpheres; ++s) {
        const float4 cr = __ldg(sc.spheres + 2 * s);
        const int4 meta = __ldg(reinterpret_cast<const int4*>(sc.spheres + 2 * s + 1));
        float t;
        if (meta.y == 0 && sphereT(cr, o, d, t) && t < tmax) return true;
    }
    return false;
}

// ---------------------------------------------------------------------------------------------------------
// persistent-kernel helpers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t laneId() { return threadIdx.x & 31u; }

// warp-aggregated append: returns the slot for this lane if `want`, one atomic per warp
__device__ __forceinline__ uint32_t warpAppend(uint32_t* counter, bool want)
{
    const uint32_t mask = __ballot_sync(0xffffffffu, want);
    if (mask == 0) return 0;
    uint32_t base = 0;
    const uint32_t leader = __ffs(mask) - 1;
    if (laneId() == leader) base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + __popc(mask & ((1u << laneId()) - 1u));
}

__device__ __forceinline__ void statAdd(unsigned long long* stats, int which, uint32_t v)
{
    // warp-reduce then one atomic
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (laneId() == 0 && v) atomicAdd(stats + which, (unsigned long long)v);
}

// ---------------------------------------------------------------------------------------------------------
// raygen: renderer.cpp:42-52 + PinholeCamera::sampleRay (camera.h:49-60)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cameraRay(const DCamera& c, float u, float v, V3& o, V3& d)
{
    const V3 dir = mk((2 * u - 1) * c.scale, (1 - 2 * v) * c.scale / c.aspect, -1.0f);
    const float* m = c.c2w;
    const V3 w = mk(dir.x * m[0] + dir.y * m[4] + dir.z * m[8], dir.x * m[1] + dir.y * m[5] + dir.z * m[9],
                    dir.x * m[2] + dir.y * m[6] + dir.z * m[10]);
    d = normalize(w);
    o = mk(m[12], m[13], m[14]);
}

// path id = s * nPixels + pixel, so consecutive threads own consecutive pixels of the same sample.
// jitter (optional, device): [(pixel * spp + s) * 2] floats supplied by the parity hook.
__global__ void __launch_bounds__(kBlock) k_raygen(DCamera cam, DQueues q, DWave w, const float* __restrict__ jitter)
{
    const uint32_t n = w.nPaths;
    for (uint32_t pid = blockIdx.x * blockDim.x + threadIdx.x; pid < n; pid += gridDim.x * blockDim.x) {
        const uint32_t pix = pid % w.nPixels, s = pid / w.nPixels;
        const uint32_t i = pix / uint32_t(w.width), j = pix % uint32_t(w.width);
        float r0, r1;
        uint32_t ctr = 0;
        if (jitter) {
            r0 = jitter[(size_t(pix) * w.samplesThisWave + s) * 2];
            r1 = jitter[(size_t(pix) * w.samplesThisWave + s) * 2 + 1];
        }
        else {
            Rng rng;
            rng.open(w, pid, 0);
            r0 = rng.next();
            r1 = rng.next();
            ctr = rng.close();
        }
        const float u = (float(j) + r0) / float(uint32_t(w.width));
        const float v = (float(i) + r1) / float(uint32_t(w.height));
        V3 o, d;
        cameraRay(cam, u, v, o, d);
        q.q0[0][pid] = make_float4(o.x, o.y, o.z, 1.0f);
        q.q1[0][pid] = make_float4(d.x, d.y, d.z, 1.0f);
        q.q2[0][pid] = make_float4(1.0f, __int_as_float(int(pid)), __int_as_float(0), __int_as_float(int(ctr)));
        q.radiance[pid] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) q.ctrl[kCtrlRays] = n;
}

// ---------------------------------------------------------------------------------------------------------
// extend / connect: ONE persistent traversal kernel for closest hit (ANY=false, ray queue -> hit queue) and any
// hit (ANY=true, shadow queue -> radiance). Traversal is a resumable per-lane state machine (one BVH node or one
// leaf per step) so that a warp can REFILL lanes whose ray has finished with fresh rays from the queue while the
// other lanes keep their traversal state: incoherent bounces otherwise run at 8-11 of 32 active threads per
// instruction (ncu, profiles/r01_ncu_full_c3_before_opt.csv).
// ---------------------------------------------------------------------------------------------------------
// per-sample radiance accumulator (one thread owns a path at a time; shadow contributions use atomics)
__device__ __forceinline__ void addRadiance(const DQueues& q, uint32_t pid, V3 c)
{
    float4 r = q.radiance[pid];
    r.x += c.x; r.y += c.y; r.z += c.z;
    q.radiance[pid] = r;
}

constexpr int kSentinel = 0x7fffffff;
constexpr uint32_t kFetchChunk = 64;  // queue entries a warp reserves per atomic (256 costs up to 16 % in tail imbalance)

struct RayState {
    V3 o, d, idir, ood;
    Hit h;      // closest: current best (t, u, v, prim); any: h.t = tmax
    int minId;  // only primitives with id > minId are candidates (BoxMesh overwrite rule)
    int node;   // >= 0 inner node, < 0 leaf code ~((first << 2) | (count - 1)), kSentinel = finished
    int sp;
    uint32_t qidx;
};

__device__ __forceinline__ void stackPush(int* sstack, int* lstack, int& sp, int v)
{
    if (sp < kStackSmem) sstack[sp * kBlock] = v;
    else lstack[sp - kStackSmem] = v;
    ++sp;
}
__device__ __forceinline__ int stackPop(const int* sstack, const int* lstack, int& sp)
{
    if (sp == 0) return kSentinel;
    --sp;
    return (sp < kStackSmem) ? sstack[sp * kBlock] : lstack[sp - kStackSmem];
}

__device__ __forceinline__ void beginTraversal(RayState& r, const DScene& sc)
{
    // box tests only: clamp zero/denormal direction components (see traverse())
    const float kTiny = 1e-20f;
    const V3 ds = mk(fabsf(r.d.x) < kTiny ? copysignf(kTiny, r.d.x) : r.d.x, fabsf(r.d.y) < kTiny ? copysignf(kTiny, r.d.y) : r.d.y,
                     fabsf(r.d.z) < kTiny ? copysignf(kTiny, r.d.z) : r.d.z);
    r.idir = 1.0f / ds;
    r.ood = mk(r.o.x * r.idir.x, r.o.y * r.idir.y, r.o.z * r.idir.z);
    r.sp = 0;
    r.node = sc.nTris > 0 ? 0 : kSentinel;
}

// The two kinds of step of the state machine. nodeStep: one inner node (both children's boxes, nearer child first, farther one
// pushed). leafStep: the 1-4 triangles of one leaf; returns true when ANY and an occluder was found.
__device__ __forceinline__ bool atInner(const RayState& r) { return r.node >= 0 && r.node != kSentinel; }
template <bool COUNT>
__device__ __forceinline__ void nodeStep(const DScene& sc, RayState& r, int* sstack, int* lstack, TraceCounters& tc)
{
    const float4* __restrict__ nodes = sc.nodes;
    const float4 n0 = __ldg(nodes + 4 * r.node), n1 = __ldg(nodes + 4 * r.node + 1), n2 = __ldg(nodes + 4 * r.node + 2);
    const int4 n3 = __ldg(reinterpret_cast<const int4*>(nodes + 4 * r.node + 3));
    if (COUNT) tc.nodes++;
    // both slab tests unconditionally and the decisions as predicates / selects: the short-circuit form compiled into four
    // divergent branches per node (empty slots have count < 0; their inverted boxes must still be masked explicitly)
    float t0n, t1n;
    const bool s0 = slab(n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, r.idir, r.ood, r.h.t, t0n);
    const bool s1 = slab(n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, r.idir, r.ood, r.h.t, t1n);
    const bool h0 = s0 & (n3.z >= 0), h1 = s1 & (n3.w >= 0);
    const int e0 = n3.z > 0 ? ~((n3.x << 2) | (n3.z - 1)) : n3.x;
    const int e1 = n3.w > 0 ? ~((n3.y << 2) | (n3.w - 1)) : n3.y;
    const bool both = h0 & h1;
    const bool swap = both & (t1n < t0n); // nearer child first
    const int nearE = (swap | !h0) ? e1 : e0;
    if (both) stackPush(sstack, lstack, r.sp, swap ? e0 : e1);
    if (h0 | h1) r.node = nearE;
    else r.node = stackPop(sstack, lstack, r.sp);
}
template <bool ANY, bool COUNT>
__device__ __forceinline__ bool leafTest(const DScene& sc, RayState& r, int leafCode, TraceCounters& tc)
{
    const int code = ~leafCode;
    const int first = code >> 2, cnt = (code & 3) + 1;
    const float4* __restrict__ tris = triArray(sc, false);
    for (int i = 0; i < cnt; ++i) {
        if (COUNT) tc.tris++;
        if (triangleRecord<ANY, true>(tris + kTriF4 * (first + i), r.o, r.d, r.h, r.minId)) return true;
    }
    return false;
}

// anyOut != nullptr (parity hook): write the occlusion flag instead of adding the contribution.
template <bool ANY, bool COUNT>
__global__ void __launch_bounds__(kBlock) k_trace(DScene sc, DQueues q, int src, int bounce, int brute, unsigned long long* stats, float4* anyOut,
                                                  int refillThreshold, int stepsPerVote, int leafThreshold)
{
    __shared__ int s_stack[kStackSmem * kBlock];
    int lstack[kStackLocal];
    int* sstack = s_stack + threadIdx.x;
    uint32_t* ctrl = q.ctrl + bounce * kCtrlStride;
    const uint32_t n = ANY ? ctrl[kCtrlShadow] : ctrl[kCtrlRays];
    uint32_t* cursor = ctrl + (ANY ? kCtrlFetchConnect : kCtrlFetchExtend);
    const uint32_t lane = laneId();
    TraceCounters tc;
    RayState r;
    r.node = kSentinel; r.sp = 0; r.qidx = 0; r.minId = -1;
    bool active = false, exhausted = false;
    uint32_t resNext = 0, resEnd = 0; // the warp's current reservation of queue entries
    while (true) {
        // ---- refill the idle lanes with consecutive queue entries (one atomic per warp) ----
        const uint32_t need = __ballot_sync(0xffffffffu, !active);
        if (need != 0 && !exhausted) {
            const uint32_t nNeed = __popc(need);
            // The warp owns a private reservation [resNext, resEnd) of kFetchChunk consecutive queue entries and serves
            // its refills from it: same-address L2 atomics retire at ~1/ns, so one atomic per 32 rays (260 k per 8.3 M-ray
            // launch) was the whole duration of the bounce-0 launch (profiles/r01_notes.md).
            const uint32_t rank = __popc(need & ((1u << lane) - 1u));
            const uint32_t left = resEnd - resNext;
            uint32_t nb = 0;
            if (nNeed > left) { // serve the rest of the old reservation first, the remaining lanes from a new one
                if (lane == 0) nb = atomicAdd(cursor, kFetchChunk);
                nb = __shfl_sync(0xffffffffu, nb, 0);
                if (nb >= n) exhausted = true;
            }
            const uint32_t i = rank < left ? resNext + rank : nb + (rank - left);
            if (nNeed > left) { resNext = nb + (nNeed - left); resEnd = nb + kFetchChunk; }
            else resNext += nNeed;
            if (!active) {
                if (i < n) {
                    r.qidx = i;
                    bool done = false;
                    if (ANY) {
                        const float4 s0 = q.s0[i], s1 = q.s1[i];
                        r.o = xyz(s0); r.d = xyz(s1);
                        r.h.t = s0.w; r.h.u = 0.f; r.h.v = 0.f; r.h.prim = 0; // prim = occluded flag
                        r.minId = -1;
                        if (sc.nBoxes > 0) { r.h.prim = 1; done = true; } // BoxMesh::occluded is always true
                        else if (brute) { r.h.prim = bruteTris<true>(sc, r.o, r.d, r.h, -1) ? 1 : 0; done = true; }
                    }
                    else {
                        const float4 r0 = q.q0[src][i], r1 = q.q1[src][i];
                        r.o = xyz(r0); r.d = xyz(r1);
                        r.h.t = FLT_MAX; r.h.u = 0.f; r.h.v = 0.f; r.h.prim = kSentinel;
                        r.minId = -1;
                        for (int b = 0; b < sc.nBoxes; ++b) { // last box hit in object order wins (primitive.h:259-261)
                            const float4 bl = __ldg(sc.boxes + 2 * b), bh = __ldg(sc.boxes + 2 * b + 1);
                            float t0, t1;
                            if (boxSlabs(xyz(bl), xyz(bh), r.o, r.d, t0, t1)) { r.h.t = t0; r.h.u = t1; r.h.v = 0.f; r.h.prim = __float_as_int(bl.w); r.minId = r.h.prim; }
                        }
                        if (brute) { bruteTris<false>(sc, r.o, r.d, r.h, r.minId); done = true; }
                    }
                    beginTraversal(r, sc);
                    if (done) r.node = kSentinel;
                    active = true;
                }
            }
        }
        if (__ballot_sync(0xffffffffu, active) == 0) break;
        const uint32_t threshold = exhausted ? 1u : uint32_t(refillThreshold);
        // ---- traverse until too few lanes are still busy ----
        // Leaves are POSTPONED: a lane that reaches a leaf waits until at least `leafThreshold` lanes of the warp stand at one (or
        // no lane has an inner node left), then they test their triangles together: run immediately, the triangle code executed
        // at 4 of 32 lanes (about 2 lanes reach a leaf per node step; ncu source view, profiles/r01_notes.md). Measured on the
        // 1 M-triangle scene: threshold 1 / 4 / 8 / 12 / 16 -> 1378 / 1414 / 1376 / 1330 / 1251 Msamples/s; parking the leaf and
        // walking on instead of waiting (speculative traversal) was no better (1366 at best).
        uint32_t busy;
        do {
            for (int sv = 0; sv < stepsPerVote; ++sv) {
                if (active && atInner(r)) nodeStep<COUNT>(sc, r, sstack, lstack, tc);
                const bool atLeaf = active && r.node < 0;
                const uint32_t leafMask = __ballot_sync(0xffffffffu, atLeaf);
                const uint32_t advancing = __ballot_sync(0xffffffffu, active && atInner(r));
                if (leafMask != 0 && (__popc(leafMask) >= leafThreshold || advancing == 0)) {
                    if (atLeaf) {
                        if (leafTest<ANY, COUNT>(sc, r, r.node, tc)) { r.h.prim = 1; r.node = kSentinel; }
                        else r.node = stackPop(sstack, lstack, r.sp);
                    }
                }
                if (active && r.node == kSentinel) {
                    if (ANY) {
                        if (r.h.prim == 0) { // analytic spheres that are not emitter proxies (scene.cpp:206)
                            for (int s = 0; s < sc.nSpheres; ++s) {
                                const float4 cr = __ldg(sc.spheres + 2 * s);
                                const int4 meta = __ldg(reinterpret_cast<const int4*>(sc.spheres + 2 * s + 1));
                                float t;
                                if (meta.y == 0 && sphereT(cr, r.o, r.d, t) && t < r.h.t) { r.h.prim = 1; break; }
                            }
                        }
                        if (anyOut) anyOut[r.qidx] = make_float4(0.f, 0.f, 0.f, __int_as_float(r.h.prim));
                        else if (r.h.prim == 0) {
                            const float4 c = q.s2[r.qidx];
                            float* rad = reinterpret_cast<float*>(q.radiance + __float_as_int(q.s1[r.qidx].w));
                            atomicAdd(rad + 0, c.x); atomicAdd(rad + 1, c.y); atomicAdd(rad + 2, c.z);
                        }
                    }
                    else {
                        for (int s = 0; s < sc.nSpheres; ++s) {
                            const float4 cr = __ldg(sc.spheres + 2 * s);
                            const int4 meta = __ldg(reinterpret_cast<const int4*>(sc.spheres + 2 * s + 1));
                            float t;
                            if (meta.x > r.minId && sphereT(cr, r.o, r.d, t)) consider(r.h, t, 0.f, 0.f, meta.x);
                        }
                        const int prim = r.h.prim == kSentinel ? -1 : r.h.prim;
                        q.hits[r.qidx] = make_float4(prim >= 0 ? r.h.t : FLT_MAX, r.h.u, r.h.v, __int_as_float(prim));
                    }
                    active = false;
                }
            }
            busy = __popc(__ballot_sync(0xffffffffu, active));
        } while (busy >= threshold);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(stats + (ANY ? kStatShadow : kStatClosest), (unsigned long long)n);
    if (COUNT) { statAdd(stats, ANY ? kStatNodesAny : kStatNodes, tc.nodes); statAdd(stats, ANY ? kStatTrisAny : kStatTris, tc.tris); }
}

// parity hook: stage caller-supplied rays into the ray queue (closest) or the shadow queue (any hit)
__global__ void __launch_bounds__(kBlock) k_pack_rays(DQueues q, const float* __restrict__ org, const float* __restrict__ dir,
                                                       const float* __restrict__ tmax, uint32_t n, int anyhit)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 o = make_float4(org[3 * size_t(i)], org[3 * size_t(i) + 1], org[3 * size_t(i) + 2], anyhit ? (tmax ? tmax[i] : FLT_MAX) : 1.f);
        const float4 d = make_float4(dir[3 * size_t(i)], dir[3 * size_t(i) + 1], dir[3 * size_t(i) + 2], __int_as_float(int(i)));
        if (anyhit) { q.s0[i] = o; q.s1[i] = d; }
        else { q.q0[0][i] = o; q.q1[0][i] = d; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) q.ctrl[anyhit ? kCtrlShadow : kCtrlRays] = n;
}

// ---------------------------------------------------------------------------------------------------------
// Simple variants for SHALLOW BVHs (a few hundred nodes, e.g. the 36-triangle Cornell box): one ray per lane run to
// completion with the plain while-while traverse(). On such scenes rays visit ~7 nodes, and the resumable state machine
// of k_trace costs ~30 % more instructions than it recovers in lane utilisation (ncu: bounce 0 297 us vs 231 us,
// profiles/r01_notes.md). Work is reserved kFetchChunk entries at a time per warp.
// ---------------------------------------------------------------------------------------------------------
template <uint32_t CHUNK = kFetchChunk>
__device__ __forceinline__ bool warpNextBatch(uint32_t* cursor, uint32_t n, uint32_t& resNext, uint32_t& resEnd, uint32_t& base)
{
    if (resNext >= resEnd) {
        uint32_t nb = 0;
        if (laneId() == 0) nb = atomicAdd(cursor, CHUNK);
        nb = __shfl_sync(0xffffffffu, nb, 0);
        resNext = nb;
        resEnd = nb + CHUNK;
    }
    base = resNext;
    resNext += 32;
    return base < n;
}

template <bool COUNT>
__global__ void __launch_bounds__(kBlock) k_extend_simple(DScene sc, DQueues q, int src, int bounce, int brute, unsigned long long* stats)
{
    __shared__ int s_stack[kStackSmem * kBlock];
    __shared__ float4 s_tris[kTriF4 * kSmallSceneTris];
    const float4* smallTris = nullptr;
    if (brute == 2 && sc.nBruteTris <= kSmallSceneTris) { // small-scene mode: stage every triangle once per CTA
        for (int k = threadIdx.x; k < kTriF4 * sc.nBruteTris; k += blockDim.x) s_tris[k] = triArray(sc, true)[k];
        __syncthreads();
        smallTris = s_tris;
    }
    uint32_t* ctrl = q.ctrl + bounce * kCtrlStride;
    const uint32_t n = ctrl[kCtrlRays];
    TraceCounters tc;
    uint32_t resNext = 0, resEnd = 0, base;
    while (warpNextBatch(ctrl + kCtrlFetchExtend, n, resNext, resEnd, base)) {
        const uint32_t i = base + laneId();
        if (i < n) {
            const float4 r0 = q.q0[src][i], r1 = q.q1[src][i];
            Hit h;
            closestHit<COUNT>(sc, xyz(r0), xyz(r1), brute == 1, h, s_stack + threadIdx.x, tc, smallTris);
            q.hits[i] = make_float4(h.prim >= 0 ? h.t : FLT_MAX, h.u, h.v, __int_as_float(h.prim));
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(stats + kStatClosest, (unsigned long long)n);
    if (COUNT) { statAdd(stats, kStatNodes, tc.nodes); statAdd(stats, kStatTris, tc.tris); }
}

template <bool COUNT>
__global__ void __launch_bounds__(kBlock) k_connect_simple(DScene sc, DQueues q, int bounce, int brute, unsigned long long* stats)
{
    __shared__ int s_stack[kStackSmem * kBlock];
    __shared__ float4 s_tris[kTriF4 * kSmallSceneTris];
    const float4* smallTris = nullptr;
    if (brute == 2 && sc.nBruteTris <= kSmallSceneTris) {
        for (int k = threadIdx.x; k < kTriF4 * sc.nBruteTris; k += blockDim.x) s_tris[k] = triArray(sc, true)[k];
        __syncthreads();
        smallTris = s_tris;
    }
    uint32_t* ctrl = q.ctrl + bounce * kCtrlStride;
    const uint32_t n = ctrl[kCtrlShadow];
    TraceCounters tc;
    uint32_t resNext = 0, resEnd = 0, base;
    while (warpNextBatch(ctrl + kCtrlFetchConnect, n, resNext, resEnd, base)) {
        const uint32_t i = base + laneId();
        if (i < n) {
            const float4 s0 = q.s0[i], s1 = q.s1[i];
            const bool occ = anyHit<COUNT>(sc, xyz(s0), xyz(s1), s0.w, brute == 1, s_stack + threadIdx.x, tc, smallTris);
            if (!occ) {
                const float4 c = q.s2[i];
                float* r = reinterpret_cast<float*>(q.radiance + __float_as_int(s1.w));
                atomicAdd(r + 0, c.x); atomicAdd(r + 1, c.y); atomicAdd(r + 2, c.z);
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(stats + kStatShadow, (unsigned long long)n);
    if (COUNT) { statAdd(stats, kStatNodesAny, tc.nodes); statAdd(stats, kStatTrisAny, tc.tris); }
}

// Block-aggregated append: ONE global atomic per CTA per call (same-address L2 atomics retire at ~1/ns; with one atomic per
// warp the three queue counters were the whole duration of the bounce-0 shade launch). Must be called by every thread of
// the CTA; `scratch` is kShadeWarps + 1 words of shared memory owned by this call site.
constexpr int kShadeBlock = 256;
constexpr int kShadeWarps = kShadeBlock / 32;
template <int NWARPS = kShadeWarps>
__device__ __forceinline__ uint32_t blockAppend(uint32_t* counter, bool want, uint32_t* scratch)
{
    const uint32_t mask = __ballot_sync(0xffffffffu, want);
    const uint32_t warp = threadIdx.x >> 5, lane = laneId();
    if (lane == 0) scratch[warp] = __popc(mask);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t total = 0;
#pragma unroll
        for (int w = 0; w < NWARPS; ++w) { const uint32_t c = scratch[w]; scratch[w] = total; total += c; }
        scratch[NWARPS] = total ? atomicAdd(counter, total) : 0u;
    }
    __syncthreads();
    const uint32_t slot = scratch[NWARPS] + scratch[warp] + __popc(mask & ((1u << lane) - 1u));
    __syncthreads(); // scratch may be reused by the next call
    return slot;
}
// ---------------------------------------------------------------------------------------------------------
// primary: ray generation (renderer.cpp:42-52, camera.h:49-60) FUSED with the bounce-0 closest hit. Primary rays are
// coherent and most of them miss in the benchmark views (59 % Cornell, 86 % volume), so instead of writing 8.3 M rays,
// reading them back, writing 8.3 M hit records and letting shade skip the misses, this kernel resolves misses inline
// (background 0, DirectIntegrator's 0.18 grey integrator.h:114, Whitted's sky integrator.h:385-389) and appends ONLY the
// hits — ray + hit record — to a COMPACT bounce-0 queue (one atomic per CTA per 128 rays). It also initialises the
// per-path radiance. Static tile partition: CTA b owns path ids [128 b, 128 b + 128), b += grid.
// ---------------------------------------------------------------------------------------------------------
template <bool COUNT>
__global__ void __launch_bounds__(kBlock) k_primary(DScene sc, DCamera cam, DQueues q, DWave w, int brute, int missMode, unsigned long long* stats)
{
    __shared__ int s_stack[kStackSmem * kBlock];
    __shared__ uint32_t s_scratch[kBlock / 32 + 1];
    const uint32_t n = w.nPaths;
    TraceCounters tc;
    uint32_t nHits = 0;
    for (uint32_t tile = blockIdx.x; uint64_t(tile) * kBlock < n; tile += gridDim.x) {
        const uint32_t pid = tile * kBlock + threadIdx.x;
        bool hit = false;
        V3 o = mk(0.f), d = mk(0.f);
        Hit h{FLT_MAX, 0.f, 0.f, -1};
        uint32_t ctr = 0;
        if (pid < n) {
            const uint32_t pix = pid % w.nPixels;
            const uint32_t i = pix / uint32_t(w.width), j = pix % uint32_t(w.width);
            Rng rng;
            rng.open(w, pid, 0);
            const float r0 = rng.next();
            const float r1 = rng.next();
            ctr = rng.close();
            const float u = (float(j) + r0) / float(uint32_t(w.width));
            const float v = (float(i) + r1) / float(uint32_t(w.height));
            cameraRay(cam, u, v, o, d);
            closestHit<COUNT>(sc, o, d, brute != 0, h, s_stack + threadIdx.x, tc);
            hit = h.prim >= 0;
            V3 c = mk(0.f);
            if (!hit) {
                if (missMode == 1) c = mk(float(0.18));
                else if (missMode == 2) c = mk(1.f) * mk(float(0.235294), float(0.67451), float(0.843137));
            }
            q.radiance[pid] = make_float4(c.x, c.y, c.z, 0.f);
        }
        const uint32_t slot = blockAppend<kBlock / 32>(q.ctrl + kCtrlRays, hit, s_scratch);
        nHits += hit ? 1u : 0u;
        if (hit) {
            q.q0[0][slot] = make_float4(o.x, o.y, o.z, 1.0f);
            q.q1[0][slot] = make_float4(d.x, d.y, d.z, 1.0f);
            q.q2[0][slot] = make_float4(1.0f, __int_as_float(int(pid)), __int_as_float(0), __int_as_float(int(ctr)));
            q.hits[slot] = make_float4(h.t, h.u, h.v, __int_as_float(h.prim));
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(stats + kStatClosest, (unsigned long long)n);
    statAdd(stats, kStatPrimaryHits, nHits);
    if (COUNT) { statAdd(stats, kStatNodes, tc.nodes); statAdd(stats, kStatTris, tc.tris); }
}

// ---------------------------------------------------------------------------------------------------------
// shading
// ---------------------------------------------------------------------------------------------------------
struct Surf {
    V3 pos, ng, ns, dpdu, dpdv, albedo;
    uint32_t meta;
    float t1;
};

// Reconstructs what Mesh::intersect (primitive.cpp:100-110) / Sphere::intersect (primitive.h:112-122) store in
// IntersectInfo. ng was normalised on the host with the reference's expression; ns is interpolated and NOT
// re-normalised. A sphere hit leaves dpdu/dpdv at zero (the reference leaves them stale).
__device__ __forceinline__ void makeSurf(const DScene& sc, V3 o, V3 d, const Hit& h, Surf& s)
{
    const float4 p3 = __ldg(sc.prims + 4 * h.prim + 3);
    s.meta = __float_as_uint(p3.w);
    s.albedo = xyz(p3);
    s.pos = o + h.t * d;
    s.t1 = h.u;
    const uint32_t kind = s.meta & kMetaKindMask;
    if (kind == XRTG_OBJ_MESH) {
        const float4 p0 = __ldg(sc.prims + 4 * h.prim), p1 = __ldg(sc.prims + 4 * h.prim + 1), p2 = __ldg(sc.prims + 4 * h.prim + 2);
        s.ng = mk(p0.w, p1.w, p2.w);
        s.ns = xyz(p0) * (1.0f - h.u - h.v) + xyz(p1) * h.u + xyz(p2) * h.v;
        orthonormalBasis(s.ns, s.dpdu, s.dpdv);
    }
    else if (kind == XRTG_OBJ_SPHERE) {
        const float4 p0 = __ldg(sc.prims + 4 * h.prim);
        s.ng = normalize(s.pos - xyz(p0));
        s.ns = s.ng;
        s.dpdu = mk(0.f); s.dpdv = mk(0.f);
    }
    else {
        s.ng = mk(0.f); s.ns = mk(0.f); s.dpdu = mk(0.f); s.dpdv = mk(0.f);
    }
}
__device__ __forceinline__ bool hasMaterial(const Surf& s) { return (s.meta & kMetaHasMaterial) != 0; }
__device__ __forceinline__ int lightOf(const Surf& s) { return int((s.meta >> kMetaLightShift) & 0xfffu) - 1; }
__device__ __forceinline__ int mediumOf(const Surf& s) { return int((s.meta >> kMetaMediumShift) & 0xfffu) - 1; }

// AreaLight::Le (light.h:62-69)
__device__ __forceinline__ V3 emitted(const DScene& sc, const Surf& s, V3 rayDir)
{
    const int li = lightOf(s);
    if (li < 0) return mk(0.f);
    return (dot(rayDir, s.ns) < 0) ? xyz(sc.lights[li].Le) : mk(0.f);
}

// AreaLight::sample: quad light.cpp:59-68, triangle light.cpp:21-30 + :43-47, sphere light.h:158-197.
// Draw order as compiled by g++ (second-written getNext1D() draws first), pinned by the oracle KATs.
__device__ __forceinline__ V3 sampleLight(const DLight& L, V3 position, V3& wi, float& pdf, float& tmax, Rng& rng)
{
    const int kind = __float_as_int(L.v0_kind.w);
    const V3 v0 = xyz(L.v0_kind);
    if (kind == XRTG_LIGHT_QUAD) {
        const float rb = rng.next();
        const float ra = rng.next();
        const V3 d = (v0 + xyz(L.e1_r) * ra + xyz(L.e2) * rb) - position;
        tmax = length(d);
        const float dn = dot(d, xyz(L.Ng));
        if (dn >= 0) return mk(0.f);
        wi = d / tmax;
        pdf = (tmax * tmax * tmax) / fabsf(dn);
        return xyz(L.Le);
    }
    if (kind == XRTG_LIGHT_TRIANGLE) {
        const float v = rng.next();
        const float u = rng.next();
        const float su = sqrtf(u);
        const V3 A = v0, B = xyz(L.v1), C = xyz(L.v2);
        const V3 p = C + (1.f - su) * (A - C) + (v * su) * (B - C);
        const V3 d = p - position;
        tmax = length(d);
        const float dn = dot(d, xyz(L.Ng));
        if (dn >= 0) return mk(0.f);
        wi = d / tmax;
        pdf = (2.f * tmax * tmax * tmax) / fabsf(dn);
        return xyz(L.Le);
    }
    const float radius = L.e1_r.w;
    V3 dz = v0 - position;
    const float dz_len_2 = dot(dz, dz);
    const float dz_len = sqrtf(dz_len_2);
    dz = dz / mk(-dz_len);
    V3 dx, dy;
    orthonormalBasis(dz, dx, dy);
    const float sin_theta_max_2 = radius * radius / dz_len_2;
    const float sin_theta_max = sqrtf(sin_theta_max_2);
    const float cos_theta_max = sqrtf(smax(0.f, 1.f - sin_theta_max_2));
    const float cos_theta = 1 + (cos_theta_max - 1) * rng.next();
    const float sin_theta_2 = 1.f - cos_theta * cos_theta;
    const float cos_alpha = sin_theta_2 / sin_theta_max + cos_theta * sqrtf(smax(0.0f, 1 - sin_theta_2 / sin_theta_max_2));
    const float sin_alpha = sqrtf(smax(0.0f, 1 - cos_alpha * cos_alpha));
    const float phi = 2 * kPI * rng.next();
    const V3 n = cosf(phi) * sin_alpha * dx + sinf(phi) * sin_alpha * dy + cos_alpha * dz;
    const V3 p = v0 + n * radius;
    const V3 d = p - position;
    tmax = length(d);
    const float d_dot_n = dot(d, n);
    if (d_dot_n >= 0) return mk(0.f);
    pdf = 1.f / (2.f * kPI * (1.f - cos_theta_max));
    wi = d / tmax;
    return xyz(L.Le);
}

// Lambert::sampleDir (material.h:55-73): UNIFORM hemisphere about ng using ns's tangent frame, pdf = 1/2PI
__device__ __forceinline__ V3 lambertSampleDir(const Surf& s, Rng& rng, float& pdf)
{
    const float r1 = rng.next();
    const float r2 = rng.next();
    pdf = 1 / (2 * kPI);
    const float sinTheta = sqrtf(1 - r1 * r1);
    const float phi = 2 * kPI * r2;
    const float x = sinTheta * cosf(phi);
    const float z = sinTheta * sinf(phi);
    return localToWorld(mk(x, r1, z), s.dpdu, s.ng, s.dpdv);
}
__device__ __forceinline__ V3 evalBxDF(const Surf& s) { return hasMaterial(s) ? s.albedo / kPI : mk(0.f); }
__device__ __forceinline__ V3 sampleBxDF(const Surf& s, Rng& rng, V3& wi, float& pdf)
{
    if (!hasMaterial(s)) return mk(0.f);
    wi = lambertSampleDir(s, rng, pdf);
    return evalBxDF(s);
}

struct ShadeOut {
    DQueues q;
    uint32_t* ctrlNext; // ctrl block of bounce+1 (ray count)
    uint32_t* ctrlCur;  // ctrl block of this bounce (shadow count)
    int dst;
};

__device__ __forceinline__ void pushShadow(const ShadeOut& so, bool want, V3 o, V3 d, float tmax, uint32_t pid, V3 c, uint32_t* scratch)
{
    const uint32_t slot = blockAppend(so.ctrlCur + kCtrlShadow, want, scratch);
    if (want) {
        so.q.s0[slot] = make_float4(o.x, o.y, o.z, tmax);
        so.q.s1[slot] = make_float4(d.x, d.y, d.z, __int_as_float(int(pid)));
        so.q.s2[slot] = make_float4(c.x, c.y, c.z, 0.f);
    }
}
// warp-aggregated variant (one atomic per warp) for the volume kernel, whose warps run independently
__device__ __forceinline__ void pushRayWarp(const ShadeOut& so, bool want, V3 o, V3 d, V3 T, uint32_t pid, int depth, uint32_t ctr)
{
    const uint32_t slot = warpAppend(so.ctrlNext + kCtrlRays, want);
    if (want) {
        so.q.q0[so.dst][slot] = make_float4(o.x, o.y, o.z, T.x);
        so.q.q1[so.dst][slot] = make_float4(d.x, d.y, d.z, T.y);
        so.q.q2[so.dst][slot] = make_float4(T.z, __int_as_float(int(pid)), __int_as_float(depth), __int_as_float(int(ctr)));
    }
}
__device__ __forceinline__ void pushRay(const ShadeOut& so, bool want, V3 o, V3 d, V3 T, uint32_t pid, int depth, uint32_t ctr, uint32_t* scratch)
{
    const uint32_t slot = blockAppend(so.ctrlNext + kCtrlRays, want, scratch);
    if (want) {
        so.q.q0[so.dst][slot] = make_float4(o.x, o.y, o.z, T.x);
        so.q.q1[so.dst][slot] = make_float4(d.x, d.y, d.z, T.y);
        so.q.q2[so.dst][slot] = make_float4(T.z, __int_as_float(int(pid)), __int_as_float(depth), __int_as_float(int(ctr)));
    }
}

// Surface integrators: Normal (integrator.h:29-36), furnace (:59-66), Direct (:82-119), Indirect (:129-186),
// GI (:205-287), Whitted's Lambert/delta-light branch (:302-394). One thread per ray-queue entry of bounce b.
__global__ void __launch_bounds__(kShadeBlock, 4) k_shade_surface(DScene sc, DQueues q, DWave w, int src, int bounce)
{
    __shared__ uint32_t s_scratch[kShadeWarps + 1];
    uint32_t* ctrl = q.ctrl + bounce * kCtrlStride;
    const uint32_t n = ctrl[kCtrlRays];
    ShadeOut so{q, ctrl + kCtrlStride, ctrl, src ^ 1};
    const int kind = w.integrator;
    // static partition: CTA b owns tiles b, b+grid, ... of kShadeBlock consecutive queue entries (uniform cost per entry,
    // no work-fetch atomics); the trip count is uniform across the CTA, as the block-level appends require
    for (uint32_t tile = blockIdx.x; uint64_t(tile) * kShadeBlock < n; tile += gridDim.x) {
        const uint32_t i = tile * kShadeBlock + threadIdx.x;
        const bool live = i < n;
        // per-lane outputs, appended collectively at the end of the iteration
        bool wantRay = false;
        V3 no = mk(0.f), nd = mk(0.f), nT = mk(0.f);
        uint32_t pid = 0, ctr = 0;
        int depth = 0;
        // hit record + path word first; the 32 B origin/direction only for rays that hit something (41 % of the primary
        // rays of the 1080p Cornell view): a miss costs 32 B instead of 64 B of HBM reads
        float4 r0, r1, r2, hv;
        r0 = r1 = r2 = hv = make_float4(0, 0, 0, 0);
        hv.w = __int_as_float(-1);
        if (live) {
            hv = q.hits[i]; r2 = q.q2[src][i];
            if (__float_as_int(hv.w) >= 0) { r0 = q.q0[src][i]; r1 = q.q1[src][i]; }
        }
        const V3 o = xyz(r0), d = xyz(r1);
        V3 T = mk(r0.w, r1.w, r2.x);
        pid = uint32_t(__float_as_int(r2.y));
        depth = __float_as_int(r2.z);
        Hit h{hv.x, hv.y, hv.z, __float_as_int(hv.w)};
        Rng rng;
        bool shadeLights = false, shadeDelta = false;
        Surf s = {};
        if (live) {
            rng.open(w, pid, uint32_t(__float_as_int(r2.w)));
            if (h.prim < 0) {
                if (kind == XRTG_INT_DIRECT) addRadiance(q, pid, mk(float(0.18)));
                else if (kind == XRTG_INT_WHITTED) addRadiance(q, pid, mk(1.f) * mk(float(0.235294), float(0.67451), float(0.843137)));
            }
            else {
                makeSurf(sc, o, d, h, s);
                if (kind == XRTG_INT_NORMAL) {
                    addRadiance(q, pid, 0.5f * (s.ns + 1.0f));
                }
                else if (kind == XRTG_INT_FURNACE) {
                    float pdf = 1.0f;
                    V3 nextDir = mk(0.f);
                    const V3 fr = sampleBxDF(s, rng, nextDir, pdf);
                    const float cs = smax(0.0f, dot(nextDir, s.ng));
                    addRadiance(q, pid, fr * cs * mk(1.0f) / pdf);
                }
                else if (kind == XRTG_INT_DIRECT) {
                    if (lightOf(s) >= 0) addRadiance(q, pid, emitted(sc, s, d));
                    else shadeLights = true;
                }
                else if (kind == XRTG_INT_WHITTED) {
                    shadeDelta = hasMaterial(s);
                }
                else { // Indirect / GI
                    bool alive = true;
                    if (depth > 0) { // russian roulette (integrator.h:223-231)
                        const float p = smin((T.x + T.y + T.z) / 3.0f, 1.0f);
                        if (rng.next() >= p) alive = false;
                        else T = T / mk(p);
                    }
                    if (alive && lightOf(s) >= 0) {
                        if (kind == XRTG_INT_INDIRECT || depth == 0) addRadiance(q, pid, T * emitted(sc, s, d));
                        alive = false;
                    }
                    if (alive) {
                        shadeLights = (kind == XRTG_INT_GI);
                        wantRay = true; // BSDF sampling happens after the light loop (draw order!)
                    }
                }
            }
        }
        // ---- NEE over EVERY area light (integrator.h:95-108, :250-267) ----
        if (kind == XRTG_INT_DIRECT || kind == XRTG_INT_GI) { // CTA-uniform: every thread takes part in the block appends
            for (int li = 0; li < sc.nLights; ++li) {
                bool want = false;
                V3 wi = mk(0.f), c = mk(0.f);
                float tmax = 0.f;
                if (shadeLights) {
                    float pdf = 0.0f;
                    const V3 Lr = sampleLight(sc.lights[li], s.pos, wi, pdf, tmax, rng);
                    if (pdf != 0) {
                        const float cs = smax(0.0f, dot(s.ng, wi));
                        const V3 fr = evalBxDF(s);
                        c = T * (fr * Lr * cs / pdf);
                        want = true;
                    }
                }
                const float bias = 0.01f;
                pushShadow(so, want, s.pos + s.ng * bias, wi, tmax - bias, pid, c, s_scratch);
            }
        }
        // ---- Whitted diffuse term over delta lights (integrator.h:328-343; PointLight/DistantLight light.cpp:120-142)
        if (kind == XRTG_INT_WHITTED) {
            for (int li = 0; li < sc.nDelta; ++li) {
                bool want = false;
                V3 wi = mk(0.f), c = mk(0.f);
                float tmax = 0.f;
                if (shadeDelta) {
                    const DDelta L = sc.dlights[li];
                    float pdf;
                    if (__float_as_int(L.p_kind.w) == XRTG_DLIGHT_POINT) {
                        const V3 ld = xyz(L.p_kind) - s.pos;
                        const float dist = length(ld);
                        wi = ld / dist; pdf = dist * dist; tmax = dist;
                    }
                    else { wi = -xyz(L.p_kind); pdf = 1.0f; tmax = FLT_MAX; }
                    c = evalBxDF(s) * xyz(L.L) * smax(0.f, dot(s.ns, wi)) / pdf;
                    want = true;
                }
                pushShadow(so, want, s.pos + s.ng * float(0.1), wi, tmax, pid, c, s_scratch);
            }
        }
        // ---- BSDF bounce (integrator.h:271-283) ----
        if (wantRay) {
            float pdf = 1.0f;
            V3 nextDir = mk(0.f);
            const V3 fr = sampleBxDF(s, rng, nextDir, pdf);
            const float cs = smax(.0f, dot(nextDir, s.ng));
            nT = T * (fr * cs / pdf);
            no = s.pos + s.ng * 0.01f;
            nd = nextDir;
            wantRay = (depth + 1 < w.maxDepth);
        }
        if (live) ctr = rng.close();
        if (kind == XRTG_INT_INDIRECT || kind == XRTG_INT_GI) pushRay(so, wantRay, no, nd, nT, pid, depth + 1, ctr, s_scratch);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Small scenes (<= kSmallSceneTris triangles — every scene the reference ships): ONE kernel per bounce that shades the hit,
// traces the NEE shadow rays, samples the BSDF, traces the next closest hit and applies the NEXT depth's Russian roulette and
// emitter test, all against the triangle list in shared memory. Only paths that go on to shade at depth+1 are appended
// (ray + hit record, one atomic per CTA per 128 paths), so every lane that enters the kernel does useful work in every phase:
// the shadow queue, the separate connect / extend launches, the hit-record round trip and the radiance atomics of the
// three-kernel pipeline disappear (per path and bounce: 64 B in, <= 64 B out, one 16 B radiance read-modify-write).
// The per-path draw order is the reference's: [RR] -> light samples -> BSDF sample (integrator.h:223-283); the RR draw of
// depth+1 simply happens at the end of depth's kernel. Hit records are double-buffered (q.hits / q.s0) because CTAs append
// to bounce b+1 while others still read bounce b.
// ---------------------------------------------------------------------------------------------------------
// Stages the scene's triangle list (primitive-id order) in shared memory; the throughput instantiation also builds the list of
// OCCLUDERS (everything that is not an emitter proxy, scene.cpp:206) so the shadow loop carries no per-triangle flag test.
__device__ __forceinline__ void stageSmallScene(const DScene& sc, float4* s_tris, float4* s_occ, int* s_nOcc)
{
    const float4* __restrict__ src = triArray(sc, true);
    for (int k = threadIdx.x; k < kTriF4 * sc.nBruteTris; k += blockDim.x) s_tris[k] = src[k];
    if constexpr (kExact) { if (threadIdx.x == 0) *s_nOcc = sc.nBruteTris; }
    else if (threadIdx.x < 32) { // warp 0: order-preserving compaction, 32 triangles per round
        int nOcc = 0;
        for (int base = 0; base < sc.nBruteTris; base += 32) {
            const int i = base + int(threadIdx.x);
            const bool keep = i < sc.nBruteTris && (__float_as_int(src[4 * i + 3].y) & 1) == 0;
            const uint32_t m = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                const int slot = nOcc + __popc(m & ((1u << threadIdx.x) - 1u));
                for (int k = 0; k < 4; ++k) s_occ[4 * slot + k] = src[4 * i + k];
            }
            nOcc += __popc(m);
        }
        if (threadIdx.x == 0) *s_nOcc = nOcc;
    }
    __syncthreads();
}
// Scene::occluded (scene.cpp:202-211) on a staged small scene
__device__ __forceinline__ bool anyHitSmall(const DScene& sc, V3 o, V3 d, float tmax, const float4* occTris, int nOcc)
{
    if (sc.nBoxes > 0) return true; // BoxMesh::occluded is always true (primitive.h:266-268)
    Hit h;
    h.t = tmax; h.prim = 0x7fffffff; h.u = h.v = 0.f;
    if (smallSceneTris<true, !kExact>(occTris, nOcc, o, d, h, -1)) return true;
    for (int s = 0; s < sc.nSpheres; ++s) {
        const float4 cr = __ldg(sc.spheres + 2 * s);
        const int4 meta = __ldg(reinterpret_cast<const int4*>(sc.spheres + 2 * s + 1));
        float t;
        if (meta.y == 0 && sphereT(cr, o, d, t) && t < tmax) return true;
    }
    return false;
}
__device__ __forceinline__ float4* hitBuffer(const DQueues& q, int bounce) { return (bounce & 1) ? q.s0 : q.hits; }

// Plane-paired small scene (small_scene.h), throughput instantiation: one record = one supporting plane + two triangles in it;
// one ray/plane intersection and two barycentric plane equations per triangle. Branch-free per lane.
struct SmallSection {
    const float4* recs; // 5 per record: N|d , A: n1|d1 , n2|d2 , B: n1|d1 , n2|d2
    const int* ids;     // 2 per record
    int nRecords;
};
__device__ __forceinline__ SmallSection smallSection(const float4* blk, int off)
{
    const int4 h = *reinterpret_cast<const int4*>(blk + off);
    return SmallSection{blk + h.z, reinterpret_cast<const int*>(blk + h.w), h.x};
}
__device__ __forceinline__ float insideness(const float4 a, const float4 b, V3 P)
{
    const float u = fmaf(P.x, a.x, fmaf(P.y, a.y, fmaf(P.z, a.z, a.w)));
    const float v = fmaf(P.x, b.x, fmaf(P.y, b.y, fmaf(P.z, b.z, b.w)));
    return fminf(fminf(u, v), 1.f - (u + v)); // >= 0 <=> u >= 0, v >= 0, u + v <= 1
}
// Any hit with eps < t < tmax among the occluders; lanes without a shadow ray pass tmax < 0. Must be called by all 32 lanes of
// a converged warp: a plane that no lane can hit (behind every ray or beyond every tmax — the floor and the ceiling for every
// shadow ray towards the Cornell light) is skipped with one vote.
__device__ __forceinline__ bool groupedAnyHit(const SmallSection& S, V3 o, V3 d, float tmax)
{
    float acc = -1.f;
    for (int r = 0; r < S.nRecords; ++r) {
        const float4* rec = S.recs + 5 * r;
        const float4 pl = rec[0];
        const float det = dot(xyz(pl), d);
        const float t = fmaf(-o.x, pl.x, fmaf(-o.y, pl.y, fmaf(-o.z, pl.z, pl.w))) * (1.0f / det);
        const bool vt = !(fabsf(det) < FLT_EPSILON) && t > FLT_EPSILON && t < tmax;
        if (!__any_sync(0xffffffffu, vt)) continue;
        const V3 P = o + t * d;
        const float m = fmaxf(insideness(rec[1], rec[2], P), insideness(rec[3], rec[4], P));
        acc = fmaxf(acc, vt ? m : -1.f);
    }
    return acc >= 0.f;
}
// Closest hit: strictly smaller t wins between records, A before B inside one (scene.cpp:193-197); lanes without a ray pass
// want = false.
__device__ __forceinline__ void groupedClosest(const SmallSection& S, V3 o, V3 d, bool want, Hit& h)
{
    float best = want ? FLT_MAX : -1.f;
    int bi = -1;
#pragma unroll 2
    for (int r = 0; r < S.nRecords; ++r) {
        const float4* rec = S.recs + 5 * r;
        const float4 pl = rec[0];
        const float det = dot(xyz(pl), d);
        const float t = fmaf(-o.x, pl.x, fmaf(-o.y, pl.y, fmaf(-o.z, pl.z, pl.w))) * (1.0f / det);
        const V3 P = o + t * d;
        const float m0 = insideness(rec[1], rec[2], P), m1 = insideness(rec[3], rec[4], P);
        const bool take = fmaxf(m0, m1) >= 0.f && !(fabsf(det) < FLT_EPSILON) && t > FLT_EPSILON && t < best;
        best = take ? t : best;
        bi = take ? (m0 >= 0.f ? 2 * r : 2 * r + 1) : bi;
    }
    if (bi >= 0) {
        const float4* rec = S.recs + 5 * (bi >> 1) + 1 + 2 * (bi & 1);
        const float4 a = rec[0], b = rec[1];
        const V3 P = o + best * d;
        h.t = best;
        h.u = fmaf(P.x, a.x, fmaf(P.y, a.y, fmaf(P.z, a.z, a.w)));
        h.v = fmaf(P.x, b.x, fmaf(P.y, b.y, fmaf(P.z, b.z, b.w)));
        h.prim = S.ids[bi];
    }
}

// GROUPED: the scene carries a plane-paired block (throughput instantiation, no BoxMesh); otherwise the per-triangle lists.
template <bool GROUPED>
__global__ void __launch_bounds__(kBlock, 5) k_bounce_small(DScene sc, DQueues q, DWave w, int src, int bounce, unsigned long long* stats)
{
    constexpr int kListF4 = GROUPED ? 1 : kTriF4 * kSmallSceneTris;
    __shared__ float4 s_tris[kListF4];
    __shared__ float4 s_occ[(kExact || GROUPED) ? 1 : kTriF4 * kSmallSceneTris];
    __shared__ float4 s_block[GROUPED ? kSmallBlockF4 : 1];
    __shared__ uint32_t s_scratch[kBlock / 32 + 1];
    __shared__ int s_nOcc;
    SmallSection secAll{}, secOcc{};
    const float4* occTris = nullptr;
    int nOcc = 0;
    if constexpr (GROUPED) {
        for (int k = threadIdx.x; k < sc.smallBlockF4; k += blockDim.x) s_block[k] = sc.smallBlock[k];
        __syncthreads();
        const int4 hd = *reinterpret_cast<const int4*>(s_block);
        secAll = smallSection(s_block, hd.x);
        secOcc = smallSection(s_block, hd.y);
    }
    else {
        stageSmallScene(sc, s_tris, s_occ, &s_nOcc);
        occTris = kExact ? s_tris : s_occ;
        nOcc = s_nOcc;
    }
    // Scene::occluded / Scene::intersect on the staged scene. GROUPED: executed by the whole (converged) warp.
    auto occluded = [&](bool want, V3 o, V3 d, float tmax) -> bool {
        if constexpr (GROUPED) {
            bool occ = groupedAnyHit(secOcc, o, d, want ? tmax : -1.f);
            if (want && !occ)
                for (int k = 0; k < sc.nSpheres; ++k) {
                    const float4 cr = __ldg(sc.spheres + 2 * k);
                    const int4 meta = __ldg(reinterpret_cast<const int4*>(sc.spheres + 2 * k + 1));
                    float t;
                    if (meta.y == 0 && sphereT(cr, o, d, t) && t < tmax) { occ = true; break; }
                }
            return occ;
        }
        else return want && anyHitSmall(sc, o, d, tmax, occTris, nOcc);
    };
    auto closest = [&](bool want, V3 o, V3 d, Hit& h) {
        if constexpr (GROUPED) {
            h.t = FLT_MAX; h.u = 0.f; h.v = 0.f; h.prim = 0x7fffffff;
            groupedClosest(secAll, o, d, want, h);
            if (want)
                for (int k = 0; k < sc.nSpheres; ++k) {
                    const float4 cr = __ldg(sc.spheres + 2 * k);
                    const int4 meta = __ldg(reinterpret_cast<const int4*>(sc.spheres + 2 * k + 1));
                    float t;
                    if (sphereT(cr, o, d, t)) consider(h, t, 0.f, 0.f, meta.x);
                }
            if (h.prim == 0x7fffffff) h.prim = -1;
        }
        else if (want) {
            TraceCounters tc;
            closestHit<false, true>(sc, o, d, false, h, nullptr, tc, s_tris);
        }
    };

    uint32_t* ctrl = q.ctrl + bounce * kCtrlStride;
    const uint32_t n = ctrl[kCtrlRays];
    uint32_t* nextCount = ctrl + kCtrlStride + kCtrlRays;
    const float4* __restrict__ hitsIn = hitBuffer(q, bounce);
    float4* __restrict__ hitsOut = hitBuffer(q, bounce + 1);
    // (ternaries instead of q.q0[src]: a dynamic index would force a local-memory copy of the kernel parameter)
    const float4* __restrict__ in0 = src ? q.q0[1] : q.q0[0];
    const float4* __restrict__ in1 = src ? q.q1[1] : q.q1[0];
    const float4* __restrict__ in2 = src ? q.q2[1] : q.q2[0];
    float4* __restrict__ out0 = src ? q.q0[0] : q.q0[1];
    float4* __restrict__ out1 = src ? q.q1[0] : q.q1[1];
    float4* __restrict__ out2 = src ? q.q2[0] : q.q2[1];
    const int kind = w.integrator;
    uint32_t nClosest = 0, nShadow = 0;
    for (uint32_t tile = blockIdx.x; uint64_t(tile) * kBlock < n; tile += gridDim.x) {
        const uint32_t i = tile * kBlock + threadIdx.x;
        const bool live = i < n;
        // ---- phase 1: load the path, rebuild the surface, emitter test of depth 0 ----
        V3 d = mk(0.f), T = mk(0.f);
        uint32_t pid = 0, ctr = 0;
        int depth = 0;
        float4 rad = make_float4(0.f, 0.f, 0.f, 0.f);
        bool radDirty = false;
        auto add = [&](V3 c) { rad.x += c.x; rad.y += c.y; rad.z += c.z; radDirty = true; };
        Rng rng;
        Surf s = {};
        bool shadeLights = false, shadeDelta = false, bsdf = false;
        if (live) {
            const float4 hv = hitsIn[i], r0 = in0[i], r1 = in1[i], r2 = in2[i];
            const V3 o = xyz(r0);
            d = xyz(r1);
            T = mk(r0.w, r1.w, r2.x);
            pid = uint32_t(__float_as_int(r2.y));
            depth = __float_as_int(r2.z);
            rad = q.radiance[pid];
            const Hit h{hv.x, hv.y, hv.z, __float_as_int(hv.w)};
            rng.open(w, pid, uint32_t(__float_as_int(r2.w)));
            makeSurf(sc, o, d, h, s);
            if (kind == XRTG_INT_DIRECT) {
                if (lightOf(s) >= 0) add(emitted(sc, s, d));
                else shadeLights = true;
            }
            else if (kind == XRTG_INT_WHITTED) shadeDelta = hasMaterial(s);
            else { // Indirect / GI. Entries of bounce > 0 already passed RR and the emitter test in the kernel that traced them.
                bool alive = true;
                if (bounce == 0 && lightOf(s) >= 0) { add(T * emitted(sc, s, d)); alive = false; }
                shadeLights = alive && kind == XRTG_INT_GI;
                bsdf = alive;
            }
        }
        // ---- phase 2: NEE over EVERY area light (integrator.h:95-108, :250-267), shadow ray traced inline ----
        if (kind == XRTG_INT_DIRECT || kind == XRTG_INT_GI) {
            for (int li = 0; li < sc.nLights; ++li) {
                bool want = false;
                V3 wi = mk(0.f), c = mk(0.f);
                float tmax = 0.f;
                if (shadeLights) {
                    float pdf = 0.0f;
                    const V3 Lr = sampleLight(sc.lights[li], s.pos, wi, pdf, tmax, rng);
                    if (pdf != 0) {
                        const float cs = smax(0.0f, dot(s.ng, wi));
                        const V3 fr = evalBxDF(s);
                        c = T * (fr * Lr * cs / pdf);
                        want = true;
                        ++nShadow;
                    }
                }
                const float bias = 0.01f;
                if (!occluded(want, s.pos + s.ng * bias, wi, tmax - bias) && want) add(c);
            }
        }
        // ---- Whitted diffuse term over delta lights (integrator.h:328-343; light.cpp:120-142) ----
        if (kind == XRTG_INT_WHITTED) {
            for (int li = 0; li < sc.nDelta; ++li) {
                V3 wi = mk(0.f), c = mk(0.f);
                float tmax = 0.f;
                if (shadeDelta) {
                    const DDelta L = sc.dlights[li];
                    float pdf;
                    if (__float_as_int(L.p_kind.w) == XRTG_DLIGHT_POINT) {
                        const V3 ld = xyz(L.p_kind) - s.pos;
                        const float dist = length(ld);
                        wi = ld / dist; pdf = dist * dist; tmax = dist;
                    }
                    else { wi = -xyz(L.p_kind); pdf = 1.0f; tmax = FLT_MAX; }
                    c = evalBxDF(s) * xyz(L.L) * smax(0.f, dot(s.ns, wi)) / pdf;
                    ++nShadow;
                }
                if (!occluded(shadeDelta, s.pos + s.ng * float(0.1), wi, tmax) && shadeDelta) add(c);
            }
        }
        // ---- phase 3: BSDF bounce (integrator.h:271-283), then intersect + RR + emitter test of depth+1 (integrator.h:214-245) ----
        bool wantNext = false, trace = false;
        V3 no = mk(0.f), nd = mk(0.f), nT = mk(0.f);
        Hit nh{FLT_MAX, 0.f, 0.f, -1};
        if (bsdf) {
            float pdf = 1.0f;
            V3 nextDir = mk(0.f);
            const V3 fr = sampleBxDF(s, rng, nextDir, pdf);
            const float cs = smax(.0f, dot(nextDir, s.ng));
            nT = T * (fr * cs / pdf);
            no = s.pos + s.ng * 0.01f;
            nd = nextDir;
            trace = depth + 1 < w.maxDepth;
        }
        if (kind == XRTG_INT_INDIRECT || kind == XRTG_INT_GI) closest(trace, no, nd, nh);
        if (trace) {
            ++nClosest;
            if (nh.prim >= 0) {
                const float p = smin((nT.x + nT.y + nT.z) / 3.0f, 1.0f);
                if (!(rng.next() >= p)) {
                    nT = nT / mk(p);
                    const uint32_t meta = __float_as_uint(__ldg(sc.prims + 4 * nh.prim + 3).w);
                    if (((meta >> kMetaLightShift) & 0xfffu) == 0) wantNext = true;
                    else if (kind == XRTG_INT_INDIRECT) { // Le at any depth (integrator.h:150-160); GI only at depth 0
                        Surf s2;
                        makeSurf(sc, no, nd, nh, s2);
                        add(nT * emitted(sc, s2, nd));
                    }
                }
            }
        }
        if (live) {
            ctr = rng.close();
            if (radDirty) q.radiance[pid] = rad;
        }
        const uint32_t slot = blockAppend<kBlock / 32>(nextCount, wantNext, s_scratch);
        if (wantNext) {
            out0[slot] = make_float4(no.x, no.y, no.z, nT.x);
            out1[slot] = make_float4(nd.x, nd.y, nd.z, nT.y);
            out2[slot] = make_float4(nT.z, __int_as_float(int(pid)), __int_as_float(depth + 1), __int_as_float(int(ctr)));
            hitsOut[slot] = make_float4(nh.t, nh.u, nh.v, __int_as_float(nh.prim));
        }
    }
    statAdd(stats, kStatClosest, nClosest);
    statAdd(stats, kStatShadow, nShadow);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(stats + kStatBounceEntries, (unsigned long long)n);
}

// ---------------------------------------------------------------------------------------------------------
// participating media (medium.h, medium.cpp) — used by the volume shade kernel
// ---------------------------------------------------------------------------------------------------------

// DenseGrid lookup: fp32 restatement of OpenVDBGrid::getDensity (grid.h:71-77) — (p-origin)/voxel, floor,
// eight point fetches with `background` outside the block, lerp z then y then x as a + (b-a)*w.
__device__ __forceinline__ float gridVoxel(const DGrid& g, int x, int y, int z)
{
    if (x < 0 || y < 0 || z < 0 || x >= g.nx || y >= g.ny || z >= g.nz) return g.background;
    return __ldg(g.data + (size_t(z) * g.ny + y) * g.nx + x);
}
__device__ __forceinline__ float gridDensity(const DGrid& g, V3 p)
{
    const float fx = (p.x - g.origin[0]) / g.voxel, fy = (p.y - g.origin[1]) / g.voxel, fz = (p.z - g.origin[2]) / g.voxel;
    const float bx = floorf(fx), by = floorf(fy), bz = floorf(fz);
    const float wx = fx - bx, wy = fy - by, wz = fz - bz;
    const int x = int(bx), y = int(by), z = int(bz);
    float v000, v001, v010, v011, v100, v101, v110, v111;
    if (x >= 0 && y >= 0 && z >= 0 && x + 1 < g.nx && y + 1 < g.ny && z + 1 < g.nz) { // interior cell: one address, eight fixed offsets
        const float* __restrict__ c = g.data + (size_t(z) * g.ny + y) * g.nx + x;
        const size_t sy = size_t(g.nx), sz = size_t(g.nx) * g.ny;
        v000 = __ldg(c); v100 = __ldg(c + 1); v010 = __ldg(c + sy); v110 = __ldg(c + sy + 1);
        v001 = __ldg(c + sz); v101 = __ldg(c + sz + 1); v011 = __ldg(c + sz + sy); v111 = __ldg(c + sz + sy + 1);
    }
    else {
        v000 = gridVoxel(g, x, y, z); v001 = gridVoxel(g, x, y, z + 1);
        v010 = gridVoxel(g, x, y + 1, z); v011 = gridVoxel(g, x, y + 1, z + 1);
        v100 = gridVoxel(g, x + 1, y, z); v101 = gridVoxel(g, x + 1, y, z + 1);
        v110 = gridVoxel(g, x + 1, y + 1, z); v111 = gridVoxel(g, x + 1, y + 1, z + 1);
    }
    const float c00 = v000 + (v001 - v000) * wz;
    const float c01 = v010 + (v011 - v010) * wz;
    const float c10 = v100 + (v101 - v100) * wz;
    const float c11 = v110 + (v111 - v110) * wz;
    const float c0 = c00 + (c01 - c00) * wy;
    const float c1 = c10 + (c11 - c10) * wy;
    return c0 + (c1 - c0) * wx;
}

// HenyeyGreenstein::evaluate / sampleDirection (medium.h:29-67); u[1] is drawn first (g++ argument order)
__device__ __forceinline__ float hgEval(float g, V3 wo, V3 wi)
{
    const float cosTheta = dot(wo, wi);
    const float denom = 1 + g * g - 2 * g * cosTheta;
    const float pi4inv = 1.0f / (4.0f * kPI);
    return pi4inv * (1 - g * g) / (denom * sqrtf(denom));
}
__device__ __forceinline__ void hgSample(float g, V3 wo, Rng& rng, V3& wi)
{
    const float u1 = rng.next();
    const float u0 = rng.next();
    float cosTheta;
    if (fabsf(g) < 1e-3) cosTheta = 2 * u0 - 1.0f;
    else {
        const float sqrTerm = (1 - g * g) / (1 - g + 2 * g * u0);
        cosTheta = (1 + g * g - sqrTerm * sqrTerm) / (2 * g);
    }
    const float sinTheta = sqrtf(smax(1.0f - cosTheta * cosTheta, 0.0f));
    const float phi = 2 * kPI * u1;
    const V3 local = mk(cosf(phi) * sinTheta, cosTheta, sinf(phi) * sinTheta);
    V3 t, b;
    orthonormalBasis(wo, t, b);
    wi = localToWorld(local, t, wo, b);
}

// Medium::sampleWavelength (medium.h:102-115) + DiscreteEmpiricalDistribution1D (sampler.h:53-97); the
// std::lower_bound probe order over the 4-entry cdf is unrolled; channel clamped to 2 where the reference reads
// past the cdf.
__device__ __forceinline__ uint32_t sampleWavelength(V3 throughput, V3 albedo, Rng& rng, V3& pmf)
{
    const V3 ta = throughput * albedo;
    float sum = 0;
    sum += ta.x; sum += ta.y; sum += ta.z;
    const float c0 = 0;
    const float c1 = c0 + ta.x / sum;
    const float c2 = c1 + ta.y / sum;
    const float c3 = c2 + ta.z / sum;
    pmf = mk(c1 - c0, c2 - c1, c3 - c2);
    const float u = rng.next();
    int x;
    if (c2 < u) x = (c3 < u) ? 4 : 3;
    else if (c1 < u) x = 2;
    else x = (c0 < u) ? 1 : 0;
    if (x == 0) x++;
    if (x > 3) x = 3;
    return uint32_t(x - 1);
}
__device__ __forceinline__ V3 analyticTr(float t, V3 sigma) { return vexp(-sigma * t); } // medium.h:95-98
__device__ __forceinline__ V3 v3(const float* p) { return mk(p[0], p[1], p[2]); }

// HomogeneousMedium{MIS,Achromatic,NoMIS}::sampleMedium (medium.h:154-191, 202-228, 239-276)
__device__ __forceinline__ bool sampleHomogeneous(const DMedium& m, V3 o, V3 d, V3 rayT, float t0, float t1, Rng& rng, V3& pos, V3& dir, V3& thr)
{
    const V3 sa = v3(m.sigma_a), ss = v3(m.sigma_s), st = v3(m.sigma_t);
    (void)sa;
    const float distToSurface = t1 - t0;
    if (m.kind == XRTG_MEDIUM_HOMOGENEOUS_MIS) {
        V3 pmf = mk(1.0f);
        const uint32_t ch = sampleWavelength(rayT, ss / st, rng, pmf);
        const float t = -logf(smax(1.0f - rng.next(), 0.0f)) / comp(st, ch);
        if (t > distToSurface - kRayEps) {
            pos = o + (t1 + kRayEps) * d; dir = d;
            const V3 tr = analyticTr(distToSurface, st);
            const V3 pdf = pmf * tr;
            thr = tr / (pdf.x + pdf.y + pdf.z);
            return false;
        }
        hgSample(m.g, d, rng, dir);
        pos = o + (t0 + t) * d;
        const V3 tr = analyticTr(t, st);
        const V3 pdf = pmf * (st * tr);
        thr = (tr * ss) / (pdf.x + pdf.y + pdf.z);
        return true;
    }
    if (m.kind == XRTG_MEDIUM_HOMOGENEOUS_ACHROMATIC) {
        const float t = -logf(smax(1.0f - rng.next(), 0.0f)) / st.x;
        if (t > distToSurface - kRayEps) { pos = o + (t1 + kRayEps) * d; dir = d; thr = mk(1.0f); return false; }
        hgSample(m.g, d, rng, dir);
        pos = o + (t0 + t) * d;
        thr = ss / st;
        return true;
    }
    int ch = int(3 * rng.next());
    if (ch == 3) ch--;
    const float pmfw = 1.0f / 3.0f;
    const float sc_ = comp(st, ch);
    const float t = -logf(smax(1.0f - rng.next(), 0.0f)) / sc_;
    const float pdf_distance = sc_ * expf(-sc_ * t);
    if (t > distToSurface - kRayEps) {
        pos = o + (t1 + kRayEps) * d; dir = d;
        const V3 tr = analyticTr(distToSurface, st);
        const float p_surface = expf(-sc_ * distToSurface);
        thr = 1.0f / 3.0f * tr / (pmfw * p_surface);
        return false;
    }
    hgSample(m.g, d, rng, dir);
    pos = o + (t0 + t) * d;
    thr = 1.0f / 3.0f * analyticTr(t, st) * ss / (pmfw * pdf_distance);
    return true;
}

// HeterogeneousMedium::sampleMedium — spectral delta tracking (medium.cpp:45-133), as a resumable loop: trackBegin() = the
// set-up before the loop (:52-60), trackStep() = one iteration (:62-132), returning true when the walk ended (scatter or exit).
struct TrackState {
    float t, t1, density; // density = multiplier * grid density at the current position (sigma_a of the next wavelength pick)
    V3 tt;                // throughput accumulated by the walk
    float sd;             // last sampled distance and wavelength pmf: what trackFinish() needs from the final step
    V3 pmf;
};
struct TrackResult {
    V3 pos, dir, thr;
    bool scattered;
};
enum { kTrackContinue = 0, kTrackExit = 1, kTrackScatter = 2 };
__device__ __forceinline__ void trackBegin(const DMedium& m, const DGrid& g, V3 o, V3 d, float tEntry, float t1, TrackState& ts)
{
    ts.tt = mk(1.f);
    ts.t = tEntry;
    ts.t1 = t1;
    ts.density = m.densityMul * gridDensity(g, o + tEntry * d);
}
// One iteration of the loop up to the decision (medium.cpp:62-72, :84-99): null collisions update the throughput and continue;
// leaving the medium or a real scattering event only RECORD the step (ts.sd, ts.pmf) — the few lanes that end their walk in a
// given step would otherwise run the long exit / scatter epilogues at 2-3 of 32 lanes inside the lockstep loop.
__device__ __forceinline__ int trackStep(const DMedium& m, const DGrid& g, V3 o, V3 d, V3 rayT, TrackState& ts, Rng& rng, uint32_t& steps)
{
    const V3 absC = v3(m.sigma_a), scatC = v3(m.sigma_s);
    const V3 maj = mk(m.majorant);
    V3 sigma_a = absC * ts.density;
    ++steps;
    rng.alignBlock(); // the three draws of one step come from one Philox block
    const uint32_t ch = sampleWavelength(rayT * ts.tt, (maj - sigma_a) * m.invMajorant, rng, ts.pmf);
    ts.sd = -logf(smax(1.0f - rng.next(), 0.0f)) * m.invMajorant;
    ts.t += ts.sd;
    if (ts.t > ts.t1 - kRayEps) return kTrackExit;
    ts.density = m.densityMul * gridDensity(g, o + ts.t * d);
    const V3 sigma_s = scatC * ts.density;
    sigma_a = absC * ts.density;
    const V3 sigma_n = maj - sigma_a - sigma_s;
    const V3 P_s = sigma_s / (sigma_s + sigma_n);
    if (rng.next() < comp(P_s, ch)) return kTrackScatter;
    const V3 P_n = sigma_n / (sigma_s + sigma_n);
    const V3 tr = analyticTr(ts.sd, maj);
    const V3 pdf_distance = m.majorant * tr;
    const V3 pdf = ts.pmf * pdf_distance * P_n;
    ts.tt = ts.tt * ((tr * sigma_n) / (pdf.x + pdf.y + pdf.z));
    return kTrackContinue;
}
// The epilogue of the walk: medium.cpp:73-83 (left the medium) or :100-112 (scattered, new direction from the phase function)
__device__ __forceinline__ void trackFinish(const DMedium& m, V3 o, V3 d, TrackState& ts, int how, Rng& rng, TrackResult& r)
{
    const V3 maj = mk(m.majorant);
    if (how == kTrackExit) {
        r.pos = o + (ts.t1 + kRayEps) * d; r.dir = d;
        const float rest = ts.sd - (ts.t - (ts.t1 - kRayEps));
        const V3 tr = analyticTr(rest, maj);
        const V3 pdf = ts.pmf * tr;
        ts.tt = ts.tt * (tr / (pdf.x + pdf.y + pdf.z));
        r.scattered = false;
    }
    else {
        const V3 absC = v3(m.sigma_a), scatC = v3(m.sigma_s);
        const V3 sigma_s = scatC * ts.density, sigma_a = absC * ts.density;
        const V3 sigma_n = maj - sigma_a - sigma_s;
        const V3 P_s = sigma_s / (sigma_s + sigma_n);
        r.pos = o + ts.t * d;
        hgSample(m.g, d, rng, r.dir);
        const V3 tr = analyticTr(ts.sd, maj);
        const V3 pdf_distance = m.majorant * tr;
        const V3 pdf = ts.pmf * pdf_distance * P_s;
        ts.tt = ts.tt * ((tr * sigma_s) / (pdf.x + pdf.y + pdf.z));
        r.scattered = true;
    }
    r.thr = anyNan(ts.tt) ? mk(0.f) : ts.tt;
}

// Medium::transmittance: analytic (medium.h:134-139) or ratio tracking (medium.h:360-386)
__device__ __forceinline__ V3 transmittance(const DScene& sc, const DMedium& m, V3 p1, V3 p2, Rng& rng, uint32_t& steps)
{
    if (m.kind != XRTG_MEDIUM_HETEROGENEOUS) return analyticTr(length(p1 - p2), v3(m.sigma_t));
    const DGrid g = sc.grids[m.grid];
    const float distToEnd = length(p1 - p2);
    float t = 0;
    const V3 dir = normalize(p2 - p1);
    V3 tr = mk(1.f);
    while (true) {
        const float sd = -logf(smax(1.0f - rng.next(), 0.0f)) * m.invMajorant;
        t += sd;
        if (t > distToEnd) break;
        ++steps;
        const float density = m.densityMul * gridDensity(g, p1 + t * dir);
        const V3 sigma_n = mk(m.majorant) - v3(m.sigma_a) * density - v3(m.sigma_s) * density;
        tr = tr * (sigma_n * m.invMajorant);
    }
    return tr;
}

// VolumePathTracing (integrator.h:409-473) and VolumePathTracingNEE (integrator.h:489-631): ONE iteration of the reference's loop
// for one path whose ray (o, d) has the closest hit h, in three pieces so that the tracking walk in the middle can be driven
// either inline (wavefront kernel) or one step at a time by a whole warp (path kernel):
//   volumePre  : Russian roulette, emitter test, medium lookup. Returns kVolEnd (path over; at most one contribution),
//                kVolTrack (heterogeneous medium: walk initialised in ts) or kVolSampled (homogeneous medium: r is final).
//   trackStep  : see above.
//   volumePost : NEE through the medium (VolumePathTracingNEE), next ray. Returns true if the path continues.
enum { kVolEnd = 0, kVolTrack = 1, kVolSampled = 2 };
__device__ __forceinline__ int volumePre(const DScene& sc, const DMedium* media, const DGrid* grids, bool nee, V3 o, V3 d, V3& T, const Hit& h, int depth,
                                         Rng& rng, int& mi, TrackState& ts, TrackResult& r, bool& hasContrib, V3& contrib)
{
    hasContrib = false;
    if (h.prim < 0) return kVolEnd; // a miss adds throughput*background*(depth!=0) = 0
    Surf s;
    makeSurf(sc, o, d, h, s);
    if (depth > 0) {
        const float p = smin((T.x + T.y + T.z) / 3.0f, 1.0f);
        if (rng.next() >= p) return kVolEnd;
        T = T / mk(p);
    }
    if (lightOf(s) >= 0) {
        if (!nee || depth == 0) { contrib = T * emitted(sc, s, d); hasContrib = true; }
        return kVolEnd;
    }
    mi = mediumOf(s);
    if (mi < 0) {
        // A plain surface: the reference never advances here (infinite loop).
        // The sample is poisoned so that it is DROPPED and counted, like the oracle port does.
        contrib = mk(__int_as_float(0x7fc00000)); hasContrib = true;
        return kVolEnd;
    }
    const DMedium& m = media[mi];
    if (m.kind == XRTG_MEDIUM_HETEROGENEOUS) {
        trackBegin(m, grids[m.grid], o, d, h.t, s.t1, ts);
        return kVolTrack;
    }
    r.scattered = sampleHomogeneous(m, o, d, T, h.t, s.t1, rng, r.pos, r.dir, r.thr);
    return kVolSampled;
}
template <bool COUNT>
__device__ __forceinline__ bool volumePost(const DScene& sc, const DWave& w, const DMedium* media, bool nee, bool brute, V3 d, V3 T, int mi,
                                           const TrackResult& r, int& depth, Rng& rng, int* sstack, TraceCounters& tc, uint32_t& steps,
                                           uint32_t& extraClosest, V3& no, V3& nd, V3& nT, bool& hasContrib, V3& contrib)
{
    hasContrib = false;
    const DMedium& m = media[mi];
    if (nee && r.scattered) {
        // sampleDirectionToLight (integrator.h:583-602), Scene::sampleAreaLight (scene.cpp:182-188)
        unsigned int li = (unsigned int)(float(sc.nLights) * rng.next());
        if (li == (unsigned int)sc.nLights) li--;
        const float choose = 1.0f / float(sc.nLights);
        V3 dl = mk(0.f);
        float dist, lp = 0.0f;
        const V3 Le = sampleLight(sc.lights[li], r.pos, dl, lp, dist, rng);
        const float pdf_dir = choose * lp;
        if (pdf_dir > 0.0f) {
            // isVisible (integrator.h:604-631): ONE closest-hit query, dist_to_light ignored
            V3 trn = mk(1.0f);
            bool visible = true;
            Hit sh;
            closestHit<COUNT>(sc, r.pos, dl, brute, sh, sstack, tc);
            ++extraClosest;
            if (sh.prim >= 0) {
                Surf ss;
                makeSurf(sc, r.pos, dl, sh, ss);
                if (hasMaterial(ss)) visible = false;
                else if (mediumOf(ss) >= 0)
                    trn = trn * transmittance(sc, media[mediumOf(ss)], r.pos + sh.t * dl, r.pos + ss.t1 * dl, rng, steps);
            }
            if (visible) {
                const V3 f = mk(hgEval(m.g, d, dl));
                const V3 Ls = trn * f * Le / pdf_dir;
                contrib = T * r.thr * Ls; hasContrib = true;
            }
        }
    }
    nT = T * r.thr;
    no = r.pos; nd = r.dir;
    if (r.scattered) depth++;
    return depth < w.maxDepth;
}

// Wavefront form: one iteration per launch, continuing paths appended to the next ray queue (deep BVHs: the closest hits of the
// next iteration go through the refillable traversal kernel).
template <bool COUNT>
__global__ void __launch_bounds__(kBlock) k_shade_volume(DScene sc, DQueues q, DWave w, int src, int bounce, int brute, unsigned long long* stats)
{
    __shared__ int s_stack[kStackSmem * kBlock];
    uint32_t* ctrl = q.ctrl + bounce * kCtrlStride;
    const uint32_t n = ctrl[kCtrlRays];
    ShadeOut so{q, ctrl + kCtrlStride, ctrl, src ^ 1};
    const bool nee = (w.integrator == XRTG_INT_VOLUME_NEE);
    uint32_t steps = 0, extraClosest = 0;
    TraceCounters tc;
    // warp-granular dynamic fetch: tracking cost varies by orders of magnitude between paths, so neither a static
    // partition nor CTA-sized tiles keep the SMs busy (measured: 5.1 ms vs 3.5 ms per 8 spp on workload c5)
    uint32_t resNext = 0, resEnd = 0, base;
    while (warpNextBatch<32>(ctrl + kCtrlFetchShade, n, resNext, resEnd, base)) { // finest grain: path cost varies wildly
        const uint32_t i = base + laneId();
        bool wantRay = false;
        V3 no = mk(0.f), nd = mk(0.f), nT = mk(0.f);
        uint32_t pid = 0, ctr = 0;
        int depth = 0;
        if (i < n) {
            const float4 r0 = q.q0[src][i], r1 = q.q1[src][i], r2 = q.q2[src][i], hv = q.hits[i];
            pid = uint32_t(__float_as_int(r2.y));
            depth = __float_as_int(r2.z);
            const Hit h{hv.x, hv.y, hv.z, __float_as_int(hv.w)};
            const V3 o = xyz(r0), d = xyz(r1);
            V3 T = mk(r0.w, r1.w, r2.x);
            Rng rng;
            rng.open(w, pid, uint32_t(__float_as_int(r2.w)));
            bool hasContrib;
            V3 contrib;
            int mi = -1;
            TrackState ts;
            TrackResult r;
            const int what = volumePre(sc, sc.media, sc.grids, nee, o, d, T, h, depth, rng, mi, ts, r, hasContrib, contrib);
            if (hasContrib) addRadiance(q, pid, contrib);
            if (what != kVolEnd) {
                if (what == kVolTrack) {
                    const DMedium m = sc.media[mi];
                    const DGrid g = sc.grids[m.grid];
                    int how;
                    while ((how = trackStep(m, g, o, d, T, ts, rng, steps)) == kTrackContinue) {}
                    trackFinish(m, o, d, ts, how, rng, r);
                }
                wantRay = volumePost<COUNT>(sc, w, sc.media, nee, brute != 0, d, T, mi, r, depth, rng, s_stack + threadIdx.x, tc, steps, extraClosest, no, nd,
                                            nT, hasContrib, contrib);
                if (hasContrib) addRadiance(q, pid, contrib);
            }
            ctr = rng.close();
        }
        pushRayWarp(so, wantRay, no, nd, nT, pid, depth, ctr);
    }
    if (extraClosest) atomicAdd(stats + kStatClosest, (unsigned long long)extraClosest);
    statAdd(stats, kStatSteps, steps);
    if (COUNT) { statAdd(stats, kStatNodes, tc.nodes); statAdd(stats, kStatTris, tc.tris); }
}

// Shallow BVHs (every volume scene of the reference: a box or a sphere, a light, a few walls): after the primary hit a path
// only ever produces ONE next ray, and only ~14 % of the 1080p paths enter the medium at all, so the wavefront form degenerates
// into dozens of launches over a few ten thousand rays each plus a host poll per iteration (workload c5: 105 launches per wave,
// 10.5 M closest hits of which 8.3 M are primary). Here every lane owns one path at a time and runs it to completion — the same
// draws in the same order — and the warp is a small state machine around the one hot loop, the delta-tracking walk:
//   kLanePre   : the lane has a ray + hit: volumePre(), then kLaneTrack, kLanePost or finished
//   kLaneTrack : trackStep() executed in lockstep by every tracking lane, stepsPerVote steps per vote, while at least
//                `threshold` lanes are still walking (walk lengths differ by orders of magnitude: run per thread the loop
//                keeps 7.6 of 32 lanes busy, ncu profiles/r01_notes.md)
//   kLaneWalked: trackFinish() — the exit / scatter epilogue, outside the lockstep loop
//   kLanePost  : volumePost(), inline closest hit of the next ray, back to kLanePre or finished
//   kLaneIdle  : refilled from the compact bounce-0 queue (one atomic per 32 entries per warp)
// One launch per wave, no host round trip.
enum { kLaneIdle = 0, kLanePre, kLaneTrack, kLaneWalked, kLanePost };
constexpr int kMediaSmem = 8; // media / grids staged in shared memory by k_volume_paths (the host falls back to the wavefront form above that)
template <bool COUNT, int MINB>
__global__ void __launch_bounds__(kBlock, MINB) k_volume_paths(DScene sc, DQueues q, DWave w, int brute, int maxIter, int threshold, int stepsPerVote, unsigned long long* stats)
{
    __shared__ int s_stack[kStackSmem * kBlock];
    __shared__ DMedium s_media[kMediaSmem];
    __shared__ DGrid s_grids[kMediaSmem];
    if (int(threadIdx.x) < min(sc.nMedia, kMediaSmem)) s_media[threadIdx.x] = sc.media[threadIdx.x];
    if (int(threadIdx.x) < min(sc.nGrids, kMediaSmem)) s_grids[threadIdx.x] = sc.grids[threadIdx.x];
    __syncthreads();
    const DMedium* media = s_media;
    const DGrid* grids = s_grids;
    uint32_t* ctrl = q.ctrl;
    const uint32_t n = ctrl[kCtrlRays];
    const bool nee = (w.integrator == XRTG_INT_VOLUME_NEE);
    const uint32_t lane = laneId();
    uint32_t steps = 0, nClosest = 0;
    TraceCounters tc;
    // per-lane path state
    int state = kLaneIdle, depth = 0, it = 0, mi = -1, walkEnd = kTrackContinue;
    uint32_t pid = 0;
    V3 o = mk(0.f), d = mk(0.f), T = mk(0.f);
    Hit h{FLT_MAX, 0.f, 0.f, -1};
    Rng rng;
    TrackState ts;
    TrackResult r;
    float4 rad = make_float4(0.f, 0.f, 0.f, 0.f);
    bool dirty = false, exhausted = false;
    uint32_t resNext = 0, resEnd = 0;
    auto finish = [&]() { // the lane's path is over
        rng.close();
        if (dirty) q.radiance[pid] = rad;
        state = kLaneIdle;
    };
    auto add = [&](V3 c) { rad.x += c.x; rad.y += c.y; rad.z += c.z; dirty = true; };
    while (true) {
        // ---- refill idle lanes from the warp's reservation of 32 consecutive queue entries ----
        const uint32_t need = __ballot_sync(0xffffffffu, state == kLaneIdle);
        if (need != 0 && !exhausted) {
            const uint32_t nNeed = __popc(need), rank = __popc(need & ((1u << lane) - 1u)), left = resEnd - resNext;
            uint32_t nb = 0;
            if (nNeed > left) {
                if (lane == 0) nb = atomicAdd(ctrl + kCtrlFetchShade, 32u);
                nb = __shfl_sync(0xffffffffu, nb, 0);
                if (nb >= n) exhausted = true;
            }
            const uint32_t i = rank < left ? resNext + rank : nb + (rank - left);
            if (nNeed > left) { resNext = nb + (nNeed - left); resEnd = nb + 32u; }
            else resNext += nNeed;
            if (state == kLaneIdle && i < n) {
                const float4 r0 = q.q0[0][i], r1 = q.q1[0][i], r2 = q.q2[0][i], hv = q.hits[i];
                pid = uint32_t(__float_as_int(r2.y));
                depth = __float_as_int(r2.z);
                o = xyz(r0); d = xyz(r1); T = mk(r0.w, r1.w, r2.x);
                h = Hit{hv.x, hv.y, hv.z, __float_as_int(hv.w)};
                rng.open(w, pid, uint32_t(__float_as_int(r2.w)));
                rad = q.radiance[pid];
                dirty = false;
                it = 0;
                state = kLanePre;
            }
        }
        if (__ballot_sync(0xffffffffu, state != kLaneIdle) == 0) break;
        // ---- lanes between walks: epilogue of the last walk, closest hit of the next ray, prologue of the next walk ----
        if (state == kLaneWalked) {
            trackFinish(media[mi], o, d, ts, walkEnd, rng, r);
            state = kLanePost;
        }
        if (state == kLanePost) {
            V3 no, nd, nT, contrib;
            bool hasContrib;
            const bool cont = volumePost<COUNT>(sc, w, media, nee, brute != 0, d, T, mi, r, depth, rng, s_stack + threadIdx.x, tc, steps, nClosest, no, nd,
                                                nT, hasContrib, contrib);
            if (hasContrib) add(contrib);
            if (!cont || ++it == maxIter) finish();
            else {
                o = no; d = nd; T = nT;
                closestHit<COUNT>(sc, o, d, brute != 0, h, s_stack + threadIdx.x, tc);
                ++nClosest;
                state = kLanePre;
            }
        }
        if (state == kLanePre) {
            V3 contrib;
            bool hasContrib;
            const int what = volumePre(sc, media, grids, nee, o, d, T, h, depth, rng, mi, ts, r, hasContrib, contrib);
            if (hasContrib) add(contrib);
            if (what == kVolEnd) finish();
            else state = what == kVolTrack ? kLaneTrack : kLanePost;
        }
        // ---- the walk: every tracking lane takes kTrackSteps steps per vote ----
        const uint32_t pending = __ballot_sync(0xffffffffu, state == kLanePre || state == kLanePost || state == kLaneWalked || (state == kLaneIdle && !exhausted));
        const uint32_t thr = pending ? uint32_t(threshold) : 1u;
        uint32_t busy = __popc(__ballot_sync(0xffffffffu, state == kLaneTrack));
        while (busy >= thr && busy > 0) {
#pragma unroll 1
            for (int k = 0; k < stepsPerVote; ++k)
                if (state == kLaneTrack) {
                    const DMedium& m = media[mi];
                    walkEnd = trackStep(m, grids[m.grid], o, d, T, ts, rng, steps);
                    if (walkEnd != kTrackContinue) state = kLaneWalked;
                }
            busy = __popc(__ballot_sync(0xffffffffu, state == kLaneTrack));
        }
    }
    statAdd(stats, kStatClosest, nClosest);
    statAdd(stats, kStatSteps, steps);
    if (COUNT) { statAdd(stats, kStatNodes, tc.nodes); statAdd(stats, kStatTris, tc.tris); }
}

// ---------------------------------------------------------------------------------------------------------
// accumulate: validate each sample like renderer.cpp:57-73 (NaN / inf / any negative channel -> dropped, the
// divisor is unchanged) and add the wave's samples to the pixel sum IN SAMPLE ORDER (bit-exact vs
// Image::addPixel order). One thread per pixel, no atomics.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_accumulate(DQueues q, DWave w, float* __restrict__ accum, unsigned long long* stats)
{
    uint32_t dropped = 0;
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < w.nPixels; p += gridDim.x * blockDim.x) {
        float ax = accum[3 * size_t(p)], ay = accum[3 * size_t(p) + 1], az = accum[3 * size_t(p) + 2];
        for (uint32_t s = 0; s < w.samplesThisWave; ++s) {
            const float4 r = q.radiance[size_t(s) * w.nPixels + p];
            if (isnan(r.x) || isnan(r.y) || isnan(r.z)) { ++dropped; continue; }
            else if (isinf(r.x) || isinf(r.y) || isinf(r.z)) { ++dropped; continue; }
            else if (r.x < 0 || r.y < 0 || r.z < 0) { ++dropped; continue; }
            ax += r.x; ay += r.y; az += r.z;
        }
        accum[3 * size_t(p)] = ax; accum[3 * size_t(p) + 1] = ay; accum[3 * size_t(p) + 2] = az;
    }
    if (dropped) atomicAdd(stats + kStatDropped, (unsigned long long)dropped);
}

// image /= Vec3f(n_samples) (renderer.cpp:98) — IEEE division like the reference; divisor 0 = leave the sum
__global__ void __launch_bounds__(kBlock) k_finalize(const float* __restrict__ accum, float* __restrict__ out, size_t n, float divisor)
{
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
        out[i] = divisor > 0.f ? accum[i] / divisor : accum[i];
}

// ---------------------------------------------------------------------------------------------------------
// host-side launchers (called from api.cpp through the table in kernels.h)
// ---------------------------------------------------------------------------------------------------------
// persistent grids: (resident CTAs per SM for this kernel) x (number of SMs) — 148 on B200
inline int gridFor(const void* fn, int block = kBlock)
{
    static thread_local int cachedDev = -1;
    static thread_local int sms = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != cachedDev) {
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cachedDev = dev;
    }
    int perSm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, fn, block, 0);
    return sms * (perSm < 1 ? 1 : perSm);
}

inline void launchSeedMt(cudaStream_t st, const DWave& w)
{
    const int grid = int((w.nPixels + kBlock - 1) / kBlock);
    k_seed_mt<<<grid, kBlock, 0, st>>>(w.mt, w.mti, w.nPixels);
}
inline void launchGenJitter(cudaStream_t st, const DWave& w, int spp, float* jitter)
{
    const int grid = int((w.nPixels + kBlock - 1) / kBlock);
    k_gen_jitter<<<grid, kBlock, 0, st>>>(w, spp, jitter);
}
inline void launchRaygen(cudaStream_t st, const DCamera& cam, const DQueues& q, const DWave& w, const float* jitter)
{
    static thread_local int grid = 0;
    if (!grid) grid = gridFor((const void*)k_raygen);
    k_raygen<<<grid, kBlock, 0, st>>>(cam, q, w, jitter);
}
inline void launchPrimary(cudaStream_t st, const DScene& sc, const DCamera& cam, const DQueues& q, const DWave& w, bool brute, int missMode,
                          bool count, unsigned long long* stats)
{
    static thread_local int g0 = 0, g1 = 0;
    if (!g0) { g0 = gridFor((const void*)k_primary<false>); g1 = gridFor((const void*)k_primary<true>); }
    if (count) k_primary<true><<<g1, kBlock, 0, st>>>(sc, cam, q, w, brute, missMode, stats);
    else k_primary<false><<<g0, kBlock, 0, st>>>(sc, cam, q, w, brute, missMode, stats);
}
inline void launchExtend(cudaStream_t st, const DScene& sc, const DQueues& q, int src, int bounce, int brute, bool count, unsigned long long* stats,
                         int thr, int spv, int leafThr)
{
    static thread_local int g0 = 0, g1 = 0, h0 = 0, h1 = 0;
    if (!g0) {
        g0 = gridFor((const void*)k_trace<false, false>); g1 = gridFor((const void*)k_trace<false, true>);
        h0 = gridFor((const void*)k_extend_simple<false>); h1 = gridFor((const void*)k_extend_simple<true>);
    }
    if (thr <= 0) { // shallow BVH: simple run-to-completion kernel
        if (count) k_extend_simple<true><<<h1, kBlock, 0, st>>>(sc, q, src, bounce, brute, stats);
        else k_extend_simple<false><<<h0, kBlock, 0, st>>>(sc, q, src, bounce, brute, stats);
        return;
    }
    if (count) k_trace<false, true><<<g1, kBlock, 0, st>>>(sc, q, src, bounce, brute, stats, nullptr, thr, spv, leafThr);
    else k_trace<false, false><<<g0, kBlock, 0, st>>>(sc, q, src, bounce, brute, stats, nullptr, thr, spv, leafThr);
}
inline void launchConnect(cudaStream_t st, const DScene& sc, const DQueues& q, int bounce, int brute, bool count, unsigned long long* stats,
                          int thr, int spv, int leafThr)
{
    static thread_local int g0 = 0, g1 = 0, h0 = 0, h1 = 0;
    if (!g0) {
        g0 = gridFor((const void*)k_trace<true, false>); g1 = gridFor((const void*)k_trace<true, true>);
        h0 = gridFor((const void*)k_connect_simple<false>); h1 = gridFor((const void*)k_connect_simple<true>);
    }
    if (thr <= 0) {
        if (count) k_connect_simple<true><<<h1, kBlock, 0, st>>>(sc, q, bounce, brute, stats);
        else k_connect_simple<false><<<h0, kBlock, 0, st>>>(sc, q, bounce, brute, stats);
        return;
    }
    if (count) k_trace<true, true><<<g1, kBlock, 0, st>>>(sc, q, 0, bounce, brute, stats, nullptr, thr, spv, leafThr);
    else k_trace<true, false><<<g0, kBlock, 0, st>>>(sc, q, 0, bounce, brute, stats, nullptr, thr, spv, leafThr);
}
inline void launchShadeSurface(cudaStream_t st, const DScene& sc, const DQueues& q, const DWave& w, int src, int bounce)
{
    static thread_local int grid = 0;
    if (!grid) grid = gridFor((const void*)k_shade_surface, kShadeBlock);
    k_shade_surface<<<grid, kShadeBlock, 0, st>>>(sc, q, w, src, bounce);
}
inline void launchBounceSmall(cudaStream_t st, const DScene& sc, const DQueues& q, const DWave& w, int src, int bounce, unsigned long long* stats)
{
    static thread_local int g0 = 0, g1 = 0;
    if (!g0) { g0 = gridFor((const void*)k_bounce_small<false>); g1 = gridFor((const void*)k_bounce_small<true>); }
    const bool grouped = !kExact && sc.smallBlockF4 > 0 && sc.nBoxes == 0;
    if (grouped) k_bounce_small<true><<<g1, kBlock, 0, st>>>(sc, q, w, src, bounce, stats);
    else k_bounce_small<false><<<g0, kBlock, 0, st>>>(sc, q, w, src, bounce, stats);
}
inline void launchShadeVolume(cudaStream_t st, const DScene& sc, const DQueues& q, const DWave& w, int src, int bounce, bool brute, bool count,
                              unsigned long long* stats)
{
    static thread_local int g0 = 0, g1 = 0;
    if (!g0) { g0 = gridFor((const void*)k_shade_volume<false>); g1 = gridFor((const void*)k_shade_volume<true>); }
    if (count) k_shade_volume<true><<<g1, kBlock, 0, st>>>(sc, q, w, src, bounce, brute, stats);
    else k_shade_volume<false><<<g0, kBlock, 0, st>>>(sc, q, w, src, bounce, brute, stats);
}
inline void launchVolumePaths(cudaStream_t st, const DScene& sc, const DQueues& q, const DWave& w, bool brute, int maxIter, int threshold, int stepsPerVote, bool count,
                              unsigned long long* stats)
{
    static thread_local int g0 = 0, g1 = 0, g2 = 0, sel = 0;
    if (!g0) {
        g0 = gridFor((const void*)k_volume_paths<false, 5>); g1 = gridFor((const void*)k_volume_paths<true, 4>);
        g2 = gridFor((const void*)k_volume_paths<false, 4>);
        const char* e = std::getenv("XRT_VOLUME_MINB");
        sel = e ? std::atoi(e) : 4;
    }
    if (count) k_volume_paths<true, 4><<<g1, kBlock, 0, st>>>(sc, q, w, brute, maxIter, threshold, stepsPerVote, stats);
    else if (sel == 4) k_volume_paths<false, 4><<<g2, kBlock, 0, st>>>(sc, q, w, brute, maxIter, threshold, stepsPerVote, stats);
    else k_volume_paths<false, 5><<<g0, kBlock, 0, st>>>(sc, q, w, brute, maxIter, threshold, stepsPerVote, stats);
}
inline void launchAccumulate(cudaStream_t st, const DQueues& q, const DWave& w, float* accum, unsigned long long* stats)
{
    const int grid = int((w.nPixels + kBlock - 1) / kBlock);
    k_accumulate<<<grid, kBlock, 0, st>>>(q, w, accum, stats);
}
inline void launchFinalize(cudaStream_t st, const float* accum, float* out, size_t n, float divisor)
{
    const int grid = int(std::min<size_t>((n + kBlock - 1) / kBlock, 148 * 16));
    k_finalize<<<grid, kBlock, 0, st>>>(accum, out, n, divisor);
}
// parity hook: rays go through the SAME persistent traversal kernel the renderer uses. `out` = n float4 (device):
// closest -> the hit queue itself is returned by the caller; any hit -> occlusion flags are written to `out`.
inline void launchTraceRays(cudaStream_t st, const DScene& sc, const DQueues& q, const float* org, const float* dir, const float* tmax,
                            long long n, bool anyhit, bool brute, float4* out, unsigned long long* stats)
{
    const int grid = int(std::min<long long>((n + kBlock - 1) / kBlock, 148 * 16));
    k_pack_rays<<<grid > 0 ? grid : 1, kBlock, 0, st>>>(q, org, dir, tmax, uint32_t(n), anyhit ? 1 : 0);
    if (anyhit) k_trace<true, false><<<gridFor((const void*)k_trace<true, false>), kBlock, 0, st>>>(sc, q, 0, 0, brute, stats, out, 16, 1, 8);
    else k_trace<false, false><<<gridFor((const void*)k_trace<false, false>), kBlock, 0, st>>>(sc, q, 0, 0, brute, stats, nullptr, 16, 1, 8);
}

} // namespace XRT_NS
} // namespace xrt
