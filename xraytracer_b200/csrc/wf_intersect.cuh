// wf_intersect.cuh — part of wavefront.cuh (included inside namespace xrt::XRT_NS, in this order): triangle / sphere / box tests, BVH2 traversal, the small-scene triangle loops, Scene::intersect and Scene::occluded for one ray.
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
constexpr int kSmallPrims = 96;    // shading records / area lights k_bounce_small stages in shared memory (the host uses that
constexpr int kSmallLights = 16;   // kernel only below these limits)
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
