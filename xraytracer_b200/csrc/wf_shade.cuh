// wf_shade.cuh — part of wavefront.cuh (included inside namespace xrt::XRT_NS, in this order): surface shading: lights, Lambert, k_shade_surface and the fused per-bounce kernel of small scenes (k_bounce_small).
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
// re-normalised. A sphere hit keeps the reference's stale dpdu/dpdv in the exact instantiation (sphereStaleBasis), zero in the
// throughput one.
// Sphere::intersect never writes dpdu / dpdv (primitive.h:112-122): after Scene::intersect (scene.cpp:190-200) they still hold what
// the last MESH hit that was the closest so far wrote while the objects before the sphere were visited — or zero if there was
// none. The hits that "were the closest so far" form a decreasing sequence, so the last one is simply the closest hit among
// the primitives with a smaller id; if that is another sphere, the question repeats for it. Exact instantiation only (the
// throughput one keeps zeros: both are an artefact); BoxMesh objects (which reset info.t) are not modelled.
// O(#triangles) per Lambert-sphere hit: the parity path, not the fast one.
__device__ __noinline__ void sphereStaleBasis(const DScene& sc, V3 o, V3 d, int sphereId, V3& dpdu, V3& dpdv)
{
    dpdu = mk(0.f); dpdv = mk(0.f);
    int cur = sphereId;
    for (int guard = 0; guard <= sc.nSpheres; ++guard) {
        Hit best{FLT_MAX, 0.f, 0.f, 0x7fffffff};
        const float4* __restrict__ tris = sc.tris_id;
        for (int i = 0; i < sc.nBruteTris; ++i) {
            const float4 q0 = __ldg(tris + 3 * i);
            const int id = __float_as_int(q0.w);
            if (id >= cur) break; // primitive-id order
            const float4 q1 = __ldg(tris + 3 * i + 1), q2 = __ldg(tris + 3 * i + 2);
            float t, u, v;
            if (rayTriangle(o, d, xyz(q0), xyz(q1), xyz(q2), t, u, v)) consider(best, t, u, v, id);
        }
        bool sphereWins = false;
        for (int k = 0; k < sc.nSpheres; ++k) {
            const float4 cr = __ldg(sc.spheres + 2 * k);
            const int id = __ldg(reinterpret_cast<const int4*>(sc.spheres + 2 * k + 1)).x;
            float t;
            if (id < cur && sphereT(cr, o, d, t) && (t < best.t || (t == best.t && id < best.prim))) { best.t = t; best.prim = id; sphereWins = true; }
        }
        if (best.prim == 0x7fffffff) return;          // nothing in front of the sphere in visiting order: zero-initialised
        if (sphereWins && best.prim < cur) {           // an earlier sphere was the closest so far: look before it
            bool isSphere = false;
            for (int k = 0; k < sc.nSpheres; ++k) isSphere = isSphere || __ldg(reinterpret_cast<const int4*>(sc.spheres + 2 * k + 1)).x == best.prim;
            if (isSphere) { cur = best.prim; continue; }
        }
        const float4 p0 = __ldg(sc.prims + 4 * best.prim), p1 = __ldg(sc.prims + 4 * best.prim + 1), p2 = __ldg(sc.prims + 4 * best.prim + 2);
        const V3 ns = xyz(p0) * (1.0f - best.u - best.v) + xyz(p1) * best.u + xyz(p2) * best.v;
        orthonormalBasis(ns, dpdu, dpdv);
        return;
    }
}

template <bool SMEM>
__device__ __forceinline__ float4 primWord(const float4* prims, int i) { return SMEM ? prims[i] : __ldg(prims + i); }
// SMEM: `prims` is the CTA's shared-memory copy of the shading records (small scenes), otherwise sc.prims in global memory
template <bool SMEM>
__device__ __forceinline__ void makeSurfT(const DScene& sc, const float4* prims, V3 o, V3 d, const Hit& h, Surf& s)
{
    const float4 p3 = primWord<SMEM>(prims, 4 * h.prim + 3);
    s.meta = __float_as_uint(p3.w);
    s.albedo = xyz(p3);
    s.pos = o + h.t * d;
    s.t1 = h.u;
    const uint32_t kind = s.meta & kMetaKindMask;
    if (kind == XRTG_OBJ_MESH) {
        const float4 p0 = primWord<SMEM>(prims, 4 * h.prim), p1 = primWord<SMEM>(prims, 4 * h.prim + 1), p2 = primWord<SMEM>(prims, 4 * h.prim + 2);
        s.ng = mk(p0.w, p1.w, p2.w);
        s.ns = xyz(p0) * (1.0f - h.u - h.v) + xyz(p1) * h.u + xyz(p2) * h.v;
        orthonormalBasis(s.ns, s.dpdu, s.dpdv);
    }
    else if (kind == XRTG_OBJ_SPHERE) {
        const float4 p0 = primWord<SMEM>(prims, 4 * h.prim);
        s.ng = normalize(s.pos - xyz(p0));
        s.ns = s.ng;
        s.dpdu = mk(0.f); s.dpdv = mk(0.f);
        if constexpr (kExact) {
            if ((s.meta & kMetaHasMaterial) != 0) sphereStaleBasis(sc, o, d, h.prim, s.dpdu, s.dpdv);
        }
    }
    else {
        s.ng = mk(0.f); s.ns = mk(0.f); s.dpdu = mk(0.f); s.dpdv = mk(0.f);
    }
}
__device__ __forceinline__ void makeSurf(const DScene& sc, V3 o, V3 d, const Hit& h, Surf& s) { makeSurfT<false>(sc, sc.prims, o, d, h, s); }
__device__ __forceinline__ bool hasMaterial(const Surf& s) { return (s.meta & kMetaHasMaterial) != 0; }
__device__ __forceinline__ int lightOf(const Surf& s) { return int((s.meta >> kMetaLightShift) & 0xfffu) - 1; }
__device__ __forceinline__ int mediumOf(const Surf& s) { return int((s.meta >> kMetaMediumShift) & 0xfffu) - 1; }

// AreaLight::Le (light.h:62-69)
__device__ __forceinline__ V3 emittedFrom(const DLight* lights, const Surf& s, V3 rayDir)
{
    const int li = lightOf(s);
    if (li < 0) return mk(0.f);
    return (dot(rayDir, s.ns) < 0) ? xyz(lights[li].Le) : mk(0.f);
}
__device__ __forceinline__ V3 emitted(const DScene& sc, const Surf& s, V3 rayDir) { return emittedFrom(sc.lights, s, rayDir); }

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

__device__ __forceinline__ void pushShadow(const ShadeOut& so, bool want, V3 o, V3 d, float tmax, uint32_t pid, V3 c, uint32_t (*scratch2)[kShadeWarps + 1], int& phase)
{
    const uint32_t slot = blockAppendAlt(so.ctrlCur + kCtrlShadow, want, scratch2, phase);
    if (want) {
        qstore(so.q.s0 + slot, make_float4(o.x, o.y, o.z, tmax));
        qstore(so.q.s1 + slot, make_float4(d.x, d.y, d.z, __int_as_float(int(pid))));
        qstore(so.q.s2 + slot, make_float4(c.x, c.y, c.z, 0.f));
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
__device__ __forceinline__ void pushRay(const ShadeOut& so, bool want, V3 o, V3 d, V3 T, uint32_t pid, int depth, uint32_t ctr, uint32_t (*scratch2)[kShadeWarps + 1], int& phase)
{
    const uint32_t slot = blockAppendAlt(so.ctrlNext + kCtrlRays, want, scratch2, phase);
    if (want) {
        qstore(so.q.q0[so.dst] + slot, make_float4(o.x, o.y, o.z, T.x));
        qstore(so.q.q1[so.dst] + slot, make_float4(d.x, d.y, d.z, T.y));
        qstore(so.q.q2[so.dst] + slot, make_float4(T.z, __int_as_float(int(pid)), __int_as_float(depth), __int_as_float(int(ctr))));
    }
}

// Surface integrators: Normal (integrator.h:29-36), furnace (:59-66), Direct (:82-119), Indirect (:129-186),
// GI (:205-287), Whitted's Lambert/delta-light branch (:302-394). One thread per ray-queue entry of bounce b.
__global__ void __launch_bounds__(kShadeBlock, XRT_SHADE_MINB) k_shade_surface(DScene sc, DQueues q, DWave w, int src, int bounce)
{
    __shared__ uint32_t s_scratch[2][kShadeWarps + 1]; // two buffers used alternately: two barriers per append (blockAppendAlt)
    int appendPhase = 0;
    uint32_t nUntraced = 0, nUntracedClosest = 0;
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
        bool liveEntry = live;
        if (live) {
            hv = qload(q.hits + i); r2 = qload(q.q2[src] + i);
            if (XRT_WARP_APPEND_PRIMARY && deadEntry(r2)) liveEntry = false; // (unused tail slot of k_primary's warp-private chunks)
            else if (__float_as_int(hv.w) >= 0) { r0 = qload(q.q0[src] + i); r1 = qload(q.q1[src] + i); }
        }
        const V3 o = xyz(r0), d = xyz(r1);
        V3 T = mk(r0.w, r1.w, r2.x);
        pid = uint32_t(__float_as_int(r2.y));
        depth = __float_as_int(r2.z);
        Hit h{hv.x, hv.y, hv.z, __float_as_int(hv.w)};
        Rng rng;
        bool shadeLights = false, shadeDelta = false;
        Surf s = {};
        if (liveEntry) {
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
                    if (kExact && depth > 0) { // russian roulette (integrator.h:223-231); throughput instantiation: already applied, see below
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
                        // Throughput instantiation: a contribution that is exactly zero (surface facing away from the light: cs = 0,
                        // or the light's back side: Lr = 0) cannot change the image whatever Scene::occluded (scene.cpp:202-211)
                        // returns, so the shadow ray is counted as the reference's call but not traced (NaN contributions compare
                        // unequal to zero and keep their ray). On the 1 M-triangle scene that is a quarter of the shadow rays.
                        if (!kExact && c.x == 0.f && c.y == 0.f && c.z == 0.f) { want = false; ++nUntraced; }
                    }
                }
                const float bias = 0.01f;
                pushShadow(so, want, s.pos + s.ng * bias, wi, tmax - bias, pid, c, s_scratch, appendPhase);
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
                    if (!kExact && c.x == 0.f && c.y == 0.f && c.z == 0.f) { want = false; ++nUntraced; } // (as above)
                }
                pushShadow(so, want, s.pos + s.ng * float(0.1), wi, tmax, pid, c, s_scratch, appendPhase);
            }
        }
        // ---- BSDF bounce (integrator.h:271-283) ----
        if (!kExact && depth + 1 >= w.maxDepth) wantRay = false; // (last depth: nothing follows the BSDF sample on the path's counter stream)
        if (wantRay) {
            float pdf = 1.0f;
            V3 nextDir = mk(0.f);
            const V3 fr = sampleBxDF(s, rng, nextDir, pdf);
            const float cs = smax(.0f, dot(nextDir, s.ng));
            nT = T * (fr * cs / pdf);
            no = s.pos + s.ng * 0.01f;
            nd = nextDir;
            wantRay = (depth + 1 < w.maxDepth);
            // Throughput instantiation: the Russian roulette of depth + 1 (integrator.h:223-231) is decided HERE, before the ray is
            // traced. The reference intersects first and rolls afterwards, but a path that loses the roll contributes nothing more
            // whatever it hit (the background of these integrators is black), and the roll is the next draw of this path's own
            // counter stream either way — same draws, same decisions, same image; the rays of the losers (a third of the
            // bounce >= 1 rays on the 1 M-triangle scene) are counted as the reference's Scene::intersect calls but never traced.
            if (!kExact && wantRay) {
                const float p = smin((nT.x + nT.y + nT.z) / 3.0f, 1.0f);
                if (rng.next() >= p) { wantRay = false; ++nUntracedClosest; }
                else nT = nT / mk(p);
            }
        }
        if (liveEntry) ctr = rng.close();
        if (kind == XRTG_INT_INDIRECT || kind == XRTG_INT_GI) pushRay(so, wantRay, no, nd, nT, pid, depth + 1, ctr, s_scratch, appendPhase);
    }    if (!kExact) { // reference-equivalent rays that were not traced: counted as calls, subtracted from rays_traced
        statAdd(q.stats, kStatShadow, nUntraced);
        statAdd(q.stats, kStatClosest, nUntracedClosest);
        statAdd(q.stats, kStatScissored, nUntraced + nUntracedClosest);
        statAdd(q.stats, kStatUntracedShadow, nUntraced);
        statAdd(q.stats, kStatUntracedClosest, nUntracedClosest);
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
// one ray/plane intersection and two barycentric plane equations per triangle. Records come in component-interleaved PAIRS and
// all the fp32 arithmetic of a pair runs as packed fp32x2 instructions (FFMA2 / FMUL2 / FADD2, sm_100): 26 packed
// instructions do what 52 scalar ones did, and the loops are bound by issue slots, not by the FMA pipe. Branch-free per lane.
struct SmallSection {
    const float4* recs; // 10 per pair of records (float 2c + j = component c of record j)
    const int4* ids;    // per pair: (A, B) of record 0, (A, B) of record 1
    int nPairs;
};
__device__ __forceinline__ SmallSection smallSection(const float4* blk, int off)
{
    const int4 h = *reinterpret_cast<const int4*>(blk + off);
    return SmallSection{blk + h.z, reinterpret_cast<const int4*>(blk + h.w), h.y};
}
__device__ __forceinline__ float2 lo2(const float4 v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(const float4 v) { return make_float2(v.z, v.w); }
// the ray, broadcast into both halves once per ray
struct Ray2 {
    float2 nox, noy, noz, dx, dy, dz, ox, oy, oz;
};
__device__ __forceinline__ Ray2 makeRay2(V3 o, V3 d)
{
    return Ray2{make_float2(-o.x, -o.x), make_float2(-o.y, -o.y), make_float2(-o.z, -o.z), make_float2(d.x, d.x), make_float2(d.y, d.y),
                make_float2(d.z, d.z), make_float2(o.x, o.x), make_float2(o.y, o.y), make_float2(o.z, o.z)};
}
// det = N.d and t = (d - N.o) / det of the two planes of a pair
__device__ __forceinline__ void planes2(const float4 q0, const float4 q1, const Ray2& r, float2& det, float2& t)
{
    const float2 Nx = lo2(q0), Ny = hi2(q0), Nz = lo2(q1), dd = hi2(q1);
    det = __ffma2_rn(Nz, r.dz, __ffma2_rn(Ny, r.dy, __fmul2_rn(Nx, r.dx)));
    const float2 num = __ffma2_rn(Nx, r.nox, __ffma2_rn(Ny, r.noy, __ffma2_rn(Nz, r.noz, dd)));
    t = __fmul2_rn(num, make_float2(1.0f / det.x, 1.0f / det.y));
}
// min(u, v, 1 - u - v) of one triangle slot (A: k = 2, B: k = 6) for both records of the pair; >= 0 <=> inside
__device__ __forceinline__ float2 insideness2(const float4* q, int k, float2 Px, float2 Py, float2 Pz)
{
    const float4 a = q[k], b = q[k + 1], c = q[k + 2], e = q[k + 3];
    const float2 u = __ffma2_rn(Px, lo2(a), __ffma2_rn(Py, hi2(a), __ffma2_rn(Pz, lo2(b), hi2(b))));
    const float2 v = __ffma2_rn(Px, lo2(c), __ffma2_rn(Py, hi2(c), __ffma2_rn(Pz, lo2(e), hi2(e))));
    const float2 w = __ffma2_rn(__fadd2_rn(u, v), make_float2(-1.f, -1.f), make_float2(1.f, 1.f));
    return make_float2(fminf(fminf(u.x, v.x), w.x), fminf(fminf(u.y, v.y), w.y));
}
// Any hit with eps < t < tmax among the occluders; lanes without a shadow ray pass tmax < 0. Must be called by all 32 lanes of
// a converged warp: a pair of planes that no lane can hit (behind every ray or beyond every tmax) is skipped with one vote.
__device__ __forceinline__ bool groupedAnyHit(const SmallSection& S, V3 o, V3 d, float tmax)
{
    const Ray2 r = makeRay2(o, d);
    float acc = -1.f;
    for (int p = 0; p < S.nPairs; ++p) {
        const float4* q = S.recs + 10 * p;
        float2 det, t;
        planes2(q[0], q[1], r, det, t);
        const bool v0 = !(fabsf(det.x) < FLT_EPSILON) && t.x > FLT_EPSILON && t.x < tmax;
        const bool v1 = !(fabsf(det.y) < FLT_EPSILON) && t.y > FLT_EPSILON && t.y < tmax;
        if (!__any_sync(0xffffffffu, v0 || v1)) continue;
        const float2 Px = __ffma2_rn(t, r.dx, r.ox), Py = __ffma2_rn(t, r.dy, r.oy), Pz = __ffma2_rn(t, r.dz, r.oz);
        const float2 mA = insideness2(q, 2, Px, Py, Pz), mB = insideness2(q, 6, Px, Py, Pz);
        acc = fmaxf(acc, v0 ? fmaxf(mA.x, mB.x) : -1.f);
        acc = fmaxf(acc, v1 ? fmaxf(mA.y, mB.y) : -1.f);
    }
    return acc >= 0.f;
}
// Closest hit: strictly smaller t wins between records, A before B inside one (scene.cpp:193-197); lanes without a ray pass
// want = false.
__device__ __forceinline__ void groupedClosest(const SmallSection& S, V3 o, V3 d, bool want, Hit& h)
{
    const Ray2 r = makeRay2(o, d);
    float best = want ? FLT_MAX : -1.f;
    int bi = -1;
#pragma unroll 2
    for (int p = 0; p < S.nPairs; ++p) {
        const float4* q = S.recs + 10 * p;
        float2 det, t;
        planes2(q[0], q[1], r, det, t);
        const float2 Px = __ffma2_rn(t, r.dx, r.ox), Py = __ffma2_rn(t, r.dy, r.oy), Pz = __ffma2_rn(t, r.dz, r.oz);
        const float2 mA = insideness2(q, 2, Px, Py, Pz), mB = insideness2(q, 6, Px, Py, Pz);
        const bool take0 = fmaxf(mA.x, mB.x) >= 0.f && !(fabsf(det.x) < FLT_EPSILON) && t.x > FLT_EPSILON && t.x < best;
        best = take0 ? t.x : best;
        bi = take0 ? (mA.x >= 0.f ? 4 * p : 4 * p + 1) : bi;
        const bool take1 = fmaxf(mA.y, mB.y) >= 0.f && !(fabsf(det.y) < FLT_EPSILON) && t.y > FLT_EPSILON && t.y < best;
        best = take1 ? t.y : best;
        bi = take1 ? (mA.y >= 0.f ? 4 * p + 2 : 4 * p + 3) : bi;
    }
    if (bi >= 0) { // re-derive u, v and the primitive id of the winner: slot bi = 4 * pair + 2 * record + (A / B)
        const float* f = reinterpret_cast<const float*>(S.recs + 10 * (bi >> 2));
        const int j = (bi >> 1) & 1, k = 8 + 16 * (bi & 1); // float index of n1.x of triangle A / B, record j in the odd floats
        const V3 P = o + best * d;
        h.t = best;
        h.u = fmaf(P.x, f[k + j], fmaf(P.y, f[k + 2 + j], fmaf(P.z, f[k + 4 + j], f[k + 6 + j])));
        h.v = fmaf(P.x, f[k + 8 + j], fmaf(P.y, f[k + 10 + j], fmaf(P.z, f[k + 12 + j], f[k + 14 + j])));
        h.prim = reinterpret_cast<const int*>(S.ids)[bi];
    }
}

// Scene::intersect / Scene::occluded on a small scene staged in shared memory — the ONE implementation behind k_bounce_small
// (production) and k_hook_small (parity hook xrtg_trace_rays with XRTG_FLAG_FAST_HOOK), so the hit ids the hook reports are the
// ids the timed kernel shades. GROUPED: the scene carries a plane-paired block (throughput instantiation, no BoxMesh) and both
// functions must be called by all 32 lanes of a converged warp (lanes without a ray pass want = false); otherwise the
// per-triangle lists.
template <bool GROUPED>
struct SmallTracer {
    SmallSection secAll, secOcc, secOccFull;
    const float4* tris;    // per-triangle list, primitive-id order (closest hit)
    const float4* occTris; // occluders (exact: the same list, emitter proxies skipped by flag)
    int nOcc;
    // outsideHull (GROUPED only): this lane's ray starts on a primitive flagged kMetaShadowOutside — its origin may lie behind a
    // hull plane that was pruned from secOcc, so the whole warp tests the unpruned section for this ray batch
    __device__ __forceinline__ bool occluded(const DScene& sc, bool want, V3 o, V3 d, float tmax, bool outsideHull = false) const
    {
        if constexpr (GROUPED) {
            if (!__any_sync(0xffffffffu, want)) return false; // (e.g. a warp of paths that all face away from the light)
            const bool full = __any_sync(0xffffffffu, want && outsideHull);
            bool occ = groupedAnyHit(full ? secOccFull : secOcc, o, d, want ? tmax : -1.f);
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
    }
    __device__ __forceinline__ void closest(const DScene& sc, bool want, V3 o, V3 d, Hit& h) const
    {
        // (no lane has a ray — the LAST bounce of a path never traces its BSDF ray: that loop was 43 % of the last launch's instructions)
        if (GROUPED && !__any_sync(0xffffffffu, want)) return;
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
            closestHit<false, true>(sc, o, d, false, h, nullptr, tc, tris);
        }
    }
};
// Copies the scene's triangles (plane-paired block or per-triangle lists) into the CTA's shared memory; ends with a barrier.
template <bool GROUPED>
__device__ __forceinline__ SmallTracer<GROUPED> stageSmallTracer(const DScene& sc)
{
    constexpr int kListF4 = GROUPED ? 1 : kTriF4 * kSmallSceneTris;
    __shared__ float4 s_tris[kListF4];
    __shared__ float4 s_occ[(kExact || GROUPED) ? 1 : kTriF4 * kSmallSceneTris];
    __shared__ float4 s_block[GROUPED ? kSmallBlockF4 : 1];
    __shared__ int s_nOcc;
    SmallTracer<GROUPED> tr{};
    if constexpr (GROUPED) {
        for (int k = threadIdx.x; k < sc.smallBlockF4; k += blockDim.x) s_block[k] = sc.smallBlock[k];
        __syncthreads();
        const int4 hd = *reinterpret_cast<const int4*>(s_block);
        tr.secAll = smallSection(s_block, hd.x);
        tr.secOcc = smallSection(s_block, hd.y);
        tr.secOccFull = smallSection(s_block, hd.w);
    }
    else {
        stageSmallScene(sc, s_tris, s_occ, &s_nOcc);
        tr.tris = s_tris;
        tr.occTris = kExact ? s_tris : s_occ;
        tr.nOcc = s_nOcc;
    }
    return tr;
}

// Parity hook (xrtg_trace_rays + XRTG_FLAG_FAST_HOOK on a small scene): caller-supplied rays through the SAME SmallTracer the
// fused bounce kernel uses. Closest hit: rays in q.q0[0] / q.q1[0] -> q.hits; any hit: q.s0 (origin | tmax) / q.s1 -> out[i].w.
template <bool GROUPED>
__global__ void __launch_bounds__(kBlock) k_hook_small(DScene sc, DQueues q, uint32_t n, int anyhit, float4* __restrict__ out)
{
    const SmallTracer<GROUPED> tr = stageSmallTracer<GROUPED>(sc);
    for (uint32_t tile = blockIdx.x; uint64_t(tile) * kBlock < n; tile += gridDim.x) {
        const uint32_t i = tile * kBlock + threadIdx.x;
        const bool live = i < n;
        const float4 a = live ? (anyhit ? q.s0[i] : q.q0[0][i]) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 b = live ? (anyhit ? q.s1[i] : q.q1[0][i]) : make_float4(0.f, 0.f, 1.f, 0.f);
        if (anyhit) {
            // out[i].w holds the primitive the ray starts on (or -1): the renderer's flag for origins behind a pruned hull plane
            const int srcPrim = live ? __float_as_int(out[i].w) : -1;
            const bool outside = srcPrim >= 0 && srcPrim < sc.nPrims && (__float_as_uint(__ldg(sc.prims + 4 * srcPrim + 3).w) & kMetaShadowOutside) != 0;
            const bool occ = tr.occluded(sc, live, xyz(a), xyz(b), a.w, outside);
            if (live) out[i] = make_float4(0.f, 0.f, 0.f, __int_as_float(occ ? 1 : 0));
        }
        else {
            Hit h{FLT_MAX, 0.f, 0.f, -1};
            tr.closest(sc, live, xyz(a), xyz(b), h);
            if (live) q.hits[i] = make_float4(h.prim >= 0 ? h.t : FLT_MAX, h.u, h.v, __int_as_float(h.prim));
        }
    }
}

template <bool GROUPED>
__global__ void __launch_bounds__(kBlock, 5) k_bounce_small(DScene sc, DQueues q, DWave w, int src, int bounce, unsigned long long* stats)
{
    __shared__ uint32_t s_scratch[kBlock / 32 + 1];
    // shading records and lights of a small scene live in shared memory too: the queue traffic streams through L1 and would keep
    // evicting them (every entry reads 4-5 of these words on its dependency chain)
    __shared__ float4 s_prims[4 * kSmallPrims];
    __shared__ DLight s_lights[kSmallLights];
    for (int k = threadIdx.x; k < 4 * sc.nPrims; k += blockDim.x) s_prims[k] = sc.prims[k];
    for (int k = threadIdx.x; k < sc.nLights * int(sizeof(DLight) / sizeof(float4)); k += blockDim.x)
        reinterpret_cast<float4*>(s_lights)[k] = reinterpret_cast<const float4*>(sc.lights)[k];
    const SmallTracer<GROUPED> tracer = stageSmallTracer<GROUPED>(sc);
    auto occluded = [&](bool want, V3 o, V3 d, float tmax, bool outsideHull) -> bool { return tracer.occluded(sc, want, o, d, tmax, outsideHull); };
    auto closest = [&](bool want, V3 o, V3 d, Hit& h) { tracer.closest(sc, want, o, d, h); };

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
    uint32_t nClosest = 0, nShadow = 0, nUntraced = 0;
#if XRT_WARP_APPEND
    const uint32_t lane = laneId();
    uint32_t resNext = 0, resEnd = 0, nEntries = 0; // this warp's reservation of output slots; live entries consumed
#endif
    // The queue entries of the CTA's NEXT tile are copied into shared memory while the current tile is being worked on: the HBM
    // latency of the four queue loads at the head of every tile's dependency chain is hidden behind a whole tile of work.
    // kTmaStage: ONE thread issues four bulk copies (cp.async.bulk, the TMA unit: 4 x 2 KB contiguous rows of the SoA queues) that
    // complete on an mbarrier; otherwise every thread stages its own 4 x 16 B with cp.async (LDGSTS) and waits for its own group.
    __shared__ __align__(128) float4 s_stage[2][4][kBlock];
#if XRT_TMA_STAGE
    __shared__ __align__(8) unsigned long long s_bar[2];
    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; ++b) {
            const uint32_t bar = uint32_t(__cvta_generic_to_shared(&s_bar[b]));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t phaseBits = 0; // bit b = parity the next wait on s_bar[b] expects
    auto stage = [&](int buf, uint64_t tile) {
        const uint64_t e0 = tile * kBlock;
        if (threadIdx.x == 0 && e0 < n) {
            // (the previous readers of this buffer are behind the CTA barriers of the last tile's append)
            const uint32_t bytes = uint32_t(min(uint64_t(kBlock), uint64_t(n) - e0)) * 16u;
            const uint32_t bar = uint32_t(__cvta_generic_to_shared(&s_bar[buf]));
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(4u * bytes) : "memory");
            const float4* src[4] = {hitsIn + e0, in0 + e0, in1 + e0, in2 + e0};
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const uint32_t dst = uint32_t(__cvta_generic_to_shared(&s_stage[buf][a][0]));
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src[a]), "r"(bytes), "r"(bar)
                             : "memory");
            }
        }
    };
    auto stageWait = [&](int buf) {
        const uint32_t bar = uint32_t(__cvta_generic_to_shared(&s_bar[buf])), parity = (phaseBits >> buf) & 1u;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "WAIT_%=:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@!p bra WAIT_%=;\n"
            "}\n" ::"r"(bar), "r"(parity)
            : "memory");
        phaseBits ^= 1u << buf;
    };
#else
    auto stage = [&](int buf, uint64_t tile) {
        const uint64_t e = tile * kBlock + threadIdx.x;
        if (e < n) {
            const float4* src[4] = {hitsIn + e, in0 + e, in1 + e, in2 + e};
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const uint32_t dst = uint32_t(__cvta_generic_to_shared(&s_stage[buf][a][threadIdx.x]));
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src[a]) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto stageWait = [&](int) { asm volatile("cp.async.wait_group 1;" ::: "memory"); }; // everything but the group just committed has landed
#endif
    stage(0, blockIdx.x);
    int buf = 0;
    for (uint32_t tile = blockIdx.x; uint64_t(tile) * kBlock < n; tile += gridDim.x, buf ^= 1) {
        const uint32_t i = tile * kBlock + threadIdx.x;
        bool live = i < n;
        stage(buf ^ 1, uint64_t(tile) + gridDim.x);
        stageWait(buf);
        // (warp-chunked appends leave up to kAppendChunk - 1 unused slots per warp at the end of the previous launch: marked dead)
        if (XRT_WARP_APPEND && live && deadEntry(s_stage[buf][3][threadIdx.x])) live = false;
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
#if XRT_WARP_APPEND
            ++nEntries;
#endif
            const float4 hv = s_stage[buf][0][threadIdx.x], r0 = s_stage[buf][1][threadIdx.x], r1 = s_stage[buf][2][threadIdx.x], r2 = s_stage[buf][3][threadIdx.x];
            const V3 o = xyz(r0);
            d = xyz(r1);
            T = mk(r0.w, r1.w, r2.x);
            pid = uint32_t(__float_as_int(r2.y));
            depth = __float_as_int(r2.z);
            rad = q.radiance[pid];
            const Hit h{hv.x, hv.y, hv.z, __float_as_int(hv.w)};
            rng.open(w, pid, uint32_t(__float_as_int(r2.w)));
            makeSurfT<true>(sc, s_prims, o, d, h, s);
            if (kind == XRTG_INT_DIRECT) {
                if (lightOf(s) >= 0) add(emittedFrom(s_lights, s, d));
                else shadeLights = true;
            }
            else if (kind == XRTG_INT_WHITTED) shadeDelta = hasMaterial(s);
            else { // Indirect / GI. Entries of bounce > 0 already passed RR and the emitter test in the kernel that traced them.
                bool alive = true;
                if (bounce == 0 && lightOf(s) >= 0) { add(T * emittedFrom(s_lights, s, d)); alive = false; }
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
                    const V3 Lr = sampleLight(s_lights[li], s.pos, wi, pdf, tmax, rng);
                    if (pdf != 0) {
                        const float cs = smax(0.0f, dot(s.ng, wi));
                        const V3 fr = evalBxDF(s);
                        c = T * (fr * Lr * cs / pdf);
                        want = true;
                        ++nShadow;
                        // an exactly-zero contribution needs no shadow ray (see k_shade_surface); a warp whose lanes all face away
                        // from the light skips the occluder loop altogether, and the plane votes see fewer reachable planes
                        if (!kExact && c.x == 0.f && c.y == 0.f && c.z == 0.f) { want = false; ++nUntraced; }
                    }
                }
                const float bias = 0.01f;
                if (!occluded(want, s.pos + s.ng * bias, wi, tmax - bias, (s.meta & kMetaShadowOutside) != 0) && want) add(c);
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
                if (!occluded(shadeDelta, s.pos + s.ng * float(0.1), wi, tmax, (s.meta & kMetaShadowOutside) != 0) && shadeDelta) add(c);
            }
        }
        // ---- phase 3: BSDF bounce (integrator.h:271-283), then intersect + RR + emitter test of depth+1 (integrator.h:214-245) ----
        bool wantNext = false, trace = false;
        V3 no = mk(0.f), nd = mk(0.f), nT = mk(0.f);
        Hit nh{FLT_MAX, 0.f, 0.f, -1};
        // (throughput instantiation: at the last depth the BSDF sample is not even drawn — nothing follows it on the path's counter
        //  stream; the exact one must consume the draws, the next sample of the pixel continues the same mt19937 stream)
        if (bsdf && (kExact || bounce + 1 < w.maxDepth)) {
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
                    const uint32_t meta = __float_as_uint(s_prims[4 * nh.prim + 3].w);
                    if (((meta >> kMetaLightShift) & 0xfffu) == 0) wantNext = true;
                    else if (kind == XRTG_INT_INDIRECT) { // Le at any depth (integrator.h:150-160); GI only at depth 0
                        Surf s2;
                        makeSurfT<true>(sc, s_prims, no, nd, nh, s2);
                        add(nT * emittedFrom(s_lights, s2, nd));
                    }
                }
            }
        }
        if (live) {
            ctr = rng.close();
            if (radDirty) q.radiance[pid] = rad;
        }
#if XRT_WARP_APPEND
        // Survivors are appended per WARP from a warp-private reservation of kAppendChunk output slots (one global atomic per
        // kAppendChunk survivors of the warp): no CTA barrier anywhere in the tile loop, so the four warps of a CTA drift apart
        // and hide each other's latencies. The slots a warp has not used when it runs out of tiles are marked dead (below).
        uint32_t slot = 0;
        {
            const uint32_t mask = __ballot_sync(0xffffffffu, wantNext);
            if (mask) slot = warpReserve(nextCount, __popc(mask), resNext, resEnd).at(__popc(mask & ((1u << lane) - 1u)));
        }
#else
        const uint32_t slot = blockAppend<kBlock / 32>(nextCount, wantNext, s_scratch);
#endif
        if (wantNext) {
            out0[slot] = make_float4(no.x, no.y, no.z, nT.x);
            out1[slot] = make_float4(nd.x, nd.y, nd.z, nT.y);
            out2[slot] = make_float4(nT.z, __int_as_float(int(pid)), __int_as_float(depth + 1), __int_as_float(int(ctr)));
            hitsOut[slot] = make_float4(nh.t, nh.u, nh.v, __int_as_float(nh.prim));
        }
    }
#if XRT_WARP_APPEND
    warpMarkDead(out2, hitsOut, resNext, resEnd); // the unused tail of this warp's last reservation
    statAdd(stats, kStatBounceEntries, nEntries);
#else
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(stats + kStatBounceEntries, (unsigned long long)n);
#endif
    statAdd(stats, kStatClosest, nClosest);
    statAdd(stats, kStatShadow, nShadow);
    if (!kExact) { statAdd(stats, kStatScissored, nUntraced); statAdd(stats, kStatUntracedShadow, nUntraced); } // counted as the reference's calls, not traced
}
