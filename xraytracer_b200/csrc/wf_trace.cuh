// wf_trace.cuh — part of wavefront.cuh (included inside namespace xrt::XRT_NS, in this order): queue helpers, ray generation, the traversal kernels (k_trace, k_extend_simple, k_connect_simple) and k_primary.
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

// path id = s * wavePixels + (pixel - pixelBase), so consecutive threads own consecutive pixels of the same sample.
// jitter (optional, device): [(pixel * spp + s) * 2] floats supplied by the parity hook.
__global__ void __launch_bounds__(kBlock) k_raygen(DCamera cam, DQueues q, DWave w, const float* __restrict__ jitter)
{
    const uint32_t n = w.nPaths;
    for (uint32_t pid = blockIdx.x * blockDim.x + threadIdx.x; pid < n; pid += gridDim.x * blockDim.x) {
        uint32_t s, lp, i, j;
        w.byWavePixels.divmod(pid, s, lp);
        const uint32_t pix = w.pixelBase + lp;
        w.byWidth.divmod(pix, i, j);
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
        qstore(q.q0[0] + pid, make_float4(o.x, o.y, o.z, 1.0f));
        qstore(q.q1[0] + pid, make_float4(d.x, d.y, d.z, 1.0f));
        qstore(q.q2[0] + pid, make_float4(1.0f, __int_as_float(int(pid)), __int_as_float(0), __int_as_float(int(ctr))));
        qstore(q.radiance + pid, make_float4(0.f, 0.f, 0.f, 0.f));
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
#ifndef XRT_FETCH_CHUNK
#define XRT_FETCH_CHUNK 64
#endif
constexpr uint32_t kFetchChunk = XRT_FETCH_CHUNK;  // queue entries a warp reserves per atomic (256 costs up to 16 % in tail imbalance)

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
// Four-child node step (Bvh4Node): four slab tests — two children per packed fp32x2 instruction —, the hit children sorted by
// entry distance with a 5-comparator network, the nearest becomes the current node, the others are pushed farthest first.
__device__ __forceinline__ float2 f2lo(const float4 v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 f2hi(const float4 v) { return make_float2(v.z, v.w); }
__device__ __forceinline__ void cmpSwap(float& ka, int& ea, float& kb, int& eb)
{
    const bool sw = kb < ka;
    const float k0 = fminf(ka, kb), k1 = fmaxf(ka, kb);
    const int e0 = sw ? eb : ea, e1 = sw ? ea : eb;
    ka = k0; kb = k1; ea = e0; eb = e1;
}
template <bool COUNT>
__device__ __forceinline__ void nodeStep4(const DScene& sc, RayState& r, int* sstack, int* lstack, TraceCounters& tc)
{
    const float4* __restrict__ nd = sc.nodes4 + 8 * size_t(r.node);
    const float4 lx = __ldg(nd), ly = __ldg(nd + 1), lz = __ldg(nd + 2), hx = __ldg(nd + 3), hy = __ldg(nd + 4), hz = __ldg(nd + 5);
    const int4 ch = __ldg(reinterpret_cast<const int4*>(nd + 6)), cn = __ldg(reinterpret_cast<const int4*>(nd + 7));
    if (COUNT) tc.nodes++;
    const float2 ix = make_float2(r.idir.x, r.idir.x), iy = make_float2(r.idir.y, r.idir.y), iz = make_float2(r.idir.z, r.idir.z);
    const float2 ox = make_float2(-r.ood.x, -r.ood.x), oy = make_float2(-r.ood.y, -r.ood.y), oz = make_float2(-r.ood.z, -r.ood.z);
    // children (0,1) in the low halves, (2,3) in the high halves
    const float2 ax0 = __ffma2_rn(f2lo(lx), ix, ox), ax1 = __ffma2_rn(f2lo(hx), ix, ox), bx0 = __ffma2_rn(f2hi(lx), ix, ox), bx1 = __ffma2_rn(f2hi(hx), ix, ox);
    const float2 ay0 = __ffma2_rn(f2lo(ly), iy, oy), ay1 = __ffma2_rn(f2lo(hy), iy, oy), by0 = __ffma2_rn(f2hi(ly), iy, oy), by1 = __ffma2_rn(f2hi(hy), iy, oy);
    const float2 az0 = __ffma2_rn(f2lo(lz), iz, oz), az1 = __ffma2_rn(f2lo(hz), iz, oz), bz0 = __ffma2_rn(f2hi(lz), iz, oz), bz1 = __ffma2_rn(f2hi(hz), iz, oz);
    auto nearFar = [&](float x0, float x1, float y0, float y1, float z0, float z1, float& tn) -> bool {
        tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));
        const float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), r.h.t));
        return tn <= tf;
    };
    float k0, k1, k2, k3;
    const bool h0 = nearFar(ax0.x, ax1.x, ay0.x, ay1.x, az0.x, az1.x, k0) & (cn.x >= 0);
    const bool h1 = nearFar(ax0.y, ax1.y, ay0.y, ay1.y, az0.y, az1.y, k1) & (cn.y >= 0);
    const bool h2 = nearFar(bx0.x, bx1.x, by0.x, by1.x, bz0.x, bz1.x, k2) & (cn.z >= 0);
    const bool h3 = nearFar(bx0.y, bx1.y, by0.y, by1.y, bz0.y, bz1.y, k3) & (cn.w >= 0);
    int e0 = cn.x > 0 ? ~((ch.x << 2) | (cn.x - 1)) : ch.x, e1 = cn.y > 0 ? ~((ch.y << 2) | (cn.y - 1)) : ch.y;
    int e2 = cn.z > 0 ? ~((ch.z << 2) | (cn.z - 1)) : ch.z, e3 = cn.w > 0 ? ~((ch.w << 2) | (cn.w - 1)) : ch.w;
    k0 = h0 ? k0 : FLT_MAX; k1 = h1 ? k1 : FLT_MAX; k2 = h2 ? k2 : FLT_MAX; k3 = h3 ? k3 : FLT_MAX;
    const int nHit = int(h0) + int(h1) + int(h2) + int(h3);
    cmpSwap(k0, e0, k1, e1); cmpSwap(k2, e2, k3, e3); cmpSwap(k0, e0, k2, e2); cmpSwap(k1, e1, k3, e3); cmpSwap(k1, e1, k2, e2);
    if (nHit > 3) stackPush(sstack, lstack, r.sp, e3);
    if (nHit > 2) stackPush(sstack, lstack, r.sp, e2);
    if (nHit > 1) stackPush(sstack, lstack, r.sp, e1);
    if (nHit > 0) r.node = e0;
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
template <bool ANY, bool COUNT, bool WIDE>
__global__ void __launch_bounds__(kBlock, 8) k_trace(DScene sc, DQueues q, int src, int bounce, int brute, unsigned long long* stats, float4* anyOut,
                                                  int refillThreshold, int stepsPerVote, int leafThreshold)
{
    __shared__ int s_stack[kStackSmem * kBlock];
    int lstack[kStackLocalDeep]; // (local memory, touched only as deep as a ray's stack actually grows)
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
                if (active && atInner(r)) {
                    if (WIDE) nodeStep4<COUNT>(sc, r, sstack, lstack, tc); // deep trees carry a four-child form of the tree (api.cu)
                    else nodeStep<COUNT>(sc, r, sstack, lstack, tc);
                }
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
__global__ void __launch_bounds__(kBlock) k_connect_simple(DScene sc, DQueues q, int bounce, int brute, unsigned long long* stats, float4* anyOut)
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
            if (anyOut) anyOut[i] = make_float4(0.f, 0.f, 0.f, __int_as_float(occ ? 1 : 0)); // parity hook: the flag, no contribution
            else if (!occ) {
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
#ifndef XRT_SHADE_MINB
#define XRT_SHADE_MINB 4
#endif
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
// The same with TWO scratch buffers used alternately (`phase` flips on every call): the barriers of the next call protect the
// previous call's buffer, so the third barrier goes away — two barriers per append instead of three.
template <int NWARPS = kShadeWarps>
__device__ __forceinline__ uint32_t blockAppendAlt(uint32_t* counter, bool want, uint32_t (*scratch2)[NWARPS + 1], int& phase)
{
    uint32_t* scratch = scratch2[phase];
    phase ^= 1;
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
    return scratch[NWARPS] + scratch[warp] + __popc(mask & ((1u << lane) - 1u));
}
// Two candidates per thread, one atomic and one barrier round per CTA: slots of a warp are laid out as [its A entries][its B entries].
template <int NWARPS>
__device__ __forceinline__ void blockAppend2(uint32_t* counter, bool wantA, bool wantB, uint32_t* scratch, uint32_t& slotA, uint32_t& slotB)
{
    const uint32_t ma = __ballot_sync(0xffffffffu, wantA), mb = __ballot_sync(0xffffffffu, wantB);
    const uint32_t warp = threadIdx.x >> 5, lane = laneId();
    if (lane == 0) scratch[warp] = __popc(ma) + __popc(mb);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t total = 0;
#pragma unroll
        for (int w = 0; w < NWARPS; ++w) { const uint32_t c = scratch[w]; scratch[w] = total; total += c; }
        scratch[NWARPS] = total ? atomicAdd(counter, total) : 0u;
    }
    __syncthreads();
    const uint32_t base = scratch[NWARPS] + scratch[warp], below = (1u << lane) - 1u;
    slotA = base + __popc(ma & below);
    slotB = base + __popc(ma) + __popc(mb & below);
    __syncthreads(); // scratch may be reused by the next call
}
// Warp-private output chunks (k_primary, k_bounce_small): a warp owns a reservation [resNext, resEnd) of kAppendChunk consecutive queue
// slots and takes a new one with ONE global atomic when its `cnt` new entries do not fit; entries that still fit go to the old
// chunk, so the only unused slots are the tail of a warp's LAST chunk, which warpMarkDead() flags with kDeadPath in the path word
// (every consumer of such a queue skips those). No CTA barrier, no shared memory; must be called by all 32 lanes.
struct WarpSlots {
    uint32_t base0, left, base1;
    __device__ __forceinline__ uint32_t at(uint32_t rank) const { return rank < left ? base0 + rank : base1 + (rank - left); }
};
__device__ __forceinline__ WarpSlots warpReserve(uint32_t* counter, uint32_t cnt, uint32_t& resNext, uint32_t& resEnd)
{
    WarpSlots s{resNext, resEnd - resNext, 0u};
    if (cnt > s.left) {
        uint32_t nb = 0;
        if (laneId() == 0) nb = atomicAdd(counter, kAppendChunk);
        nb = __shfl_sync(0xffffffffu, nb, 0);
        s.base1 = nb;
        resNext = nb + (cnt - s.left);
        resEnd = nb + kAppendChunk;
    }
    else resNext += cnt;
    return s;
}
__device__ __forceinline__ void warpMarkDead(float4* __restrict__ pathWords, float4* __restrict__ hits, uint32_t resNext, uint32_t resEnd)
{
    for (uint32_t sl = resNext + laneId(); sl < resEnd; sl += 32u) {
        pathWords[sl] = make_float4(0.f, __uint_as_float(kDeadPath), 0.f, 0.f);
        hits[sl] = make_float4(FLT_MAX, 0.f, 0.f, __int_as_float(-1));
    }
}
__device__ __forceinline__ bool deadEntry(const float4& pathWord) { return __float_as_uint(pathWord.y) == kDeadPath; }
// ---------------------------------------------------------------------------------------------------------
// primary: ray generation (renderer.cpp:42-52, camera.h:49-60) FUSED with the bounce-0 closest hit. Primary rays are
// coherent and most of them miss in the benchmark views (59 % Cornell, 86 % volume), so instead of writing 8.3 M rays,
// reading them back, writing 8.3 M hit records and letting shade skip the misses, this kernel resolves misses inline
// (background 0, DirectIntegrator's 0.18 grey integrator.h:114, Whitted's sky integrator.h:385-389) and appends ONLY the
// hits — ray + hit record — to a COMPACT bounce-0 queue (one atomic per CTA per 128 rays). It also initialises the
// per-path radiance. Static tile partition: CTA b owns path ids [128 b, 128 b + 128), b += grid.
// ---------------------------------------------------------------------------------------------------------
// Screen-space candidate masks of the primary rays (small scenes): bit t of segment s (= 32 consecutive pixels in linear index)
// is set when the pixel bounding box of mesh triangle t, bbox[t] = (x0, y0, x1, y1) with exclusive upper bounds, computed on
// the host by projecting its vertices through PinholeCamera::sampleRay's inverse (camera.h:49-60) with a pixel of margin, touches
// the segment; x0 > x1 marks a triangle with a vertex beside / behind the camera (no valid projection): every segment gets it.
__global__ void __launch_bounds__(kBlock) k_primary_masks(const TriBoxes boxes, int nTris, int width, uint32_t nPixels, unsigned long long* __restrict__ masks)
{
    const int4* bbox = boxes.b;
    const uint32_t nSeg = (nPixels + 31u) / 32u;
    for (uint32_t sgm = blockIdx.x * blockDim.x + threadIdx.x; sgm < nSeg; sgm += gridDim.x * blockDim.x) {
        const uint32_t p0 = 32u * sgm, p1 = min(p0 + 31u, nPixels - 1u);
        const int r0 = int(p0 / uint32_t(width)), r1 = int(p1 / uint32_t(width));
        unsigned long long m = 0ull;
        for (int t = 0; t < nTris; ++t) {
            const int4 b = bbox[t];
            bool hit = b.x > b.z;
            for (int r = max(r0, b.y); r <= r1 && r < b.w && !hit; ++r) { // rows of the segment inside the box's rows
                const int c0 = r == r0 ? int(p0 - uint32_t(r0) * uint32_t(width)) : 0;
                const int c1 = r == r1 ? int(p1 - uint32_t(r1) * uint32_t(width)) : width - 1;
                hit = c0 < b.z && c1 >= b.x;
            }
            if (hit) m |= 1ull << t;
        }
        masks[sgm] = m;
    }
}

// Scene::intersect (scene.cpp:190-200) restricted to the mesh triangles in `mask` (bit = index in primitive-id order): boxes and
// spheres as in closestHit(). The mask is made warp-uniform by the caller so that the warp stays converged.
template <bool COUNT>
__device__ __forceinline__ void closestHitMasked(const DScene& sc, V3 o, V3 d, Hit& h, unsigned long long mask, TraceCounters& tc)
{
    h.t = FLT_MAX; h.u = 0.f; h.v = 0.f; h.prim = 0x7fffffff;
    int minId = -1;
    for (int b = 0; b < sc.nBoxes; ++b) { // boxes[] are in object order
        const float4 bl = __ldg(sc.boxes + 2 * b), bh = __ldg(sc.boxes + 2 * b + 1);
        float t0, t1;
        if (boxSlabs(xyz(bl), xyz(bh), o, d, t0, t1)) { h.t = t0; h.u = t1; h.v = 0.f; h.prim = __float_as_int(bl.w); minId = h.prim; }
    }
    const float4* __restrict__ tris = triArray(sc, true);
    while (mask) {
        const int t = __ffsll((long long)mask) - 1;
        mask &= mask - 1ull;
        if (COUNT) tc.tris++;
        triangleRecord<false, true>(tris + kTriF4 * t, o, d, h, minId);
    }
    for (int s = 0; s < sc.nSpheres; ++s) {
        const float4 cr = __ldg(sc.spheres + 2 * s);
        const int4 meta = __ldg(reinterpret_cast<const int4*>(sc.spheres + 2 * s + 1));
        float t;
        if (meta.x > minId && sphereT(cr, o, d, t)) consider(h, t, 0.f, 0.f, meta.x);
    }
    if (h.prim == 0x7fffffff) h.prim = -1;
}

// JITTER: the two primary-sample offsets come from `jitter` (parity hook xrtg_trace_primary) instead of the path's RNG — the only
// difference between the hook's instantiation and the production one.
template <bool COUNT, bool JITTER>
__global__ void __launch_bounds__(kBlock) k_primary(DScene sc, DCamera cam, DQueues q, DWave w, int brute, int missMode, unsigned long long* stats,
                                                     const float* __restrict__ jitter)
{
    __shared__ int s_stack[kStackSmem * kBlock];
    __shared__ uint32_t s_scratch[kBlock / 32 + 1];
    const uint32_t n = w.nPaths;
    TraceCounters tc;
    uint32_t nHits = 0, nScissored = 0;
    uint32_t resNext = 0, resEnd = 0; // this warp's reservation of slots in the compact bounce-0 queue (XRT_WARP_APPEND_PRIMARY)
    (void)resNext; (void)resEnd;
    // one path: ray generation, scissor, closest hit; misses are resolved here
    auto path = [&](uint32_t pid, V3& d, Hit& h, uint32_t& ctr) -> bool {
        if (pid >= n) return false;
        uint32_t smp, lp, i, j;
        w.byWavePixels.divmod(pid, smp, lp);
        const uint32_t pix = w.pixelBase + lp;
        w.byWidth.divmod(pix, i, j);
        const bool inView = int(j) >= w.sx0 && int(j) < w.sx1 && int(i) >= w.sy0 && int(i) < w.sy1;
        bool hit = false;
        // small scenes: the triangles that can be seen through this warp's pixels (one mask per 32 pixels; OR over the warp's
        // lanes, which may straddle two segments, keeps the candidate loop converged)
        unsigned long long cand = 0ull;
        if (w.primMask != nullptr && !brute) {
            const unsigned long long mine = inView ? __ldg(w.primMask + (pix >> 5)) : 0ull;
            const uint32_t act = __activemask();
            cand = (unsigned long long)__reduce_or_sync(act, uint32_t(mine)) | ((unsigned long long)__reduce_or_sync(act, uint32_t(mine >> 32)) << 32);
        }
        if (inView) { // (outside: the pixel cannot see the scene's bounding box, every sample is a miss, no ray needed)
            float r0, r1;
            if constexpr (JITTER) { // parity hook (xrtg_trace_primary): caller-supplied jitter, laid out [(pixel * spp + s) * 2]
                const size_t k = (size_t(pix) * w.samplesThisWave + smp) * 2;
                r0 = jitter[k]; r1 = jitter[k + 1];
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
            V3 o;
            cameraRay(cam, u, v, o, d);
            if (w.primMask != nullptr && !brute) closestHitMasked<COUNT>(sc, o, d, h, cand, tc);
            else closestHit<COUNT>(sc, o, d, brute != 0, h, s_stack + threadIdx.x, tc);
            hit = h.prim >= 0;
        }
        else ++nScissored;
        V3 c = mk(0.f);
        if (!hit) {
            if (missMode == 1) c = mk(float(0.18));
            else if (missMode == 2) c = mk(1.f) * mk(float(0.235294), float(0.67451), float(0.843137));
        }
        if (inView || !(w.flags & kWaveScissorSkip)) q.radiance[pid] = make_float4(c.x, c.y, c.z, 0.f); // (see k_accumulate)
        return hit;
    };
    const V3 org = mk(cam.c2w[12], cam.c2w[13], cam.c2w[14]); // every primary ray starts at the camera position (camera.h:57)
    auto emit = [&](uint32_t slot, uint32_t pid, V3 d, const Hit& h, uint32_t ctr) {
        q.q0[0][slot] = make_float4(org.x, org.y, org.z, 1.0f);
        q.q1[0][slot] = make_float4(d.x, d.y, d.z, 1.0f);
        q.q2[0][slot] = make_float4(1.0f, __int_as_float(int(pid)), __int_as_float(0), __int_as_float(int(ctr)));
        q.hits[slot] = make_float4(h.t, h.u, h.v, __int_as_float(h.prim));
    };
    // two tiles of 128 paths per round: half the barriers and atomics of the queue append per path
    for (uint32_t tile = 2 * blockIdx.x; uint64_t(tile) * kBlock < n; tile += 2 * gridDim.x) {
        const uint32_t pidA = tile * kBlock + threadIdx.x, pidB = pidA + kBlock;
        V3 dA = mk(0.f), dB = mk(0.f);
        Hit hA{FLT_MAX, 0.f, 0.f, -1}, hB{FLT_MAX, 0.f, 0.f, -1};
        uint32_t cA = 0, cB = 0;
        const bool hitA = path(pidA, dA, hA, cA);
        const bool hitB = path(pidB, dB, hB, cB);
        uint32_t slotA, slotB;
#if XRT_WARP_APPEND_PRIMARY
        {
            const uint32_t ma = __ballot_sync(0xffffffffu, hitA), mb = __ballot_sync(0xffffffffu, hitB), below = (1u << laneId()) - 1u;
            const WarpSlots ws = warpReserve(q.ctrl + kCtrlRays, __popc(ma) + __popc(mb), resNext, resEnd);
            slotA = ws.at(__popc(ma & below));
            slotB = ws.at(__popc(ma) + __popc(mb & below));
        }
#else
        blockAppend2<kBlock / 32>(q.ctrl + kCtrlRays, hitA, hitB, s_scratch, slotA, slotB);
#endif
        nHits += (hitA ? 1u : 0u) + (hitB ? 1u : 0u);
        if (hitA) emit(slotA, pidA, dA, hA, cA);
        if (hitB) emit(slotB, pidB, dB, hB, cB);
    }
#if XRT_WARP_APPEND_PRIMARY
    warpMarkDead(q.q2[0], q.hits, resNext, resEnd);
#endif
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(stats + kStatClosest, (unsigned long long)n);
    statAdd(stats, kStatPrimaryHits, nHits);
    statAdd(stats, kStatScissored, nScissored); // reference-equivalent rays that were never traced
    statAdd(stats, kStatUntracedClosest, nScissored);
    if (COUNT) { statAdd(stats, kStatNodes, tc.nodes); statAdd(stats, kStatTris, tc.tris); }
}
