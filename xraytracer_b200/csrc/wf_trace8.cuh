// wf_trace8.cuh — part of wavefront.cuh (included inside namespace xrt::XRT_NS, after wf_trace.cuh): k_trace8, the traversal kernel of
// deep scenes in the THROUGHPUT instantiation — eight-child nodes with 8-bit quantised child boxes (bvh.h: Bvh8Node, 80 B) and the
// node-ordered plane-equation triangle records (ftris8). Same outer structure as k_trace (persistent warps, per-lane resumable
// state, idle lanes refilled from a warp-private reservation of queue entries, triangle tests postponed until enough lanes of the
// warp have some), but the traversal state is the compressed-wide-BVH one (Ylitie, Karras & Laine 2017):
//   * a lane's stack holds GROUPS, not nodes: (childBase, hit bits | imask) = all children of one node that are still to visit
//     -> at most one entry per tree level;
//   * children are visited in the fixed order "decreasing (slot ^ octant-inverse)" the builder arranged the slots for: no
//     per-ray distance sort, the 8 slab tests produce one 8-bit mask;
//   * a node test costs ONE dependent 80-byte fetch (five 16 B loads) for eight children: ~4.5 dependent node fetches per ray
//     on the 1 M-triangle scene instead of ~7.8 four-child nodes of 128 B, and the whole tree is 6 MB instead of 25 MB.
// The child boxes are supersets of the padded boxes of the two-child tree and the slab test keeps nodes whose entry distance
// equals the current best, so the candidates that reach the triangle tests are a superset of the exact traversal's and the
// result is decided by the same triangleRecord() as everywhere else (ties -> lowest primitive id through consider()).
constexpr int kStack8Smem = 12;  // group-stack entries (8 B) per thread in shared memory
constexpr int kStack8Local = 52; // overflow; api.cu only builds the eight-child tree when 2 * depth fits

struct Ray8State {
    V3 o, d, idir;
    Hit h;           // closest: current best; any: h.t = tmax, h.prim = occluded flag
    int minId;
    uint32_t gBase, gBits; // current node group: first child node | (hit bits in visiting priority, bits 0..7) | imask << 8
    uint32_t tBase, tBits, tValid; // current triangle group: first triangle | hit bits | the node's validTri (bit -> triangle index)
    uint32_t octinv;       // 7 - direction octant (bit a set <=> d[a] >= 0)
    int sp;
    uint32_t qidx;
    bool done;
};

__device__ __forceinline__ void push8(uint2* sstack, uint2* lstack, int& sp, uint32_t x, uint32_t y)
{
    if (sp < kStack8Smem) sstack[sp * kBlock] = make_uint2(x, y);
    else lstack[sp - kStack8Smem] = make_uint2(x, y);
    ++sp;
}
__device__ __forceinline__ uint2 pop8(const uint2* sstack, const uint2* lstack, int& sp)
{
    --sp;
    return (sp < kStack8Smem) ? sstack[sp * kBlock] : lstack[sp - kStack8Smem];
}

// Byte k of `w` as the float 2^23 + b: ONE PRMT builds the bit pattern. The "- 2^23" is not executed per byte: it is folded into
// the additive term of the slab FMA, t = (2^23 + b) * s + (c - 2^23 * s). The product 2^23 * s is exact (power of two) and the
// FMA rounds once, so the only extra error is the rounding of c - 2^23 * s: at most half a quantisation step (s / 2) in t. The
// builder pays for it with one step of margin on each side of every child box (bvh.cpp: collapseBvh8), which keeps the test
// conservative; 48 FADD per node disappear.
__device__ __forceinline__ float byteMagic(uint32_t w, int k) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650 + k)); }

// One node: the eight quantised slab tests, two children per packed fp32x2 FMA. Writes the node's groups into r.
template <bool COUNT>
__device__ __forceinline__ void nodeStep8(const DScene& sc, Ray8State& r, uint32_t nodeIdx, TraceCounters& tc)
{
    const uint4* __restrict__ nd = sc.nodes8 + 5 * size_t(nodeIdx);
    const uint4 n0 = __ldg(nd), n1 = __ldg(nd + 1), n2 = __ldg(nd + 2), n3 = __ldg(nd + 3), n4 = __ldg(nd + 4);
    if (COUNT) tc.nodes++;
    // per-axis scale 2^e folded into the reciprocal direction; origin of the node frame relative to the ray, minus 2^23 steps
    const float sx = r.idir.x * __uint_as_float((n0.w & 0xffu) << 23), sy = r.idir.y * __uint_as_float(((n0.w >> 8) & 0xffu) << 23),
                sz = r.idir.z * __uint_as_float(((n0.w >> 16) & 0xffu) << 23);
    const float cx = fmaf(-8388608.0f, sx, (__uint_as_float(n0.x) - r.o.x) * r.idir.x), cy = fmaf(-8388608.0f, sy, (__uint_as_float(n0.y) - r.o.y) * r.idir.y),
                cz = fmaf(-8388608.0f, sz, (__uint_as_float(n0.z) - r.o.z) * r.idir.z);
    const float2 SX = make_float2(sx, sx), SY = make_float2(sy, sy), SZ = make_float2(sz, sz);
    const float2 CX = make_float2(cx, cx), CY = make_float2(cy, cy), CZ = make_float2(cz, cz);
    const uint32_t imask = n0.w >> 24;
    // near / far planes by the sign of the direction: lo is near where d >= 0
    const bool px = (r.octinv & 1u) != 0, py = (r.octinv & 2u) != 0, pz = (r.octinv & 4u) != 0;
    const uint32_t nx[2] = {px ? n2.x : n3.z, px ? n2.y : n3.w}, fx[2] = {px ? n3.z : n2.x, px ? n3.w : n2.y}; // qlo.x = n2.xy, qhi.x = n3.zw
    const uint32_t ny[2] = {py ? n2.z : n4.x, py ? n2.w : n4.y}, fy[2] = {py ? n4.x : n2.z, py ? n4.y : n2.w}; // qlo.y = n2.zw, qhi.y = n4.xy
    const uint32_t nz[2] = {pz ? n3.x : n4.z, pz ? n3.y : n4.w}, fz[2] = {pz ? n4.z : n3.x, pz ? n4.w : n3.y}; // qlo.z = n3.xy, qhi.z = n4.zw
    // H: 0xF in nibble k for every child k whose box the ray enters before the current best (empty slots carry inverted boxes)
    uint32_t H = 0;
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
        const int w = k >> 2, b = k & 3;
        const float2 tnx = __ffma2_rn(make_float2(byteMagic(nx[w], b), byteMagic(nx[w], b + 1)), SX, CX);
        const float2 tfx = __ffma2_rn(make_float2(byteMagic(fx[w], b), byteMagic(fx[w], b + 1)), SX, CX);
        const float2 tny = __ffma2_rn(make_float2(byteMagic(ny[w], b), byteMagic(ny[w], b + 1)), SY, CY);
        const float2 tfy = __ffma2_rn(make_float2(byteMagic(fy[w], b), byteMagic(fy[w], b + 1)), SY, CY);
        const float2 tnz = __ffma2_rn(make_float2(byteMagic(nz[w], b), byteMagic(nz[w], b + 1)), SZ, CZ);
        const float2 tfz = __ffma2_rn(make_float2(byteMagic(fz[w], b), byteMagic(fz[w], b + 1)), SZ, CZ);
        const float tn0 = fmaxf(fmaxf(tnx.x, tny.x), fmaxf(tnz.x, 0.0f)), tf0 = fminf(fminf(tfx.x, tfy.x), fminf(tfz.x, r.h.t));
        const float tn1 = fmaxf(fmaxf(tnx.y, tny.y), fmaxf(tnz.y, 0.0f)), tf1 = fminf(fminf(tfx.y, tfy.y), fminf(tfz.y, r.h.t));
        H |= (tn0 <= tf0 ? (0xFu << (4 * k)) : 0u) | (tn1 <= tf1 ? (0xFu << (4 * k + 4)) : 0u);
    }
    // triangle hits: the hit slots' nibbles of validTri. Node hits: bit 0 of every hit nibble, compressed to one bit per slot,
    // restricted to inner children, then permuted to visiting priority (bit k -> bit k ^ octinv) with three masked swaps.
    uint32_t y = H & 0x11111111u;
    y = (y | (y >> 3)) & 0x03030303u;
    y = (y | (y >> 6)) & 0x000F000Fu;
    y = (y | (y >> 12)) & imask;
    y = (r.octinv & 1u) ? (((y >> 1) & 0x55u) | ((y << 1) & 0xAAu)) : y;
    y = (r.octinv & 2u) ? (((y >> 2) & 0x33u) | ((y << 2) & 0xCCu)) : y;
    y = (r.octinv & 4u) ? (((y >> 4) & 0x0Fu) | ((y << 4) & 0xF0u)) : y;
    r.gBase = n1.x;
    r.gBits = y | (imask << 8);
    r.tBase = n1.y;
    r.tBits = H & n1.z;
    r.tValid = n1.z;
}

template <bool ANY, bool COUNT, int MINB>
__global__ void __launch_bounds__(kBlock, MINB) k_trace8(DScene sc, DQueues q, int src, int bounce, unsigned long long* stats, float4* anyOut, int refillThreshold,
                                                      int stepsPerVote, int leafThreshold)
{
    __shared__ uint2 s_stack[kStack8Smem * kBlock];
    uint2 lstack[kStack8Local];
    uint2* sstack = s_stack + threadIdx.x;
    uint32_t* ctrl = q.ctrl + bounce * kCtrlStride;
    const uint32_t n = ANY ? ctrl[kCtrlShadow] : ctrl[kCtrlRays];
    uint32_t* cursor = ctrl + (ANY ? kCtrlFetchConnect : kCtrlFetchExtend);
    const uint32_t lane = laneId();
    const float4* __restrict__ tris = sc.ftris8;
    TraceCounters tc;
    Ray8State r;
    r.gBits = r.tBits = 0u; r.gBase = r.tBase = 0u; r.tValid = 0u; r.sp = 0; r.qidx = 0; r.minId = -1; r.done = false; r.octinv = 0u;
    bool active = false, exhausted = false;
    uint32_t resNext = 0, resEnd = 0; // the warp's current reservation of queue entries
    while (true) {
        // ---- refill the idle lanes with consecutive queue entries (one atomic per kFetchChunk entries per warp; see k_trace) ----
        const uint32_t need = __ballot_sync(0xffffffffu, !active);
        if (need != 0 && !exhausted) {
            const uint32_t nNeed = __popc(need);
            const uint32_t rank = __popc(need & ((1u << lane) - 1u));
            const uint32_t left = resEnd - resNext;
            uint32_t nb = 0;
            if (nNeed > left) {
                if (lane == 0) nb = atomicAdd(cursor, kFetchChunk);
                nb = __shfl_sync(0xffffffffu, nb, 0);
                if (nb >= n) exhausted = true;
            }
            const uint32_t i = rank < left ? resNext + rank : nb + (rank - left);
            if (nNeed > left) { resNext = nb + (nNeed - left); resEnd = nb + kFetchChunk; }
            else resNext += nNeed;
            if (!active && i < n) {
                r.qidx = i;
                r.done = false;
                if (ANY) {
                    const float4 s0 = qload(q.s0 + i), s1 = qload(q.s1 + i);
                    r.o = xyz(s0); r.d = xyz(s1);
                    r.h.t = s0.w; r.h.u = 0.f; r.h.v = 0.f; r.h.prim = 0; // prim = occluded flag
                    r.minId = -1;
                    if (sc.nBoxes > 0) { r.h.prim = 1; r.done = true; } // BoxMesh::occluded is always true
                }
                else {
                    const float4 r0 = qload(q.q0[src] + i), r1 = qload(q.q1[src] + i);
                    r.o = xyz(r0); r.d = xyz(r1);
                    r.h.t = FLT_MAX; r.h.u = 0.f; r.h.v = 0.f; r.h.prim = kSentinel;
                    r.minId = -1;
                    for (int b = 0; b < sc.nBoxes; ++b) { // last box hit in object order wins (primitive.h:259-261)
                        const float4 bl = __ldg(sc.boxes + 2 * b), bh = __ldg(sc.boxes + 2 * b + 1);
                        float t0, t1;
                        if (boxSlabs(xyz(bl), xyz(bh), r.o, r.d, t0, t1)) { r.h.t = t0; r.h.u = t1; r.h.v = 0.f; r.h.prim = __float_as_int(bl.w); r.minId = r.h.prim; }
                    }
                }
                // box tests only: clamp zero / denormal direction components (see traverse())
                const float kTiny = 1e-20f;
                const V3 ds = mk(fabsf(r.d.x) < kTiny ? copysignf(kTiny, r.d.x) : r.d.x, fabsf(r.d.y) < kTiny ? copysignf(kTiny, r.d.y) : r.d.y,
                                 fabsf(r.d.z) < kTiny ? copysignf(kTiny, r.d.z) : r.d.z);
                r.idir = 1.0f / ds;
                r.octinv = (ds.x >= 0.f ? 1u : 0u) | (ds.y >= 0.f ? 2u : 0u) | (ds.z >= 0.f ? 4u : 0u);
                r.sp = 0;
                r.tBits = 0u; r.tBase = 0u;
                // the root as a one-child group: node 0, priority bit 0 ^ octinv ... simply "visit node gBase + 0"
                r.gBase = 0u;
                r.gBits = (r.done || sc.nTris == 0) ? 0u : (1u << (0u ^ r.octinv)) | (1u << 8); // imask bit 0: slot 0 is an inner node = the root
                active = true;
            }
        }
        if (__ballot_sync(0xffffffffu, active) == 0) break;
        const uint32_t threshold = exhausted ? 1u : uint32_t(refillThreshold);
        uint32_t busy;
        do {
            for (int sv = 0; sv < stepsPerVote; ++sv) {
                // ---- node phase: lanes without pending triangles take the next child of their current group ----
                if (active && r.tBits == 0u && (r.gBits & 0xffu) != 0u) {
                    const uint32_t pri = 31u - __clz(r.gBits & 0xffu);       // highest priority = nearest slot for this octant
                    const uint32_t slot = pri ^ r.octinv;
                    const uint32_t rest = r.gBits & ~(1u << pri);
                    const uint32_t nodeIdx = r.gBase + __popc((r.gBits >> 8) & ((1u << slot) - 1u));
                    if ((rest & 0xffu) != 0u) push8(sstack, lstack, r.sp, r.gBase, rest);   // the siblings wait on the stack as ONE entry
                    nodeStep8<COUNT>(sc, r, nodeIdx, tc);
                }
                // ---- triangle phase: postponed until enough lanes have some (or nobody can advance through nodes) ----
                const bool hasTri = active && r.tBits != 0u;
                const uint32_t triMask = __ballot_sync(0xffffffffu, hasTri);
                const uint32_t advancing = __ballot_sync(0xffffffffu, active && r.tBits == 0u && (r.gBits & 0xffu) != 0u);
                if (triMask != 0u && (__popc(triMask) >= leafThreshold || advancing == 0u)) {
                    // one triangle per lane and round, rounds while at least half of the lanes that started still have one
                    uint32_t left = triMask;
                    const int stop = __popc(triMask) / 2;
                    do {
                        if (r.tBits != 0u) {
                            const uint32_t bit = __ffs(r.tBits) - 1u;
                            r.tBits &= r.tBits - 1u;
                            const uint32_t k = __popc(r.tValid & ((1u << bit) - 1u)); // triangles of this node stored before it
                            if (COUNT) tc.tris++;
                            if (triangleRecord<ANY, true>(tris + 4 * size_t(r.tBase + k), r.o, r.d, r.h, r.minId)) { r.h.prim = 1; r.done = true; r.tBits = 0u; r.gBits = 0u; r.sp = 0; }
                        }
                        left = __ballot_sync(0xffffffffu, r.tBits != 0u);
                    } while (__popc(left) > stop);
                }
                // ---- group exhausted: next group from the stack, or the ray is finished ----
                if (active && r.tBits == 0u && (r.gBits & 0xffu) == 0u) {
                    if (r.sp > 0 && !r.done) {
                        const uint2 e = pop8(sstack, lstack, r.sp);
                        r.gBase = e.x; r.gBits = e.y;
                    }
                    else r.done = true;
                }
                if (active && r.done && r.tBits == 0u) {
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
                            const float4 c = qload(q.s2 + r.qidx);
                            float* rad = reinterpret_cast<float*>(q.radiance + __float_as_int(qload(q.s1 + r.qidx).w));
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
                        qstore(q.hits + r.qidx, make_float4(prim >= 0 ? r.h.t : FLT_MAX, r.h.u, r.h.v, __int_as_float(prim)));
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
