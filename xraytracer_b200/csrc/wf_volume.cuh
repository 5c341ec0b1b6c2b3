// wf_volume.cuh — part of wavefront.cuh (included inside namespace xrt::XRT_NS, in this order): participating media: density lookup, phase function, delta / ratio tracking, k_shade_volume and k_volume_paths.
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
    float fx, fy, fz;
    if constexpr (kExact) { fx = (p.x - g.origin[0]) / g.voxel; fy = (p.y - g.origin[1]) / g.voxel; fz = (p.z - g.origin[2]) / g.voxel; }
    else {
        fx = (p.x - g.origin[0]) * g.invVoxel; fy = (p.y - g.origin[1]) * g.invVoxel; fz = (p.z - g.origin[2]) * g.invVoxel;
        // Throughput instantiation: ONE texture fetch with hardware trilinear filtering instead of eight loads and seven lerps.
        // Texel centres sit at i + 0.5 in unnormalised coordinates, so f + 0.5 reproduces grid.h:71-77's "voxel values at integer
        // index coordinates"; border addressing returns 0 = the background outside the grid. The filter weights carry 8
        // fractional bits (the exact instantiation and the oracle interpolate in fp32): the images stay within the stated
        // statistical bounds of the oracle's (tests/test_gpu_parity.py).
        if (g.tex) return tex3D<float>(cudaTextureObject_t(g.tex), fx + 0.5f, fy + 0.5f, fz + 0.5f);
    }
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
    ++steps;
    rng.alignBlock(); // the three draws of one step come from one Philox block
    if constexpr (!kExact) {
        // Grey medium and a grey throughput (every medium of the reference's examples, and c5): the three channels of the spectral
        // tracker behave alike — the wavelength pmf is (1/3, 1/3, 1/3) whatever channel is drawn, the scattering probability is one
        // number, and the null-collision update sigma_n / (maj * sum(pmf * P_n)) collapses to (sigma_s + sigma_n) / maj. Same
        // draws in the same order, one third of the arithmetic.
        if (m.grey && rayT.x == rayT.y && rayT.y == rayT.z && ts.tt.x == ts.tt.y && ts.tt.y == ts.tt.z) {
            (void)rng.next(); // the wavelength draw
            ts.pmf = mk(1.0f / 3.0f);
            ts.sd = -logf(smax(1.0f - rng.next(), 0.0f)) * m.invMajorant;
            ts.t += ts.sd;
            if (ts.t > ts.t1 - kRayEps) return kTrackExit;
            ts.density = m.densityMul * gridDensity(g, o + ts.t * d);
            const float ss = m.sigma_s[0] * ts.density, sn = m.majorant - m.sigma_a[0] * ts.density - ss;
            if (rng.next() < ss / (ss + sn)) return kTrackScatter;
            ts.tt = ts.tt * ((ss + sn) * m.invMajorant);
            return kTrackContinue;
        }
    }
    const V3 absC = v3(m.sigma_a), scatC = v3(m.sigma_s);
    const V3 maj = mk(m.majorant);
    V3 sigma_a = absC * ts.density;
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
    if constexpr (kExact) {
        const V3 tr = analyticTr(ts.sd, maj);
        const V3 pdf_distance = m.majorant * tr;
        const V3 pdf = ts.pmf * pdf_distance * P_n;
        ts.tt = ts.tt * ((tr * sigma_n) / (pdf.x + pdf.y + pdf.z));
    }
    else {
        // the majorant is the same for the three channels, so is tr = exp(-maj * sd), and it cancels:
        // tr * sigma_n / sum(pmf * maj * tr * P_n) = sigma_n / (maj * sum(pmf * P_n))
        const V3 pn = ts.pmf * P_n;
        ts.tt = ts.tt * (sigma_n / (m.majorant * (pn.x + pn.y + pn.z)));
    }
    return kTrackContinue;
}
// The epilogue of the walk: medium.cpp:73-83 (left the medium) or :100-112 (scattered, new direction from the phase function)
__device__ __forceinline__ void trackFinish(const DMedium& m, V3 o, V3 d, TrackState& ts, int how, Rng& rng, TrackResult& r)
{
    const V3 maj = mk(m.majorant);
    if (how == kTrackExit) {
        r.pos = o + (ts.t1 + kRayEps) * d; r.dir = d;
        if constexpr (kExact) {
            const float rest = ts.sd - (ts.t - (ts.t1 - kRayEps));
            const V3 tr = analyticTr(rest, maj);
            const V3 pdf = ts.pmf * tr;
            ts.tt = ts.tt * (tr / (pdf.x + pdf.y + pdf.z));
        }
        // (throughput instantiation: tr is channel-uniform and the pmf sums to one, so tr / sum(pmf * tr) = 1)
        r.scattered = false;
    }
    else {
        const V3 absC = v3(m.sigma_a), scatC = v3(m.sigma_s);
        const V3 sigma_s = scatC * ts.density, sigma_a = absC * ts.density;
        const V3 sigma_n = maj - sigma_a - sigma_s;
        const V3 P_s = sigma_s / (sigma_s + sigma_n);
        r.pos = o + ts.t * d;
        hgSample(m.g, d, rng, r.dir);
        if constexpr (kExact) {
            const V3 tr = analyticTr(ts.sd, maj);
            const V3 pdf_distance = m.majorant * tr;
            const V3 pdf = ts.pmf * pdf_distance * P_s;
            ts.tt = ts.tt * ((tr * sigma_s) / (pdf.x + pdf.y + pdf.z));
        }
        else {
            const V3 ps = ts.pmf * P_s; // tr cancels as in trackStep
            ts.tt = ts.tt * (sigma_s / (m.majorant * (ps.x + ps.y + ps.z)));
        }
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
// (prims: the shading records, sc.prims or a shared-memory copy of them — read through a generic pointer)
__device__ __forceinline__ int volumePre(const DScene& sc, const float4* prims, const DMedium* media, const DGrid* grids, bool nee, V3 o, V3 d, V3& T, const Hit& h, int depth,
                                         Rng& rng, int& mi, TrackState& ts, TrackResult& r, bool& hasContrib, V3& contrib)
{
    hasContrib = false;
    if (h.prim < 0) return kVolEnd; // a miss adds throughput*background*(depth!=0) = 0
    Surf s;
    makeSurfT<true>(sc, prims, o, d, h, s);
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
        if (i < n && !(XRT_WARP_APPEND_PRIMARY && deadEntry(q.q2[src][i]))) {
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
            const int what = volumePre(sc, sc.prims, sc.media, sc.grids, nee, o, d, T, h, depth, rng, mi, ts, r, hasContrib, contrib);
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
constexpr int kVolPrimsSmem = 16; // shading records (64 B each) k_volume_paths stages in shared memory
constexpr int kMediaSmem = 8; // media / grids staged in shared memory by k_volume_paths (the host falls back to the wavefront form above that)
template <bool COUNT, int MINB>
__global__ void __launch_bounds__(kBlock, MINB) k_volume_paths(DScene sc, DQueues q, DWave w, int brute, int maxIter, int threshold, int stepsPerVote, unsigned long long* stats)
{
    __shared__ int s_stack[kStackSmem * kBlock];
    __shared__ DMedium s_media[kMediaSmem];
    __shared__ DGrid s_grids[kMediaSmem];
    __shared__ float4 s_prims[4 * kVolPrimsSmem]; // the shading records of a scene with few primitives (a box, a light, a few walls)
    if (int(threadIdx.x) < min(sc.nMedia, kMediaSmem)) s_media[threadIdx.x] = sc.media[threadIdx.x];
    if (int(threadIdx.x) < min(sc.nGrids, kMediaSmem)) s_grids[threadIdx.x] = sc.grids[threadIdx.x];
    const bool primsInSmem = sc.nPrims <= kVolPrimsSmem;
    if (primsInSmem)
        for (int k = threadIdx.x; k < 4 * sc.nPrims; k += blockDim.x) s_prims[k] = sc.prims[k];
    const float4* prims = primsInSmem ? s_prims : sc.prims;
    __syncthreads();
    const DMedium* media = s_media;
    const DGrid* grids = s_grids;
    uint32_t* ctrl = q.ctrl;
    const uint32_t n = ctrl[kCtrlRays];
    const bool nee = (w.integrator == XRTG_INT_VOLUME_NEE);
    const uint32_t lane = laneId();
    const int spv = stepsPerVote & 0xff, refillRounds = max(1, stepsPerVote >> 8); // (two small integers in one kernel parameter)
    uint32_t steps = 0, nClosest = 0, nTruncated = 0;
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
        // ---- lanes that just ended a walk: its epilogue, NEE, the closest hit of the next ray ----
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
            // maxIter bounds the medium crossings of one path (4 * maxDepth + 8): the reference's loop (integrator.h:418) has no
            // such bound, so a cut path is COUNTED (xrtg_stats.truncated_paths) instead of disappearing silently
            if (cont && it + 1 == maxIter) ++nTruncated;
            if (!cont || ++it == maxIter) finish();
            else {
                o = no; d = nd; T = nT;
                closestHit<COUNT>(sc, o, d, brute != 0, h, s_stack + threadIdx.x, tc);
                ++nClosest;
                state = kLanePre;
            }
        }
        // ---- refill + prologue rounds. A path usually ENDS in volumePre (the ray that left the medium misses, or Russian roulette) and
        // such a lane sits out the next walk. refillRounds > 1 hands it a fresh path from the queue before the warp enters the
        // lockstep loop — measured SLOWER on c5 (12.16 vs 12.95 Gsamples/s): the extra round runs refill + prologue at ~20 % of
        // the lanes, while waiting batches those lanes with everything else that ends during the walk. Default: one round ----
        bool allDone = false;
        for (int round = 0;; ++round) {
            // refill idle lanes (including those whose path just ended) from the warp's reservation of 32 queue entries
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
                if (state == kLaneIdle && i < n && !(XRT_WARP_APPEND_PRIMARY && deadEntry(q.q2[0][i]))) {
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
            if (__ballot_sync(0xffffffffu, state != kLaneIdle) == 0) { allDone = true; break; }
            // prologue of the next walk, for continuing and for fresh paths alike
            if (state == kLanePre) {
                V3 contrib;
                bool hasContrib;
                const int what = volumePre(sc, prims, media, grids, nee, o, d, T, h, depth, rng, mi, ts, r, hasContrib, contrib);
                if (hasContrib) add(contrib);
                if (what == kVolEnd) finish();
                else state = what == kVolTrack ? kLaneTrack : kLanePost;
            }
            if (round + 1 >= refillRounds || exhausted || __ballot_sync(0xffffffffu, state == kLaneIdle) == 0) break;
        }
        if (allDone) break;
        // ---- the walk: every tracking lane takes kTrackSteps steps per vote ----
        const uint32_t pending = __ballot_sync(0xffffffffu, state == kLanePre || state == kLanePost || state == kLaneWalked || (state == kLaneIdle && !exhausted));
        const uint32_t thr = pending ? uint32_t(threshold) : 1u;
        uint32_t busy = __popc(__ballot_sync(0xffffffffu, state == kLaneTrack));
        while (busy >= thr && busy > 0) {
#pragma unroll 1
            for (int k = 0; k < spv; ++k)
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
    statAdd(stats, kStatTruncated, nTruncated);
    if (COUNT) { statAdd(stats, kStatNodes, tc.nodes); statAdd(stats, kStatTris, tc.tris); }
}
