// wf_math_rng.cuh — part of wavefront.cuh (included inside namespace xrt::XRT_NS, in this order): vector algebra in the reference's operation order, the two RNGs (mt19937 restated / Philox4x32-7), seeding and jitter kernels.
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
        uint32_t s, lp;
        w.byWavePixels.divmod(pid, s, lp);
        pix = w.pixelBase + lp;
        if constexpr (kExact) {
            mt = w.mt; mti = w.mti; stride = w.nPixels;
            idx = mti[pix];
        }
        else {
            k0 = w.seed; k1 = pix;
            sample = w.sampleBase + s;
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
        w1.samplesThisWave = 1; w1.pixelBase = 0; w1.wavePixels = w.nPixels; w1.byWavePixels = makeFastDiv(w.nPixels);
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
