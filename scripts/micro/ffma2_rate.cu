// Microbenchmark: issue rate of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu ; run: ./ffma2_rate
#include <cstdio>
#include <cuda_runtime.h>
constexpr int kIters = 4096, kChains = 8;
__global__ void k_scalar(float* out, float a, float b)
{
    float acc[2 * kChains];
    for (int i = 0; i < 2 * kChains; ++i) acc[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < kIters; ++it)
#pragma unroll
        for (int i = 0; i < 2 * kChains; ++i) acc[i] = fmaf(acc[i], a, b);
    float s = 0;
    for (int i = 0; i < 2 * kChains; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed(float* out, float a, float b)
{
    float2 acc[kChains];
    for (int i = 0; i < kChains; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
    const float2 A = make_float2(a, a), B = make_float2(b, b);
    for (int it = 0; it < kIters; ++it)
#pragma unroll
        for (int i = 0; i < kChains; ++i) acc[i] = __ffma2_rn(acc[i], A, B);
    float s = 0;
    for (int i = 0; i < kChains; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    float* out; cudaMalloc(&out, sizeof(float) * 148 * 8 * 256);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
        float ms;
        cudaEventRecord(e0); k_scalar<<<148 * 8, 256>>>(out, 1.0001f, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        const double fma = double(148) * 8 * 256 * kIters * 2 * kChains;
        printf("scalar FFMA : %.3f ms  %.1f TFMA/s (%.1f TFLOP/s)\n", ms, fma / ms / 1e9, 2 * fma / ms / 1e9);
        cudaEventRecord(e0); k_packed<<<148 * 8, 256>>>(out, 1.0001f, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("packed FFMA2: %.3f ms  %.1f TFMA/s (%.1f TFLOP/s)\n", ms, fma / ms / 1e9, 2 * fma / ms / 1e9);
    }
    return 0;
}
