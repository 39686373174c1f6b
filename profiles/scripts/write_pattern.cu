// write_pattern.cu -- how fast can HBM take the headline kernel's OUTPUT PATTERN with no arithmetic at all?
// k_mb3d_rep<2,1,2,2> on config 3: 8192 CTAs of 256 threads, two per SM (108 KB of shared memory each); a CTA owns a
// 128 x 8 x 32 brick of each of the four x/y replicas of a 1024^3 float volume, and every z step a warp writes one
// 512-byte row piece per replica with st.global.cs.v4.  Variants: the same pattern at 2 / 4 / 8 CTAs per SM, and one
// contiguous 4 GiB stream for reference.  Prints one JSON object.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o bin/write_pattern write_pattern.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void __launch_bounds__(256) k_brick_writes(float *out)
{
    extern __shared__ float4 pad[];
    const int b = blockIdx.x, bx = b & 3, yb = (b >> 2) & 63, zb = b >> 8;       // 4 x 64 x 32 bricks
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t plane = (size_t)1024 * 1024;
    float *o = out + (size_t)(bx * 128 + 4 * lane) + (size_t)1024 * (yb * 8 + warp) + plane * (zb * 32);
    float4 v = make_float4((float)b, (float)lane, 1.0f, 2.0f);
    if (out == nullptr) pad[threadIdx.x] = v;                                    // keep the allocation alive
    for (int k = 0; k < 32; ++k) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
            __stcs(reinterpret_cast<float4 *>(o + (r & 1) * 512 + (size_t)(r >> 1) * 512 * 1024), v);
        o += plane;
        v.x += 1.0f;
    }
}

__global__ void k_stream(float4 *dst, size_t n4)
{
    const float4 v = make_float4(1.0f, 2.0f, 3.0f, 4.0f);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) __stcs(dst + i, v);
}

template <class F> static float best_ms(F f, int reps)
{
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    float *vol; const size_t bytes = (size_t)4 << 30;
    CK(cudaMalloc(&vol, bytes)); CK(cudaMemset(vol, 0, bytes));
    CK(cudaFuncSetAttribute(k_brick_writes, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    const int smem[3] = { 108 * 1024, 54 * 1024, 24 * 1024 };                     // 2, 4, 8 CTAs per SM
    double gbs[3];
    for (int s = 0; s < 3; ++s) {
        const float ms = best_ms([&] { k_brick_writes<<<8192, 256, smem[s]>>>(vol); }, 10);
        gbs[s] = (double)bytes / (ms * 1e-3) / 1e9;
    }
    const float ms_s = best_ms([&] { k_stream<<<prop.multiProcessorCount * 16, 512>>>((float4 *)vol, bytes / 16); }, 10);
    printf("{\"gpu\": \"%s\", \"brick_pattern_2cta_gbs\": %.1f, \"brick_pattern_4cta_gbs\": %.1f, \"brick_pattern_8cta_gbs\": %.1f, "
           "\"contiguous_stream_gbs\": %.1f}\n", prop.name, gbs[0], gbs[1], gbs[2], (double)bytes / (ms_s * 1e-3) / 1e9);
    return 0;
}
