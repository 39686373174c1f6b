// peaks.cu -- micro-benchmarks for the roofline denominators MEASURED_PEAKS.json does not hold:
//   fp32_fma      : dependent-chain FFMA throughput, 8 chains per thread (TFLOP/s)
//   fp32_fma2     : the same with packed FFMA2 (sm_100)
//   l2_read_8MiB  : coalesced LDG.128 streaming of an L2-resident 8 MiB buffer (the tile) by every SM (GB/s)
//   l2_read_64MiB : the same over 64 MiB (the 256^3 period block)
//   hbm_write     : st.global.cs float4 stream over 4 GiB (what the headline kernel's output stream can reach) (GB/s)
//   hbm_copy      : float4 copy 2 GiB -> 2 GiB (read + write bytes), the MEASURED_PEAKS.json method with a plain kernel
//   d2h_pinned    : cudaMemcpyAsync device -> pinned host, 1 GiB (GB/s): ceiling of the e2e arm
// Prints one JSON object.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o bin/peaks peaks.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void k_fma(float *out, int iters, float a, float b)
{
    float c0 = threadIdx.x, c1 = c0 + 1, c2 = c0 + 2, c3 = c0 + 3, c4 = c0 + 4, c5 = c0 + 5, c6 = c0 + 6, c7 = c0 + 7;
    for (int i = 0; i < iters; ++i) {
        c0 = fmaf(c0, a, b); c1 = fmaf(c1, a, b); c2 = fmaf(c2, a, b); c3 = fmaf(c3, a, b);
        c4 = fmaf(c4, a, b); c5 = fmaf(c5, a, b); c6 = fmaf(c6, a, b); c7 = fmaf(c7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7;
}

__global__ void k_fma2(float *out, int iters, float a, float b)
{
    float2 c0 = make_float2(threadIdx.x, 1), c1 = make_float2(threadIdx.x, 2), c2 = make_float2(threadIdx.x, 3),
           c3 = make_float2(threadIdx.x, 4), c4 = make_float2(threadIdx.x, 5), c5 = make_float2(threadIdx.x, 6),
           c6 = make_float2(threadIdx.x, 7), c7 = make_float2(threadIdx.x, 8);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int i = 0; i < iters; ++i) {
        c0 = __ffma2_rn(c0, a2, b2); c1 = __ffma2_rn(c1, a2, b2); c2 = __ffma2_rn(c2, a2, b2); c3 = __ffma2_rn(c3, a2, b2);
        c4 = __ffma2_rn(c4, a2, b2); c5 = __ffma2_rn(c5, a2, b2); c6 = __ffma2_rn(c6, a2, b2); c7 = __ffma2_rn(c7, a2, b2);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0.x + c1.x + c2.x + c3.x + c4.x + c5.x + c6.x + c7.x + c0.y + c1.y + c2.y +
                                                 c3.y + c4.y + c5.y + c6.y + c7.y;
}

// every CTA streams the whole buffer `passes` times (L2-resident when it is small)
__global__ void k_l2_read(const float4 *buf, size_t n4, int passes, float *out)
{
    float acc = 0.0f;
    for (int p = 0; p < passes; ++p) {
        // CTAs start at different offsets so they do not all hit the same L2 slices at the same time
        const size_t start = ((size_t)blockIdx.x * 7919u * blockDim.x) % n4;
        for (size_t i = threadIdx.x; i < n4; i += blockDim.x) {
            size_t j = start + i;
            if (j >= n4) j -= n4;
            const float4 v = __ldcg(buf + j);
            acc += v.x + v.y + v.z + v.w;
        }
    }
    if (acc == 123.456f) out[0] = acc;
}

__global__ void k_write(float4 *dst, size_t n4)
{
    const float4 v = make_float4(1.0f, 2.0f, 3.0f, 4.0f);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) __stcs(dst + i, v);
}

__global__ void k_copy(const float4 *src, float4 *dst, size_t n4)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
        __stcs(dst + i, __ldcs(src + i));
}

template <class F>
static float best_ms(F f, int reps)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(a));
        f();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    float *out; CK(cudaMalloc(&out, (size_t)sms * 8 * 256 * sizeof(float)));
    const int iters = 1 << 16;
    const double flop = (double)sms * 8 * 256 * 8 * 2.0 * iters;
    const float ms_fma = best_ms([&] { k_fma<<<sms * 8, 256>>>(out, iters, 1.0001f, 0.5f); }, 5);
    const float ms_fma2 = best_ms([&] { k_fma2<<<sms * 8, 256>>>(out, iters, 1.0001f, 0.5f); }, 5);
    float4 *big; const size_t big_bytes = (size_t)4 << 30;
    CK(cudaMalloc(&big, big_bytes));
    CK(cudaMemset(big, 0, big_bytes));
    double l2[2];
    const size_t l2_sizes[2] = { (size_t)8 << 20, (size_t)64 << 20 };
    for (int s = 0; s < 2; ++s) {
        const size_t n4 = l2_sizes[s] / 16;
        const int passes = s == 0 ? 16 : 2;
        const float ms = best_ms([&] { k_l2_read<<<sms * 2, 1024>>>(big, n4, passes, out); }, 5);
        l2[s] = (double)l2_sizes[s] * passes * sms * 2 / (ms * 1e-3) / 1e9;
    }
    const float ms_w = best_ms([&] { k_write<<<sms * 16, 512>>>(big, big_bytes / 16); }, 10);
    const float ms_c = best_ms([&] { k_copy<<<sms * 16, 512>>>(big, big + big_bytes / 32, big_bytes / 32); }, 10);
    void *host; const size_t hbytes = (size_t)1 << 30;
    CK(cudaMallocHost(&host, hbytes));
    const float ms_d2h = best_ms([&] { CK(cudaMemcpyAsync(host, big, hbytes, cudaMemcpyDeviceToHost, 0)); }, 5);
    const float ms_h2d = best_ms([&] { CK(cudaMemcpyAsync(big, host, hbytes, cudaMemcpyHostToDevice, 0)); }, 5);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"fp32_fma_tflops\": %.2f, \"fp32_fma2_tflops\": %.2f, \"l2_read_8MiB_gbs\": %.1f, "
           "\"l2_read_64MiB_gbs\": %.1f, \"hbm_write_gbs\": %.1f, \"hbm_copy_gbs\": %.1f, \"d2h_pinned_gbs\": %.2f, "
           "\"h2d_pinned_gbs\": %.2f}\n",
           prop.name, sms, flop / (ms_fma * 1e-3) / 1e12, 2.0 * flop / (ms_fma2 * 1e-3) / 1e12, l2[0], l2[1],
           (double)big_bytes / (ms_w * 1e-3) / 1e9, (double)big_bytes / (ms_c * 1e-3) / 1e9,
           (double)hbytes / (ms_d2h * 1e-3) / 1e9, (double)hbytes / (ms_h2d * 1e-3) / 1e9);
    return 0;
}
