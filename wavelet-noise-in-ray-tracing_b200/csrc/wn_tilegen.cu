// wn_tilegen.cu -- wavelet-noise tile construction on the GPU (kernels K1..K4 of DESIGN.md).
//
// What it computes (reference: WaveletNoise.cpp:37-66 filters, :87-107 2D sweeps, :153-182 3D sweeps):
//   per axis, every line of the tile is down-sampled with the 32-tap analysis filter (stride 2,
//   periodic) and up-sampled again with the 4-tap refinement filter; after the last axis the
//   result is subtracted from the Gaussian field R, giving the band-limited tile N.
//
// Parity: all arithmetic uses __fmul_rn/__fadd_rn in the reference's summation order
// (k = -16..15 starting from 0.0f; then the two refinement taps starting from 0.0f), so for the
// same R the tile is BIT-IDENTICAL to the CPU reference.  (This file is also built with -fmad=false.)
//
// Mapping to the machine:
//   x axis   : lines are contiguous.  One warp owns one line at a time: coalesced 128 B loads into a
//              warp-private shared-memory row, coarse values into a second row, coalesced stores.
//   y/z axis : lines are strided by n / n*n, but neighbouring lines are adjacent in memory.  A CTA owns a
//              slab of W (=32) neighbouring lines x the full axis: every global access is a coalesced
//              row of W floats, shared-memory rows are W wide so the column filter is conflict-free.
//   The last sweep fuses the subtraction N = R - (...) into its store.
#include "wn_internal.h"

namespace {

// Cook & DeRose 2005, Appendix 1 analysis filter as used by the reference (WaveletNoise.cpp:11-16);
// entry [16] is tap k = 0.  Entry 28 really is 0.003546 (not 0.003545) in the reference.
__constant__ float c_A[32] = {
    0.000334f, -0.001528f,  0.000410f,  0.003545f, -0.000938f, -0.008233f,  0.002172f,  0.019120f,
   -0.005040f, -0.044412f,  0.011655f,  0.103311f, -0.025936f, -0.243780f,  0.033979f,  0.655340f,
    0.655340f,  0.033979f, -0.243780f, -0.025936f,  0.103311f,  0.011655f, -0.044412f, -0.005040f,
    0.019120f,  0.002172f, -0.008233f, -0.000938f,  0.003546f,  0.000410f, -0.001528f,  0.000334f
};

__device__ __forceinline__ int wrap_any(int i, int n)     // Mod(), WaveletNoise.cpp:31-34
{
    int m = i % n;
    return m < 0 ? m + n : m;
}

// coarse[i] of one line.  `line` has element stride `ls` in shared memory.
__device__ __forceinline__ float down_one(const float *line, int ls, int i, int n)
{
    float sum = 0.0f;
    if (n >= 32) {                       // 2i+k lies in [-16, n+14]: one conditional wrap is enough
#pragma unroll
        for (int k = -16; k < 16; ++k) {
            int idx = 2 * i + k;
            idx = idx < 0 ? idx + n : (idx >= n ? idx - n : idx);
            sum = __fadd_rn(sum, __fmul_rn(c_A[16 + k], line[idx * ls]));
        }
    } else {
#pragma unroll
        for (int k = -16; k < 16; ++k)
            sum = __fadd_rn(sum, __fmul_rn(c_A[16 + k], line[wrap_any(2 * i + k, n) * ls]));
    }
    return sum;
}

// refined[i] from the coarse line (WaveletNoise.cpp:51-66): taps k = i/2 then i/2+1,
// weights P[2+(i-2k)] = {0.75, 0.25} for even i and {0.25, 0.75} for odd i.
__device__ __forceinline__ float up_one(const float *coarse, int ls, int i, int half)
{
    const int k0 = i >> 1;
    int k1 = k0 + 1;
    if (k1 >= half) k1 -= half;
    const bool odd = i & 1;
    float sum = __fadd_rn(0.0f, __fmul_rn(odd ? 0.25f : 0.75f, coarse[k0 * ls]));
    return __fadd_rn(sum, __fmul_rn(odd ? 0.75f : 0.25f, coarse[k1 * ls]));
}

// ---- x axis: one warp per line ---------------------------------------------------------------
__global__ void k_filter_x(const float *__restrict__ src, float *__restrict__ dst,
                           const float *__restrict__ minuend, int n, int nlines)
{
    extern __shared__ float smem[];
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int half = n >> 1;
    float *row = smem + (size_t)warp * (n + half);
    float *coarse = row + n;
    for (int line = blockIdx.x * warps + warp; line < nlines; line += gridDim.x * warps) {
        const size_t base = (size_t)line * n;
        for (int i = lane; i < n; i += 32) row[i] = src[base + i];
        __syncwarp();
        for (int i = lane; i < half; i += 32) coarse[i] = down_one(row, 1, i, n);
        __syncwarp();
        for (int i = lane; i < n; i += 32) {
            float v = up_one(coarse, 1, i, half);
            if (minuend) v = __fsub_rn(minuend[base + i], v);
            dst[base + i] = v;
        }
        __syncwarp();
    }
}

// ---- y / z axis: a CTA owns W neighbouring lines --------------------------------------------------
// Line l (0 <= l < nlines) starts at  (l / inner) * outer_stride + (l % inner)  and has element stride
// `stride`:   y axis of a 3D tile: inner = n, outer_stride = n*n, stride = n     (l = z*n + x)
//             z axis of a 3D tile: inner = n*n, outer_stride = 0, stride = n*n   (l = y*n + x)
//             y axis of a 2D tile: inner = n, outer_stride = 0, stride = n       (l = x)
__global__ void k_filter_strided(const float *__restrict__ src, float *__restrict__ dst,
                                 const float *__restrict__ minuend, int n, int nlines,
                                 int inner, size_t outer_stride, size_t stride)
{
    extern __shared__ float smem[];
    const int W = blockDim.x, rows = blockDim.y;
    const int lx = threadIdx.x, ly = threadIdx.y;
    const int half = n >> 1;
    float *slab = smem;                        // [n][W]
    float *coarse = smem + (size_t)n * W;      // [half][W]
    const int l = blockIdx.x * W + lx;
    const bool live = l < nlines;
    const size_t base = live ? (size_t)(l / inner) * outer_stride + (size_t)(l % inner) : 0;
    if (live)
        for (int i = ly; i < n; i += rows) slab[i * W + lx] = src[base + (size_t)i * stride];
    __syncthreads();
    if (live)
        for (int i = ly; i < half; i += rows) coarse[i * W + lx] = down_one(slab + lx, W, i, n);
    __syncthreads();
    if (live)
        for (int i = ly; i < n; i += rows) {
            float v = up_one(coarse + lx, W, i, half);
            const size_t g = base + (size_t)i * stride;
            if (minuend) v = __fsub_rn(minuend[g], v);
            dst[g] = v;
        }
}

// ---- paper odd-offset step (NOT in the reference; opt-in, WN_TILE_ODD_OFFSET) -------------------------
// Cook & DeRose App. 1: offset = n/2 made odd; temp[ix*n*n + iy*n + iz] = noise[Mod(ix+o) + Mod(iy+o)*n +
// Mod(iz+o)*n*n]; noise[i] += temp[i].  Output index i = a*n*n + b*n + c reads the shifted cell (a,b,c) as (x,y,z).
__global__ void k_odd_offset3d(const float *__restrict__ src, float *__restrict__ dst, int n, int off)
{
    const size_t cnt = (size_t)n * n * n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < cnt; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % n), b = (int)((i / n) % n), a = (int)(i / ((size_t)n * n));
        const size_t g = (size_t)wrap_any(a + off, n) + (size_t)wrap_any(b + off, n) * n +
                         (size_t)wrap_any(c + off, n) * n * n;
        dst[i] = __fadd_rn(src[i], src[g]);
    }
}

} // namespace

int wn_launch_filter_axis(const float *src, float *dst, const float *minuend, int n, int dims, int axis,
                          cudaStream_t st)
{
    const int half = n / 2;
    const size_t nlines = (dims == 3) ? (size_t)n * n : (size_t)n;
    if (axis == 0) {
        int warps = 8;
        size_t per_warp = (size_t)(n + half) * sizeof(float);
        while (warps > 1 && per_warp * warps > 200 * 1024) warps >>= 1;
        size_t smem = per_warp * warps;
        if (smem > 227 * 1024) return -1;
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(k_filter_x, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        size_t blocks = (nlines + warps - 1) / warps;
        if (blocks > 148 * 16) blocks = 148 * 16;
        k_filter_x<<<(unsigned)blocks, warps * 32, smem, st>>>(src, dst, minuend, n, (int)nlines);
    } else {
        int W = 32;
        while (W > 1 && (size_t)(n + half) * W * sizeof(float) > 200 * 1024) W >>= 1;
        size_t smem = (size_t)(n + half) * W * sizeof(float);
        if (smem > 227 * 1024) return -1;
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(k_filter_strided, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int inner; size_t outer_stride, stride;
        if (dims == 3 && axis == 1)      { inner = n;     outer_stride = (size_t)n * n; stride = (size_t)n; }
        else if (dims == 3 && axis == 2) { inner = n * n; outer_stride = 0;             stride = (size_t)n * n; }
        else                             { inner = n;     outer_stride = 0;             stride = (size_t)n; }
        dim3 block(W, 256 / W > 32 ? 32 : 256 / W);
        unsigned blocks = (unsigned)((nlines + W - 1) / W);
        k_filter_strided<<<blocks, block, smem, st>>>(src, dst, minuend, n, (int)nlines, inner, outer_stride, stride);
    }
    return 1;
}

int wn_launch_odd_offset3d(const float *src, float *dst, int n, cudaStream_t st)
{
    int off = n / 2;
    if (off % 2 == 0) off++;
    k_odd_offset3d<<<148 * 8, 256, 0, st>>>(src, dst, n, off);
    return 1;
}
