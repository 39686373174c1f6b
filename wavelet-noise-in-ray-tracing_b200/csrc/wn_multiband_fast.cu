// wn_multiband_fast.cu -- throughput path for dense multiband 3D evaluation on a lattice
// (kernel K6 of DESIGN.md; BASELINE config 3: 1024^3 samples x 5 bands).
//
// Reference semantics: out = post * sum_b w_b * evaluate3D(p * s_b)   (WaveletNoise.cpp:185-215 applied per
// band; composition per Cook & DeRose App. 2).  This file may reorder the 27-tap sum (separable form) and use
// FMAs; tests bound the difference to the CPU reference by 1e-5 * (tile max - tile min).
#include "wn_internal.h"

namespace {

__device__ __forceinline__ int tmodf(int x, int n, int pow2)
{
    if (pow2) return x & (n - 1);
    int m = x % n;
    return m < 0 ? m + n : m;
}

__device__ __forceinline__ void basisf(float p, int &mid, float &w0, float &w1, float &w2)
{
    const float a = p - 0.5f;
    mid = (int)ceilf(a);
    const float t = (float)mid - a;
    w0 = t * t * 0.5f;
    const float s = 1.0f - t;
    w2 = s * s * 0.5f;
    w1 = 1.0f - w0 - w2;
}

// v1: one sample per thread, separable contraction, taps through the read-only path (tile is L2/L1 resident).
__global__ void __launch_bounds__(256)
k_mb3d_lattice_v1(WnTileView t, WnLattice c, WnBands b, int k0, int nk, float *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    const int k = blockIdx.z;
    if (i >= c.nx || k >= nk) return;
    const float x = __ldg(c.xs + i), y = __ldg(c.ys + j), z = __ldg(c.zs + k0 + k);
    const int n = t.n;
    float acc = 0.0f;
    for (int bi = 0; bi < b.nbands; ++bi) {
        const float s = b.scale[bi];
        int mx, my, mz; float wx[3], wy[3], wz[3];
        basisf(x * s, mx, wx[0], wx[1], wx[2]);
        basisf(y * s, my, wy[0], wy[1], wy[2]);
        basisf(z * s, mz, wz[0], wz[1], wz[2]);
        int cx[3], cy[3], cz[3];
#pragma unroll
        for (int f = 0; f < 3; ++f) {
            cx[f] = tmodf(mx + f - 1, n, t.pow2);
            cy[f] = tmodf(my + f - 1, n, t.pow2) * n;
            cz[f] = tmodf(mz + f - 1, n, t.pow2) * n * n;
        }
        float vz = 0.0f;
#pragma unroll
        for (int fz = 0; fz < 3; ++fz) {
            float vy = 0.0f;
#pragma unroll
            for (int fy = 0; fy < 3; ++fy) {
                const float *row = t.N + cy[fy] + cz[fz];
                float vx = wx[0] * __ldg(row + cx[0]);
                vx = fmaf(wx[1], __ldg(row + cx[1]), vx);
                vx = fmaf(wx[2], __ldg(row + cx[2]), vx);
                vy = fmaf(wy[fy], vx, vy);
            }
            vz = fmaf(wz[fz], vy, vz);
        }
        acc = fmaf(b.weight[bi], vz, acc);
    }
    out[(size_t)i + (size_t)c.nx * ((size_t)j + (size_t)c.ny * k)] = acc * b.post;
}

} // namespace

int wn_launch_mb3d_lattice_fast(WnTileView t, WnLattice c, WnBands b, int k0, int nk, float *out, cudaStream_t st)
{
    if (nk <= 0 || c.nx <= 0 || c.ny <= 0) return 0;
    int launches = 0;
    // gridDim.y/z are limited to 65535: walk z in slabs if needed
    for (int kk = 0; kk < nk; kk += 65535) {
        const int cnt = (nk - kk) < 65535 ? (nk - kk) : 65535;
        for (int jj = 0; jj < c.ny; jj += 65535) {
            WnLattice cc = c;
            cc.ys = c.ys + jj;
            const int cy = (c.ny - jj) < 65535 ? (c.ny - jj) : 65535;
            dim3 grid((c.nx + 255) / 256, cy, cnt);
            // note: out index uses c.ny as row pitch, so offset the pointer for the y sub-range
            k_mb3d_lattice_v1<<<grid, 256, 0, st>>>(t, WnLattice{cc.xs, cc.ys, c.zs, c.nx, c.ny, c.nz}, b, k0 + kk, cnt,
                                                    out + (size_t)c.nx * ((size_t)jj + (size_t)c.ny * kk));
            ++launches;
        }
    }
    return launches;
}
