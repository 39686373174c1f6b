// wn_multiband_fast.cu -- throughput path for dense multiband 3D evaluation on a lattice
// (kernel K6 of DESIGN.md; BASELINE config 3: 1024^3 samples x 5 bands).
//
// Reference semantics: out = post * sum_b w_b * evaluate3D(p * s_b)   (WaveletNoise.cpp:185-215 applied per
// band; composition per Cook & DeRose App. 2).  This file may reorder the 27-tap sum (separable form) and use
// FMAs; tests bound the difference to the CPU reference by 1e-5 * (tile max - tile min).
//
// Design (k_mb3d_brick).  On an axis-aligned lattice the quadratic B-spline weights factor per axis, so the
// 27-tap gather is a tensor-product resampling  out = (Wz (x) Wy (x) Wx) N  with 3 non-zeros per row.
//   * k_axis_tables turns every axis coordinate of every band into {w0,w1,w2, first tap cell} once per launch
//     (un-fused arithmetic, so the tap cells are the reference's integers exactly).
//   * one CTA owns a brick of 32 x BY x BZ samples and loops over the bands.  Per band:
//       X pass : every tile row (cy,cz) the brick touches is contracted along x for the CTA's 32 x-samples
//                (lane = x sample; 3 read-only loads of the L2-resident tile, 3 FMA) into shared memory U[cz][cy][32].
//       YZ pass: thread (lane, j) walks its z column with a 3-deep sliding register window of y-contracted
//                values V[cz] = sum_f wy[f] U[cz][cy_j+f][lane]; each sample is 3 FMAs of the window with the
//                z weights (band weight and post scale folded in).  The window only advances when the next
//                sample's first tap cell advances, so low bands (many samples per cell) cost ~3 FMA per sample.
//     Shared-memory traffic is conflict-free (lane = fastest index), stores are 128 B per warp instruction.
//   * requirements: y and z tap cells non-decreasing along the axis and U fitting in shared memory; anything else
//     (unsorted axes, huge steps) runs k_mb3d_gather, the plain one-sample-per-thread form.
#include "wn_internal.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

namespace {

__device__ __forceinline__ int tmodf(int x, int n, int pow2)
{
    if (pow2) return x & (n - 1);
    int m = x % n;
    return m < 0 ? m + n : m;
}

// {w0, w1, w2, first tap cell (= mid - 1, NOT wrapped)} of one coordinate; un-fused like the reference (cpp:194-200)
__device__ __forceinline__ float4 axis_entry(float q, float wscale)
{
    const float a = __fsub_rn(q, 0.5f);
    const int mid = (int)ceilf(a);
    const float t = __fsub_rn((float)mid, a);
    const float w0 = __fmul_rn(__fmul_rn(t, t), 0.5f);
    const float s = __fsub_rn(1.0f, t);
    const float w2 = __fmul_rn(__fmul_rn(s, s), 0.5f);
    const float w1 = __fsub_rn(__fsub_rn(1.0f, w0), w2);
    return make_float4(w0 * wscale, w1 * wscale, w2 * wscale, __int_as_float(mid - 1));
}

// tables: tab[b * len + i] for the x, y axes and the launch's z slab
__global__ void k_axis_tables(WnLattice c, WnBands b, int k0, int nk, float4 *__restrict__ tx, float4 *__restrict__ ty,
                              float4 *__restrict__ tz)
{
    const int per_band = c.nx + c.ny + nk;
    const int total = per_band * b.nbands;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int band = e / per_band;
        int i = e - band * per_band;
        const float s = b.scale[band];
        if (i < c.nx) { tx[band * c.nx + i] = axis_entry(__fmul_rn(__ldg(c.xs + i), s), 1.0f); continue; }
        i -= c.nx;
        if (i < c.ny) { ty[band * c.ny + i] = axis_entry(__fmul_rn(__ldg(c.ys + i), s), 1.0f); continue; }
        i -= c.ny;
        tz[band * nk + i] = axis_entry(__fmul_rn(__ldg(c.zs + k0 + i), s), b.weight[band] * b.post);
    }
}

template <int BY, int BZ, int NT>
__global__ void __launch_bounds__(NT)
k_mb3d_brick(const float *__restrict__ N, int n, int pow2, const float4 *__restrict__ tabX,
             const float4 *__restrict__ tabY, const float4 *__restrict__ tabZ, int nx, int ny, int nk, int nbands,
             int max_rows, float *__restrict__ out)
{
    constexpr int NW = NT / 32;
    constexpr int C = BY / NW;                  // y columns per thread
    static_assert(BY % NW == 0, "BY must be a multiple of the warp count");
    extern __shared__ float U[];                // [Ez][Ey][32] followed by int rowoff[max_rows]
    int *s_rowoff = reinterpret_cast<int *>(U + (size_t)max_rows * 32);
    __shared__ float4 s_z[BZ];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + lane;
    const int j0 = blockIdx.y * BY, k0 = blockIdx.z * BZ;
    const int jl = min(j0 + BY, ny) - 1, kl = min(k0 + BZ, nk) - 1;     // last valid sample of the brick
    const int ic = min(i, nx - 1);

    float acc[C][BZ];
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
        for (int k = 0; k < BZ; ++k) acc[c][k] = 0.0f;

    for (int b = 0; b < nbands; ++b) {
        const float4 *tY = tabY + b * ny, *tZ = tabZ + b * nk;
        const float4 tx = __ldg(tabX + b * nx + ic);
        const int my0 = __float_as_int(__ldg(&tY[j0].w)), mz0 = __float_as_int(__ldg(&tZ[k0].w));
        const int Ey = __float_as_int(__ldg(&tY[jl].w)) - my0 + 3;
        const int Ez = __float_as_int(__ldg(&tZ[kl].w)) - mz0 + 3;
        const int rows = Ey * Ez;

        // row r = cz*Ey + cy of the brick's footprint -> offset of the (x-padded) tile row (wrapped y, wrapped z)
        for (int r = threadIdx.x; r < rows; r += NT) {
            const int cz = r / Ey, cy = r - cz * Ey;
            s_rowoff[r] = (tmodf(mz0 + cz, n, pow2) * n + tmodf(my0 + cy, n, pow2)) * (n + 2);
        }
        if (threadIdx.x < BZ) s_z[threadIdx.x] = __ldg(tZ + min(k0 + (int)threadIdx.x, nk - 1));
        __syncthreads();

        // ---- X pass: U[cz][cy][lane] = sum_f wx[f] * N[cz][cy][cx + f]
        {
            // rows of the padded tile hold cells 0..n+1, so the three taps are x0, x0+1, x0+2 of one address
            const int x0 = tmodf(__float_as_int(tx.w), n, pow2);
            float *u = U + warp * 32 + lane;
            int r = warp;
            for (; r + 3 * NW < rows; r += 4 * NW, u += 4 * NW * 32) {
                const float *q0 = N + (unsigned)(s_rowoff[r] + x0), *q1 = N + (unsigned)(s_rowoff[r + NW] + x0);
                const float *q2 = N + (unsigned)(s_rowoff[r + 2 * NW] + x0), *q3 = N + (unsigned)(s_rowoff[r + 3 * NW] + x0);
                const float a0 = __ldg(q0), a1 = __ldg(q0 + 1), a2 = __ldg(q0 + 2);
                const float b0 = __ldg(q1), b1 = __ldg(q1 + 1), b2 = __ldg(q1 + 2);
                const float c0 = __ldg(q2), c1 = __ldg(q2 + 1), c2 = __ldg(q2 + 2);
                const float d0 = __ldg(q3), d1 = __ldg(q3 + 1), d2 = __ldg(q3 + 2);
                u[0]           = fmaf(tx.z, a2, fmaf(tx.y, a1, tx.x * a0));
                u[NW * 32]     = fmaf(tx.z, b2, fmaf(tx.y, b1, tx.x * b0));
                u[2 * NW * 32] = fmaf(tx.z, c2, fmaf(tx.y, c1, tx.x * c0));
                u[3 * NW * 32] = fmaf(tx.z, d2, fmaf(tx.y, d1, tx.x * d0));
            }
            for (; r < rows; r += NW, u += NW * 32) {
                const float *q0 = N + (unsigned)(s_rowoff[r] + x0);
                u[0] = fmaf(tx.z, __ldg(q0 + 2), fmaf(tx.y, __ldg(q0 + 1), tx.x * __ldg(q0)));
            }
        }
        __syncthreads();

        // ---- YZ pass
        {
            float4 ty[C];
            const float *ucol[C];
            float v[C][3];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                ty[c] = __ldg(tY + min(j0 + warp + c * NW, ny - 1));
                ucol[c] = U + (__float_as_int(ty[c].w) - my0) * 32 + lane + 2 * Ey * 32;
                v[c][0] = v[c][1] = v[c][2] = 0.0f;
            }
            const int slab = Ey * 32;
            int base = -3;                       // window holds V[base .. base+2]
#pragma unroll
            for (int k = 0; k < BZ; ++k) {
                const float4 tz = s_z[k];
                const int rel = __float_as_int(tz.w) - mz0;
                while (base < rel) {
                    ++base;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const float *uu = ucol[c] + base * slab;
                        float nv = ty[c].x * uu[0];
                        nv = fmaf(ty[c].y, uu[32], nv);
                        nv = fmaf(ty[c].z, uu[64], nv);
                        v[c][0] = v[c][1]; v[c][1] = v[c][2]; v[c][2] = nv;
                    }
                }
#pragma unroll
                for (int c = 0; c < C; ++c)
                    acc[c][k] = fmaf(tz.x, v[c][0], fmaf(tz.y, v[c][1], fmaf(tz.z, v[c][2], acc[c][k])));
            }
        }
        __syncthreads();
    }

    if (i < nx) {
        const size_t plane = (size_t)nx * ny;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j = j0 + warp + c * NW;
            if (j >= ny) continue;
            float *o = out + ((size_t)i + (size_t)nx * j + plane * k0);
            if (k0 + BZ <= nk) {
#pragma unroll
                for (int k = 0; k < BZ; ++k) o[plane * k] = acc[c][k];
            } else {
#pragma unroll
                for (int k = 0; k < BZ; ++k)
                    if (k0 + k < nk) o[plane * k] = acc[c][k];
            }
        }
    }
}

__global__ void k_pad_tile(const float *__restrict__ N, float *__restrict__ P, int n)
{
    const int pitch = n + 2;
    const size_t total = (size_t)n * n * pitch;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const size_t row = e / pitch;
        int x = (int)(e - row * pitch);
        if (x >= n) x -= n;
        P[e] = N[row * n + x];
    }
}

// fallback: one sample per thread, separable contraction, taps through the read-only path
__global__ void __launch_bounds__(256)
k_mb3d_gather(WnTileView t, const float4 *__restrict__ tabX, const float4 *__restrict__ tabY,
              const float4 *__restrict__ tabZ, int nx, int ny, int nk, int nbands, float *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y, k = blockIdx.z;
    if (i >= nx) return;
    const int n = t.n;
    float acc = 0.0f;
    for (int b = 0; b < nbands; ++b) {
        const float4 ax = __ldg(tabX + (size_t)b * nx + i), ay = __ldg(tabY + (size_t)b * ny + j),
                     az = __ldg(tabZ + (size_t)b * nk + k);
        const float wx[3] = { ax.x, ax.y, ax.z }, wy[3] = { ay.x, ay.y, ay.z }, wz[3] = { az.x, az.y, az.z };
        int cx[3], cy[3], cz[3];
#pragma unroll
        for (int f = 0; f < 3; ++f) {
            cx[f] = tmodf(__float_as_int(ax.w) + f, n, t.pow2);
            cy[f] = tmodf(__float_as_int(ay.w) + f, n, t.pow2) * n;
            cz[f] = tmodf(__float_as_int(az.w) + f, n, t.pow2) * n * n;
        }
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
            acc = fmaf(wz[fz], vy, acc);
        }
    }
    out[(size_t)i + (size_t)nx * ((size_t)j + (size_t)ny * k)] = acc;
}

// host-side plan: first tap cells with exactly the device formula (this TU's host code is built with
// -ffp-contract=off), monotonicity of y/z and the largest Ey*Ez any brick needs.
inline int first_cell(float coord, float scale) { return (int)std::ceil(coord * scale - 0.5f) - 1; }

struct BrickPlan { bool ok; size_t smem; int max_rows; };

BrickPlan plan_bricks(const float *ys, int ny, const float *zs, int nk, const WnBands &b, int BY, int BZ)
{
    BrickPlan p{true, 0, 0};
    std::vector<int> my(ny), mz(nk);
    for (int band = 0; band < b.nbands; ++band) {
        for (int j = 0; j < ny; ++j) my[j] = first_cell(ys[j], b.scale[band]);
        for (int k = 0; k < nk; ++k) mz[k] = first_cell(zs[k], b.scale[band]);
        for (int j = 1; j < ny; ++j) if (my[j] < my[j - 1]) { p.ok = false; return p; }
        for (int k = 1; k < nk; ++k) if (mz[k] < mz[k - 1]) { p.ok = false; return p; }
        long long ey = 0, ez = 0;
        for (int j0 = 0; j0 < ny; j0 += BY) ey = std::max<long long>(ey, (long long)my[std::min(j0 + BY, ny) - 1] - my[j0] + 3);
        for (int k0 = 0; k0 < nk; k0 += BZ) ez = std::max<long long>(ez, (long long)mz[std::min(k0 + BZ, nk) - 1] - mz[k0] + 3);
        if (ey * ez > 1500) { p.ok = false; return p; }                 // 1500 rows * 132 B = 198 KB
        p.max_rows = std::max(p.max_rows, (int)(ey * ez));
    }
    p.smem = (size_t)p.max_rows * (32 * sizeof(float) + sizeof(int));
    return p;
}

template <int BY, int BZ, int NT>
int launch_brick(WnTileView t, const float4 *tx, const float4 *ty, const float4 *tz, int nx, int ny, int nk, int nbands,
                 float *out, const BrickPlan &plan, cudaStream_t st)
{
    auto kern = k_mb3d_brick<BY, BZ, NT>;
    const size_t smem = plan.smem;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    dim3 grid((nx + 31) / 32, (ny + BY - 1) / BY, (nk + BZ - 1) / BZ);
    kern<<<grid, NT, smem, st>>>(t.Npad, t.n, t.pow2, tx, ty, tz, nx, ny, nk, nbands, plan.max_rows, out);
    return 1;
}

} // namespace

int wn_launch_pad_tile(const float *N, float *Npad, int n, cudaStream_t st)
{
    k_pad_tile<<<148 * 8, 256, 0, st>>>(N, Npad, n);
    return 1;
}

int wn_launch_mb3d_lattice_fast(WnTileView t, WnLattice c, const float *h_ys, const float *h_zs, WnBands b, int k0, int nk,
                                float *out, cudaStream_t st)
{
    if (nk <= 0 || c.nx <= 0 || c.ny <= 0) return 0;
    int launches = 0;
    const size_t per_band = (size_t)c.nx + c.ny + nk;
    float4 *tab = nullptr;
    if (cudaMallocAsync(&tab, per_band * b.nbands * sizeof(float4), st) != cudaSuccess) return -1;
    float4 *tx = tab, *ty = tab + (size_t)b.nbands * c.nx, *tz = ty + (size_t)b.nbands * c.ny;
    {
        const int total = (int)(per_band * b.nbands);
        k_axis_tables<<<std::min((total + 255) / 256, 1184), 256, 0, st>>>(c, b, k0, nk, tx, ty, tz);
        ++launches;
    }
    // brick shapes (BY, BZ, threads); WN_BRICK=<index> overrides the default for tuning runs
    struct Shape { int by, bz; };
    static const Shape shapes[] = { {16, 8}, {16, 16}, {8, 8}, {8, 16}, {32, 8}, {16, 4} };
    int pick = 0;
    if (const char *e = getenv("WN_BRICK")) pick = atoi(e);
    if (pick < 0 || pick >= (int)(sizeof(shapes) / sizeof(shapes[0]))) pick = 0;
    const int BY = shapes[pick].by, BZ = shapes[pick].bz;
    const BrickPlan plan = plan_bricks(h_ys, c.ny, h_zs + k0, nk, b, BY, BZ);
    const bool grid_ok = (c.ny + BY - 1) / BY <= 65535 && (nk + BZ - 1) / BZ <= 65535;
    if (plan.ok && grid_ok && plan.smem <= 200 * 1024) {
        int r = -1;
        switch (pick) {
        case 0: r = launch_brick<16, 8, 256>(t, tx, ty, tz, c.nx, c.ny, nk, b.nbands, out, plan, st); break;
        case 1: r = launch_brick<16, 16, 256>(t, tx, ty, tz, c.nx, c.ny, nk, b.nbands, out, plan, st); break;
        case 2: r = launch_brick<8, 8, 256>(t, tx, ty, tz, c.nx, c.ny, nk, b.nbands, out, plan, st); break;
        case 3: r = launch_brick<8, 16, 256>(t, tx, ty, tz, c.nx, c.ny, nk, b.nbands, out, plan, st); break;
        case 4: r = launch_brick<32, 8, 256>(t, tx, ty, tz, c.nx, c.ny, nk, b.nbands, out, plan, st); break;
        case 5: r = launch_brick<16, 4, 256>(t, tx, ty, tz, c.nx, c.ny, nk, b.nbands, out, plan, st); break;
        }
        if (r < 0) { cudaFreeAsync(tab, st); return -1; }
        launches += r;
    } else {
        // gridDim.y/z are limited to 65535: walk y and z in slabs
        for (int kk = 0; kk < nk; kk += 65535)
            for (int jj = 0; jj < c.ny; jj += 65535) {
                const int cz = std::min(nk - kk, 65535), cy = std::min(c.ny - jj, 65535);
                if (cy != c.ny || cz != nk) { cudaFreeAsync(tab, st); return -1; }    // lattices that large are rejected
                dim3 grid((c.nx + 255) / 256, cy, cz);
                k_mb3d_gather<<<grid, 256, 0, st>>>(t, tx, ty, tz, c.nx, c.ny, nk, b.nbands, out);
                ++launches;
            }
    }
    cudaFreeAsync(tab, st);
    return launches;
}
