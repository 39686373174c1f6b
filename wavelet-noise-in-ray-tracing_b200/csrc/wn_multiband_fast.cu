// wn_multiband_fast.cu -- throughput path for dense multiband 3D evaluation on a lattice
// (kernel K6 of DESIGN.md; BASELINE config 3: 1024^3 samples x 5 bands).
//
// Reference semantics: out = post * sum_b w_b * evaluate3D(p * s_b)   (WaveletNoise.cpp:185-215 applied per
// band; composition per Cook & DeRose App. 2).  This file may reorder the 27-tap sum (separable form) and use
// FMAs; tests bound the difference to the CPU reference by 1e-5 * (tile max - tile min).
//
// Design.  On an axis-aligned lattice the quadratic B-spline weights factor per axis, so the 27-tap gather is a
// tensor-product resampling  out = (Wz (x) Wy (x) Wx) N  with 3 non-zeros per row.
//   * k_axis_tables turns every axis coordinate of every band into {w0,w1,w2, first tap cell} once per call
//     (un-fused arithmetic, so the tap cells are the reference's integers exactly); the host builds the same table
//     (HostAxes) for period detection and shared-memory planning.
//   * Periodic folding (wn_mb3d_fast_prepare): a band whose table entries repeat with period P along every axis is
//     evaluated once on its period block and added by index mod period; blocks nest.  See the comment there.
//   * Brick kernels.  A CTA owns a brick of samples and works per band in two passes:
//       X pass : every tile row (cy,cz) the brick touches is contracted along x for the CTA's x-samples (read-only
//                loads of the L2-resident, x-padded tile) into shared memory U[cz][cy][x];
//       YZ pass: a thread walks its z column with a 3-deep sliding register window of y-contracted values
//                V[cz] = sum_f wy[f] U[cz][cy+f][x]; each sample is 3 FMAs of the window with the z weights (band weight
//                and post scale folded in).  The window only advances when the next sample's first tap cell advances,
//                so low bands (many samples per cell) cost ~3 FMA per sample.
//     k_mb3d_col4  : 128 x 8 x 32 samples, four x-samples per lane, band loop INSIDE the z loop (one or two bands
//                    left after folding); the period-block column streams through a cp.async ring.  Main kernel.
//     k_mb3d_brick4: 128 x BY x BZ samples, four x-samples per lane, all bands resident, BZ accumulators per thread.
//     k_mb3d_brick : 32 x BY x BZ samples, one x-sample per lane, per-band phases (large footprints).
//     k_mb3d_gather: one sample per thread, for lattices that do not qualify (unsorted y/z axes, huge steps).
//     All four produce bit-identical samples (same per-sample operation order).
//   * The launches of a call are chained with programmatic dependent launch (chain_wait / chain_release).
#include "wn_internal.h"

#include <cuda.h>                 // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

namespace {

// sum of the periodic ("folded") bands on their common period block, added by index mod period; P == nullptr: none
struct WnFold {
    const float *P;
    int Lx, Ly, Lz;
    int kphase;                 // (first z index of this launch) mod Lz
    int xmask, ymask;           // L-1 when L is a power of two, else -1 (use %)
};

// Axis tables of one top-level call: row r of x/y/z holds {w0, w1, w2, first tap cell} of every coordinate of the FULL
// axis for original band r (k_axis_tables, one launch per call).  A kernel launch works on a subset of the bands
// (row[]) and on a window of the lattice: the x/y windows always start at 0 (period blocks are prefixes of the
// axes), the z window starts at kz0 (chunks of a WN_HOST call).
struct WnTabs {
    const float4 *x, *y, *z;
    int sx, sy, sz;             // row strides = full axis lengths
    int kz0;                    // first z entry of this launch
    int nb;                     // bands of this launch
    unsigned long long rowbits;  // table row of band b in bits [4b, 4b+4): no indexed kernel parameter, no local copy
    int wait_first;             // first kernel after k_axis_tables: must wait for its predecessor before reading the tables
    int pdl;                    // host side only: launch with the programmatic stream serialisation attribute
};
static_assert(WN_MAX_BANDS <= 16, "rows are packed four bits each");
__host__ __device__ __forceinline__ int tabs_row(const WnTabs &t, int b) { return (int)((t.rowbits >> (4 * b)) & 15ull); }

// Programmatic dependent launch.  The kernels of a call are a chain of dependent launches, several of them tiny
// (period blocks of a few CTAs), and back-to-back calls continue the chain.  Each launch carries the programmatic
// stream serialisation attribute, so its CTAs may start while the previous kernel drains and run the part that does
// not depend on it (tables from two kernels back, footprints, the X pass from the constant tile) up to
// chain_wait(), which returns once the previous kernel has completed and its writes are visible.  Every kernel
// releases its dependents only AFTER its own wait, so "my predecessor's predecessor is complete" holds at every
// start, and no kernel writes global memory before its wait (stream-ordered scratch may be reused across calls).
__device__ __forceinline__ void chain_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void chain_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Canonical summation.  The bands of a call are ordered by ascending scale (ties: band index), c0 < c1 < ... < ck, and a
// sample is ALWAYS formed as  B_c0 + (B_c1 + (... + (B_ck + 0)))  where each band value B is computed from zero by
// band_value().  Folded bands are a suffix of that order and their period block holds the canonical sum of the suffix, so
// a kernel starts from the period-block value (or 0) and adds its own bands from the highest scale down.  The result
// of a sample therefore does not depend on which bands were folded, how period blocks nest, which kernel ran, or how
// the lattice was partitioned into calls: shards, chunks and the single call are bit-identical.
__device__ __forceinline__ float band_value(const float4 tz, float v0, float v1, float v2)
{
    return fmaf(tz.x, v0, fmaf(tz.y, v1, __fmul_rn(tz.z, v2)));
}

inline WnFold make_fold(const float *P, int Lx, int Ly, int Lz, int kphase)
{
    WnFold f{P, Lx, Ly, Lz, kphase, -1, -1};
    if ((Lx & (Lx - 1)) == 0) f.xmask = Lx - 1;
    if ((Ly & (Ly - 1)) == 0) f.ymask = Ly - 1;
    return f;
}

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
    chain_wait();
    chain_release();
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

template <int BY, int BZ, int NT, bool POW2>
__global__ void __launch_bounds__(NT)
k_mb3d_brick(const float *__restrict__ N, int n, WnTabs tabs, int nx, int ny, int nk,
             int max_rows, WnFold fold, float *__restrict__ out)
{
    const int nbands = tabs.nb;
    constexpr int NW = NT / 32;
    constexpr int C = BY / NW;                  // y columns per thread
    constexpr int PER_BAND = 32 + BY + BZ;      // table entries of one band this brick needs
    constexpr int pow2 = POW2 ? 1 : 0;          // tile edge is a power of two: Mod() is a mask
    static_assert(BY % NW == 0, "BY must be a multiple of the warp count");
    static_assert(PER_BAND <= 64 && NT % 64 == 0, "phase 0 maps 64 threads to one band");
    // dynamic shared memory: U[max_rows][32] | band tables: nbands x (x[32], y[BY], z[BZ]) float4 | rowoff[max_rows]
    // (max_rows is a multiple of 4)
    extern __shared__ float4 smem4[];
    float *U = reinterpret_cast<float *>(smem4);
    const float4 *s_tab = smem4 + max_rows * 8;
    int *s_rowoff = reinterpret_cast<int *>(smem4 + max_rows * 8 + nbands * PER_BAND);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i0 = blockIdx.x * 32, j0 = blockIdx.y * BY, k0 = blockIdx.z * BZ;
    const int pitch = n + WN_TILE_PAD;
    if (tabs.wait_first || fold.P) { chain_wait(); chain_release(); }   // tables / period block of the previous kernel

    // ---- phase 0: every table entry of every band in one round of independent loads (64 threads per band)
    {
        const int q = threadIdx.x & 63;
        for (int b = threadIdx.x >> 6; b < nbands; b += NT / 64) {
            if (q < PER_BAND) {
                float4 v;
                const int row = tabs_row(tabs, b);
                if (q < 32)           v = __ldg(tabs.x + row * tabs.sx + min(i0 + q, nx - 1));
                else if (q < 32 + BY) v = __ldg(tabs.y + row * tabs.sy + min(j0 + q - 32, ny - 1));
                else                  v = __ldg(tabs.z + row * tabs.sz + tabs.kz0 + min(k0 + q - 32 - BY, nk - 1));
                smem4[max_rows * 8 + b * PER_BAND + q] = v;
            }
        }
    }
    __syncthreads();

    // ---- running sums start from the period-block value (canonical summation: folded bands first)
    const int i = i0 + lane;
    const size_t plane = (size_t)nx * ny;
    float acc[C][BZ];
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
        for (int k = 0; k < BZ; ++k) acc[c][k] = 0.0f;
    if (fold.P) {
        const unsigned pplane = (unsigned)(fold.Lx * fold.Ly);
        const int ic = min(i, nx - 1);
        const unsigned pi = (unsigned)(fold.xmask >= 0 ? (ic & fold.xmask) : ic % fold.Lx);
        int kz = fold.kphase + k0;
        if (kz >= fold.Lz) kz %= fold.Lz;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j = min(j0 + warp + c * NW, ny - 1);
            const unsigned pj = pi + (unsigned)(fold.ymask >= 0 ? (j & fold.ymask) : j % fold.Ly) * (unsigned)fold.Lx;
            int kk = kz;
#pragma unroll
            for (int k = 0; k < BZ; ++k) {
                acc[c][k] = __ldg(fold.P + (pj + (unsigned)kk * pplane));
                if (++kk == fold.Lz) kk = 0;
            }
        }
    }

    for (int b = nbands - 1; b >= 0; --b) {                    // highest scale first
        const float4 *tX = s_tab + b * PER_BAND, *tY = tX + 32, *tZ = tY + BY;
        // clamped entries repeat the last valid one, so the last entry carries the brick's last tap cell
        const int my0 = __float_as_int(tY[0].w), mz0 = __float_as_int(tZ[0].w);
        const int Ey = __float_as_int(tY[BY - 1].w) - my0 + 3;
        const int Ez = __float_as_int(tZ[BZ - 1].w) - mz0 + 3;
        const int slab = Ey * 32;
        const int rows = Ey * Ez, rows4 = (rows + 3) & ~3;

        // row r = cz*Ey + cy of the brick's footprint -> offset of the (x-padded) tile row (wrapped y, wrapped z);
        // the table is padded to a multiple of 4 rows by repeating the last row
        {
            const float inv_ey = 1.0f / (float)Ey;
            for (int r = threadIdx.x; r < rows4; r += NT) {
                const int rr = min(r, rows - 1);
                const int cz = (int)(((float)rr + 0.5f) * inv_ey), cy = rr - cz * Ey;   // exact: rows <= 1500
                s_rowoff[r] = (tmodf(mz0 + cz, n, pow2) * n + tmodf(my0 + cy, n, pow2)) * pitch;
            }
        }
        __syncthreads();

        // ---- X pass: U[row][lane] = sum_f wx[f] * N[row][cx + f]; each warp takes groups of 4 consecutive rows
        {
            const float4 tx = tX[lane];
            // padded rows hold cells 0..n+1: the three taps are x0, x0+1, x0+2 of one address
            const float *base = N + tmodf(__float_as_int(tx.w), n, pow2);
            for (int r = warp * 4; r < rows4; r += NW * 4) {
                const int4 o = *reinterpret_cast<const int4 *>(s_rowoff + r);
                const float *q0 = base + (unsigned)o.x, *q1 = base + (unsigned)o.y;
                const float *q2 = base + (unsigned)o.z, *q3 = base + (unsigned)o.w;
                const float a0 = __ldg(q0), a1 = __ldg(q0 + 1), a2 = __ldg(q0 + 2);
                const float b0 = __ldg(q1), b1 = __ldg(q1 + 1), b2 = __ldg(q1 + 2);
                const float c0 = __ldg(q2), c1 = __ldg(q2 + 1), c2 = __ldg(q2 + 2);
                const float d0 = __ldg(q3), d1 = __ldg(q3 + 1), d2 = __ldg(q3 + 2);
                float *u = U + r * 32 + lane;                  // rows past `rows` are scratch (max_rows % 4 == 0)
                u[0]  = fmaf(tx.z, a2, fmaf(tx.y, a1, tx.x * a0));
                u[32] = fmaf(tx.z, b2, fmaf(tx.y, b1, tx.x * b0));
                u[64] = fmaf(tx.z, c2, fmaf(tx.y, c1, tx.x * c0));
                u[96] = fmaf(tx.z, d2, fmaf(tx.y, d1, tx.x * d0));
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
                ty[c] = tY[warp + c * NW];
                ucol[c] = U + (__float_as_int(ty[c].w) - my0) * 32 + lane + 2 * slab;
                v[c][0] = v[c][1] = v[c][2] = 0.0f;
            }
            int base = -3;                       // window holds V[base .. base+2]
#pragma unroll
            for (int k = 0; k < BZ; ++k) {
                const float4 tz = tZ[k];
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
                    acc[c][k] = __fadd_rn(band_value(tz, v[c][0], v[c][1], v[c][2]), acc[c][k]);
            }
        }
        if (b > 0) __syncthreads();
    }

    if (!(tabs.wait_first || fold.P)) { chain_wait(); chain_release(); }   // nothing was read from the previous kernel so far
    if (i < nx) {
        const bool full = k0 + BZ <= nk;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j = j0 + warp + c * NW;
            if (j >= ny) continue;
            float *o = out + ((size_t)i + (size_t)nx * j + plane * k0);
            if (full) {
#pragma unroll
                for (int k = 0; k < BZ; ++k) { __stcs(o, acc[c][k]); o += plane; }
            } else {
#pragma unroll
                for (int k = 0; k < BZ; ++k) {
                    if (k0 + k < nk) __stcs(o, acc[c][k]);
                    o += plane;
                }
            }
        }
    }
}

// ---- pieces shared by the float4 kernels (k_mb3d_brick4, k_mb3d_col4): a CTA owns 128 x BY x BZ samples, lane l owns
// x = 4l..4l+3.  Shared memory (float4 units): U[max_rows][32] | tables nbands x (128 + BY + BZ) | rowoff[max_rows] (int)
struct Q4Foot {
    int ey[WN_MAX_BANDS], ez[WN_MAX_BANDS], row0[WN_MAX_BANDS + 1];
    // k_mb3d_rep: window advances of the first two bands at z step k in bits [2k, 2k+2) (planes entering the 3-deep
    // window: first tap cell of step k minus that of step k-1); slow != 0 when some step advances by more than 3
    unsigned long long adv[2];
    int slow;
};

// axis-table entries of every band for this brick; ends with a barrier
template <int BY, int BZ, int NT>
__device__ __forceinline__ void q4_load_tables(float4 *s_tab, const WnTabs &tabs, int i0, int j0, int k0,
                                               int nx, int ny, int nk)
{
    constexpr int BX = 128, PER_BAND = BX + BY + BZ;
    for (int e = threadIdx.x; e < tabs.nb * PER_BAND; e += NT) {
        const int b = e / PER_BAND, q = e - b * PER_BAND;
        float4 v;
        const int row = tabs_row(tabs, b);
        if (q < BX)           v = __ldg(tabs.x + row * tabs.sx + min(i0 + q, nx - 1));
        else if (q < BX + BY) v = __ldg(tabs.y + row * tabs.sy + min(j0 + q - BX, ny - 1));
        else                  v = __ldg(tabs.z + row * tabs.sz + tabs.kz0 + min(k0 + q - BX - BY, nk - 1));
        // x entries are stored slot-major ([sample slot 0..3][lane]) so a lane's four LDS.128 are conflict-free
        s_tab[q < BX ? b * PER_BAND + (q & 3) * 32 + (q >> 2) : e] = v;
    }
    __syncthreads();
}

// footprints of all bands (Ey, Ez, first U row: U holds every band at once) and the tile-row offset of every U row;
// ends with a barrier
template <int BY, int BZ, int NT, bool POW2>
__device__ __forceinline__ void q4_footprints(const float4 *s_tab, int nbands, int n, Q4Foot &ft, int *s_rowoff)
{
    constexpr int BX = 128, PER_BAND = BX + BY + BZ, pow2 = POW2 ? 1 : 0;
    const int pitch = n + WN_TILE_PAD;
    if (threadIdx.x == 0) {
        int row0 = 0;
        for (int b = 0; b < nbands; ++b) {
            const float4 *tY = s_tab + b * PER_BAND + BX, *tZ = tY + BY;
            // clamped entries repeat the last valid one, so the last entry carries the brick's last tap cell
            const int ey = __float_as_int(tY[BY - 1].w) - __float_as_int(tY[0].w) + 3;
            const int ez = __float_as_int(tZ[BZ - 1].w) - __float_as_int(tZ[0].w) + 3;
            ft.ey[b] = ey; ft.ez[b] = ez; ft.row0[b] = row0;
            row0 += (ey * ez + 1) & ~1;                      // each band starts on an even row
        }
        ft.row0[nbands] = row0;
    }
    __syncthreads();
    for (int b = 0; b < nbands; ++b) {
        const float4 *tY = s_tab + b * PER_BAND + BX, *tZ = tY + BY;
        const int my0 = __float_as_int(tY[0].w), mz0 = __float_as_int(tZ[0].w);
        const int Ey = ft.ey[b], rows = Ey * ft.ez[b], rows2 = (rows + 1) & ~1, row0 = ft.row0[b];
        const float inv_ey = 1.0f / (float)Ey;
        for (int r = threadIdx.x; r < rows2; r += NT) {
            const int rr = min(r, rows - 1);
            const int cz = (int)(((float)rr + 0.5f) * inv_ey), cy = rr - cz * Ey;       // exact for rows < 2^20
            s_rowoff[row0 + r] = (tmodf(mz0 + cz, n, pow2) * n + tmodf(my0 + cy, n, pow2)) * pitch;
        }
    }
    __syncthreads();
}

// X pass of one band: U[row][4 lanes-samples] = sum_f wx[f] * N[row][cx + f] for rows [row0, row1)
template <int NT, bool POW2>
__device__ __forceinline__ void q4_xpass(const float *__restrict__ N, int n, const float4 *tX, float4 *U4,
                                         const int *s_rowoff, int row0, int row1)
{
    constexpr int NW = NT / 32, pow2 = POW2 ? 1 : 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float4 t0 = tX[lane], t1 = tX[32 + lane], t2 = tX[64 + lane], t3 = tX[96 + lane];
    const int c0 = __float_as_int(t0.w), c1 = __float_as_int(t1.w), c2 = __float_as_int(t2.w), c3 = __float_as_int(t3.w);
    const int cmin = min(min(c0, c1), min(c2, c3)), cmax = max(max(c0, c1), max(c2, c3));
    // Bands with <= 1/2 cell per sample: a lane's four samples touch at most the four cells cmin..cmin+3, so
    // four loads serve all of them.  Each sample's three weights are placed on that 4-cell footprint (the
    // unused slot is an exact 0, fmaf(0, a, x) == x), which keeps the result bit-identical to the 12-load form.
    const bool narrow = __all_sync(0xffffffffu, cmax - cmin <= 1);
    if (narrow) {
        const float *base = N + tmodf(cmin, n, pow2);     // padded rows hold cells 0..n+2, so cmin+3 stays in the row
        const bool s0 = c0 != cmin, s1 = c1 != cmin, s2 = c2 != cmin, s3 = c3 != cmin;
        // W[e][0..3]: shifted by one slot when the sample's first cell is cmin + 1
        const float a00 = s0 ? 0.0f : t0.x, a01 = s0 ? t0.x : t0.y, a02 = s0 ? t0.y : t0.z, a03 = s0 ? t0.z : 0.0f;
        const float a10 = s1 ? 0.0f : t1.x, a11 = s1 ? t1.x : t1.y, a12 = s1 ? t1.y : t1.z, a13 = s1 ? t1.z : 0.0f;
        const float a20 = s2 ? 0.0f : t2.x, a21 = s2 ? t2.x : t2.y, a22 = s2 ? t2.y : t2.z, a23 = s2 ? t2.z : 0.0f;
        const float a30 = s3 ? 0.0f : t3.x, a31 = s3 ? t3.x : t3.y, a32 = s3 ? t3.y : t3.z, a33 = s3 ? t3.z : 0.0f;
        for (int r = row0 + warp * 2; r < row1; r += NW * 2) {
            const int2 o = *reinterpret_cast<const int2 *>(s_rowoff + r);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float *q = base + (unsigned)(h ? o.y : o.x);
                const float v0 = __ldg(q), v1 = __ldg(q + 1), v2 = __ldg(q + 2), v3 = __ldg(q + 3);
                float4 u;
                u.x = fmaf(a03, v3, fmaf(a02, v2, fmaf(a01, v1, a00 * v0)));
                u.y = fmaf(a13, v3, fmaf(a12, v2, fmaf(a11, v1, a10 * v0)));
                u.z = fmaf(a23, v3, fmaf(a22, v2, fmaf(a21, v1, a20 * v0)));
                u.w = fmaf(a33, v3, fmaf(a32, v2, fmaf(a31, v1, a30 * v0)));
                U4[(r + h) * 32 + lane] = u;
            }
        }
    } else {
        const float *b0 = N + tmodf(c0, n, pow2), *b1 = N + tmodf(c1, n, pow2);
        const float *b2 = N + tmodf(c2, n, pow2), *b3 = N + tmodf(c3, n, pow2);
        for (int r = row0 + warp * 2; r < row1; r += NW * 2) {
            const int2 o = *reinterpret_cast<const int2 *>(s_rowoff + r);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const unsigned off = (unsigned)(h ? o.y : o.x);
                const float *q0 = b0 + off, *q1 = b1 + off, *q2 = b2 + off, *q3 = b3 + off;
                float4 u;
                u.x = fmaf(t0.z, __ldg(q0 + 2), fmaf(t0.y, __ldg(q0 + 1), t0.x * __ldg(q0)));
                u.y = fmaf(t1.z, __ldg(q1 + 2), fmaf(t1.y, __ldg(q1 + 1), t1.x * __ldg(q1)));
                u.z = fmaf(t2.z, __ldg(q2 + 2), fmaf(t2.y, __ldg(q2 + 1), t2.x * __ldg(q2)));
                u.w = fmaf(t3.z, __ldg(q3 + 2), fmaf(t3.y, __ldg(q3 + 1), t3.x * __ldg(q3)));
                U4[(r + h) * 32 + lane] = u;
            }
        }
    }
}

// y contraction of one tile-z plane for a lane's four samples: V = sum_f wy[f] * U[cz][cy + f]
__device__ __forceinline__ float4 q4_ycontract(const float4 *uu, const float4 ty)
{
    const float4 u0 = uu[0], u1 = uu[32], u2 = uu[64];
    float4 nv;
    nv.x = fmaf(ty.z, u2.x, fmaf(ty.y, u1.x, ty.x * u0.x));
    nv.y = fmaf(ty.z, u2.y, fmaf(ty.y, u1.y, ty.x * u0.y));
    nv.z = fmaf(ty.z, u2.z, fmaf(ty.y, u1.z, ty.x * u0.z));
    nv.w = fmaf(ty.z, u2.w, fmaf(ty.y, u1.w, ty.x * u0.w));
    return nv;
}

// band value of a lane's four samples (see band_value) and the canonical running sum
__device__ __forceinline__ float4 q4_band_value(const float4 tz, const float4 (&v)[3])
{
    float4 a;
    a.x = band_value(tz, v[0].x, v[1].x, v[2].x);
    a.y = band_value(tz, v[0].y, v[1].y, v[2].y);
    a.z = band_value(tz, v[0].z, v[1].z, v[2].z);
    a.w = band_value(tz, v[0].w, v[1].w, v[2].w);
    return a;
}
__device__ __forceinline__ float4 q4_band_add(const float4 bv, const float4 s)
{
    return make_float4(__fadd_rn(bv.x, s.x), __fadd_rn(bv.y, s.y), __fadd_rn(bv.z, s.z), __fadd_rn(bv.w, s.w));
}

// ---- k_mb3d_brick4: same algorithm, four x-samples per thread --------------------------------------------------
// A CTA owns 128 x BY x BZ samples; lane l owns x = 4l..4l+3, so U rows, the y-contracted window, the accumulators,
// the period-block loads and the output stores are all float4 (LDS.128 / LDG.128 / STG.128): the per-sample FMA
// count is unchanged but every other instruction is amortised over four samples.  Used for small footprints
// (bands with <= ~1 cell per sample) when three or more bands are left per sample, and for the inner period blocks;
// the per-sample arithmetic is identical to k_mb3d_brick.
template <int BY, int BZ, int NT, bool POW2>
__global__ void __launch_bounds__(NT, BY * BZ <= 64 ? 3 : 1)
k_mb3d_brick4(const float *__restrict__ N, int n, WnTabs tabs, int nx, int ny, int nk,
              int max_rows, WnFold fold, float *__restrict__ out)
{
    const int nbands = tabs.nb;
    constexpr int NW = NT / 32;
    constexpr int C = BY / NW;
    constexpr int BX = 128;
    constexpr int PER_BAND = BX + BY + BZ;
    static_assert(BY % NW == 0, "BY must be a multiple of the warp count");
    // dynamic shared memory (float4 units): U[max_rows][32] | tables nbands x PER_BAND | rowoff[max_rows] (int)
    extern __shared__ float4 smem4[];
    float4 *U4 = smem4;
    float4 *s_tab = smem4 + max_rows * 32;
    int *s_rowoff = reinterpret_cast<int *>(s_tab + nbands * PER_BAND);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i0 = blockIdx.x * BX, j0 = blockIdx.y * BY, k0 = blockIdx.z * BZ;
    if (tabs.wait_first || fold.P) { chain_wait(); chain_release(); }   // tables / period block of the previous kernel

    q4_load_tables<BY, BZ, NT>(s_tab, tabs, i0, j0, k0, nx, ny, nk);
    __shared__ Q4Foot ft;
    q4_footprints<BY, BZ, NT, POW2>(s_tab, nbands, n, ft, s_rowoff);

    // ---- X pass for every band
    for (int b = 0; b < nbands; ++b)
        q4_xpass<NT, POW2>(N, n, s_tab + b * PER_BAND, U4, s_rowoff, ft.row0[b], ft.row0[b + 1]);
    __syncthreads();

    // ---- running sums start from the period-block value (canonical summation: folded bands first).  Host guarantees
    // nx % 4 == 0 and, when folding, Lx % 4 == 0: float4 loads and stores.
    const int i = i0 + 4 * lane;
    float4 acc[C][BZ];
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
        for (int k = 0; k < BZ; ++k) acc[c][k] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (fold.P) {
        const unsigned pplane = (unsigned)(fold.Lx * fold.Ly);
        const int ic = min(i, nx - 4);
        const unsigned pi = (unsigned)(fold.xmask >= 0 ? (ic & fold.xmask) : ic % fold.Lx);
        int kz = fold.kphase + k0;
        if (kz >= fold.Lz) kz %= fold.Lz;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j = min(j0 + warp + c * NW, ny - 1);
            const unsigned pj = pi + (unsigned)(fold.ymask >= 0 ? (j & fold.ymask) : j % fold.Ly) * (unsigned)fold.Lx;
            int kk = kz;
#pragma unroll
            for (int k = 0; k < BZ; ++k) {
                acc[c][k] = __ldg(reinterpret_cast<const float4 *>(fold.P + (pj + (unsigned)kk * pplane)));
                if (++kk == fold.Lz) kk = 0;
            }
        }
    }


    // ---- YZ pass for every band, highest scale first
    for (int b = nbands - 1; b >= 0; --b) {
        const float4 *tY = s_tab + b * PER_BAND + BX, *tZ = tY + BY;
        const int my0 = __float_as_int(tY[0].w), mz0 = __float_as_int(tZ[0].w);
        const int slab4 = ft.ey[b] * 32;                         // one cz plane of this band's U in float4 units
        float4 ty[C];
        const float4 *ucol[C];
        float4 v[C][3];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            ty[c] = tY[warp + c * NW];
            ucol[c] = U4 + ft.row0[b] * 32 + (__float_as_int(ty[c].w) - my0) * 32 + lane + 2 * slab4;
            v[c][0] = v[c][1] = v[c][2] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
        int base = -3;
#pragma unroll
        for (int k = 0; k < BZ; ++k) {
            const float4 tz = tZ[k];
            const int rel = __float_as_int(tz.w) - mz0;
            while (base < rel) {
                ++base;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float4 nv = q4_ycontract(ucol[c] + base * slab4, ty[c]);
                    v[c][0] = v[c][1]; v[c][1] = v[c][2]; v[c][2] = nv;
                }
            }
#pragma unroll
            for (int c = 0; c < C; ++c) acc[c][k] = q4_band_add(q4_band_value(tz, v[c]), acc[c][k]);
        }
    }

    // ---- epilogue: float4 streaming stores
    if (!(tabs.wait_first || fold.P)) { chain_wait(); chain_release(); }   // nothing was read from the previous kernel so far
    if (i < nx) {
        const size_t plane = (size_t)nx * ny;
        const bool full = k0 + BZ <= nk;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j = j0 + warp + c * NW;
            if (j >= ny) continue;
            float *o = out + ((size_t)i + (size_t)nx * j + plane * k0);
#pragma unroll
            for (int k = 0; k < BZ; ++k) {
                if (full || k0 + k < nk) __stcs(reinterpret_cast<float4 *>(o), acc[c][k]);
                o += plane;
            }
        }
    }
}

// ---- k_mb3d_col4: one or two direct bands, streaming along z -----------------------------------------------------
// Same tables, X pass and per-sample arithmetic as k_mb3d_brick4, but the band loop is inside the z loop: each band
// keeps only its 3-deep y-contracted window in registers and every sample is finished (period-block value added,
// stored) as soon as it is computed.  Without BZ accumulators per thread the brick can be long in z (BZ = 32), which
// amortises the tables, the footprint set-up and the X pass over four times as many samples and leaves registers for
// five CTAs per SM.  This is the main kernel whenever folding leaves at most two bands to evaluate per sample.
//
// Block order: when the period block divides the lattice, the (y, z) bricks are enumerated replica-first
// (blockIdx.y = ((y brick in period) * yrep + y replica) * zrep + z replica, blockIdx.z = z brick in period), so the
// CTAs that add the same period-block lines run back to back and the block is read from HBM once even when it is
// much larger than L2.
struct WnOrder { int yrep, yper, zrep, zper; };          // replicas and bricks per period; {1, nyb, 1, nzb} = plain order

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(PENDING) : "memory"); }

template <int NB, int BY, int BZ, int NT, int RING, bool POW2>
__global__ void __launch_bounds__(NT, NB == 1 ? 4 : 3)
k_mb3d_col4(const float *__restrict__ N, int n, WnTabs tabs, int nx, int ny, int nk,
            int max_rows, WnFold fold, WnOrder ord, float *__restrict__ out)
{
    constexpr int BX = 128;
    constexpr int PER_BAND = BX + BY + BZ;
    static_assert(BY == NT / 32, "one y row per warp");
    extern __shared__ float4 smem4[];
    float4 *U4 = smem4;
    float4 *s_tab = smem4 + max_rows * 32;
    int *s_rowoff = reinterpret_cast<int *>(s_tab + NB * PER_BAND);

    chain_wait();                                              // the ring prefetch below reads the previous kernel's output
    chain_release();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int yb, zb;
    {
        const int gy = blockIdx.y, rz = gy % ord.zrep, t = gy / ord.zrep;
        yb = t / ord.yrep + (t % ord.yrep) * ord.yper;
        zb = rz * ord.zper + blockIdx.z;
    }
    const int i0 = blockIdx.x * BX, j0 = yb * BY, k0 = zb * BZ;
    const int i = i0 + 4 * lane, j = j0 + warp;
    const bool active = i < nx && j < ny;
    const int kmax = min(BZ, nk - k0);

    // Period-block column of this thread (host guarantees nx % 4 == 0 and Lx % 4 == 0).  Its values travel through a
    // per-thread ring of RING float4 slots in shared memory filled by cp.async: the first RING-1 planes are requested
    // here, before the tables and the X pass, and plane k+RING-1 is requested when plane k is consumed, so the loads
    // have the whole prologue and RING-1 z steps to land and cost no registers.
    // RING < 0: the same look-ahead in registers instead (PF = -RING planes, plain loads): fewer LSU wavefronts per
    // sample, 4 registers per plane.
    constexpr int PF = RING < 0 ? -RING : 0;
    static_assert(RING != 0 && (PF == 0 || BZ % PF == 0), "look-ahead depth");
    float4 *ring = reinterpret_cast<float4 *>(s_rowoff + max_rows) + threadIdx.x;
    float4 pq[PF > 0 ? PF : 1];
#pragma unroll
    for (int d = 0; d < (PF > 0 ? PF : 1); ++d) pq[d] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    const size_t pplane = (size_t)fold.Lx * fold.Ly;
    const float *pcol = nullptr, *pp = nullptr;
    int kk = 0;
    const bool folded = fold.P != nullptr;
    if (folded && active) {
        const unsigned pi = (unsigned)(fold.xmask >= 0 ? (i & fold.xmask) : i % fold.Lx);
        const unsigned pj = (unsigned)(fold.ymask >= 0 ? (j & fold.ymask) : j % fold.Ly);
        pcol = fold.P + (pi + pj * (unsigned)fold.Lx);
        kk = fold.kphase + k0;
        if (kk >= fold.Lz) kk %= fold.Lz;
        pp = pcol + kk * pplane;
        if constexpr (RING > 0) {
#pragma unroll
            for (int d = 0; d < RING - 1; ++d) {
                if (d < kmax) {
                    cp_async16(ring + d * NT, pp);
                    if (++kk == fold.Lz) { kk = 0; pp = pcol; } else pp += pplane;
                }
                cp_async_commit();
            }
        } else {
#pragma unroll
            for (int d = 0; d < PF; ++d) {
                if (d < kmax) {
                    pq[d] = __ldg(reinterpret_cast<const float4 *>(pp));
                    if (++kk == fold.Lz) { kk = 0; pp = pcol; } else pp += pplane;
                }
            }
        }
    }

    q4_load_tables<BY, BZ, NT>(s_tab, tabs, i0, j0, k0, nx, ny, nk);
    __shared__ Q4Foot ft;
    q4_footprints<BY, BZ, NT, POW2>(s_tab, NB, n, ft, s_rowoff);
#pragma unroll
    for (int b = 0; b < NB; ++b)
        q4_xpass<NT, POW2>(N, n, s_tab + b * PER_BAND, U4, s_rowoff, ft.row0[b], ft.row0[b + 1]);
    __syncthreads();
    if (!active) return;                                       // no barrier below

    // per band: y weights, the 3-deep window of y-contracted tile-z planes (starts at the brick's first tap plane) and
    // the next plane to contract
    float4 ty[NB], v[NB][3];
    const float4 *unext[NB], *tZ[NB];
    int slab4[NB], mz0[NB], base[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const float4 *tY = s_tab + b * PER_BAND + BX;
        tZ[b] = tY + BY;
        ty[b] = tY[warp];
        slab4[b] = ft.ey[b] * 32;
        mz0[b] = __float_as_int(tZ[b][0].w);
        const float4 *u = U4 + ft.row0[b] * 32 + (__float_as_int(ty[b].w) - __float_as_int(tY[0].w)) * 32 + lane;
#pragma unroll
        for (int f = 0; f < 3; ++f) v[b][f] = q4_ycontract(u + f * slab4[b], ty[b]);
        unext[b] = u + 3 * slab4[b];
        base[b] = 0;
    }

    const size_t plane = (size_t)nx * ny;
    float *o = out + ((size_t)i + (size_t)nx * j + plane * k0);
    // one z step: the value of every band at sample plane k (each from zero, see band_value)
    auto zstep = [&](int k, float4 (&bv)[NB]) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const float4 tz = tZ[b][k];
            const int rel = __float_as_int(tz.w) - mz0[b];
#pragma unroll 1
            while (base[b] < rel) {
                ++base[b];
                const float4 nv = q4_ycontract(unext[b], ty[b]);
                unext[b] += slab4[b];
                v[b][0] = v[b][1]; v[b][1] = v[b][2]; v[b][2] = nv;
            }
            bv[b] = q4_band_value(tz, v[b]);
        }
    };
    // canonical sum: period-block value (or 0) first, then the bands from the highest scale down
    auto combine = [&](const float4 (&bv)[NB], const float4 pv) {
        float4 s = pv;                                         // zero when nothing is folded
#pragma unroll
        for (int b = NB - 1; b >= 0; --b) s = q4_band_add(bv[b], s);
        return s;
    };
    if constexpr (RING > 0) {
#pragma unroll 4
        for (int k = 0; k < kmax; ++k) {
            float4 bv[NB];
            zstep(k, bv);
            float4 pv = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (folded) {
                if (k + RING - 1 < kmax) {                     // slot of plane k-1, consumed in the previous step
                    cp_async16(ring + ((k + RING - 1) % RING) * NT, pp);
                    if (++kk == fold.Lz) { kk = 0; pp = pcol; } else pp += pplane;
                }
                cp_async_commit();
                cp_async_wait<RING - 1>();                     // plane k has landed
                pv = ring[(k % RING) * NT];
            }
            __stcs(reinterpret_cast<float4 *>(o), combine(bv, pv));
            o += plane;
        }
    } else {
        for (int kb = 0; kb < kmax; kb += PF) {
#pragma unroll
            for (int d = 0; d < PF; ++d) {
                const int k = kb + d;
                if (k < kmax) {
                    float4 bv[NB];
                    zstep(k, bv);
                    const float4 pv = pq[d];
                    if (folded && k + PF < kmax) {
                        pq[d] = __ldg(reinterpret_cast<const float4 *>(pp));
                        if (++kk == fold.Lz) { kk = 0; pp = pcol; } else pp += pplane;
                    }
                    __stcs(reinterpret_cast<float4 *>(o), combine(bv, pv));
                    o += plane;
                }
            }
        }
    }
}

// ---- k_mb3d_rep: k_mb3d_col4 with the period-block value shared between replicas -------------------------------------
// The main kernel is bound by data movement, not arithmetic: every sample costs one period-block read (through L2)
// and one output write.  Samples that are a whole number of period blocks apart read the SAME period-block value, so a
// thread here owns R = RX * RY of them: sample (i, j, k) of the "base region" [0, bx) x [0, by) and its replicas
// (i + qx*bx, j + qy*by, k).  The period-block value is fetched once per R outputs (L2 -> SM traffic, ring fill and
// read-back divided by R), the z-table entry and the window bookkeeping are shared too.
//   * Host guarantee (plan_replicas): for every direct band the axis-table entries at i and i + bx carry bit-identical
//     weights and first tap cells that differ by a constant (cxs[b] cells; cys[b] along y).  A replica is then the same
//     computation on a tile shifted by that many cells: own rows of U (X pass), own 3-deep window, same weights.
//   * Thread mapping: a warp covers 32/YPW float4 columns x YPW y rows.  With YPW = 4 the lanes of a warp that share
//     x read the same U rows while a band advances less than one cell per y sample, so a window advance costs one
//     shared-memory wavefront instead of four; stores stay whole 128-byte lines (32 samples per row and warp).
//   * Arithmetic: the same operations in the same order as k_mb3d_col4 / brick4 / brick (bit-identical samples), issued
//     as packed pairs (FFMA2 / FMUL2 / FADD2, sm_100): half the issue slots for the same IEEE results.
//   * Period-block ring, RINGMODE 0: per-thread cp.async slots as in k_mb3d_col4 (any brick, any period block).
//     RINGMODE 1: one elected thread fetches the CTA's 128 x 8 x 4 sub-box of the period block with ONE TMA tensor copy
//     (cp.async.bulk.tensor.3d, 16 KB) per four z planes into one half of an 8-plane ring; full/empty mbarriers per
//     half, consumers wait once per four planes and the LSU pipe carries no fill traffic.  Needs whole bricks and a
//     period block whose edges are multiples of the box (checked by the host).
struct WnRep {
    int bx, by;                 // base region in samples; the grid covers it, replica (qx, qy) adds (qx*bx, qy*by)
    int cxs[2], cys[2];         // per direct band: first-tap-cell shift of one replica step along x / y
    long long roff[4];          // byte offset of replica r = qy*RX + qx in the output: 4 * (qx*bx + nx*qy*by)
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WN_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra WN_DONE;\n"
        "bra WN_WAIT;\n"
        "WN_DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// one 3D box of the period block -> shared memory; completion is signalled on `bar` (complete_tx)
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, int x, int y, int z, unsigned long long *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}

// packed forms of q4_ycontract / q4_band_value / q4_band_add: the same IEEE operations, two lanes per instruction
__device__ __forceinline__ float2 lo2(const float4 v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(const float4 v) { return make_float2(v.z, v.w); }
__device__ __forceinline__ float4 cat4(const float2 a, const float2 b) { return make_float4(a.x, a.y, b.x, b.y); }
__device__ __forceinline__ float4 p4_contract3(const float w0, const float w1, const float w2, const float4 a0, const float4 a1,
                                               const float4 a2)
{
    // per component: fmaf(w2, a2, fmaf(w1, a1, w0 * a0))
    const float2 s0 = make_float2(w0, w0), s1 = make_float2(w1, w1), s2 = make_float2(w2, w2);
    const float2 l = __ffma2_rn(s2, lo2(a2), __ffma2_rn(s1, lo2(a1), __fmul2_rn(s0, lo2(a0))));
    const float2 h = __ffma2_rn(s2, hi2(a2), __ffma2_rn(s1, hi2(a1), __fmul2_rn(s0, hi2(a0))));
    return cat4(l, h);
}
__device__ __forceinline__ float4 p4_band_value(const float4 tz, const float4 (&v)[3])
{
    // per component: fmaf(tz.x, v0, fmaf(tz.y, v1, tz.z * v2))  == band_value()
    return p4_contract3(tz.z, tz.y, tz.x, v[2], v[1], v[0]);
}
__device__ __forceinline__ float4 p4_add(const float4 a, const float4 b)
{
    return cat4(__fadd2_rn(lo2(a), lo2(b)), __fadd2_rn(hi2(a), hi2(b)));
}

// footprints of all bands with R replicas each: band b owns U rows [row0[b], row0[b+1]) = R blocks of rows2 rows, the
// block of replica r = qy*RX + qx holds the tile rows shifted by qy*cys[b] cells in y (the x shift is applied by the X
// pass); ends with a barrier
template <int BY, int BZ, int NT, int RX, int RY, int NSH, bool POW2>
__device__ __forceinline__ void q4_footprints_rep(const float4 *s_tab, int nbands, int n, Q4Foot &ft, int *s_rowoff,
                                                  const WnRep &rep)
{
    constexpr int BX = 128, PER_BAND = BX + BY + BZ, pow2 = POW2 ? 1 : 0, R = RX * RY;
    // the last NSH bands are SHARED: their replicas coincide (cell shift = a whole number of tile periods), one block
    const int pitch = n + WN_TILE_PAD;
    if (threadIdx.x == 0) {
        int row0 = 0;
        for (int b = 0; b < nbands; ++b) {
            const float4 *tY = s_tab + b * PER_BAND + BX, *tZ = tY + BY;
            const int ey = __float_as_int(tY[BY - 1].w) - __float_as_int(tY[0].w) + 3;
            const int ez = __float_as_int(tZ[BZ - 1].w) - __float_as_int(tZ[0].w) + 3;
            ft.ey[b] = ey; ft.ez[b] = ez; ft.row0[b] = row0;
            row0 += (b >= nbands - NSH ? 1 : R) * ((ey * ez + 1) & ~1);
        }
        ft.row0[nbands] = row0;
    }
    if (threadIdx.x < 32) {                                    // advance masks (BZ == 32 steps, one lane per step)
        int slow = 0;
        for (int b = 0; b < nbands && b < 2; ++b) {
            const float4 *tZ = s_tab + b * PER_BAND + BX + BY;
            const int k = threadIdx.x;
            const int a = k == 0 ? 0 : __float_as_int(tZ[k].w) - __float_as_int(tZ[k - 1].w);
            slow |= __any_sync(0xffffffffu, a > 3 || a < 0);
            unsigned long long m = (unsigned long long)(a & 3) << (2 * k);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m |= __shfl_xor_sync(0xffffffffu, m, o);
            if (k == 0) ft.adv[b] = m;
        }
        if (threadIdx.x == 0) ft.slow = slow;
    }
    __syncthreads();
    for (int b = 0; b < nbands; ++b) {
        const float4 *tY = s_tab + b * PER_BAND + BX, *tZ = tY + BY;
        const int my0 = __float_as_int(tY[0].w), mz0 = __float_as_int(tZ[0].w);
        const int Ey = ft.ey[b], rows = Ey * ft.ez[b], rows2 = (rows + 1) & ~1, row0 = ft.row0[b];
        const float inv_ey = 1.0f / (float)Ey;
        const bool shared = b >= nbands - NSH;
        for (int r = threadIdx.x; r < rows2; r += NT) {
            const int rr = min(r, rows - 1);
            const int cz = (int)(((float)rr + 0.5f) * inv_ey), cy = rr - cz * Ey;       // exact for rows < 2^20
            const int zoff = tmodf(mz0 + cz, n, pow2) * n;
            if (shared) { s_rowoff[row0 + r] = (zoff + tmodf(my0 + cy, n, pow2)) * pitch; continue; }
#pragma unroll
            for (int qy = 0; qy < RY; ++qy) {
                const int off = (zoff + tmodf(my0 + cy + qy * rep.cys[b], n, pow2)) * pitch;
#pragma unroll
                for (int qx = 0; qx < RX; ++qx) s_rowoff[row0 + (qy * RX + qx) * rows2 + r] = off;
            }
        }
    }
    __syncthreads();
}

// X pass of one band for all its replicas: rows [row0, row0 + R*rows2), replica r = (row - row0) / rows2 reads the tile
// shifted by (r % RX) * cxs cells in x.  Same arithmetic as q4_xpass.
template <int NT, int RX, int RY, bool POW2>
__device__ __forceinline__ void q4_xpass_rep(const float *__restrict__ N, int n, const float4 *tX, float4 *U4,
                                             const int *s_rowoff, int row0, int rows2, int cxs)
{
    constexpr int NW = NT / 32, pow2 = POW2 ? 1 : 0, R = RX * RY;
    static_assert(RX == 1 || RX == 2, "x replicas: 1 or 2");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float4 t0 = tX[lane], t1 = tX[32 + lane], t2 = tX[64 + lane], t3 = tX[96 + lane];
    const int c0 = __float_as_int(t0.w), c1 = __float_as_int(t1.w), c2 = __float_as_int(t2.w), c3 = __float_as_int(t3.w);
    const int cmin = min(min(c0, c1), min(c2, c3)), cmax = max(max(c0, c1), max(c2, c3));
    const bool narrow = __all_sync(0xffffffffu, cmax - cmin <= 1);
    const int half = rows2 >> 1, pairs = half * R;
    if (narrow) {
        const float *base0 = N + tmodf(cmin, n, pow2);
        const float *base1 = N + tmodf(cmin + cxs, n, pow2);
        const bool s0 = c0 != cmin, s1 = c1 != cmin, s2 = c2 != cmin, s3 = c3 != cmin;
        const float a00 = s0 ? 0.0f : t0.x, a01 = s0 ? t0.x : t0.y, a02 = s0 ? t0.y : t0.z, a03 = s0 ? t0.z : 0.0f;
        const float a10 = s1 ? 0.0f : t1.x, a11 = s1 ? t1.x : t1.y, a12 = s1 ? t1.y : t1.z, a13 = s1 ? t1.z : 0.0f;
        const float a20 = s2 ? 0.0f : t2.x, a21 = s2 ? t2.x : t2.y, a22 = s2 ? t2.y : t2.z, a23 = s2 ? t2.z : 0.0f;
        const float a30 = s3 ? 0.0f : t3.x, a31 = s3 ? t3.x : t3.y, a32 = s3 ? t3.y : t3.z, a33 = s3 ? t3.z : 0.0f;
        // four rows (two pairs) per iteration: 16 independent tile loads in flight per lane
        for (int p0 = warp; p0 < pairs; p0 += 2 * NW) {
            float vv[4][4];
            int rw[4];
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const int p = min(p0 + g * NW, pairs - 1);     // the tail repeats the last pair (same values rewritten)
                const int rr = (p >= half) + (R > 2 ? (p >= 2 * half) + (p >= 3 * half) : 0);
                const float *base = (RX == 2 && (rr & 1)) ? base1 : base0;
                const int r = row0 + 2 * p;
                const int2 o = *reinterpret_cast<const int2 *>(s_rowoff + r);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float *q = base + (unsigned)(h ? o.y : o.x);
                    rw[2 * g + h] = r + h;
#pragma unroll
                    for (int e = 0; e < 4; ++e) vv[2 * g + h][e] = __ldg(q + e);
                }
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const float v0 = vv[g][0], v1 = vv[g][1], v2 = vv[g][2], v3 = vv[g][3];
                float4 u;
                u.x = fmaf(a03, v3, fmaf(a02, v2, fmaf(a01, v1, a00 * v0)));
                u.y = fmaf(a13, v3, fmaf(a12, v2, fmaf(a11, v1, a10 * v0)));
                u.z = fmaf(a23, v3, fmaf(a22, v2, fmaf(a21, v1, a20 * v0)));
                u.w = fmaf(a33, v3, fmaf(a32, v2, fmaf(a31, v1, a30 * v0)));
                U4[rw[g] * 32 + lane] = u;
            }
        }
    } else {
        for (int p = warp; p < pairs; p += NW) {
            const int rr = (p >= half) + (R > 2 ? (p >= 2 * half) + (p >= 3 * half) : 0);
            const int sh = (RX == 2 && (rr & 1)) ? cxs : 0;
            const float *b0 = N + tmodf(c0 + sh, n, pow2), *b1 = N + tmodf(c1 + sh, n, pow2);
            const float *b2 = N + tmodf(c2 + sh, n, pow2), *b3 = N + tmodf(c3 + sh, n, pow2);
            const int r = row0 + 2 * p;
            const int2 o = *reinterpret_cast<const int2 *>(s_rowoff + r);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const unsigned off = (unsigned)(h ? o.y : o.x);
                const float *q0 = b0 + off, *q1 = b1 + off, *q2 = b2 + off, *q3 = b3 + off;
                float4 u;
                u.x = fmaf(t0.z, __ldg(q0 + 2), fmaf(t0.y, __ldg(q0 + 1), t0.x * __ldg(q0)));
                u.y = fmaf(t1.z, __ldg(q1 + 2), fmaf(t1.y, __ldg(q1 + 1), t1.x * __ldg(q1)));
                u.z = fmaf(t2.z, __ldg(q2 + 2), fmaf(t2.y, __ldg(q2 + 1), t2.x * __ldg(q2)));
                u.w = fmaf(t3.z, __ldg(q3 + 2), fmaf(t3.y, __ldg(q3 + 1), t3.x * __ldg(q3)));
                U4[(r + h) * 32 + lane] = u;
            }
        }
    }
}

constexpr int rep_min_blocks(int windows) { return windows <= 2 ? 3 : 2; }

// NB direct bands; the last NSH of them are shared by the replicas (see q4_footprints_rep), the others have one window
// per replica
template <int NB, int NSH, int RX, int RY, int YPW, int RINGMODE, int BY, bool POW2>
__global__ void __launch_bounds__(32 * BY, BY == 8 ? rep_min_blocks((NB - NSH) * RX * RY + NSH) : 1)
k_mb3d_rep(const float *__restrict__ N, int n, WnTabs tabs, int nx, int ny, int nk, int max_rows, WnFold fold,
           WnOrder ord, WnRep rep, const __grid_constant__ CUtensorMap pmap, float *__restrict__ out)
{
    constexpr int BX = 128, BZ = 32, NT = 32 * BY, R = RX * RY;
    // period-block planes in flight (two bands in 8-row bricks: U needs the room for two CTAs per SM)
    constexpr int RING = (NB == 1 || BY == 16) ? 8 : 4;
    constexpr int HP = RING / 2;                                // RINGMODE 1: planes per TMA box = half a ring
    constexpr int PER_BAND = BX + BY + BZ;
    constexpr int LPR = 32 / YPW;                               // lanes per y row of a warp
    static_assert(YPW == 1 || YPW == 2 || YPW == 4, "y rows per warp");
    static_assert(NB >= 1 && NB <= 2 && NSH >= 0 && NSH < NB && R <= 4, "windows live in registers");
    constexpr int NR = NB - NSH;                                 // bands [0, NR) have a window per replica
    const bool folded = fold.P != nullptr;
    // dynamic shared memory (float4 units): [ring RING x NT, only when folded] | U[max_rows][32] | tables NB x PER_BAND |
    // rowoff[max_rows] (int) | mbarriers
    extern __shared__ __align__(128) float4 smem_rep[];
    float4 *ringbuf = smem_rep;
    float4 *U4 = smem_rep + (folded ? RING * NT : 0);
    float4 *s_tab = U4 + max_rows * 32;
    int *s_rowoff = reinterpret_cast<int *>(s_tab + NB * PER_BAND);
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(s_rowoff + max_rows);   // full[2], empty[2]

    chain_wait();                                              // the ring prefetch below reads the previous kernel's output
    chain_release();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int yb, zb;
    {
        const int gy = blockIdx.y, rz = gy % ord.zrep, t = gy / ord.zrep;
        yb = t / ord.yrep + (t % ord.yrep) * ord.yper;
        zb = rz * ord.zper + blockIdx.z;
    }
    const int i0 = blockIdx.x * BX, j0 = yb * BY, k0 = zb * BZ;
    const int xq = (warp % YPW) * LPR + (lane % LPR);            // float4 column of this thread in the 128-sample row
    const int jl = (warp / YPW) * YPW + lane / LPR;              // y row of this thread in the brick
    const int i = i0 + 4 * xq, j = j0 + jl;
    const bool active = i < rep.bx && j < rep.by;
    const int kmax = min(BZ, nk - k0);

    // ---- period-block prefetch (before the tables and the X pass)
    const size_t pplane = (size_t)fold.Lx * fold.Ly;
    const float *pcol = nullptr, *pp = nullptr;
    int kk = 0;
    float4 *ring = ringbuf + threadIdx.x;                        // RINGMODE 0: this thread's private slots
    int tma_x = 0, tma_y = 0, tma_z = 0;
    if (folded) {
        kk = fold.kphase + k0;
        if (kk >= fold.Lz) kk %= fold.Lz;
        if constexpr (RINGMODE == 1) {
            if (threadIdx.x == 0) {
                tma_x = fold.xmask >= 0 ? (i0 & fold.xmask) : i0 % fold.Lx;
                tma_y = fold.ymask >= 0 ? (j0 & fold.ymask) : j0 % fold.Ly;
                tma_z = kk;
                mbar_init(bars + 0, 1); mbar_init(bars + 1, 1);
                mbar_init(bars + 2, NT / 32); mbar_init(bars + 3, NT / 32);
                mbar_fence_init();
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    mbar_arrive_expect_tx(bars + h, HP * BY * BX * (unsigned)sizeof(float));
                    tma_load_3d(ringbuf + h * HP * NT, &pmap, tma_x, tma_y, tma_z, bars + h);
                    tma_z += HP;
                    if (tma_z >= fold.Lz) tma_z -= fold.Lz;
                }
            }
        } else if (active) {
            const unsigned pi = (unsigned)(fold.xmask >= 0 ? (i & fold.xmask) : i % fold.Lx);
            const unsigned pj = (unsigned)(fold.ymask >= 0 ? (j & fold.ymask) : j % fold.Ly);
            pcol = fold.P + (pi + pj * (unsigned)fold.Lx);
            pp = pcol + kk * pplane;
#pragma unroll
            for (int d = 0; d < RING - 1; ++d) {
                if (d < kmax) {
                    cp_async16(ring + d * NT, pp);
                    if (++kk == fold.Lz) { kk = 0; pp = pcol; } else pp += pplane;
                }
                cp_async_commit();
            }
        }
    }

    // clamped at the base region: the footprints then match the host's plan (plan_bricks over [0, by))
    q4_load_tables<BY, BZ, NT>(s_tab, tabs, i0, j0, k0, rep.bx, rep.by, nk);
    __shared__ Q4Foot ft;
    q4_footprints_rep<BY, BZ, NT, RX, RY, NSH, POW2>(s_tab, NB, n, ft, s_rowoff, rep);
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        if (b < NR)
            q4_xpass_rep<NT, RX, RY, POW2>(N, n, s_tab + b * PER_BAND, U4, s_rowoff, ft.row0[b],
                                           (ft.row0[b + 1] - ft.row0[b]) / R, rep.cxs[b]);
        else
            q4_xpass_rep<NT, 1, 1, POW2>(N, n, s_tab + b * PER_BAND, U4, s_rowoff, ft.row0[b], ft.row0[b + 1] - ft.row0[b], 0);
    }
    __syncthreads();
    if (RINGMODE == 0 && !active) return;                      // no barrier below (RINGMODE 1 runs on whole bricks only)

    // per band: y weights, per replica the 3-deep window of y-contracted tile-z planes, the next plane to contract
    float4 ty[NB], v[NB][R][3];
    const float4 *unext[NB], *tZ[NB];
    int slab4[NB], rstride[NB], prevw[NB];
    unsigned long long advm[NB];
    const bool slow = ft.slow != 0;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const float4 *tY = s_tab + b * PER_BAND + BX;
        tZ[b] = tY + BY;
        ty[b] = tY[jl];
        slab4[b] = ft.ey[b] * 32;
        rstride[b] = b < NR ? (ft.row0[b + 1] - ft.row0[b]) / R * 32 : 0;
        prevw[b] = __float_as_int(tZ[b][0].w);
        advm[b] = ft.adv[b];
        const float4 *u = U4 + ft.row0[b] * 32 + (__float_as_int(ty[b].w) - __float_as_int(tY[0].w)) * 32 + xq;
#pragma unroll
        for (int r = 0; r < (b < NR ? R : 1); ++r)
#pragma unroll
            for (int f = 0; f < 3; ++f) {
                const float4 *uu = u + r * rstride[b] + f * slab4[b];
                v[b][r][f] = p4_contract3(ty[b].x, ty[b].y, ty[b].z, uu[0], uu[32], uu[64]);
            }
        unext[b] = u + 3 * slab4[b];
    }

    const size_t plane = (size_t)nx * ny;
    float *o = out + ((size_t)i + (size_t)nx * j + plane * k0);
    // one z step: advance the windows, then finish the R samples of this thread (canonical sum: period-block value or
    // 0 first, then the bands from the highest scale down, each band value formed from zero)
    auto zstep_a = [&](int k, float4 (&tz)[NB]) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            tz[b] = tZ[b][k];
            // planes entering the window at this step: from the brick's advance mask (no dependence on the table load
            // above, so the branch does not wait for shared memory); steps of more than 3 cells take the table value
            int adv = (int)(advm[b] >> (2 * k)) & 3;
            if (slow) { adv = __float_as_int(tz[b].w) - prevw[b]; prevw[b] = __float_as_int(tz[b].w); }
#pragma unroll 1
            for (; adv > 0; --adv) {
#pragma unroll
                for (int r = 0; r < (b < NR ? R : 1); ++r) {
                    const float4 *uu = unext[b] + r * rstride[b];
                    const float4 nv = p4_contract3(ty[b].x, ty[b].y, ty[b].z, uu[0], uu[32], uu[64]);
                    v[b][r][0] = v[b][r][1]; v[b][r][1] = v[b][r][2]; v[b][r][2] = nv;
                }
                unext[b] += slab4[b];
            }
        }
    };
    auto zstep_b = [&](const float4 (&tz)[NB], const float4 pv) {
        float4 sh = pv;                                        // shared bands: the same value for every replica
#pragma unroll
        for (int b = NB - 1; b >= NR; --b) sh = p4_add(p4_band_value(tz[b], v[b][0]), sh);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float4 s = sh;
#pragma unroll
            for (int b = NR - 1; b >= 0; --b) s = p4_add(p4_band_value(tz[b], v[b][r]), s);
            __stcs(reinterpret_cast<float4 *>(reinterpret_cast<char *>(o) + rep.roff[r]), s);
        }
        o += plane;
    };
    const float4 zero4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (!folded) {
        for (int k = 0; k < kmax; ++k) {
            float4 tz[NB];
            zstep_a(k, tz);
            zstep_b(tz, zero4);
        }
    } else if constexpr (RINGMODE == 1) {
        const float4 *mine = ringbuf + jl * 32 + xq;              // plane layout [8 y][128 x]
#pragma unroll 1
        for (int g = 0; g < BZ / HP; ++g) {                      // one group = one TMA box = HP planes
            const int h = g & 1;
            const unsigned par = (unsigned)(g >> 1) & 1u;
            bool waited = false;
#pragma unroll
            for (int d = 0; d < HP; ++d) {
                float4 tz[NB];
                zstep_a(g * HP + d, tz);
                if (!waited) { mbar_wait(bars + h, par); waited = true; }
                zstep_b(tz, mine[(h * HP + d) * NT]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + 2 + h);
            if (threadIdx.x == 0 && g + 2 < BZ / HP) {            // refill this half with the planes two groups ahead
                mbar_wait(bars + 2 + h, par);
                mbar_arrive_expect_tx(bars + h, HP * BY * BX * (unsigned)sizeof(float));
                tma_load_3d(ringbuf + h * HP * NT, &pmap, tma_x, tma_y, tma_z, bars + h);
                tma_z += HP;
                if (tma_z >= fold.Lz) tma_z -= fold.Lz;
            }
        }
    } else {
#pragma unroll 4
        for (int k = 0; k < kmax; ++k) {
            float4 tz[NB];
            zstep_a(k, tz);
            if (k + RING - 1 < kmax) {                         // slot of plane k-1, consumed in the previous step
                cp_async16(ring + ((k + RING - 1) % RING) * NT, pp);
                if (++kk == fold.Lz) { kk = 0; pp = pcol; } else pp += pplane;
            }
            cp_async_commit();
            cp_async_wait<RING - 1>();                         // plane k has landed
            zstep_b(tz, ring[(k % RING) * NT]);
        }
    }
}

__global__ void k_pad_tile(const float *__restrict__ N, float *__restrict__ P, int n)
{
    const int pitch = n + WN_TILE_PAD;
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
k_mb3d_gather(WnTileView t, WnTabs tabs, int nx, int ny, int nk, float *__restrict__ out)
{
    chain_wait();
    chain_release();
    const int nbands = tabs.nb;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y, k = blockIdx.z;
    if (i >= nx) return;
    const int n = t.n;
    float acc = 0.0f;
    for (int b = nbands - 1; b >= 0; --b) {                    // canonical summation, highest scale first
        const int row = tabs_row(tabs, b);
        const float4 ax = __ldg(tabs.x + row * tabs.sx + i), ay = __ldg(tabs.y + row * tabs.sy + j),
                     az = __ldg(tabs.z + row * tabs.sz + tabs.kz0 + k);
        const float wx[3] = { ax.x, ax.y, ax.z }, wy[3] = { ay.x, ay.y, ay.z };
        int cx[3], cy[3], cz[3];
#pragma unroll
        for (int f = 0; f < 3; ++f) {
            cx[f] = tmodf(__float_as_int(ax.w) + f, n, t.pow2);
            cy[f] = tmodf(__float_as_int(ay.w) + f, n, t.pow2) * n;
            cz[f] = tmodf(__float_as_int(az.w) + f, n, t.pow2) * n * n;
        }
        float vz[3];
#pragma unroll
        for (int fz = 0; fz < 3; ++fz) {
            float ux[3];
#pragma unroll
            for (int fy = 0; fy < 3; ++fy) {
                const float *row = t.N + cy[fy] + cz[fz];
                ux[fy] = fmaf(wx[2], __ldg(row + cx[2]), fmaf(wx[1], __ldg(row + cx[1]), wx[0] * __ldg(row + cx[0])));
            }
            vz[fz] = fmaf(wy[2], ux[2], fmaf(wy[1], ux[1], wy[0] * ux[0]));     // same order as q4_ycontract
        }
        acc = __fadd_rn(band_value(az, vz[0], vz[1], vz[2]), acc);
    }
    out[(size_t)i + (size_t)nx * ((size_t)j + (size_t)ny * k)] = acc;
}

// ---- host-side tables of one call ------------------------------------------------------------------------------
// Entry of every coordinate of every band, computed once per top-level call with exactly the device formula (this
// TU's host code is built with -ffp-contract=off): the weights (compared bitwise by the period detection), the first
// tap cell reduced mod n, and the unwrapped first tap cell (brick footprints).  Period blocks are prefixes of the
// axes, so every nesting level reads the same tables.
struct HostEntry {
    float w0, w1, w2;
    int cell;
    bool operator==(const HostEntry &o) const
    {
        return std::memcmp(&w0, &o.w0, 3 * sizeof(float)) == 0 && cell == o.cell;
    }
};

static_assert(sizeof(HostEntry) == 16, "HostEntry is compared with memcmp");

struct FoldDecision { int nfold; bool folded[WN_MAX_BANDS]; long long Lx, Ly, Lz; };

struct HostAxes {
    int nb = 0, nx = 0, ny = 0, nz = 0;
    mutable int runs = 0;           // lattice kernels launched so far in this call (the first one follows k_axis_tables)
    // Plan cache (one slot per tile object, see wn_mb3d_fast_prepare): a call whose axes, bands and tile edge are
    // bitwise those of the previous call on the same tile reuses these host tables and the fold decisions made on them.
    // Pure host-side planning -- every kernel of the call still runs.
    bool cached = false;            // owned by a tile's cache slot, not by the call
    int key_n = 0;
    WnBands key_b;
    std::vector<float> key_axes;    // xs | ys | zs
    struct Memo { unsigned char rows[WN_MAX_BANDS]; int nbands, nx, ny, nz; long long budget; double level; FoldDecision fd; };
    mutable std::vector<Memo> memo;
    bool matches(const float *xs, int nx_, const float *ys, int ny_, const float *zs, int nz_, const WnBands &b, int n) const
    {
        if (nx_ != nx || ny_ != ny || nz_ != nz || n != key_n || b.nbands != nb) return false;
        if (std::memcmp(b.scale, key_b.scale, sizeof(float) * b.nbands) != 0) return false;
        if (key_axes.size() != (size_t)nx + ny + nz) return false;
        return std::memcmp(key_axes.data(), xs, sizeof(float) * nx) == 0 &&
               std::memcmp(key_axes.data() + nx, ys, sizeof(float) * ny) == 0 &&
               std::memcmp(key_axes.data() + nx + ny, zs, sizeof(float) * nz) == 0;
    }
    std::vector<HostEntry> e;       // [band][x | y | z]
    std::vector<int> first;         // unwrapped first tap cell, same layout
    size_t per_band() const { return (size_t)nx + ny + nz; }
    const HostEntry *ex(int row) const { return e.data() + row * per_band(); }
    const HostEntry *ey(int row) const { return ex(row) + nx; }
    const HostEntry *ez(int row) const { return ey(row) + ny; }
    const int *fx(int row) const { return first.data() + row * per_band(); }
    const int *fy(int row) const { return first.data() + row * per_band() + nx; }
    const int *fz(int row) const { return fy(row) + ny; }
};

// ceilf without the libm call (the table loops below are on the per-call host path): exact for |a| < 2^31, which
// holds for every coordinate the int conversion in the reference (cpp:196) is defined for
inline int ceil_to_int(float a)
{
    if (!(a > -2147483000.0f && a < 2147483000.0f)) return (int)std::ceil(a);
    const int m = (int)a;                                      // truncates toward zero
    return m + ((float)m < a ? 1 : 0);
}

inline void host_entry(float coord, float scale, int n, HostEntry &e, int &first)   // mirrors axis_entry()
{
    const float a = coord * scale - 0.5f;
    const int mid = ceil_to_int(a);
    const float tt = (float)mid - a;
    e.w0 = tt * tt * 0.5f;
    const float s1 = 1.0f - tt;
    e.w2 = s1 * s1 * 0.5f;
    e.w1 = 1.0f - e.w0 - e.w2;
    first = mid - 1;
    if ((n & (n - 1)) == 0) { e.cell = first & (n - 1); return; }
    const int m = first % n;
    e.cell = m < 0 ? m + n : m;
}

// number of leading coordinates of a[0..na) that are bitwise equal to b[0..nb)
inline int same_prefix(const float *a, int na, const float *b, int nb)
{
    const int m = std::min(na, nb);
    if (m > 0 && std::memcmp(a, b, (size_t)m * sizeof(float)) == 0) return m;
    return 0;
}

// The tables are a few hundred KB: a fresh allocation per call would be an mmap / page-fault / munmap round trip
// (about a third of the host cost of a call), so each host thread keeps one spare instance with its storage.
thread_local HostAxes *t_spare_axes = nullptr;
HostAxes *acquire_host_axes()
{
    HostAxes *h = t_spare_axes;
    t_spare_axes = nullptr;
    return h ? h : new HostAxes;
}
void release_host_axes(HostAxes *h)
{
    if (!h || h->cached) return;
    if (!t_spare_axes) t_spare_axes = h; else delete h;
}

HostAxes *make_host_axes(const float *xs, int nx, const float *ys, int ny, const float *zs, int nz, const WnBands &b, int n,
                         void **cache_slot = nullptr)
{
    static const bool cache_on = [] { const char *e = getenv("WN_PLAN_CACHE"); return !e || atoi(e) != 0; }();
    HostAxes *h = nullptr;
    if (cache_slot && cache_on) {
        h = static_cast<HostAxes *>(*cache_slot);
        if (h && h->matches(xs, nx, ys, ny, zs, nz, b, n)) { h->runs = 0; return h; }
        if (!h) { h = new HostAxes; h->cached = true; *cache_slot = h; }
        h->memo.clear();
        h->key_n = n; h->key_b = b;
        h->key_axes.resize((size_t)nx + ny + nz);
        std::copy(xs, xs + nx, h->key_axes.begin());
        std::copy(ys, ys + ny, h->key_axes.begin() + nx);
        std::copy(zs, zs + nz, h->key_axes.begin() + nx + ny);
    } else {
        h = acquire_host_axes();
    }
    h->runs = 0;
    h->nb = b.nbands; h->nx = nx; h->ny = ny; h->nz = nz;
    h->e.resize(h->per_band() * b.nbands);
    h->first.resize(h->per_band() * b.nbands);
    for (int band = 0; band < b.nbands; ++band) {
        HostEntry *e = h->e.data() + band * h->per_band();
        int *f = h->first.data() + band * h->per_band();
        const float s = b.scale[band];
        for (int i = 0; i < nx; ++i) host_entry(xs[i], s, n, e[i], f[i]);
        // volumes usually reuse one coordinate array for several axes: copy instead of recomputing (the entries of a
        // coordinate do not depend on the axis it is used for)
        const int ny_same = same_prefix(ys, ny, xs, nx), nz_same_x = same_prefix(zs, nz, xs, nx);
        std::copy(e, e + ny_same, e + nx);
        std::copy(f, f + ny_same, f + nx);
        for (int i = ny_same; i < ny; ++i) host_entry(ys[i], s, n, e[nx + i], f[nx + i]);
        std::copy(e, e + nz_same_x, e + nx + ny);
        std::copy(f, f + nz_same_x, f + nx + ny);
        for (int i = nz_same_x; i < nz; ++i) host_entry(zs[i], s, n, e[nx + ny + i], f[nx + ny + i]);
    }
    return h;
}

// brick plan of a lattice window (y in [0, ny), z in [k0, k0 + nk)) for the bands with table rows rows[0..nb):
// monotonicity of the y/z tap cells and the largest Ey*Ez any brick needs
struct BrickPlan { bool ok; size_t smem; int max_rows; };

BrickPlan plan_bricks(const HostAxes &h, const unsigned char *rows, int nb, int ny, int k0, int nk, int BY, int BZ,
                      int BX = 32, bool all_bands_resident = false)
{
    BrickPlan p{true, 0, 0};
    for (int band = 0; band < nb; ++band) {
        const int *my = h.fy(rows[band]), *mz = h.fz(rows[band]) + k0;
        for (int j = 1; j < ny; ++j) if (my[j] < my[j - 1]) { p.ok = false; return p; }
        for (int k = 1; k < nk; ++k) if (mz[k] < mz[k - 1]) { p.ok = false; return p; }
        long long ey = 0, ez = 0;
        for (int j0 = 0; j0 < ny; j0 += BY) ey = std::max<long long>(ey, (long long)my[std::min(j0 + BY, ny) - 1] - my[j0] + 3);
        for (int kb = 0; kb < nk; kb += BZ) ez = std::max<long long>(ez, (long long)mz[std::min(kb + BZ, nk) - 1] - mz[kb] + 3);
        if (ey * ez * BX > 1500 * 32) { p.ok = false; return p; }       // U must stay below ~192 KB
        const int rows_b = (int)((ey * ez + 3) & ~3LL);
        p.max_rows = all_bands_resident ? p.max_rows + rows_b : std::max(p.max_rows, rows_b);
    }
    if ((long long)p.max_rows * BX > 1500 * 32) { p.ok = false; return p; }
    p.smem = (size_t)p.max_rows * (BX * sizeof(float) + sizeof(int)) + (size_t)nb * (BX + BY + BZ) * sizeof(float4);
    return p;
}

// raises a kernel's dynamic shared memory limit when a launch needs more than it has been granted so far (48 KB by
// default); remembered per kernel instantiation and device, so the driver call happens once, not per launch
template <typename K>
bool allow_smem(K kern, size_t smem)
{
    if (smem <= 48 * 1024) return true;
    static std::mutex mu;
    static std::map<std::pair<const void *, int>, size_t> granted;     // (kernel, device ordinal) -> bytes
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    std::lock_guard<std::mutex> lock(mu);
    size_t &g = granted[std::make_pair(reinterpret_cast<const void *>(kern), dev)];
    if (smem <= g) return true;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
    g = smem;
    return true;
}

// launch with the programmatic stream serialisation attribute (WN_PDL=0: plain stream order, for A/B runs)
template <typename... KArgs, typename... Args>
cudaError_t launch_chained(bool want_pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                           Args... args)
{
    static const bool pdl_env = [] { const char *e = getenv("WN_PDL"); return !e || atoi(e) != 0; }();
    const bool pdl = want_pdl && pdl_env;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

template <int BY, int BZ, int NT>
int launch_brick(WnTileView t, const WnTabs &tabs, int nx, int ny, int nk,
                 float *out, const BrickPlan &plan, WnFold fold, cudaStream_t st)
{
    auto kern = t.pow2 ? k_mb3d_brick<BY, BZ, NT, true> : k_mb3d_brick<BY, BZ, NT, false>;
    const size_t smem = plan.smem;
    if (!allow_smem(kern, smem)) return -1;
    dim3 grid((nx + 31) / 32, (ny + BY - 1) / BY, (nk + BZ - 1) / BZ);
    launch_chained(tabs.pdl != 0, kern, grid, dim3(NT), smem, st, t.Npad, t.n, tabs, nx, ny, nk, plan.max_rows, fold, out);
    return 1;
}

template <int BY, int BZ, int NT>
int launch_brick4(WnTileView t, const WnTabs &tabs, int nx, int ny, int nk,
                  float *out, const BrickPlan &plan, WnFold fold, cudaStream_t st)
{
    auto kern = t.pow2 ? k_mb3d_brick4<BY, BZ, NT, true> : k_mb3d_brick4<BY, BZ, NT, false>;
    const size_t smem = plan.smem;
    if (!allow_smem(kern, smem)) return -1;
    dim3 grid((nx + 127) / 128, (ny + BY - 1) / BY, (nk + BZ - 1) / BZ);
    launch_chained(tabs.pdl != 0, kern, grid, dim3(NT), smem, st, t.Npad, t.n, tabs, nx, ny, nk, plan.max_rows, fold, out);
    return 1;
}

template <int NB, int RING>
int launch_col4(WnTileView t, const WnTabs &tabs, int nx, int ny, int nk,
                float *out, const BrickPlan &plan, WnFold fold, cudaStream_t st)
{
    constexpr int BY = 8, BZ = 32, NT = 256;
    auto kern = t.pow2 ? k_mb3d_col4<NB, BY, BZ, NT, RING, true> : k_mb3d_col4<NB, BY, BZ, NT, RING, false>;
    const size_t smem = plan.smem + (fold.P && RING > 0 ? (size_t)RING * NT * sizeof(float4) : 0);   // + the period-block ring
    if (!allow_smem(kern, smem)) return -1;
    const int nyb = (ny + BY - 1) / BY, nzb = (nk + BZ - 1) / BZ;
    WnOrder ord{1, nyb, 1, nzb};
    // replica-first order when the period block cannot stay in L2 by itself (config 3, 512 MiB block: 1.08 vs 1.17 ms
    // per 1024^3); WN_REPLICA_ORDER=0/1 forces it for A/B runs
    bool replica = fold.P && (long long)fold.Lx * fold.Ly * fold.Lz > (8LL << 20);
    if (const char *e = getenv("WN_REPLICA_ORDER")) replica = fold.P && atoi(e) != 0;
    if (replica) {
        if (fold.Ly % BY == 0 && ny % fold.Ly == 0) { ord.yrep = ny / fold.Ly; ord.yper = fold.Ly / BY; }
        if (fold.kphase == 0 && fold.Lz % BZ == 0 && nk % fold.Lz == 0 && (long long)nyb * (nk / fold.Lz) <= 65535) {
            ord.zrep = nk / fold.Lz; ord.zper = fold.Lz / BZ;
        }
    }
    dim3 grid((nx + 127) / 128, nyb * ord.zrep, ord.zper);
    launch_chained(tabs.pdl != 0, kern, grid, dim3(NT), smem, st, t.Npad, t.n, tabs, nx, ny, nk, plan.max_rows, fold, ord, out);
    return 1;
}

// ---- replica kernel (k_mb3d_rep): host side ----------------------------------------------------------------------
// Two halves of an axis are replicas of each other for a band when entry i + S carries bit-identical weights and a
// first tap cell that differs from entry i's by one constant, for every i in [0, S) (S = len / 2).
bool axis_replica_shift(const HostEntry *e, const int *first, int len, int *shift)
{
    if (len < 2 || (len & 1)) return false;
    const int S = len / 2, c = first[S] - first[0];
    for (int i = 0; i < S; ++i)
        if (std::memcmp(&e[i].w0, &e[i + S].w0, 3 * sizeof(float)) != 0 || first[i + S] - first[i] != c) return false;
    *shift = c;
    return true;
}

// tensor map of the period block float32[Lz][Ly][Lx] with a 128 x by x planes box (one half of the kernel's ring)
bool make_period_map(const float *P, int Lx, int Ly, int Lz, int by, int planes, CUtensorMap *map)
{
    typedef CUresult (*Encode)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                               const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static const Encode encode = [] {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            fn = nullptr;
        }
        return reinterpret_cast<Encode>(fn);
    }();
    if (!encode) return false;
    const cuuint64_t dims[3] = { (cuuint64_t)Lx, (cuuint64_t)Ly, (cuuint64_t)Lz };
    const cuuint64_t strides[2] = { (cuuint64_t)Lx * sizeof(float), (cuuint64_t)Lx * Ly * sizeof(float) };
    const cuuint32_t box[3] = { 128, (cuuint32_t)by, (cuuint32_t)planes }, estr[3] = { 1, 1, 1 };
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(P), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int NB, int NSH, int RX, int RY, int YPW, int RINGMODE, int BY>
int launch_rep(WnTileView t, const WnTabs &tabs, int nx, int ny, int nk, float *out, int max_rows, size_t smem, WnFold fold,
               const WnRep &rep, const CUtensorMap &pmap, cudaStream_t st)
{
    constexpr int BZ = 32, NT = 32 * BY;
    auto kern = t.pow2 ? k_mb3d_rep<NB, NSH, RX, RY, YPW, RINGMODE, BY, true>
                       : k_mb3d_rep<NB, NSH, RX, RY, YPW, RINGMODE, BY, false>;
    if (!allow_smem(kern, smem)) return -1;
    const int nyb = (rep.by + BY - 1) / BY, nzb = (nk + BZ - 1) / BZ;
    WnOrder ord{1, nyb, 1, nzb};
    bool replica = fold.P && (long long)fold.Lx * fold.Ly * fold.Lz > (8LL << 20);
    if (const char *e = getenv("WN_REPLICA_ORDER")) replica = fold.P && atoi(e) != 0;
    if (replica) {
        if (fold.Ly % BY == 0 && rep.by % fold.Ly == 0) { ord.yrep = rep.by / fold.Ly; ord.yper = fold.Ly / BY; }
        if (fold.kphase == 0 && fold.Lz % BZ == 0 && nk % fold.Lz == 0 && (long long)nyb * (nk / fold.Lz) <= 65535) {
            ord.zrep = nk / fold.Lz; ord.zper = fold.Lz / BZ;
        }
    }
    dim3 grid((rep.bx + 127) / 128, nyb * ord.zrep, ord.zper);
    launch_chained(tabs.pdl != 0, kern, grid, dim3(NT), smem, st, t.Npad, t.n, tabs, nx, ny, nk, max_rows, fold, ord, rep, pmap, out);
    return 1;
}

// Replica kernel for a lattice window with one or two direct bands.  Returns kernels launched, 0 when the window does
// not qualify (the caller falls back to k_mb3d_col4), -1 on a launch error.  dry: launch nothing, return 1 where a
// real call would launch.
// Knobs (A/B runs): WN_REP=0 off, "11" / "12" / "22" = at most that many replicas along x and y (default "22");
// WN_REP_YPW=1|4 thread mapping (NB = 1 only); WN_REP_TMA=0|1 period-block ring; WN_REP_SHARE=0: no shared bands;
// WN_REP_BY=16: 128 x 16 x 32 bricks, 512 threads, one CTA per SM with an 8-plane ring (measured 0.843 ms against
// 0.830 ms for the default two 8-row CTAs per SM on config 3: fewer X-pass rows, but the prologue is no longer hidden
// behind a second CTA).
int rep_pass(WnTileView t, const WnTabs &tabs, const HostAxes &h, const unsigned char *rows, int k0, int nx, int ny, int nk,
             const WnBands &b, WnFold fold, float *out, cudaStream_t st, bool dry = false)
{
    const int nb = b.nbands;
    if (nb < 1 || nb > 2 || nx % 4 != 0 || (fold.P && fold.Lx % 4 != 0)) return 0;
    if (ny > 8 * 65535 || nk > 32 * 65535) return 0;
    int want_rx = 2, want_ry = 2, share = 1;
    if (const char *e = getenv("WN_REP")) {
        if (!strcmp(e, "0")) return 0;
        if (strlen(e) == 2) { want_rx = e[0] - '0'; want_ry = e[1] - '0'; }
    }
    if (const char *e = getenv("WN_REP_SHARE")) share = atoi(e) != 0;
    // which axes can be halved into replicas, and the first-tap-cell shift of every band between the halves
    int cxs[2] = { 0, 0 }, cys[2] = { 0, 0 };
    bool ok_x = want_rx == 2 && nx % 8 == 0 && (!fold.P || (nx / 2) % fold.Lx == 0);
    bool ok_y = want_ry == 2 && ny % 2 == 0 && (!fold.P || (ny / 2) % fold.Ly == 0);
    for (int i = 0; i < nb && ok_x; ++i) ok_x = axis_replica_shift(h.ex(rows[i]), h.fx(rows[i]), nx, &cxs[i]);
    for (int i = 0; i < nb && ok_y; ++i) ok_y = axis_replica_shift(h.ey(rows[i]), h.fy(rows[i]), ny, &cys[i]);
    int want_by = 8;                                           // brick rows (= warps per CTA); 16: one CTA per SM
    if (const char *e = getenv("WN_REP_BY")) want_by = atoi(e) == 16 ? 16 : 8;
    // candidates in order of preference; the first one whose windows fit the registers and whose U rows leave room for
    // two CTAs per SM (<= ~113 KB each; 16-row bricks: one CTA, <= ~226 KB) is used
    const int cand[3][2] = { {2, 2}, {1, 2}, {1, 1} };
    for (int ci = 0; ci < 3; ++ci) {
        const int RX = cand[ci][0], RY = cand[ci][1], R = RX * RY;
        if ((RX == 2 && !ok_x) || (RY == 2 && !ok_y)) continue;
        WnRep rep{nx / RX, ny / RY, {cxs[0], cxs[1]}, {cys[0], cys[1]}, {0, 0, 0, 0}};
        // trailing bands whose replicas coincide (shift = whole tile periods along every halved axis) are shared
        int nsh = 0;
        if (share && R > 1)
            for (int i = nb - 1; i >= 1; --i) {
                const bool same = (RX == 1 || cxs[i] % t.n == 0) && (RY == 1 || cys[i] % t.n == 0);
                if (!same) break;
                ++nsh;
            }
        const int windows = (nb - nsh) * R + nsh;
        if (windows > (nsh ? 5 : 4)) continue;
        int ypw = 1, tma = 1;
        if (const char *e = getenv("WN_REP_YPW")) ypw = (atoi(e) == 4 && nb == 1) ? 4 : 1;
        if (const char *e = getenv("WN_REP_TMA")) tma = atoi(e) != 0;
        // the 16-row brick exists for the TMA ring only; anything else runs the 8-row kernel
        const bool tma_ok16 = tma && ypw == 1 && fold.P && fold.Lx % 128 == 0 && fold.Ly % 16 == 0 && fold.Lz % 4 == 0 &&
                              fold.kphase % 4 == 0 && rep.bx % 128 == 0 && rep.by % 16 == 0 && nk % 32 == 0;
        const int by = (want_by == 16 && tma_ok16) ? 16 : 8;
        const int ring_planes = (nb == 1 || by == 16) ? 8 : 4;     // RING of k_mb3d_rep
        const size_t ring_bytes = fold.P ? (size_t)ring_planes * 32 * by * sizeof(float4) : 0;
        int max_rows = 0;
        bool ok = true;
        for (int i = 0; i < nb && ok; ++i) {
            const BrickPlan pb = plan_bricks(h, rows + i, 1, rep.by, k0, nk, by, 32, 128, true);
            ok = pb.ok;
            max_rows += pb.max_rows * (i < nb - nsh ? R : 1);
        }
        if (!ok) return 0;                                     // unsorted axes / huge footprints: not a brick lattice
        const size_t smem = ring_bytes + (size_t)max_rows * (128 * sizeof(float) + sizeof(int)) +
                            (size_t)nb * (128 + by + 32) * sizeof(float4) + 64;
        if (smem > (size_t)(by == 8 ? 113 : 226) * 1024) continue;
        if (dry) return 1;                                     // the caller only asks whether this window streams
        CUtensorMap pmap;
        std::memset(&pmap, 0, sizeof(pmap));
        tma = tma && fold.P && fold.Lx % 128 == 0 && fold.Ly % by == 0 && fold.Lz % 4 == 0 && fold.kphase % 4 == 0 &&
              rep.bx % 128 == 0 && rep.by % by == 0 && nk % 32 == 0 &&
              make_period_map(fold.P, fold.Lx, fold.Ly, fold.Lz, by, ring_planes / 2, &pmap);
        if (by == 16 && !tma) return -1;                       // the driver entry point vanished between two calls
        for (int r = 0; r < 4; ++r)
            rep.roff[r] = 4LL * ((long long)(r % RX) * rep.bx + (long long)nx * ((long long)(r / RX) * rep.by));
#define WN_REP_CASE(NB_, NSH_, RX_, RY_, YPW_, TMA_, BY_)                                                                \
        if (nb == NB_ && nsh == NSH_ && RX == RX_ && RY == RY_ && ypw == YPW_ && tma == TMA_ && by == BY_)              \
            return launch_rep<NB_, NSH_, RX_, RY_, YPW_, TMA_, BY_>(t, tabs, nx, ny, nk, out, max_rows, smem, fold, rep,   \
                                                                     pmap, st);
#define WN_REP_CASES(NB_, NSH_, RX_, RY_)                                                                               \
        WN_REP_CASE(NB_, NSH_, RX_, RY_, 1, 1, 8) WN_REP_CASE(NB_, NSH_, RX_, RY_, 1, 0, 8)                             \
        WN_REP_CASE(NB_, NSH_, RX_, RY_, 1, 1, 16)
        WN_REP_CASES(1, 0, 1, 1) WN_REP_CASES(1, 0, 1, 2) WN_REP_CASES(1, 0, 2, 2)
        WN_REP_CASE(1, 0, 1, 2, 4, 1, 8) WN_REP_CASE(1, 0, 1, 2, 4, 0, 8)
        WN_REP_CASES(2, 0, 1, 1) WN_REP_CASES(2, 0, 1, 2)
        WN_REP_CASES(2, 1, 1, 2) WN_REP_CASES(2, 1, 2, 2)
#undef WN_REP_CASES
#undef WN_REP_CASE
    }
    return 0;
}

// brick shapes (BY, BZ, threads); WN_BRICK=<index> overrides the default for tuning runs
struct Shape { int by, bz, nt; };
const Shape kShapes[] = { {16, 8, 256}, {16, 16, 256}, {8, 8, 256}, {8, 16, 256}, {8, 4, 256}, {16, 4, 256},
                          /* 6..9: four x-samples per thread (k_mb3d_brick4) */
                          {8, 8, 256}, {16, 8, 256}, {8, 16, 256}, {8, 4, 256} };
const int kFirstShape4 = 6;
const int kNumShapes = (int)(sizeof(kShapes) / sizeof(kShapes[0]));
const int kDefaultShape = 5;

// -1: choose by footprint (larger bricks while several CTAs still fit an SM)
int forced_shape()
{
    if (const char *e = getenv("WN_BRICK")) {
        const int pick = atoi(e);
        if (pick >= 0 && pick < kNumShapes) return pick;
    }
    return -1;
}

// One brick-kernel pass over the lattice window y in [0, ny), z in [k0, k0 + nk) (h: the call's host tables, rows: the
// table rows of the bands in b).
// Returns kernels launched, or -1 when the lattice does not qualify (unsorted y/z, footprint too large, ...).
int brick_pass(WnTileView t, const WnTabs &tabs, const HostAxes &h, const unsigned char *rows, int k0,
               int nx, int ny, int nk, const WnBands &b, WnFold fold, float *out, cudaStream_t st)
{
    int pick = forced_shape();
    BrickPlan plan{false, 0, 0};
    // the float4 kernels need 16-byte aligned rows of the output and of the period block
    const bool can4 = (nx % 4 == 0) && (!fold.P || fold.Lx % 4 == 0);
    // one or two bands left after folding: the z-streaming kernel (WN_COL4=0 disables it for A/B runs)
    const char *col4_env = getenv("WN_COL4");
    const bool col4_on = !col4_env || atoi(col4_env) != 0;
    if (pick < 0 && can4 && col4_on && b.nbands >= 1 && b.nbands <= 2) {
        const int r = rep_pass(t, tabs, h, rows, k0, nx, ny, nk, b, fold, out, st);
        if (r != 0) return r;
    }
    if (pick < 0 && can4 && col4_on && b.nbands >= 1 && b.nbands <= 2 && (ny + 7) / 8 <= 65535 && (nk + 31) / 32 <= 65535) {
        plan = plan_bricks(h, rows, b.nbands, ny, k0, nk, 8, 32, 128, true);
        if (plan.ok && plan.smem <= 56 * 1024) {
            // ring of 8 period-block planes per thread (32 KB) while four CTAs still fit an SM, else 4 planes;
            // WN_RING=-2/-4: register look-ahead instead (A/B runs)
            int ring = plan.smem <= 24 * 1024 ? 8 : 4;
            if (const char *e = getenv("WN_RING")) ring = atoi(e);
#define WN_COL4_CASE(nb, rg) if (b.nbands == nb && ring == rg) return launch_col4<nb, rg>(t, tabs, nx, ny, nk, out, plan, fold, st);
            WN_COL4_CASE(1, 8) WN_COL4_CASE(1, 4) WN_COL4_CASE(1, -2) WN_COL4_CASE(1, -4)
            WN_COL4_CASE(2, 8) WN_COL4_CASE(2, 4) WN_COL4_CASE(2, -2) WN_COL4_CASE(2, -4)
#undef WN_COL4_CASE
            return b.nbands == 1 ? launch_col4<1, 8>(t, tabs, nx, ny, nk, out, plan, fold, st)
                                 : launch_col4<2, 8>(t, tabs, nx, ny, nk, out, plan, fold, st);
        }
    }
    // Three or more direct bands: the brick kernels below pay for every band's whole y-z footprint, while the lowest one
    // or two bands usually still stream through k_mb3d_rep.  The canonical sum takes the bands from the highest scale
    // down, so the upper bands can be evaluated first (recursively, same rule) and the lowest ones added on top IN PLACE:
    // the second pass reads `out` as its period block (period = the window itself; a CTA's ring only ever reads planes
    // of its own column ahead of the plane it writes).  Same operations in the same order: bit-identical to one pass.
    // WN_SPLIT=0 turns it off (A/B runs).
    if (pick < 0 && can4 && col4_on && b.nbands >= 3 && (ny + 7) / 8 <= 65535 && (nk + 31) / 32 <= 65535) {
        const char *split_env = getenv("WN_SPLIT");
        if (!split_env || atoi(split_env) != 0) {
            const WnFold inplace = make_fold(out, nx, ny, nk, 0);
            for (int m = 2; m >= 1; --m) {
                WnBands low = b, up = b;
                low.nbands = m;
                up.nbands = b.nbands - m;
                for (int i = 0; i < up.nbands; ++i) { up.scale[i] = b.scale[i + m]; up.weight[i] = b.weight[i + m]; }
                if (rep_pass(t, tabs, h, rows, k0, nx, ny, nk, low, inplace, out, st, true) != 1) continue;
                WnTabs uptabs = tabs, lowtabs = tabs;
                uptabs.rowbits = tabs.rowbits >> (4 * m);
                uptabs.nb = up.nbands;
                lowtabs.nb = m;
                const int r1 = brick_pass(t, uptabs, h, rows + m, k0, nx, ny, nk, up, fold, out, st);
                if (r1 < 0) break;                             // nothing launched: go on with the unsplit kernels
                const int r2 = rep_pass(t, lowtabs, h, rows, k0, nx, ny, nk, low, inplace, out, st);
                return r2 > 0 ? r1 + r2 : -1;                  // -1: the caller recomputes the window by direct gathers
            }
        }
    }
    if (pick >= kFirstShape4 && !can4) pick = -1;
    if (pick >= 0) {                                           // forced shape (tuning): only where it fits
        plan = plan_bricks(h, rows, b.nbands, ny, k0, nk, kShapes[pick].by, kShapes[pick].bz, pick >= kFirstShape4 ? 128 : 32,
                           pick >= kFirstShape4);
        if (!plan.ok || plan.smem > 100 * 1024) pick = -1;
    }
    if (pick < 0) {
        // WN_BRICK_FALLBACK=<index>: shape of the passes the streaming kernels above did not take (A/B runs)
        if (const char *e = getenv("WN_BRICK_FALLBACK")) {
            const int p = atoi(e);
            if (p >= 0 && p < kNumShapes && (p < kFirstShape4 || can4)) {
                plan = plan_bricks(h, rows, b.nbands, ny, k0, nk, kShapes[p].by, kShapes[p].bz, p >= kFirstShape4 ? 128 : 32,
                                   p >= kFirstShape4);
                if (plan.ok && plan.smem <= 100 * 1024) pick = p;
            }
        }
    }
    if (pick < 0) {
        if (can4) {
            pick = kFirstShape4;                               // 128 x 8 x 8 samples, small footprints only
            plan = plan_bricks(h, rows, b.nbands, ny, k0, nk, kShapes[pick].by, kShapes[pick].bz, 128, true);
        }
        if (!plan.ok || plan.smem > 72 * 1024) {
            pick = 0;                                          // 32 x 16 x 8 samples
            plan = plan_bricks(h, rows, b.nbands, ny, k0, nk, kShapes[0].by, kShapes[0].bz);
        }
        if (!plan.ok || plan.smem > 56 * 1024) {               // big footprints: halve the brick in z
            pick = kDefaultShape;
            plan = plan_bricks(h, rows, b.nbands, ny, k0, nk, kShapes[pick].by, kShapes[pick].bz);
        }
    }
    const int BY = kShapes[pick].by, BZ = kShapes[pick].bz;
    const bool grid_ok = (ny + BY - 1) / BY <= 65535 && (nk + BZ - 1) / BZ <= 65535;
    if (!plan.ok || !grid_ok || plan.smem > 200 * 1024) return -1;
    int r = -1;
#define WN_BRICK_CASE(idx, by, bz, nt) \
    case idx: r = launch_brick<by, bz, nt>(t, tabs, nx, ny, nk, out, plan, fold, st); break;
    switch (pick) {
        WN_BRICK_CASE(0, 16, 8, 256)  WN_BRICK_CASE(1, 16, 16, 256) WN_BRICK_CASE(2, 8, 8, 256)
        WN_BRICK_CASE(3, 8, 16, 256)  WN_BRICK_CASE(4, 8, 4, 256)   WN_BRICK_CASE(5, 16, 4, 256)
#define WN_BRICK4_CASE(idx, by, bz, nt) \
    case idx: r = launch_brick4<by, bz, nt>(t, tabs, nx, ny, nk, out, plan, fold, st); break;
        WN_BRICK4_CASE(6, 8, 8, 256)  WN_BRICK4_CASE(7, 16, 8, 256) WN_BRICK4_CASE(8, 8, 16, 256)
        WN_BRICK4_CASE(9, 8, 4, 256)
#undef WN_BRICK4_CASE
    }
#undef WN_BRICK_CASE
    return r < 0 ? -1 : r;
}

// ---- periodic folding ---------------------------------------------------------------------------------------
// The tile is periodic (n cells), so whenever a band's sample lattice is commensurate with it -- coordinate i+P
// lands on the same cell (mod n) with bit-identical weights as coordinate i -- that band's contribution repeats
// with period P along the axis.  Such bands are evaluated once on their common period block (Lx x Ly x Lz samples,
// by the same kernels) and their block value, read by index mod period, is the start of every sample's running sum
// (canonical summation, see band_value).  Detection is exact (bitwise on the table entries the device will compute),
// so arbitrary coordinates simply do not fold.
bool axis_replica_shift(const HostEntry *e, const int *first, int len, int *shift);

// How the replica kernel (k_mb3d_rep) can treat a band on an nx x ny lattice: 0 = one window per sample column,
// 1 = the x and y halves of the lattice are replicas (same weights, constant cell shift: four samples share the
// period-block value), 2 = replicas that coincide (shift = whole tile periods: the band itself is shared by the four).
int band_replica_class(const HostAxes &hax, int row, int nx, int ny)
{
    int sx = 0, sy = 0;
    if (nx % 8 != 0 || ny % 2 != 0) return 0;
    if (!axis_replica_shift(hax.ex(row), hax.fx(row), nx, &sx)) return 0;
    if (!axis_replica_shift(hax.ey(row), hax.fy(row), ny, &sy)) return 0;
    const bool same = hax.ex(row)[nx / 2].cell == hax.ex(row)[0].cell && hax.ey(row)[ny / 2].cell == hax.ey(row)[0].cell;
    return same ? 2 : 1;
}

// smallest P <= len/2 with entry[i+P] == entry[i] for all i; len when the axis is not periodic
int axis_period(const HostEntry *e, int len)
{
    if (len < 4) return len;
    for (int P = 1; P <= len / 2; ++P) {
        if (!(e[P] == e[0])) continue;
        if (std::memcmp(e + P, e, (size_t)(len - P) * sizeof(HostEntry)) == 0) return P;   // entries are 16 packed bytes
    }
    return len;
}

long long lcm_capped(long long a, long long b, long long cap)
{
    long long x = a, y = b;
    while (y) { long long r = x % y; x = y; y = r; }
    const long long l = a / x * b;
    return l > cap ? cap + 1 : l;
}

// ---- fold decision (pure host code; also behind wn_debug_fold_plan for the CPU tests) ------------------------------
// b: the bands of this (sub)lattice in canonical order, rows[i] = table row of band i; the lattice is the prefix
// nx x ny x nz of the call's axes
FoldDecision decide_fold_uncached(const HostAxes &hax, const unsigned char *rows, const WnBands &b, const float *h_xs,
                                  const float *h_ys, const float *h_zs, int nx, int ny, int nz, long long budget,
                                  double level_overhead);

// the decision depends on the host tables, the band subset, the sub-lattice and the two tuning knobs: memoised on the
// (cached) tables
FoldDecision decide_fold(const HostAxes &hax, const unsigned char *rows, const WnBands &b, const float *h_xs,
                         const float *h_ys, const float *h_zs, int nx, int ny, int nz)
{
    long long budget = 1LL << 27;                              // samples in the period block: 512 MiB of scratch at most
    if (const char *e = getenv("WN_FOLD_BUDGET")) budget = atoll(e);
    double level_overhead = 8e6;
    if (const char *e = getenv("WN_FOLD_LEVEL_COST")) level_overhead = atof(e);   // tests fold tiny lattices with 0
    if (hax.cached)
        for (const HostAxes::Memo &m : hax.memo)
            if (m.nbands == b.nbands && m.nx == nx && m.ny == ny && m.nz == nz && m.budget == budget && m.level == level_overhead &&
                std::memcmp(m.rows, rows, b.nbands) == 0)
                return m.fd;
    const FoldDecision fd = decide_fold_uncached(hax, rows, b, h_xs, h_ys, h_zs, nx, ny, nz, budget, level_overhead);
    if (hax.cached) {
        HostAxes::Memo m;
        std::memset(&m, 0, sizeof(m));
        std::memcpy(m.rows, rows, b.nbands);
        m.nbands = b.nbands; m.nx = nx; m.ny = ny; m.nz = nz; m.budget = budget; m.level = level_overhead; m.fd = fd;
        hax.memo.push_back(m);
    }
    return fd;
}

FoldDecision decide_fold_uncached(const HostAxes &hax, const unsigned char *rows, const WnBands &b, const float *h_xs,
                                  const float *h_ys, const float *h_zs, int nx, int ny, int nz, long long budget,
                                  double level_overhead)
{
    const long long total = (long long)nx * ny * nz;
    struct Cand { int band; int px, py, pz; long long vol; };
    std::vector<Cand> cand;
    // relative per-sample cost of evaluating a band directly, from its step in tile cells per sample (fitted to the
    // measured single-band times: 1 : 1.1 : 1.3 : 2 : 5.4 for steps 1/8 .. 2)
    double cost[WN_MAX_BANDS];
    // typical step of an axis: the median of up to 15 evenly spaced consecutive differences (a block-cyclic z axis jumps
    // at every chunk boundary; a single probe in the middle of the axis would land on such a jump)
    auto median_step = [](const float *a, int len, float scale) {
        if (len < 2) return 0.0;
        double d[15];
        const int probes = std::min(15, len - 1);
        for (int q = 0; q < probes; ++q) {
            const int m = 1 + (int)((long long)(len - 2) * q / std::max(1, probes - 1));
            d[q] = std::fabs((double)a[m] - (double)a[m - 1]);
        }
        std::sort(d, d + probes);
        return d[probes / 2] * (double)scale;
    };
    double urows[WN_MAX_BANDS];                                // U rows of a 128 x 8 x 32 brick of the replica kernel
    for (int i = 0; i < b.nbands; ++i) {
        const double sy = median_step(h_ys, ny, b.scale[i]), sz = median_step(h_zs, nz, b.scale[i]);
        const double st = std::max(median_step(h_xs, nx, b.scale[i]), std::max(sy, sz));
        cost[i] = 1.0 + 1.2 * std::pow(st, 1.6);
        urows[i] = std::ceil(8.0 * sy + 3.0) * std::ceil(32.0 * sz + 3.0);
    }
    // candidates: the folded bands must be a suffix of the canonical order (their block holds the canonical sum of that
    // suffix), so walk down from the highest scale and stop at the first band that does not repeat on this lattice
    if (budget > 0)
        for (int i = b.nbands - 1; i >= 0; --i) {
            const int row = rows[i];
            Cand cd{i, axis_period(hax.ex(row), nx), axis_period(hax.ey(row), ny), axis_period(hax.ez(row), nz), 0};
            cd.vol = (long long)cd.px * cd.py * cd.pz;
            if (cd.vol * 4 > total) break;
            cand.push_back(cd);
        }
    // grow the folded set one band at a time (highest scale first); keep the prefix with the lowest estimated cost.  Period blocks
    // nest (the block is evaluated by this same function), so band i of the prefix is charged on the block it
    // enlarges the fold to, plus the add of the previous level:
    //   cost = (sum_direct c_b + 0.3) * total + sum_{i folded} ((c_i + 0.3) * block_i + level)
    // in units of one band-sample (~1 ps of GPU time); level = the latency of one more small dependent launch (~8 us).
    // Per-sample cost of the bands left direct when the first nd canonical bands (lowest scales) stay direct:
    // the sum of their costs plus 0.3 for adding the period-block value.  The replica kernel changes that for one or two
    // direct bands on a lattice whose halves are replicas: the period-block value (and a coinciding higher band) is
    // fetched / evaluated once per four samples (measured on config 3: bands 4 + 5 direct over a 256^3 block 0.85 ms,
    // band 4 over a 512^3 block 0.98 ms, without sharing 1.04 / 1.11 ms).
    int rclass[WN_MAX_BANDS];
    for (int i = 0; i < b.nbands; ++i) rclass[i] = (i < 2 && nx % 4 == 0) ? band_replica_class(hax, rows[i], nx, ny) : 0;
    auto direct_cost = [&](int nd, bool with_block) {
        double sum = 0.0;
        for (int i = 0; i < nd; ++i) sum += cost[i];
        const double add = with_block ? 0.3 : 0.0;
        // (only while the replicas' U rows fit the kernel's shared memory: rep_pass falls back otherwise)
        if (nd == 1 && rclass[0] >= 1 && 4.0 * urows[0] <= 150.0) return cost[0] + 0.25 * add;
        if (nd == 2 && rclass[0] >= 1 && rclass[1] == 2 && 4.0 * urows[0] + urows[1] <= 180.0) return cost[0] + 0.25 * (cost[1] + add);
        if (nd == 2 && rclass[0] >= 1 && rclass[1] >= 1 && 2.0 * (urows[0] + urows[1]) <= 180.0) return sum + 0.5 * add;
        return sum + add;
    };
    double direct_sum = 0.0;
    for (int i = 0; i < b.nbands; ++i) direct_sum += cost[i];
    double best = direct_sum * (double)total, nested_sum = 0.0;
    long long Lx = 1, Ly = 1, Lz = 1, bLx = 1, bLy = 1, bLz = 1;
    bool folded[WN_MAX_BANDS] = { false }, trial[WN_MAX_BANDS] = { false };
    int nfold = 0, ntrial = 0;
    for (const Cand &cd : cand) {
        const long long lx = lcm_capped(Lx, cd.px, nx), ly = lcm_capped(Ly, cd.py, ny), lz = lcm_capped(Lz, cd.pz, nz);
        if (lx > nx || ly > ny || lz > nz) break;
        if (lx * ly * lz > budget || lx * ly * lz * 4 > total) break;
        Lx = lx; Ly = ly; Lz = lz;
        trial[cd.band] = true;
        ++ntrial;
        nested_sum += (cost[cd.band] + 0.3) * (double)(Lx * Ly * Lz) + level_overhead;
        // a block that does not stay in L2 (> 32 MiB) is written to and read back from HBM: 8 bytes per block sample next
        // to the 4 bytes per sample of the output stream the main kernel is bound by
        if (Lx * Ly * Lz > (1LL << 23)) nested_sum += 2.0 * (double)(Lx * Ly * Lz);
        direct_sum -= cost[cd.band];
        const double est = direct_cost(b.nbands - ntrial, true) * (double)total + nested_sum;
        if (est < best) {
            best = est;
            for (int i = 0; i < b.nbands; ++i) folded[i] = trial[i];
            nfold = ntrial;
            bLx = Lx; bLy = Ly; bLz = Lz;
        }
    }
    FoldDecision fd;
    fd.nfold = nfold;
    for (int i = 0; i < WN_MAX_BANDS; ++i) fd.folded[i] = folded[i];
    fd.Lx = bLx; fd.Ly = bLy; fd.Lz = bLz;
    return fd;
}

// canonical order of the bands (see band_value): ascending scale, ties by index
void canonical_order(const WnBands &b, int order[WN_MAX_BANDS])
{
    for (int i = 0; i < b.nbands; ++i) order[i] = i;
    std::stable_sort(order, order + b.nbands, [&](int l, int r) { return b.scale[l] < b.scale[r]; });
}

} // namespace

// diagnostics: the axis-table entries {w0, w1, w2, first tap cell (unwrapped)} the lattice kernels use for `count`
// coordinates of one axis at one band scale (device pointers), computed by the same kernel as in a real call
int wn_mb3d_debug_axis_table(const float *coords, int count, float scale, float4 *entries, cudaStream_t st)
{
    if (count <= 0) return 0;
    WnLattice c{coords, nullptr, nullptr, count, 0, 0};
    WnBands b;
    std::memset(&b, 0, sizeof(b));
    b.nbands = 1; b.scale[0] = scale; b.weight[0] = 1.0f; b.post = 1.0f;
    k_axis_tables<<<std::min((count + 255) / 256, 1184), 256, 0, st>>>(c, b, 0, 0, entries, nullptr, nullptr);
    return 1;
}

int wn_launch_pad_tile(const float *N, float *Npad, int n, cudaStream_t st)
{
    k_pad_tile<<<148 * 8, 256, 0, st>>>(N, Npad, n);
    return 1;
}

// A fast-lattice call = prepare (fold decision on the WHOLE lattice + evaluation of the period block, once),
// any number of z-slab launches (the chunks of a WN_HOST call), finish.  Deciding per call, not per slab, keeps
// the result of a sample independent of how the call is chunked.
namespace {

WnTabs plan_tabs(const WnFastPlan *plan, const unsigned char *rows, int nb, int kz0)
{
    WnTabs tb;
    tb.x = plan->tab; tb.y = tb.x + (size_t)plan->tab_bands * plan->sx; tb.z = tb.y + (size_t)plan->tab_bands * plan->sy;
    tb.sx = plan->sx; tb.sy = plan->sy; tb.sz = plan->sz;
    tb.kz0 = kz0; tb.nb = nb;
    tb.wait_first = 1;
    tb.pdl = plan->pdl;
    tb.rowbits = 0;
    for (int i = 0; i < nb; ++i) tb.rowbits |= (unsigned long long)(rows[i] & 15) << (4 * i);
    return tb;
}

}  // namespace

// depth 0: `b` are the call's bands and the axis tables are computed here (one launch for every band and the full
// axes); depth > 0 (period blocks): `plan` arrives with the parent's tables and rows[] = the original band indices of b.
int wn_mb3d_fast_prepare(WnTileView t, WnLattice c, const float *h_xs, const float *h_ys, const float *h_zs, WnBands b,
                         WnFastPlan *plan, cudaStream_t st, int depth)
{
    plan->direct = b;
    plan->P = nullptr;
    plan->Lx = plan->Ly = plan->Lz = 1;
    const int nx = c.nx, ny = c.ny, nz = c.nz;
    int launched = 0;
    if (depth == 0) {
        plan->tab = nullptr;
        plan->owns_tab = 0;
        plan->host_axes = nullptr;
        plan->tab_bands = b.nbands; plan->sx = nx; plan->sy = ny; plan->sz = nz;
        for (int i = 0; i < WN_MAX_BANDS; ++i) plan->direct_rows[i] = (unsigned char)i;
        plan->owns_tab = 1;
        plan->host_axes = make_host_axes(h_xs, std::max(nx, 0), h_ys, std::max(ny, 0), h_zs, std::max(nz, 0), b, t.n,
                                         plan->axes_cache);
        if (nx <= 0 || ny <= 0 || nz <= 0 || b.nbands <= 0) return 0;
        const size_t per_band = (size_t)nx + ny + nz;
        if (wn_scratch_alloc((void **)&plan->tab, per_band * b.nbands * sizeof(float4), st) != cudaSuccess) return -1;
        float4 *tx = plan->tab, *ty = tx + (size_t)b.nbands * nx, *tz = ty + (size_t)b.nbands * ny;
        const int total_e = (int)(per_band * b.nbands);
        launch_chained(plan->pdl != 0, k_axis_tables, dim3(std::min((total_e + 255) / 256, 1184)), dim3(256), 0, st, c, b, 0, nz, tx, ty, tz);
        launched = 1;
        // canonical order of the bands (see band_value): ascending scale, ties by index; table rows keep the caller's order
        int order[WN_MAX_BANDS];
        canonical_order(b, order);
        WnBands sorted = b;
        for (int i = 0; i < b.nbands; ++i) {
            sorted.scale[i] = b.scale[order[i]];
            sorted.weight[i] = b.weight[order[i]];
            plan->direct_rows[i] = (unsigned char)order[i];
        }
        b = sorted;
        plan->direct = b;
    }
    if (nx <= 0 || ny <= 0 || nz <= 0) return 0;
    const HostAxes &hax = *static_cast<const HostAxes *>(plan->host_axes);
    const FoldDecision fd = decide_fold(hax, plan->direct_rows, b, h_xs, h_ys, h_zs, nx, ny, nz);
    const bool *folded = fd.folded;
    const int nfold = fd.nfold;
    const long long Lx = fd.Lx, Ly = fd.Ly, Lz = fd.Lz;
    if (nfold == 0) return launched;
    WnBands bf = b, bd = b;
    bf.nbands = bd.nbands = 0;
    unsigned char rows_f[WN_MAX_BANDS] = { 0 }, rows_d[WN_MAX_BANDS] = { 0 };
    for (int i = 0; i < b.nbands; ++i) {
        WnBands &dst = folded[i] ? bf : bd;
        (folded[i] ? rows_f : rows_d)[dst.nbands] = plan->direct_rows[i];
        dst.scale[dst.nbands] = b.scale[i];
        dst.weight[dst.nbands] = b.weight[i];
        ++dst.nbands;
    }
    float *P = nullptr;
    if (wn_scratch_alloc((void **)&P, (size_t)(Lx * Ly * Lz) * sizeof(float), st) != cudaSuccess) {
        cudaGetLastError();                                    // no scratch for the block: evaluate every band per sample
        return launched;
    }
    // The period block is itself a lattice, and the folded bands with the shortest periods repeat inside it: evaluate
    // it with the same machinery (nested folding), e.g. bands 7, 8 of config 3 on 128^3 inside band 6's 256^3 block.
    const WnLattice lc{c.xs, c.ys, c.zs, (int)Lx, (int)Ly, (int)Lz};
    WnFastPlan inner = *plan;                                  // shares the parent's axis tables
    inner.owns_tab = 0;
    for (int i = 0; i < WN_MAX_BANDS; ++i) inner.direct_rows[i] = rows_f[i];
    int max_depth = 3;
    if (const char *e = getenv("WN_FOLD_NEST")) max_depth = atoi(e);
    int r = depth < max_depth ? wn_mb3d_fast_prepare(t, lc, h_xs, h_ys, h_zs, bf, &inner, st, depth + 1) : 0;
    if (r < 0) { cudaFreeAsync(P, st); return -1; }
    if (depth >= max_depth) { inner.direct = bf; inner.P = nullptr; inner.Lx = inner.Ly = inner.Lz = 1; }
    const int r2 = wn_mb3d_fast_run(t, lc, h_ys, h_zs, bf, rows_f, &inner, 0, (int)Lz, P, st);
    wn_mb3d_fast_finish(&inner, st);
    if (r2 < 0) { cudaFreeAsync(P, st); return -1; }
    plan->direct = bd;
    for (int i = 0; i < WN_MAX_BANDS; ++i) plan->direct_rows[i] = rows_d[i];
    plan->P = P; plan->Lx = (int)Lx; plan->Ly = (int)Ly; plan->Lz = (int)Lz;
    return launched + r + r2;
}

// finish without releasing the device scratch: hands the axis tables and the period block to the caller, which frees
// them (cudaFreeAsync) once the kernels that read them are known to be complete
void wn_mb3d_fast_detach(WnFastPlan *plan, void **tab, void **P)
{
    *tab = plan->owns_tab ? plan->tab : nullptr;
    *P = plan->P;
    if (plan->owns_tab) release_host_axes(static_cast<HostAxes *>(plan->host_axes));
    plan->P = nullptr;
    plan->tab = nullptr;
    plan->host_axes = nullptr;
    plan->owns_tab = 0;
}

// Host-only: the top-level fold decision of a FAST lattice call (no device needed).  folded[i] refers to the caller's
// band order; block = period block Lx, Ly, Lz (1, 1, 1 when nothing folds).
int wn_mb3d_fast_plan_host(const float *h_xs, int nx, const float *h_ys, int ny, const float *h_zs, int nz, WnBands b,
                           int tile_n, int *folded, int block[3])
{
    block[0] = block[1] = block[2] = 1;
    for (int i = 0; i < b.nbands; ++i) folded[i] = 0;
    if (nx <= 0 || ny <= 0 || nz <= 0 || b.nbands <= 0 || tile_n < 2) return 0;
    HostAxes *hax = make_host_axes(h_xs, nx, h_ys, ny, h_zs, nz, b, tile_n);
    int order[WN_MAX_BANDS];
    canonical_order(b, order);
    WnBands sorted = b;
    unsigned char rows[WN_MAX_BANDS] = { 0 };
    for (int i = 0; i < b.nbands; ++i) {
        sorted.scale[i] = b.scale[order[i]];
        sorted.weight[i] = b.weight[order[i]];
        rows[i] = (unsigned char)order[i];
    }
    const FoldDecision fd = decide_fold(*hax, rows, sorted, h_xs, h_ys, h_zs, nx, ny, nz);
    release_host_axes(hax);
    if (fd.nfold > 0) {
        for (int i = 0; i < b.nbands; ++i) folded[order[i]] = fd.folded[i] ? 1 : 0;
        block[0] = (int)fd.Lx; block[1] = (int)fd.Ly; block[2] = (int)fd.Lz;
    }
    return fd.nfold;
}

void wn_mb3d_fast_cache_free(void *slot) { delete static_cast<HostAxes *>(slot); }

void wn_mb3d_fast_finish(WnFastPlan *plan, cudaStream_t st)
{
    if (plan->P) cudaFreeAsync(plan->P, st);
    plan->P = nullptr;
    if (plan->owns_tab && plan->tab) cudaFreeAsync(plan->tab, st);
    if (plan->owns_tab) release_host_axes(static_cast<HostAxes *>(plan->host_axes));
    plan->tab = nullptr;
    plan->host_axes = nullptr;
    plan->owns_tab = 0;
}

// all_bands / all_rows: every band of this (sub)lattice with its table row, used only by the generic fallback
int wn_mb3d_fast_run(WnTileView t, WnLattice c, const float *h_ys, const float *h_zs, const WnBands &all_bands,
                     const unsigned char *all_rows, const WnFastPlan *plan, int k0, int nk, float *out, cudaStream_t st)
{
    if (nk <= 0 || c.nx <= 0 || c.ny <= 0) return 0;
    const int nx = c.nx, ny = c.ny;
    (void)h_ys; (void)h_zs;
    const HostAxes &hax = *static_cast<const HostAxes *>(plan->host_axes);
    const WnFold fold = make_fold(plan->P, plan->Lx, plan->Ly, plan->Lz, plan->P ? k0 % plan->Lz : 0);
    WnTabs tabs = plan_tabs(plan, plan->direct_rows, plan->direct.nbands, k0);
    tabs.wait_first = hax.runs++ == 0;
    const int r = brick_pass(t, tabs, hax, plan->direct_rows, k0, nx, ny, nk, plan->direct, fold, out, st);
    if (r >= 0) return r;

    // ---- generic path (unsorted y/z axes or huge steps): every band by direct gathers, no folding
    if (ny > 65535 || nk > 65535) return -1;                 // lattices that large are rejected
    unsigned char ident[WN_MAX_BANDS];
    for (int i = 0; i < WN_MAX_BANDS; ++i) ident[i] = (unsigned char)i;
    const WnTabs all = plan_tabs(plan, all_rows ? all_rows : ident, all_bands.nbands, k0);
    dim3 grid((nx + 255) / 256, ny, nk);
    launch_chained(all.pdl != 0, k_mb3d_gather, grid, dim3(256), 0, st, t, all, nx, ny, nk, out);
    return 1;
}
