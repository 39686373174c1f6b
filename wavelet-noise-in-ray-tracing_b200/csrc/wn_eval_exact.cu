// wn_eval_exact.cu -- reference-order evaluators (kernels K5..K10 of DESIGN.md), one sample per thread.
//
// Everything here reproduces the reference's floating-point operation ORDER and WIDTH with explicit
// round-to-nearest intrinsics (__fmul_rn / __fadd_rn / __dmul_rn ...), which nvcc never contracts into
// FMAs; the file is additionally built with -fmad=false.  ceilf/floorf/sqrtf/fabsf and float<->int
// conversions are exact IEEE operations on both sides, so for in-range inputs the results are
// BIT-IDENTICAL to the CPU reference (checked against the 15 shipped .raw images).
//
//   evaluate2D            WaveletNoise.cpp:111-140      eval2d()
//   evaluate3D            WaveletNoise.cpp:185-215      eval3d()
//   evaluate3DProjected   WaveletNoise.cpp:218-265      eval3d_projected()
//   PerlinNoise::noise    experient/PerlinNoise.hpp:13-56 (== perlin.h:17-62)   perlin3()
//   wavelet_texture/noise_texture::value   texture.h:67-107 / :37-43
//   calculateStats        WaveletNoise.cpp:268-288
//
// The tile (8 MiB at n = 128) stays L2-resident; taps are read through the read-only path.
#include "wn_internal.h"

#include <cstdlib>

namespace {

#define FMUL __fmul_rn
#define FADD __fadd_rn
#define FSUB __fsub_rn

__device__ __forceinline__ int tmod(int x, const WnTileView &t)          // Mod(), cpp:31-34
{
    if (t.pow2) return x & (t.n - 1);
    int m = x % t.n;
    return m < 0 ? m + t.n : m;
}

// quadratic B-spline weights of one axis (cpp:121-127 / :194-200)
__device__ __forceinline__ void basis(float p, int &mid, float w[3])
{
    const float a = FSUB(p, 0.5f);
    mid = (int)ceilf(a);
    const float t = FSUB((float)mid, a);
    w[0] = FMUL(FMUL(t, t), 0.5f);                       // t*t/2.0f  (x/2 == x*0.5 exactly)
    const float s = FSUB(1.0f, t);
    w[2] = FMUL(FMUL(s, s), 0.5f);
    w[1] = FSUB(FSUB(1.0f, w[0]), w[2]);
}

__device__ __forceinline__ float eval2d(const WnTileView &t, float px, float py)
{
    int mx, my; float wx[3], wy[3];
    basis(px, mx, wx); basis(py, my, wy);
    int cx[3], cy[3];
#pragma unroll
    for (int f = 0; f < 3; ++f) { cx[f] = tmod(mx + f - 1, t); cy[f] = tmod(my + f - 1, t) * t.n; }
    float r = 0.0f;
#pragma unroll
    for (int fy = 0; fy < 3; ++fy)
#pragma unroll
        for (int fx = 0; fx < 3; ++fx)
            r = FADD(r, FMUL(FMUL(wx[fx], wy[fy]), __ldg(t.N + cx[fx] + cy[fy])));
    return r;
}

__device__ __forceinline__ float eval3d(const WnTileView &t, float px, float py, float pz)
{
    int mx, my, mz; float wx[3], wy[3], wz[3];
    basis(px, mx, wx); basis(py, my, wy); basis(pz, mz, wz);
    int cx[3], cy[3], cz[3];
#pragma unroll
    for (int f = 0; f < 3; ++f) {
        cx[f] = tmod(mx + f - 1, t);
        cy[f] = tmod(my + f - 1, t) * t.n;
        cz[f] = tmod(mz + f - 1, t) * t.n * t.n;
    }
    float r = 0.0f;
#pragma unroll
    for (int fz = 0; fz < 3; ++fz)
#pragma unroll
        for (int fy = 0; fy < 3; ++fy)
#pragma unroll
            for (int fx = 0; fx < 3; ++fx)          // weight = (wx*wy)*wz, cpp:206
                r = FADD(r, FMUL(FMUL(FMUL(wx[fx], wy[fy]), wz[fz]), __ldg(t.N + cx[fx] + cy[fy] + cz[fz])));
    return r;
}

__device__ __forceinline__ float multiband3d(const WnTileView &t, const WnBands &b, float x, float y, float z)
{
    float acc = 0.0f;
    for (int k = 0; k < b.nbands; ++k) {
        const float s = b.scale[k];
        acc = FADD(acc, FMUL(b.weight[k], eval3d(t, FMUL(x, s), FMUL(y, s), FMUL(z, s))));
    }
    return FMUL(acc, b.post);
}

__device__ float eval3d_projected(const WnTileView &t, const float p[3], const float nrm[3])
{
    int lo[3], hi[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        // support = 3|n_i| + 3 sqrt((1 - n_i^2)/2), cpp:229
        const float support = FADD(FMUL(3.0f, fabsf(nrm[i])),
                                   FMUL(3.0f, __fsqrt_rn(FMUL(FSUB(1.0f, FMUL(nrm[i], nrm[i])), 0.5f))));
        lo[i] = (int)ceilf(FSUB(p[i], support));
        hi[i] = (int)floorf(FADD(p[i], support));
    }
    // |support| <= 3 + 3/sqrt(2) for a unit normal, so a box wider than 12 cells means non-finite or
    // out-of-int-range input (undefined behaviour in the reference); refuse instead of looping forever.
#pragma unroll
    for (int i = 0; i < 3; ++i)
        if (hi[i] - lo[i] > 12 || hi[i] < lo[i] - 1 || hi[i] == 0x7fffffff) return 0.0f;
    const float q0 = FSUB(p[0], 1.5f), q1 = FSUB(p[1], 1.5f), q2 = FSUB(p[2], 1.5f);
    // Row culling.  A candidate contributes only if every t_i lies in (0,3); in real arithmetic
    //   t_i = 1.5 - D_i + n_i*dot/2,  D = p - c,  dot = n.D,
    // which is affine in D_0 for a fixed (c1, c2) row, so the x range that can contribute is an interval.  It is
    // evaluated in float with a slack `eps` that covers the reference's own rounding of t_i (it forms c_i + ... and
    // p_i - 1.5 in float: a few ulp of |p_i|), so only candidates the reference rejects (weight 0) are skipped and
    // the surviving ones are evaluated with the reference's exact operation order: the result stays bit-identical.
    const float pmax = fmaxf(fmaxf(fabsf(p[0]), fabsf(p[1])), fabsf(p[2])) + 16.0f;
    const float eps = pmax * 9.6e-7f + 1e-4f;                  // 8 ulp of the largest coordinate + a constant
    const float b0 = 1.0f - 0.5f * nrm[0] * nrm[0];           // -d t_0 / d D_0  (in [0.5, 1])
    const float b1 = 0.5f * nrm[0] * nrm[1];                   //  d t_1 / d D_0
    const float b2 = 0.5f * nrm[0] * nrm[2];                   //  d t_2 / d D_0
    // the bounds are only a conservative filter (eps absorbs a few ulp), so reciprocals replace the five divisions
    // per row: about 40 instructions less per row of the bounding box
    const bool use1 = fabsf(b1) > 1e-6f, use2 = fabsf(b2) > 1e-6f;
    const float ib0 = 1.0f / b0, ib1 = use1 ? 1.0f / b1 : 0.0f, ib2 = use2 ? 1.0f / b2 : 0.0f;
    float result = 0.0f;
    for (int c2 = lo[2]; c2 <= hi[2]; ++c2) {
        const float f2 = (float)c2;
        const int i2 = tmod(c2, t) * t.n * t.n;
        const float D2 = p[2] - f2;
        for (int c1 = lo[1]; c1 <= hi[1]; ++c1) {
            const float f1 = (float)c1;
            const int i1 = tmod(c1, t) * t.n;
            // admissible D_0 interval of this row
            const float D1 = p[1] - f1;
            const float K = nrm[1] * D1 + nrm[2] * D2;
            float dlo = -1e30f, dhi = 1e30f;
            {   // t_0 = (1.5 + n0 K/2) - b0 D0 in (-eps, 3+eps)
                const float a = 1.5f + 0.5f * nrm[0] * K;
                dlo = fmaxf(dlo, (a - 3.0f - eps) * ib0);
                dhi = fminf(dhi, (a + eps) * ib0);
            }
            bool row_empty = false;
            {   // t_1 = (1.5 - D1 + n1 K/2) + b1 D0
                const float a = 1.5f - D1 + 0.5f * nrm[1] * K;
                if (use1) {
                    const float x0 = (-eps - a) * ib1, x1 = (3.0f + eps - a) * ib1;
                    dlo = fmaxf(dlo, fminf(x0, x1)); dhi = fminf(dhi, fmaxf(x0, x1));
                } else if (a <= -eps - 8.0f * fabsf(b1) || a >= 3.0f + eps + 8.0f * fabsf(b1)) row_empty = true;
            }
            {   // t_2 = (1.5 - D2 + n2 K/2) + b2 D0
                const float a = 1.5f - D2 + 0.5f * nrm[2] * K;
                if (use2) {
                    const float x0 = (-eps - a) * ib2, x1 = (3.0f + eps - a) * ib2;
                    dlo = fmaxf(dlo, fminf(x0, x1)); dhi = fminf(dhi, fmaxf(x0, x1));
                } else if (a <= -eps - 8.0f * fabsf(b2) || a >= 3.0f + eps + 8.0f * fabsf(b2)) row_empty = true;
            }
            if (row_empty || !(dlo <= dhi)) continue;
            // c0 = p0 - D0: widen by eps again for the float evaluation of the bounds themselves
            const int r_lo = max(lo[0], (int)floorf(p[0] - dhi - eps));
            const int r_hi = min(hi[0], (int)ceilf(p[0] - dlo + eps));
            for (int c0 = r_lo; c0 <= r_hi; ++c0) {
                const float f0 = (float)c0;
                // dot = ((0 + n0 (p0-c0)) + n1 (p1-c1)) + n2 (p2-c2), cpp:239-240
                float dot = FADD(0.0f, FMUL(nrm[0], FSUB(p[0], f0)));
                dot = FADD(dot, FMUL(nrm[1], FSUB(p[1], f1)));
                dot = FADD(dot, FMUL(nrm[2], FSUB(p[2], f2)));
                float weight = 1.0f;
                const float fc[3] = { f0, f1, f2 };
                const float qq[3] = { q0, q1, q2 };
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    // t = (c_i + n_i*dot/2) - (p_i - 1.5), cpp:245
                    const float tt = FSUB(FADD(fc[i], FMUL(FMUL(nrm[i], dot), 0.5f)), qq[i]);
                    if (tt <= 0.0f || tt >= 3.0f) { weight = 0.0f; break; }
                    float piece;
                    if (tt < 1.0f) piece = FMUL(FMUL(tt, tt), 0.5f);
                    else if (tt < 2.0f) {
                        const float t1 = FSUB(tt, 1.0f), t2 = FSUB(2.0f, tt);
                        piece = FSUB(1.0f, FMUL(FADD(FMUL(t1, t1), FMUL(t2, t2)), 0.5f));
                    } else {
                        const float t3 = FSUB(3.0f, tt);
                        piece = FMUL(FMUL(t3, t3), 0.5f);
                    }
                    weight = FMUL(weight, piece);
                }
                // cpp:257 compares the float against the DOUBLE literal 1e-6.  float(1e-6) = 9.99999997e-7 is the
                // largest float below that double, so `(double)w > 1e-6` <=> `w > 1e-6f` exactly (no FP64 needed).
                if (weight > 1e-6f)
                    result = FADD(result, FMUL(weight, __ldg(t.N + tmod(c0, t) + i1 + i2)));
            }
        }
    }
    return result;
}

// ---- Perlin, double precision -------------------------------------------------------------------
__device__ __forceinline__ double pfade(double t)    // t*t*t*(t*(t*6-15)+10)
{
    const double inner = __dadd_rn(__dmul_rn(t, __dsub_rn(__dmul_rn(t, 6.0), 15.0)), 10.0);
    return __dmul_rn(__dmul_rn(__dmul_rn(t, t), t), inner);
}
__device__ __forceinline__ double plerp(double t, double a, double b) { return __dadd_rn(a, __dmul_rn(t, __dsub_rn(b, a))); }
__device__ __forceinline__ double pgrad(int hash, double x, double y, double z)
{
    const int h = hash & 15;
    const double u = h < 8 ? x : y;
    const double v = h < 4 ? y : ((h == 12 || h == 14) ? x : z);
    return __dadd_rn((h & 1) == 0 ? u : -u, (h & 2) == 0 ? v : -v);
}
__device__ double perlin3(const int *p /* 512 ints, shared memory */, double x, double y, double z)
{
    const double fx = floor(x), fy = floor(y), fz = floor(z);
    const int X = (int)fx & 255, Y = (int)fy & 255, Z = (int)fz & 255;
    x = __dsub_rn(x, fx); y = __dsub_rn(y, fy); z = __dsub_rn(z, fz);
    const double u = pfade(x), v = pfade(y), w = pfade(z);
    const int A = p[X] + Y, AA = p[A] + Z, AB = p[A + 1] + Z;
    const int B = p[X + 1] + Y, BA = p[B] + Z, BB = p[B + 1] + Z;
    const double x1 = __dsub_rn(x, 1.0), y1 = __dsub_rn(y, 1.0), z1 = __dsub_rn(z, 1.0);
    return plerp(w,
        plerp(v, plerp(u, pgrad(p[AA], x, y, z),      pgrad(p[BA], x1, y, z)),
                 plerp(u, pgrad(p[AB], x, y1, z),     pgrad(p[BB], x1, y1, z))),
        plerp(v, plerp(u, pgrad(p[AA + 1], x, y, z1), pgrad(p[BA + 1], x1, y, z1)),
                 plerp(u, pgrad(p[AB + 1], x, y1, z1), pgrad(p[BB + 1], x1, y1, z1))));
}

// ---- Perlin, single precision (fast mode, opt-in) -----------------------------------------------------
// The same algorithm in FP32 with FMAs: every reference caller passes float-valued coordinates, so x - floor(x) is
// exact in FP32 too and the result differs from the FP64 kernel by rounding only (tests bound it by 1e-5 * range 2).
// The float selects of grad() are single ALU ops where the FP64 ones are two, and the arithmetic leaves the FP64 pipe.
__device__ __forceinline__ float pfadef(float t) { return t * t * t * fmaf(t, fmaf(t, 6.0f, -15.0f), 10.0f); }
__device__ __forceinline__ float plerpf(float t, float a, float b) { return fmaf(t, b - a, a); }
// gradient of hash h as coefficients of (x, y, z): grad(h, x, y, z) = cx x + cy y + cz z with one coefficient 0
__device__ __forceinline__ float4 pgrad_coeff(int hash)
{
    const int h = hash & 15;
    const float su = (h & 1) ? -1.0f : 1.0f, sv = (h & 2) ? -1.0f : 1.0f;
    float cx = 0.0f, cy = 0.0f, cz = 0.0f;
    if (h < 8) cx += su; else cy += su;
    if (h < 4) cy += sv; else if (h == 12 || h == 14) cx += sv; else cz += sv;
    return make_float4(cx, cy, cz, 0.0f);
}
// fast mode on image grids and batches: g[i] = pgrad_coeff(p[i]) (512 entries, built per CTA), so a corner is one
// 16-byte shared-memory read and three FMAs instead of an integer lookup and nine selects
__device__ float perlin3f_table(const int *p, const float4 *g, float x, float y, float z)
{
    const float fx = floorf(x), fy = floorf(y), fz = floorf(z);
    const int X = (int)fx & 255, Y = (int)fy & 255, Z = (int)fz & 255;
    x -= fx; y -= fy; z -= fz;
    const float u = pfadef(x), v = pfadef(y), w = pfadef(z);
    const int A = p[X] + Y, AA = p[A] + Z, AB = p[A + 1] + Z;
    const int B = p[X + 1] + Y, BA = p[B] + Z, BB = p[B + 1] + Z;
    const float x1 = x - 1.0f, y1 = y - 1.0f, z1 = z - 1.0f;
    auto dot = [](const float4 c, float a, float b, float d) { return fmaf(c.z, d, fmaf(c.y, b, c.x * a)); };
    return plerpf(w,
        plerpf(v, plerpf(u, dot(g[AA], x, y, z),      dot(g[BA], x1, y, z)),
                  plerpf(u, dot(g[AB], x, y1, z),     dot(g[BB], x1, y1, z))),
        plerpf(v, plerpf(u, dot(g[AA + 1], x, y, z1), dot(g[BA + 1], x1, y, z1)),
                  plerpf(u, dot(g[AB + 1], x, y1, z1), dot(g[BB + 1], x1, y1, z1))));
}

// ---- coordinate generators -------------------------------------------------------------------------
__device__ __forceinline__ void coord2(const WnPointsAoS &c, size_t s, float p[3])
{
    p[0] = FMUL(__ldg(c.p + s * 2), c.pre);
    p[1] = FMUL(__ldg(c.p + s * 2 + 1), c.pre);
}
__device__ __forceinline__ void coord(const WnPointsAoS &c, size_t s, float p[3])
{
#pragma unroll
    for (int i = 0; i < 3; ++i) p[i] = FMUL(__ldg(c.p + s * 3 + i), c.pre);
}
// s = q * n + r.  Sample indices below 2^32 (every configuration in BASELINE.json) take a 32-bit division: the 64-bit
// one is a ~70-instruction sequence, a sixth of the Perlin kernel.
__device__ __forceinline__ void split_index(size_t s, int n, size_t &q, size_t &r)
{
    if (s <= 0xffffffffull) {
        const unsigned s32 = (unsigned)s, q32 = s32 / (unsigned)n;
        q = q32;
        r = s32 - q32 * (unsigned)n;
    } else {
        q = s / (size_t)n;
        r = s - q * (size_t)n;
    }
}
__device__ __forceinline__ void coord(const WnLattice &c, size_t s, float p[3])
{
    size_t row, i, k, j;
    split_index(s, c.nx, row, i);
    p[0] = __ldg(c.xs + i);
    split_index(row, c.ny, k, j);
    p[1] = __ldg(c.ys + j);
    p[2] = c.zs ? __ldg(c.zs + k) : 0.0f;
}
__device__ __forceinline__ void coord(const WnAffine &c, size_t s, float p[3])
{
    size_t j, i;
    split_index(s, c.nu, j, i);
    const float u = __ldg(c.us + i), v = __ldg(c.vs + j);
#pragma unroll
    for (int i = 0; i < 3; ++i)
        p[i] = FMUL(FADD(FADD(c.o[i], FMUL(u, c.e1[i])), FMUL(v, c.e2[i])), c.pre);
}

#define WN_TID_OR_RETURN(count)                                                        \
    const size_t s = blockIdx.x * (size_t)blockDim.x + threadIdx.x;                    \
    if (s >= (count)) return

__global__ void k_eval2d_points(WnTileView t, WnPointsAoS c, size_t first, size_t count, float post, float *out)
{
    WN_TID_OR_RETURN(count);
    float p[3]; coord2(c, first + s, p);
    out[s] = FMUL(eval2d(t, p[0], p[1]), post);
}
__global__ void k_eval2d_lattice(WnTileView t, WnLattice c, float pre, size_t first, size_t count, float post, float *out)
{
    WN_TID_OR_RETURN(count);
    float p[3]; coord(c, first + s, p);
    out[s] = FMUL(eval2d(t, FMUL(p[0], pre), FMUL(p[1], pre)), post);
}
template <class C>
__global__ void k_mb3d(WnTileView t, C c, WnBands b, size_t first, size_t count, float *out)
{
    WN_TID_OR_RETURN(count);
    float p[3];
    coord(c, first + s, p);
    out[s] = multiband3d(t, b, p[0], p[1], p[2]);
}
template <class C>
__global__ void k_proj(WnTileView t, C c, const float *normals, float n0, float n1, float n2, size_t first,
                       size_t count, float post, float *out)
{
    WN_TID_OR_RETURN(count);
    float p[3];
    coord(c, first + s, p);
    float nrm[3] = { n0, n1, n2 };
    if (normals) { nrm[0] = __ldg(normals + 3 * s); nrm[1] = __ldg(normals + 3 * s + 1); nrm[2] = __ldg(normals + 3 * s + 2); }
    out[s] = FMUL(eval3d_projected(t, p, nrm), post);
}
// Cook & DeRose App. 2 WMultibandNoise(p, s, normal, firstBand, nbands, w), statement by statement: active bands b < nb_active (s + firstBand + b < 0, decided on the host), q = 2 p
// 2^(firstBand+b) evaluated in double like the listing's pow(), result /= sqrt(variance * (normal ? 0.296 : 0.210)).
__global__ void k_wmultiband(WnTileView t, const float *p, size_t count, WnBands b /* scale[] = 2^(firstBand+b) */,
                             int nb_active, int projected, float n0, float n1, float n2, double denom, float *out)
{
    WN_TID_OR_RETURN(count);
    const float nrm[3] = { n0, n1, n2 };
    float result = 0.0f;
    for (int k = 0; k < nb_active; ++k) {
        float q[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) q[i] = (float)__dmul_rn(__dmul_rn(2.0, (double)__ldg(p + 3 * s + i)), (double)b.scale[k]);
        const float v = projected ? eval3d_projected(t, q, nrm) : eval3d(t, q[0], q[1], q[2]);
        result = FADD(result, FMUL(b.weight[k], v));
    }
    if (denom != 0.0) result = (float)__ddiv_rn((double)result, denom);
    out[s] = result;
}

// Image grids: the 32 lanes of a warp are neighbouring pixels, a fraction of a tile cell apart, so they weigh almost
// the same cells.  Per lane, eval3d_projected() spends most of its instructions FINDING its ~52 contributing cells (row
// bounds for every row of its bounding box) and the lanes' loops diverge (ncu: 25 of 32 lanes active).  Here the warp
// builds ONE candidate list cooperatively -- each lane takes one (y, z) row of the union bounding box and solves the
// exact support region A (c - p) in (-1.5, 1.5)^3, A = I - n n^T / 2, widened by the extent of the warp's points and a
// rounding slack, for the row's interval of x; the rows' counts are scanned across the warp so the cells land in shared
// memory in the reference's visiting order (z outer, x inner) -- and then every lane evaluates every listed cell for
// its own point with the reference's arithmetic (cpp:239-260), two cells per iteration on the packed FP32 pipe.  A
// listed cell that is outside a lane's own reference box, or that the reference weighs with 0, contributes nothing, so
// each lane adds its contributing cells in the reference's order: bit-identical to eval3d_projected().
constexpr int PROJ_LIST = 192;                                 // candidate cells per warp (config 4 needs ~95)

__global__ void __launch_bounds__(128, 8) k_proj_grid(WnTileView t, WnAffine c, float n0, float n1, float n2, size_t first,
                                                   size_t count, float post, float *out)
{
    // candidate cells in pairs: {x_a, x_b, y_a, y_b}, {z_a, z_b, tile index a, b (int bits)} -- cell coordinates as floats,
    // laid out so that two LDS.128 deliver the operand pairs of the packed evaluation below
    __shared__ float4 s_list[4][PROJ_LIST];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *const my_list = reinterpret_cast<float *>(s_list[warp]);
    auto entry = [&](int k) {                                  // cell k of this warp's list as {x, y, z, index}
        const float *q = my_list + (k >> 1) * 8 + (k & 1);
        return make_float4(q[0], q[2], q[4], q[6]);
    };
    // a warp owns an 8 x 4 patch of pixels (not 32 pixels of one row): its points are closer together, so the union of
    // their candidate cells is smaller (about 1.15x a single point's instead of 1.6x)
    const size_t row0 = first / (size_t)c.nu;                  // first image row of this launch's sample window
    const int patches_x = (c.nu + 7) / 8;
    const size_t wid = blockIdx.x * (size_t)(blockDim.x >> 5) + warp;
    const size_t pi = (wid % patches_x) * 8 + (lane & 7), pj = row0 + (wid / patches_x) * 4 + (lane >> 3);
    const size_t sg = pi + (size_t)c.nu * pj;                  // global sample index
    bool live = pi < (size_t)c.nu && pj < (size_t)c.nv && sg >= first && sg < first + count;
    const bool store = live;
    const size_t s = sg - first;
    const unsigned livemask = __ballot_sync(full, live);
    if (!livemask) return;
    float p[3] = { 0.0f, 0.0f, 0.0f };
    if (live) coord(c, sg, p);
    const float nrm[3] = { n0, n1, n2 };
    // the reference's own box of this lane (cpp:228-232)
    int lo[3], hi[3];
    float sup[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        sup[i] = FADD(FMUL(3.0f, fabsf(nrm[i])), FMUL(3.0f, __fsqrt_rn(FMUL(FSUB(1.0f, FMUL(nrm[i], nrm[i])), 0.5f))));
        lo[i] = (int)ceilf(FSUB(p[i], sup[i]));
        hi[i] = (int)floorf(FADD(p[i], sup[i]));
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
        if (hi[i] - lo[i] > 12 || hi[i] < lo[i] - 1 || hi[i] == 0x7fffffff) live = false;     // see eval3d_projected
    const unsigned okmask = __ballot_sync(full, live);
    float result = 0.0f;
    if (okmask) {
        // union box and the extent of the warp's points (dead lanes borrow a live lane's values)
        const int src = __ffs((int)okmask) - 1;
        int ulo[3], uhi[3];
        float pc[3], hw[3];
        float pmax = 0.0f;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float borrowed = __shfl_sync(full, p[i], src);       // executed by every lane (never inside a branch)
            const float pl = live ? p[i] : borrowed;
            float mn = pl, mx = pl;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                mn = fminf(mn, __shfl_xor_sync(full, mn, o));
                mx = fmaxf(mx, __shfl_xor_sync(full, mx, o));
            }
            pc[i] = 0.5f * (mn + mx);
            hw[i] = 0.5f * (mx - mn);
            ulo[i] = __reduce_min_sync(full, live ? lo[i] : 0x7fffffff);
            uhi[i] = __reduce_max_sync(full, live ? hi[i] : (int)0x80000000);
            pmax = fmaxf(pmax, fmaxf(fabsf(mn), fabsf(mx)));
        }
        const int e0 = uhi[0] - ulo[0] + 1, e1 = uhi[1] - ulo[1] + 1, e2 = uhi[2] - ulo[2] + 1;
        const int total = e0 * e1 * e2;
        // A lane's own box (the only cells the reference visits, cpp:235-237) is the union box minus, per face, at most
        // one layer of cells when the warp's points are less than a cell apart: `excl` has bit 2i / 2i+1 set when the
        // lane's box starts one cell after / ends one cell before the union's on axis i, and a listed cell carries the
        // faces of the union box it lies on in bits 24..29 of its tile index, so "inside my box" is one AND.
        unsigned excl = 0;
        bool near = true;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            excl |= (lo[i] > ulo[i] ? 1u : 0u) << (2 * i) | (hi[i] < uhi[i] ? 1u : 0u) << (2 * i + 1);
            near = near && lo[i] - ulo[i] <= 1 && uhi[i] - hi[i] <= 1;
        }
        excl <<= 24;
        const bool nearbox = __all_sync(full, !live || near) && (long long)t.n * t.n * t.n <= (1LL << 24);
        int listed = PROJ_LIST + 1;                            // > PROJ_LIST: fall back to the per-lane walk
        if (e0 <= 32 && e1 <= 32 && e2 <= 32 && total <= 8192) {
            // |A (c - p)|_i < 1.5 for some lane  =>  |A (c - pc)|_i < 1.5 + sum_j |A_ij| hw_j (+ rounding slack)
            const float eps = (pmax + 16.0f) * 4.0e-6f + 1.0e-4f;
            float bound[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                float sl = 0.0f;
#pragma unroll
                for (int j = 0; j < 3; ++j) sl += fabsf((i == j ? 1.0f : 0.0f) - 0.5f * nrm[i] * nrm[j]) * hw[j];
                bound[i] = 1.5f + sl + eps;
            }
            listed = 0;
            // One (y, z) row of the union box per lane.  (A d)_i is affine in d_0 = c_0 - pc_0 for a fixed row:
            //   (A d)_0 = a0 d_0 + k0,  (A d)_1 = a1 d_0 + k1,  (A d)_2 = a2 d_0 + k2   (a0 = 1 - n0^2/2 in [0.5, 1]),
            // so the cells of the row that can pass |(A d)_i| < bound_i form one interval of x; it is widened by eps again
            // for the rounding of its own end points (a listed cell that does not contribute is only wasted work).  The
            // rows' cell counts are scanned across the warp, so the list keeps the reference's visiting order.
            const float a0 = 1.0f - 0.5f * nrm[0] * nrm[0], a1 = -0.5f * nrm[1] * nrm[0], a2 = -0.5f * nrm[2] * nrm[0];
            const bool use1 = fabsf(a1) > 1e-6f, use2 = fabsf(a2) > 1e-6f;
            const float ia0 = 1.0f / a0, ia1 = use1 ? 1.0f / a1 : 0.0f, ia2 = use2 ? 1.0f / a2 : 0.0f;
            const int nrows = e1 * e2;
            const float inv1 = 1.0f / (float)e1;               // row -> (y, z) by a float reciprocal: exact for rows < 2^10
            for (int base = 0; base < nrows; base += 32) {
                const int rr = base + lane;
                const int z = (int)(((float)rr + 0.5f) * inv1), y = rr - z * e1;
                int cnt = 0, xlo = 0;
                if (rr < nrows) {
                    const float d1 = (float)(ulo[1] + y) - pc[1], d2 = (float)(ulo[2] + z) - pc[2];
                    const float mk = 0.5f * (nrm[1] * d1 + nrm[2] * d2);
                    const float k0 = -nrm[0] * mk, k1 = d1 - nrm[1] * mk, k2 = d2 - nrm[2] * mk;
                    float dlo = fmaxf((-bound[0] - k0) * ia0, -32.0f), dhi = fminf((bound[0] - k0) * ia0, 32.0f);
                    bool empty = false;
                    if (use1) {
                        const float x0 = (-bound[1] - k1) * ia1, x1 = (bound[1] - k1) * ia1;
                        dlo = fmaxf(dlo, fminf(x0, x1)); dhi = fminf(dhi, fmaxf(x0, x1));
                    } else empty = fabsf(k1) >= bound[1] + 1.0e-4f;   // |a1 d_0| <= 1e-6 * 32
                    if (use2) {
                        const float x0 = (-bound[2] - k2) * ia2, x1 = (bound[2] - k2) * ia2;
                        dlo = fmaxf(dlo, fminf(x0, x1)); dhi = fminf(dhi, fmaxf(x0, x1));
                    } else empty = empty || fabsf(k2) >= bound[2] + 1.0e-4f;
                    if (!empty && dlo <= dhi) {
                        xlo = max(ulo[0], (int)ceilf(pc[0] + dlo - eps));
                        cnt = max(0, min(uhi[0], (int)floorf(pc[0] + dhi + eps)) - xlo + 1);
                    }
                }
                int inc = cnt;                                 // inclusive scan of the rows' counts
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(full, inc, o);
                    if (lane >= o) inc += v;
                }
                const int start = listed + inc - cnt;
                listed += __shfl_sync(full, inc, 31);
                if (cnt > 0) {
                    const float fy = (float)(ulo[1] + y), fz = (float)(ulo[2] + z);
                    const int iyz = tmod(ulo[1] + y, t) * t.n + tmod(ulo[2] + z, t) * t.n * t.n;
                    const int faces_yz = nearbox ? ((y == 0) << 2 | (y == e1 - 1) << 3 | (z == 0) << 4 | (z == e2 - 1) << 5) << 24 : 0;
                    int ix = tmod(xlo, t);
                    for (int j = 0; j < cnt; ++j) {
                        const int pos = start + j, x = xlo + j;
                        if (pos < PROJ_LIST) {
                            float *q = my_list + (pos >> 1) * 8 + (pos & 1);
                            q[0] = (float)x; q[2] = fy; q[4] = fz;
                            const int faces_x = nearbox ? ((x == ulo[0]) | (x == uhi[0]) << 1) << 24 : 0;
                            q[6] = __int_as_float((ix + iyz) | faces_yz | faces_x);
                        }
                        if (++ix == t.n) ix = 0;
                    }
                }
            }
            __syncwarp();
            if (lane < 4 && (listed & 1) && listed < PROJ_LIST) {  // odd count: the last pair's second cell repeats the first
                float *q = my_list + (listed >> 1) * 8 + 2 * lane;
                q[1] = q[0];
            }
            __syncwarp();
        }
        if (listed > PROJ_LIST) {
            if (live) result = eval3d_projected(t, p, nrm);   // incoherent warp (not an image grid after all)
        } else if (live) {
            const float q0 = FSUB(p[0], 1.5f), q1 = FSUB(p[1], 1.5f), q2 = FSUB(p[2], 1.5f);
            const float flo[3] = { (float)lo[0], (float)lo[1], (float)lo[2] }, fhi[3] = { (float)hi[0], (float)hi[1], (float)hi[2] };
            // one candidate cell, evaluated exactly like the reference's inner statement (cpp:239-260)
            auto cell = [&](const float4 e, bool ok) {
                const float fc[3] = { e.x, e.y, e.z };
                // dot = ((0 + n0 (p0-c0)) + n1 (p1-c1)) + n2 (p2-c2), cpp:239-240
                float dot = FADD(0.0f, FMUL(nrm[0], FSUB(p[0], fc[0])));
                dot = FADD(dot, FMUL(nrm[1], FSUB(p[1], fc[1])));
                dot = FADD(dot, FMUL(nrm[2], FSUB(p[2], fc[2])));
                const float qq[3] = { q0, q1, q2 };
                // The reference stops at the first axis whose t leaves (0, 3) and skips the cell (cpp:247-250); here all
                // three axes are evaluated without branches and the cell is skipped through `ok`: the same cells
                // contribute, with the same weight ((1 * piece0) * piece1) * piece2 (1 * x is x exactly).
                float weight = 0.0f;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    // t = (c_i + n_i*dot/2) - (p_i - 1.5), cpp:245
                    const float tt = FSUB(FADD(fc[i], FMUL(FMUL(nrm[i], dot), 0.5f)), qq[i]);
                    ok = ok && tt > 0.0f && tt < 3.0f;
                    const float edge = tt < 1.0f ? tt : FSUB(3.0f, tt);                      // t^2/2 or (3-t)^2/2
                    const float pe = FMUL(FMUL(edge, edge), 0.5f);
                    const float t1 = FSUB(tt, 1.0f), t2 = FSUB(2.0f, tt);
                    const float pm = FSUB(1.0f, FMUL(FADD(FMUL(t1, t1), FMUL(t2, t2)), 0.5f));
                    const float piece = (tt >= 1.0f && tt < 2.0f) ? pm : pe;
                    weight = i == 0 ? piece : FMUL(weight, piece);
                }
                if (ok && weight > 1e-6f) result = FADD(result, FMUL(weight, __ldg(t.N + __float_as_int(e.w))));
            };
            if (nearbox) {
                // Two listed cells per iteration on the packed FP32 pipe: FADD2 / FMUL2 are the same IEEE operations as
                // the scalar ones two at a time (a - b is formed as a + (-b), exact), so every cell keeps the reference's
                // arithmetic; the two contributions are then added in list order.
                // ptxas (12.9) contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under --fmad=false, which the
                // scalar .rn forms never are.  Where a sum takes a rounded product (the dot product, t1^2 + t2^2) the
                // add is therefore done with two scalar FADDs (add2s); where the product is a multiplication by 0.5
                // (exact: x/2 + y rounds like fl(x/2) + y) a contraction cannot change the result.
                auto add2s = [](const float2 a, const float2 b) { return make_float2(FADD(a.x, b.x), FADD(a.y, b.y)); };
                const float2 n2[3] = { make_float2(nrm[0], nrm[0]), make_float2(nrm[1], nrm[1]), make_float2(nrm[2], nrm[2]) };
                const float2 p2[3] = { make_float2(p[0], p[0]), make_float2(p[1], p[1]), make_float2(p[2], p[2]) };
                const float2 nq2[3] = { make_float2(-q0, -q0), make_float2(-q1, -q1), make_float2(-q2, -q2) };
                const float2 half2 = make_float2(0.5f, 0.5f), one2 = make_float2(1.0f, 1.0f), two2 = make_float2(2.0f, 2.0f),
                             three2 = make_float2(3.0f, 3.0f), mone2 = make_float2(-1.0f, -1.0f);
                auto neg2 = [](const float2 a) { return make_float2(-a.x, -a.y); };
#pragma unroll 2
                for (int k = 0; k < listed; k += 2) {
                    const float4 e01 = s_list[warp][k], e23 = s_list[warp][k + 1];
                    const float2 fc[3] = { make_float2(e01.x, e01.y), make_float2(e01.z, e01.w), make_float2(e23.x, e23.y) };
                    float2 dot = __fadd2_rn(make_float2(0.0f, 0.0f), __fmul2_rn(n2[0], __fadd2_rn(p2[0], neg2(fc[0]))));
                    dot = add2s(dot, __fmul2_rn(n2[1], __fadd2_rn(p2[1], neg2(fc[1]))));
                    dot = add2s(dot, __fmul2_rn(n2[2], __fadd2_rn(p2[2], neg2(fc[2]))));
                    const unsigned wa = (unsigned)__float_as_int(e23.z), wb = (unsigned)__float_as_int(e23.w);
                    bool oka = (wa & excl) == 0, okb = (wb & excl) == 0 && k + 1 < listed;
                    float2 weight = make_float2(0.0f, 0.0f);
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const float2 tt = __fadd2_rn(__fadd2_rn(fc[i], __fmul2_rn(__fmul2_rn(n2[i], dot), half2)), nq2[i]);
                        // The reference's tests (0 < t < 3; t < 1, t < 2: cpp:247-256) through u = t - 1.5, which is
                        // exact for t in [0.75, 3]: |u| < 1.5 drops, besides t <= 0 and t >= 3, only t < 2^-24 (piece
                        // t^2/2 < 2^-49: the cell fails the 1e-6 weight test either way); |u| < 0.5 is 1 < t < 2, and
                        // at t = 1 the edge polynomial t^2/2 and the middle one give the same 0.5.
                        const float2 u = __fadd2_rn(tt, make_float2(-1.5f, -1.5f));
                        oka = oka && fabsf(u.x) < 1.5f;
                        okb = okb && fabsf(u.y) < 1.5f;
                        const float2 t3 = __fadd2_rn(three2, neg2(tt));
                        const float2 edge = make_float2(u.x < 0.0f ? tt.x : t3.x, u.y < 0.0f ? tt.y : t3.y);
                        const float2 pe = __fmul2_rn(__fmul2_rn(edge, edge), half2);
                        const float2 t1 = __fadd2_rn(tt, mone2), t2 = __fadd2_rn(two2, neg2(tt));
                        const float2 pm = __fadd2_rn(one2, neg2(__fmul2_rn(add2s(__fmul2_rn(t1, t1), __fmul2_rn(t2, t2)), half2)));
                        const float2 piece = make_float2(fabsf(u.x) < 0.5f ? pm.x : pe.x, fabsf(u.y) < 0.5f ? pm.y : pe.y);
                        weight = i == 0 ? piece : __fmul2_rn(weight, piece);
                    }
                    if (oka && weight.x > 1e-6f) result = FADD(result, FMUL(weight.x, __ldg(t.N + (wa & 0xffffffu))));
                    if (okb && weight.y > 1e-6f) result = FADD(result, FMUL(weight.y, __ldg(t.N + (wb & 0xffffffu))));
                }
            } else {
                for (int k = 0; k < listed; ++k) {
                    const float4 e = entry(k);
                    // the reference only visits its own box (cpp:235-237)
                    cell(e, e.x >= flo[0] && e.x <= fhi[0] && e.y >= flo[1] && e.y <= fhi[1] && e.z >= flo[2] && e.z <= fhi[2]);
                }
            }
        }
    }
    if (store) out[s] = FMUL(result, post);
}

template <class C, bool FAST>
__global__ void k_perlin(const int32_t *perm, C c, size_t first, size_t count, float *out)
{
    __shared__ int sp[512];
    __shared__ float4 sg[FAST ? 512 : 1];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) {
        sp[i] = perm[i];
        if (FAST) sg[i] = pgrad_coeff(perm[i]);
    }
    __syncthreads();
    WN_TID_OR_RETURN(count);
    float p[3];
    coord(c, first + s, p);
    if (FAST) out[s] = perlin3f_table(sp, sg, p[0], p[1], p[2]);
    else out[s] = (float)perlin3(sp, (double)p[0], (double)p[1], (double)p[2]);
}
// double coordinates in, double noise out: PerlinNoise::noise(double, double, double) as declared
// (experient/PerlinNoise.hpp:36, perlin.h:42) for callers whose coordinates are not float-valued
__global__ void k_perlin_f64(const int32_t *perm, const double *p, size_t count, double *out)
{
    __shared__ int sp[512];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) sp[i] = perm[i];
    __syncthreads();
    WN_TID_OR_RETURN(count);
    out[s] = perlin3(sp, p[3 * s], p[3 * s + 1], p[3 * s + 2]);
}

// texture.h:67-107 (3D branch).  oct2 = octave_scale * 2.0f (float), inv_std = 1.0f/sqrt(0.18402f)
__global__ void k_wavelet_texture(WnTileView t, const float *p, size_t count, double scale, float oct2, float inv_std,
                                  float *grey)
{
    WN_TID_OR_RETURN(count);
    float q[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
        q[i] = FMUL(__double2float_rn(__dmul_rn((double)__ldg(p + 3 * s + i), scale)), oct2);
    double v = (double)eval3d(t, q[0], q[1], q[2]);
    v = __dmul_rn(v, (double)inv_std);
    double c = __ddiv_rn(v, 4.0);
    c = (c < -1.0) ? -1.0 : ((1.0 < c) ? 1.0 : c);                 // std::clamp
    grey[s] = __double2float_rn(__dmul_rn(0.5, __dadd_rn(1.0, c)));
}
// texture.h:86-99 (2D branch): evaluate2D of the xy components; inv_std = 1.0f/sqrt(0.19686f)
__global__ void k_wavelet_texture2d(WnTileView t, const float *p, size_t count, double scale, float oct2, float inv_std,
                                    float *grey)
{
    WN_TID_OR_RETURN(count);
    const float qx = FMUL(__double2float_rn(__dmul_rn((double)__ldg(p + 3 * s), scale)), oct2);
    const float qy = FMUL(__double2float_rn(__dmul_rn((double)__ldg(p + 3 * s + 1), scale)), oct2);
    double v = (double)eval2d(t, qx, qy);
    v = __dmul_rn(v, (double)inv_std);
    double c = __ddiv_rn(v, 4.0);
    c = (c < -1.0) ? -1.0 : ((1.0 < c) ? 1.0 : c);
    grey[s] = __double2float_rn(__dmul_rn(0.5, __dadd_rn(1.0, c)));
}
// texture.h:37-43: (p * float(scale)) * octave_scale in float, Perlin in double, 0.5*(1+v)
__global__ void k_perlin_texture(const int32_t *perm, const float *p, size_t count, float scale_f, float oct, float *grey)
{
    __shared__ int sp[512];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) sp[i] = perm[i];
    __syncthreads();
    WN_TID_OR_RETURN(count);
    double q[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) q[i] = (double)FMUL(FMUL(__ldg(p + 3 * s + i), scale_f), oct);
    const double v = perlin3(sp, q[0], q[1], q[2]);
    grey[s] = __double2float_rn(__dmul_rn(0.5, __dadd_rn(1.0, v)));
}

// calculateStats (cpp:268-288): double sums, float min/max.  Per-block partials, finished on the host.
__global__ void k_stats(const float *__restrict__ data, size_t count, double *__restrict__ partial)
{
    double sum = 0.0, sq = 0.0;
    float mn = 3.402823466e+38f, mx = -3.402823466e+38f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        const float v = data[i];
        sum += (double)v;
        sq += (double)v * (double)v;
        mn = fminf(mn, v); mx = fmaxf(mx, v);
    }
    __shared__ double s_sum[8], s_sq[8];
    __shared__ float s_mn[8], s_mx[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_down_sync(0xffffffffu, sum, o);
        sq += __shfl_down_sync(0xffffffffu, sq, o);
        mn = fminf(mn, __shfl_down_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_down_sync(0xffffffffu, mx, o));
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s_sum[warp] = sum; s_sq[warp] = sq; s_mn[warp] = mn; s_mx[warp] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
            sum += s_sum[w]; sq += s_sq[w]; mn = fminf(mn, s_mn[w]); mx = fmaxf(mx, s_mx[w]);
        }
        partial[4 * blockIdx.x + 0] = sum;
        partial[4 * blockIdx.x + 1] = sq;
        partial[4 * blockIdx.x + 2] = (double)mn;
        partial[4 * blockIdx.x + 3] = (double)mx;
    }
}

inline unsigned blocks_for(size_t count, int threads) { return (unsigned)((count + threads - 1) / threads); }

} // namespace

#define WN_T 256

int wn_launch_eval2d_points(WnTileView t, WnPointsAoS c, size_t first, size_t count, float post, float *out, cudaStream_t st)
{
    if (!count) return 0;
    k_eval2d_points<<<blocks_for(count, WN_T), WN_T, 0, st>>>(t, c, first, count, post, out);
    return 1;
}
int wn_launch_eval2d_lattice(WnTileView t, WnLattice c, float pre, size_t first, size_t count, float post, float *out, cudaStream_t st)
{
    if (!count) return 0;
    k_eval2d_lattice<<<blocks_for(count, WN_T), WN_T, 0, st>>>(t, c, pre, first, count, post, out);
    return 1;
}
int wn_launch_mb3d_points(WnTileView t, WnPointsAoS c, WnBands b, size_t first, size_t count, float *out, cudaStream_t st)
{
    if (!count) return 0;
    k_mb3d<WnPointsAoS><<<blocks_for(count, WN_T), WN_T, 0, st>>>(t, c, b, first, count, out);
    return 1;
}
int wn_launch_mb3d_lattice_exact(WnTileView t, WnLattice c, WnBands b, size_t first, size_t count, float *out, cudaStream_t st)
{
    if (!count) return 0;
    k_mb3d<WnLattice><<<blocks_for(count, WN_T), WN_T, 0, st>>>(t, c, b, first, count, out);
    return 1;
}
int wn_launch_mb3d_affine(WnTileView t, WnAffine c, WnBands b, size_t first, size_t count, float *out, cudaStream_t st)
{
    if (!count) return 0;
    k_mb3d<WnAffine><<<blocks_for(count, WN_T), WN_T, 0, st>>>(t, c, b, first, count, out);
    return 1;
}
int wn_launch_proj_points(WnTileView t, WnPointsAoS c, const float *normals, const float nrm[3], size_t first,
                          size_t count, float post, float *out, cudaStream_t st)
{
    if (!count) return 0;
    k_proj<WnPointsAoS><<<blocks_for(count, 128), 128, 0, st>>>(t, c, normals, nrm[0], nrm[1], nrm[2], first, count, post, out);
    return 1;
}
int wn_launch_proj_affine(WnTileView t, WnAffine c, const float nrm[3], size_t first, size_t count, float post,
                          float *out, cudaStream_t st)
{
    if (!count) return 0;
    bool per_lane = false;                                     // WN_PROJ_WARP=0: every lane on its own (A/B runs, tests)
    if (const char *e = getenv("WN_PROJ_WARP")) per_lane = atoi(e) == 0;
    if (per_lane) k_proj<WnAffine><<<blocks_for(count, 128), 128, 0, st>>>(t, c, nullptr, nrm[0], nrm[1], nrm[2], first, count, post, out);
    else {
        // one warp per 8 x 4 pixel patch over the image rows the sample window [first, first + count) touches
        const size_t row0 = first / (size_t)c.nu, row1 = (first + count - 1) / (size_t)c.nu;
        const size_t patches = (size_t)((c.nu + 7) / 8) * ((row1 - row0) / 4 + 1);
        k_proj_grid<<<(unsigned)((patches + 3) / 4), 128, 0, st>>>(t, c, nrm[0], nrm[1], nrm[2], first, count, post, out);
    }
    return 1;
}
int wn_launch_perlin_points(const int32_t *perm, WnPointsAoS c, size_t first, size_t count, float *out, int fast, cudaStream_t st)
{
    if (!count) return 0;
    if (fast) k_perlin<WnPointsAoS, true><<<blocks_for(count, WN_T), WN_T, 0, st>>>(perm, c, first, count, out);
    else k_perlin<WnPointsAoS, false><<<blocks_for(count, WN_T), WN_T, 0, st>>>(perm, c, first, count, out);
    return 1;
}
int wn_launch_perlin_lattice(const int32_t *perm, WnLattice c, size_t first, size_t count, float *out, int fast, cudaStream_t st)
{
    if (!count) return 0;
    if (fast) k_perlin<WnLattice, true><<<blocks_for(count, WN_T), WN_T, 0, st>>>(perm, c, first, count, out);
    else k_perlin<WnLattice, false><<<blocks_for(count, WN_T), WN_T, 0, st>>>(perm, c, first, count, out);
    return 1;
}
int wn_launch_perlin_affine(const int32_t *perm, WnAffine c, size_t first, size_t count, float *out, int fast, cudaStream_t st)
{
    if (!count) return 0;
    if (fast) k_perlin<WnAffine, true><<<blocks_for(count, WN_T), WN_T, 0, st>>>(perm, c, first, count, out);
    else k_perlin<WnAffine, false><<<blocks_for(count, WN_T), WN_T, 0, st>>>(perm, c, first, count, out);
    return 1;
}
int wn_launch_wmultiband(WnTileView t, const float *p, size_t count, WnBands b, int nb_active, const float *normal,
                         double denom, float *out, cudaStream_t st)
{
    if (!count) return 0;
    k_wmultiband<<<blocks_for(count, 128), 128, 0, st>>>(t, p, count, b, nb_active, normal != nullptr, normal ? normal[0] : 0.0f,
                                                         normal ? normal[1] : 0.0f, normal ? normal[2] : 0.0f, denom, out);
    return 1;
}
int wn_launch_perlin_points_f64(const int32_t *perm, const double *p, size_t count, double *out, cudaStream_t st)
{
    if (!count) return 0;
    k_perlin_f64<<<blocks_for(count, WN_T), WN_T, 0, st>>>(perm, p, count, out);
    return 1;
}
int wn_launch_wavelet_texture(WnTileView t, const float *p, size_t count, double scale, float oct2, float inv_std,
                              float *grey, cudaStream_t st)
{
    if (!count) return 0;
    k_wavelet_texture<<<blocks_for(count, WN_T), WN_T, 0, st>>>(t, p, count, scale, oct2, inv_std, grey);
    return 1;
}
int wn_launch_wavelet_texture2d(WnTileView t, const float *p, size_t count, double scale, float oct2, float inv_std,
                                float *grey, cudaStream_t st)
{
    if (!count) return 0;
    k_wavelet_texture2d<<<blocks_for(count, WN_T), WN_T, 0, st>>>(t, p, count, scale, oct2, inv_std, grey);
    return 1;
}
int wn_launch_perlin_texture(const int32_t *perm, const float *p, size_t count, float scale_f, float oct,
                             float *grey, cudaStream_t st)
{
    if (!count) return 0;
    k_perlin_texture<<<blocks_for(count, WN_T), WN_T, 0, st>>>(perm, p, count, scale_f, oct, grey);
    return 1;
}
int wn_launch_stats(const float *data, size_t count, double *partial, cudaStream_t st)
{
    k_stats<<<WN_STATS_BLOCKS, 256, 0, st>>>(data, count, partial);
    return 1;
}
