/*
 * wn_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY, see wn_oracle.h for the contract).
 *
 * Every function cites the reference lines it restates (paths relative to the reference
 * repository root).  Arithmetic is kept in the reference's exact operation order and
 * width: float where the reference uses float, double where it uses double, no fused
 * multiply-add (compile with -ffp-contract=off).
 */
#include "wn_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* =========================================================================================
 * libstdc++ pieces
 * ========================================================================================= */

/* std::mt19937 (ISO C++ [rand.eng.mers]; identical on every conforming library).
 * Used by WaveletNoise.h:47 (rng) and PerlinNoise.hpp:32. */
void orc_rng_seed(orc_rng *g, uint32_t seed)
{
    g->mt[0] = seed;
    for (int i = 1; i < 624; ++i)
        g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
    g->idx = 624;
    g->has_saved = 0;
    g->saved = 0.0f;
    g->draws = 0;
}

static void orc_rng_refill(orc_rng *g)
{
    uint32_t *mt = g->mt;
    for (int k = 0; k < 624; ++k) {
        uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
        uint32_t v = mt[(k + 397) % 624] ^ (y >> 1);
        if (y & 1u) v ^= 0x9908b0dfu;
        mt[k] = v;
    }
    g->idx = 0;
}

uint32_t orc_rng_u32(orc_rng *g)
{
    if (g->idx >= 624) orc_rng_refill(g);
    uint32_t y = g->mt[g->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    g->draws++;
    return y;
}

/* std::generate_canonical<float,24>(mt19937) -- libstdc++ 13 bits/random.tcc:3346-3382.
 * One 32-bit draw, converted to float (round-to-nearest), divided by 2^32; a result that
 * rounds up to 1.0f is replaced by nextafterf(1,0). */
float orc_rng_canonical(orc_rng *g)
{
    float sum = (float)orc_rng_u32(g) * 1.0f;
    float ret = sum / 4294967296.0f;
    if (ret >= 1.0f) ret = nextafterf(1.0f, 0.0f);
    return ret;
}

/* std::normal_distribution<float>(0,1)::operator() -- libstdc++ 13 bits/random.tcc:1811-1843.
 * Marsaglia polar method.  Returns y*mult first and keeps x*mult for the next call.
 * `2.0f*u - 1.0` is evaluated in double in libstdc++ (the literal 1.0 is double) and then
 * narrowed; the value is exactly representable either way, the double form is kept anyway. */
float orc_rng_normal(orc_rng *g)
{
    float ret;
    if (g->has_saved) {
        g->has_saved = 0;
        ret = g->saved;
    } else {
        float x, y, r2;
        do {
            x = (float)((double)(2.0f * orc_rng_canonical(g)) - 1.0);
            y = (float)((double)(2.0f * orc_rng_canonical(g)) - 1.0);
            r2 = x * x + y * y;
        } while ((double)r2 > 1.0 || (double)r2 == 0.0);
        const float mult = sqrtf(-2.0f * logf(r2) / r2);
        g->saved = x * mult;
        g->has_saved = 1;
        ret = y * mult;
    }
    ret = ret * 1.0f + 0.0f;           /* stddev 1, mean 0 */
    return ret;
}

/* WaveletNoise.cpp:74-77 / :146-147 -- tile cells are filled in memory order. */
void orc_gaussian_fill(orc_rng *g, float *out, size_t count)
{
    for (size_t i = 0; i < count; ++i) out[i] = orc_rng_normal(g);
}

/* uniform_int_distribution<unsigned long>{0,range-1}(mt19937): Lemire's nearly-divisionless
 * method, libstdc++ 13 bits/uniform_int_dist.h:255-283 (32-bit generator, 64-bit product). */
static uint32_t orc_lemire(orc_rng *g, uint32_t range)
{
    uint64_t product = (uint64_t)orc_rng_u32(g) * (uint64_t)range;
    uint32_t low = (uint32_t)product;
    if (low < range) {
        uint32_t threshold = (uint32_t)(0u - range) % range;
        while (low < threshold) {
            product = (uint64_t)orc_rng_u32(g) * (uint64_t)range;
            low = (uint32_t)product;
        }
    }
    return (uint32_t)(product >> 32);
}

/* PerlinNoise.hpp:29-34 == perlin.h:34-39: iota(256), std::shuffle with a fresh mt19937(seed),
 * then the table is appended to itself.  std::shuffle = libstdc++ 13 bits/stl_algo.h:3742-3806
 * (two swap positions per generator call). */
void orc_perlin_perm(uint32_t seed, int32_t perm512[512])
{
    orc_rng g;
    orc_rng_seed(&g, seed);
    int32_t *p = perm512;
    const uint32_t count = 256;
    for (uint32_t i = 0; i < count; ++i) p[i] = (int32_t)i;

    uint32_t i = 1;
    if ((count % 2u) == 0u) {                       /* even length: one single swap up front */
        uint32_t j = orc_lemire(&g, 2u);
        int32_t t = p[i]; p[i] = p[j]; p[j] = t;
        ++i;
    }
    while (i != count) {
        const uint32_t swap_range = i + 1u;
        const uint32_t b1 = swap_range + 1u;
        const uint32_t x = orc_lemire(&g, swap_range * b1);
        const uint32_t j0 = x / b1, j1 = x % b1;
        int32_t t = p[i]; p[i] = p[j0]; p[j0] = t; ++i;
        t = p[i]; p[i] = p[j1]; p[j1] = t; ++i;
    }
    for (uint32_t k = 0; k < count; ++k) p[count + k] = p[k];
}

/* =========================================================================================
 * Tile construction
 * ========================================================================================= */

/* WaveletNoise.cpp:11-16 -- Cook & DeRose Appendix 1 analysis filter, 32 taps, index 16 is
 * tap k=0.  (Values are published filter coefficients; note entry 28 is 0.003546, not 0.003545.) */
static const float ORC_A[32] = {
    0.000334f, -0.001528f,  0.000410f,  0.003545f, -0.000938f, -0.008233f,  0.002172f,  0.019120f,
   -0.005040f, -0.044412f,  0.011655f,  0.103311f, -0.025936f, -0.243780f,  0.033979f,  0.655340f,
    0.655340f,  0.033979f, -0.243780f, -0.025936f,  0.103311f,  0.011655f, -0.044412f, -0.005040f,
    0.019120f,  0.002172f, -0.008233f, -0.000938f,  0.003546f,  0.000410f, -0.001528f,  0.000334f
};
/* WaveletNoise.cpp:18 -- refinement filter, index 2 is tap 0. */
static const float ORC_P[4] = { 0.25f, 0.75f, 0.75f, 0.25f };

/* WaveletNoise.cpp:31-34 */
static inline int orc_mod(int x, int n)
{
    int m = x % n;
    return (m < 0) ? m + n : m;
}

/* WaveletNoise.cpp:20-26 */
int orc_adjust_tile_size(int n) { return (n % 2 != 0) ? n + 1 : n; }

/* WaveletNoise.cpp:37-48 -- to[i] = sum_{k=-16}^{15} A[16+k] * from[Mod(2i+k, n)], summed in k order. */
void orc_downsample1d(const float *from, float *to, int n)
{
    for (int i = 0; i < n / 2; ++i) {
        float sum = 0.0f;
        for (int k = -16; k < 16; ++k)
            sum += ORC_A[16 + k] * from[orc_mod(2 * i + k, n)];
        to[i] = sum;
    }
}

/* WaveletNoise.cpp:51-66 -- k runs over {i/2, i/2+1}; tap index i-2k lies in [-2,1] always. */
void orc_upsample1d(const float *from, float *to, int n)
{
    const int half = n / 2;
    for (int i = 0; i < n; ++i) {
        float sum = 0.0f;
        for (int k = i / 2; k <= i / 2 + 1; ++k) {
            int pidx = i - 2 * k;
            if (pidx >= -2 && pidx <= 1)
                sum += ORC_P[2 + pidx] * from[orc_mod(k, half)];
        }
        to[i] = sum;
    }
}

/* One separable sweep: every line of `count` lines (start offsets enumerated by the caller)
 * is copied out, down- then up-sampled, and written to the same place in dst. */
static void orc_sweep_line(const float *src, float *dst, int n, size_t base, size_t stride,
                           float *lin, float *lds, float *lout)
{
    for (int i = 0; i < n; ++i) lin[i] = src[base + (size_t)i * stride];
    orc_downsample1d(lin, lds, n);
    orc_upsample1d(lds, lout, n);
    for (int i = 0; i < n; ++i) dst[base + (size_t)i * stride] = lout[i];
}

/* WaveletNoise.cpp:69-108 minus the fill: X rows R->T1, Y columns T1->T2, N = R - T2. */
void orc_tile2d_from_field(const float *R, float *N, int n)
{
    size_t cnt = (size_t)n * n;
    float *t1 = (float *)malloc(cnt * sizeof(float));
    float *t2 = (float *)malloc(cnt * sizeof(float));
    float *lin = (float *)malloc(3 * (size_t)n * sizeof(float));
    float *lds = lin + n, *lout = lin + 2 * n;
    for (int iy = 0; iy < n; ++iy) orc_sweep_line(R, t1, n, (size_t)iy * n, 1, lin, lds, lout);
    for (int ix = 0; ix < n; ++ix) orc_sweep_line(t1, t2, n, (size_t)ix, (size_t)n, lin, lds, lout);
    for (size_t i = 0; i < cnt; ++i) N[i] = R[i] - t2[i];
    free(lin); free(t2); free(t1);
}

/* WaveletNoise.cpp:142-183 minus the fill: X: R->T1, Y: T1->T2, Z: T2->T1, N = R - T1. */
void orc_tile3d_from_field(const float *R, float *N, int n)
{
    size_t n2 = (size_t)n * n, cnt = n2 * n;
    float *t1 = (float *)malloc(cnt * sizeof(float));
    float *t2 = (float *)malloc(cnt * sizeof(float));
#ifdef _OPENMP
#pragma omp parallel
#endif
    {
        float *lin = (float *)malloc(3 * (size_t)n * sizeof(float));
        float *lds = lin + n, *lout = lin + 2 * n;
#ifdef _OPENMP
#pragma omp for
#endif
        for (int line = 0; line < n * n; ++line) {            /* (iz, iy) rows along x */
            int iz = line / n, iy = line % n;
            orc_sweep_line(R, t1, n, (size_t)iy * n + (size_t)iz * n2, 1, lin, lds, lout);
        }
#ifdef _OPENMP
#pragma omp for
#endif
        for (int line = 0; line < n * n; ++line) {            /* (iz, ix) columns along y */
            int iz = line / n, ix = line % n;
            orc_sweep_line(t1, t2, n, (size_t)ix + (size_t)iz * n2, (size_t)n, lin, lds, lout);
        }
#ifdef _OPENMP
#pragma omp for
#endif
        for (int line = 0; line < n * n; ++line) {            /* (iy, ix) columns along z */
            int iy = line / n, ix = line % n;
            orc_sweep_line(t2, t1, n, (size_t)ix + (size_t)iy * n, n2, lin, lds, lout);
        }
        free(lin);
    }
    for (size_t i = 0; i < cnt; ++i) N[i] = R[i] - t1[i];
    free(t2); free(t1);
}

void orc_generate_tile2d(orc_rng *g, float *N, int n)
{
    size_t cnt = (size_t)n * n;
    float *R = (float *)malloc(cnt * sizeof(float));
    orc_gaussian_fill(g, R, cnt);
    orc_tile2d_from_field(R, N, n);
    free(R);
}

void orc_generate_tile3d(orc_rng *g, float *N, int n)
{
    size_t cnt = (size_t)n * n * n;
    float *R = (float *)malloc(cnt * sizeof(float));
    orc_gaussian_fill(g, R, cnt);
    orc_tile3d_from_field(R, N, n);
    free(R);
}

/* Cook & DeRose 2005, Appendix 1, last step of GenerateNoiseTile (the reference stops before it,
 * WaveletNoise.cpp:179-182): offset = n/2, made odd; temp[ix*n*n + iy*n + iz] =
 * noise[Mod(ix+o) + Mod(iy+o)*n + Mod(iz+o)*n*n]; noise[i] += temp[i].  The transposed
 * destination index is in the paper's listing and is kept. UNPINNED. */
void orc_odd_offset3d(float *N, int n)
{
    size_t n2 = (size_t)n * n, cnt = n2 * n;
    float *tmp = (float *)malloc(cnt * sizeof(float));
    int off = n / 2;
    if (off % 2 == 0) off++;
    for (int ix = 0; ix < n; ++ix)
        for (int iy = 0; iy < n; ++iy)
            for (int iz = 0; iz < n; ++iz)
                tmp[(size_t)ix * n2 + (size_t)iy * n + iz] =
                    N[orc_mod(ix + off, n) + (size_t)orc_mod(iy + off, n) * n +
                      (size_t)orc_mod(iz + off, n) * n2];
    for (size_t i = 0; i < cnt; ++i) N[i] += tmp[i];
    free(tmp);
}

/* =========================================================================================
 * Evaluation
 * ========================================================================================= */

/* Quadratic B-spline weights for one axis, WaveletNoise.cpp:121-127 / :194-200. */
static inline void orc_basis(float p, int *mid, float w[3])
{
    *mid = (int)ceilf(p - 0.5f);
    float t = (float)*mid - (p - 0.5f);
    w[0] = t * t / 2.0f;
    w[2] = (1.0f - t) * (1.0f - t) / 2.0f;
    w[1] = 1.0f - w[0] - w[2];
}

/* WaveletNoise.cpp:111-140 */
float orc_eval2d(const float *N, int n, const float p[2])
{
    if (n == 0) return 0.0f;
    int mid[2]; float w[2][3];
    for (int i = 0; i < 2; ++i) orc_basis(p[i], &mid[i], w[i]);
    float result = 0.0f;
    for (int fy = -1; fy <= 1; ++fy)
        for (int fx = -1; fx <= 1; ++fx) {
            float weight = w[0][fx + 1] * w[1][fy + 1];
            int cx = orc_mod(mid[0] + fx, n), cy = orc_mod(mid[1] + fy, n);
            result += weight * N[cx + cy * n];
        }
    return result;
}

/* WaveletNoise.cpp:185-215 */
float orc_eval3d(const float *N, int n, const float p[3])
{
    if (n == 0) return 0.0f;
    int mid[3]; float w[3][3];
    for (int i = 0; i < 3; ++i) orc_basis(p[i], &mid[i], w[i]);
    float result = 0.0f;
    for (int fz = -1; fz <= 1; ++fz)
        for (int fy = -1; fy <= 1; ++fy)
            for (int fx = -1; fx <= 1; ++fx) {
                float weight = w[0][fx + 1] * w[1][fy + 1] * w[2][fz + 1];
                int cx = orc_mod(mid[0] + fx, n), cy = orc_mod(mid[1] + fy, n),
                    cz = orc_mod(mid[2] + fz, n);
                result += weight * N[cx + cy * n + cz * n * n];
            }
    return result;
}

void orc_eval3d_taps(int n, const float p[3], int32_t idx27[27])
{
    int mid[3]; float w[3];
    for (int i = 0; i < 3; ++i) orc_basis(p[i], &mid[i], w);
    int t = 0;
    for (int fz = -1; fz <= 1; ++fz)
        for (int fy = -1; fy <= 1; ++fy)
            for (int fx = -1; fx <= 1; ++fx)
                idx27[t++] = orc_mod(mid[0] + fx, n) + orc_mod(mid[1] + fy, n) * n +
                             orc_mod(mid[2] + fz, n) * n * n;
}

/* WaveletNoise.cpp:218-265 */
float orc_eval3d_projected(const float *N, int n, const float p[3], const float normal[3])
{
    if (n == 0) return 0.0f;
    float result = 0.0f;
    int c[3], lo[3], hi[3];
    for (int i = 0; i < 3; ++i) {
        float support = 3.0f * fabsf(normal[i]) + 3.0f * sqrtf((1.0f - normal[i] * normal[i]) / 2.0f);
        lo[i] = (int)ceilf(p[i] - support);
        hi[i] = (int)floorf(p[i] + support);
    }
    for (c[2] = lo[2]; c[2] <= hi[2]; ++c[2])
        for (c[1] = lo[1]; c[1] <= hi[1]; ++c[1])
            for (c[0] = lo[0]; c[0] <= hi[0]; ++c[0]) {
                float dot = 0.0f;
                for (int i = 0; i < 3; ++i) dot += normal[i] * (p[i] - (float)c[i]);
                float weight = 1.0f;
                for (int i = 0; i < 3; ++i) {
                    float t = ((float)c[i] + normal[i] * dot / 2.0f) - (p[i] - 1.5f);
                    if (t <= 0.0f || t >= 3.0f) { weight = 0.0f; break; }
                    float t1 = t - 1.0f, t2 = 2.0f - t, t3 = 3.0f - t;
                    if (t < 1.0f)       weight *= (t * t / 2.0f);
                    else if (t < 2.0f)  weight *= (1.0f - (t1 * t1 + t2 * t2) / 2.0f);
                    else                weight *= (t3 * t3 / 2.0f);
                }
                if ((double)weight > 1e-6) {
                    int idx = orc_mod(c[0], n) + orc_mod(c[1], n) * n + orc_mod(c[2], n) * n * n;
                    result += weight * N[idx];
                }
            }
    return result;
}

static int orc_threads(int threads)
{
#ifdef _OPENMP
    return threads > 0 ? threads : omp_get_max_threads();
#else
    (void)threads; return 1;
#endif
}

void orc_set_threads(int threads)
{
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#else
    (void)threads;
#endif
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* experient/main.cpp:20-28 shape: the coordinate is scaled by one float multiply
 * ((u*s)*2 == u*(2s) exactly), evaluated, then multiplied by the float 1/sqrt(var). */
void orc_eval2d_points(const float *N, int n, const float *p, size_t count, float pre, float post,
                       float *out, int threads)
{
    int nt = orc_threads(threads); (void)nt;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nt) schedule(static)
#endif
    for (long long i = 0; i < (long long)count; ++i) {
        float q[2] = { p[2 * i] * pre, p[2 * i + 1] * pre };
        out[i] = orc_eval2d(N, n, q) * post;
    }
}

void orc_eval3d_points(const float *N, int n, const float *p, size_t count, float pre, float post,
                       float *out, int threads)
{
    int nt = orc_threads(threads); (void)nt;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nt) schedule(static)
#endif
    for (long long i = 0; i < (long long)count; ++i) {
        float q[3] = { p[3 * i] * pre, p[3 * i + 1] * pre, p[3 * i + 2] * pre };
        out[i] = orc_eval3d(N, n, q) * post;
    }
}

void orc_eval3d_projected_points(const float *N, int n, const float *p, const float *nrm,
                                 int shared, size_t count, float pre, float post, float *out,
                                 int threads)
{
    int nt = orc_threads(threads); (void)nt;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nt) schedule(static)
#endif
    for (long long i = 0; i < (long long)count; ++i) {
        float q[3] = { p[3 * i] * pre, p[3 * i + 1] * pre, p[3 * i + 2] * pre };
        const float *nn = shared ? nrm : nrm + 3 * i;
        out[i] = orc_eval3d_projected(N, n, q, nn) * post;
    }
}

/* Cook & DeRose App. 2 WMultibandNoise restated over the reference's evaluate3D
 * (WaveletNoise.cpp:185-215): result = sum_b w[b] * WNoise(p * band_scale[b]), accumulated in
 * band order in float, then one multiply by post_scale.  UNPINNED composition;
 * with nbands=1, w={1} it is experient/main.cpp:45-58. */
static inline float orc_multiband_at(const float *N, int n, float x, float y, float z,
                                     const float *bs, const float *w, int nb)
{
    float acc = 0.0f;
    for (int b = 0; b < nb; ++b) {
        float q[3] = { x * bs[b], y * bs[b], z * bs[b] };
        acc += w[b] * orc_eval3d(N, n, q);
    }
    return acc;
}

void orc_multiband3d_lattice(const float *N, int n, const float *xs, int nx, const float *ys, int ny,
                             const float *zs, int nz, const float *bs, const float *w, int nb,
                             float post, float *out, int threads)
{
    int nt = orc_threads(threads); (void)nt;
    long long rows = (long long)ny * nz;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nt) schedule(dynamic, 4)
#endif
    for (long long r = 0; r < rows; ++r) {
        int j = (int)(r % ny), k = (int)(r / ny);
        float *o = out + (size_t)r * nx;
        for (int i = 0; i < nx; ++i)
            o[i] = orc_multiband_at(N, n, xs[i], ys[j], zs[k], bs, w, nb) * post;
    }
}

void orc_multiband3d_points(const float *N, int n, const float *p, size_t count, const float *bs,
                            const float *w, int nb, float post, float *out, int threads)
{
    int nt = orc_threads(threads); (void)nt;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nt) schedule(static)
#endif
    for (long long i = 0; i < (long long)count; ++i)
        out[i] = orc_multiband_at(N, n, p[3 * i], p[3 * i + 1], p[3 * i + 2], bs, w, nb) * post;
}

/* Cook & DeRose 2005, Appendix 2, WMultibandNoise(p, s, normal, firstBand, nbands, w) restated statement by statement
 * over the reference's evaluate3D / evaluate3DProjected (the reference itself has no multiband function):
 *   for (b = 0; b < nbands && s + firstBand + b < 0; b++) { q = 2 p 2^(firstBand+b); result += w[b] * (normal ?
 *       WProjectedNoise(q, normal) : WNoise(q)); }
 *   variance = sum over ALL nbands of w[b]^2;  if (variance) result /= sqrt(variance * (normal ? 0.296 : 0.210));
 * PARITY UNPINNED: no reference code or artefact exists for this composition; pow() is exact for integer exponents. */
float orc_wmultiband(const float *N, int n, const float p[3], float s, const float *normal, int first_band, int nbands,
                     const float *w)
{
    float q[3], result = 0, variance = 0;
    int i, b;
    for (b = 0; b < nbands && s + first_band + b < 0; b++) {
        for (i = 0; i <= 2; i++) q[i] = (float)(2 * p[i] * pow(2, first_band + b));
        result += normal ? w[b] * orc_eval3d_projected(N, n, q, normal) : w[b] * orc_eval3d(N, n, q);
    }
    for (b = 0; b < nbands; b++) variance += w[b] * w[b];
    if (variance) result = (float)(result / sqrt(variance * (normal ? 0.296 : 0.210)));
    return result;
}

void orc_wmultiband_points(const float *N, int n, const float *p, size_t count, float s, const float *normal,
                           int first_band, int nbands, const float *w, float *out, int threads)
{
    int nt = orc_threads(threads); (void)nt;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nt) schedule(static)
#endif
    for (long long i = 0; i < (long long)count; ++i) out[i] = orc_wmultiband(N, n, p + 3 * i, s, normal, first_band, nbands, w);
}

/* experient/main.cpp:18-30 */
void orc_eval2d_lattice(const float *N, int n, const float *xs, int nx, const float *ys, int ny,
                        float pre, float post, float *out, int threads)
{
    int nt = orc_threads(threads); (void)nt;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nt) schedule(static)
#endif
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) {
            float q[2] = { xs[i] * pre, ys[j] * pre };
            out[(size_t)j * nx + i] = orc_eval2d(N, n, q) * post;
        }
}

/* experient/main.cpp:74-87 */
void orc_eval3d_projected_lattice(const float *N, int n, const float *xs, int nx, const float *ys,
                                  int ny, float z, const float normal[3], float pre, float post,
                                  float *out, int threads)
{
    int nt = orc_threads(threads); (void)nt;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nt) schedule(static)
#endif
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) {
            float q[3] = { xs[i] * pre, ys[j] * pre, z * pre };
            out[(size_t)j * nx + i] = orc_eval3d_projected(N, n, q, normal) * post;
        }
}

/* =========================================================================================
 * Perlin -- experient/PerlinNoise.hpp:13-60 (== perlin.h:17-72), all double
 * ========================================================================================= */

static inline double orc_fade(double t) { return t * t * t * (t * (t * 6 - 15) + 10); }
static inline double orc_lerp(double t, double a, double b) { return a + t * (b - a); }
static inline double orc_grad(int hash, double x, double y, double z)
{
    const int h = hash & 15;
    const double u = h < 8 ? x : y;
    const double v = h < 4 ? y : (h == 12 || h == 14) ? x : z;
    return ((h & 1) == 0 ? u : -u) + ((h & 2) == 0 ? v : -v);
}

double orc_perlin_noise(const int32_t p[512], double x, double y, double z)
{
    const int X = (int)floor(x) & 255, Y = (int)floor(y) & 255, Z = (int)floor(z) & 255;
    x -= floor(x); y -= floor(y); z -= floor(z);
    const double u = orc_fade(x), v = orc_fade(y), w = orc_fade(z);
    const int A = p[X] + Y, AA = p[A] + Z, AB = p[A + 1] + Z;
    const int B = p[X + 1] + Y, BA = p[B] + Z, BB = p[B + 1] + Z;
    return orc_lerp(w,
        orc_lerp(v, orc_lerp(u, orc_grad(p[AA], x, y, z),         orc_grad(p[BA], x - 1, y, z)),
                    orc_lerp(u, orc_grad(p[AB], x, y - 1, z),     orc_grad(p[BB], x - 1, y - 1, z))),
        orc_lerp(v, orc_lerp(u, orc_grad(p[AA + 1], x, y, z - 1), orc_grad(p[BA + 1], x - 1, y, z - 1)),
                    orc_lerp(u, orc_grad(p[AB + 1], x, y - 1, z - 1),
                                orc_grad(p[BB + 1], x - 1, y - 1, z - 1))));
}

/* noise_texture-style points: coordinate = p * pre (float multiply), promoted to double. */
void orc_perlin_points(const int32_t perm[512], const float *p, size_t count, float pre, float *out,
                       int threads)
{
    int nt = orc_threads(threads); (void)nt;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nt) schedule(static)
#endif
    for (long long i = 0; i < (long long)count; ++i) {
        float qx = p[3 * i] * pre, qy = p[3 * i + 1] * pre, qz = p[3 * i + 2] * pre;
        out[i] = (float)orc_perlin_noise(perm, (double)qx, (double)qy, (double)qz);
    }
}

/* experient/main.cpp:100-106, 118-124: float coordinates promoted to double, result cast to float. */
void orc_perlin_lattice(const int32_t perm[512], const float *xs, int nx, const float *ys, int ny,
                        const float *zs, int nz, float *out, int threads)
{
    int nt = orc_threads(threads); (void)nt;
    long long rows = (long long)ny * nz;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nt) schedule(static)
#endif
    for (long long r = 0; r < rows; ++r) {
        int j = (int)(r % ny), k = (int)(r / ny);
        for (int i = 0; i < nx; ++i)
            out[(size_t)r * nx + i] =
                (float)orc_perlin_noise(perm, (double)xs[i], (double)ys[j], (double)zs[k]);
    }
}

/* =========================================================================================
 * Texture hooks
 * ========================================================================================= */

/* texture.h:67-107, 3D branch. p holds the hit point's float components (vec3 stores float). */
double orc_wavelet_texture_value(const float *N, int n, const float p[3], double scale, int octave)
{
    float pos[3] = { (float)((double)p[0] * scale), (float)((double)p[1] * scale),
                     (float)((double)p[2] * scale) };
    const float octave_scale = (float)pow(2.0, (double)octave);   /* std::pow(2.0f,int) -> double */
    pos[0] *= octave_scale * 2.0f; pos[1] *= octave_scale * 2.0f; pos[2] *= octave_scale * 2.0f;
    double v = (double)orc_eval3d(N, n, pos);
    const float inv_stddev = 1.0f / sqrtf(0.18402f);
    v *= (double)inv_stddev;
    double c = v / 4.0;
    if (c < -1.0) c = -1.0;
    if (c > 1.0) c = 1.0;
    return 0.5 * (1.0 + c);
}

/* texture.h:86-99, the 2D branch (use_3d = false; never taken by the reference renderer): evaluate2D of the xy
 * components, 1/sqrt(0.19686f), same clamp remap. */
double orc_wavelet_texture2d_value(const float *N2, int n, const float p[3], double scale, int octave)
{
    float pos[2] = { (float)((double)p[0] * scale), (float)((double)p[1] * scale) };
    const float octave_scale = (float)pow(2.0, (double)octave);
    pos[0] *= octave_scale * 2.0f; pos[1] *= octave_scale * 2.0f;
    double v = (double)orc_eval2d(N2, n, pos);
    const float inv_stddev = 1.0f / sqrtf(0.19686f);
    v *= (double)inv_stddev;
    double c = v / 4.0;
    if (c < -1.0) c = -1.0;
    if (c > 1.0) c = 1.0;
    return 0.5 * (1.0 + c);
}

/* texture.h:37-43: p * float(scale) * octave_scale in float (vec3 ops), Perlin in double. */
double orc_perlin_texture_value(const int32_t perm[512], const float p[3], double scale, int octave)
{
    const float octave_scale = (float)pow(2.0, (double)octave);
    const float s = (float)scale;
    float q[3] = { p[0] * s * octave_scale, p[1] * s * octave_scale, p[2] * s * octave_scale };
    double v = orc_perlin_noise(perm, (double)q[0], (double)q[1], (double)q[2]);
    return 0.5 * (1.0 + v);
}

/* =========================================================================================
 * Stats -- WaveletNoise.cpp:268-288
 * ========================================================================================= */
void orc_calculate_stats(const float *data, size_t count, orc_stats *out)
{
    out->avg = 0.0f; out->var = 0.0f;
    out->min_val = 3.402823466e+38f; out->max_val = -3.402823466e+38f;
    if (count == 0) return;
    double sum = 0.0, sum_sq = 0.0;
    for (size_t i = 0; i < count; ++i) {
        float v = data[i];
        sum += v;
        sum_sq += (double)v * v;
        if (v < out->min_val) out->min_val = v;
        if (v > out->max_val) out->max_val = v;
    }
    out->avg = (float)(sum / (double)count);
    out->var = (float)((sum_sq / (double)count) - (double)out->avg * out->avg);
}

/* glibc logf (sysdeps/ieee754/flt-32/e_logf.c, 2.27+; ARM optimized-routines algorithm, N = 16 table) restated.
 * The device Gaussian fill (csrc/wn_rng.cu) carries the same constants; tests compare this function with the
 * libm logf the reference links against, bit for bit, to pin them.  Normal positive inputs only. */
float orc_logf_restated(float x)
{
    static const double invc[16] = {
        0x1.661ec79f8f3bep+0, 0x1.571ed4aaf883dp+0, 0x1.49539f0f010bp+0, 0x1.3c995b0b80385p+0, 0x1.30d190c8864a5p+0,
        0x1.25e227b0b8eap+0, 0x1.1bb4a4a1a343fp+0, 0x1.12358f08ae5bap+0, 0x1.0953f419900a7p+0, 0x1p+0,
        0x1.e608cfd9a47acp-1, 0x1.ca4b31f026aap-1, 0x1.b2036576afce6p-1, 0x1.9c2d163a1aa2dp-1, 0x1.886e6037841edp-1,
        0x1.767dcf5534862p-1 };
    static const double logc[16] = {
        -0x1.57bf7808caadep-2, -0x1.2bef0a7c06ddbp-2, -0x1.01eae7f513a67p-2, -0x1.b31d8a68224e9p-3, -0x1.6574f0ac07758p-3,
        -0x1.1aa2bc79c81p-3, -0x1.a4e76ce8c0e5ep-4, -0x1.1973c5a611cccp-4, -0x1.252f438e10c1ep-5, 0x0p+0,
        0x1.aa5aa5df25984p-5, 0x1.c5e53aa362eb4p-4, 0x1.526e57720db08p-3, 0x1.bc2860d22477p-3, 0x1.1058bc8a07ee1p-2,
        0x1.4043057b6ee09p-2 };
    uint32_t ix;
    memcpy(&ix, &x, 4);
    if (ix == 0x3f800000u) return 0.0f;
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (int)((tmp >> 19) & 15u);
    const int k = (int32_t)tmp >> 23;
    const uint32_t iz = ix - (tmp & 0xff800000u);
    float zf;
    memcpy(&zf, &iz, 4);
    const double z = (double)zf;
    const double r = z * invc[i] - 1;
    const double y0 = logc[i] + (double)k * 0x1.62e42fefa39efp-1;
    const double r2 = r * r;
    double y = 0x1.5575b0be00b6ap-2 * r + -0x1.ffffef20a4123p-2;
    y = -0x1.00ea348b88334p-2 * r2 + y;
    y = y * r2 + (y0 + r);
    return (float)y;
}

/* number of floats in [lo, hi] (bit patterns stepped by `step`) where orc_logf_restated differs from libm logf */
uint64_t orc_logf_mismatches(float lo, float hi, uint32_t step)
{
    uint32_t a, b;
    memcpy(&a, &lo, 4);
    memcpy(&b, &hi, 4);
    uint64_t bad = 0;
    for (uint64_t u = a; u <= b; u += step) {
        uint32_t uu = (uint32_t)u;
        float x;
        memcpy(&x, &uu, 4);
        float p = logf(x), q = orc_logf_restated(x);
        if (memcmp(&p, &q, 4) != 0) ++bad;
    }
    return bad;
}

uint64_t orc_fnv1a64(const void *bytes, size_t len)
{
    const unsigned char *b = (const unsigned char *)bytes;
    uint64_t h = 0xcbf29ce484222325ull;
    for (size_t i = 0; i < len; ++i) { h ^= b[i]; h *= 0x100000001b3ull; }
    return h;
}
