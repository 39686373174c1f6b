/*
 * wn_oracle.h -- CPU oracle for the wavelet-noise hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the reference algorithm
 * (Jason9339/Wavelet-Noise-in-ray-tracing: WaveletNoise.cpp, experient/PerlinNoise.hpp,
 * texture.h, experient/main.cpp) plus the parts of libstdc++ 13 the reference leans on
 * (std::mt19937, std::normal_distribution<float>, std::shuffle).  It exists so that the
 * CUDA product path can be checked; nothing in the product path may call it.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs load this library.
 *
 * Parity status: PINNED.  tests/test_oracle_cpu.py checks this file against
 *   (1) the 15 golden .raw images the reference ships (tests/golden/, byte-exact),
 *   (2) the known-answer vectors of SURVEY.md Appendix B,
 *   (3) oracle/_ref/libwnref.so = the unmodified reference sources compiled here
 *       (when present).
 * UNPINNED parts (no reference code or artefact exists; restated from the paper):
 *   orc_multiband3d_* (Cook & DeRose App. 2 composition over the reference evaluate3D)
 *   orc_odd_offset3d  (Cook & DeRose App. 1 final step, omitted by the reference).
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fopenmp -shared -fPIC wn_oracle.c -lm
 * (-ffp-contract=off is mandatory: the reference results are un-fused IEEE float.)
 */
#ifndef WN_ORACLE_H
#define WN_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- libstdc++ restatements ------------------------------------------------------------ */

typedef struct {
    uint32_t mt[624];
    int      idx;
    /* std::normal_distribution<float> carries one cached variate between calls */
    int      has_saved;
    float    saved;
    uint64_t draws;       /* number of 32-bit outputs consumed so far */
} orc_rng;

void     orc_rng_seed(orc_rng *g, uint32_t seed);
uint32_t orc_rng_u32(orc_rng *g);
float    orc_rng_canonical(orc_rng *g);            /* generate_canonical<float,24>          */
float    orc_rng_normal(orc_rng *g);               /* normal_distribution<float>(0,1)       */
void     orc_gaussian_fill(orc_rng *g, float *out, size_t count);
/* std::shuffle(iota(256), mt19937(seed)) duplicated to 512 entries (PerlinNoise.hpp:29-34) */
void     orc_perlin_perm(uint32_t seed, int32_t perm512[512]);

/* ---- tile construction (WaveletNoise.cpp:37-108, 142-183) ------------------------------ */

int  orc_adjust_tile_size(int n);                               /* odd n -> n+1 (cpp:22-25)  */
void orc_downsample1d(const float *from, float *to, int n);     /* cpp:37-48                 */
void orc_upsample1d(const float *from, float *to, int n);       /* cpp:51-66                 */
/* From a caller-supplied Gaussian field R (n^d floats) to the noise tile N = R - R(down,up). */
void orc_tile2d_from_field(const float *R, float *N, int n);
void orc_tile3d_from_field(const float *R, float *N, int n);
/* Whole generateNoiseTile{2,3}D: fill from the rng (continuing its stream), then filter.    */
void orc_generate_tile2d(orc_rng *g, float *N, int n);
void orc_generate_tile3d(orc_rng *g, float *N, int n);
/* Paper Appendix 1 odd-offset step (NOT in the reference; unpinned). In place.              */
void orc_odd_offset3d(float *N, int n);

/* ---- evaluation (WaveletNoise.cpp:111-140, 185-265) ------------------------------------ */

float orc_eval2d(const float *N, int n, const float p[2]);
float orc_eval3d(const float *N, int n, const float p[3]);
float orc_eval3d_projected(const float *N, int n, const float p[3], const float normal[3]);

/* Batch forms (OpenMP-parallel when threads > 1; threads <= 0 -> all cores).
 * points are AoS; out[i] = eval(p_i * pre_scale) * post_scale, pre_scale a single float mul. */
void orc_eval2d_points(const float *N, int n, const float *p_aos, size_t count,
                       float pre_scale, float post_scale, float *out, int threads);
void orc_eval3d_points(const float *N, int n, const float *p_aos, size_t count,
                       float pre_scale, float post_scale, float *out, int threads);
void orc_eval3d_projected_points(const float *N, int n, const float *p_aos,
                                 const float *normal_aos, int normal_is_shared, size_t count,
                                 float pre_scale, float post_scale, float *out, int threads);

/* Lattice = axis-aligned grid given by three coordinate arrays; sample (i,j,k) sits at
 * (xs[i], ys[j], zs[k]); out index i + nx*(j + ny*k).
 * out = post_scale * sum_b weights[b] * evaluate3D(p * band_scale[b]).  (paper App. 2 WMultibandNoise
 * restated over the reference's evaluate3D; nbands=1, weights={1} is the reference's single band.) */
void orc_multiband3d_lattice(const float *N, int n,
                             const float *xs, int nx, const float *ys, int ny,
                             const float *zs, int nz,
                             const float *band_scale, const float *weights, int nbands,
                             float post_scale, float *out, int threads);
/* paper Appendix 2 WMultibandNoise (s = scale cut-off, normal NULL -> WNoise); parity unpinned */
float orc_wmultiband(const float *N, int n, const float p[3], float s, const float *normal, int first_band, int nbands,
                     const float *w);
void orc_wmultiband_points(const float *N, int n, const float *p_aos, size_t count, float s, const float *normal,
                           int first_band, int nbands, const float *w, float *out, int threads);
void orc_multiband3d_points(const float *N, int n, const float *p_aos, size_t count,
                            const float *band_scale, const float *weights, int nbands,
                            float post_scale, float *out, int threads);
void orc_eval2d_lattice(const float *N, int n, const float *xs, int nx, const float *ys, int ny,
                        float pre_scale, float post_scale, float *out, int threads);
void orc_eval3d_projected_lattice(const float *N, int n, const float *xs, int nx,
                                  const float *ys, int ny, float z, const float normal[3],
                                  float pre_scale, float post_scale, float *out, int threads);

/* Integer tap dump for bit-exact index tests: 27 tile indices of evaluate3D(p). */
void orc_eval3d_taps(int n, const float p[3], int32_t idx27[27]);

/* ---- Perlin (experient/PerlinNoise.hpp:13-60, perlin.h:42-72) -------------------------- */

double orc_perlin_noise(const int32_t perm512[512], double x, double y, double z);
void   orc_perlin_points(const int32_t perm512[512], const float *p_aos, size_t count,
                         float pre_scale, float *out, int threads);
void   orc_perlin_lattice(const int32_t perm512[512], const float *xs, int nx,
                          const float *ys, int ny, const float *zs, int nz,
                          float *out, int threads);

/* ---- texture remaps (texture.h:37-43, 67-107) ------------------------------------------ */

/* wavelet_texture::value, 3D branch: p (float xyz) -> grey in [0,1] (returned as double) */
double orc_wavelet_texture_value(const float *N, int n, const float p[3], double scale, int octave);
/* wavelet_texture::value, 2D branch (texture.h:86-99) */
double orc_wavelet_texture2d_value(const float *N2, int n, const float p[3], double scale, int octave);
/* noise_texture::value: Perlin grey */
double orc_perlin_texture_value(const int32_t perm512[512], const float p[3], double scale, int octave);

/* ---- stats (WaveletNoise.cpp:268-288) -------------------------------------------------- */
typedef struct { float avg, var, min_val, max_val; } orc_stats;
void orc_calculate_stats(const float *data, size_t count, orc_stats *out);

float    orc_logf_restated(float x);      /* glibc logf restatement (pins the device table of csrc/wn_rng.cu) */
uint64_t orc_logf_mismatches(float lo, float hi, uint32_t step);
uint64_t orc_fnv1a64(const void *bytes, size_t len);
int      orc_max_threads(void);
void     orc_set_threads(int threads);   /* OpenMP team size for the tile sweeps */

#ifdef __cplusplus
}
#endif
#endif
