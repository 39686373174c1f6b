// ref_shim.cpp -- extern "C" access to the UNMODIFIED reference classes.  TEST INFRASTRUCTURE ONLY.
//
// This file contains no reference code.  It is compiled together with the reference's own
// sources *where they lie* under $(REF) (= /root/reference) by oracle/Makefile:
//     $(REF)/experient/WaveletNoise.cpp   (class WaveletNoise)
//     $(REF)/experient/PerlinNoise.hpp    (class PerlinNoise)
//     $(REF)/texture.h (+ vec3.h, perlin.h ...)  (wavelet_texture / noise_texture value() hooks)
// into oracle/_ref/libwnref.so, which is git-ignored but travels to the GPU box.  It is used
// to validate oracle/wn_oracle.c and as the "reference" CPU baseline in bench.py.
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "WaveletNoise.h"      // from $(REF)/experient
#include "PerlinNoise.hpp"     // from $(REF)/experient
#ifdef WNREF_WITH_TEXTURE
#include "texture.h"           // from $(REF)
#endif

static int pick_threads(int t)
{
#ifdef _OPENMP
    return t > 0 ? t : omp_get_max_threads();
#else
    (void)t; return 1;
#endif
}

extern "C" {

void *ref_wn_create(int n, unsigned seed) { return new WaveletNoise(n, seed); }
void  ref_wn_destroy(void *h) { delete static_cast<WaveletNoise *>(h); }
void  ref_wn_generate2d(void *h) { static_cast<WaveletNoise *>(h)->generateNoiseTile2D(); }
void  ref_wn_generate3d(void *h) { static_cast<WaveletNoise *>(h)->generateNoiseTile3D(); }
int   ref_wn_tile_size(void *h) { return static_cast<WaveletNoise *>(h)->getTileSize(); }
size_t ref_wn_tile_count(void *h) { return static_cast<WaveletNoise *>(h)->getNoiseCoefficients().size(); }
void  ref_wn_tile_copy(void *h, float *out)
{
    const std::vector<float> &v = static_cast<WaveletNoise *>(h)->getNoiseCoefficients();
    std::memcpy(out, v.data(), v.size() * sizeof(float));
}

// out[i] = evaluate*(p_i * pre) * post, the adapter shape of experient/main.cpp
void ref_wn_eval2d_points(void *h, const float *p, size_t count, float pre, float post, float *out, int threads)
{
    const WaveletNoise *w = static_cast<WaveletNoise *>(h);
    int nt = pick_threads(threads); (void)nt;
#pragma omp parallel for num_threads(nt) schedule(static)
    for (long long i = 0; i < (long long)count; ++i) {
        float q[2] = { p[2 * i] * pre, p[2 * i + 1] * pre };
        out[i] = w->evaluate2D(q) * post;
    }
}
void ref_wn_eval3d_points(void *h, const float *p, size_t count, float pre, float post, float *out, int threads)
{
    const WaveletNoise *w = static_cast<WaveletNoise *>(h);
    int nt = pick_threads(threads); (void)nt;
#pragma omp parallel for num_threads(nt) schedule(static)
    for (long long i = 0; i < (long long)count; ++i) {
        float q[3] = { p[3 * i] * pre, p[3 * i + 1] * pre, p[3 * i + 2] * pre };
        out[i] = w->evaluate3D(q) * post;
    }
}
void ref_wn_eval3d_projected_points(void *h, const float *p, const float *nrm, int shared, size_t count,
                                    float pre, float post, float *out, int threads)
{
    const WaveletNoise *w = static_cast<WaveletNoise *>(h);
    int nt = pick_threads(threads); (void)nt;
#pragma omp parallel for num_threads(nt) schedule(static)
    for (long long i = 0; i < (long long)count; ++i) {
        float q[3] = { p[3 * i] * pre, p[3 * i + 1] * pre, p[3 * i + 2] * pre };
        out[i] = w->evaluate3DProjected(q, shared ? nrm : nrm + 3 * i) * post;
    }
}
// composed multiband over the reference's own evaluate3D (the reference has no WMultibandNoise)
void ref_wn_multiband3d_lattice(void *h, const float *xs, int nx, const float *ys, int ny, const float *zs, int nz,
                                const float *bs, const float *wts, int nb, float post, float *out, int threads)
{
    const WaveletNoise *w = static_cast<WaveletNoise *>(h);
    int nt = pick_threads(threads); (void)nt;
    long long rows = (long long)ny * nz;
#pragma omp parallel for num_threads(nt) schedule(dynamic, 4)
    for (long long r = 0; r < rows; ++r) {
        int j = (int)(r % ny), k = (int)(r / ny);
        for (int i = 0; i < nx; ++i) {
            float acc = 0.0f;
            for (int b = 0; b < nb; ++b) {
                float q[3] = { xs[i] * bs[b], ys[j] * bs[b], zs[k] * bs[b] };
                acc += wts[b] * w->evaluate3D(q);
            }
            out[(size_t)r * nx + i] = acc * post;
        }
    }
}

void *ref_perlin_create(unsigned seed) { return new PerlinNoise(seed); }
void  ref_perlin_destroy(void *h) { delete static_cast<PerlinNoise *>(h); }
void  ref_perlin_points(void *h, const float *p, size_t count, float pre, float *out, int threads)
{
    const PerlinNoise *pn = static_cast<PerlinNoise *>(h);
    int nt = pick_threads(threads); (void)nt;
#pragma omp parallel for num_threads(nt) schedule(static)
    for (long long i = 0; i < (long long)count; ++i) {
        float x = p[3 * i] * pre, y = p[3 * i + 1] * pre, z = p[3 * i + 2] * pre;
        out[i] = static_cast<float>(pn->noise(x, y, z));
    }
}
double ref_perlin_noise(void *h, double x, double y, double z)
{
    return static_cast<PerlinNoise *>(h)->noise(x, y, z);
}

#ifdef WNREF_WITH_TEXTURE
void *ref_wavelet_texture_create(double scale, int octave) { return new wavelet_texture(scale, octave, true); }
void *ref_wavelet_texture2d_create(double scale, int octave) { return new wavelet_texture(scale, octave, false); }
void  ref_wavelet_texture_destroy(void *h) { delete static_cast<wavelet_texture *>(h); }
void *ref_perlin_texture_create(double scale, int octave) { return new noise_texture(scale, octave); }
void  ref_perlin_texture_destroy(void *h) { delete static_cast<noise_texture *>(h); }
// grey[i] = value(0,0,p_i).x()  (the three channels are equal)
void ref_texture_values(void *h, const float *p, size_t count, float *grey, int threads)
{
    const texture *t = static_cast<texture *>(h);
    int nt = pick_threads(threads); (void)nt;
#pragma omp parallel for num_threads(nt) schedule(static)
    for (long long i = 0; i < (long long)count; ++i) {
        color c = t->value(0.0, 0.0, point3(p[3 * i], p[3 * i + 1], p[3 * i + 2]));
        grey[i] = c.x();
    }
}
#endif

int ref_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

} // extern "C"
