// wn_internal.h -- declarations shared by the translation units of libwn_b200.so.
// Kernels live in wn_tilegen.cu / wn_eval_exact.cu (compiled with -fmad=false: reference operation
// order, un-fused IEEE arithmetic) and wn_multiband_fast.cu (FMA allowed).  wn_capi.cu is the only
// file that implements the extern "C" surface of include/wn_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#define WN_MAX_BANDS 16
#define WN_TILE_PAD 3         // extra cells per row of the x-padded tile replica: cells n, n+1, n+2 repeat 0, 1, 2

// ---- sample-coordinate generators (device pointers) ---------------------------------------------
// Every batch kernel is "for sample s in [first, first+count): out[s-first] = op(coord(s))".
struct WnPointsAoS {            // scattered points, p*pre (one float multiply per component)
    const float *p;             // xyz (or xy) per point
    float pre;
};
struct WnLattice {              // axis-aligned: s = i + nx*(j + ny*k) -> (xs[i], ys[j], zs[k])
    const float *xs, *ys, *zs;
    int nx, ny, nz;
};
struct WnAffine {               // s = i + nu*j -> origin + us[i]*e1 + vs[j]*e2, then *pre
    const float *us, *vs;
    int nu, nv;
    float o[3], e1[3], e2[3];
    float pre;
};

struct WnBands {
    int   nbands;
    float scale[WN_MAX_BANDS];
    float weight[WN_MAX_BANDS];
    float post;
};

struct WnTileView {
    const float *N;
    int n;                      // tile edge
    int pow2;                   // n is a power of two -> Mod is a mask
    const float *Npad;          // 3D only: rows padded to n+WN_TILE_PAD floats, the extra cells wrap around (fast lattice kernel)
};

// ---- stream-ordered scratch (wn_capi.cu) ------------------------------------------------------------
// Filter temporaries, MT19937 draws, axis tables and period blocks come from a memory pool PRIVATE to this library
// (one per device, release threshold = keep): recycled across calls like the default pool would, without touching the
// attributes of the device's default pool that the host application (torch, ...) may be using.  Free with cudaFreeAsync.
cudaError_t wn_scratch_alloc(void **p, size_t bytes, cudaStream_t st);

// ---- tile construction (wn_tilegen.cu) ------------------------------------------------------------
// dst = filter_axis(src) for axis in {0,1,2}; when minuend != nullptr: dst = minuend - filter_axis(src).
// tmp smem sizes are handled inside.  Returns the number of kernels launched.
int wn_launch_filter_axis(const float *src, float *dst, const float *minuend, int n, int dims, int axis,
                          cudaStream_t st);
// paper odd-offset step: dst[x,y,z-transposed] = src + shifted src (see wn_tilegen.cu)
int wn_launch_odd_offset3d(const float *src, float *dst, int n, cudaStream_t st);

// ---- exact evaluators (wn_eval_exact.cu) ----------------------------------------------------------
int wn_launch_eval2d_points(WnTileView t, WnPointsAoS c, size_t first, size_t count, float post, float *out, cudaStream_t st);
int wn_launch_eval2d_lattice(WnTileView t, WnLattice c, float pre, size_t first, size_t count, float post, float *out, cudaStream_t st);
int wn_launch_mb3d_points(WnTileView t, WnPointsAoS c, WnBands b, size_t first, size_t count, float *out, cudaStream_t st);
int wn_launch_mb3d_lattice_exact(WnTileView t, WnLattice c, WnBands b, size_t first, size_t count, float *out, cudaStream_t st);
int wn_launch_mb3d_affine(WnTileView t, WnAffine c, WnBands b, size_t first, size_t count, float *out, cudaStream_t st);
// projected: normals == nullptr -> shared normal `nrm`
int wn_launch_proj_points(WnTileView t, WnPointsAoS c, const float *normals, const float nrm[3], size_t first,
                          size_t count, float post, float *out, cudaStream_t st);
int wn_launch_proj_affine(WnTileView t, WnAffine c, const float nrm[3], size_t first, size_t count, float post,
                          float *out, cudaStream_t st);
// fast != 0: FP32 arithmetic with FMAs (<= 1e-5 * range of the FP64 kernel), else the reference's FP64 un-fused order
int wn_launch_perlin_points(const int32_t *perm, WnPointsAoS c, size_t first, size_t count, float *out, int fast, cudaStream_t st);
int wn_launch_perlin_lattice(const int32_t *perm, WnLattice c, size_t first, size_t count, float *out, int fast, cudaStream_t st);
int wn_launch_perlin_affine(const int32_t *perm, WnAffine c, size_t first, size_t count, float *out, int fast, cudaStream_t st);
// paper App. 2 WMultibandNoise on points: b.scale[k] = 2^(firstBand+k), b.weight[k] = w[k]; normal = host pointer or nullptr
int wn_launch_wmultiband(WnTileView t, const float *p, size_t count, WnBands b, int nb_active, const float *normal,
                         double denom, float *out, cudaStream_t st);
int wn_launch_perlin_points_f64(const int32_t *perm, const double *p, size_t count, double *out, cudaStream_t st);
// texture hooks: scale (double) and octave as in texture.h
int wn_launch_wavelet_texture(WnTileView t, const float *p, size_t count, double scale, float oct2, float inv_std,
                              float *grey, cudaStream_t st);
int wn_launch_wavelet_texture2d(WnTileView t, const float *p, size_t count, double scale, float oct2, float inv_std,
                                float *grey, cudaStream_t st);
int wn_launch_perlin_texture(const int32_t *perm, const float *p, size_t count, float scale_f, float oct,
                             float *grey, cudaStream_t st);
// stats: partial[] must hold 4*WN_STATS_BLOCKS doubles; result read back by the host
#define WN_STATS_BLOCKS 592
int wn_launch_stats(const float *data, size_t count, double *partial, cudaStream_t st);

// ---- fast multiband lattice (wn_multiband_fast.cu) ------------------------------------------------
// prepare: decides which bands are periodic on this lattice ("folded"), evaluates their period block once.
// run    : computes the z-range [k0, k0+nk) of the lattice into out (out points at sample (0,0,k0)).
// finish : releases the period block.  All stream ordered on `st`.
// c holds DEVICE axis pointers; h_xs / h_ys / h_zs are the same axes on the host (brick planning, period detection).
struct WnFastPlan {
    WnBands direct;             // bands evaluated per sample
    unsigned char direct_rows[WN_MAX_BANDS];   // their rows in the call's axis tables (= original band indices)
    float *P;                   // sum of the folded bands on the period block Lx x Ly x Lz, or nullptr
    int Lx, Ly, Lz;
    float4 *tab;                // axis tables of the whole call: [x | y | z] x tab_bands rows (see WnTabs)
    int tab_bands, sx, sy, sz;
    int owns_tab;               // the top-level plan frees the tables
    int pdl;                    // launch this plan's kernels with programmatic dependent launch (set before prepare / run)
    void *host_axes;            // host copy of the tables (period detection, brick planning), shared by nested plans
    void **axes_cache;          // set before prepare: the tile's plan-cache slot (nullptr: no caching across calls)
};
int  wn_mb3d_fast_prepare(WnTileView t, WnLattice c, const float *h_xs, const float *h_ys, const float *h_zs, WnBands b,
                          WnFastPlan *plan, cudaStream_t st, int depth = 0);
int  wn_mb3d_fast_run(WnTileView t, WnLattice c, const float *h_ys, const float *h_zs, const WnBands &all_bands,
                      const unsigned char *all_rows, const WnFastPlan *plan, int k0, int nk, float *out, cudaStream_t st);
void wn_mb3d_fast_finish(WnFastPlan *plan, cudaStream_t st);
void wn_mb3d_fast_detach(WnFastPlan *plan, void **tab, void **P);
void wn_mb3d_fast_cache_free(void *slot);      // releases a plan-cache slot (tile destruction)
// host-only fold decision of the top level (diagnostics / CPU tests); returns the number of folded bands
int  wn_mb3d_fast_plan_host(const float *h_xs, int nx, const float *h_ys, int ny, const float *h_zs, int nz, WnBands b,
                            int tile_n, int *folded, int block[3]);

// diagnostics: axis-table entries of `count` coordinates at one band scale (all device pointers); kernels launched
int wn_mb3d_debug_axis_table(const float *coords, int count, float scale, float4 *entries, cudaStream_t st);

// 3D tile -> x-padded replica (row pitch n+WN_TILE_PAD, the extra cells wrap around)
int wn_launch_pad_tile(const float *N, float *Npad, int n, cudaStream_t st);

// ---- device Gaussian fill (wn_rng.cu) ---------------------------------------------------------------
// Fills out[0..count) with the libstdc++ normal_distribution<float>(mt19937(seed)) sequence, bit-identical to the
// host objects.  accepted (device, 16 bytes): [0] = accepted polar attempts -- the fill is complete iff
// 2 * accepted >= count (check after synchronising; retry with a larger margin_permille otherwise); [1] = index of the
// attempt that produced the last output (raw MT draws consumed = 2 * ([1] + 1)).
// Scratch is stream-ordered (cudaMallocAsync).  Returns kernels launched, < 0 on error.
int wn_launch_gaussian_fill(unsigned seed, float *out, size_t count, unsigned long long *accepted, int margin_permille,
                            cudaStream_t st);
