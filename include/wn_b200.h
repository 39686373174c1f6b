/*
 * wn_b200.h -- C ABI of the B200-native wavelet-noise path (libwn_b200.so).
 *
 * The reference (Jason9339/Wavelet-Noise-in-ray-tracing) has no FFI layer: its boundary is the
 * C++ class `WaveletNoise` (WaveletNoise.h:20-59), `PerlinNoise`/`perlin`
 * (experient/PerlinNoise.hpp:9-61, perlin.h:14-91) and the `texture::value()` hook
 * (texture.h:14-18).  This header is what a binding for that path calls; every entry point
 * names the reference interface it replaces.  The drop-in host layers built on top of it are
 *   - C++   : wavelet-noise-in-ray-tracing_b200/cpp/WaveletNoise.cpp (+ PerlinNoise.hpp, batch drivers)
 *   - Python: wavelet-noise-in-ray-tracing_b200/ (ctypes mirror used by tests/ and bench.py)
 *
 * Conventions
 *   - every function returns 0 on success, a negative WN_E* code on failure, and never throws;
 *     wn_last_error() returns the message of the calling thread's last failure.
 *   - there is NO CPU fallback: without a CUDA device wn_ctx_create fails with WN_ENODEVICE.
 *   - one wn_ctx = one GPU + one stream; one host thread drives a context at a time.
 *     Multi-GPU = one process per GPU (bench.py under torchrun), or one wn_group in one process (below);
 *     samples shard with no exchange.
 *   - `space` says where the caller's sample buffers live:
 *       WN_HOST   : host memory (pinned is faster).  The call copies in, computes, copies out and
 *                   returns after the result is in `out` (large lattices are chunked so the D2H of
 *                   one chunk overlaps the kernel of the next).
 *       WN_DEVICE : device memory on the context's GPU.  Work is only enqueued on the context's
 *                   stream; the caller synchronises (wn_ctx_synchronize or its own stream sync).
 *     Small parameter arrays (coordinate axes, band scales, weights, normals when shared, perm
 *     tables) are always HOST pointers and are copied at call time.
 *   - tile layout: idx = x + n*y + n*n*z, float32 (WaveletNoise.cpp:154,163,172,209).
 *   - lattice / grid output layout: out[i + nx*(j + ny*k)], float32 -- the `.raw` layout of
 *     experient/main.cpp:28-34 for nz = 1.
 */
#ifndef WN_B200_H
#define WN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WN_OK            0
#define WN_EINVAL       -1   /* bad argument                                  */
#define WN_ENODEVICE    -2   /* no CUDA device / wrong architecture           */
#define WN_ECUDA        -3   /* a CUDA runtime call failed                    */
#define WN_ENOMEM       -4
#define WN_ESTATE       -5   /* e.g. evaluating a tile that was never built   */

#define WN_HOST          0
#define WN_DEVICE        1

/* wn_tile_create flags */
#define WN_TILE_DEFAULT      0u
#define WN_TILE_ODD_OFFSET   1u   /* add the paper's odd-offset copy (Cook&DeRose App.1); the
                                     reference omits it (WaveletNoise.cpp:179-182), default off */

/* wn_multiband3d_lattice `mode` */
#define WN_EVAL_FAST     0   /* separable, fused-multiply-add kernel (<= 1e-5 * tile range)      */
#define WN_EVAL_EXACT    1   /* reference operation order, un-fused: bit-identical to the CPU     */

typedef struct wn_ctx    wn_ctx;
typedef struct wn_tile   wn_tile;
typedef struct wn_perlin wn_perlin;
typedef struct wn_rng    wn_rng;

/* mirrors `struct DataStats` (WaveletNoise.h:11-18); count_nan_inf/energy are never written by
 * the reference and are left 0 here as well. */
typedef struct {
    float     avg, var, min_val, max_val;
    long long count_nan_inf;
    float     energy;
} wn_stats;

/* ---- diagnostics -------------------------------------------------------------------------- */
const char *wn_last_error(void);
const char *wn_version(void);
/* number of kernels this library launched since the context was created (bench.py gpu_launches) */
uint64_t    wn_kernel_launches(const wn_ctx *ctx);
/* CUDA-event time of the kernels enqueued by the most recent WN_HOST call on this context (ms) */
float       wn_timing_last_ms(const wn_ctx *ctx);
/* Per-call CUDA-event timing of the MAIN kernel (the launch that writes the output) of WN_DEVICE, WN_EVAL_FAST
 * wn_multiband3d_lattice calls, for roofline reports: enable, run calls, then collect (synchronises the stream; ms[i] =
 * duration of call i's main kernel; *count = calls recorded, which may exceed capacity).  The events sit between the
 * period-block chain and the main kernel, so dependent-launch overlap across that boundary is lost while enabled. */
int         wn_timing_main_kernel_enable(wn_ctx *ctx, int on);
int         wn_timing_main_kernel_collect(wn_ctx *ctx, float *ms, int capacity, int *count);
/* Host-only (no GPU needed): which bands a WN_EVAL_FAST wn_multiband3d_lattice call on these axes would evaluate once
 * per period ("fold") at its top level for a tile of edge tile_n.  band_folded[nbands] receives 0/1 in the caller's
 * band order, block[3] the period block Lx, Ly, Lz in samples (1,1,1 when nothing folds); *nfolded the count.  The
 * decision never changes a result (canonical summation), only where the work is done. */
int         wn_debug_fold_plan(const float *xs, int nx, const float *ys, int ny, const float *zs, int nz,
                               const float *band_scale, int nbands, int tile_n,
                               int *band_folded, int block[3], int *nfolded);

/* Diagnostics (GPU): the axis-table entries the FAST lattice kernels use for `count` coordinates of one axis at one
 * band scale, computed by the same kernel as in a real call: weights3[3i..3i+2] = the three quadratic B-spline weights
 * (WaveletNoise.cpp:194-200, un-fused), first_cell[i] = mid - 1 (the first of the three tap cells, NOT wrapped; the taps
 * are Mod(first_cell + f, n), f = 0..2 -- the reference's indices cpp:202-209).  Host pointers. */
int         wn_debug_axis_entries(wn_ctx *ctx, const float *coords, int count, float band_scale,
                                  float *weights3, int32_t *first_cell);

/* ---- context ------------------------------------------------------------------------------ */
int  wn_ctx_create(int device /* -1 = current */, wn_ctx **out);
int  wn_ctx_destroy(wn_ctx *ctx);
int  wn_ctx_set_stream(wn_ctx *ctx, void *cuda_stream /* cudaStream_t, NULL = library's own */);
int  wn_ctx_synchronize(wn_ctx *ctx);
int  wn_ctx_device(const wn_ctx *ctx, int *device, int *sm_count);
/* pinned host buffers for callers that want full-speed copies */
int  wn_host_alloc(size_t bytes, void **out);
int  wn_host_free(void *p);

/* ---- host RNG: the reference's own generator objects ---------------------------------------
 * replaces: the `std::mt19937 rng; std::normal_distribution<float> gaussianDist` members
 * (WaveletNoise.h:47-48) and the fill loops WaveletNoise.cpp:74-77 / :146-147.  The state
 * persists across calls so a second generate* continues the stream like the reference does. */
int  wn_rng_create(unsigned seed, wn_rng **out);
int  wn_rng_destroy(wn_rng *rng);
int  wn_rng_fill_gaussian(wn_rng *rng, float *out_host, size_t count);
/* advance the engine by `raw_draws` 32-bit outputs and clear the distribution's cached variate: brings the host
 * objects to the state they would have after a fill that wn_tile_build_seeded performed on the GPU */
int  wn_rng_discard(wn_rng *rng, unsigned long long raw_draws);
/* std::shuffle(iota(256), std::mt19937(seed)) duplicated to 512 ints: the constructors
 * PerlinNoise.hpp:29-34 and perlin.h:34-39 */
int  wn_perlin_make_perm(unsigned seed, int32_t perm512_host[512]);

/* ---- tile construction ----------------------------------------------------------------------
 * replaces: WaveletNoise::WaveletNoise (cpp:20-26), generateNoiseTile2D (cpp:69-108),
 * generateNoiseTile3D (cpp:142-183), getNoiseCoefficients / getTileSize (cpp:290-291). */
int  wn_adjust_tile_size(int n);                       /* odd n -> n+1, the ctor rule cpp:22-25 */
int  wn_tile_create(wn_ctx *ctx, int n, int dims /* 2|3 */, unsigned flags, wn_tile **out);
int  wn_tile_destroy(wn_tile *tile);
int  wn_tile_info(const wn_tile *tile, int *n, int *dims, size_t *count, int *built);
/* R = the Gaussian field (n^dims floats, memory order).  Runs the separable down/up passes and the
 * subtraction on the GPU.  Arithmetic is un-fused and in the reference's summation order, so the
 * tile is bit-identical to the CPU one for the same R. */
int  wn_tile_build_from_gaussian(wn_tile *tile, const float *R, int space);
/* device-side fill: MT19937 + libstdc++'s polar method + a restatement of glibc's logf, all on the GPU; same
 * accept/reject sequence and the same bits as `std::normal_distribution<float>` over a fresh
 * `std::mt19937(seed)` (i.e. a freshly constructed WaveletNoise), so the tile is bit-identical.
 * *mt_draws (nullable) receives the number of raw engine outputs the fill consumed, so a host generator can be
 * brought to the same state with wn_rng_discard / std::mt19937::discard. */
int  wn_tile_build_seeded(wn_tile *tile, unsigned seed, unsigned long long *mt_draws);
/* adopt finished coefficients (e.g. a copy of a WaveletNoise object, or a cached tile) */
int  wn_tile_upload(wn_tile *tile, const float *N, int space);
int  wn_tile_download(const wn_tile *tile, float *out, int space);
int  wn_tile_device_ptr(const wn_tile *tile, void **dptr);   /* for NCCL broadcast of the tile */
int  wn_tile_mark_built(wn_tile *tile);                      /* after writing through the ptr  */

/* ---- evaluation: scattered points -----------------------------------------------------------
 * out[i] = evaluate*(p_i * pre_scale) * post_scale.
 * replaces: WaveletNoise::evaluate2D (cpp:111-140), evaluate3D (cpp:185-215),
 * evaluate3DProjected (cpp:218-265) called in a loop; the scalar methods are count = 1.
 * p is AoS (xy / xyz per point).  Reference operation order, bit-identical. */
int  wn_eval2d_points(const wn_tile *tile, const float *p, size_t count,
                      float pre_scale, float post_scale, float *out, int space);
int  wn_eval3d_points(const wn_tile *tile, const float *p, size_t count,
                      float pre_scale, float post_scale, float *out, int space);
/* normals: 3 floats shared by all points (normal_is_shared=1, HOST pointer) or one xyz per point
 * (normal_is_shared=0, same space as p).  Assumed unit length, like the reference. */
int  wn_eval3d_projected_points(const wn_tile *tile, const float *p, const float *normals,
                                int normal_is_shared, size_t count,
                                float pre_scale, float post_scale, float *out, int space);
/* out[i] = post_scale * sum_b weights[b] * evaluate3D(p_i * band_scale[b])   (paper App.2
 * WMultibandNoise over the reference's evaluate3D; the reference itself only ever uses nbands=1:
 * experient/main.cpp:45-58, texture.h:77-85). */
int  wn_multiband3d_points(const wn_tile *tile, const float *p, size_t count,
                           const float *band_scale, const float *weights, int nbands,
                           float post_scale, float *out, int space);

/* The paper's own multiband entry point (Cook & DeRose 2005, Appendix 2), which the reference omits (it only ever
 * evaluates one band):  WMultibandNoise(p, s, normal, firstBand, nbands, w) =
 *   sum over b < nbands while s + firstBand + b < 0 of w[b] * (normal ? WProjectedNoise(q, normal) : WNoise(q)),
 *   q = 2 p 2^(firstBand+b), divided by sqrt(sum_b w[b]^2 * (normal ? 0.296 : 0.210)) over ALL nbands weights.
 * s is the scale cut-off (log2 of the sample footprint: bands finer than the footprint are dropped), normal is NULL or
 * three floats shared by the batch (HOST pointer).  WNoise / WProjectedNoise are the reference's evaluate3D /
 * evaluate3DProjected (WaveletNoise.cpp:185-265).  Note the paper's variance constant 0.210 where the reference's
 * drivers use 0.18402 (experient/main.cpp:43). */
int  wn_wmultiband_points(const wn_tile *tile, const float *p, size_t count, float s, const float *normal,
                          int first_band, int nbands, const float *w, float *out, int space);

/* ---- evaluation: lattices (axis-aligned grids given by coordinate arrays) -------------------
 * sample (i,j,k) = (xs[i], ys[j], zs[k]); the caller computes the axes with whatever float formula
 * it uses (the reference: u = (float(x)/size)*4.0f, experient/main.cpp:20-21), so coordinates are
 * bit-identical to the CPU loop.  replaces the per-pixel loops experient/main.cpp:18-30, :45-58. */
int  wn_eval2d_lattice(const wn_tile *tile, const float *xs, int nx, const float *ys, int ny,
                       float pre_scale, float post_scale, float *out, int space);
/* wn_multiband3d_lattice: out[i + nx*(j + ny*k)] = post_scale * sum_b weights[b] * evaluate3D(p_ijk * band_scale[b])
 * (the composition of wn_multiband3d_points on a lattice; the axes are always HOST arrays).
 * mode WN_EVAL_FAST: separable evaluation; bands whose samples repeat with the tile period on this lattice are
 *   evaluated once per period ("folded") in stream-ordered scratch memory (up to 512 MiB per nesting level, released
 *   after the call; device-output calls of <= 2^29 samples keep two generations of it between calls).  Results are
 *   within 1e-5 * (tile max - tile min) per unit of band weight of the reference, and a sample's value depends only on
 *   its coordinates, the bands and the tile: not on folding, chunking, the band order given, or how a volume is cut
 *   into calls (slabs / shards are bit-identical to the single call).
 * mode WN_EVAL_EXACT: reference operation order, bit-identical to the CPU loop.
 * space WN_DEVICE: the call only enqueues.  `out` is ready in the order of the context's compute stream
 *   (wn_ctx_set_stream); the library may run the part of the work that does not depend on earlier work of that stream
 *   on an internal stream, which the compute stream then waits for. */
int  wn_multiband3d_lattice(const wn_tile *tile,
                            const float *xs, int nx, const float *ys, int ny, const float *zs, int nz,
                            const float *band_scale, const float *weights, int nbands,
                            float post_scale, int mode, float *out, int space);

/* ---- evaluation: affine grids ----------------------------------------------------------------
 * sample (i,j) = origin + us[i]*e1 + vs[j]*e2, each component evaluated as
 * fadd(fadd(origin, fmul(us[i],e1)), fmul(vs[j],e2)) -- un-fused, fixed order -- then * pre_scale.
 * replaces the loop experient/main.cpp:74-87 (and generalises it to oblique planes). */
int  wn_eval3d_projected_grid(const wn_tile *tile, const float origin[3],
                              const float e1[3], const float *us, int nu,
                              const float e2[3], const float *vs, int nv,
                              const float normal[3], float pre_scale, float post_scale,
                              float *out, int space);
int  wn_eval3d_grid(const wn_tile *tile, const float origin[3],
                    const float e1[3], const float *us, int nu,
                    const float e2[3], const float *vs, int nv,
                    float pre_scale, float post_scale, float *out, int space);

/* ---- Perlin reference noise -------------------------------------------------------------------
 * replaces: PerlinNoise::noise / perlin::noise (PerlinNoise.hpp:36-60, perlin.h:42-72), double
 * precision, un-fused, result narrowed to float on store (experient/main.cpp:104,122). */
int  wn_perlin_create(wn_ctx *ctx, const int32_t perm512_host[512], wn_perlin **out);
int  wn_perlin_destroy(wn_perlin *pn);
/* Arithmetic of the float batch calls below (points / lattice / grid).  WN_PERLIN_F64 (default): the reference's
 * double-precision, un-fused operation order -- bit-identical to PerlinNoise::noise narrowed to float.
 * WN_PERLIN_F32: single precision with FMAs, within 1e-5 * 2 (the noise range) of the FP64 result for float-valued
 * coordinates; opt-in fast mode (hard part 6).  Texture hooks always run FP64. */
#define WN_PERLIN_F64    0
#define WN_PERLIN_F32    1
int  wn_perlin_set_precision(wn_perlin *pn, int precision);
/* double coordinates in (xyz per point), double noise out: the scalar signature of PerlinNoise::noise(double x,
 * double y, double z) (experient/PerlinNoise.hpp:36-56, perlin.h:42-62) without narrowing at either end. */
int  wn_perlin_points_f64(const wn_perlin *pn, const double *p, size_t count, double *out, int space);
int  wn_perlin_points(const wn_perlin *pn, const float *p, size_t count, float pre_scale,
                      float *out, int space);
int  wn_perlin_lattice(const wn_perlin *pn, const float *xs, int nx, const float *ys, int ny,
                       const float *zs, int nz, float *out, int space);
int  wn_perlin_grid(const wn_perlin *pn, const float origin[3],
                    const float e1[3], const float *us, int nu,
                    const float e2[3], const float *vs, int nv,
                    float pre_scale, float *out, int space);

/* ---- texture hooks (batched) ------------------------------------------------------------------
 * grey[i] = the value wavelet_texture::value / noise_texture::value (texture.h:67-107, :37-43)
 * returns in each colour channel for hit point p_i (xyz float, as stored in vec3). */
int  wn_wavelet_texture_values(const wn_tile *tile3d, const float *p, size_t count,
                               double scale, int octave, float *grey, int space);
/* the 2D branch of wavelet_texture::value (texture.h:86-99; use_3d = false): xy of p on a 2D tile */
int  wn_wavelet_texture2d_values(const wn_tile *tile2d, const float *p, size_t count,
                                 double scale, int octave, float *grey, int space);
int  wn_perlin_texture_values(const wn_perlin *pn, const float *p, size_t count,
                              double scale, int octave, float *grey, int space);

/* ---- statistics --------------------------------------------------------------------------------
 * replaces: WaveletNoise::calculateStats (cpp:268-288) without the printing (the host layer prints). */
int  wn_stats_compute(wn_ctx *ctx, const float *data, size_t count, int space, wn_stats *out);

/* ---- device groups: single-process multi-GPU -------------------------------------------------------
 * north_star: volumes shard by z-slab and images by row-band across the GPUs of one box, each GPU holds a replica of
 * the tile (broadcast once over NVLink with NCCL), output is gathered only for file write.  A wn_group owns one wn_ctx
 * per GPU, driven by one host thread; sharded calls enqueue on every GPU before waiting for any, and there is no
 * data-path collective (every sample reads only its GPU's tile replica).  The reference is single-threaded and has no
 * counterpart; the loops that are sharded are experient/main.cpp:45-58 (volume, along z), :74-87 (projected plane, by
 * row) and main.cpp:175-204 (render rows, through the batched texture hook).  NCCL is loaded at run time (libnccl.so.2)
 * and only for groups of two or more GPUs. */
typedef struct wn_group wn_group;
typedef struct wn_gtile wn_gtile;
#define WN_SHARD_SLAB    0   /* contiguous z-slabs / row-bands                                             */
#define WN_SHARD_CYCLIC  1   /* volumes: 32-slice chunks dealt round-robin (every rank keeps the periodic
                                structure of the whole volume, so per-rank work stays 1/N) */
/* Host-only (no GPU needed): the slices / rows rank `rank` of `world` owns under `sharding`, ascending; *count = how
 * many (may exceed capacity). */
int  wn_debug_shard_indices(int total, int rank, int world, int sharding, int *indices, int capacity, int *count);
int  wn_group_create(int ngpus /* <= 0: all visible */, const int *devices /* NULL: 0..ngpus-1 */, wn_group **out);
int  wn_group_destroy(wn_group *g);
int  wn_group_size(const wn_group *g);
int  wn_group_ctx(wn_group *g, int rank, wn_ctx **ctx);
int  wn_group_synchronize(wn_group *g);
/* one tile replica per GPU: rank 0 builds (wn_tile_build_seeded) or receives the coefficients, one ncclBroadcast
 * replicates them on the ranks' streams */
int  wn_group_tile_create(wn_group *g, int n, int dims, unsigned flags, wn_gtile **out);
int  wn_group_tile_destroy(wn_gtile *t);
int  wn_group_tile_build_seeded(wn_gtile *t, unsigned seed, unsigned long long *mt_draws);
int  wn_group_tile_upload(wn_gtile *t, const float *N_host);
int  wn_group_tile_rank(wn_gtile *t, int rank, wn_tile **tile);
/* wn_multiband3d_lattice sharded along z (BASELINE config 3).  The shards stay on the devices (wn_group_shard); when
 * out_host is not NULL the volume is also gathered there in the lattice layout (pinned memory recommended).  *gpu_ms
 * (nullable) = max over the GPUs of the CUDA-event time of the enqueued kernels.  With out_host == NULL and gpu_ms ==
 * NULL the call only enqueues (wn_group_synchronize waits), so back-to-back calls keep every GPU busy.  Results are
 * bit-identical to the single-GPU call for either sharding (canonical summation, see wn_multiband3d_lattice). */
int  wn_group_multiband3d_lattice(wn_gtile *t, const float *xs, int nx, const float *ys, int ny, const float *zs, int nz,
                                  const float *band_scale, const float *weights, int nbands, float post_scale, int mode,
                                  int sharding, float *out_host, float *gpu_ms);
/* device buffer and float count of `rank`'s shard of the most recent sharded call */
int  wn_group_shard(wn_group *g, int rank, float **dptr, size_t *count);
/* wn_eval3d_projected_grid / wn_perlin_grid sharded into contiguous row-bands of the v axis (BASELINE config 4) */
int  wn_group_eval3d_projected_grid(wn_gtile *t, const float origin[3], const float e1[3], const float *us, int nu,
                                    const float e2[3], const float *vs, int nv, const float normal[3], float pre_scale,
                                    float post_scale, float *out_host, float *gpu_ms);
int  wn_group_perlin_grid(wn_group *g, wn_perlin *const *perlin_per_rank, const float origin[3], const float e1[3],
                          const float *us, int nu, const float e2[3], const float *vs, int nv, float pre_scale,
                          float *out_host, float *gpu_ms);
/* wn_wavelet_texture_values with the points cut into one contiguous run per GPU (BASELINE config 5: the hit points of a
 * row band).  Copies in, kernels and copies out of all GPUs are in flight together; wait == 0 returns after enqueueing
 * (wn_group_synchronize before reading grey_host), so the caller can trace the next band meanwhile. */
int  wn_group_wavelet_texture_values(wn_gtile *t, const float *p_host, size_t count, double scale, int octave,
                                     float *grey_host, int wait);

int  wn_group_perlin_texture_values(wn_group *g, wn_perlin *const *perlin_per_rank, const float *p_host, size_t count,
                                    double scale, int octave, float *grey_host, int wait);

#ifdef __cplusplus
}
#endif
#endif /* WN_B200_H */
