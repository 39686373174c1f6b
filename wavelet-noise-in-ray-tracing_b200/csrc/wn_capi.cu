// wn_capi.cu -- the extern "C" surface of include/wn_b200.h: contexts, tiles, staging and launches.
// No kernels live here (see wn_tilegen.cu, wn_eval_exact.cu, wn_multiband_fast.cu, wn_rng.cu).
//
// There is deliberately no CPU compute path in this file: every evaluation entry point ends in a kernel
// launch, and wn_ctx_create fails when no sm_100 device is present.  The only host-side arithmetic is the
// reference's own generator objects (std::mt19937 / std::normal_distribution<float> / std::shuffle), which
// is exactly what the reference's constructors and fill loops run (WaveletNoise.h:47-48, cpp:74-77,146-147;
// PerlinNoise.hpp:29-34).
#include "../../include/wn_b200.h"
#include "wn_internal.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <numeric>
#include <random>
#include <vector>

// ---------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int wn_fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define WN_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return wn_fail(WN_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#define WN_REQUIRE(cond, ...)                                                                      \
    do {                                                                                           \
        if (!(cond)) return wn_fail(WN_EINVAL, __VA_ARGS__);                                       \
    } while (0)

// used by the other host translation units (wn_group.cu)
int wn_set_error(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

extern "C" const char *wn_last_error(void) { return g_err; }
extern "C" const char *wn_version(void) { return "wn_b200 0.1 (sm_100a)"; }

// ---------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------
struct WnBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct wn_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr;   // compute stream created by the library
    cudaStream_t stream = nullptr;       // stream in use (own or caller's)
    cudaStream_t h2d = nullptr, d2h = nullptr;
    // High-priority stream for the period-block chain of a fast lattice call (axis tables + nested period blocks).
    // That chain does not depend on earlier work of the compute stream, so it runs while the PREVIOUS call's main
    // kernel still occupies the GPU; the compute stream waits for ev_side before the main kernel.
    cudaStream_t side = nullptr;
    cudaEvent_t ev_side = nullptr, ev_tile = nullptr;
    WnBuf params_side;                   // the chain's own parameter block (uploaded on `side`)
    std::vector<char> h_params_side;
    // Scratch of the last two side-chain calls (axis tables, outer period block).  It is allocated AND freed on the
    // side stream so the pool recycles it without cross-stream waits; the free of generation g is issued at the start
    // of the next call that uses g, after the side stream has been ordered behind ev_main[g] (recorded on the compute
    // stream after the main kernel that read the scratch) -- by then that kernel is two calls in the past.
    struct Deferred { void *tab = nullptr, *P = nullptr; bool pending = false; } deferred[2];
    cudaEvent_t ev_main[2] = { nullptr, nullptr };
    uint64_t side_calls = 0;
    // double-buffered staging for WN_HOST calls
    WnBuf in[2], aux[2], outb[2];
    cudaEvent_t ev_in[2] = { nullptr, nullptr }, ev_k[2] = { nullptr, nullptr }, ev_out[2] = { nullptr, nullptr };
    WnBuf params;                        // coordinate axes etc. for the call in flight
    std::vector<char> h_params;          // host staging of the same block
    WnBuf stats_partial;
    std::vector<cudaEvent_t> tev;        // timing event pool (pairs)
    size_t tev_used = 0;
    float last_ms = 0.0f;
    uint64_t launches = 0;
    // optional per-call timing of the MAIN kernel of WN_DEVICE fast lattice calls (wn_timing_main_kernel_enable):
    // one event pair per call, read back by wn_timing_main_kernel_collect after a synchronisation
    bool time_main = false;
    std::vector<cudaEvent_t> main_ev;    // pairs, in call order
};

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

static int buf_reserve(WnBuf &b, size_t bytes)
{
    if (bytes <= b.cap) return WN_OK;
    if (b.p) WN_CUDA(cudaFree(b.p));
    b.p = nullptr; b.cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return wn_fail(WN_ENOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    }
    b.cap = want;
    return WN_OK;
}

// library-private scratch pool of a device (see wn_internal.h); created on first use, lives until process exit
static cudaMemPool_t scratch_pool(int device)
{
    static std::mutex mu;
    static cudaMemPool_t pools[64] = {};
    if (device < 0 || device >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (!pools[device]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        cudaMemPool_t pool = nullptr;
        if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        unsigned long long keep = ~0ull;             // recycle, do not return scratch to the OS between calls
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        pools[device] = pool;
    }
    return pools[device];
}

cudaError_t wn_scratch_alloc(void **p, size_t bytes, cudaStream_t st)
{
    int device = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess) return e;
    cudaMemPool_t pool = scratch_pool(device);
    if (!pool) return cudaErrorMemoryAllocation;
    return cudaMallocFromPoolAsync(p, bytes, pool, st);
}

static void ctx_release(wn_ctx *c);

extern "C" int wn_ctx_create(int device, wn_ctx **out)
{
    WN_REQUIRE(out, "wn_ctx_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return wn_fail(WN_ENODEVICE, "no CUDA device available (%s); this library has no CPU fallback",
                       e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device < 0) WN_CUDA(cudaGetDevice(&device));
    WN_REQUIRE(device < count, "wn_ctx_create: device %d out of range (%d devices)", device, count);
    cudaDeviceProp prop;
    WN_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return wn_fail(WN_ENODEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                       prop.major, prop.minor);
    DeviceGuard g(device);
    wn_ctx *c = new (std::nothrow) wn_ctx();
    if (!c) return wn_fail(WN_ENOMEM, "out of host memory");
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    // a failure below releases what was created so far (streams, events, the context itself)
#define WN_CUDA_CTX(call)                                                                          \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            ctx_release(c);                                                                        \
            return wn_fail(WN_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
        }                                                                                          \
    } while (0)
    WN_CUDA_CTX(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    WN_CUDA_CTX(cudaStreamCreateWithFlags(&c->h2d, cudaStreamNonBlocking));
    WN_CUDA_CTX(cudaStreamCreateWithFlags(&c->d2h, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    {
        int least = 0, greatest = 0;
        WN_CUDA_CTX(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        WN_CUDA_CTX(cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, greatest));
        WN_CUDA_CTX(cudaEventCreateWithFlags(&c->ev_side, cudaEventDisableTiming));
        WN_CUDA_CTX(cudaEventCreateWithFlags(&c->ev_tile, cudaEventDisableTiming));
        for (int i = 0; i < 2; ++i) WN_CUDA_CTX(cudaEventCreateWithFlags(&c->ev_main[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < 2; ++i) {
        WN_CUDA_CTX(cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming));
        WN_CUDA_CTX(cudaEventCreateWithFlags(&c->ev_k[i], cudaEventDisableTiming));
        WN_CUDA_CTX(cudaEventCreateWithFlags(&c->ev_out[i], cudaEventDisableTiming));
    }
#undef WN_CUDA_CTX
    if (!scratch_pool(device)) {
        ctx_release(c);
        return wn_fail(WN_ECUDA, "could not create the library's scratch memory pool on device %d", device);
    }
    *out = c;
    return WN_OK;
}

// everything a context owns; members that were never created are null (partial construction)
static void ctx_release(wn_ctx *c)
{
    for (int i = 0; i < 2; ++i) {
        cudaFree(c->in[i].p); cudaFree(c->aux[i].p); cudaFree(c->outb[i].p);
        if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]);
        if (c->ev_k[i]) cudaEventDestroy(c->ev_k[i]);
        if (c->ev_out[i]) cudaEventDestroy(c->ev_out[i]);
    }
    for (int i = 0; i < 2; ++i) {
        if (c->deferred[i].pending && c->side) { cudaFreeAsync(c->deferred[i].tab, c->side); cudaFreeAsync(c->deferred[i].P, c->side); }
        if (c->ev_main[i]) cudaEventDestroy(c->ev_main[i]);
    }
    if (c->side) cudaStreamSynchronize(c->side);
    cudaFree(c->params.p);
    cudaFree(c->params_side.p);
    cudaFree(c->stats_partial.p);
    if (c->ev_side) cudaEventDestroy(c->ev_side);
    if (c->ev_tile) cudaEventDestroy(c->ev_tile);
    if (c->side) cudaStreamDestroy(c->side);
    for (cudaEvent_t e : c->tev) cudaEventDestroy(e);
    for (cudaEvent_t e : c->main_ev) cudaEventDestroy(e);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    if (c->h2d) cudaStreamDestroy(c->h2d);
    if (c->d2h) cudaStreamDestroy(c->d2h);
    cudaGetLastError();
    delete c;
}

extern "C" int wn_ctx_destroy(wn_ctx *c)
{
    if (!c) return WN_OK;
    DeviceGuard g(c->device);
    cudaDeviceSynchronize();
    ctx_release(c);
    return WN_OK;
}

cudaStream_t wn_ctx_stream_internal(const wn_ctx *c) { return c->stream; }

extern "C" int wn_ctx_set_stream(wn_ctx *c, void *s)
{
    WN_REQUIRE(c, "wn_ctx_set_stream: ctx is NULL");
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return WN_OK;
}

extern "C" int wn_ctx_synchronize(wn_ctx *c)
{
    WN_REQUIRE(c, "wn_ctx_synchronize: ctx is NULL");
    DeviceGuard g(c->device);
    WN_CUDA(cudaStreamSynchronize(c->stream));
    return WN_OK;
}

extern "C" int wn_ctx_device(const wn_ctx *c, int *device, int *sm_count)
{
    WN_REQUIRE(c, "wn_ctx_device: ctx is NULL");
    if (device) *device = c->device;
    if (sm_count) *sm_count = c->sm_count;
    return WN_OK;
}

extern "C" uint64_t wn_kernel_launches(const wn_ctx *c) { return c ? c->launches : 0; }
extern "C" float wn_timing_last_ms(const wn_ctx *c) { return c ? c->last_ms : 0.0f; }

extern "C" int wn_timing_main_kernel_enable(wn_ctx *c, int on)
{
    WN_REQUIRE(c, "wn_timing_main_kernel_enable: ctx is NULL");
    c->time_main = on != 0;
    return WN_OK;
}

extern "C" int wn_timing_main_kernel_collect(wn_ctx *c, float *ms, int capacity, int *count)
{
    WN_REQUIRE(c && count, "wn_timing_main_kernel_collect: NULL argument");
    DeviceGuard g(c->device);
    WN_CUDA(cudaStreamSynchronize(c->stream));
    const int pairs = (int)(c->main_ev.size() / 2);
    *count = pairs;
    for (int i = 0; i < pairs; ++i) {
        float t = 0.0f;
        WN_CUDA(cudaEventElapsedTime(&t, c->main_ev[2 * i], c->main_ev[2 * i + 1]));
        if (ms && i < capacity) ms[i] = t;
    }
    for (cudaEvent_t e : c->main_ev) cudaEventDestroy(e);
    c->main_ev.clear();
    return WN_OK;
}

extern "C" int wn_host_alloc(size_t bytes, void **out)
{
    WN_REQUIRE(out, "wn_host_alloc: out is NULL");
    WN_CUDA(cudaMallocHost(out, bytes ? bytes : 1));
    return WN_OK;
}
extern "C" int wn_host_free(void *p)
{
    if (p) WN_CUDA(cudaFreeHost(p));
    return WN_OK;
}

// timing helpers: pairs of events around kernel groups of a WN_HOST call
static int timing_begin(wn_ctx *c) { c->tev_used = 0; c->last_ms = 0.0f; return WN_OK; }
static int timing_mark(wn_ctx *c, cudaStream_t st)
{
    if (c->tev_used == c->tev.size()) {
        cudaEvent_t e;
        WN_CUDA(cudaEventCreate(&e));
        c->tev.push_back(e);
    }
    WN_CUDA(cudaEventRecord(c->tev[c->tev_used++], st));
    return WN_OK;
}
static int timing_end(wn_ctx *c)      // call after the streams have been synchronised
{
    float total = 0.0f;
    for (size_t i = 0; i + 1 < c->tev_used; i += 2) {
        float ms = 0.0f;
        WN_CUDA(cudaEventElapsedTime(&ms, c->tev[i], c->tev[i + 1]));
        total += ms;
    }
    c->last_ms = total;
    return WN_OK;
}

// Small host parameter arrays (coordinate axes, ...) of one call: packed into one host staging block and sent to the
// context's device parameter block with a SINGLE stream-ordered copy (flush), instead of one copy per array.
struct ParamWriter {
    WnBuf &dev;
    std::vector<char> &host;
    cudaStream_t stream;
    size_t off = 0;
    explicit ParamWriter(wn_ctx *ctx) : dev(ctx->params), host(ctx->h_params), stream(ctx->stream) {}
    // the side-stream chain of a fast lattice call has its own block, ordered on that stream
    ParamWriter(wn_ctx *ctx, bool side) : dev(side ? ctx->params_side : ctx->params),
                                          host(side ? ctx->h_params_side : ctx->h_params),
                                          stream(side ? ctx->side : ctx->stream) {}
    int reserve(size_t bytes)
    {
        if (bytes + 1024 > dev.cap) {
            // growing frees the old block: nothing may still read it
            WN_CUDA(cudaStreamSynchronize(stream));
        }
        int r = buf_reserve(dev, bytes + 1024);
        if (r) return r;
        if (host.size() < bytes + 1024) host.resize(bytes + 1024);
        return WN_OK;
    }
    int put(const void *src, size_t bytes, const void **dptr)
    {
        off = (off + 255) & ~(size_t)255;
        if (off + bytes > dev.cap || off + bytes > host.size())
            return wn_fail(WN_EINVAL, "internal: parameter block overflow");
        std::memcpy(host.data() + off, src, bytes);
        *dptr = (char *)dev.p + off;
        off += bytes;
        return WN_OK;
    }
    int flush()
    {
        if (!off) return WN_OK;
        // pageable source: the runtime stages it before returning, so the staging block can be reused by the next call
        cudaError_t e = cudaMemcpyAsync(dev.p, host.data(), off, cudaMemcpyHostToDevice, stream);
        if (e != cudaSuccess) return wn_fail(WN_ECUDA, "parameter upload failed: %s", cudaGetErrorString(e));
        return WN_OK;
    }
};

// ---------------------------------------------------------------------------------------------------
// chunked HOST pipeline:  H2D(in, aux) -> kernel -> D2H(out), double buffered on three streams
// ---------------------------------------------------------------------------------------------------
struct ChunkIO {
    const void *in = nullptr;  size_t in_item = 0;     // bytes per item (0 = no per-item input)
    const void *aux = nullptr; size_t aux_item = 0;
    void *out = nullptr;       size_t out_item = sizeof(float);
};

template <class Launch>
static int run_chunked_host(wn_ctx *c, size_t total, size_t chunk, const ChunkIO &io, Launch launch)
{
    if (total == 0) return WN_OK;
    if (chunk == 0 || chunk > total) chunk = total;
    for (int s = 0; s < 2; ++s) {
        if (io.in_item) { int r = buf_reserve(c->in[s], chunk * io.in_item); if (r) return r; }
        if (io.aux_item) { int r = buf_reserve(c->aux[s], chunk * io.aux_item); if (r) return r; }
        int r = buf_reserve(c->outb[s], chunk * io.out_item); if (r) return r;
    }
    // the parameter block was written on c->stream; the h2d stream needs no ordering with it.
    timing_begin(c);
    size_t nchunks = (total + chunk - 1) / chunk;
    if (nchunks == 1) {
        // small calls (the scalar evaluate*() methods are count == 1): everything in order on the compute stream
        if (io.in_item)
            WN_CUDA(cudaMemcpyAsync(c->in[0].p, io.in, total * io.in_item, cudaMemcpyHostToDevice, c->stream));
        if (io.aux_item)
            WN_CUDA(cudaMemcpyAsync(c->aux[0].p, io.aux, total * io.aux_item, cudaMemcpyHostToDevice, c->stream));
        int r = timing_mark(c, c->stream); if (r) return r;
        int nl = launch(c->in[0].p, c->aux[0].p, (float *)c->outb[0].p, 0, total, c->stream);
        if (nl < 0) return wn_fail(WN_EINVAL, "kernel launch rejected the configuration");
        c->launches += (uint64_t)nl;
        WN_CUDA(cudaGetLastError());
        r = timing_mark(c, c->stream); if (r) return r;
        WN_CUDA(cudaMemcpyAsync(io.out, c->outb[0].p, total * io.out_item, cudaMemcpyDeviceToHost, c->stream));
        WN_CUDA(cudaStreamSynchronize(c->stream));
        return timing_end(c);
    }
    for (size_t ci = 0; ci < nchunks; ++ci) {
        const int s = (int)(ci & 1);
        const size_t first = ci * chunk, cnt = std::min(chunk, total - first);
        const bool has_in = io.in_item || io.aux_item;
        if (has_in) {
            if (ci >= 2) WN_CUDA(cudaStreamWaitEvent(c->h2d, c->ev_k[s], 0));   // kernel ci-2 done with in[s]
            if (io.in_item)
                WN_CUDA(cudaMemcpyAsync(c->in[s].p, (const char *)io.in + first * io.in_item, cnt * io.in_item,
                                        cudaMemcpyHostToDevice, c->h2d));
            if (io.aux_item)
                WN_CUDA(cudaMemcpyAsync(c->aux[s].p, (const char *)io.aux + first * io.aux_item, cnt * io.aux_item,
                                        cudaMemcpyHostToDevice, c->h2d));
            WN_CUDA(cudaEventRecord(c->ev_in[s], c->h2d));
            WN_CUDA(cudaStreamWaitEvent(c->stream, c->ev_in[s], 0));
        }
        if (ci >= 2) WN_CUDA(cudaStreamWaitEvent(c->stream, c->ev_out[s], 0));    // D2H ci-2 done with outb[s]
        int r = timing_mark(c, c->stream); if (r) return r;
        int nl = launch(c->in[s].p, c->aux[s].p, (float *)c->outb[s].p, first, cnt, c->stream);
        if (nl < 0) return wn_fail(WN_EINVAL, "kernel launch rejected the configuration");
        c->launches += (uint64_t)nl;
        WN_CUDA(cudaGetLastError());
        r = timing_mark(c, c->stream); if (r) return r;
        WN_CUDA(cudaEventRecord(c->ev_k[s], c->stream));
        WN_CUDA(cudaStreamWaitEvent(c->d2h, c->ev_k[s], 0));
        WN_CUDA(cudaMemcpyAsync((char *)io.out + first * io.out_item, c->outb[s].p, cnt * io.out_item,
                                cudaMemcpyDeviceToHost, c->d2h));
        WN_CUDA(cudaEventRecord(c->ev_out[s], c->d2h));
    }
    WN_CUDA(cudaStreamSynchronize(c->d2h));
    WN_CUDA(cudaStreamSynchronize(c->stream));
    return timing_end(c);
}

template <class Launch>
static int run_device(wn_ctx *c, Launch launch)
{
    int nl = launch(c->stream);
    if (nl < 0) return wn_fail(WN_EINVAL, "kernel launch rejected the configuration");
    c->launches += (uint64_t)nl;
    WN_CUDA(cudaGetLastError());
    return WN_OK;
}

static const size_t kChunkSamples = (size_t)32 << 20;     // 32 Mi samples = 128 MiB of output per chunk

// ---------------------------------------------------------------------------------------------------
// host RNG (the reference's own generator objects)
// ---------------------------------------------------------------------------------------------------
struct wn_rng {
    std::mt19937 engine;
    std::normal_distribution<float> gauss;
    explicit wn_rng(unsigned seed) : engine(seed), gauss(0.0f, 1.0f) {}
};

extern "C" int wn_rng_create(unsigned seed, wn_rng **out)
{
    WN_REQUIRE(out, "wn_rng_create: out is NULL");
    *out = new (std::nothrow) wn_rng(seed);
    return *out ? WN_OK : wn_fail(WN_ENOMEM, "out of host memory");
}
extern "C" int wn_rng_destroy(wn_rng *r) { delete r; return WN_OK; }
extern "C" int wn_rng_discard(wn_rng *r, unsigned long long raw_draws)
{
    WN_REQUIRE(r, "wn_rng_discard: rng is NULL");
    r->engine.discard(raw_draws);
    r->gauss.reset();
    return WN_OK;
}
extern "C" int wn_rng_fill_gaussian(wn_rng *r, float *out, size_t count)
{
    WN_REQUIRE(r && (out || !count), "wn_rng_fill_gaussian: NULL argument");
    for (size_t i = 0; i < count; ++i) out[i] = r->gauss(r->engine);
    return WN_OK;
}
extern "C" int wn_perlin_make_perm(unsigned seed, int32_t perm[512])
{
    WN_REQUIRE(perm, "wn_perlin_make_perm: perm is NULL");
    std::vector<int> p(256);
    std::iota(p.begin(), p.end(), 0);
    std::shuffle(p.begin(), p.end(), std::mt19937(seed));
    for (int i = 0; i < 256; ++i) perm[i] = perm[256 + i] = p[i];
    return WN_OK;
}

// ---------------------------------------------------------------------------------------------------
// tiles
// ---------------------------------------------------------------------------------------------------
struct wn_tile {
    wn_ctx *ctx = nullptr;
    int n = 0, dims = 0;
    unsigned flags = 0;
    size_t count = 0;
    float *d = nullptr;
    float *dpad = nullptr;               // 3D: x-padded replica for the fast lattice kernel
    bool built = false;
    uint64_t version = 0;                // bumped whenever the coefficients change (on the compute stream)
    mutable uint64_t side_seen = ~0ull;  // version the side stream has been ordered after
    mutable void *plan_cache = nullptr;  // host-side plan of the last fast lattice call on this tile (axes tables, fold decisions)
};

static WnTileView tile_view(const wn_tile *t)
{
    WnTileView v;
    v.N = t->d; v.n = t->n; v.pow2 = (t->n & (t->n - 1)) == 0;
    v.Npad = t->dpad;
    return v;
}

extern "C" int wn_adjust_tile_size(int n) { return (n % 2 != 0) ? n + 1 : n; }

extern "C" int wn_tile_create(wn_ctx *c, int n, int dims, unsigned flags, wn_tile **out)
{
    WN_REQUIRE(c && out, "wn_tile_create: NULL argument");
    *out = nullptr;
    WN_REQUIRE(dims == 2 || dims == 3, "wn_tile_create: dims must be 2 or 3 (got %d)", dims);
    WN_REQUIRE(n >= 2, "wn_tile_create: tile size must be >= 2 (got %d)", n);
    n = wn_adjust_tile_size(n);
    WN_REQUIRE(dims == 3 ? n <= 1024 : n <= 16384, "wn_tile_create: tile size %d too large for %dD", n, dims);
    WN_REQUIRE(!(flags & WN_TILE_ODD_OFFSET) || dims == 3, "wn_tile_create: WN_TILE_ODD_OFFSET needs a 3D tile");
    DeviceGuard g(c->device);
    wn_tile *t = new (std::nothrow) wn_tile();
    if (!t) return wn_fail(WN_ENOMEM, "out of host memory");
    t->ctx = c; t->n = n; t->dims = dims; t->flags = flags;
    t->count = (dims == 3) ? (size_t)n * n * n : (size_t)n * n;
    cudaError_t e = cudaMalloc(&t->d, t->count * sizeof(float));
    if (e != cudaSuccess) {
        cudaGetLastError();
        delete t;
        return wn_fail(WN_ENOMEM, "cudaMalloc of the %d^%d tile failed: %s", n, dims, cudaGetErrorString(e));
    }
    if (dims == 3) {
        e = cudaMalloc(&t->dpad, (size_t)n * n * (n + WN_TILE_PAD) * sizeof(float));
        if (e != cudaSuccess) {
            cudaGetLastError();
            cudaFree(t->d);
            delete t;
            return wn_fail(WN_ENOMEM, "cudaMalloc of the padded tile failed: %s", cudaGetErrorString(e));
        }
    }
    *out = t;
    return WN_OK;
}

extern "C" int wn_tile_destroy(wn_tile *t)
{
    if (!t) return WN_OK;
    DeviceGuard g(t->ctx->device);
    cudaStreamSynchronize(t->ctx->stream);
    cudaStreamSynchronize(t->ctx->side);
    cudaFree(t->d);
    cudaFree(t->dpad);
    wn_mb3d_fast_cache_free(t->plan_cache);
    delete t;
    return WN_OK;
}

// the tile changed: refresh the padded replica (stream ordered) and mark it usable
static int tile_finish(wn_tile *t)
{
    if (t->dpad) {
        t->ctx->launches += (uint64_t)wn_launch_pad_tile(t->d, t->dpad, t->n, t->ctx->stream);
        WN_CUDA(cudaGetLastError());
    }
    t->built = true;
    ++t->version;
    return WN_OK;
}

extern "C" int wn_tile_info(const wn_tile *t, int *n, int *dims, size_t *count, int *built)
{
    WN_REQUIRE(t, "wn_tile_info: tile is NULL");
    if (n) *n = t->n;
    if (dims) *dims = t->dims;
    if (count) *count = t->count;
    if (built) *built = t->built ? 1 : 0;
    return WN_OK;
}

// R (device) -> tile.  Pass chaining follows WaveletNoise.cpp:153-182 (3D) and :87-107 (2D).
static int tile_build_device(wn_tile *t, const float *dR)
{
    wn_ctx *c = t->ctx;
    cudaStream_t st = c->stream;
    const size_t bytes = t->count * sizeof(float);
    float *t1 = nullptr, *t2 = nullptr;
    WN_CUDA(wn_scratch_alloc((void **)&t1, bytes, st));
    WN_CUDA(wn_scratch_alloc((void **)&t2, bytes, st));
    int nl = 0, r;
    const bool odd = (t->flags & WN_TILE_ODD_OFFSET) != 0;
    if (t->dims == 3) {
        if ((r = wn_launch_filter_axis(dR, t1, nullptr, t->n, 3, 0, st)) < 0) goto bad; nl += r;
        if ((r = wn_launch_filter_axis(t1, t2, nullptr, t->n, 3, 1, st)) < 0) goto bad; nl += r;
        // z pass writes N = R - (...) directly; with the odd-offset flag it goes through t1 -> t->d
        if (odd) {
            if ((r = wn_launch_filter_axis(t2, t1, dR, t->n, 3, 2, st)) < 0) goto bad; nl += r;
            nl += wn_launch_odd_offset3d(t1, t->d, t->n, st);
        } else {
            if ((r = wn_launch_filter_axis(t2, t->d, dR, t->n, 3, 2, st)) < 0) goto bad; nl += r;
        }
    } else {
        if ((r = wn_launch_filter_axis(dR, t1, nullptr, t->n, 2, 0, st)) < 0) goto bad; nl += r;
        if ((r = wn_launch_filter_axis(t1, t->d, dR, t->n, 2, 1, st)) < 0) goto bad; nl += r;
    }
    c->launches += (uint64_t)nl;
    WN_CUDA(cudaGetLastError());
    WN_CUDA(cudaFreeAsync(t1, st));
    WN_CUDA(cudaFreeAsync(t2, st));
    return tile_finish(t);
bad:
    cudaFreeAsync(t1, st); cudaFreeAsync(t2, st);
    return wn_fail(WN_EINVAL, "tile size %d does not fit the filter kernels' shared memory", t->n);
}

extern "C" int wn_tile_build_from_gaussian(wn_tile *t, const float *R, int space)
{
    WN_REQUIRE(t && R, "wn_tile_build_from_gaussian: NULL argument");
    wn_ctx *c = t->ctx;
    DeviceGuard g(c->device);
    if (space == WN_DEVICE) return tile_build_device(t, R);
    WN_REQUIRE(space == WN_HOST, "bad space %d", space);
    float *dR = nullptr;
    const size_t bytes = t->count * sizeof(float);
    WN_CUDA(wn_scratch_alloc((void **)&dR, bytes, c->stream));
    WN_CUDA(cudaMemcpyAsync(dR, R, bytes, cudaMemcpyHostToDevice, c->stream));
    timing_begin(c);
    timing_mark(c, c->stream);
    int r = tile_build_device(t, dR);
    timing_mark(c, c->stream);
    cudaFreeAsync(dR, c->stream);
    if (r) return r;
    WN_CUDA(cudaStreamSynchronize(c->stream));
    return timing_end(c);
}

extern "C" int wn_tile_build_seeded(wn_tile *t, unsigned seed, unsigned long long *mt_draws)
{
    WN_REQUIRE(t, "wn_tile_build_seeded: tile is NULL");
    wn_ctx *c = t->ctx;
    DeviceGuard g(c->device);
    cudaStream_t st = c->stream;
    float *dR = nullptr;
    unsigned long long *dacc = nullptr;
    WN_CUDA(wn_scratch_alloc((void **)&dR, t->count * sizeof(float), st));
    WN_CUDA(wn_scratch_alloc((void **)&dacc, 2 * sizeof(unsigned long long), st));
    int rc = WN_OK;
    for (int margin = 20; ; margin *= 4) {                 // 2 % more attempts than expected; never short in practice
        timing_begin(c);
        timing_mark(c, st);
        const int nl = wn_launch_gaussian_fill(seed, dR, t->count, dacc, margin, st);
        if (nl < 0) { rc = wn_fail(WN_ECUDA, "device Gaussian fill failed: %s", cudaGetErrorString(cudaGetLastError())); break; }
        c->launches += (uint64_t)nl;
        unsigned long long info[2] = { 0, 0 };
        if (cudaMemcpyAsync(info, dacc, sizeof(info), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess) {
            rc = wn_fail(WN_ECUDA, "device Gaussian fill failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        if (2 * info[0] >= t->count) {
            if (mt_draws) *mt_draws = 2 * (info[1] + 1);
            rc = tile_build_device(t, dR);
            timing_mark(c, st);
            break;
        }
        if (margin > 2000) { rc = wn_fail(WN_ESTATE, "device Gaussian fill could not accept enough attempts"); break; }
    }
    cudaFreeAsync(dR, st);
    cudaFreeAsync(dacc, st);
    if (rc) return rc;
    WN_CUDA(cudaStreamSynchronize(st));
    return timing_end(c);
}

extern "C" int wn_tile_upload(wn_tile *t, const float *N, int space)
{
    WN_REQUIRE(t && N, "wn_tile_upload: NULL argument");
    DeviceGuard g(t->ctx->device);
    WN_CUDA(cudaMemcpyAsync(t->d, N, t->count * sizeof(float),
                            space == WN_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, t->ctx->stream));
    int r = tile_finish(t);
    if (r) return r;
    if (space != WN_DEVICE) WN_CUDA(cudaStreamSynchronize(t->ctx->stream));
    return WN_OK;
}

extern "C" int wn_tile_download(const wn_tile *t, float *out, int space)
{
    WN_REQUIRE(t && out, "wn_tile_download: NULL argument");
    if (!t->built) return wn_fail(WN_ESTATE, "wn_tile_download: the tile has not been built");
    DeviceGuard g(t->ctx->device);
    WN_CUDA(cudaMemcpyAsync(out, t->d, t->count * sizeof(float),
                            space == WN_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, t->ctx->stream));
    if (space != WN_DEVICE) WN_CUDA(cudaStreamSynchronize(t->ctx->stream));
    return WN_OK;
}

extern "C" int wn_tile_device_ptr(const wn_tile *t, void **dptr)
{
    WN_REQUIRE(t && dptr, "wn_tile_device_ptr: NULL argument");
    *dptr = t->d;
    return WN_OK;
}

extern "C" int wn_tile_mark_built(wn_tile *t)
{
    WN_REQUIRE(t, "wn_tile_mark_built: tile is NULL");
    DeviceGuard g(t->ctx->device);
    return tile_finish(t);
}

#define WN_NEED_TILE(t, d, who)                                                                    \
    do {                                                                                           \
        WN_REQUIRE((t), who ": tile is NULL");                                                     \
        WN_REQUIRE((t)->dims == (d), who ": needs a %dD tile (this one is %dD)", (d), (t)->dims);  \
        if (!(t)->built) return wn_fail(WN_ESTATE, who ": the tile has not been built");           \
    } while (0)

#define WN_NEED_SPACE(space) WN_REQUIRE((space) == WN_HOST || (space) == WN_DEVICE, "bad space %d", (space))

static int make_bands(const float *scale, const float *weights, int nbands, float post, WnBands *b)
{
    WN_REQUIRE(nbands >= 1 && nbands <= WN_MAX_BANDS, "nbands must be in [1,%d] (got %d)", WN_MAX_BANDS, nbands);
    WN_REQUIRE(scale && weights, "band_scale / weights is NULL");
    b->nbands = nbands;
    for (int i = 0; i < WN_MAX_BANDS; ++i) { b->scale[i] = i < nbands ? scale[i] : 0.0f; b->weight[i] = i < nbands ? weights[i] : 0.0f; }
    b->post = post;
    return WN_OK;
}

// ---------------------------------------------------------------------------------------------------
// points
// ---------------------------------------------------------------------------------------------------
extern "C" int wn_eval2d_points(const wn_tile *t, const float *p, size_t count, float pre, float post, float *out, int space)
{
    WN_NEED_TILE(t, 2, "wn_eval2d_points");
    WN_NEED_SPACE(space);
    if (!count) return WN_OK;
    WN_REQUIRE(p && out, "wn_eval2d_points: NULL buffer");
    wn_ctx *c = t->ctx;
    DeviceGuard g(c->device);
    const WnTileView tv = tile_view(t);
    if (space == WN_DEVICE)
        return run_device(c, [&](cudaStream_t st) { return wn_launch_eval2d_points(tv, WnPointsAoS{p, pre}, 0, count, post, out, st); });
    ChunkIO io; io.in = p; io.in_item = 2 * sizeof(float); io.out = out;
    return run_chunked_host(c, count, kChunkSamples, io, [&](void *din, void *, float *dout, size_t, size_t cnt, cudaStream_t st) {
        return wn_launch_eval2d_points(tv, WnPointsAoS{(const float *)din, pre}, 0, cnt, post, dout, st);
    });
}

extern "C" int wn_multiband3d_points(const wn_tile *t, const float *p, size_t count, const float *band_scale,
                                     const float *weights, int nbands, float post, float *out, int space)
{
    WN_NEED_TILE(t, 3, "wn_multiband3d_points");
    WN_NEED_SPACE(space);
    WnBands b;
    int r = make_bands(band_scale, weights, nbands, post, &b);
    if (r) return r;
    if (!count) return WN_OK;
    WN_REQUIRE(p && out, "wn_multiband3d_points: NULL buffer");
    wn_ctx *c = t->ctx;
    DeviceGuard g(c->device);
    const WnTileView tv = tile_view(t);
    if (space == WN_DEVICE)
        return run_device(c, [&](cudaStream_t st) { return wn_launch_mb3d_points(tv, WnPointsAoS{p, 1.0f}, b, 0, count, out, st); });
    ChunkIO io; io.in = p; io.in_item = 3 * sizeof(float); io.out = out;
    return run_chunked_host(c, count, kChunkSamples, io, [&](void *din, void *, float *dout, size_t, size_t cnt, cudaStream_t st) {
        return wn_launch_mb3d_points(tv, WnPointsAoS{(const float *)din, 1.0f}, b, 0, cnt, dout, st);
    });
}

// Cook & DeRose App. 2: WMultibandNoise(p, s, normal, firstBand, nbands, w)
extern "C" int wn_wmultiband_points(const wn_tile *t, const float *p, size_t count, float s, const float *normal,
                                    int first_band, int nbands, const float *w, float *out, int space)
{
    WN_NEED_TILE(t, 3, "wn_wmultiband_points");
    WN_NEED_SPACE(space);
    WN_REQUIRE(nbands >= 1 && nbands <= WN_MAX_BANDS && w, "wn_wmultiband_points: nbands must be in [1,%d] and w not NULL", WN_MAX_BANDS);
    WN_REQUIRE(first_band > -60 && first_band + nbands < 60, "wn_wmultiband_points: band range out of range");
    if (!count) return WN_OK;
    WN_REQUIRE(p && out, "wn_wmultiband_points: NULL buffer");
    WnBands b;
    std::memset(&b, 0, sizeof(b));
    b.nbands = nbands; b.post = 1.0f;
    float variance = 0.0f;
    int active = 0;
    for (int k = 0; k < nbands; ++k) {
        b.scale[k] = (float)std::pow(2.0, first_band + k);
        b.weight[k] = w[k];
        variance += w[k] * w[k];
        if (active == k && s + first_band + k < 0) ++active;       // the listing's loop stops at the first band cut off
    }
    const double denom = variance ? std::sqrt(variance * (normal ? 0.296 : 0.210)) : 0.0;
    wn_ctx *c = t->ctx;
    DeviceGuard g(c->device);
    const WnTileView tv = tile_view(t);
    if (space == WN_DEVICE)
        return run_device(c, [&](cudaStream_t st) { return wn_launch_wmultiband(tv, p, count, b, active, normal, denom, out, st); });
    ChunkIO io; io.in = p; io.in_item = 3 * sizeof(float); io.out = out;
    return run_chunked_host(c, count, kChunkSamples / 4, io, [&](void *din, void *, float *dout, size_t, size_t cnt, cudaStream_t st) {
        return wn_launch_wmultiband(tv, (const float *)din, cnt, b, active, normal, denom, dout, st);
    });
}

// evaluate3D(p*pre)*post == multiband with one band {scale=pre, weight=1}: 0 + 1*v == v and v*post, bit-exact
extern "C" int wn_eval3d_points(const wn_tile *t, const float *p, size_t count, float pre, float post, float *out, int space)
{
    const float one = 1.0f;
    return wn_multiband3d_points(t, p, count, &pre, &one, 1, post, out, space);
}

extern "C" int wn_eval3d_projected_points(const wn_tile *t, const float *p, const float *normals, int shared,
                                          size_t count, float pre, float post, float *out, int space)
{
    WN_NEED_TILE(t, 3, "wn_eval3d_projected_points");
    WN_NEED_SPACE(space);
    if (!count) return WN_OK;
    WN_REQUIRE(p && out && normals, "wn_eval3d_projected_points: NULL buffer");
    wn_ctx *c = t->ctx;
    DeviceGuard g(c->device);
    const WnTileView tv = tile_view(t);
    float nrm[3] = { 0, 0, 0 };
    if (shared) { nrm[0] = normals[0]; nrm[1] = normals[1]; nrm[2] = normals[2]; }
    if (space == WN_DEVICE)
        return run_device(c, [&](cudaStream_t st) {
            return wn_launch_proj_points(tv, WnPointsAoS{p, pre}, shared ? nullptr : normals, nrm, 0, count, post, out, st);
        });
    ChunkIO io; io.in = p; io.in_item = 3 * sizeof(float); io.out = out;
    if (!shared) { io.aux = normals; io.aux_item = 3 * sizeof(float); }
    return run_chunked_host(c, count, kChunkSamples / 4, io, [&](void *din, void *daux, float *dout, size_t, size_t cnt, cudaStream_t st) {
        return wn_launch_proj_points(tv, WnPointsAoS{(const float *)din, pre}, shared ? nullptr : (const float *)daux, nrm,
                                     0, cnt, post, dout, st);
    });
}

// ---------------------------------------------------------------------------------------------------
// lattices and affine grids
// ---------------------------------------------------------------------------------------------------
static int check_axes(const float *xs, int nx, const float *ys, int ny, const float *zs, int nz, bool need_z)
{
    WN_REQUIRE(nx >= 0 && ny >= 0 && nz >= 0, "negative lattice size");
    WN_REQUIRE((xs || !nx) && (ys || !ny) && (!need_z || zs || !nz), "lattice axis pointer is NULL");
    return WN_OK;
}

extern "C" int wn_eval2d_lattice(const wn_tile *t, const float *xs, int nx, const float *ys, int ny, float pre, float post,
                                 float *out, int space)
{
    WN_NEED_TILE(t, 2, "wn_eval2d_lattice");
    WN_NEED_SPACE(space);
    int r = check_axes(xs, nx, ys, ny, nullptr, 1, false);
    if (r) return r;
    const size_t total = (size_t)nx * ny;
    if (!total) return WN_OK;
    WN_REQUIRE(out, "wn_eval2d_lattice: out is NULL");
    wn_ctx *c = t->ctx;
    DeviceGuard g(c->device);
    ParamWriter pw(c);
    if ((r = pw.reserve(((size_t)nx + ny) * sizeof(float) + 1024))) return r;
    WnLattice L{nullptr, nullptr, nullptr, nx, ny, 1};
    if ((r = pw.put(xs, nx * sizeof(float), (const void **)&L.xs))) return r;
    if ((r = pw.put(ys, ny * sizeof(float), (const void **)&L.ys))) return r;
    if ((r = pw.flush())) return r;
    const WnTileView tv = tile_view(t);
    if (space == WN_DEVICE)
        return run_device(c, [&](cudaStream_t st) { return wn_launch_eval2d_lattice(tv, L, pre, 0, total, post, out, st); });
    ChunkIO io; io.out = out;
    return run_chunked_host(c, total, kChunkSamples, io, [&](void *, void *, float *dout, size_t first, size_t cnt, cudaStream_t st) {
        return wn_launch_eval2d_lattice(tv, L, pre, first, cnt, post, dout, st);
    });
}

extern "C" int wn_multiband3d_lattice(const wn_tile *t, const float *xs, int nx, const float *ys, int ny, const float *zs,
                                      int nz, const float *band_scale, const float *weights, int nbands, float post,
                                      int mode, float *out, int space)
{
    WN_NEED_TILE(t, 3, "wn_multiband3d_lattice");
    WN_NEED_SPACE(space);
    WN_REQUIRE(mode == WN_EVAL_FAST || mode == WN_EVAL_EXACT, "bad mode %d", mode);
    int r = check_axes(xs, nx, ys, ny, zs, nz, true);
    if (r) return r;
    WnBands b;
    if ((r = make_bands(band_scale, weights, nbands, post, &b))) return r;
    const size_t slice = (size_t)nx * ny, total = slice * nz;
    if (!total) return WN_OK;
    WN_REQUIRE(out, "wn_multiband3d_lattice: out is NULL");
    wn_ctx *c = t->ctx;
    DeviceGuard g(c->device);
    // FAST + device output: the chain (parameter upload, axis tables, period blocks) goes to the side stream so it
    // overlaps the previous call's main kernel; WN_SIDE_CHAIN=0 keeps everything on the compute stream (A/B runs)
    static const bool side_env = [] { const char *e = getenv("WN_SIDE_CHAIN"); return !e || atoi(e) != 0; }();
    // Measured on config 3 shards: 157 -> 147 us (1/8 of the volume), 291 -> 280 us (1/4), 575 -> 557 us (1/2).  For the
    // whole 1024^3 it cost 2 % in round 1 (1.12 -> 1.14 ms: the 512^3 period block made the chain bandwidth-bound); with
    // the 256^3 / 128^3 chain of round 2 it is 0.825 -> 0.819 ms (profiles/r2_side_chain_full_size.log), so the side
    // stream is used up to 2^30 samples per call.
    size_t side_max = (size_t)1 << 30;
    if (const char *e = getenv("WN_SIDE_MAX_LOG2")) side_max = (size_t)1 << atoi(e);     // read per call (A/B runs)
    const bool use_side = mode == WN_EVAL_FAST && space == WN_DEVICE && side_env && total <= side_max;
    ParamWriter pw(c, use_side);
    if ((r = pw.reserve(((size_t)nx + ny + nz) * sizeof(float) + 2048))) return r;
    WnLattice L{nullptr, nullptr, nullptr, nx, ny, nz};
    if ((r = pw.put(xs, nx * sizeof(float), (const void **)&L.xs))) return r;
    if ((r = pw.put(ys, ny * sizeof(float), (const void **)&L.ys))) return r;
    if ((r = pw.put(zs, nz * sizeof(float), (const void **)&L.zs))) return r;
    if (use_side && t->side_seen != t->version) {
        // the tile changed on the compute stream since the side stream last read it
        WN_CUDA(cudaEventRecord(c->ev_tile, c->stream));
        WN_CUDA(cudaStreamWaitEvent(c->side, c->ev_tile, 0));
        t->side_seen = t->version;
    }
    if ((r = pw.flush())) return r;
    const WnTileView tv = tile_view(t);
    if (mode == WN_EVAL_EXACT) {
        if (space == WN_DEVICE)
            return run_device(c, [&](cudaStream_t st) { return wn_launch_mb3d_lattice_exact(tv, L, b, 0, total, out, st); });
        ChunkIO io; io.out = out;
        return run_chunked_host(c, total, kChunkSamples, io, [&](void *, void *, float *dout, size_t first, size_t cnt, cudaStream_t st) {
            return wn_launch_mb3d_lattice_exact(tv, L, b, first, cnt, dout, st);
        });
    }
    WnFastPlan plan;
    // Programmatic dependent launch helps when everything is on one stream (N=8-sized shard: 176 -> 157 us) and hurts
    // with the side-stream chain (157 vs 145 us): early-launched CTAs of the next main kernel then sit on SM slots the
    // high-priority chain needs during the tail of the current one.
    static const int pdl_side_chain = [] { const char *e = getenv("WN_PDL_SIDE_CHAIN"); return e ? atoi(e) : 0; }();
    static const int pdl_side_main = [] { const char *e = getenv("WN_PDL_SIDE_MAIN"); return e ? atoi(e) : 0; }();
    plan.pdl = use_side ? pdl_side_chain : 1;
    plan.axes_cache = &t->plan_cache;
    cudaStream_t chain = use_side ? c->side : c->stream;
    const int gen = (int)(c->side_calls & 1);
    if (use_side) {
        ++c->side_calls;
        wn_ctx::Deferred &d = c->deferred[gen];
        if (d.pending) {                               // scratch of the call two side-chain calls ago
            WN_CUDA(cudaStreamWaitEvent(c->side, c->ev_main[gen], 0));
            if (d.tab) WN_CUDA(cudaFreeAsync(d.tab, c->side));
            if (d.P) WN_CUDA(cudaFreeAsync(d.P, c->side));
            d = wn_ctx::Deferred();
        }
    }
    int np = wn_mb3d_fast_prepare(tv, L, xs, ys, zs, b, &plan, chain);
    if (np < 0) {
        wn_mb3d_fast_finish(&plan, chain);
        return wn_fail(WN_ECUDA, "fast lattice prepare failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    c->launches += (uint64_t)np;
    if (use_side) {
        WN_CUDA(cudaEventRecord(c->ev_side, c->side));
        WN_CUDA(cudaStreamWaitEvent(c->stream, c->ev_side, 0));
        plan.pdl = pdl_side_main;
    }
    if (space == WN_DEVICE) {
        cudaEvent_t ta = nullptr, tb = nullptr;
        if (c->time_main && cudaEventCreate(&ta) == cudaSuccess && cudaEventCreate(&tb) == cudaSuccess) {
            c->main_ev.push_back(ta); c->main_ev.push_back(tb);
            WN_CUDA(cudaEventRecord(ta, c->stream));
        }
        r = run_device(c, [&](cudaStream_t st) { return wn_mb3d_fast_run(tv, L, ys, zs, b, nullptr, &plan, 0, nz, out, st); });
        if (tb) cudaEventRecord(tb, c->stream);
    } else {
        // HOST: chunk by whole z slices so each chunk is a lattice slab
        size_t slices_per_chunk = std::max<size_t>(1, kChunkSamples / slice);
        ChunkIO io; io.out = out; io.out_item = slice * sizeof(float);
        r = run_chunked_host(c, (size_t)nz, slices_per_chunk, io, [&](void *, void *, float *dout, size_t first, size_t cnt, cudaStream_t st) {
            return wn_mb3d_fast_run(tv, L, ys, zs, b, nullptr, &plan, (int)first, (int)cnt, dout, st);
        });
    }
    if (use_side) {
        wn_ctx::Deferred &d = c->deferred[gen];
        wn_mb3d_fast_detach(&plan, &d.tab, &d.P);
        d.pending = true;
        cudaError_t e = cudaEventRecord(c->ev_main[gen], c->stream);
        if (e != cudaSuccess && r == WN_OK) r = wn_fail(WN_ECUDA, "event record failed: %s", cudaGetErrorString(e));
    } else {
        wn_mb3d_fast_finish(&plan, c->stream);
    }
    return r;
}

extern "C" int wn_debug_fold_plan(const float *xs, int nx, const float *ys, int ny, const float *zs, int nz,
                                  const float *band_scale, int nbands, int tile_n, int *band_folded, int block[3],
                                  int *nfolded)
{
    WN_REQUIRE(band_folded && block && nfolded, "wn_debug_fold_plan: NULL output");
    WN_REQUIRE(tile_n >= 2, "wn_debug_fold_plan: tile_n must be >= 2 (got %d)", tile_n);
    int r = check_axes(xs, nx, ys, ny, zs, nz, true);
    if (r) return r;
    WnBands b;
    std::vector<float> ones((size_t)std::max(nbands, 1), 1.0f);
    if ((r = make_bands(band_scale, ones.data(), nbands, 1.0f, &b))) return r;
    *nfolded = wn_mb3d_fast_plan_host(xs, nx, ys, ny, zs, nz, b, tile_n, band_folded, block);
    return WN_OK;
}

extern "C" int wn_debug_axis_entries(wn_ctx *c, const float *coords, int count, float band_scale, float *weights3,
                                     int32_t *first_cell)
{
    WN_REQUIRE(c && coords && weights3 && first_cell, "wn_debug_axis_entries: NULL argument");
    WN_REQUIRE(count > 0 && count <= (1 << 24), "wn_debug_axis_entries: count out of range");
    DeviceGuard g(c->device);
    float *dc = nullptr;
    float4 *de = nullptr;
    WN_CUDA(wn_scratch_alloc((void **)&dc, (size_t)count * sizeof(float), c->stream));
    WN_CUDA(wn_scratch_alloc((void **)&de, (size_t)count * sizeof(float4), c->stream));
    WN_CUDA(cudaMemcpyAsync(dc, coords, (size_t)count * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    c->launches += (uint64_t)wn_mb3d_debug_axis_table(dc, count, band_scale, de, c->stream);
    std::vector<float4> h((size_t)count);
    WN_CUDA(cudaMemcpyAsync(h.data(), de, (size_t)count * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
    WN_CUDA(cudaStreamSynchronize(c->stream));
    cudaFreeAsync(dc, c->stream);
    cudaFreeAsync(de, c->stream);
    for (int i = 0; i < count; ++i) {
        weights3[3 * i] = h[i].x; weights3[3 * i + 1] = h[i].y; weights3[3 * i + 2] = h[i].z;
        std::memcpy(&first_cell[i], &h[i].w, sizeof(int32_t));
    }
    return WN_OK;
}

static int make_affine(ParamWriter &pw, const float origin[3], const float e1[3], const float *us, int nu,
                       const float e2[3], const float *vs, int nv, float pre, WnAffine *A)
{
    WN_REQUIRE(origin && e1 && e2, "affine grid: origin/e1/e2 is NULL");
    WN_REQUIRE(nu >= 0 && nv >= 0 && (us || !nu) && (vs || !nv), "affine grid: bad axes");
    int r;
    if ((r = pw.reserve(((size_t)nu + nv) * sizeof(float) + 1024))) return r;
    A->nu = nu; A->nv = nv; A->pre = pre;
    for (int i = 0; i < 3; ++i) { A->o[i] = origin[i]; A->e1[i] = e1[i]; A->e2[i] = e2[i]; }
    if ((r = pw.put(us, nu * sizeof(float), (const void **)&A->us))) return r;
    if ((r = pw.put(vs, nv * sizeof(float), (const void **)&A->vs))) return r;
    return pw.flush();
}

extern "C" int wn_eval3d_projected_grid(const wn_tile *t, const float origin[3], const float e1[3], const float *us, int nu,
                                        const float e2[3], const float *vs, int nv, const float normal[3], float pre,
                                        float post, float *out, int space)
{
    WN_NEED_TILE(t, 3, "wn_eval3d_projected_grid");
    WN_NEED_SPACE(space);
    WN_REQUIRE(normal, "wn_eval3d_projected_grid: normal is NULL");
    wn_ctx *c = t->ctx;
    DeviceGuard g(c->device);
    ParamWriter pw(c);
    WnAffine A;
    int r = make_affine(pw, origin, e1, us, nu, e2, vs, nv, pre, &A);
    if (r) return r;
    const size_t total = (size_t)nu * nv;
    if (!total) return WN_OK;
    WN_REQUIRE(out, "wn_eval3d_projected_grid: out is NULL");
    const WnTileView tv = tile_view(t);
    const float nrm[3] = { normal[0], normal[1], normal[2] };
    if (space == WN_DEVICE)
        return run_device(c, [&](cudaStream_t st) { return wn_launch_proj_affine(tv, A, nrm, 0, total, post, out, st); });
    ChunkIO io; io.out = out;
    return run_chunked_host(c, total, kChunkSamples / 4, io, [&](void *, void *, float *dout, size_t first, size_t cnt, cudaStream_t st) {
        return wn_launch_proj_affine(tv, A, nrm, first, cnt, post, dout, st);
    });
}

extern "C" int wn_eval3d_grid(const wn_tile *t, const float origin[3], const float e1[3], const float *us, int nu,
                              const float e2[3], const float *vs, int nv, float pre, float post, float *out, int space)
{
    WN_NEED_TILE(t, 3, "wn_eval3d_grid");
    WN_NEED_SPACE(space);
    wn_ctx *c = t->ctx;
    DeviceGuard g(c->device);
    ParamWriter pw(c);
    WnAffine A;
    int r = make_affine(pw, origin, e1, us, nu, e2, vs, nv, pre, &A);
    if (r) return r;
    const size_t total = (size_t)nu * nv;
    if (!total) return WN_OK;
    WN_REQUIRE(out, "wn_eval3d_grid: out is NULL");
    const WnTileView tv = tile_view(t);
    WnBands b;
    const float one = 1.0f;
    make_bands(&one, &one, 1, post, &b);        // the pre-scale is applied by the affine generator
    if (space == WN_DEVICE)
        return run_device(c, [&](cudaStream_t st) { return wn_launch_mb3d_affine(tv, A, b, 0, total, out, st); });
    ChunkIO io; io.out = out;
    return run_chunked_host(c, total, kChunkSamples, io, [&](void *, void *, float *dout, size_t first, size_t cnt, cudaStream_t st) {
        return wn_launch_mb3d_affine(tv, A, b, first, cnt, dout, st);
    });
}

// ---------------------------------------------------------------------------------------------------
// Perlin
// ---------------------------------------------------------------------------------------------------
struct wn_perlin {
    wn_ctx *ctx = nullptr;
    int32_t *d = nullptr;
    int fast = 0;                        // WN_PERLIN_F32: FP32 kernel for the float batch calls
};

extern "C" int wn_perlin_create(wn_ctx *c, const int32_t perm[512], wn_perlin **out)
{
    WN_REQUIRE(c && perm && out, "wn_perlin_create: NULL argument");
    *out = nullptr;
    for (int i = 0; i < 512; ++i)
        WN_REQUIRE(perm[i] >= 0 && perm[i] < 256, "wn_perlin_create: perm[%d]=%d outside [0,255]", i, perm[i]);
    DeviceGuard g(c->device);
    wn_perlin *p = new (std::nothrow) wn_perlin();
    if (!p) return wn_fail(WN_ENOMEM, "out of host memory");
    p->ctx = c;
    if (cudaMalloc(&p->d, 512 * sizeof(int32_t)) != cudaSuccess) { delete p; cudaGetLastError(); return wn_fail(WN_ENOMEM, "cudaMalloc failed"); }
    WN_CUDA(cudaMemcpyAsync(p->d, perm, 512 * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    WN_CUDA(cudaStreamSynchronize(c->stream));
    *out = p;
    return WN_OK;
}

extern "C" int wn_perlin_destroy(wn_perlin *p)
{
    if (!p) return WN_OK;
    DeviceGuard g(p->ctx->device);
    cudaStreamSynchronize(p->ctx->stream);
    cudaFree(p->d);
    delete p;
    return WN_OK;
}

extern "C" int wn_perlin_set_precision(wn_perlin *pn, int precision)
{
    WN_REQUIRE(pn, "wn_perlin_set_precision: perlin is NULL");
    WN_REQUIRE(precision == WN_PERLIN_F64 || precision == WN_PERLIN_F32, "wn_perlin_set_precision: bad precision %d", precision);
    pn->fast = precision == WN_PERLIN_F32;
    return WN_OK;
}

// double coordinates in, double noise out: the scalar signature of PerlinNoise::noise (PerlinNoise.hpp:36, perlin.h:42)
extern "C" int wn_perlin_points_f64(const wn_perlin *pn, const double *p, size_t count, double *out, int space)
{
    WN_REQUIRE(pn, "wn_perlin_points_f64: perlin is NULL");
    WN_NEED_SPACE(space);
    if (!count) return WN_OK;
    WN_REQUIRE(p && out, "wn_perlin_points_f64: NULL buffer");
    wn_ctx *c = pn->ctx;
    DeviceGuard g(c->device);
    const int32_t *perm = pn->d;
    if (space == WN_DEVICE)
        return run_device(c, [&](cudaStream_t st) { return wn_launch_perlin_points_f64(perm, p, count, out, st); });
    ChunkIO io; io.in = p; io.in_item = 3 * sizeof(double); io.out = out; io.out_item = sizeof(double);
    return run_chunked_host(c, count, kChunkSamples / 2, io, [&](void *din, void *, float *dout, size_t, size_t cnt, cudaStream_t st) {
        return wn_launch_perlin_points_f64(perm, (const double *)din, cnt, (double *)dout, st);
    });
}

extern "C" int wn_perlin_points(const wn_perlin *pn, const float *p, size_t count, float pre, float *out, int space)
{
    WN_REQUIRE(pn, "wn_perlin_points: perlin is NULL");
    WN_NEED_SPACE(space);
    if (!count) return WN_OK;
    WN_REQUIRE(p && out, "wn_perlin_points: NULL buffer");
    wn_ctx *c = pn->ctx;
    DeviceGuard g(c->device);
    const int32_t *perm = pn->d;
    const int fast = pn->fast;
    if (space == WN_DEVICE)
        return run_device(c, [&](cudaStream_t st) { return wn_launch_perlin_points(perm, WnPointsAoS{p, pre}, 0, count, out, fast, st); });
    ChunkIO io; io.in = p; io.in_item = 3 * sizeof(float); io.out = out;
    return run_chunked_host(c, count, kChunkSamples, io, [&](void *din, void *, float *dout, size_t, size_t cnt, cudaStream_t st) {
        return wn_launch_perlin_points(perm, WnPointsAoS{(const float *)din, pre}, 0, cnt, dout, fast, st);
    });
}

extern "C" int wn_perlin_lattice(const wn_perlin *pn, const float *xs, int nx, const float *ys, int ny, const float *zs, int nz,
                                 float *out, int space)
{
    WN_REQUIRE(pn, "wn_perlin_lattice: perlin is NULL");
    WN_NEED_SPACE(space);
    int r = check_axes(xs, nx, ys, ny, zs, nz, true);
    if (r) return r;
    const size_t total = (size_t)nx * ny * nz;
    if (!total) return WN_OK;
    WN_REQUIRE(out, "wn_perlin_lattice: out is NULL");
    wn_ctx *c = pn->ctx;
    DeviceGuard g(c->device);
    ParamWriter pw(c);
    if ((r = pw.reserve(((size_t)nx + ny + nz) * sizeof(float) + 2048))) return r;
    WnLattice L{nullptr, nullptr, nullptr, nx, ny, nz};
    if ((r = pw.put(xs, nx * sizeof(float), (const void **)&L.xs))) return r;
    if ((r = pw.put(ys, ny * sizeof(float), (const void **)&L.ys))) return r;
    if ((r = pw.put(zs, nz * sizeof(float), (const void **)&L.zs))) return r;
    if ((r = pw.flush())) return r;
    const int32_t *perm = pn->d;
    const int fast = pn->fast;
    if (space == WN_DEVICE)
        return run_device(c, [&](cudaStream_t st) { return wn_launch_perlin_lattice(perm, L, 0, total, out, fast, st); });
    ChunkIO io; io.out = out;
    return run_chunked_host(c, total, kChunkSamples, io, [&](void *, void *, float *dout, size_t first, size_t cnt, cudaStream_t st) {
        return wn_launch_perlin_lattice(perm, L, first, cnt, dout, fast, st);
    });
}

extern "C" int wn_perlin_grid(const wn_perlin *pn, const float origin[3], const float e1[3], const float *us, int nu,
                              const float e2[3], const float *vs, int nv, float pre, float *out, int space)
{
    WN_REQUIRE(pn, "wn_perlin_grid: perlin is NULL");
    WN_NEED_SPACE(space);
    wn_ctx *c = pn->ctx;
    DeviceGuard g(c->device);
    ParamWriter pw(c);
    WnAffine A;
    int r = make_affine(pw, origin, e1, us, nu, e2, vs, nv, pre, &A);
    if (r) return r;
    const size_t total = (size_t)nu * nv;
    if (!total) return WN_OK;
    WN_REQUIRE(out, "wn_perlin_grid: out is NULL");
    const int32_t *perm = pn->d;
    const int fast = pn->fast;
    if (space == WN_DEVICE)
        return run_device(c, [&](cudaStream_t st) { return wn_launch_perlin_affine(perm, A, 0, total, out, fast, st); });
    ChunkIO io; io.out = out;
    return run_chunked_host(c, total, kChunkSamples, io, [&](void *, void *, float *dout, size_t first, size_t cnt, cudaStream_t st) {
        return wn_launch_perlin_affine(perm, A, first, cnt, dout, fast, st);
    });
}

// ---------------------------------------------------------------------------------------------------
// texture hooks
// ---------------------------------------------------------------------------------------------------
extern "C" int wn_wavelet_texture_values(const wn_tile *t, const float *p, size_t count, double scale, int octave,
                                         float *grey, int space)
{
    WN_NEED_TILE(t, 3, "wn_wavelet_texture_values");
    WN_NEED_SPACE(space);
    if (!count) return WN_OK;
    WN_REQUIRE(p && grey, "wn_wavelet_texture_values: NULL buffer");
    wn_ctx *c = t->ctx;
    DeviceGuard g(c->device);
    // texture.h:77-80,84: octave_scale = std::pow(2.0f, octave) narrowed to float; pos *= octave_scale*2.0f
    const float octave_scale = (float)std::pow(2.0, (double)octave);
    const float oct2 = octave_scale * 2.0f;
    const float inv_std = 1.0f / std::sqrt(0.18402f);
    const WnTileView tv = tile_view(t);
    if (space == WN_DEVICE)
        return run_device(c, [&](cudaStream_t st) { return wn_launch_wavelet_texture(tv, p, count, scale, oct2, inv_std, grey, st); });
    ChunkIO io; io.in = p; io.in_item = 3 * sizeof(float); io.out = grey;
    return run_chunked_host(c, count, kChunkSamples, io, [&](void *din, void *, float *dout, size_t, size_t cnt, cudaStream_t st) {
        return wn_launch_wavelet_texture(tv, (const float *)din, cnt, scale, oct2, inv_std, dout, st);
    });
}

extern "C" int wn_wavelet_texture2d_values(const wn_tile *t, const float *p, size_t count, double scale, int octave,
                                           float *grey, int space)
{
    WN_NEED_TILE(t, 2, "wn_wavelet_texture2d_values");
    WN_NEED_SPACE(space);
    if (!count) return WN_OK;
    WN_REQUIRE(p && grey, "wn_wavelet_texture2d_values: NULL buffer");
    wn_ctx *c = t->ctx;
    DeviceGuard g(c->device);
    const float octave_scale = (float)std::pow(2.0, (double)octave);       // texture.h:91
    const float oct2 = octave_scale * 2.0f;
    const float inv_std = 1.0f / std::sqrt(0.19686f);                      // texture.h:97
    const WnTileView tv = tile_view(t);
    if (space == WN_DEVICE)
        return run_device(c, [&](cudaStream_t st) { return wn_launch_wavelet_texture2d(tv, p, count, scale, oct2, inv_std, grey, st); });
    ChunkIO io; io.in = p; io.in_item = 3 * sizeof(float); io.out = grey;
    return run_chunked_host(c, count, kChunkSamples, io, [&](void *din, void *, float *dout, size_t, size_t cnt, cudaStream_t st) {
        return wn_launch_wavelet_texture2d(tv, (const float *)din, cnt, scale, oct2, inv_std, dout, st);
    });
}

extern "C" int wn_perlin_texture_values(const wn_perlin *pn, const float *p, size_t count, double scale, int octave,
                                        float *grey, int space)
{
    WN_REQUIRE(pn, "wn_perlin_texture_values: perlin is NULL");
    WN_NEED_SPACE(space);
    if (!count) return WN_OK;
    WN_REQUIRE(p && grey, "wn_perlin_texture_values: NULL buffer");
    wn_ctx *c = pn->ctx;
    DeviceGuard g(c->device);
    const float octave_scale = (float)std::pow(2.0, (double)octave);
    const float scale_f = (float)scale;                      // vec3 * double -> operator*(vec3, float)
    const int32_t *perm = pn->d;
    if (space == WN_DEVICE)
        return run_device(c, [&](cudaStream_t st) { return wn_launch_perlin_texture(perm, p, count, scale_f, octave_scale, grey, st); });
    ChunkIO io; io.in = p; io.in_item = 3 * sizeof(float); io.out = grey;
    return run_chunked_host(c, count, kChunkSamples, io, [&](void *din, void *, float *dout, size_t, size_t cnt, cudaStream_t st) {
        return wn_launch_perlin_texture(perm, (const float *)din, cnt, scale_f, octave_scale, dout, st);
    });
}

// ---------------------------------------------------------------------------------------------------
// stats
// ---------------------------------------------------------------------------------------------------
extern "C" int wn_stats_compute(wn_ctx *c, const float *data, size_t count, int space, wn_stats *out)
{
    WN_REQUIRE(c && out, "wn_stats_compute: NULL argument");
    WN_NEED_SPACE(space);
    wn_stats s;
    s.avg = 0.0f; s.var = 0.0f;
    s.min_val = 3.402823466e+38f; s.max_val = -3.402823466e+38f;      // DataStats defaults, WaveletNoise.h:12-17
    s.count_nan_inf = 0; s.energy = 0.0f;
    *out = s;
    if (!count) return WN_OK;
    WN_REQUIRE(data, "wn_stats_compute: data is NULL");
    DeviceGuard g(c->device);
    int r = buf_reserve(c->stats_partial, 4 * WN_STATS_BLOCKS * sizeof(double));
    if (r) return r;
    const float *dd = data;
    if (space == WN_HOST) {
        if ((r = buf_reserve(c->in[0], count * sizeof(float)))) return r;
        WN_CUDA(cudaMemcpyAsync(c->in[0].p, data, count * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        dd = (const float *)c->in[0].p;
    }
    c->launches += (uint64_t)wn_launch_stats(dd, count, (double *)c->stats_partial.p, c->stream);
    WN_CUDA(cudaGetLastError());
    std::vector<double> part(4 * WN_STATS_BLOCKS);
    WN_CUDA(cudaMemcpyAsync(part.data(), c->stats_partial.p, part.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    WN_CUDA(cudaStreamSynchronize(c->stream));
    double sum = 0.0, sq = 0.0;
    for (int b = 0; b < WN_STATS_BLOCKS; ++b) {
        sum += part[4 * b]; sq += part[4 * b + 1];
        s.min_val = std::min(s.min_val, (float)part[4 * b + 2]);
        s.max_val = std::max(s.max_val, (float)part[4 * b + 3]);
    }
    s.avg = (float)(sum / (double)count);
    s.var = (float)((sq / (double)count) - (double)s.avg * s.avg);
    *out = s;
    return WN_OK;
}
