// wn_group.cu -- single-process multi-GPU layer of the C ABI (include/wn_b200.h, "device groups").
//
// north_star: "The multi-sample volume and image workloads shard naturally by z-slab or image row-band across the 8 GPUs
// of one box.  Each GPU holds a replica of the tile, broadcast once over NVLink with NCCL, and output is gathered to rank 0
// only for file write."  A wn_group is that: one wn_ctx per GPU driven by ONE host thread, a tile replicated with
// ncclBroadcast (single-process communicators from ncclCommInitAll), and sharded evaluation calls that only enqueue on
// every GPU's stream before anything is waited for -- there is no data-path collective, every sample reads only its
// GPU's tile replica.  The reference has no counterpart (it is single-threaded); the loops that
// are sharded are experient/main.cpp:45-58 (volume, by z), :74-87 (projected plane, by row) and main.cpp:175-204
// (render rows, through the batched texture hook).
//
// NCCL is loaded at run time (dlopen "libnccl.so.2"): the library has no link-time dependency on it, a process that
// already holds an NCCL (PyTorch bundles its own) shares that copy, and a one-GPU group never touches it.
#include "../../include/wn_b200.h"
#include "wn_internal.h"

#include <dlfcn.h>

#include <algorithm>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

int wn_set_error(int code, const char *fmt, ...);               // wn_capi.cu: fills the thread's wn_last_error text
cudaStream_t wn_ctx_stream_internal(const wn_ctx *ctx);          // wn_capi.cu

namespace {

struct NcclApi {
    void *lib = nullptr;
    int (*CommInitAll)(void **, int, const int *) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool load()
    {
        if (lib) return true;
        for (const char *name : { "libnccl.so.2", "libnccl.so" }) {
            lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) return false;
        CommInitAll = reinterpret_cast<decltype(CommInitAll)>(dlsym(lib, "ncclCommInitAll"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
        Broadcast = reinterpret_cast<decltype(Broadcast)>(dlsym(lib, "ncclBroadcast"));
        GroupStart = reinterpret_cast<decltype(GroupStart)>(dlsym(lib, "ncclGroupStart"));
        GroupEnd = reinterpret_cast<decltype(GroupEnd)>(dlsym(lib, "ncclGroupEnd"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
        return CommInitAll && CommDestroy && Broadcast && GroupStart && GroupEnd && GetErrorString;
    }
};
const int kNcclFloat = 7;                                        // ncclFloat32 (nccl.h)

struct RankBuf { float *p = nullptr; size_t cap = 0; };          // floats

// One host thread per GPU of a group: planning and enqueueing a rank's share of a call costs tens of microseconds of
// host time, which a single thread would pay N times in a row while the GPUs wait (measured: 8 ranks enqueued by one
// thread = 0.41 ms per 1024^3 call, slower than 4).  The caller's thread hands every rank's closure to its worker and
// waits for all of them to have ENQUEUED; the GPU work itself stays asynchronous.
struct Worker {
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    std::function<int()> job;
    bool has_job = false, stop = false;
    int rc = 0;
    std::string err;
    void loop(int device)
    {
        cudaSetDevice(device);
        std::unique_lock<std::mutex> lk(m);
        for (;;) {
            cv.wait(lk, [&] { return has_job || stop; });
            if (stop) return;
            lk.unlock();
            const int r = job();
            std::string e = r != WN_OK ? wn_last_error() : "";
            lk.lock();
            rc = r; err.swap(e);
            has_job = false;
            cv.notify_all();
        }
    }
};

// the group functions switch the current device while they enqueue; the caller's device is restored on return
struct DeviceRestore {
    int d = 0;
    DeviceRestore() { if (cudaGetDevice(&d) != cudaSuccess) { cudaGetLastError(); d = -1; } }
    ~DeviceRestore() { if (d >= 0) cudaSetDevice(d); }
};

}  // namespace

struct wn_group {
    int n = 0;
    std::vector<int> dev;
    std::vector<wn_ctx *> ctx;
    std::vector<void *> comm;                                    // ncclComm_t per rank (empty for one GPU)
    NcclApi nccl;
    std::vector<RankBuf> out, in;                                // per-rank device output / input shards
    std::vector<size_t> out_count;                               // floats of the last sharded call per rank
    std::vector<cudaEvent_t> ev0, ev1;                           // timing of the last sharded call per rank
    std::vector<std::unique_ptr<Worker>> workers;               // one per rank when n > 1
};

struct wn_gtile {
    wn_group *g = nullptr;
    std::vector<wn_tile *> t;
    size_t count = 0;
};

namespace {

#define WG_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return wn_set_error(WN_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define WG_REQUIRE(cond, ...)                                                                      \
    do {                                                                                           \
        if (!(cond)) return wn_set_error(WN_EINVAL, __VA_ARGS__);                                  \
    } while (0)
#define WG_OK(call)                                                                                \
    do {                                                                                           \
        int r_ = (call);                                                                           \
        if (r_ != WN_OK) return r_;                                                                \
    } while (0)

int reserve(wn_group *g, std::vector<RankBuf> &bufs, int rank, size_t floats)
{
    RankBuf &b = bufs[rank];
    if (floats <= b.cap) return WN_OK;
    WG_CUDA(cudaSetDevice(g->dev[rank]));
    if (b.p) { WG_CUDA(cudaStreamSynchronize(wn_ctx_stream_internal(g->ctx[rank]))); WG_CUDA(cudaFree(b.p)); }
    b.p = nullptr; b.cap = 0;
    const size_t want = floats + floats / 16 + 64;
    if (cudaMalloc(&b.p, want * sizeof(float)) != cudaSuccess) {
        cudaGetLastError();
        return wn_set_error(WN_ENOMEM, "cudaMalloc of a %zu-float shard buffer failed on device %d", want, g->dev[rank]);
    }
    b.cap = want;
    return WN_OK;
}

// [begin, end) of `total` units for rank: contiguous, remainders to the lowest ranks
void slab_range(size_t total, int rank, int world, size_t *begin, size_t *end)
{
    const size_t base = total / world, rem = total % world;
    *begin = rank * base + std::min<size_t>(rank, rem);
    *end = *begin + base + ((size_t)rank < rem ? 1 : 0);
}

// z indices a rank owns: contiguous slab, or 32-slice chunks dealt round-robin (the z extent of one CTA brick; keeps the
// slices a rank owns congruent modulo the periods of the folded bands)
std::vector<int> shard_indices(int total, int rank, int world, int sharding)
{
    std::vector<int> idx;
    if (sharding == WN_SHARD_SLAB) {
        size_t b, e;
        slab_range((size_t)total, rank, world, &b, &e);
        for (size_t k = b; k < e; ++k) idx.push_back((int)k);
    } else {
        const int chunk = 32;
        for (int c = rank; c * chunk < total; c += world)
            for (int k = c * chunk; k < std::min((c + 1) * chunk, total); ++k) idx.push_back(k);
    }
    return idx;
}

// f(rank) for every rank, concurrently on the rank workers (inline for a one-GPU group); returns the first failure
template <class F>
int for_ranks(wn_group *g, F f)
{
    if (g->workers.empty()) {
        for (int r = 0; r < g->n; ++r) WG_OK(f(r));
        return WN_OK;
    }
    for (int r = 0; r < g->n; ++r) {
        Worker &w = *g->workers[r];
        std::lock_guard<std::mutex> lk(w.m);
        w.job = [&f, r] { return f(r); };
        w.has_job = true;
        w.cv.notify_all();
    }
    int rc = WN_OK;
    for (int r = 0; r < g->n; ++r) {
        Worker &w = *g->workers[r];
        std::unique_lock<std::mutex> lk(w.m);
        w.cv.wait(lk, [&] { return !w.has_job; });
        if (w.rc != WN_OK && rc == WN_OK) rc = wn_set_error(w.rc, "%s", w.err.c_str());
    }
    return rc;
}

// waits for every rank; *ms (nullable) = max over ranks of the GPU time between begin_timing and end_timing
int finish(wn_group *g, float *ms)
{
    float worst = 0.0f;
    for (int r = 0; r < g->n; ++r) {
        WG_CUDA(cudaSetDevice(g->dev[r]));
        WG_CUDA(cudaStreamSynchronize(wn_ctx_stream_internal(g->ctx[r])));
        float t = 0.0f;
        if (cudaEventElapsedTime(&t, g->ev0[r], g->ev1[r]) == cudaSuccess) worst = std::max(worst, t);
        else cudaGetLastError();
    }
    if (ms) *ms = worst;
    return WN_OK;
}

}  // namespace

// host-only: the z indices (or rows) rank `rank` of `world` owns under `sharding` -- what the sharded calls use
extern "C" int wn_debug_shard_indices(int total, int rank, int world, int sharding, int *indices, int capacity, int *count)
{
    WG_REQUIRE(count && total >= 0 && world > 0 && rank >= 0 && rank < world, "wn_debug_shard_indices: bad argument");
    WG_REQUIRE(sharding == WN_SHARD_SLAB || sharding == WN_SHARD_CYCLIC, "bad sharding %d", sharding);
    const std::vector<int> idx = shard_indices(total, rank, world, sharding);
    *count = (int)idx.size();
    for (int i = 0; i < (int)idx.size() && i < capacity && indices; ++i) indices[i] = idx[i];
    return WN_OK;
}

extern "C" int wn_group_create(int ngpus, const int *devices, wn_group **out)
{
    WG_REQUIRE(out, "wn_group_create: out is NULL");
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return wn_set_error(WN_ENODEVICE, "no CUDA device available; this library has no CPU fallback");
    }
    if (ngpus <= 0) ngpus = count;
    WG_REQUIRE(ngpus <= count, "wn_group_create: %d GPUs requested, %d visible", ngpus, count);
    wn_group *g = new (std::nothrow) wn_group();
    if (!g) return wn_set_error(WN_ENOMEM, "out of host memory");
    g->n = ngpus;
    g->dev.resize(ngpus); g->ctx.assign(ngpus, nullptr); g->out.resize(ngpus); g->in.resize(ngpus);
    g->out_count.assign(ngpus, 0); g->ev0.assign(ngpus, nullptr); g->ev1.assign(ngpus, nullptr);
    int prev = 0;
    cudaGetDevice(&prev);
    int rc = WN_OK;
    for (int r = 0; r < ngpus && rc == WN_OK; ++r) {
        g->dev[r] = devices ? devices[r] : r;
        rc = wn_ctx_create(g->dev[r], &g->ctx[r]);
        if (rc == WN_OK && (cudaSetDevice(g->dev[r]) != cudaSuccess || cudaEventCreate(&g->ev0[r]) != cudaSuccess ||
                            cudaEventCreate(&g->ev1[r]) != cudaSuccess))
            rc = wn_set_error(WN_ECUDA, "event creation failed on device %d", g->dev[r]);
    }
    if (rc == WN_OK && ngpus > 1) {
        if (!g->nccl.load())
            rc = wn_set_error(WN_ESTATE, "wn_group_create: NCCL (libnccl.so.2) could not be loaded: %s", dlerror());
        else {
            g->comm.assign(ngpus, nullptr);
            const int e = g->nccl.CommInitAll(g->comm.data(), ngpus, g->dev.data());
            if (e != 0) { rc = wn_set_error(WN_ECUDA, "ncclCommInitAll failed: %s", g->nccl.GetErrorString(e)); g->comm.clear(); }
        }
    }
    cudaSetDevice(prev);
    if (rc != WN_OK) { wn_group_destroy(g); return rc; }
    if (ngpus > 1)
        for (int r = 0; r < ngpus; ++r) {
            g->workers.emplace_back(new Worker());
            Worker *w = g->workers.back().get();
            const int device = g->dev[r];
            w->th = std::thread([w, device] { w->loop(device); });
        }
    *out = g;
    return WN_OK;
}

extern "C" int wn_group_destroy(wn_group *g)
{
    if (!g) return WN_OK;
    for (auto &w : g->workers) {
        { std::lock_guard<std::mutex> lk(w->m); w->stop = true; w->cv.notify_all(); }
        if (w->th.joinable()) w->th.join();
    }
    g->workers.clear();
    int prev = 0;
    cudaGetDevice(&prev);
    for (int r = 0; r < g->n; ++r) {
        if (!g->ctx[r]) continue;
        cudaSetDevice(g->dev[r]);
        cudaStreamSynchronize(wn_ctx_stream_internal(g->ctx[r]));
        if (r < (int)g->comm.size() && g->comm[r]) g->nccl.CommDestroy(g->comm[r]);
        cudaFree(g->out[r].p); cudaFree(g->in[r].p);
        if (g->ev0[r]) cudaEventDestroy(g->ev0[r]);
        if (g->ev1[r]) cudaEventDestroy(g->ev1[r]);
        wn_ctx_destroy(g->ctx[r]);
    }
    cudaSetDevice(prev);
    cudaGetLastError();
    delete g;
    return WN_OK;
}

extern "C" int wn_group_size(const wn_group *g) { return g ? g->n : 0; }

extern "C" int wn_group_ctx(wn_group *g, int rank, wn_ctx **ctx)
{
    WG_REQUIRE(g && ctx && rank >= 0 && rank < g->n, "wn_group_ctx: bad argument");
    *ctx = g->ctx[rank];
    return WN_OK;
}

extern "C" int wn_group_synchronize(wn_group *g)
{
    DeviceRestore keep;
    WG_REQUIRE(g, "wn_group_synchronize: group is NULL");
    for (int r = 0; r < g->n; ++r) WG_OK(wn_ctx_synchronize(g->ctx[r]));
    return WN_OK;
}

extern "C" int wn_group_shard(wn_group *g, int rank, float **dptr, size_t *count)
{
    WG_REQUIRE(g && rank >= 0 && rank < g->n, "wn_group_shard: bad argument");
    if (dptr) *dptr = g->out[rank].p;
    if (count) *count = g->out_count[rank];
    return WN_OK;
}

// ---- tiles ------------------------------------------------------------------------------------------------------------
extern "C" int wn_group_tile_create(wn_group *g, int n, int dims, unsigned flags, wn_gtile **out)
{
    WG_REQUIRE(g && out, "wn_group_tile_create: NULL argument");
    *out = nullptr;
    wn_gtile *t = new (std::nothrow) wn_gtile();
    if (!t) return wn_set_error(WN_ENOMEM, "out of host memory");
    t->g = g;
    t->t.assign(g->n, nullptr);
    for (int r = 0; r < g->n; ++r) {
        const int rc = wn_tile_create(g->ctx[r], n, dims, flags, &t->t[r]);
        if (rc != WN_OK) { wn_group_tile_destroy(t); return rc; }
    }
    wn_tile_info(t->t[0], nullptr, nullptr, &t->count, nullptr);
    *out = t;
    return WN_OK;
}

extern "C" int wn_group_tile_destroy(wn_gtile *t)
{
    if (!t) return WN_OK;
    for (wn_tile *x : t->t) wn_tile_destroy(x);
    delete t;
    return WN_OK;
}

extern "C" int wn_group_tile_rank(wn_gtile *t, int rank, wn_tile **tile)
{
    WG_REQUIRE(t && tile && rank >= 0 && rank < t->g->n, "wn_group_tile_rank: bad argument");
    *tile = t->t[rank];
    return WN_OK;
}

// rank 0 holds the finished coefficients: replicate them (one ncclBroadcast of n^dims floats over NVLink, enqueued on
// every rank's stream right behind the build) and refresh each replica's derived state
static int broadcast_tile(wn_gtile *t)
{
    wn_group *g = t->g;
    if (g->n == 1) return WN_OK;
    std::vector<void *> ptr(g->n);
    for (int r = 0; r < g->n; ++r) WG_OK(wn_tile_device_ptr(t->t[r], &ptr[r]));
    int e = g->nccl.GroupStart();
    for (int r = 0; r < g->n && e == 0; ++r) {
        WG_CUDA(cudaSetDevice(g->dev[r]));
        e = g->nccl.Broadcast(ptr[r], ptr[r], t->count, kNcclFloat, 0, g->comm[r], wn_ctx_stream_internal(g->ctx[r]));
    }
    const int e2 = g->nccl.GroupEnd();
    if (e != 0 || e2 != 0) return wn_set_error(WN_ECUDA, "ncclBroadcast of the tile failed: %s", g->nccl.GetErrorString(e ? e : e2));
    for (int r = 1; r < g->n; ++r) WG_OK(wn_tile_mark_built(t->t[r]));
    return WN_OK;
}

extern "C" int wn_group_tile_build_seeded(wn_gtile *t, unsigned seed, unsigned long long *mt_draws)
{
    DeviceRestore keep;
    WG_REQUIRE(t, "wn_group_tile_build_seeded: tile is NULL");
    WG_OK(wn_tile_build_seeded(t->t[0], seed, mt_draws));
    return broadcast_tile(t);
}

extern "C" int wn_group_tile_upload(wn_gtile *t, const float *N_host)
{
    DeviceRestore keep;
    WG_REQUIRE(t && N_host, "wn_group_tile_upload: NULL argument");
    WG_OK(wn_tile_upload(t->t[0], N_host, WN_HOST));
    return broadcast_tile(t);
}

// ---- sharded evaluation ---------------------------------------------------------------------------------------------
// Config 3: the volume is cut along z; every rank evaluates its slices as ONE lattice call on its own (sub-)axis.
extern "C" int wn_group_multiband3d_lattice(wn_gtile *t, const float *xs, int nx, const float *ys, int ny, const float *zs,
                                            int nz, const float *band_scale, const float *weights, int nbands, float post,
                                            int mode, int sharding, float *out_host, float *gpu_ms)
{
    DeviceRestore keep;
    WG_REQUIRE(t, "wn_group_multiband3d_lattice: tile is NULL");
    WG_REQUIRE(sharding == WN_SHARD_SLAB || sharding == WN_SHARD_CYCLIC, "bad sharding %d", sharding);
    WG_REQUIRE(nx >= 0 && ny >= 0 && nz >= 0 && (zs || !nz), "bad lattice");
    wn_group *g = t->g;
    const size_t slice = (size_t)nx * ny;
    std::vector<std::vector<int>> idx(g->n);
    std::vector<std::vector<float>> zr(g->n);
    for (int r = 0; r < g->n; ++r) {
        idx[r] = shard_indices(nz, r, g->n, sharding);
        zr[r].resize(idx[r].size());
        for (size_t k = 0; k < idx[r].size(); ++k) zr[r][k] = zs[idx[r][k]];
        WG_OK(reserve(g, g->out, r, slice * idx[r].size()));
        g->out_count[r] = slice * idx[r].size();
    }
    // every rank plans and enqueues its share on its own host thread: timing event, kernels, timing event, gather copies
    WG_OK(for_ranks(g, [&](int r) -> int {
        cudaStream_t st = wn_ctx_stream_internal(g->ctx[r]);
        WG_CUDA(cudaSetDevice(g->dev[r]));
        WG_CUDA(cudaEventRecord(g->ev0[r], st));
        if (!idx[r].empty())
            WG_OK(wn_multiband3d_lattice(t->t[r], xs, nx, ys, ny, zr[r].data(), (int)zr[r].size(), band_scale, weights, nbands,
                                         post, mode, g->out[r].p, WN_DEVICE));
        WG_CUDA(cudaEventRecord(g->ev1[r], st));
        if (out_host) {                                          // gather for file output: runs of consecutive slices
            size_t k = 0;
            while (k < idx[r].size()) {
                size_t e = k + 1;
                while (e < idx[r].size() && idx[r][e] == idx[r][e - 1] + 1) ++e;
                WG_CUDA(cudaMemcpyAsync(out_host + slice * (size_t)idx[r][k], g->out[r].p + slice * k,
                                        slice * (e - k) * sizeof(float), cudaMemcpyDeviceToHost, st));
                k = e;
            }
        }
        return WN_OK;
    }));
    if (!out_host && !gpu_ms) return WN_OK;                     // enqueue only: the caller synchronises the group later
    return finish(g, gpu_ms);
}

// Configs 4 / 5: image rows (the v axis) are cut into contiguous row-bands
extern "C" int wn_group_eval3d_projected_grid(wn_gtile *t, const float origin[3], const float e1[3], const float *us, int nu,
                                              const float e2[3], const float *vs, int nv, const float normal[3], float pre,
                                              float post, float *out_host, float *gpu_ms)
{
    DeviceRestore keep;
    WG_REQUIRE(t && (vs || !nv), "wn_group_eval3d_projected_grid: NULL argument");
    wn_group *g = t->g;
    std::vector<size_t> b(g->n), e(g->n);
    for (int r = 0; r < g->n; ++r) {
        slab_range((size_t)std::max(nv, 0), r, g->n, &b[r], &e[r]);
        WG_OK(reserve(g, g->out, r, (size_t)nu * (e[r] - b[r])));
        g->out_count[r] = (size_t)nu * (e[r] - b[r]);
    }
    WG_OK(for_ranks(g, [&](int r) -> int {
        cudaStream_t st = wn_ctx_stream_internal(g->ctx[r]);
        WG_CUDA(cudaSetDevice(g->dev[r]));
        WG_CUDA(cudaEventRecord(g->ev0[r], st));
        if (e[r] > b[r])
            WG_OK(wn_eval3d_projected_grid(t->t[r], origin, e1, us, nu, e2, vs + b[r], (int)(e[r] - b[r]), normal, pre, post,
                                           g->out[r].p, WN_DEVICE));
        WG_CUDA(cudaEventRecord(g->ev1[r], st));
        if (out_host && e[r] > b[r])
            WG_CUDA(cudaMemcpyAsync(out_host + (size_t)nu * b[r], g->out[r].p, g->out_count[r] * sizeof(float),
                                    cudaMemcpyDeviceToHost, st));
        return WN_OK;
    }));
    return finish(g, gpu_ms);
}

extern "C" int wn_group_perlin_grid(wn_group *g, wn_perlin *const *perlin_per_rank, const float origin[3], const float e1[3],
                                    const float *us, int nu, const float e2[3], const float *vs, int nv, float pre,
                                    float *out_host, float *gpu_ms)
{
    DeviceRestore keep;
    WG_REQUIRE(g && perlin_per_rank && (vs || !nv), "wn_group_perlin_grid: NULL argument");
    std::vector<size_t> b(g->n), e(g->n);
    for (int r = 0; r < g->n; ++r) {
        slab_range((size_t)std::max(nv, 0), r, g->n, &b[r], &e[r]);
        WG_OK(reserve(g, g->out, r, (size_t)nu * (e[r] - b[r])));
        g->out_count[r] = (size_t)nu * (e[r] - b[r]);
    }
    WG_OK(for_ranks(g, [&](int r) -> int {
        cudaStream_t st = wn_ctx_stream_internal(g->ctx[r]);
        WG_CUDA(cudaSetDevice(g->dev[r]));
        WG_CUDA(cudaEventRecord(g->ev0[r], st));
        if (e[r] > b[r])
            WG_OK(wn_perlin_grid(perlin_per_rank[r], origin, e1, us, nu, e2, vs + b[r], (int)(e[r] - b[r]), pre, g->out[r].p,
                                 WN_DEVICE));
        WG_CUDA(cudaEventRecord(g->ev1[r], st));
        if (out_host && e[r] > b[r])
            WG_CUDA(cudaMemcpyAsync(out_host + (size_t)nu * b[r], g->out[r].p, g->out_count[r] * sizeof(float),
                                    cudaMemcpyDeviceToHost, st));
        return WN_OK;
    }));
    return finish(g, gpu_ms);
}

// Config 5: the hit points of a batch (image row-bands in the renderer) are cut into contiguous runs, one per GPU; copies
// in, kernels and copies out of all ranks are in flight together.  p_host / grey_host should be pinned (wn_host_alloc)
// for the copies to overlap.  wait == 0: returns after enqueueing (call wn_group_synchronize before reading grey_host).
extern "C" int wn_group_wavelet_texture_values(wn_gtile *t, const float *p_host, size_t count, double scale, int octave,
                                               float *grey_host, int wait)
{
    DeviceRestore keep;
    WG_REQUIRE(t && (p_host || !count) && (grey_host || !count), "wn_group_wavelet_texture_values: NULL argument");
    wn_group *g = t->g;
    std::vector<size_t> b(g->n), e(g->n);
    for (int r = 0; r < g->n; ++r) {
        slab_range(count, r, g->n, &b[r], &e[r]);
        WG_OK(reserve(g, g->in, r, 3 * (e[r] - b[r])));
        WG_OK(reserve(g, g->out, r, e[r] - b[r]));
        g->out_count[r] = e[r] - b[r];
    }
    WG_OK(for_ranks(g, [&](int r) -> int {
        const size_t cnt = e[r] - b[r];
        cudaStream_t st = wn_ctx_stream_internal(g->ctx[r]);
        WG_CUDA(cudaSetDevice(g->dev[r]));
        WG_CUDA(cudaEventRecord(g->ev0[r], st));
        if (cnt) {
            WG_CUDA(cudaMemcpyAsync(g->in[r].p, p_host + 3 * b[r], 3 * cnt * sizeof(float), cudaMemcpyHostToDevice, st));
            WG_OK(wn_wavelet_texture_values(t->t[r], g->in[r].p, cnt, scale, octave, g->out[r].p, WN_DEVICE));
            WG_CUDA(cudaMemcpyAsync(grey_host + b[r], g->out[r].p, cnt * sizeof(float), cudaMemcpyDeviceToHost, st));
        }
        WG_CUDA(cudaEventRecord(g->ev1[r], st));
        return WN_OK;
    }));
    return wait ? finish(g, nullptr) : WN_OK;
}

extern "C" int wn_group_perlin_texture_values(wn_group *g, wn_perlin *const *perlin_per_rank, const float *p_host,
                                              size_t count, double scale, int octave, float *grey_host, int wait)
{
    DeviceRestore keep;
    WG_REQUIRE(g && perlin_per_rank && (p_host || !count) && (grey_host || !count), "wn_group_perlin_texture_values: NULL argument");
    std::vector<size_t> b(g->n), e(g->n);
    for (int r = 0; r < g->n; ++r) {
        slab_range(count, r, g->n, &b[r], &e[r]);
        WG_OK(reserve(g, g->in, r, 3 * (e[r] - b[r])));
        WG_OK(reserve(g, g->out, r, e[r] - b[r]));
        g->out_count[r] = e[r] - b[r];
    }
    WG_OK(for_ranks(g, [&](int r) -> int {
        const size_t cnt = e[r] - b[r];
        cudaStream_t st = wn_ctx_stream_internal(g->ctx[r]);
        WG_CUDA(cudaSetDevice(g->dev[r]));
        WG_CUDA(cudaEventRecord(g->ev0[r], st));
        if (cnt) {
            WG_CUDA(cudaMemcpyAsync(g->in[r].p, p_host + 3 * b[r], 3 * cnt * sizeof(float), cudaMemcpyHostToDevice, st));
            WG_OK(wn_perlin_texture_values(perlin_per_rank[r], g->in[r].p, cnt, scale, octave, g->out[r].p, WN_DEVICE));
            WG_CUDA(cudaMemcpyAsync(grey_host + b[r], g->out[r].p, cnt * sizeof(float), cudaMemcpyDeviceToHost, st));
        }
        WG_CUDA(cudaEventRecord(g->ev1[r], st));
        return WN_OK;
    }));
    return wait ? finish(g, nullptr) : WN_OK;
}
