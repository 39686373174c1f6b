// sharded_main.cpp -- single-process multi-GPU drivers for BASELINE configs 3 and 4 over the C ABI's device groups.
//
//   sharded_b200 volume [--gpus N] [--size 1024] [--bands 4 8] [--tile 128] [--seed 12345] [--sharding cyclic|slab]
//                       [--reps 10] [--gather] [--out volume.raw]
//       config 3: WMultibandNoise on a size^3 lattice, p = (idx/size)*4, q_b = 2 p 2^b, w_b = 2^-(b-first),
//       post = 1/sqrt(sum w^2 * 0.18402) -- the loop a caller would write over WaveletNoise::evaluate3D
//       (reference WaveletNoise.cpp:185-215, coordinates as experient/main.cpp:45-58), sharded along z.
//   sharded_b200 plane  [--gpus N] [--size 8192] [--tile 128] [--seed 12345] [--reps 5] [--out plane]
//       config 4: evaluate3DProjected on an oblique plane (normal (1,2,3)/sqrt14) vs Perlin(12345) octave 4 on the same
//       grid (reference experient/main.cpp:74-87, :113-129), sharded into row-bands.
// Prints one JSON line: Gsamples/s from the max-over-GPUs CUDA-event time, wall time per call, and an FNV-1a hash of the
// gathered output (the same for every GPU count and sharding: samples are independent and summed canonically).
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/wn_b200.h"

static void die(const char *what)
{
    std::fprintf(stderr, "sharded_b200: %s: %s\n", what, wn_last_error());
    std::exit(1);
}
#define CHECK(call) do { if ((call) != WN_OK) die(#call); } while (0)

static uint64_t fnv1a(const float *p, size_t n)
{
    uint64_t h = 1469598103934665603ull;
    const unsigned char *b = reinterpret_cast<const unsigned char *>(p);
    for (size_t i = 0; i < n * sizeof(float); ++i) h = (h ^ b[i]) * 1099511628211ull;
    return h;
}

struct Args {
    std::string mode = "volume", sharding = "cyclic", out;
    int gpus = 0, size = -1, b0 = 4, b1 = 8, tile = 128, reps = -1;
    unsigned seed = 12345;
    bool gather = false;
};

static Args parse(int argc, char **argv)
{
    Args a;
    int i = 1;
    if (i < argc && argv[i][0] != '-') a.mode = argv[i++];
    for (; i < argc; ++i) {
        const std::string k = argv[i];
        auto next = [&]() -> const char * { if (i + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", k.c_str()); std::exit(2); } return argv[++i]; };
        if (k == "--gpus") a.gpus = std::atoi(next());
        else if (k == "--size") a.size = std::atoi(next());
        else if (k == "--bands") { a.b0 = std::atoi(next()); a.b1 = std::atoi(next()); }
        else if (k == "--tile") a.tile = std::atoi(next());
        else if (k == "--seed") a.seed = (unsigned)std::strtoul(next(), nullptr, 10);
        else if (k == "--sharding") a.sharding = next();
        else if (k == "--reps") a.reps = std::atoi(next());
        else if (k == "--out") a.out = next();
        else if (k == "--gather") a.gather = true;
        else { std::fprintf(stderr, "unknown option %s\n", k.c_str()); std::exit(2); }
    }
    return a;
}

static int run_volume(const Args &a)
{
    const int S = a.size > 0 ? a.size : 1024, reps = a.reps > 0 ? a.reps : 10;
    wn_group *g = nullptr;
    CHECK(wn_group_create(a.gpus, nullptr, &g));
    const int N = wn_group_size(g);
    wn_gtile *t = nullptr;
    CHECK(wn_group_tile_create(g, a.tile, 3, WN_TILE_DEFAULT, &t));
    auto t0 = std::chrono::steady_clock::now();
    CHECK(wn_group_tile_build_seeded(t, a.seed, nullptr));
    CHECK(wn_group_synchronize(g));
    const double tile_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    std::vector<float> ax(S), scale, w;
    for (int i = 0; i < S; ++i) ax[i] = ((float)i / (float)S) * 4.0f;          // experient/main.cpp:20-21
    float sw2 = 0.0f;
    for (int b = a.b0; b <= a.b1; ++b) {
        scale.push_back(2.0f * std::pow(2.0f, (float)b));
        w.push_back(std::pow(2.0f, -(float)(b - a.b0)));
        sw2 += w.back() * w.back();
    }
    const float post = 1.0f / std::sqrt(sw2 * 0.18402f);
    const int sharding = a.sharding == "slab" ? WN_SHARD_SLAB : WN_SHARD_CYCLIC;
    const size_t total = (size_t)S * S * S;
    const bool gather = a.gather || !a.out.empty();
    float *host = nullptr;
    if (gather) CHECK(wn_host_alloc(total * sizeof(float), (void **)&host));
    float gpu_ms = 0.0f, best_ms = 1e30f, sum_ms = 0.0f;
    for (int r = 0; r < 3; ++r)                                                // warm-up
        CHECK(wn_group_multiband3d_lattice(t, ax.data(), S, ax.data(), S, ax.data(), S, scale.data(), w.data(), (int)w.size(),
                                           post, WN_EVAL_FAST, sharding, nullptr, &gpu_ms));
    t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; ++r) {
        CHECK(wn_group_multiband3d_lattice(t, ax.data(), S, ax.data(), S, ax.data(), S, scale.data(), w.data(), (int)w.size(),
                                           post, WN_EVAL_FAST, sharding, nullptr, &gpu_ms));
        best_ms = std::min(best_ms, gpu_ms);
        sum_ms += gpu_ms;
    }
    const double wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / reps;
    // throughput: the same calls enqueued back to back without waiting in between (the shards stay on the devices)
    CHECK(wn_group_synchronize(g));
    t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < 4 * reps; ++r)
        CHECK(wn_group_multiband3d_lattice(t, ax.data(), S, ax.data(), S, ax.data(), S, scale.data(), w.data(), (int)w.size(),
                                           post, WN_EVAL_FAST, sharding, nullptr, nullptr));
    CHECK(wn_group_synchronize(g));
    const double stream_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / (4 * reps);
    double gather_ms = 0.0;
    uint64_t hash = 0;
    if (gather) {
        t0 = std::chrono::steady_clock::now();
        CHECK(wn_group_multiband3d_lattice(t, ax.data(), S, ax.data(), S, ax.data(), S, scale.data(), w.data(), (int)w.size(),
                                           post, WN_EVAL_FAST, sharding, host, &gpu_ms));
        gather_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        hash = fnv1a(host, total);
        if (!a.out.empty()) {
            FILE *f = std::fopen(a.out.c_str(), "wb");
            if (!f || std::fwrite(host, sizeof(float), total, f) != total) { std::fprintf(stderr, "cannot write %s\n", a.out.c_str()); return 1; }
            std::fclose(f);
        }
    }
    std::printf("{\"driver\": \"sharded_b200 volume\", \"config\": \"WMultibandNoise %d^3 bands %d-%d tile n=%d seed %u\", "
                "\"n_gpus\": %d, \"sharding\": \"%s\", \"reps\": %d, \"gpu_ms_mean\": %.4f, \"gpu_ms_best\": %.4f, "
                "\"wall_ms_per_call\": %.4f, \"gsamples_s\": %.2f, \"gsamples_s_wall\": %.2f, \"back_to_back_ms_per_call\": %.4f, "
                "\"gsamples_s_back_to_back\": %.2f, \"tile_build_and_broadcast_ms\": %.3f, "
                "\"gather_ms\": %.2f, \"fnv1a64\": \"%016llx\"}\n",
                S, a.b0, a.b1, a.tile, a.seed, N, a.sharding.c_str(), reps, sum_ms / reps, best_ms, wall_ms,
                (double)total / (sum_ms / reps) / 1e6, (double)total / wall_ms / 1e6, stream_ms, (double)total / stream_ms / 1e6,
                tile_ms, gather_ms,
                (unsigned long long)hash);
    if (host) wn_host_free(host);
    wn_group_tile_destroy(t);
    wn_group_destroy(g);
    return 0;
}

static int run_plane(const Args &a)
{
    const int S = a.size > 0 ? a.size : 8192, reps = a.reps > 0 ? a.reps : 5;
    wn_group *g = nullptr;
    CHECK(wn_group_create(a.gpus, nullptr, &g));
    const int N = wn_group_size(g);
    wn_gtile *t = nullptr;
    CHECK(wn_group_tile_create(g, a.tile, 3, WN_TILE_DEFAULT, &t));
    CHECK(wn_group_tile_build_seeded(t, a.seed, nullptr));
    int32_t perm[512];
    CHECK(wn_perlin_make_perm(12345u, perm));
    std::vector<wn_perlin *> pn(N, nullptr);
    for (int r = 0; r < N; ++r) {
        wn_ctx *c = nullptr;
        CHECK(wn_group_ctx(g, r, &c));
        CHECK(wn_perlin_create(c, perm, &pn[r]));
    }
    // the plane of BASELINE config 4: vectors rounded to float once, coordinates formed un-fused on the device
    const float nrm[3] = { (float)(1.0 / std::sqrt(14.0)), (float)(2.0 / std::sqrt(14.0)), (float)(3.0 / std::sqrt(14.0)) };
    const float e1[3] = { (float)(2.0 / std::sqrt(5.0)), (float)(-1.0 / std::sqrt(5.0)), 0.0f };
    const float e2[3] = { (float)(3.0 / std::sqrt(70.0)), (float)(6.0 / std::sqrt(70.0)), (float)(-5.0 / std::sqrt(70.0)) };
    const float origin[3] = { 0.0f, 0.0f, 1.0f };
    std::vector<float> ax(S);
    for (int i = 0; i < S; ++i) ax[i] = ((float)i / (float)S) * 4.0f;
    const float pre_w = 2.0f * 16.0f, pre_p = 16.0f, inv = 1.0f / std::sqrt(0.296f);     // experient/main.cpp:72
    const size_t total = (size_t)S * S;
    float *hproj = nullptr, *hperl = nullptr;
    CHECK(wn_host_alloc(total * sizeof(float), (void **)&hproj));
    CHECK(wn_host_alloc(total * sizeof(float), (void **)&hperl));
    float ms = 0.0f, proj_ms = 0.0f, perl_ms = 0.0f;
    CHECK(wn_group_eval3d_projected_grid(t, origin, e1, ax.data(), S, e2, ax.data(), S, nrm, pre_w, inv, nullptr, &ms));
    for (int r = 0; r < reps; ++r) {
        CHECK(wn_group_eval3d_projected_grid(t, origin, e1, ax.data(), S, e2, ax.data(), S, nrm, pre_w, inv, nullptr, &ms));
        proj_ms += ms / reps;
    }
    CHECK(wn_group_eval3d_projected_grid(t, origin, e1, ax.data(), S, e2, ax.data(), S, nrm, pre_w, inv, hproj, &ms));
    CHECK(wn_group_perlin_grid(g, pn.data(), origin, e1, ax.data(), S, e2, ax.data(), S, pre_p, nullptr, &ms));
    for (int r = 0; r < reps; ++r) {
        CHECK(wn_group_perlin_grid(g, pn.data(), origin, e1, ax.data(), S, e2, ax.data(), S, pre_p, nullptr, &ms));
        perl_ms += ms / reps;
    }
    CHECK(wn_group_perlin_grid(g, pn.data(), origin, e1, ax.data(), S, e2, ax.data(), S, pre_p, hperl, &ms));
    if (!a.out.empty())
        for (int k = 0; k < 2; ++k) {
            const std::string name = a.out + (k ? "_perlin.raw" : "_projected.raw");
            FILE *f = std::fopen(name.c_str(), "wb");
            if (!f || std::fwrite(k ? hperl : hproj, sizeof(float), total, f) != total) { std::fprintf(stderr, "cannot write %s\n", name.c_str()); return 1; }
            std::fclose(f);
        }
    std::printf("{\"driver\": \"sharded_b200 plane\", \"config\": \"WProjectedNoise vs Perlin octave 4, %dx%d plane, normal (1,2,3)/sqrt14\", "
                "\"n_gpus\": %d, \"sharding\": \"row-band\", \"reps\": %d, \"projected_gpu_ms\": %.4f, \"projected_gsamples_s\": %.3f, "
                "\"perlin_gpu_ms\": %.4f, \"perlin_gsamples_s\": %.2f, \"fnv1a64_projected\": \"%016llx\", \"fnv1a64_perlin\": \"%016llx\"}\n",
                S, S, N, reps, proj_ms, (double)total / proj_ms / 1e6, perl_ms, (double)total / perl_ms / 1e6,
                (unsigned long long)fnv1a(hproj, total), (unsigned long long)fnv1a(hperl, total));
    for (wn_perlin *p : pn) wn_perlin_destroy(p);
    wn_host_free(hproj); wn_host_free(hperl);
    wn_group_tile_destroy(t);
    wn_group_destroy(g);
    return 0;
}

int main(int argc, char **argv)
{
    const Args a = parse(argc, argv);
    if (a.mode == "volume") return run_volume(a);
    if (a.mode == "plane") return run_plane(a);
    std::fprintf(stderr, "usage: sharded_b200 volume|plane [options] (see the header of sharded_main.cpp)\n");
    return 2;
}
