// render_deferred.cpp -- the reference's renderer (reference: main.cpp:80-215) with the noise-sampling path on the GPU.
//
// BASELINE config 5.  The reference evaluates its noise texture once per diffuse hit, inside the recursive
// trace() (main.cpp:38-59 -> material.h:63-74 -> texture.h:67-107 / :37-43): ~0.59 scalar calls per primary ray.
// The texture value only scales colour -- it never changes a ray path or the rand() stream (material.h:65-72) --
// so tracing and texturing can be separated without changing a single pixel:
//   pass 1 (CPU, unchanged reference geometry code): for a band of image rows, trace every sample in the reference's
//           loop order; a recording texture notes each hit point instead of evaluating the noise, and each sample keeps
//           its chain of attenuation factors (constant colour | lookup #i) and its terminal radiance;
//   pass 2 (GPUs): ONE batched call per band evaluates all recorded lookups
//           (wn_group_wavelet_texture_values / wn_group_perlin_texture_values = texture.h:67-107 / :37-43, bit-exact);
//           with --gpus G the band's points are cut into G contiguous runs, one per GPU, all in flight together;
//   pass 3 (CPU): each sample's product is re-formed right to left exactly like `attenuation * trace(...)`
//           (main.cpp:55), summed per pixel in sample order and quantised with the reference's expression.
// The passes are software-pipelined over the bands: the GPUs shade band k (from pinned staging, enqueue-only call)
// while the CPU recombines band k-1 and traces band k+1, so the noise path costs no wall time next to the tracer.
// The PNG/PPM are byte-identical to the reference's (tests compare with result_raytracing/*.png).
//
// This file contains no reference code: the RTIOW substrate (vec3/ray/sphere/quad/material ..., CC0) and
// stb_image_write.h are #included from the reference tree where it lies (-I$(REF)) at build time, like oracle/_ref.
// Protocol kept from the reference: noise type and octave are read from stdin (main.cpp:85-109).
// Extra argv: --width W --height H --spp S --band ROWS --out DIR --gpus G (one device group of G GPUs).
#define STB_IMAGE_WRITE_IMPLEMENTATION
#include "stb_image_write.h"

#include "rtweekend.h"

#include <cfloat>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "hittable.h"
#include "hittable_list.h"
#include "material.h"      // pulls in the reference's texture.h (class texture, the hook)
#include "quad.h"
#include "sphere.h"

#include "../../include/wn_b200.h"

namespace {

const int MAX_DEPTH = 10;                       // main.cpp:30

struct Factor {                                 // one `attenuation` of the chain
    int lookup;                                 // >= 0: index into the band's lookup batch; -1: constant colour
    color constant;
};

// ---- the recording texture: same hook signature, defers the evaluation ------------------------------------------
struct Recorder {
    std::vector<float> points;                  // xyz per lookup (vec3 stores float)
    int last = -1;                              // index of the lookup made by the most recent value() call
    void clear() { points.clear(); last = -1; }
};
Recorder g_rec;

class deferred_noise_texture : public texture {
  public:
    color value(double, double, const point3& p) const override
    {
        g_rec.last = (int)(g_rec.points.size() / 3);
        g_rec.points.push_back(p.x());
        g_rec.points.push_back(p.y());
        g_rec.points.push_back(p.z());
        return color(1, 1, 1);                  // placeholder; the real grey value arrives in pass 2
    }
};

inline double clamp01(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }   // main.cpp:24-28

void die(int rc)
{
    if (rc != WN_OK) {
        std::cerr << "error: " << wn_last_error() << std::endl;
        std::exit(1);
    }
}

// pinned host staging that grows on demand (the lookups of a band go to the GPUs from here, the grey values come back)
struct Pinned {
    float* p = nullptr;
    size_t cap = 0;
    void ensure(size_t n)
    {
        if (n <= cap) return;
        if (p) wn_host_free(p);
        cap = n + n / 4 + 1024;
        die(wn_host_alloc(cap * sizeof(float), (void**)&p));
    }
    ~Pinned() { if (p) wn_host_free(p); }
};

// everything pass 3 needs about one band of rows
struct Band {
    int jtop = 0, jbot = 0;
    std::vector<Factor> factors;                // all chains of the band, concatenated
    std::vector<int> chain_begin;               // per sample: first factor
    std::vector<color> terminal;                // per sample: innermost radiance
    Pinned points, grey;
    size_t nlook = 0;
};

}  // namespace

int main(int argc, char** argv)
{
    int width = 1000, height = 500, spp = 100, band_rows = 10, ngpus = 1;     // main.cpp:119-121
    std::string out_dir = "result_raytracing";
    for (int a = 1; a + 1 < argc; a += 2) {
        const std::string k = argv[a];
        if (k == "--width") width = std::atoi(argv[a + 1]);
        else if (k == "--height") height = std::atoi(argv[a + 1]);
        else if (k == "--spp") spp = std::atoi(argv[a + 1]);
        else if (k == "--band") band_rows = std::atoi(argv[a + 1]);
        else if (k == "--out") out_dir = argv[a + 1];
        else if (k == "--gpus") ngpus = std::atoi(argv[a + 1]);
    }

    // ---- stdin protocol of the reference (main.cpp:85-109): out-of-range / unreadable -> defaults
    std::cout << "\n=== Wavelet Noise in Ray Tracing (deferred texturing on GPU) ===" << std::endl;
    int noise_choice = 1;
    if (!(std::cin >> noise_choice)) { noise_choice = 1; std::cin.clear(); std::cin.ignore(10000, '\n'); }
    if (noise_choice < 0 || noise_choice > 1) noise_choice = 1;
    int octave_level = 4;
    if (!(std::cin >> octave_level)) { octave_level = 4; std::cin.clear(); std::cin.ignore(10000, '\n'); }
    if (octave_level < 3 || octave_level > 5) octave_level = 4;
    const bool wavelet = noise_choice == 1;
    const std::string noise_name = (wavelet ? "Wavelet3D_octave" : "Perlin_octave") + std::to_string(octave_level);
    const double tex_scale = 1.0;               // create_noise_texture(selected_noise, 1.0, octave_level), main.cpp:144

    // ---- GPU side: the texture's noise objects (texture.h:53-65 builds n=128 seed 12345; texture.h:46 default-seeds perlin)
    wn_group* group = nullptr;
    die(wn_group_create(ngpus, nullptr, &group));
    ngpus = wn_group_size(group);
    wn_gtile* gtile = nullptr;
    std::vector<wn_perlin*> perlins(ngpus, nullptr);
    if (wavelet) {
        die(wn_group_tile_create(group, 128, 3, WN_TILE_DEFAULT, &gtile));
        die(wn_group_tile_build_seeded(gtile, 12345, nullptr));      // rank 0 builds, one NCCL broadcast replicates
    } else {
        int32_t perm[512];
        die(wn_perlin_make_perm(5489u, perm));                       // std::mt19937::default_seed
        for (int g = 0; g < ngpus; ++g) {
            wn_ctx* c = nullptr;
            die(wn_group_ctx(group, g, &c));
            die(wn_perlin_create(c, perm, &perlins[g]));
        }
    }

    // ---- scene, exactly main.cpp:126-163 (only the noise texture object differs: it records instead of evaluating)
    vec3 lower_left_corner(-2, -1, -1);
    vec3 origin(0, 0, 1);
    vec3 horizontal(4, 0, 0);
    vec3 vertical(0, 2, 0);
    auto light_material = make_shared<diffuse_light>(color(4.0, 4.0, 4.0));
    auto noise_material = make_shared<lambertian>(make_shared<deferred_noise_texture>());
    hittable_list world;
    world.add(make_shared<quad>(point3(-10, -0.5, -10), vec3(20, 0, 0), vec3(0, 0, 20), noise_material));
    world.add(make_shared<sphere>(vec3(-5, 5, 0), 0.8, light_material));
    world.add(make_shared<sphere>(vec3(1, 0, -1.75), 0.5, noise_material));

    std::vector<unsigned char> image((size_t)width * height * 3);
    std::string cmd = "mkdir -p " + out_dir;
    if (std::system(cmd.c_str()) != 0) std::cerr << "warning: could not create " << out_dir << std::endl;
    const std::string output_ppm = out_dir + "/raytrace_" + noise_name + ".ppm";
    const std::string output_png = out_dir + "/raytrace_" + noise_name + ".png";
    std::ofstream file(output_ppm);
    file << "P3\n" << width << " " << height << "\n255\n";
    std::cout << "Processing " << width << "x" << height << " @ " << spp << " spp, " << noise_name << std::endl;

    Band bands[2];                              // double buffer: the GPUs work on one band while the CPU fills the other
    size_t total_lookups = 0;
    double t_trace = 0, t_wait = 0, t_combine = 0;
    const auto t_start = std::chrono::steady_clock::now();

    // pass 3 of a band whose grey values have arrived
    auto combine = [&](Band& B) {
        size_t sample = 0;
        const float* grey = B.grey.p;
        for (int j = B.jtop; j >= B.jbot; --j)
            for (int i = 0; i < width; ++i) {
                vec3 color_sum(0, 0, 0);
                for (int s = 0; s < spp; ++s, ++sample) {
                    color c = B.terminal[sample];
                    for (int f = B.chain_begin[sample + 1] - 1; f >= B.chain_begin[sample]; --f) {
                        const Factor& fa = B.factors[f];
                        const color att = fa.lookup >= 0 ? color(grey[fa.lookup], grey[fa.lookup], grey[fa.lookup]) : fa.constant;
                        c = att * c;
                    }
                    color_sum += c;
                }
                vec3 c = color_sum / float(spp);
                int r = static_cast<int>(255.99 * clamp01(c.x(), 0.0f, 1.0f));
                int g = static_cast<int>(255.99 * clamp01(c.y(), 0.0f, 1.0f));
                int b = static_cast<int>(255.99 * clamp01(c.z(), 0.0f, 1.0f));
                file << r << " " << g << " " << b << "\n";
                int index = ((height - 1 - j) * width + i) * 3;
                image[index + 0] = r; image[index + 1] = g; image[index + 2] = b;
            }
    };

    int band_index = 0;
    Band* in_flight = nullptr;                  // band whose lookups are on the GPUs
    for (int jtop = height - 1; jtop >= 0; jtop -= band_rows, ++band_index) {
        Band& B = bands[band_index & 1];
        B.jtop = jtop;
        B.jbot = std::max(jtop - band_rows + 1, 0);
        auto c0 = std::chrono::steady_clock::now();
        // ---------------- pass 1: trace in the reference's loop order (rows top -> bottom, x, samples), main.cpp:175-191
        g_rec.clear();
        B.factors.clear(); B.chain_begin.clear(); B.terminal.clear();
        for (int j = B.jtop; j >= B.jbot; --j)
            for (int i = 0; i < width; ++i)
                for (int s = 0; s < spp; ++s) {
                    float rand_u = float(rand()) / RAND_MAX - 0.5f;
                    float rand_v = float(rand()) / RAND_MAX - 0.5f;
                    float u = float(i + 0.5f + rand_u) / width;
                    float v = float(j + 0.5f + rand_v) / height;
                    ray r(origin, unit_vector(lower_left_corner + u * horizontal + v * vertical - origin));
                    B.chain_begin.push_back((int)B.factors.size());
                    // iterative form of trace() (main.cpp:38-59); same calls in the same order
                    color last(0, 0, 0);
                    for (int step = 0;; ++step) {
                        if (step > MAX_DEPTH) { last = vec3(0, 0, 0); break; }
                        hit_record rec;
                        if (!world.hit(r, 0.001f, FLT_MAX, rec)) {
                            vec3 unit_direction = unit_vector(r.direction());
                            float t = 0.5f * (unit_direction.y() + 1.0f);
                            last = (1.0f - t) * vec3(1, 1, 1) + t * vec3(0.40, 0.50, 1.00);
                            break;
                        }
                        color attenuation;
                        ray scattered;
                        g_rec.last = -1;
                        if (rec.mat->scatter(r, rec, attenuation, scattered)) {
                            B.factors.push_back(Factor{g_rec.last, attenuation});
                            r = scattered;
                            continue;
                        }
                        last = rec.mat->emitted(rec.u, rec.v, rec.p);
                        break;
                    }
                    B.terminal.push_back(last);
                }
        B.chain_begin.push_back((int)B.factors.size());
        B.nlook = g_rec.points.size() / 3;
        B.points.ensure(3 * B.nlook);
        B.grey.ensure(B.nlook);
        if (B.nlook) std::memcpy(B.points.p, g_rec.points.data(), 3 * B.nlook * sizeof(float));
        total_lookups += B.nlook;
        auto c1 = std::chrono::steady_clock::now();

        // ---------------- pass 2: the previous band's batch ran during this trace; wait for it (normally already done),
        // hand this band to the GPUs (enqueue only), then recombine the previous band while they work
        die(wn_group_synchronize(group));
        auto c2 = std::chrono::steady_clock::now();
        if (B.nlook) {
            if (wavelet) die(wn_group_wavelet_texture_values(gtile, B.points.p, B.nlook, tex_scale, octave_level, B.grey.p, 0));
            else die(wn_group_perlin_texture_values(group, perlins.data(), B.points.p, B.nlook, tex_scale, octave_level, B.grey.p, 0));
        }
        // ---------------- pass 3 (previous band): products right to left, accumulate, quantise (main.cpp:55, 190-202)
        if (in_flight) combine(*in_flight);
        in_flight = &B;
        auto c3 = std::chrono::steady_clock::now();
        t_trace += std::chrono::duration<double>(c1 - c0).count();
        t_wait += std::chrono::duration<double>(c2 - c1).count();
        t_combine += std::chrono::duration<double>(c3 - c2).count();
    }
    {
        auto c1 = std::chrono::steady_clock::now();
        die(wn_group_synchronize(group));
        auto c2 = std::chrono::steady_clock::now();
        if (in_flight) combine(*in_flight);
        t_wait += std::chrono::duration<double>(c2 - c1).count();
        t_combine += std::chrono::duration<double>(std::chrono::steady_clock::now() - c2).count();
    }
    const double t_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
    stbi_write_png(output_png.c_str(), width, height, 3, image.data(), width * 3);

    std::printf("texture lookups: %zu (%.3f per primary ray)\n", total_lookups, (double)total_lookups / ((double)width * height * spp));
    std::printf("time: total %.3f s = trace(CPU) %.3f s + recombine(CPU) %.3f s + waiting for the GPUs %.3f s (%d GPU%s, "
                "noise batches overlap the tracing of the next band)\n", t_total, t_trace, t_combine, t_wait, ngpus, ngpus == 1 ? "" : "s");
    if (t_wait > 0) std::printf("noise-sampling path: %.1f Mlookups/s of exposed GPU time\n", total_lookups / t_wait / 1e6);
    std::cout << "- PPM: " << output_ppm << "\n- PNG: " << output_png << std::endl;
    for (wn_perlin* p : perlins) wn_perlin_destroy(p);
    wn_group_tile_destroy(gtile);
    wn_group_destroy(group);
    return 0;
}
