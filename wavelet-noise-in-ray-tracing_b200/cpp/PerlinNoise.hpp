// PerlinNoise.hpp -- GPU-backed drop-in for the reference's header-only `class PerlinNoise`
// (reference: experient/PerlinNoise.hpp:9-61; the renderer's `class perlin`, perlin.h:14-72, is the same algorithm).
// The permutation table is built exactly like the reference's constructor (iota + std::shuffle with
// std::mt19937(seed), duplicated to 512 entries); noise() runs the double-precision, un-fused kernel in
// libwn_b200.so.  Scalar noise() calls are 1-point launches (correct but slow) through the double-in /
// double-out entry wn_perlin_points_f64; drivers should use the batch methods.  Coordinates are float-valued in
// every reference caller (experient/main.cpp:104,122; texture.h:39-40), so the BATCH entries take float32
// coordinates and promote them to double on the device.
#ifndef PERLINNOISE_HPP
#define PERLINNOISE_HPP

#include <cstdint>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/wn_b200.h"
#include "wn_batch.hpp"

class PerlinNoise {
  private:
    std::vector<int> p;                 // 512 entries, same contents as the reference's member
    wn_perlin* dev = nullptr;

  public:
    explicit PerlinNoise(unsigned int seed = std::mt19937::default_seed)
    {
        int32_t perm[512];
        wnb::check(wn_perlin_make_perm(seed, perm));
        p.assign(perm, perm + 512);
        wnb::check(wn_perlin_create(wnb::context(), perm, &dev));
    }
    ~PerlinNoise() { wn_perlin_destroy(dev); }
    PerlinNoise(const PerlinNoise& o) : p(o.p)
    {
        std::vector<int32_t> perm(p.begin(), p.end());
        wnb::check(wn_perlin_create(wnb::context(), perm.data(), &dev));
    }
    PerlinNoise& operator=(const PerlinNoise&) = delete;

    // double coordinates in, double noise out, like the reference (experient/PerlinNoise.hpp:36): nothing is narrowed
    double noise(double x, double y, double z) const
    {
        const double q[3] = {x, y, z};
        double out = 0.0;
        wnb::check(wn_perlin_points_f64(dev, q, 1, &out, WN_HOST));
        return out;
    }
    double noise(double x, double y) const { return noise(x, y, 0.0); }

    // image[(k*ny + j)*nx + i] = float(noise(xs[i], ys[j], zs[k]))
    std::vector<float> noise_lattice(const std::vector<float>& xs, const std::vector<float>& ys, const std::vector<float>& zs) const
    {
        std::vector<float> out(xs.size() * ys.size() * zs.size());
        wnb::check(wn_perlin_lattice(dev, xs.data(), (int)xs.size(), ys.data(), (int)ys.size(), zs.data(), (int)zs.size(),
                                     out.data(), WN_HOST));
        return out;
    }
    std::vector<float> noise_points(const float* xyz, size_t count, float pre = 1.0f) const
    {
        std::vector<float> out(count);
        wnb::check(wn_perlin_points(dev, xyz, count, pre, out.data(), WN_HOST));
        return out;
    }
    // noise_texture::value (texture.h:37-43) for a batch of hit points
    std::vector<float> texture_values(const float* xyz, size_t count, double scale, int octave) const
    {
        std::vector<float> out(count);
        wnb::check(wn_perlin_texture_values(dev, xyz, count, scale, octave, out.data(), WN_HOST));
        return out;
    }
    const std::vector<int>& permutation() const { return p; }
};

#endif
