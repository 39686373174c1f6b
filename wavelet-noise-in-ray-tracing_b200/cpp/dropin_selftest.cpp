// dropin_selftest.cpp -- exercises the reference-facing C++ surface (scalar methods, copies, stats, Perlin) and prints
// one "name value-bits" line per check; tests/test_cpp_dropin.py compares the lines with the CPU oracle.
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "PerlinNoise.hpp"
#include "WaveletNoise.h"
#include "wn_batch.hpp"

static void line(const char* name, float v)
{
    uint32_t b;
    std::memcpy(&b, &v, 4);
    std::printf("%s %08x\n", name, b);
}

int main()
{
    try {
        const float p[3] = {3.25f, -7.5f, 100.125f};
        const float nrm[3] = {0.6f, 0.0f, 0.8f};
        WaveletNoise n3(127, 4242);                  // odd size -> 128 with the reference's warning on stderr
        line("empty_eval3d", n3.evaluate3D(p));      // no tile yet -> 0.0f
        n3.generateNoiseTile3D();
        std::printf("tile_size %d\n", n3.getTileSize());
        line("eval3d", n3.evaluate3D(p));
        line("proj", n3.evaluate3DProjected(p, nrm));
        line("coeff0", n3.getNoiseCoefficients()[0]);
        line("coeff_last", n3.getNoiseCoefficients().back());
        WaveletNoise copy = n3;                      // implicit copy: must evaluate identically (lazy re-upload)
        line("copy_eval3d", copy.evaluate3D(p));
        // copy ASSIGNMENT between two generated objects of the same size reuses the target's vector storage (same
        // pointer, same size): the device tile must still follow the new coefficients (content fingerprint)
        WaveletNoise a(16, 1), b(16, 2);
        a.generateNoiseTile3D();
        b.generateNoiseTile3D();
        line("assign_before", a.evaluate3D(p));
        a = b;
        line("assign_after", a.evaluate3D(p));
        line("assign_source", b.evaluate3D(p));
        line("assign_coeff0", a.getNoiseCoefficients()[0]);
        WaveletNoise n2(16, 7);
        n2.generateNoiseTile2D();
        n2.generateNoiseTile2D();                    // second call continues the RNG stream
        line("eval2d_second", n2.evaluate2D(p));
        DataStats st = n3.calculateStats(n3.getNoiseCoefficients(), "tile3D");
        line("stats_min", st.min_val);
        line("stats_max", st.max_val);
        PerlinNoise perlin(12345);
        line("perlin", (float)perlin.noise(3.25, -7.5, 100.125));
        line("perlin2d", (float)perlin.noise(0.3, 0.7));
        std::printf("perm0 %d\n", perlin.permutation()[0]);
        const float pts[6] = {0.5f, 1.5f, -2.25f, 9.0f, 9.5f, 10.25f};
        std::vector<float> tv = wnb::wavelet_texture_values(n3, pts, 2, 1.0, 4);
        line("tex0", tv[0]);
        line("tex1", tv[1]);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
