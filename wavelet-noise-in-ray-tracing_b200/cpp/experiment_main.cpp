// experiment_main.cpp -- batch sibling of the reference's experiment driver (reference: experient/main.cpp:11-168).
// Same tiles (n=128, seed 12345), same 15 output files with the same names, byte layout (float32 row-major
// image[y*256+x]) and stdout lines; but each image is ONE batched GPU call instead of 65 536 scalar
// evaluate*() calls.  The unmodified reference driver also links against this directory's WaveletNoise.cpp /
// PerlinNoise.hpp (scalar path); this file is the throughput path.
#include <sys/stat.h>

#include <cmath>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "PerlinNoise.hpp"
#include "WaveletNoise.h"
#include "wn_batch.hpp"

namespace {

const float kBaseRange = 4.0f;

// u = (float(x) / imageSize) * base_range for every pixel -- the reference's own expression (main.cpp:20-21)
std::vector<float> pixel_axis(int imageSize, float scale)
{
    std::vector<float> u(imageSize);
    for (int x = 0; x < imageSize; ++x) u[x] = ((static_cast<float>(x) / imageSize) * kBaseRange) * scale;
    return u;
}

void write_raw(const std::string& file, const std::vector<float>& image, const std::string& what, int octave)
{
    std::ofstream out(file, std::ios::binary);
    out.write(reinterpret_cast<const char*>(image.data()), image.size() * sizeof(float));
    out.close();
    std::cout << "Generated " << what << " Octave " << octave << " noise: " << file << std::endl;
}

}  // namespace

void generate2DOctaveBandNoise(int imageSize, int octave, const std::string& outputFile, WaveletNoise& noise)
{
    const float octave_scale = std::pow(2.0f, octave);
    const float inv_stddev_2d = 1.0f / std::sqrt(0.19686f);
    // p = (u*octave_scale)*2 == u * (octave_scale*2): one exact power-of-two factor
    const std::vector<float> u = pixel_axis(imageSize, 1.0f);
    write_raw(outputFile, wnb::evaluate2D_lattice(noise, u, u, octave_scale * 2.0f, inv_stddev_2d), "Wavelet 2D", octave);
}

void generate3DSlicedOctaveBandNoise(int imageSize, int octave, const std::string& outputFile, WaveletNoise& noise)
{
    const float octave_scale = std::pow(2.0f, octave);
    const float inv_stddev_3d = 1.0f / std::sqrt(0.18402f);
    const std::vector<float> xy = pixel_axis(imageSize, octave_scale * 2.0f);
    const std::vector<float> z = {1.0f * 2.0f};                       // p[2] = 1.0f; p[2] *= 2.0f (main.cpp:50-54)
    write_raw(outputFile, wnb::multiband3D_lattice(noise, xy, xy, z, {1.0f}, {1.0f}, inv_stddev_3d, WN_EVAL_EXACT),
              "Wavelet 3D Sliced", octave);
}

void generate3DProjectedOctaveBandNoise(int imageSize, int octave, const std::string& outputFile, WaveletNoise& noise)
{
    const float octave_scale = std::pow(2.0f, octave);
    const float normal[3] = {0.0f, 0.0f, 1.0f};
    const float inv_stddev_3d_proj = 1.0f / std::sqrt(0.296f);
    const std::vector<float> xy = pixel_axis(imageSize, octave_scale * 2.0f);
    const float origin[3] = {0.0f, 0.0f, 2.0f}, ex[3] = {1.0f, 0.0f, 0.0f}, ey[3] = {0.0f, 1.0f, 0.0f};
    write_raw(outputFile, wnb::evaluate3DProjected_grid(noise, origin, ex, xy, ey, xy, normal, 1.0f, inv_stddev_3d_proj),
              "Wavelet 3D Projected", octave);
}

void generatePerlinNoise2D(int imageSize, int octave, const std::string& outputFile, const PerlinNoise& perlin)
{
    const std::vector<float> u = pixel_axis(imageSize, std::pow(2.0f, octave));
    write_raw(outputFile, perlin.noise_lattice(u, u, {0.0f}), "Perlin 2D", octave);
}

void generatePerlinNoise3DSliced(int imageSize, int octave, const std::string& outputFile, const PerlinNoise& perlin)
{
    const float octave_scale = std::pow(2.0f, octave);
    const std::vector<float> u = pixel_axis(imageSize, octave_scale);
    write_raw(outputFile, perlin.noise_lattice(u, u, {1.0f * octave_scale}), "Perlin 3D Sliced", octave);
}

int main(int argc, char** argv)
{
    const std::string dir = argc > 1 ? argv[1] : "result_raw";
    std::cout << "=== Wavelet & Perlin Noise Comparison Generation ===" << std::endl;
    mkdir(dir.c_str(), 0755);

    const int IMAGE_SIZE = 256;
    const int TILE_SIZE = 128;
    const unsigned int SEED = 12345;

    try {
        std::cout << "\n--- Initializing Wavelet Noise ---" << std::endl;
        WaveletNoise noise2D(TILE_SIZE, SEED);
        noise2D.generateNoiseTile2D();
        WaveletNoise noise3D(TILE_SIZE, SEED);
        noise3D.generateNoiseTile3D();

        std::cout << "\n--- Initializing Perlin Noise ---" << std::endl;
        PerlinNoise perlin(SEED);

        for (int octave : {3, 4, 5}) {
            std::cout << "\n--- Generating Data for Octave " << octave << " ---" << std::endl;
            const std::string o = std::to_string(octave);
            generate2DOctaveBandNoise(IMAGE_SIZE, octave, dir + "/wavelet_noise_2D_octave_" + o + ".raw", noise2D);
            generate3DSlicedOctaveBandNoise(IMAGE_SIZE, octave, dir + "/wavelet_noise_3Dsliced_octave_" + o + ".raw", noise3D);
            generate3DProjectedOctaveBandNoise(IMAGE_SIZE, octave, dir + "/wavelet_noise_3Dprojected_octave_" + o + ".raw", noise3D);
            generatePerlinNoise2D(IMAGE_SIZE, octave, dir + "/perlin_noise_2D_octave_" + o + ".raw", perlin);
            generatePerlinNoise3DSliced(IMAGE_SIZE, octave, dir + "/perlin_noise_3Dsliced_octave_" + o + ".raw", perlin);
        }
    } catch (const std::exception& e) {
        std::cerr << "error: " << e.what() << std::endl;
        return 1;
    }
    std::cout << "\n=== Generation Complete ===" << std::endl;
    std::cout << "Generated files include both Wavelet and Perlin noise for comparison." << std::endl;
    return 0;
}
