// WaveletNoise.h -- interface-compatible declaration of the reference's `class WaveletNoise`
// (reference: WaveletNoise.h:11-59).  Same public methods, same data members in the same order, so
// callers written against the reference header (experient/main.cpp, texture.h) compile and link
// unchanged against this directory's WaveletNoise.cpp, whose methods run on the GPU through the C ABI
// of include/wn_b200.h.  A maintainer may equally keep the reference's own header: WaveletNoise.cpp
// only relies on the member names declared here.
#ifndef WAVELET_NOISE_H
#define WAVELET_NOISE_H

#include <iostream>
#include <limits>
#include <random>
#include <string>
#include <vector>

struct DataStats {                                   // reference WaveletNoise.h:11-18
    float avg = 0.0f, var = 0.0f;
    float min_val = std::numeric_limits<float>::max();
    float max_val = std::numeric_limits<float>::lowest();
    long long count_nan_inf = 0;                     // never written (as in the reference)
    float energy = 0.0f;                             // never written (as in the reference)
};

class WaveletNoise {
  public:
    WaveletNoise(int tileSize, unsigned int seed = 0);
    ~WaveletNoise();

    void generateNoiseTile2D();                                     // GPU: separable down/up passes + subtract
    void generateNoiseTile3D();

    float evaluate2D(const float p[2]) const;                       // scalar calls = 1-point GPU launches
    float evaluate3D(const float p[3]) const;
    float evaluate3DProjected(const float p[3], const float normal[3]) const;

    DataStats calculateStats(const std::vector<float>& data, const std::string& name) const;
    const std::vector<float>& getNoiseCoefficients() const;
    int getTileSize() const;

  private:
    int tileSizeN;
    std::vector<float> noiseCoefficients;            // host copy of the tile (downloaded after generation)
    unsigned int randomSeed;
    std::mt19937 rng;                                // the Gaussian field is drawn from these two objects,
    std::normal_distribution<float> gaussianDist;    // exactly as the reference does (state persists)

    static const int ARAD = 16;
    static const float A_COEFFS[2 * ARAD];           // kept for layout/ABI parity; the filters run on the GPU
    static const float P_COEFFS[4];

    int Mod(int x, int n) const;
    void downsample1D(const std::vector<float>& from, std::vector<float>& to, int n, int stride);
    void upsample1D(const std::vector<float>& from, std::vector<float>& to, int n, int stride);
};

#endif
