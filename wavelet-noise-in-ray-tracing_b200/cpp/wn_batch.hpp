// wn_batch.hpp -- additive batch entry points for the C++ drop-in layer.
// The reference's callers evaluate one point per call (experient/main.cpp:18-30, texture.h:82); these free
// functions expose the same evaluators over whole point sets / lattices / grids so a driver can hand the GPU
// one batch instead.  Everything forwards to the C ABI (include/wn_b200.h); errors throw std::runtime_error.
#ifndef WN_BATCH_HPP
#define WN_BATCH_HPP

#include <cstddef>
#include <cstdint>
#include <vector>

#include "../../include/wn_b200.h"

class WaveletNoise;

namespace wnb {

// The process-wide GPU context used by the drop-in classes (created on first use; device = current).
wn_ctx* context();
// Device tile behind a WaveletNoise object (uploads the host coefficients first if the object is a copy).
wn_tile* tile_of(const WaveletNoise& noise);
void check(int rc);                                   // throws std::runtime_error(wn_last_error()) when rc != 0

// out[i] = evaluate*(p_i * pre) * post
std::vector<float> evaluate2D_points(const WaveletNoise& n, const float* xy, size_t count, float pre = 1.0f, float post = 1.0f);
std::vector<float> evaluate3D_points(const WaveletNoise& n, const float* xyz, size_t count, float pre = 1.0f, float post = 1.0f);
std::vector<float> evaluate3DProjected_points(const WaveletNoise& n, const float* xyz, const float normal[3], size_t count,
                                              float pre = 1.0f, float post = 1.0f);
// image[j*nx + i] -- the .raw layout of experient/main.cpp:28-34
std::vector<float> evaluate2D_lattice(const WaveletNoise& n, const std::vector<float>& xs, const std::vector<float>& ys,
                                      float pre, float post);
std::vector<float> multiband3D_lattice(const WaveletNoise& n, const std::vector<float>& xs, const std::vector<float>& ys,
                                       const std::vector<float>& zs, const std::vector<float>& band_scale,
                                       const std::vector<float>& weights, float post, int mode = WN_EVAL_FAST);
std::vector<float> evaluate3DProjected_grid(const WaveletNoise& n, const float origin[3], const float e1[3],
                                            const std::vector<float>& us, const float e2[3], const std::vector<float>& vs,
                                            const float normal[3], float pre, float post);
// texture.h:67-107 for a batch of hit points (grey value per point)
std::vector<float> wavelet_texture_values(const WaveletNoise& noise3d, const float* xyz, size_t count, double scale, int octave);

}  // namespace wnb

#endif
