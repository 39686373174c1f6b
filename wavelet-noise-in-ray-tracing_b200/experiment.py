"""Drop-in for the reference's experiment driver (experient/main.cpp:131-168) on the GPU path.

Writes the same 15 float32 `.raw` images (256x256, row-major image[y*256+x], experient/main.cpp:28-34)
with the same file names and the same stdout lines, but every image is one batched call instead of
65 536 scalar evaluate*() calls.  Run:  python -m ... experiment  (see __main__ below) or call main().
"""
import os

import numpy as np

from . import PerlinNoise, WaveletNoise

IMAGE_SIZE = 256            # experient/main.cpp:136
TILE_SIZE = 128             # :137
SEED = 12345                # :138
OCTAVES = (3, 4, 5)         # :149
BASE_RANGE = np.float32(4.0)


def _axis(image_size):
    """u = (float(x)/imageSize)*base_range for every pixel, float32 (main.cpp:20-21)."""
    return (np.arange(image_size, dtype=np.float32) / np.float32(image_size)) * BASE_RANGE


def _write_raw(path, image):
    np.ascontiguousarray(image, dtype="<f4").tofile(path)


def generate2DOctaveBandNoise(imageSize, octave, outputFile, noise):
    """experient/main.cpp:11-36"""
    u = _axis(imageSize)
    pre = np.float32(2.0 ** octave) * np.float32(2.0)            # (u*octave_scale)*2 == u*(2*octave_scale)
    inv_stddev_2d = np.float32(1.0) / np.sqrt(np.float32(0.19686))
    image = noise.evaluate2D_lattice(u, u, float(pre), float(inv_stddev_2d))
    if outputFile:
        _write_raw(outputFile, image)
        print(f"Generated Wavelet 2D Octave {octave} noise: {outputFile}")
    return image


def _sliced_axes(imageSize, octave):
    u = _axis(imageSize)
    pre = np.float32(2.0 ** octave) * np.float32(2.0)
    xy = u * pre
    z = np.array([np.float32(1.0) * np.float32(2.0)], np.float32)   # p[2] = 1.0f; p[2] *= 2.0f (not octave scaled)
    return xy, z


def generate3DSlicedOctaveBandNoise(imageSize, octave, outputFile, noise):
    """experient/main.cpp:38-64"""
    xy, z = _sliced_axes(imageSize, octave)
    inv_stddev_3d = np.float32(1.0) / np.sqrt(np.float32(0.18402))
    from . import WN_EVAL_EXACT
    image = noise.multiband3D_lattice(xy, xy, z, [1.0], [1.0], float(inv_stddev_3d), mode=WN_EVAL_EXACT)[0]
    if outputFile:
        _write_raw(outputFile, image)
        print(f"Generated Wavelet 3D Sliced Octave {octave} noise: {outputFile}")
    return image


def generate3DProjectedOctaveBandNoise(imageSize, octave, outputFile, noise):
    """experient/main.cpp:66-93: normal (0,0,1), plane z = 2"""
    xy, z = _sliced_axes(imageSize, octave)
    inv = np.float32(1.0) / np.sqrt(np.float32(0.296))
    image = noise.evaluate3DProjected_grid([0.0, 0.0, float(z[0])], [1.0, 0.0, 0.0], xy, [0.0, 1.0, 0.0], xy,
                                           [0.0, 0.0, 1.0], 1.0, float(inv))
    if outputFile:
        _write_raw(outputFile, image)
        print(f"Generated Wavelet 3D Projected Octave {octave} noise: {outputFile}")
    return image


def generatePerlinNoise2D(imageSize, octave, outputFile, perlin):
    """experient/main.cpp:95-111"""
    u = _axis(imageSize) * np.float32(2.0 ** octave)
    image = perlin.noise_lattice(u, u, np.zeros(1, np.float32))[0]
    if outputFile:
        _write_raw(outputFile, image)
        print(f"Generated Perlin 2D Octave {octave} noise: {outputFile}")
    return image


def generatePerlinNoise3DSliced(imageSize, octave, outputFile, perlin):
    """experient/main.cpp:113-129"""
    os_ = np.float32(2.0 ** octave)
    u = _axis(imageSize) * os_
    image = perlin.noise_lattice(u, u, np.array([np.float32(1.0) * os_], np.float32))[0]
    if outputFile:
        _write_raw(outputFile, image)
        print(f"Generated Perlin 3D Sliced Octave {octave} noise: {outputFile}")
    return image


def radial_power_spectrum(image):
    """Radially averaged power spectrum of a square image, computed on the GPU (torch.fft = cuFFT): mean removed,
    |fftshift(fft2)|^2 summed over integer-radius rings.  Same quantity as experient/analyze.py:81-157 plots (its
    `power_spectrum` / radial profile), returned as a float64 numpy array indexed by ring radius in FFT bins."""
    import torch
    a = torch.as_tensor(np.ascontiguousarray(image, np.float32), device="cuda").double()
    n = a.shape[0]
    a = a - a.mean()
    power = torch.fft.fftshift(torch.fft.fft2(a)).abs() ** 2
    yy, xx = torch.meshgrid(torch.arange(n, device="cuda"), torch.arange(n, device="cuda"), indexing="ij")
    ring = torch.sqrt(((xx - n // 2) ** 2 + (yy - n // 2) ** 2).double()).long()
    prof = torch.zeros(int(ring.max()) + 1, dtype=torch.float64, device="cuda")
    prof.scatter_add_(0, ring.flatten(), power.flatten())
    return prof.cpu().numpy()


def band_energy(image, cells_per_pixel):
    """Where the power of a single-band noise image lies relative to the band the wavelet construction promises, [1/4, 1/2]
    cycles per tile cell (Cook & DeRose section 3; the paper's Figure 8 / analyze.py:398-463 judge this by eye):
    returns (fraction inside the band widened by 20 % / 15 %, fraction below half the band's lower edge)."""
    prof = radial_power_spectrum(image)
    n = np.asarray(image).shape[0]
    lo, hi = n * 0.25 * cells_per_pixel, n * 0.5 * cells_per_pixel
    total = prof[1:].sum()
    inband = prof[max(1, int(lo * 0.8)):int(hi * 1.15) + 1].sum()
    below = prof[1:max(1, int(lo * 0.5))].sum()
    return float(inband / total), float(below / total)


def raw_to_json_dict(image):
    """The JSON the reference's viewer loads (threejs/convert_raw_to_json.py:12-90): width, height,
    original_range{min,max,mean,std} (float64 statistics of the float32 image) and the data normalised to [0,1]."""
    a = np.asarray(image, dtype=np.float32).astype(np.float64)
    size = int(np.sqrt(a.size))
    a = a.reshape(size, size)
    lo, hi = float(np.min(a)), float(np.max(a))
    norm = (a - lo) / (hi - lo) if hi != lo else np.zeros_like(a)
    return {"width": size, "height": size,
            "original_range": {"min": lo, "max": hi, "mean": float(np.mean(a)), "std": float(np.std(a))},
            "data": norm.flatten().tolist()}


JSON_NAMES = {"wavelet_noise_2D_octave_": "wavelet_noise_2d_octave", "wavelet_noise_3Dsliced_octave_": "wavelet_noise_3d_sliced_octave",
              "wavelet_noise_3Dprojected_octave_": "wavelet_noise_3d_projected_octave",
              "perlin_noise_2D_octave_": "perlin_noise_2d_octave", "perlin_noise_3Dsliced_octave_": "perlin_noise_3d_sliced_octave"}


def export_json(images, json_dir):
    """Write the 15 viewer files (threejs/result_json/*.json naming, convert_raw_to_json.py:92-157)."""
    import json
    os.makedirs(json_dir, exist_ok=True)
    for name, image in images.items():
        for prefix, out_prefix in JSON_NAMES.items():
            if name.startswith(prefix):
                with open(os.path.join(json_dir, out_prefix + name[len(prefix):] + ".json"), "w") as f:
                    json.dump(raw_to_json_dict(image), f, separators=(",", ":"))


def main(out_dir="result_raw", ctx=None):
    """experient/main.cpp:131-168"""
    print("=== Wavelet & Perlin Noise Comparison Generation ===")
    os.makedirs(out_dir, exist_ok=True)
    print("\n--- Initializing Wavelet Noise ---")
    noise2D = WaveletNoise(TILE_SIZE, SEED, ctx)
    noise2D.generateNoiseTile2D()
    noise3D = WaveletNoise(TILE_SIZE, SEED, ctx)
    noise3D.generateNoiseTile3D()
    print("\n--- Initializing Perlin Noise ---")
    perlin = PerlinNoise(SEED, ctx)
    images = {}
    for octave in OCTAVES:
        print(f"\n--- Generating Data for Octave {octave} ---")
        o = str(octave)
        j = lambda name: os.path.join(out_dir, name) if out_dir else None   # noqa: E731
        images["wavelet_noise_2D_octave_" + o] = generate2DOctaveBandNoise(
            IMAGE_SIZE, octave, j(f"wavelet_noise_2D_octave_{o}.raw"), noise2D)
        images["wavelet_noise_3Dsliced_octave_" + o] = generate3DSlicedOctaveBandNoise(
            IMAGE_SIZE, octave, j(f"wavelet_noise_3Dsliced_octave_{o}.raw"), noise3D)
        images["wavelet_noise_3Dprojected_octave_" + o] = generate3DProjectedOctaveBandNoise(
            IMAGE_SIZE, octave, j(f"wavelet_noise_3Dprojected_octave_{o}.raw"), noise3D)
        images["perlin_noise_2D_octave_" + o] = generatePerlinNoise2D(
            IMAGE_SIZE, octave, j(f"perlin_noise_2D_octave_{o}.raw"), perlin)
        images["perlin_noise_3Dsliced_octave_" + o] = generatePerlinNoise3DSliced(
            IMAGE_SIZE, octave, j(f"perlin_noise_3Dsliced_octave_{o}.raw"), perlin)
    print("\n=== Generation Complete ===")
    print("Generated files include both Wavelet and Perlin noise for comparison.")
    return images


if __name__ == "__main__":
    main()
