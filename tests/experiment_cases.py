"""The 15 experiment images of experient/main.cpp:131-168 described as data (shared by CPU and GPU tests).

Coordinates restate experient/main.cpp:20-26: u = (float(x)/imageSize)*4.0f; p = u*2^octave; p *= 2.
"""
import numpy as np

IMAGE_SIZE = 256
TILE_SIZE = 128
SEED = 12345
OCTAVES = (3, 4, 5)
BASE_RANGE = np.float32(4.0)
INV_STD_2D = np.float32(1.0) / np.sqrt(np.float32(0.19686))      # experient/main.cpp:16
INV_STD_3D = np.float32(1.0) / np.sqrt(np.float32(0.18402))      # :43
INV_STD_PROJ = np.float32(1.0) / np.sqrt(np.float32(0.296))      # :72


def axis_coords(size=IMAGE_SIZE):
    """u for every pixel index, float32 arithmetic exactly as the reference (main.cpp:20-21)."""
    i = np.arange(size, dtype=np.float32)
    return (i / np.float32(size)) * BASE_RANGE


def octave_pre_scale(octave):
    """(u * octave_scale) * 2  ==  u * (2*octave_scale) exactly (power of two)."""
    return np.float32(2.0 ** octave) * np.float32(2.0)


def raw_name(kind, octave):
    return {
        "w2d": f"wavelet_noise_2D_octave_{octave}.raw",
        "w3d": f"wavelet_noise_3Dsliced_octave_{octave}.raw",
        "wproj": f"wavelet_noise_3Dprojected_octave_{octave}.raw",
        "p2d": f"perlin_noise_2D_octave_{octave}.raw",
        "p3d": f"perlin_noise_3Dsliced_octave_{octave}.raw",
    }[kind]


def load_raw(golden_dir, kind, octave):
    import os
    return np.fromfile(os.path.join(golden_dir, "result_raw", raw_name(kind, octave)), dtype="<f4").reshape(
        IMAGE_SIZE, IMAGE_SIZE)
