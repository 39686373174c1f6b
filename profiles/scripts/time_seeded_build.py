#!/usr/bin/env python
"""Wall time of seed -> tile entirely on the GPU (wn_tile_build_seeded: MT19937 + polar method + filters), n=128 and 256."""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
wn = importlib.import_module("wavelet-noise-in-ray-tracing_b200")
ctx = wn.Context(0)
for n in (128, 256):
    w = wn.WaveletNoise(n, 12345, ctx)
    best = 1e9
    for _ in range(6):
        t0 = time.perf_counter()
        w.generate_seeded(3)
        best = min(best, (time.perf_counter() - t0) * 1e3)
    print(f"n={n}: seed -> tile {best:.3f} ms (best of 6, wall, incl. the accepted-count readback)")
