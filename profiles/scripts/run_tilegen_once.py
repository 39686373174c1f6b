#!/usr/bin/env python
"""Builds the n=128 (and n=256) 3D tile a few times from a device-resident Gaussian field (for ncu launch lists)."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
wn = importlib.import_module("wavelet-noise-in-ray-tracing_b200")
ctx = wn.Context(0)
ctx.use_torch_stream()
for n in (128, 256):
    tg = wn.WaveletNoise(n, 12345, ctx)
    R = torch.randn(n ** 3, device="cuda")
    for _ in range(3):
        tg.generateNoiseTile3D(field=R)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        tg.generateNoiseTile3D(field=R)
    b.record()
    torch.cuda.synchronize()
    print(f"n={n}: {a.elapsed_time(b) / 10:.3f} ms per tile (device field -> tile, incl. padded replica)")
