#!/usr/bin/env python
"""One warm-up call and `reps` device-resident calls of the GENERAL path: BASELINE config 3's five bands on a lattice with
base range 4.1 (not commensurate with the tile, nothing folds) -- the command the launch list / ncu captures of the
unfolded kernels are taken from.  Prints ms per call by CUDA events.  Usage: run_unfolded_once.py [reps] [nz]"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
wn = importlib.import_module("wavelet-noise-in-ray-tracing_b200")
sh = importlib.import_module("wavelet-noise-in-ray-tracing_b200.sharding")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
nz = int(sys.argv[2]) if len(sys.argv) > 2 else 256
ctx = wn.Context(0)
ctx.use_torch_stream()
noise = wn.WaveletNoise(128, 12345, ctx)
noise.generateNoiseTile3D()
ax = (np.arange(1024, dtype=np.float32) / np.float32(1024)) * np.float32(4.1)
scale, w, post = sh.config3_bands(4, 8)
out = torch.empty((nz, 1024, 1024), dtype=torch.float32, device="cuda")
noise.multiband3D_lattice(ax, ax, ax[:nz], scale, w, float(post), out=out)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    noise.multiband3D_lattice(ax, ax, ax[:nz], scale, w, float(post), out=out)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
print(f"ok {ms:.3f} ms per call, {1024 * 1024 * nz / ms / 1e6:.1f} Gsamples/s, kernels per call {ctx.kernel_launches // (reps + 1)}")
