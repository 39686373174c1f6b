#!/usr/bin/env python
"""Runs the fast multiband lattice kernel a few times on a config-3 slab (for ncu captures).
Usage: run_multiband_once.py [nz=64] [bands=all|<index>] [reps=3]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
wn = importlib.import_module("wavelet-noise-in-ray-tracing_b200")
sh = importlib.import_module("wavelet-noise-in-ray-tracing_b200.sharding")

nz = int(sys.argv[1]) if len(sys.argv) > 1 else 64
bands = sys.argv[2] if len(sys.argv) > 2 else "all"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ctx = wn.Context(0)
ctx.use_torch_stream()
noise = wn.WaveletNoise(128, 12345, ctx)
noise.generateNoiseTile3D()
ax = sh.lattice_axes_config3(1024)
scale, w, post = sh.config3_bands(4, 8)
if bands != "all":
    b = int(bands)
    scale, w = scale[b:b + 1], w[b:b + 1]
out = torch.empty((nz, 1024, 1024), dtype=torch.float32, device="cuda")
for _ in range(reps):
    noise.multiband3D_lattice(ax, ax, ax[:nz], scale, w, float(post), out=out)
torch.cuda.synchronize()
print("done", float(out[0, 0, 0]), ctx.kernel_launches)
