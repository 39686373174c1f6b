#!/usr/bin/env python
"""Runs rank 0's block-cyclic shard of config 3 for a given world size a few times (for ncu launch lists).
Usage: run_shard_once.py <world> [reps=3]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
wn = importlib.import_module("wavelet-noise-in-ray-tracing_b200")
sh = importlib.import_module("wavelet-noise-in-ray-tracing_b200.sharding")

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = wn.Context(0)
ctx.use_torch_stream()
noise = wn.WaveletNoise(128, 12345, ctx)
noise.generateNoiseTile3D()
ax = sh.lattice_axes_config3(1024)
scale, w, post = sh.config3_bands(4, 8)
zs = ax[sh.cyclic_slab_indices(1024, 0, world)].copy()
out = torch.empty((zs.size, 1024, 1024), dtype=torch.float32, device="cuda")
for _ in range(reps):
    noise.multiband3D_lattice(ax, ax, zs, scale, w, float(post), out=out)
torch.cuda.synchronize()
print("done", zs.size, float(out[0, 0, 0]), ctx.kernel_launches)
