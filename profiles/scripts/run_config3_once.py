#!/usr/bin/env python
"""One warm-up call and `reps` device-resident calls of BASELINE config 3 (1024^3, bands 4..8) -- the command the ncu
captures of the headline kernels are taken from.  Tuning knobs come from the environment (WN_REP, WN_FOLD_BUDGET ...).
Usage: run_config3_once.py [reps] [nz]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
wn = importlib.import_module("wavelet-noise-in-ray-tracing_b200")
sh = importlib.import_module("wavelet-noise-in-ray-tracing_b200.sharding")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
nz = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
ctx = wn.Context(0)
ctx.use_torch_stream()
noise = wn.WaveletNoise(128, 12345, ctx)
noise.generateNoiseTile3D()
ax = sh.lattice_axes_config3(1024)
scale, w, post = sh.config3_bands(4, 8)
out = torch.empty((nz, 1024, 1024), dtype=torch.float32, device="cuda")
for _ in range(1 + reps):
    noise.multiband3D_lattice(ax, ax, ax[:nz], scale, w, float(post), out=out)
torch.cuda.synchronize()
print("ok", float(out[0, 0, 0]))
