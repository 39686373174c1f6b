#!/usr/bin/env python
"""Brick-shape sweep (WN_BRICK=0..9, read per call) of the general path: band subsets of BASELINE config 3 on the
non-commensurate lattice (base range 4.1), 1024 x 1024 x 256 samples, device-resident.  ms per call per shape."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
wn = importlib.import_module("wavelet-noise-in-ray-tracing_b200")
sh = importlib.import_module("wavelet-noise-in-ray-tracing_b200.sharding")
ctx = wn.Context(0); ctx.use_torch_stream()
noise = wn.WaveletNoise(128, 12345, ctx); noise.generateNoiseTile3D()
ax = (np.arange(1024, dtype=np.float32) / np.float32(1024)) * np.float32(4.1)
nz = 256
out = torch.empty((nz, 1024, 1024), dtype=torch.float32, device="cuda")
for lo, hi in ((8, 8), (7, 7), (7, 8), (4, 8)):
    scale, w, post = sh.config3_bands(lo, hi)
    row = []
    for shape in [None] + list(range(10)):
        if shape is None:
            os.environ.pop("WN_BRICK", None)
        else:
            os.environ["WN_BRICK"] = str(shape)
        for _ in range(2):
            noise.multiband3D_lattice(ax, ax, ax[:nz], scale, w, float(post), out=out)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            noise.multiband3D_lattice(ax, ax, ax[:nz], scale, w, float(post), out=out)
        b.record(); torch.cuda.synchronize()
        row.append(f"{'auto' if shape is None else shape}:{a.elapsed_time(b) / 3:.3f}")
    print(f"bands {lo}..{hi}: " + "  ".join(row), flush=True)
