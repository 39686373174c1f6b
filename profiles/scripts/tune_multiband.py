#!/usr/bin/env python
"""Times the fast multiband lattice kernel (device-resident output) for each brick shape (WN_BRICK index)
on BASELINE config 3 and checks it against the exact kernel on a sub-slab.  Usage: tune_multiband.py [nz]"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
wn = importlib.import_module("wavelet-noise-in-ray-tracing_b200")
sh = importlib.import_module("wavelet-noise-in-ray-tracing_b200.sharding")

nz = int(sys.argv[1]) if len(sys.argv) > 1 else 256
shapes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["0", "1", "2", "3", "4", "5"]
ctx = wn.Context(0)
ctx.use_torch_stream()
noise = wn.WaveletNoise(128, 12345, ctx)
noise.generateNoiseTile3D()
ax = sh.lattice_axes_config3(1024)
scale, w, post = sh.config3_bands(4, 8)
zs = ax[:nz]
out = torch.empty((nz, 1024, 1024), dtype=torch.float32, device="cuda")
ref = noise.multiband3D_lattice(ax, ax, zs[:4], scale, w, float(post), mode=wn.WN_EVAL_EXACT, device_out=True)
torch.cuda.synchronize()
for s in shapes:
    if s == "auto":
        os.environ.pop("WN_BRICK", None)
    else:
        os.environ["WN_BRICK"] = s
    for _ in range(3):
        noise.multiband3D_lattice(ax, ax, zs, scale, w, float(post), out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    reps = 20
    for _ in range(reps):
        noise.multiband3D_lattice(ax, ax, zs, scale, w, float(post), out=out)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    err = float((out[:4] - ref).abs().max())
    print(f"WN_BRICK={s}: {ms:8.3f} ms  {1024 * 1024 * nz / ms / 1e6:8.2f} Gsamples/s  max|fast-exact|={err:.3g}", flush=True)
# per-band cost (single band at a time)
if shapes[0] != "auto":
    os.environ["WN_BRICK"] = shapes[0]
for bi in range(len(scale) if os.environ.get("WN_PER_BAND") else 0):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    noise.multiband3D_lattice(ax, ax, zs, scale[bi:bi + 1], w[bi:bi + 1], float(post), out=out)
    a.record()
    for _ in range(3):
        noise.multiband3D_lattice(ax, ax, zs, scale[bi:bi + 1], w[bi:bi + 1], float(post), out=out)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    print(f"  band scale {scale[bi]:6.0f} alone (shape {shapes[0]}): {ms:8.3f} ms  {1024 * 1024 * nz / ms / 1e6:8.2f} Gsamples/s", flush=True)
