#!/usr/bin/env python
"""Device-side rate of the scattered-point entry points (the shape of BASELINE config 5: hit points in no particular
order), device-resident AoS points, 2^25 per call.  Three orders of the SAME point set: random, sorted by tile cell
(z, y, x of the integer cell of the finest band), and "surface" (a 4096 x 8192 raster of a tilted plane, what a
primary-ray hit buffer looks like).  Prints Gpoints/s per entry point and order.
Usage: points_rate.py [log2 count] [reps]"""
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
wn = importlib.import_module("wavelet-noise-in-ray-tracing_b200")

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 25
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
count = 1 << lg
ctx = wn.Context(0)
ctx.use_torch_stream()
noise = wn.WaveletNoise(128, 12345, ctx)
noise.generateNoiseTile3D()
perlin = wn.PerlinNoise(ctx=ctx)
g = torch.Generator(device="cuda").manual_seed(1)
rand = torch.rand((count, 3), device="cuda", generator=g) * 40.0 - 20.0          # texture space of the render: |p| <= 20
cell = torch.floor(rand * 16.0).to(torch.int64) + 1024                          # finest band of the 5-band sum below
key = (cell[:, 2] << 24) | (cell[:, 1] << 12) | cell[:, 0]
sorted_pts = rand[torch.argsort(key)].contiguous()
del cell, key
h, w = 1 << (lg // 2), 1 << (lg - lg // 2)
u = (torch.arange(w, device="cuda", dtype=torch.float32) + 0.5) / w * 40.0 - 20.0
v = (torch.arange(h, device="cuda", dtype=torch.float32) + 0.5) / h * 40.0 - 20.0
surface = torch.stack([u[None, :].expand(h, w), v[:, None].expand(h, w), 0.3 * u[None, :] + 0.2 * v[:, None]], dim=-1)
surface = surface.reshape(count, 3).contiguous()
out = torch.empty(count, dtype=torch.float32, device="cuda")
scale = np.array([1.0, 2.0, 4.0, 8.0, 16.0], np.float32)
weights = np.array([1.0, 0.5, 0.25, 0.125, 0.0625], np.float32)
normal = np.array([0.0, 0.6, 0.8], np.float32)
CALLS = [
    ("evaluate3D_points", lambda p: noise.evaluate3D_points(p, out=out)),
    ("multiband3D_points x5 bands", lambda p: noise.multiband3D_points(p, scale, weights, out=out)),
    ("evaluate3DProjected_points", lambda p: noise.evaluate3DProjected_points(p, normal, out=out)),
    ("wavelet texture_values (4 octaves)", lambda p: noise.texture_values(p, 4.0, 4, out=out)),
    ("perlin noise_points (FP64)", lambda p: perlin.noise_points(p, out=out)),
]
res = {"points": count, "reps": reps, "rates_gpoints_per_s": {}}
for name, call in CALLS:
    row = {}
    for order, pts in (("random", rand), ("cell-sorted", sorted_pts), ("surface raster", surface)):
        for _ in range(2):
            call(pts)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            call(pts)
        b.record()
        torch.cuda.synchronize()
        row[order] = round(count * reps / (a.elapsed_time(b) * 1e-3) / 1e9, 2)
    res["rates_gpoints_per_s"][name] = row
    print(f"{name:38s} " + "  ".join(f"{k}: {v:7.2f}" for k, v in row.items()), flush=True)
print(json.dumps(res))
