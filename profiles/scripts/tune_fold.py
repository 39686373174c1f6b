#!/usr/bin/env python
"""A/B timing of the fold budget / main-kernel choice on BASELINE config 3 (1024^3, bands 4..8, device-resident output).
Each variant is a set of tuning environment knobs (read per call by the library).  Prints ms per call, Gsamples/s,
the max difference to the exact kernel on 4 slices, and whether every variant agrees bitwise with the first one
(canonical summation order: the result does not depend on the fold set or the kernel).
Usage: tune_fold.py [nz]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
wn = importlib.import_module("wavelet-noise-in-ray-tracing_b200")
sh = importlib.import_module("wavelet-noise-in-ray-tracing_b200.sharding")

nz = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
KNOBS = ("WN_COL4", "WN_FOLD_BUDGET", "WN_REPLICA_ORDER", "WN_BRICK", "WN_RING")
VARIANTS = [
    ("default (col4, 512 MiB block, ring 8)", {}),
    ("ring 4", {"WN_RING": "4"}),
    ("ring 8 also for the 512^3 block kernel", {"WN_RING": "8"}),
    ("register look-ahead 2", {"WN_RING": "-2"}),
    ("register look-ahead 4", {"WN_RING": "-4"}),
    ("plain block order", {"WN_REPLICA_ORDER": "0"}),
    ("64 MiB block", {"WN_FOLD_BUDGET": str(1 << 24)}),
    ("brick4 main kernel", {"WN_COL4": "0"}),
    ("brick4, 64 MiB block (round-1 midpoint)", {"WN_COL4": "0", "WN_FOLD_BUDGET": str(1 << 24)}),
    ("no folding", {"WN_FOLD_BUDGET": "0"}),
]
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
keep = {}
for name, env in VARIANTS:
    for k in KNOBS:
        os.environ.pop(k, None)
    os.environ.update(env)
    for _ in range(3):
        noise.multiband3D_lattice(ax, ax, zs, scale, w, float(post), out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    a.record()
    for _ in range(reps):
        noise.multiband3D_lattice(ax, ax, zs, scale, w, float(post), out=out)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    err = float((out[:4] - ref).abs().max())
    # a strided sample of the volume, kept for the bitwise comparison between variants with the same fold set
    sample = out[:: max(1, nz // 64), ::16, :].clone()
    same = ""
    if "first" in keep:
        same = "  bitwise == first variant: %s" % bool(torch.equal(sample, keep["first"]))
    else:
        keep["first"] = sample
    print(f"{name:40s} {ms:8.3f} ms  {1024 * 1024 * nz / ms / 1e6:8.2f} Gsamples/s  max|fast-exact|={err:.3g}{same}", flush=True)
