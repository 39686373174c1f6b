#!/usr/bin/env python
"""One rank's share of BASELINE config 3 at N ranks (block-cyclic z shard), timed on one GPU: ms per call for the knob
sets given on the command line as KEY=VALUE,KEY=VALUE groups.  Usage: shard_time.py N [knobs ...]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
wn = importlib.import_module("wavelet-noise-in-ray-tracing_b200")
sh = importlib.import_module("wavelet-noise-in-ray-tracing_b200.sharding")
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
groups = sys.argv[2:] or [""]
ctx = wn.Context(0)
ctx.use_torch_stream()
noise = wn.WaveletNoise(128, 12345, ctx)
noise.generateNoiseTile3D()
ax = sh.lattice_axes_config3(1024)
scale, w, post = sh.config3_bands(4, 8)
zs = ax[sh.cyclic_slab_indices(1024, 0, world)]
out = torch.empty((zs.size, 1024, 1024), dtype=torch.float32, device="cuda")
first = None
for grp in groups:
    env = dict(kv.split("=") for kv in grp.split(",") if kv)
    os.environ.update(env)
    for _ in range(5):
        noise.multiband3D_lattice(ax, ax, zs, scale, w, float(post), out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50
    a.record()
    for _ in range(reps):
        noise.multiband3D_lattice(ax, ax, zs, scale, w, float(post), out=out)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    sample = out[::16, ::8].clone()
    same = "" if first is None else "  bitwise == first: %s" % bool(torch.equal(sample, first))
    first = sample if first is None else first
    print(f"N={world} shard, knobs [{grp}]: {ms * 1e3:8.1f} us per call = {1024 * 1024 * zs.size / ms / 1e6:8.1f} Gsamples/s "
          f"(x{world} = {1024 * 1024 * zs.size * world / ms / 1e6:8.1f}){same}", flush=True)
    for k in env:
        os.environ.pop(k, None)
