#!/usr/bin/env python
"""Host-side cost of one fast-lattice call (enqueue only) vs its GPU time, for the N=8 per-rank slab of config 3."""
import importlib
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
wn = importlib.import_module("wavelet-noise-in-ray-tracing_b200")
sh = importlib.import_module("wavelet-noise-in-ray-tracing_b200.sharding")
nz = int(sys.argv[1]) if len(sys.argv) > 1 else 128
world = int(sys.argv[2]) if len(sys.argv) > 2 else 0      # > 0: rank 0's block-cyclic shard of the 1024 axis instead of ax[:nz]
ctx = wn.Context(0)
ctx.use_torch_stream()
noise = wn.WaveletNoise(128, 12345, ctx)
noise.generateNoiseTile3D()
ax = sh.lattice_axes_config3(1024)
scale, w, post = sh.config3_bands(4, 8)
zs = ax[sh.cyclic_slab_indices(1024, 0, world)].copy() if world > 0 else ax[:nz]
nz = zs.size
out = torch.empty((nz, 1024, 1024), dtype=torch.float32, device="cuda")
for _ in range(5):
    noise.multiband3D_lattice(ax, ax, zs, scale, w, float(post), out=out)
torch.cuda.synchronize()
reps = 200
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
a.record()
for _ in range(reps):
    noise.multiband3D_lattice(ax, ax, zs, scale, w, float(post), out=out)
b.record()
t_enq = time.perf_counter() - t0
torch.cuda.synchronize()
print(f"nz={nz}: host enqueue {t_enq / reps * 1e6:.1f} us per call, GPU elapsed {a.elapsed_time(b) / reps * 1e3:.1f} us per call, "
      f"{1024 * 1024 * nz / (a.elapsed_time(b) / reps) / 1e6:.1f} Gsamples/s")
