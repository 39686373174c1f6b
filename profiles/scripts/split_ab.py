#!/usr/bin/env python
"""A/B of the in-place band split (brick_pass, WN_SPLIT=1 default / 0 off) on the general path: band subsets of BASELINE
config 3 on the non-commensurate lattice (base range 4.1), 1024 x 1024 x 256 samples, device-resident.  ms per call,
kernels per call and bitwise agreement of the two results."""
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
outs = [torch.empty((nz, 1024, 1024), dtype=torch.float32, device="cuda") for _ in range(2)]
for lo, hi in ((4, 8), (6, 8), (4, 6), (5, 8)):
    scale, w, post = sh.config3_bands(lo, hi)
    line = []
    for split in (0, 1):
        os.environ["WN_SPLIT"] = str(split)
        out = outs[split]
        out.zero_()
        for _ in range(2):
            noise.multiband3D_lattice(ax, ax, ax[:nz], scale, w, float(post), out=out)
        torch.cuda.synchronize()
        k0 = ctx.kernel_launches
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            noise.multiband3D_lattice(ax, ax, ax[:nz], scale, w, float(post), out=out)
        b.record(); torch.cuda.synchronize()
        line.append(f"split={split}: {a.elapsed_time(b) / 3:.3f} ms ({(ctx.kernel_launches - k0) // 3} kernels)")
    same = bool(torch.equal(outs[0].view(torch.int32), outs[1].view(torch.int32)))
    print(f"bands {lo}..{hi}: " + "  ".join(line) + f"  bitwise equal: {same}", flush=True)
