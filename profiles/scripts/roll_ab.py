#!/usr/bin/env python
"""A/B of the rolling-plane kernel (k_mb3d_roll, WN_ROLL=1 default / 0 off) on the general path: band subsets of BASELINE
config 3 on the non-commensurate lattice (base range 4.1), 1024 x 1024 x 256 samples, device-resident.  ms per call,
bitwise agreement of the two results, and the max difference to the exact kernel on 4 slices."""
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
for lo, hi in ((4, 8), (7, 8), (8, 8), (7, 7), (4, 6), (6, 8)):
    scale, w, post = sh.config3_bands(lo, hi)
    ref = noise.multiband3D_lattice(ax, ax, ax[:4], scale, w, float(post), mode=wn.WN_EVAL_EXACT, device_out=True)
    line = []
    for roll in (0, 1):
        os.environ["WN_ROLL"] = str(roll)
        out = outs[roll]
        out.zero_()
        for _ in range(2):
            noise.multiband3D_lattice(ax, ax, ax[:nz], scale, w, float(post), out=out)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            noise.multiband3D_lattice(ax, ax, ax[:nz], scale, w, float(post), out=out)
        b.record(); torch.cuda.synchronize()
        line.append(f"roll={roll}: {a.elapsed_time(b) / 3:.3f} ms")
    same = bool(torch.equal(outs[0].view(torch.int32), outs[1].view(torch.int32)))
    err = float((outs[1][:4] - ref).abs().max())
    print(f"bands {lo}..{hi}: " + "  ".join(line) + f"  bitwise equal: {same}  max|fast-exact| {err:.2e}", flush=True)
