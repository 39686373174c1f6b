import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
wn = importlib.import_module("wavelet-noise-in-ray-tracing_b200")
sh = importlib.import_module("wavelet-noise-in-ray-tracing_b200.sharding")
ctx = wn.Context(0); ctx.use_torch_stream()
noise = wn.WaveletNoise(128, 12345, ctx); noise.generateNoiseTile3D()
ax = (np.arange(1024, dtype=np.float32) / np.float32(1024)) * np.float32(4.1)
nz = 256
out = torch.empty((nz, 1024, 1024), dtype=torch.float32, device="cuda")
for lo, hi in ((4, 8), (6, 8), (7, 8), (8, 8), (7, 7), (6, 6), (4, 5), (4, 6), (5, 6), (4, 4)):
    scale, w, post = sh.config3_bands(lo, hi)
    for _ in range(2):
        noise.multiband3D_lattice(ax, ax, ax[:nz], scale, w, float(post), out=out)
    torch.cuda.synchronize()
    k0 = ctx.kernel_launches
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        noise.multiband3D_lattice(ax, ax, ax[:nz], scale, w, float(post), out=out)
    b.record(); torch.cuda.synchronize()
    print(f"bands {lo}..{hi}: {a.elapsed_time(b) / 3:.3f} ms, kernels per call {(ctx.kernel_launches - k0) // 3}", flush=True)
