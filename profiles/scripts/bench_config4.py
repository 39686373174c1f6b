#!/usr/bin/env python
"""BASELINE config 4: WProjectedNoise on an 8192^2 plane with an oblique normal vs Perlin 3D octave 4 on the same grid.
p = (0,0,1) + (i/8192*4) e1 + (j/8192*4) e2, n = (1,2,3)/sqrt(14), e1 = (2,-1,0)/sqrt(5), e2 = (3,6,-5)/sqrt(70);
wavelet: evaluate3DProjected(2 p 2^4, n) / sqrt(0.296); Perlin(12345): noise(p 2^4).  Device-resident outputs, CUDA events;
parity on 2^16 random pixels against the oracle (bit-exact); CPU rate of the oracle/_ref on those pixels for scale."""
import importlib
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
wn = importlib.import_module("wavelet-noise-in-ray-tracing_b200")
from oracle_lib import Oracle  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
f = np.float32
nrm = (np.array([1, 2, 3], np.float64) / np.sqrt(14.0)).astype(f)
e1 = (np.array([2, -1, 0], np.float64) / np.sqrt(5.0)).astype(f)
e2 = (np.array([3, 6, -5], np.float64) / np.sqrt(70.0)).astype(f)
origin = np.array([0, 0, 1], f)
ax = (np.arange(S, dtype=f) / f(S)) * f(4)
pre_w, pre_p = f(2.0 * 2 ** 4), f(2.0 ** 4)
inv = f(1.0) / np.sqrt(f(0.296))

ctx = wn.Context(0)
ctx.use_torch_stream()
noise = wn.WaveletNoise(128, 12345, ctx)
noise.generate_seeded(3)
perlin = wn.PerlinNoise(12345, ctx)
out = torch.empty((S, S), dtype=torch.float32, device="cuda")


def timed(fn, reps=5):
    fn(); fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


ms_proj = timed(lambda: noise.evaluate3DProjected_grid(origin, e1, ax, e2, ax, nrm, float(pre_w), float(inv), out=out))
proj = out.cpu().numpy().copy()
ms_perl = timed(lambda: perlin.noise_grid(origin, e1, ax, e2, ax, float(pre_p), out=out))
perl = out.cpu().numpy().copy()
perlin.set_precision(wn.WN_PERLIN_F32)                        # opt-in FP32 mode of the Perlin kernel
ms_perl32 = timed(lambda: perlin.noise_grid(origin, e1, ax, e2, ax, float(pre_p), out=out))
perl32_err = float(np.abs(out.cpu().numpy() - perl).max())
perlin.set_precision(wn.WN_PERLIN_F64)
ms_plain = timed(lambda: noise.evaluate3D_grid(origin, e1, ax, e2, ax, float(pre_w), 1.0, out=out))

orc = Oracle()
tile = orc.generate_tile(128, 12345, 3)
rs = np.random.RandomState(4)
ii, jj = rs.randint(0, S, 1 << 16), rs.randint(0, S, 1 << 16)
P = ((origin[None, :] + ax[ii][:, None] * e1) + ax[jj][:, None] * e2).astype(f)
t0 = time.perf_counter()
want = orc.eval3d_projected_points(tile, 128, P * pre_w, nrm, 1.0, inv)
cpu_proj_s = time.perf_counter() - t0
t0 = time.perf_counter()
wantp = orc.perlin_points(orc.perlin_perm(12345), P * pre_p)
cpu_perl_s = time.perf_counter() - t0
ok_proj = bool((proj[jj, ii].view(np.uint32) == want.view(np.uint32)).all())
ok_perl = bool((perl[jj, ii].view(np.uint32) == wantp.view(np.uint32)).all())
n = S * S
print(json.dumps({
    "config": f"BASELINE config 4, {S}x{S} plane, normal (1,2,3)/sqrt14", "samples": n,
    "projected_ms": ms_proj, "projected_gsamples_s": n / ms_proj / 1e6,
    "perlin_fp64_ms": ms_perl, "perlin_gsamples_s": n / ms_perl / 1e6,
    "perlin_fp32_ms": ms_perl32, "perlin_fp32_gsamples_s": n / ms_perl32 / 1e6, "perlin_fp32_max_abs_diff_to_fp64": perl32_err,
    "evaluate3d_grid_exact_ms": ms_plain, "evaluate3d_grid_gsamples_s": n / ms_plain / 1e6,
    "bit_exact_vs_oracle_65536_pixels": {"projected": ok_proj, "perlin": ok_perl},
    "cpu_oracle_msamples_s": {"projected": (1 << 16) / cpu_proj_s / 1e6, "perlin": (1 << 16) / cpu_perl_s / 1e6,
                              "threads": orc.max_threads()},
}))
