#!/usr/bin/env python
"""A/B timing of the replica kernel (k_mb3d_rep) on BASELINE config 3 (1024^3, bands 4..8, device-resident output).
Each variant is a set of tuning environment knobs (read per call by the library).  Prints ms per call, Gsamples/s,
the max difference to the exact kernel on 4 slices, and whether the variant agrees BITWISE with the first one on a
strided sample of the volume (canonical summation: the result must not depend on the kernel or the fold set).
Usage: tune_rep.py [nz] [reps]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
wn = importlib.import_module("wavelet-noise-in-ray-tracing_b200")
sh = importlib.import_module("wavelet-noise-in-ray-tracing_b200.sharding")

nz = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
KNOBS = ("WN_REP", "WN_REP_YPW", "WN_REP_TMA", "WN_REP_SHARE", "WN_FOLD_BUDGET", "WN_REPLICA_ORDER", "WN_COL4", "WN_REP_BY", "WN_BRICK_FALLBACK")
B24 = str(1 << 24)
VARIANTS = [
    ("col4 (round 1), 512^3 block", {"WN_REP": "0"}),
    ("col4 (round 1), 256^3 block", {"WN_REP": "0", "WN_FOLD_BUDGET": B24}),
]
for budget_name, budget in (("512^3", None), ("256^3", B24)):
    for rep in ("11", "12", "22"):
        for tma in ("1", "0"):
            for share in (("1", "0") if budget else ("1",)):
                env = {"WN_REP": rep, "WN_REP_TMA": tma, "WN_REP_SHARE": share}
                if budget:
                    env["WN_FOLD_BUDGET"] = budget
                if rep == "11" and share == "0":
                    continue
                VARIANTS.append((f"rep {rep} {'tma' if tma == '1' else 'cp.async'} share {share} {budget_name} block", env))
VARIANTS.append(("default plan, 8-row bricks", {}))
VARIANTS.append(("default plan, 16-row bricks", {"WN_REP_BY": "16"}))
for idx in range(10):
    VARIANTS.append((f"chain shape {idx}", {"WN_BRICK_FALLBACK": str(idx)}))
if os.environ.get("TUNE_ONLY"):
    keep = os.environ["TUNE_ONLY"].split(",")
    VARIANTS = [v for i, v in enumerate(VARIANTS) if i < 2 or any(k in v[0] for k in keep)]

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
first = None
for name, env in VARIANTS:
    for k in KNOBS:
        os.environ.pop(k, None)
    os.environ.update(env)
    out.zero_()
    for _ in range(3):
        noise.multiband3D_lattice(ax, ax, zs, scale, w, float(post), out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        noise.multiband3D_lattice(ax, ax, zs, scale, w, float(post), out=out)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    err = float((out[:4] - ref).abs().max())
    sample = out[:: max(1, nz // 64), ::8, :].clone()
    same = ""
    if first is None:
        first = sample
    else:
        same = "  bitwise == first: %s" % bool(torch.equal(sample, first))
    print(f"{name:44s} {ms:8.3f} ms  {1024 * 1024 * nz / ms / 1e6:8.2f} Gsamples/s  max|fast-exact|={err:.3g}{same}", flush=True)
