#!/usr/bin/env python
"""Concurrent device->host copy ceiling of one box: every visible GPU copies 1 GiB into its own pinned buffer at the
same time (the e2e arm of bench.py does exactly this with the 4 GiB/N result of a step).  Prints one JSON line:
per-GPU GB/s alone (GPU 0), aggregate GB/s with all GPUs copying, and the e2e Gsamples/s ceiling they imply (4 bytes
per sample)."""
import json
import time

import torch

n = torch.cuda.device_count()
size = 1 << 30
src = [torch.empty(size, dtype=torch.uint8, device=f"cuda:{i}") for i in range(n)]
dst = [torch.empty(size, dtype=torch.uint8, pin_memory=True) for _ in range(n)]
streams = [torch.cuda.Stream(device=i) for i in range(n)]


def run(active, reps=3):
    best = 0.0
    for _ in range(reps):
        for i in active:
            torch.cuda.synchronize(i)
        t0 = time.perf_counter()
        for i in active:
            with torch.cuda.device(i), torch.cuda.stream(streams[i]):
                dst[i].copy_(src[i], non_blocking=True)
        for i in active:
            streams[i].synchronize()
        dt = time.perf_counter() - t0
        best = max(best, len(active) * size / dt / 1e9)
    return best


run([0], 1)
out = {"gpus": n, "one_gpu_gbs": run([0])}
for k in (2, 4, 8):
    if k <= n:
        out[f"aggregate_{k}_gpus_gbs"] = run(list(range(k)))
        out[f"e2e_ceiling_gsamples_s_{k}_gpus"] = out[f"aggregate_{k}_gpus_gbs"] / 4.0
out["e2e_ceiling_gsamples_s_1_gpu"] = out["one_gpu_gbs"] / 4.0
print(json.dumps(out))
