#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline metric on BASELINE config 3.

  metric   : Gsamples/s of 3D multiband wavelet noise (one sample = one output float, all its bands)
  workload : WMultibandNoise dense volume 1024^3, bands 4..8 (q_b = 2 p 2^b, w_b = 2^-(b-4)), tile n=128 seed 12345,
             z-slab sharded over the N ranks (strong scaling: the 1024^3 volume is fixed), one process per GPU.
  step     : one pass of the hot path over the rank's slab (1024 x 1024 x 1024/N samples) into device memory.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line on rank 0 (see DESIGN.md "Measurement" for every field).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = "wavelet-noise-in-ray-tracing_b200"
sys.path.insert(0, ROOT)

# Host threads of the CPU reference arm: every core this process may run on, taken at import (before any NUMA binding)
# and passed EXPLICITLY to the reference loop.  torchrun exports OMP_NUM_THREADS=1, so omp_get_max_threads() would
# silently drop the reference to one thread at N >= 2 and make vs_reference mean something else at every N.
HOST_THREADS = len(os.sched_getaffinity(0))

VOLUME = 1024
TILE_N = 128
SEED = 12345
FLOP_PER_SAMPLE = 475.0          # 5 bands x 95 FLOP (separable form)
OUT_BYTES_PER_SAMPLE = 4.0       # one float32 written to HBM per sample
METRIC = "Gsamples/s 3D multiband wavelet noise"
UNIT = "Gsamples/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, n_gpus=1):
        self.gpu = gpu_index
        self.n_gpus = n_gpus            # > 1: sample GPUs 0..n_gpus-1 (one nvidia-smi process on rank 0 watches every rank's GPU)
        self.proc = None
        self.lines = []

    def start(self):
        try:
            which = str(self.gpu) if self.n_gpus <= 1 else ",".join(str(i) for i in range(self.n_gpus))
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", which],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        per_gpu = {}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
                per_gpu.setdefault(f[0], []).append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        out = {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
               "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}
        if len(per_gpu) > 1:
            out["per_gpu_sm_mhz"] = {k: float(np.median(v)) for k, v in sorted(per_gpu.items())}
        return out


TRAFFIC_CSV = "profiles/r2_main_raw.csv"


def extra_peaks():
    """Roofline denominators MEASURED_PEAKS.json does not hold (FP32 FMA rate, L2 read bandwidth, HBM write stream,
    pinned D2H): measured once on this pool's B200 by profiles/scripts/peaks.cu, committed as profiles/r2_peaks.json."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r2_peaks.json")))
    except (OSError, ValueError):
        return {}


def profiled_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel (the longest launch in the capture), one
    launch at N=1 bench size, from the committed `ncu --set full` capture (TRAFFIC_CSV); None when the capture is
    missing."""
    import csv
    path = os.path.join(ROOT, TRAFFIC_CSV)
    try:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        tscale = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
        d = hdr.index("gpu__time_duration.sum")
        vals = max(rows[2:], key=lambda r: float(r[d].replace(",", "")) * tscale[units[d]])
        total = 0.0
        for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(name)
            total += float(vals[i].replace(",", "")) * scale[units[i]]
        return total
    except (OSError, ValueError, KeyError, IndexError):
        return None


def config3(wnsh):
    ax = wnsh.lattice_axes_config3(VOLUME)
    scale, w, post = wnsh.config3_bands(4, 8)
    return ax, scale, w, post


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation of the path (oracle/_ref when built, else the port)
# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(ax, scale, w, post, target_seconds, threads=0):
    """Times the composed multiband loop over the reference's evaluate3D on a bounded slab.
    Returns dict(value Gsamples/s, cores, kind, sample, seconds)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle, RefLib
    if RefLib.available():
        ref = RefLib()
        noise = ref.noise(TILE_N, SEED).generate(3)
        cores = HOST_THREADS if threads <= 0 else threads
        run = lambda zs: noise.multiband3d_lattice(ax, ax, zs, scale, w, post, threads=cores)   # noqa: E731
        kind = "reference"
    else:
        orc = Oracle()
        tile = orc.generate_tile(TILE_N, SEED, 3)
        cores = HOST_THREADS if threads <= 0 else threads
        run = lambda zs: orc.multiband3d_lattice(tile, TILE_N, ax, ax, zs, scale, w, post, threads=cores)   # noqa: E731
        kind = "port"
    t0 = time.perf_counter()
    run(ax[512:513])                                         # calibration: one 1024^2 slice
    per_slice = time.perf_counter() - t0
    nz = int(max(1, min(256, round(target_seconds / max(per_slice, 1e-6)))))
    zs = ax[512:512 + nz]
    t0 = time.perf_counter()
    run(zs)
    dt = time.perf_counter() - t0
    samples = VOLUME * VOLUME * nz
    return {"value": samples / dt / 1e9, "unit": UNIT, "cores": int(cores), "kind": kind,
            "sample": f"{VOLUME}x{VOLUME}x{nz} slab of the 1024^3 volume (z index 512..), {dt:.2f} s", "seconds": dt,
            "nz": nz, "run": run}


def run_reference(args, rank):
    if rank != 0:
        return
    wnsh_ax = (np.arange(VOLUME, dtype=np.float32) / np.float32(VOLUME)) * np.float32(4.0)
    b = np.arange(4, 9)
    scale = (2.0 * 2.0 ** b).astype(np.float32)
    w = (2.0 ** -(b - 4).astype(np.float64)).astype(np.float32)
    post = np.float32(1.0) / np.sqrt(np.float32((w * w).sum()) * np.float32(0.18402))
    cal = cpu_reference_rate(wnsh_ax, scale, w, post, target_seconds=3.0)
    run, nz = cal["run"], cal["nz"]
    zs = wnsh_ax[512:512 + nz]
    for _ in range(args.warmup):
        run(zs)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run(zs)
    dt = time.perf_counter() - t0
    samples = VOLUME * VOLUME * nz * args.steps
    value = samples / dt / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "WMultibandNoise 1024^3 bands 4-8, tile n=128 seed 12345 (BASELINE config 3); "
                               f"each step is a bounded {VOLUME}x{VOLUME}x{nz} slab on the host CPU",
                   "tile_n": TILE_N, "bands": [4, 8]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cal["cores"], "kind": cal["kind"],
                         "sample": f"{VOLUME}x{VOLUME}x{nz} slab per step x {args.steps} steps"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    wn = importlib.import_module(PKG)
    wnsh = importlib.import_module(PKG + ".sharding")
    torch.cuda.set_device(local_rank)
    # host staging buffers (the e2e arm's pinned output) go on the NUMA node next to this rank's GPU
    prev_affinity = None if os.environ.get("WN_NO_NUMA_BIND") else wnsh.bind_to_gpu_numa(local_rank)
    ctx = wn.Context(local_rank)
    ctx.use_torch_stream()
    ax, scale, w, post = config3(wnsh)

    # --- tile: rank 0 builds (reference generator sequence on the host + GPU filter passes), one broadcast replicates it
    noise = wn.WaveletNoise(TILE_N, SEED, ctx)
    t0 = time.perf_counter()
    wnsh.replicate_noise(noise, 3, lambda nz: nz.generateNoiseTile3D())
    ctx.synchronize()
    tile_total_ms = (time.perf_counter() - t0) * 1e3
    tile_kernel_ms = ctx.last_kernel_ms if rank == 0 else None

    # z sharding: 32-slice chunks dealt round-robin (sharding.cyclic_slab_indices); at N=1 this is the whole axis
    zidx = wnsh.cyclic_slab_indices(VOLUME, rank, world)
    zs = np.ascontiguousarray(ax[zidx])
    nz_local = int(zidx.size)
    samples_local = VOLUME * VOLUME * nz_local
    out = torch.empty((nz_local, VOLUME, VOLUME), dtype=torch.float32, device=f"cuda:{local_rank}")

    def step():
        noise.multiband3D_lattice(ax, ax, zs, scale, w, float(post), mode=wn.WN_EVAL_FAST, out=out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the clock sampler starts BEFORE the warm-up and rank 0 keeps stepping until its first sample has arrived: the
    # start-up of nvidia-smi (NVML initialisation, first query) can hold up kernel launches for a millisecond or two,
    # which is a fifth of the timed region at N=8 (50 steps x 0.15 ms); afterwards it only samples every 100 ms
    sampler = ClockSampler(local_rank, world)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    sampler_extra_steps = 0
    if rank == 0 and sampler.proc:
        t_wait = time.perf_counter()
        while not sampler.lines and time.perf_counter() - t_wait < 3.0:
            step()
            torch.cuda.synchronize()
            sampler_extra_steps += 1
    barrier()
    launches0 = ctx.kernel_launches
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record()
    for i in range(args.steps):
        step()
        ev[i + 1].record()
    barrier()
    elapsed_ms = ev[0].elapsed_time(ev[-1])
    per_launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    launches = ctx.kernel_launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["extra_warmup_steps_while_sampler_started"] = sampler_extra_steps
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
    per_rank_ms = [elapsed_ms / args.steps]
    if world > 1:
        every = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(every, t)
        per_rank_ms = [float(x.item()) / args.steps for x in every]       # which rank, if any, is the straggler
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    total_samples = VOLUME ** 3 * args.steps
    value = total_samples / (elapsed_ms * 1e-3) / 1e9

    # --- dominant kernel alone, measured live: the same steps again with one CUDA-event pair around the main kernel of
    # every call (recorded by the library on the stream the kernel is launched on).  Kept out of the timed region above
    # because an event between the period-block chain and the main kernel removes their dependent-launch overlap.
    ctx.time_main_kernel(True)
    for _ in range(args.steps):
        step()
    main_ms = ctx.main_kernel_ms()
    main_kernel_ms = float(np.mean(main_ms)) if main_ms.size else None
    # In the steps above the period-block chain of call i+1 runs on the context's high-priority side stream WHILE the
    # main kernel of call i executes and takes SM slots from it, so that event pair spans both.  The kernel by itself:
    # the same steps with the chain kept on the compute stream (the knob is read per call).
    os.environ["WN_SIDE_MAX_LOG2"] = "20"
    for _ in range(2):
        step()
    ctx.main_kernel_ms()
    for _ in range(args.steps):
        step()
    alone_ms = ctx.main_kernel_ms()
    del os.environ["WN_SIDE_MAX_LOG2"]
    ctx.time_main_kernel(False)
    main_kernel_alone_ms = float(np.mean(alone_ms)) if alone_ms.size else None

    # --- the general path: the same five bands on a lattice that is NOT commensurate with the tile (base range 4.1
    # instead of 4: no band repeats, nothing folds, no replicas), i.e. the honest per-sample WMultibandNoise cost
    unfolded = None
    if rank == 0 and world == 1:
        axu = (np.arange(VOLUME, dtype=np.float32) / np.float32(VOLUME)) * np.float32(4.1)
        nzu = 256
        outu = torch.empty((nzu, VOLUME, VOLUME), dtype=torch.float32, device=f"cuda:{local_rank}")
        for _ in range(2):
            noise.multiband3D_lattice(axu, axu, axu[:nzu], scale, w, float(post), mode=wn.WN_EVAL_FAST, out=outu)
        torch.cuda.synchronize()
        ua, ub = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ua.record()
        for _ in range(3):
            noise.multiband3D_lattice(axu, axu, axu[:nzu], scale, w, float(post), mode=wn.WN_EVAL_FAST, out=outu)
        ub.record()
        torch.cuda.synchronize()
        u_ms = ua.elapsed_time(ub) / 3
        u_rate = VOLUME * VOLUME * nzu / (u_ms * 1e-3) / 1e9
        xp = extra_peaks()
        fp32_peak = xp.get("fp32_fma_tflops")
        u_tflops = u_rate * 1e9 * FLOP_PER_SAMPLE / 1e12
        del outu
        unfolded = {"workload": f"same bands, lattice {VOLUME}x{VOLUME}x{nzu} with base range 4.1 (not commensurate with "
                                "the tile: every band evaluated per sample)",
                    "value": u_rate, "unit": UNIT, "ms_per_call": u_ms,
                    "roofline": {"bound": "fp32", "flop_per_sample": FLOP_PER_SAMPLE, "achieved": u_tflops,
                                 "peak": fp32_peak, "unit": "TFLOP/s",
                                 "frac": (u_tflops / fp32_peak) if fp32_peak else None,
                                 "peak_source": "profiles/r2_peaks.json fp32_fma_tflops (dependent FFMA chains, measured)"}}

    # --- e2e: the public host-buffer call (axes H2D, result D2H into pinned host memory inside the timed region)
    e2e_steps = min(args.steps, 5)
    host_out = torch.empty((nz_local, VOLUME, VOLUME), dtype=torch.float32, pin_memory=True)

    def e2e_step():
        noise.multiband3D_lattice(ax, ax, zs, scale, w, float(post), mode=wn.WN_EVAL_FAST, out=host_out)

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = VOLUME ** 3 * e2e_steps / float(t.item()) / 1e9
    h2d_bytes = int((ax.nbytes * 2 + zs.nbytes) + scale.nbytes + w.nbytes)
    d2h_bytes = int(samples_local * 4)
    # sanity: the e2e result equals the device-resident result (one slice of every 32-slice staging chunk)
    same = all(bool(torch.equal(host_out[k], out[k].cpu())) for k in range(0, nz_local, 32))
    del host_out

    if rank != 0:
        return
    hbm_peak, peak_src = peaks()
    xp = extra_peaks()
    med_step_ms = float(np.median(per_launch_ms))
    alg_bytes = samples_local * OUT_BYTES_PER_SAMPLE + TILE_N ** 3 * 4
    kern_ms = main_kernel_ms if main_kernel_ms else med_step_ms
    achieved_gbs = alg_bytes / (kern_ms * 1e-3) / 1e9
    step_gbs = alg_bytes / (med_step_ms * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "kernel": "k_mb3d_rep (main kernel of the multiband lattice call: the launch that writes the output)",
        "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
        "peak_source": peak_src,
        "traffic": profiled_traffic_bytes() if world == 1 else None,
        "traffic_source": TRAFFIC_CSV + " (ncu --set full, main-kernel launch at N=1 bench size)",
        "algorithmic_bytes_per_launch": alg_bytes,
        "launch_ms": kern_ms,
        "launch_ms_source": "mean over the steps of the CUDA-event pair the library records around the main kernel; the "
                            "period-block chain of the next call runs on a side stream during that interval",
        "alone": None if not main_kernel_alone_ms else {
            "launch_ms": main_kernel_alone_ms,
            "achieved": alg_bytes / (main_kernel_alone_ms * 1e-3) / 1e9,
            "frac": alg_bytes / (main_kernel_alone_ms * 1e-3) / 1e9 / hbm_peak,
            "note": "the same kernel with the chain kept on the compute stream (nothing else on the GPU during the event pair)"},
        "step": {"achieved": step_gbs, "frac": step_gbs / hbm_peak, "ms": med_step_ms,
                 "note": "all launches of one step (axis tables, period blocks, main kernel), median over the timed steps"},
        "hbm_write_stream_gbs": xp.get("hbm_write_gbs"),
        "fp32": {"flop_per_sample": FLOP_PER_SAMPLE,
                 "achieved_tflops": samples_local * FLOP_PER_SAMPLE / (med_step_ms * 1e-3) / 1e12,
                 "note": "SURVEY 8(d) algorithmic FLOP (5 bands x 95) per sample; four of the five bands are periodic on this lattice and are evaluated once per period, so the main kernel is bound by the 4 B/sample output stream"},
    }
    # tile-gen ms at n=128 (second half of BASELINE's metric)
    tg = wn.WaveletNoise(TILE_N, SEED, ctx)
    t0 = time.perf_counter()
    R = tg.gaussian_field(TILE_N ** 3)
    fill_ms = (time.perf_counter() - t0) * 1e3
    Rd = torch.from_numpy(R).cuda()
    for _ in range(3):
        tg.generateNoiseTile3D(field=Rd)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        tg.generateNoiseTile3D(field=Rd)          # device-resident Gaussian field -> tile (3 filter passes + padded replica)
    b.record()
    torch.cuda.synchronize()
    filters_ms = a.elapsed_time(b) / 20
    h2d = []
    for _ in range(3):
        t0 = time.perf_counter()
        tg.generateNoiseTile3D(field=R)           # host field: + H2D of 8 MiB and a stream sync
        h2d.append((time.perf_counter() - t0) * 1e3)
    with_h2d_ms = min(h2d)
    seeded = []
    for _ in range(5):
        t0 = time.perf_counter()
        tg.generate_seeded(3)                      # seed -> tile through wn_tile_build_seeded (fill + filters on the GPU)
        seeded.append((time.perf_counter() - t0) * 1e3)
    seeded_ms = min(seeded)

    cpu = None
    if prev_affinity is not None:
        os.sched_setaffinity(0, prev_affinity)               # the CPU baseline uses every host core again
    if world == 1 and not args.no_cpu_baseline:
        cal = cpu_reference_rate(ax, scale, w, post, target_seconds=12.0)
        cpu = {k: cal[k] for k in ("value", "unit", "cores", "kind", "sample")}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": elapsed_ms / args.steps, "per_rank_ms_per_step": per_rank_ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "WMultibandNoise 1024^3 bands 4-8 weighted 2^-(b-4), tile n=128 seed 12345 "
                               "(BASELINE config 3), z block-cyclic sharded (32-slice chunks)",
                   "tile_n": TILE_N, "bands": [4, 8], "volume": [VOLUME] * 3, "slab_per_gpu": [VOLUME, VOLUME, nz_local],
                   "parallelism": f"z block-cyclic x{world}, no data-path collective",
                   "l2": "each step writes 4 GiB/N of fresh output (>> 126 MB L2); the 8 MiB tile is L2-resident by design"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "steps": e2e_steps, "matches_device_result": same, "numa_bound": prev_affinity is not None,
                "ceiling": {"d2h_pinned_gbs_one_gpu": xp.get("d2h_pinned_gbs"),
                            "value": (xp["d2h_pinned_gbs"] / 4.0 * min(world, 1)) if xp.get("d2h_pinned_gbs") else None,
                            "note": "4 bytes per sample cross PCIe: pinned D2H bandwidth / 4 (one GPU; profiles/r2_peaks.json)"}},
        "unfolded": unfolded,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "tile_gen_ms_n128": {"filters_device_only": filters_ms, "with_h2d": with_h2d_ms,
                             "host_gaussian_fill": fill_ms, "seeded_build_total": seeded_ms,
                             "first_build_incl_broadcast": tile_total_ms, "first_build_kernels": tile_kernel_ms},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        # NCCL prints its version banner on stdout at communicator creation: keep stdout for the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
