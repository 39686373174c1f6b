"""Multi-GPU plumbing: one process per GPU, samples sharded with no data-path exchange.

The path shards naturally: every sample reads only the replicated, read-only tile,
so a volume is cut into z-slabs (contiguous, or block-cyclic chunks -- see cyclic_slab_indices) and an image into
contiguous row-bands; the only collectives are
  (1) one broadcast of the finished tile from rank 0 (8 MiB at n=128) -- NCCL over NVLink on GPUs, and
  (2) an optional gather of the output shards to rank 0 when a file has to be written.
`torch.distributed` is plumbing only; with backend "gloo" the same code runs on CPU buffers, which is how
tests/test_sharding_cpu.py exercises it at world_size 2 without GPUs.
"""
import numpy as np


def slab_range(total, rank, world):
    """Contiguous [begin, end) of `total` units for `rank`; remainders go to the lowest ranks."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(int(total), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def all_slabs(total, world):
    return [slab_range(total, r, world) for r in range(world)]


CYCLIC_CHUNK = 32       # z slices per chunk = the z extent of one CTA brick of the fast lattice kernel


def cyclic_slab_indices(total, rank, world, chunk=CYCLIC_CHUNK):
    """Block-cyclic sharding of an axis: chunks of `chunk` consecutive units are dealt round-robin, rank r owns chunks
    r, r+world, r+2*world, ...  Returns the owned indices (int64, ascending).

    Why not contiguous slabs: the fast lattice kernel evaluates a band whose samples repeat with the tile period only
    once per period (periodic folding).  A contiguous slab of 1024/8 slices is shorter than the
    z periods of the low bands, so that saving would be lost as the rank count grows; a block-cyclic slab keeps slices
    that are congruent modulo those periods on the same rank, so every rank folds the same bands as the single-GPU run
    and its work stays 1/world of it.  Each sample still depends on the replicated tile only: no exchange."""
    if world <= 0 or not (0 <= rank < world) or chunk <= 0:
        raise ValueError(f"bad rank/world/chunk {rank}/{world}/{chunk}")
    total = int(total)
    nchunks = (total + chunk - 1) // chunk
    parts = [np.arange(c * chunk, min((c + 1) * chunk, total), dtype=np.int64) for c in range(rank, nchunks, world)]
    return np.concatenate(parts) if parts else np.zeros(0, dtype=np.int64)


def dist_info():
    """(rank, world, initialised) from torch.distributed, (0, 1, False) when not running distributed."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size(), True
    except ImportError:
        pass
    return 0, 1, False


def bind_to_gpu_numa(device_index):
    """Restrict this process to the CPUs NVML reports as local to CUDA device `device_index`, so host staging buffers
    allocated afterwards (pinned output arrays) land on the NUMA node next to that GPU's PCIe root.  With one process
    per GPU and no binding, every rank's buffers end up on the launcher's node and half of the GPUs copy across the
    socket interconnect.  Returns the previous affinity set (pass it to os.sched_setaffinity(0, ...) to undo), or None
    when NVML / the device is unavailable (nothing changed)."""
    import os
    try:
        import pynvml
        import torch
        prev = os.sched_getaffinity(0)
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(device_index)
        bus = "%08x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return prev
    except Exception:                                        # noqa: BLE001  (best effort: placement only)
        return None


def broadcast_tile(buffer, src=0):
    """Replicate the tile: `buffer` is a torch tensor (CUDA view of the tile for NCCL, CPU tensor for gloo)
    holding the coefficients on `src` and uninitialised storage elsewhere."""
    import torch.distributed as dist
    dist.broadcast(buffer, src=src)
    return buffer


def replicate_noise(noise, dims, build_on_root):
    """Rank 0 builds the tile (build_on_root(noise)), every other rank allocates an empty replica and
    receives the coefficients with one broadcast.  Returns `noise` (built on every rank)."""
    rank, world, on = dist_info()
    if not on or world == 1:
        build_on_root(noise)
        return noise
    import torch
    if rank == 0:
        build_on_root(noise)
    else:
        noise.allocate(dims)
    noise.ctx.synchronize()
    view = noise.device_tensor()
    torch.cuda.current_stream().synchronize()
    broadcast_tile(view, 0)
    torch.cuda.current_stream().synchronize()
    if rank != 0:
        noise.mark_built()
    return noise


def gather_slabs(local, total, axis_len_other, rank=None, world=None, dst=0):
    """Gather contiguous slabs (flattened float32 tensors, slab r = slab_range(total, r, world) x axis_len_other
    floats) to `dst`.  Returns the assembled tensor on dst, None elsewhere.  Used only for file output."""
    import torch
    import torch.distributed as dist
    if rank is None:
        rank, world, _ = dist_info()
    sizes = [(e - b) * axis_len_other for b, e in all_slabs(total, world)]
    pad = max(sizes)
    send = torch.zeros(pad, dtype=local.dtype, device=local.device)
    send[: local.numel()] = local.reshape(-1)
    bufs = [torch.empty(pad, dtype=local.dtype, device=local.device) for _ in range(world)] if rank == dst else None
    if dist.get_backend() == "nccl":
        # nccl has no gather-to-one with a list on every version: use all_gather (output is small relative to compute)
        bufs = [torch.empty(pad, dtype=local.dtype, device=local.device) for _ in range(world)]
        dist.all_gather(bufs, send)
    else:
        dist.gather(send, bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)])


def gather_cyclic(local, total, axis_len_other, rank=None, world=None, dst=0, chunk=CYCLIC_CHUNK):
    """gather_slabs for block-cyclic shards: rank r holds the slices cyclic_slab_indices(total, r, world, chunk) in
    ascending order; dst gets them back in axis order."""
    import torch
    import torch.distributed as dist
    if rank is None:
        rank, world, _ = dist_info()
    idx = [cyclic_slab_indices(total, r, world, chunk) for r in range(world)]
    pad = max(len(i) for i in idx) * axis_len_other
    send = torch.zeros(pad, dtype=local.dtype, device=local.device)
    send[: local.numel()] = local.reshape(-1)
    if dist.get_backend() == "nccl":
        bufs = [torch.empty(pad, dtype=local.dtype, device=local.device) for _ in range(world)]
        dist.all_gather(bufs, send)
    else:
        bufs = [torch.empty(pad, dtype=local.dtype, device=local.device) for _ in range(world)] if rank == dst else None
        dist.gather(send, bufs, dst=dst)
    if rank != dst:
        return None
    full = torch.empty((total, axis_len_other), dtype=local.dtype, device=local.device)
    for i, b in zip(idx, bufs):
        full[torch.from_numpy(i).to(local.device)] = b[: len(i) * axis_len_other].reshape(len(i), axis_len_other)
    return full.reshape(-1)


def lattice_axes_config3(size=1024, base_range=4.0):
    """Config 3 coordinates: p = (idx/size)*base_range, float32, like the reference's pixel loops."""
    return (np.arange(size, dtype=np.float32) / np.float32(size)) * np.float32(base_range)


def config3_bands(first=4, last=8, variance=0.18402):
    """bands b=first..last: q_b = 2 p 2^b, w_b = 2^-(b-first), post = 1/sqrt(sum w^2 * var) (paper App. 2)."""
    b = np.arange(first, last + 1)
    scale = (2.0 * 2.0 ** b).astype(np.float32)
    w = (2.0 ** -(b - first).astype(np.float64)).astype(np.float32)
    post = np.float32(1.0) / np.sqrt(np.float32((w * w).sum()) * np.float32(variance))
    return scale, w, np.float32(post)
