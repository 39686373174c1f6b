"""Multi-GPU plumbing: one process per GPU, samples sharded with no data-path exchange.

The path shards naturally: every sample reads only the replicated, read-only tile,
so a volume is cut into contiguous z-slabs and an image into contiguous row-bands; the only collectives are
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


def dist_info():
    """(rank, world, initialised) from torch.distributed, (0, 1, False) when not running distributed."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size(), True
    except ImportError:
        pass
    return 0, 1, False


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
