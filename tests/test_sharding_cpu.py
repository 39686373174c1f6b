"""Host-side multi-rank logic on CPU: slab partitioning, tile broadcast and output gather over gloo, world_size 2."""
import os
import socket

import numpy as np
import pytest

import wnpkg


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_slab_ranges_cover_and_balance():
    sh = wnpkg.load_sub("sharding")
    for total, world in ((1024, 1), (1024, 2), (1024, 8), (1000, 3), (5, 8), (0, 4)):
        slabs = sh.all_slabs(total, world)
        assert slabs[0][0] == 0 and slabs[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(slabs, slabs[1:]))
        sizes = [e - b for b, e in slabs]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sh.slab_range(10, 3, 3)


def test_cyclic_indices_partition_the_axis():
    sh = wnpkg.load_sub("sharding")
    for total, world, chunk in ((1024, 8, 32), (1024, 2, 32), (1000, 3, 32), (70, 4, 8), (5, 8, 32), (0, 2, 32), (64, 1, 32)):
        parts = [sh.cyclic_slab_indices(total, r, world, chunk) for r in range(world)]
        allidx = np.sort(np.concatenate(parts))
        assert np.array_equal(allidx, np.arange(total))
        for part in parts:
            assert np.all(np.diff(part) > 0)
    # 1024 slices on 8 ranks: rank 3 owns chunks 3, 11, 19, 27
    got = sh.cyclic_slab_indices(1024, 3, 8)
    assert len(got) == 128 and got[0] == 96 and got[32] == 96 + 256 and got[-1] == 96 + 768 + 31
    with pytest.raises(ValueError):
        sh.cyclic_slab_indices(10, 2, 2)


def test_config3_parameters():
    sh = wnpkg.load_sub("sharding")
    ax = sh.lattice_axes_config3(1024)
    assert ax.dtype == np.float32 and ax[1] == np.float32(4.0 / 1024) and ax[-1] == np.float32(1023 * 4.0 / 1024)
    scale, w, post = sh.config3_bands(4, 8)
    assert list(scale) == [32, 64, 128, 256, 512] and list(w) == [1, .5, .25, .125, .0625]
    assert post == np.float32(1.0) / np.sqrt(np.float32(1.33203125) * np.float32(0.18402))


def _worker(rank, world, port, tmpdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import wnpkg as wp
    sh = wp.load_sub("sharding")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert sh.dist_info() == (rank, world, True)
        # (1) tile replication: rank 0 owns the coefficients, the others receive them with one broadcast
        n = 16
        tile = torch.arange(n ** 3, dtype=torch.float32) * 0.5 if rank == 0 else torch.empty(n ** 3, dtype=torch.float32)
        sh.broadcast_tile(tile, 0)
        assert torch.equal(tile, torch.arange(n ** 3, dtype=torch.float32) * 0.5)
        # (2) every rank produces its z-slab of a 10 x 6 x 7 volume from the replicated data only
        nx, ny, nz = 10, 6, 7
        b, e = sh.slab_range(nz, rank, world)
        full = torch.arange(nx * ny * nz, dtype=torch.float32).reshape(nz, ny, nx)
        local = full[b:e].clone()
        # (3) gather for file output
        got = sh.gather_slabs(local, nz, nx * ny, rank, world, 0)
        # (4) the same with block-cyclic shards (chunks of 2 slices dealt round-robin)
        idx = torch.from_numpy(sh.cyclic_slab_indices(nz, rank, world, chunk=2))
        got_c = sh.gather_cyclic(full[idx].clone(), nz, nx * ny, rank, world, 0, chunk=2)
        if rank == 0:
            assert torch.equal(got.reshape(nz, ny, nx), full)
            assert torch.equal(got_c.reshape(nz, ny, nx), full)
            open(os.path.join(tmpdir, "ok"), "w").write("1")
        else:
            assert got is None and got_c is None
    finally:
        dist.destroy_process_group()


def test_two_rank_broadcast_and_gather_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()
