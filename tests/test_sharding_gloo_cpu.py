"""World-size-2 gloo test of the multi-GPU host logic (row sharding + all-gather),
with the CPU oracle standing in for the kernel.  No GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from horayzon_b200 import sharding, synthetic as syn


def test_row_shards_cover_domain():
    for rows, world in ((750, 8), (24001, 8), (5, 8), (97, 2), (0, 4), (1199, 4)):
        sh = sharding.row_shards(rows, world)
        assert len(sh) == world and sh[0][0] == 0 and sh[-1][1] == rows
        assert all(a[1] == b[0] for a, b in zip(sh, sh[1:]))
        per = sharding.padded_rows(rows, world)
        assert all(e - b <= per for b, e in sh)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n, K, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    c = syn.make_config("cfg1", n=n)
    ny, nx = c["ny"], c["nx"]
    b, e = sharding.row_shards(ny, world)[rank]
    per = sharding.padded_rows(ny, world)
    local = torch.zeros((per, nx, K), dtype=torch.float32)
    if e > b:  # shard = a sub-domain: shifted offset and sliced per-cell arrays
        h, _ = oracle.horizon_gridded(c["vert_grid"], n, n, c["vec_norm"][b:e], c["vec_north"][b:e],
                                      c["offset_0"] + b, c["offset_1"], c["dist_search"], azim_num=K)
        local[: e - b] = torch.from_numpy(h)
    full = sharding.allgather_rows(local, ny, dist)
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), full.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_allgather_matches_single_process(tmp_path):
    import oracle
    n, K, world = 45, 12, 2  # 13 inner rows: uneven split 7 + 6 with one padding row
    mp.spawn(_worker, args=(world, _free_port(), n, K, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "gathered.npy")
    c = syn.make_config("cfg1", n=n)
    want, _ = oracle.horizon_gridded(c["vert_grid"], n, n, c["vec_norm"], c["vec_north"], c["offset_0"],
                                     c["offset_1"], c["dist_search"], azim_num=K)
    assert got.shape == want.shape and np.array_equal(got, want)


def test_block_shards_cover_domain_and_unpack():
    """Block-interleaved sharding (round 2): shard r owns the 4-row blocks b with b % world == r; packed shards,
    concatenated like an all-gather concatenates them, unpack to the domain order."""
    for rows, world in ((1199, 8), (5998, 8), (24001, 8), (13, 2), (7, 3), (4, 4), (1, 2), (3000, 1)):
        per = sharding.padded_block_rows(rows, world)
        seen = []
        bufs = []
        for r in range(world):
            idx = sharding.shard_row_indices(rows, r, world)
            assert len(idx) == sharding.shard_block_rows(rows, r, world) <= per
            seen += [i for i in idx if i >= 0]
            b = np.full(per, -1); b[:len(idx)] = idx
            bufs.append(b)
        assert sorted(seen) == list(range(rows))
        assert np.array_equal(sharding.unpack_blocks(np.concatenate(bufs), rows, world), np.arange(rows))


def _worker_blocks(rank, world, port, n, K, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    c = syn.make_config("cfg1", n=n)
    ny, nx = c["ny"], c["nx"]
    per = sharding.padded_block_rows(ny, world)
    idx = sharding.shard_row_indices(ny, rank, world)
    rows = [int(i) for i in idx if i >= 0]
    local = torch.zeros((per, nx, K), dtype=torch.float32)
    if rows:   # the oracle stands in for the kernel's packed mode: this shard's rows back to back
        sc = oracle.Scene(c["vert_grid"], n, n)
        h = sc.horizon_rows(rows, c["vec_norm"], c["vec_north"], c["offset_0"], c["offset_1"], c["dist_search"], azim_num=K)
        sc.close()
        local[:len(rows)] = torch.from_numpy(h)     # real rows are a prefix of the packed buffer (padding only in the last block)
    gathered = torch.empty((world * per, nx, K), dtype=torch.float32)
    dist.all_gather_into_tensor(gathered, local)
    full = sharding.unpack_blocks(gathered, ny, world)
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered_blocks.npy"), full.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_block_allgather_matches_single_process(tmp_path):
    import oracle
    n, K, world = 47, 12, 2  # 15 inner rows = 4 blocks (the last one ragged): shard 0 blocks 0, 2; shard 1 blocks 1, 3
    mp.spawn(_worker_blocks, args=(world, _free_port(), n, K, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "gathered_blocks.npy")
    c = syn.make_config("cfg1", n=n)
    want, _ = oracle.horizon_gridded(c["vert_grid"], n, n, c["vec_norm"], c["vec_north"], c["offset_0"],
                                     c["offset_1"], c["dist_search"], azim_num=K)
    assert got.shape == want.shape and np.array_equal(got, want)
