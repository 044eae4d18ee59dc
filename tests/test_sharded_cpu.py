"""Host-side logic of the block-ownership sharding (voxblox_b200/sharded.py) on CPU: the ownership
function, and the variable-size block all-gather under gloo with world_size 2."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from voxblox_b200 import sharded
from voxblox_b200.api import TSDF_DTYPE


def test_block_owner_partitions_space():
    rng = np.random.default_rng(0)
    idx = rng.integers(-50, 50, size=(5000, 3)).astype(np.int32)
    for world in (1, 2, 3, 4, 8):
        o = sharded.block_owner(idx, world)
        assert o.min() >= 0 and o.max() < world
        # every rank owns a fair share, and face neighbours along x never share an owner (world > 1)
        counts = np.bincount(o, minlength=world)
        assert counts.min() > 0.5 * len(idx) / world
        if world > 1:
            nb = idx.copy()
            nb[:, 0] += 1
            assert (sharded.block_owner(nb, world) != o).all()
    # the 2 x 2 x 2 brick of 8 ranks: the eight blocks around a corner all have different owners
    corner = np.array([[x, y, z] for z in (-1, 0) for y in (-1, 0) for x in (-1, 0)], np.int32)
    assert sorted(sharded.block_owner(corner, 8).tolist()) == list(range(8))
    # negative coordinates: non-negative modulo like the device's block_owner()
    assert sharded.block_owner([[-1, 0, 0]], 8)[0] == 7 and sharded.block_owner([[0, 0, -1]], 8)[0] == 4


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _blocks_of(rank):
    rng = np.random.default_rng(100 + rank)
    m = 3 + 4 * rank  # ranks hold different numbers of blocks
    idx = rng.integers(-9, 9, size=(m, 3)).astype(np.int32)
    vox = np.zeros((m, 8), dtype=TSDF_DTYPE)
    vox["distance"] = rng.normal(size=(m, 8)).astype(np.float32)
    vox["weight"] = rank + 1
    vox["color"] = rng.integers(0, 255, size=(m, 8, 4))
    return idx, vox


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        idx, vox = _blocks_of(rank)
        all_idx, all_vox, counts = sharded.all_gather_blocks(idx, vox)
        assert counts.tolist() == [3 + 4 * r for r in range(world)]
        at = 0
        for r in range(world):
            ri, rv = _blocks_of(r)
            assert (all_idx[at:at + len(ri)] == ri).all()
            assert all_vox[at:at + len(ri)].tobytes() == rv.tobytes()
            at += len(ri)
        # a rank with nothing to send
        e_idx, e_vox, e_counts = sharded.all_gather_blocks(idx[:0] if rank == 0 else idx, vox[:0] if rank == 0 else vox)
        assert e_counts[0] == 0 and len(e_idx) == sum(e_counts)
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def test_block_all_gather_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Array("i", world)
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert list(out) == [1, 1]
