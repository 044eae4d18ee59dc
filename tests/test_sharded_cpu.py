"""world_size-2 gloo test (CPU) of the multi-GPU host logic: ray-slot partition, count exchange,
prefix-of-pack all-gather and the global record order the device back half relies on."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from voxblox_b200 import sharded


def test_slot_ranges_partition_the_scan():
    for n in (0, 1, 7, 1024, 261119, 307200):
        for world in (1, 2, 3, 4, 8):
            r = [sharded.slot_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert all(hi - lo <= (n + world - 1) // world for lo, hi in r)


def test_exchange_size_and_prefix():
    assert sharded.exchange_bytes(512, [0, 0]) == 512
    assert sharded.exchange_bytes(512, [1, 4097]) == 512 + 16 * 8192
    assert sharded.record_prefix([3, 0, 5]) == [0, 3, 3, 8]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 1000
        lo, hi = sharded.slot_range(n, rank, world)
        rng = np.random.default_rng(rank)
        count = 5000 + 3000 * rank  # ranks hold different numbers of records
        off_records = 256
        cap = 16384
        pack = torch.zeros(off_records + cap * 16, dtype=torch.uint8)
        rec = np.zeros((count, 4), dtype=np.uint32)
        rec[:, 0] = rng.integers(0, 2 ** 32, count)
        rec[:, 2] = np.sort(rng.integers(lo, hi, count))  # ray slots of this rank, ascending
        pack[off_records:off_records + count * 16] = torch.from_numpy(rec.view(np.uint8).reshape(-1))
        pack[:8] = torch.from_numpy(np.array([rank + 1], dtype=np.int64).view(np.uint8))  # fake ray table
        counts = sharded.gather_counts(count)
        assert counts == [5000 + 3000 * r for r in range(world)]
        nbytes = sharded.exchange_bytes(off_records, counts)
        gathered = torch.zeros(world * pack.numel(), dtype=torch.uint8)
        got = sharded.gather_packs(pack, nbytes, gathered).numpy().reshape(world, nbytes)
        start = sharded.record_prefix(counts)
        slots = []
        for r in range(world):
            assert int(got[r, :8].view(np.int64)[0]) == r + 1
            rr = got[r, off_records:off_records + counts[r] * 16].view(np.uint32).reshape(-1, 4)
            slots.append(rr[:, 2])
            if r == rank:
                assert (rr == rec).all()
        allslots = np.concatenate(slots)
        assert len(allslots) == start[-1]
        # rank order == ascending ray slots: the order in which the back half applies updates
        assert (np.diff(allslots.astype(np.int64)) >= 0).all()
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def test_exchange_plumbing_world2_gloo():
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
