"""-m gpu: one map sharded over W ranks by block ownership == the single-GPU map, bit for bit.

The shards exchange nothing while integrating, so the W engines of a W-way sharded map can all live
on ONE GPU for the correctness test (this also runs on a single-GPU box); the multi-process NCCL
test below exercises the gather / replica plumbing when two GPUs are visible."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import voxblox_b200 as vb
from voxblox_b200 import scenes, sharded

pytestmark = pytest.mark.gpu


def _scans():
    return scenes.c3_room_sequence(n_scans=3, width=160, height=120)


def _single(kind, **cfg_kw):
    cfg = vb.TsdfIntegratorConfig(default_truncation_distance=0.4, integrator_threads=1, **cfg_kw)
    layer = vb.Layer(0.1, 16, engine_options=vb.EngineOptions(max_blocks=4096, max_updates_per_pass=1 << 22))
    integ = vb.TsdfIntegratorFactory.create(kind, cfg, layer)
    for s in _scans():
        integ.integratePointCloud((s[2], s[3]), s[0], s[1])
    return layer, integ


@pytest.mark.parametrize("kind,cfg_kw", [(2, {}), (1, {}), (2, dict(enable_anti_grazing=1))])
@pytest.mark.parametrize("world", [2, 3, 8])
def test_union_of_shards_is_the_single_gpu_map(kind, cfg_kw, world):
    full, _ = _single(kind, **cfg_kw)
    full_blocks = full.blocks()
    seen = {}
    total_updates = 0
    for rank in range(world):
        cfg = vb.TsdfIntegratorConfig(default_truncation_distance=0.4, integrator_threads=1, **cfg_kw)
        layer = vb.Layer(0.1, 16, engine_options=sharded.shard_options(rank, world, max_blocks=4096,
                                                                      max_updates_per_pass=1 << 22))
        integ = vb.TsdfIntegratorFactory.create(kind, cfg, layer)
        for s in _scans():
            integ.integratePointCloud((s[2], s[3]), s[0], s[1])
        blocks = layer.blocks()
        idx = np.array(sorted(blocks), np.int32).reshape(-1, 3)
        if len(idx):
            assert (sharded.block_owner(idx, world) == rank).all(), "a rank created a block it does not own"
        for k, v in blocks.items():
            assert k not in seen
            seen[k] = v
    assert sorted(seen) == sorted(full_blocks)
    for k in full_blocks:
        assert seen[k].tobytes() == full_blocks[k].tobytes(), k


def test_pipelined_shard_equals_synchronous_shard():
    cfg = vb.TsdfIntegratorConfig(default_truncation_distance=0.4, integrator_threads=1)
    opts = lambda: sharded.shard_options(1, 4, max_blocks=4096, max_updates_per_pass=1 << 22)
    la, ls = vb.Layer(0.1, 16, engine_options=opts()), vb.Layer(0.1, 16, engine_options=opts())
    ia, isync = vb.TsdfIntegratorFactory.create(2, cfg, la), vb.TsdfIntegratorFactory.create(2, cfg, ls)
    keep = []
    for s in scenes.c3_room_sequence(n_scans=6, width=160, height=120):
        isync.integratePointCloud((s[2], s[3]), s[0], s[1])
        p, c = np.ascontiguousarray(s[0]), np.ascontiguousarray(s[1])
        keep.append((p, c))
        ia.integratePointCloudAsync((s[2], s[3]), p, c)
    la.sync()
    a, b = la.blocks(), ls.blocks()
    assert sorted(a) == sorted(b) and all(a[k].tobytes() == b[k].tobytes() for k in a)


# ---------------------------------------------------------------- two processes, NCCL
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        cfg = vb.TsdfIntegratorConfig(default_truncation_distance=0.4, integrator_threads=1)
        layer = vb.Layer(0.1, 16, engine_options=sharded.shard_options(rank, world, device=rank, max_blocks=4096,
                                                                      max_updates_per_pass=1 << 22))
        integ = vb.TsdfIntegratorFactory.create(2, cfg, layer)
        sl = sharded.ShardedLayer(layer)
        for s in _scans():
            integ.integratePointCloud((s[2], s[3]), s[0], s[1])
            sl.sync_replicas()
        # after sync_replicas every rank reads the whole map from its own GPU
        blocks = layer.blocks()
        gathered = sl.gather()
        assert sorted(blocks) == sorted(gathered)
        idx = np.array(sorted(blocks), np.int32)
        vox = np.stack([blocks[tuple(i)] for i in idx])
        np.savez(out + f".{rank}.npz", idx=idx, vox=vox.view(np.uint8))
    finally:
        dist.destroy_process_group()


def test_replicas_after_sync_equal_single_gpu_world2(tmp_path):
    world = 2
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    full, _ = _single(2)
    fb = full.blocks()
    out = str(tmp_path / "shard")
    port = _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    for r in range(world):
        z = np.load(out + f".{r}.npz")
        assert [tuple(i) for i in z["idx"].tolist()] == sorted(fb)
        ref = np.stack([fb[k] for k in sorted(fb)]).view(np.uint8)
        assert z["vox"].tobytes() == ref.tobytes(), f"rank {r} replica differs from the single-GPU map"
