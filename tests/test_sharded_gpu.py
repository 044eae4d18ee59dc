"""-m gpu: ray-range sharded integration == single-GPU integration, bit for bit."""
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


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _scans():
    return scenes.c3_room_sequence(n_scans=3, width=160, height=120)


def _reference_layer(kind):
    cfg = vb.TsdfIntegratorConfig(default_truncation_distance=0.4)
    layer = vb.Layer(0.1, 16, engine_options=vb.EngineOptions(max_blocks=4096, max_updates_per_pass=1 << 22))
    integ = vb.TsdfIntegratorFactory.create(kind, cfg, layer)
    for s in _scans():
        integ.integratePointCloud((s[2], s[3]), s[0], s[1])
    idx = layer.getAllAllocatedBlocks()
    return idx, layer.getBlocks(idx)


def _worker(rank, world, port, kind, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        cfg = vb.TsdfIntegratorConfig(default_truncation_distance=0.4)
        opts = vb.EngineOptions(device=rank, max_blocks=4096, max_updates_per_pass=1 << 23, rank=rank, world_size=world)
        layer = vb.Layer(0.1, 16, engine_options=opts)
        integ = vb.TsdfIntegratorFactory.create(kind, cfg, layer)
        sh = sharded.ShardedTsdfIntegrator(integ, record_capacity=1 << 21)
        for s in _scans():
            xyz = torch.from_numpy(s[0]).cuda()
            rgba = torch.from_numpy(s[1]).cuda()
            sh.integratePointCloudDevice((s[2], s[3]), xyz.data_ptr(), rgba.data_ptr(), s[0].shape[0])
        idx = layer.getAllAllocatedBlocks()
        vox, upd = layer.getBlocks(idx)
        np.savez(out + f".{rank}.npz", idx=idx, vox=vox.view(np.uint8), upd=upd)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind", [2, 1])
@pytest.mark.parametrize("world", [1, 2])
def test_sharded_equals_single_gpu(tmp_path, kind, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    idx0, (vox0, upd0) = _reference_layer(kind)
    out = str(tmp_path / "shard")
    port = _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    for r in range(world):
        z = np.load(out + f".{r}.npz")
        assert (z["idx"] == idx0).all()
        assert z["vox"].tobytes() == vox0.view(np.uint8).tobytes(), f"rank {r} replica differs from the single-GPU map"
        assert (z["upd"] == upd0).all()
