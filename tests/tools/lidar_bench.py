"""C5: MergedTsdfIntegrator on 2048x128 spinning-LiDAR scans, 0.05 m voxels, const weight,
max_ray_length 10 m.  Single GPU, or ray-range sharded when launched under torchrun."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import voxblox_b200 as vb
from voxblox_b200 import scenes, sharded
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
scans = scenes.generate_parallel(scenes.c5_lidar_scan, range(N))
kw = dict(default_truncation_distance=0.2, max_ray_length_m=10.0, use_const_weight=1)
cfg = vb.TsdfIntegratorConfig(**kw)
opts = vb.EngineOptions(device=local, max_blocks=32768, max_points_per_scan=1 << 19, max_updates_per_pass=1 << 25, rank=rank, world_size=world)
layer = vb.Layer(0.05, 16, engine_options=opts)
integ = vb.TsdfIntegratorFactory.create("merged", cfg, layer)
d = [(torch.from_numpy(s[0]).cuda(), torch.from_numpy(s[1]).cuda()) for s in scans]
sh = sharded.ShardedTsdfIntegrator(integ, record_capacity=(1 << 25) // world) if world > 1 else None
ms = []
layer.setStageProfiling(True)
for i, s in enumerate(scans):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    if sh:
        sh.integratePointCloudDevice((s[2], s[3]), d[i][0].data_ptr(), d[i][1].data_ptr(), s[0].shape[0])
    else:
        integ.integratePointCloudDevice((s[2], s[3]), d[i][0].data_ptr(), d[i][1].data_ptr(), s[0].shape[0])
    torch.cuda.synchronize(); ms.append((time.perf_counter() - t0) * 1e3)
if rank == 0:
    c = integ.counters()
    print(f"world {world}: ms/scan (wall) {np.round(ms, 2).tolist()} mean(after 2) {np.mean(ms[2:]):.3f}  points {scans[0][0].shape[0]} counters {c} blocks {layer.getNumberOfAllocatedBlocks()}")
    print({k: (round(v[0] / max(1, v[1]), 4), v[1]) for k, v in layer.stageMs().items() if v[1]})
    if world == 1 and len(sys.argv) > 2:
        from oracle import pyoracle as po
        om = po.OracleMap(po.OracleLib("reference" if po.available("reference") else "port"), po.TsdfConfig(integrator_threads=int(sys.argv[2]), **kw), 0.05, 16)
        rt = []
        for s in scans[:4]:
            om.integrate(2, s); rt.append(om.last_seconds() * 1e3)
        print("cpu reference ms/scan", np.round(rt, 1).tolist(), "threads", sys.argv[2])
if world > 1:
    dist.destroy_process_group()
