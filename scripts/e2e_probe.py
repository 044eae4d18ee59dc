"""Scratch probe: where does the host-buffer call spend its time?"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import voxblox_b200 as vb
from voxblox_b200 import scenes
scans = scenes.generate_parallel(scenes.c3_room_scan, range(12))
cfg = vb.TsdfIntegratorConfig(default_truncation_distance=0.2)
opts = vb.EngineOptions(max_blocks=16384, max_points_per_scan=1 << 19, max_updates_per_pass=1 << 24)
for mode in ("device", "pinned", "pageable"):
    layer = vb.Layer(0.05, 16, engine_options=opts)
    integ = vb.TsdfIntegratorFactory.create("merged", cfg, layer)
    bufs = []
    for s in scans:
        if mode == "device":
            bufs.append((torch.from_numpy(s[0]).cuda(), torch.from_numpy(s[1]).cuda()))
        elif mode == "pinned":
            bufs.append((torch.from_numpy(s[0]).pin_memory(), torch.from_numpy(s[1]).pin_memory()))
        else:
            bufs.append((s[0], s[1]))
    torch.cuda.synchronize()
    walls, devs = [], []
    for i, s in enumerate(scans):
        t0 = time.perf_counter()
        if mode == "device":
            integ.integratePointCloudDevice((s[2], s[3]), bufs[i][0].data_ptr(), bufs[i][1].data_ptr(), s[0].shape[0])
        elif mode == "pinned":
            integ.integratePointCloud((s[2], s[3]), bufs[i][0].numpy(), bufs[i][1].numpy())
        else:
            integ.integratePointCloud((s[2], s[3]), bufs[i][0], bufs[i][1])
        walls.append((time.perf_counter() - t0) * 1e3)
        devs.append(integ.lastDeviceMs())
    print(mode, "wall ms", np.round(walls[3:], 3).tolist(), "device ms", np.round(devs[3:], 3).tolist())
