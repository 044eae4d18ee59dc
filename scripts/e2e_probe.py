"""Scratch probe: host->device copy rate seen by the engine for different host buffers."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import voxblox_b200 as vb
from voxblox_b200 import scenes
s = scenes.c3_room_scan(5)
cfg = vb.TsdfIntegratorConfig(default_truncation_distance=0.2)
layer = vb.Layer(0.05, 16, engine_options=vb.EngineOptions(max_blocks=4096, max_points_per_scan=1 << 19, max_updates_per_pass=1 << 22))
integ = vb.TsdfIntegratorFactory.create("merged", cfg, layer)
pts = s[0]
tp = torch.from_numpy(pts).pin_memory()
own = layer.hostBuffer(pts.shape, np.float32); own[...] = pts
for name, arr in (("pageable numpy", pts), ("torch pinned", tp.numpy()), ("vbx_host_alloc", own)):
    ms = [layer.hostCopyMs(arr) for _ in range(8)]
    print(f"{name:16s} {arr.nbytes/1e6:.2f} MB  ms {np.round(ms,3).tolist()}  best GB/s {arr.nbytes/1e6/min(ms):.1f}")
for name, (a, b) in (("torch pinned", (tp.numpy(), torch.from_numpy(s[1]).pin_memory().numpy())), ):
    w = []
    for _ in range(10):
        t0 = time.perf_counter(); integ.integratePointCloud((s[2], s[3]), a, b); w.append((time.perf_counter()-t0)*1e3)
    print(name, "integrate wall ms", np.round(w,3).tolist(), "device ms", integ.lastDeviceMs())
