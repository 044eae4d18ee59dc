"""Pipelined pace (ms per scan, device clouds, timed from the first submission until the last scan is in the
map) against the number of front lanes and hand-off sets in use (VBX_ASYNC_LANES / VBX_ASYNC_SETS)."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import voxblox_b200 as vb
from voxblox_b200 import scenes

n_warm, n = 10, int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 200
scans = scenes.generate_parallel(scenes.c3_room_scan, range(60))
dev = torch.device("cuda", 0)
d_xyz = [torch.from_numpy(s[0]).to(dev) for s in scans]
d_rgba = [torch.from_numpy(s[1]).to(dev) for s in scans]
npts = [int(s[0].shape[0]) for s in scans]
out = []
CASES = [(6, 10, None), (4, 10, None), (3, 10, None), (2, 10, None), (6, 16, None), (8, 16, None), (4, 6, None)]
if "--grids" in sys.argv:   # VBX_GRID_SMS scales the grid of every integration kernel (default: the SM count)
    CASES = [(6, 10, g) for g in (37, 74, 111, 148, 222, 296)]
for lanes, sets, grid_sms in CASES:
    os.environ["VBX_ASYNC_LANES"], os.environ["VBX_ASYNC_SETS"] = str(lanes), str(sets)
    if grid_sms is None:
        os.environ.pop("VBX_GRID_SMS", None)
    else:
        os.environ["VBX_GRID_SMS"] = str(grid_sms)
    layer = vb.Layer(0.05, 16, engine_options=vb.EngineOptions(max_blocks=16384, max_points_per_scan=1 << 19, max_updates_per_pass=1 << 22))
    integ = vb.TsdfIntegratorFactory.create("merged", vb.TsdfIntegratorConfig(default_truncation_distance=0.2), layer)
    res = {}
    for steps in (20, n):
        for i in range(n_warm):
            k = i % 60
            integ.integratePointCloudAsync((scans[k][2], scans[k][3]), d_xyz[k].data_ptr(), d_rgba[k].data_ptr(), npts[k])
        layer.sync()
        t0 = time.perf_counter()
        for i in range(steps):
            k = (n_warm + i) % 60
            integ.integratePointCloudAsync((scans[k][2], scans[k][3]), d_xyz[k].data_ptr(), d_rgba[k].data_ptr(), npts[k])
        layer.sync()
        res[f"ms_per_scan_{steps}"] = round((time.perf_counter() - t0) * 1e3 / steps, 4)
    out.append({"lanes": lanes, "sets": sets, "grid_sms": grid_sms, **res})
    print(json.dumps(out[-1]), flush=True)
    del integ, layer
