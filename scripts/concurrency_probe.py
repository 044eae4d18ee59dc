"""How much of the GPU does one pipelined scan stream use?  M independent maps (contexts) on one
GPU, scans submitted round-robin through vbx_tsdf_integrate_async; aggregate points/s for M = 1..4.
If the aggregate grows with M, a deeper pipeline inside ONE context has headroom."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import voxblox_b200 as vb  # noqa: E402
from voxblox_b200 import scenes  # noqa: E402


def main():
    n_scans = int(os.environ.get("SCANS", "45"))
    warm = 5
    scans = scenes.generate_parallel(scenes.c3_room_scan, range(n_scans))
    dev = torch.device("cuda", 0)
    d_xyz = [torch.from_numpy(s[0]).to(dev) for s in scans]
    d_rgba = [torch.from_numpy(s[1]).to(dev) for s in scans]
    npts = [int(s[0].shape[0]) for s in scans]
    cfg = vb.TsdfIntegratorConfig(default_truncation_distance=0.2)
    out = {}
    for m in (1, 2, 3, 4):
        layers, integs = [], []
        for _ in range(m):
            layer = vb.Layer(0.05, 16, engine_options=vb.EngineOptions(max_blocks=16384, max_points_per_scan=1 << 19,
                                                                      max_updates_per_pass=1 << 24))
            layers.append(layer)
            integs.append(vb.TsdfIntegratorFactory.create("merged", cfg, layer))
        for i in range(warm):
            for g in integs:
                g.integratePointCloudAsync((scans[i][2], scans[i][3]), d_xyz[i].data_ptr(), d_rgba[i].data_ptr(), npts[i])
        for la in layers:
            la.sync()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(warm, n_scans):
            for g in integs:
                g.integratePointCloudAsync((scans[i][2], scans[i][3]), d_xyz[i].data_ptr(), d_rgba[i].data_ptr(), npts[i])
        for la in layers:
            la.sync()
        dt = time.perf_counter() - t0
        pts = m * sum(npts[warm:])
        out[str(m)] = {"points_per_s": pts / dt, "ms_per_scan_per_map": 1e3 * dt / (n_scans - warm)}
        del layers, integs
    print(json.dumps({"probe": "M independent pipelined maps on one GPU", **out}))


if __name__ == "__main__":
    main()
