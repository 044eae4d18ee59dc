"""cost of bringing the blocks a scan dirtied to the host, per scan.

Three ways over the bench workload (Merged, 640x480, 0.05 m):
  list+download : vbx_list_blocks + vbx_download_blocks (one copy per block into pageable memory)
  mirror        : vbx_mirror_updated into pageable memory (device gather, one copy, staging memcpy)
  mirror pinned : vbx_mirror_updated into a page-locked buffer (device gather, one copy)
Prints one JSON line.  Host wall clock around the calls (they are synchronous)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import voxblox_b200 as vb  # noqa: E402
from voxblox_b200 import scenes  # noqa: E402


def main():
    n_scans = int(os.environ.get("SCANS", "30"))
    scans = scenes.generate_parallel(scenes.c3_room_scan, range(n_scans))
    cfg = vb.TsdfIntegratorConfig(default_truncation_distance=0.2)
    out = {}
    for mode in ("list+download", "mirror", "mirror_pinned"):
        layer = vb.Layer(0.05, 16, engine_options=vb.EngineOptions(max_blocks=16384, max_points_per_scan=1 << 19,
                                                                  max_updates_per_pass=1 << 24))
        integ = vb.TsdfIntegratorFactory.create("merged", cfg, layer)
        pinned = layer.hostBuffer((512, 4096), vb.TSDF_DTYPE) if mode == "mirror_pinned" else None
        t_sum, b_sum, blocks = 0.0, 0, 0
        for i, s in enumerate(scans):
            integ.integratePointCloud((s[2], s[3]), s[0], s[1])
            t0 = time.perf_counter()
            if mode == "list+download":
                idx = layer.getAllUpdatedBlocks(1)
                vox, _ = layer.getBlocks(idx)
                layer.clearUpdated(1)  # Update::kMesh
            else:
                idx, vox, _ = layer.mirrorUpdated(2, 2, voxels_out=pinned)
            dt = time.perf_counter() - t0
            if i >= 3:
                t_sum += dt
                b_sum += vox.nbytes
                blocks += idx.shape[0]
        n = n_scans - 3
        out[mode] = {"ms_per_scan": 1e3 * t_sum / n, "blocks_per_scan": blocks / n, "MB_per_scan": b_sum / n / 1e6,
                     "GB_per_s": b_sum / t_sum / 1e9}
    print(json.dumps({"workload": "merged 640x480 0.05 m, dirty (kMesh) blocks mirrored after every scan", **out}))


if __name__ == "__main__":
    main()
