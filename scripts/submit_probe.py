"""Is the pipelined path host-bound?  Time spent inside the submission calls vs. time until the
last scan is in the map."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import voxblox_b200 as vb
from voxblox_b200 import scenes

n = 205
scans = scenes.generate_parallel(scenes.c3_room_scan, range(n))
dev = torch.device("cuda", 0)
d_xyz = [torch.from_numpy(s[0]).to(dev) for s in scans]
d_rgba = [torch.from_numpy(s[1]).to(dev) for s in scans]
npts = [int(s[0].shape[0]) for s in scans]
layer = vb.Layer(0.05, 16, engine_options=vb.EngineOptions(max_blocks=16384, max_points_per_scan=1 << 19, max_updates_per_pass=1 << 24))
integ = vb.TsdfIntegratorFactory.create("merged", vb.TsdfIntegratorConfig(default_truncation_distance=0.2), layer)
for i in range(5):
    integ.integratePointCloudAsync((scans[i][2], scans[i][3]), d_xyz[i].data_ptr(), d_rgba[i].data_ptr(), npts[i])
layer.sync()
t0 = time.perf_counter()
per = []
for i in range(5, n):
    a = time.perf_counter()
    integ.integratePointCloudAsync((scans[i][2], scans[i][3]), d_xyz[i].data_ptr(), d_rgba[i].data_ptr(), npts[i])
    per.append(time.perf_counter() - a)
t1 = time.perf_counter()
layer.sync()
t2 = time.perf_counter()
per = np.array(per) * 1e3
cn = integ.counters()
print(json.dumps({"scans": n - 5, "host_wait_for_free_set_ms_per_scan": cn["async_wait_ns_total"] / 1e6 / n,
                  "host_in_submission_calls_ms_per_scan_incl_wait": cn["async_submit_ns_total"] / 1e6 / n, "submit_ms_per_scan_mean": float(per.mean()), "submit_ms_per_scan_median": float(np.median(per)),
                  "submit_ms_first10": [round(float(v), 3) for v in per[:10]],
                  "loop_ms_per_scan": (t1 - t0) * 1e3 / (n - 5), "total_ms_per_scan": (t2 - t0) * 1e3 / (n - 5),
                  "drain_ms": (t2 - t1) * 1e3}))
