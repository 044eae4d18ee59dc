"""Where does a pipelined scan spend its time?  Timestamps of the hand-off events of the last ten scans
(VBX_ASYNC_TIMELINE, vbx_debug_async_timeline): front half / ray walk / record sort / apply per scan and
the pace of each stage from scan to scan."""
import ctypes as C, json, os, sys
os.environ["VBX_ASYNC_TIMELINE"] = "1"
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import voxblox_b200 as vb
from voxblox_b200 import scenes

n = 65
scans = scenes.generate_parallel(scenes.c3_room_scan, range(n))
dev = torch.device("cuda", 0)
d_xyz = [torch.from_numpy(s[0]).to(dev) for s in scans]
d_rgba = [torch.from_numpy(s[1]).to(dev) for s in scans]
layer = vb.Layer(0.05, 16, engine_options=vb.EngineOptions(max_blocks=16384, max_points_per_scan=1 << 19, max_updates_per_pass=1 << 24))
integ = vb.TsdfIntegratorFactory.create("merged", vb.TsdfIntegratorConfig(default_truncation_distance=0.2), layer)
host = "--host" in sys.argv   # clouds in page-locked host memory: the H2D copies are part of the pipeline
if host:
    h_xyz = [torch.from_numpy(s[0]).pin_memory() for s in scans]
    h_rgba = [torch.from_numpy(s[1]).pin_memory() for s in scans]
for i in range(n):
    if host:
        integ.integratePointCloudAsync((scans[i][2], scans[i][3]), h_xyz[i].numpy(), h_rgba[i].numpy())
    else:
        integ.integratePointCloudAsync((scans[i][2], scans[i][3]), d_xyz[i].data_ptr(), d_rgba[i].data_ptr(), int(scans[i][0].shape[0]))
layer.sync()
ctx = layer._ctx
ctx.lib.vbx_debug_async_timeline.restype = C.c_int
ctx.lib.vbx_debug_async_timeline.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
seq = np.zeros(10, np.uint64); ms = np.zeros((10, 5), np.float32)
ctx.check(ctx.lib.vbx_debug_async_timeline(ctx.handle, seq.ctypes.data, ms.ctypes.data, 10), "timeline")
o = np.argsort(seq)
seq, ms = seq[o], ms[o]
t0 = ms[0, 0]
names = ["front_start", "front_done", "walked", "sorted", "applied"]
rows = []
for k in range(10):
    r = ms[k] - t0
    rows.append({"seq": int(seq[k]), **{nm: round(float(v), 4) for nm, v in zip(names, r)},
                 "front_ms": round(float(r[1] - r[0]), 4), "wait+walk_ms": round(float(r[2] - r[1]), 4),
                 "sort_ms": round(float(r[3] - r[2]), 4), "apply_ms": round(float(r[4] - r[3]), 4)})
pace = {nm: round(float(np.diff(ms[:, j]).mean()), 4) for j, nm in enumerate(names)}
print(json.dumps({"clouds": "page-locked host" if host else "device", "scans": rows, "pace_ms_per_scan": pace}, indent=1))
