"""N4: ICP::runICP of a 640x480 cloud against a map of earlier scans (0.05 m voxels): wall / device time
per call on the GPU (host cloud, and cloud already on the device) for num_threads = 1, 8, 32, next to the
reference's own ICP on the host (one thread: its deterministic setting)."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import voxblox_b200 as vb
from oracle import pyoracle as po
from voxblox_b200 import scenes
N = 8
scans = scenes.generate_parallel(scenes.c3_room_scan, range(N + 4))
cfg = vb.TsdfIntegratorConfig(default_truncation_distance=0.2)
tsdf = vb.Layer(0.05, 16)
integ = vb.TsdfIntegratorFactory.create("merged", cfg, tsdf)
which = "reference" if po.available("reference") else "port"
om = po.OracleMap(po.OracleLib(which), po.TsdfConfig(default_truncation_distance=0.2, integrator_threads=1), 0.05, 16)
for s in scans[:4]:
    integ.integratePointCloud((s[2], s[3]), s[0], s[1])
    om.integrate(2, s)
out = {"cloud": "640x480 room scan, ~261 k points; 6.5 k mini batches of 20", "cpu_kind": which}
for T in (1, 8, 32):
    icp = vb.ICP(vb.ICPConfig(num_threads=T))
    wall, wall_dev, upd = [], [], []
    for k, s in enumerate(scans[4:]):
        t0 = s[3] + np.float32([0.03, -0.02, 0.01])
        d = torch.from_numpy(np.ascontiguousarray(s[0])).cuda()
        torch.cuda.synchronize()
        a = time.perf_counter(); n, T1 = icp.runICP(tsdf, s[0], (s[2], t0), seed=k); b = time.perf_counter()
        n2, T2 = icp.runICPDevice(tsdf, d.data_ptr(), int(s[0].shape[0]), (s[2], t0), seed=k); c = time.perf_counter()
        wall.append((b - a) * 1e3); wall_dev.append((c - b) * 1e3); upd.append(n)
    out[f"gpu_threads_{T}"] = {"wall_ms_host_cloud": round(float(np.mean(wall[2:])), 3),
                               "wall_ms_device_cloud": round(float(np.mean(wall_dev[2:])), 3), "updates": int(np.mean(upd))}
cpu = []
for k, s in enumerate(scans[4:8]):
    t0 = s[3] + np.float32([0.03, -0.02, 0.01])
    om.icp(po.IcpConfig(num_threads=1), s[0], s[2], t0, k); cpu.append(om.last_seconds() * 1e3)
out["cpu_threads_1_ms"] = round(float(np.mean(cpu[1:])), 2)
# host shuffle alone (std::shuffle of the index sequence, part of every call on both sides)
perm = np.zeros(scans[4][0].shape[0], np.uint32)
a = time.perf_counter(); po.OracleLib("port").lib.vbo_icp_shuffle(perm.size, 1, perm.ctypes.data); out["host_shuffle_ms"] = round((time.perf_counter() - a) * 1e3, 3)
print(json.dumps(out))
