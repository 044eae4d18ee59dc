"""C4: MergedTsdfIntegrator + EsdfIntegrator::updateFromTsdfLayer(true) after every scan, 640x480,
0.05 m voxels; device time per ESDF update next to the reference's on the host."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import voxblox_b200 as vb
from oracle import pyoracle as po
from voxblox_b200 import scenes
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
scans = scenes.generate_parallel(scenes.c3_room_scan, range(N))
cfg = vb.TsdfIntegratorConfig(default_truncation_distance=0.2)
ekw = dict(max_distance_m=2.0, default_distance_m=2.0, min_distance_m=0.1, min_diff_m=0.001)  # ROS defaults (ros_params.h:128-140)
tsdf = vb.Layer(0.05, 16, engine_options=vb.EngineOptions(max_blocks=16384, max_points_per_scan=1 << 19, max_updates_per_pass=1 << 24))
integ = vb.TsdfIntegratorFactory.create("merged", cfg, tsdf)
esdf = vb.Layer(0.05, 16, voxel_type="esdf")
e = vb.EsdfIntegrator(vb.EsdfIntegratorConfig(**ekw), tsdf, esdf)
tsdf.setStageProfiling(True)
t_ms, e_ms = [], []
for s in scans:
    integ.integratePointCloud((s[2], s[3]), s[0], s[1]); t_ms.append(integ.lastDeviceMs())
    e.updateFromTsdfLayer(True); e_ms.append(e.lastDeviceMs())
    cnt = e.counters()
print("gpu tsdf ms/scan", np.round(np.mean(t_ms[3:]), 3), "esdf ms/update", np.round(np.mean(e_ms[3:]), 3), "last counters", cnt)
print({k: (round(v[0] / max(1, v[1]), 4), v[1]) for k, v in tsdf.stageMs().items() if v[1]})
# reference on the host
om = po.OracleMap(po.OracleLib("reference" if po.available("reference") else "port"), po.TsdfConfig(default_truncation_distance=0.2, integrator_threads=1), 0.05, 16)
om.esdf_create(po.EsdfConfig(**ekw))
rt, re_ = [], []
for s in scans[:min(N, 10)]:
    om.integrate(2, s); rt.append(om.last_seconds() * 1e3)
    om.esdf_update(False, True); re_.append(om.last_seconds() * 1e3)
print("cpu tsdf ms/scan", np.round(np.mean(rt[3:]), 2), "esdf ms/update", np.round(np.mean(re_[3:]), 2))
