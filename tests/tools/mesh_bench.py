"""N3 measurement: MeshIntegrator::generateMesh on the device map vs the reference's CPU mesher.
  (a) the bench workload (640x480 room scans, 0.05 m): incremental mesh after every scan
      (only_mesh_updated_blocks, clear_updated_flag), device time / wall time incl. the download
  (b) one C5 LiDAR map (~730+ blocks at 0.05 m): all blocks in one call
Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import voxblox_b200 as vb
from voxblox_b200 import scenes
from oracle import pyoracle as po

N = int(sys.argv[1]) if len(sys.argv) > 1 else 12
which = "reference" if po.available("reference") else "port"
out = {"cpu_kind": which}

scans = scenes.generate_parallel(scenes.c3_room_scan, range(N))
cfg = vb.TsdfIntegratorConfig(default_truncation_distance=0.2)
layer = vb.Layer(0.05, 16)
integ = vb.TsdfIntegratorFactory.create("merged", cfg, layer)
mesh_layer = vb.MeshLayer(layer.block_size())
mesher = vb.MeshIntegrator(vb.MeshIntegratorConfig(), layer, mesh_layer)
om = po.OracleMap(po.OracleLib(which), po.TsdfConfig(default_truncation_distance=0.2, integrator_threads=1), 0.05, 16)
dev, wall, cpu, nblk, nvert = [], [], [], [], []
for s in scans:
    integ.integratePointCloud((s[2], s[3]), s[0], s[1])
    om.integrate(2, s)
    t0 = time.perf_counter()
    mesher.generateMesh(True, True)
    wall.append((time.perf_counter() - t0) * 1e3)
    dev.append(mesher.lastDeviceMs())
    nblk.append(mesher.last_blocks)
    nvert.append(mesher.last_vertices)
    om.mesh_generate(True, 1e-4, True, True)
    cpu.append(om.last_seconds() * 1e3)
k = slice(2, None)
out["incremental"] = {"scans": N, "blocks_per_call": float(np.mean(nblk[k])), "vertices_per_call": float(np.mean(nvert[k])),
                      "gpu_device_ms": float(np.mean(dev[k])), "gpu_wall_ms_incl_download": float(np.mean(wall[k])),
                      "cpu_ms_1_thread": float(np.mean(cpu[k]))}

# (b) a larger map, meshed in one call
opts = vb.EngineOptions(max_blocks=8192, max_points_per_scan=1 << 19)
kw = dict(default_truncation_distance=0.2, max_ray_length_m=10.0, use_const_weight=1)
layer2 = vb.Layer(0.05, 16, engine_options=opts)
integ2 = vb.TsdfIntegratorFactory.create("merged", vb.TsdfIntegratorConfig(**kw), layer2)
om2 = po.OracleMap(po.OracleLib(which), po.TsdfConfig(integrator_threads=1, **kw), 0.05, 16)
for i in range(3):
    s = scenes.c5_lidar_scan(i)
    integ2.integratePointCloud((s[2], s[3]), s[0], s[1])
    om2.integrate(2, s)
mesh2 = vb.MeshLayer(layer2.block_size())
mesher2 = vb.MeshIntegrator(vb.MeshIntegratorConfig(), layer2, mesh2)
d2, w2 = [], []
for _ in range(4):
    t0 = time.perf_counter()
    mesher2.generateMesh(False, False)
    w2.append((time.perf_counter() - t0) * 1e3)
    d2.append(mesher2.lastDeviceMs())
om2.mesh_generate(True, 1e-4, False, False)
nb, nv = mesher2.last_blocks, mesher2.last_vertices
alg = nb * 4096 * 12 + nv * 28
out["full_map"] = {"blocks": nb, "vertices": nv, "gpu_device_ms": float(np.min(d2)), "gpu_wall_ms_incl_download": float(np.min(w2)),
                   "cpu_ms_1_thread": om2.last_seconds() * 1e3, "alg_bytes": alg,
                   "achieved_gbs": alg / (np.min(d2) * 1e-3) / 1e9}
print(json.dumps(out))
