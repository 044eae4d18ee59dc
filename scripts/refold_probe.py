"""How often does the bundle fold fall back to IEEE division, and how big are those bundles?"""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import voxblox_b200 as vb
from voxblox_b200 import scenes

scans = scenes.generate_parallel(scenes.c3_room_scan, range(16))
layer = vb.Layer(0.05, 16, engine_options=vb.EngineOptions(max_blocks=16384, max_points_per_scan=1 << 19, max_updates_per_pass=1 << 24))
integ = vb.TsdfIntegratorFactory.create("merged", vb.TsdfIntegratorConfig(default_truncation_distance=0.2), layer)
layer.setStageProfiling(True)
rows = []
for s in scans:
    integ.integratePointCloud((s[2], s[3]), s[0], s[1])
    c = integ.counters()
    rows.append((c["rays"] + c["clear_rays"], c["refolded_bundles"], c["refolded_points"]))
print(json.dumps({"bundles_refolded_points": rows, "stage_ms": {k: v[0] / max(v[1], 1) for k, v in layer.stageMs().items() if v[1]}}))
