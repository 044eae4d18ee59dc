import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import voxblox_b200 as vb
from oracle import pyoracle as po
from tests.parity import compare_esdf
from tests.test_esdf_gpu import _setup, EKW
from voxblox_b200 import scenes
scans = scenes.c3_room_sequence(n_scans=4, width=160, height=120)
tsdf, integ, esdf, eint, omap = _setup(0.1, 0.4, EKW)
for s in scans:
    integ.integratePointCloud((s[2], s[3]), s[0], s[1]); omap.integrate(2, s, order=po.ORDER_REFERENCE)
eint.updateFromTsdfLayerBatch(); omap.esdf_update(batch=True)
rep = compare_esdf(esdf, omap, 4.0)
for k, v in rep.items(): print(k, v)
print(eint.counters())
gi = esdf.getAllAllocatedBlocks(); gv, _ = esdf.getBlocks(gi)
ov = np.stack([omap.block(i, 1)[0] for i in gi]); tv = np.stack([omap.block(i, 0)[0] for i in gi])
obs = ov["observed"] != 0
bad = obs & (np.abs(gv["distance"] - ov["distance"]) > 1e-4 * np.maximum(np.abs(ov["distance"]), 1e-4))
print("bad", bad.sum())
idx = np.argwhere(bad)[:25]
for b, l in idx:
    print(gi[b], l, "gpu", gv["distance"][b, l], gv["parent"][b, l], "oracle", ov["distance"][b, l], ov["parent"][b, l], "fixed", ov["fixed"][b, l], "tsdf", tv["distance"][b, l], tv["weight"][b, l])
# distribution of oracle values among bad
print("oracle sign of bad:", np.sign(ov["distance"][bad]).tolist()[:40])
