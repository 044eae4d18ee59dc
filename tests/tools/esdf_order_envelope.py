"""Reference-vs-reference: how much does the reference's own ESDF depend on the ORDER in which
updateFromTsdfBlocks walks the updated blocks (in the reference: the TSDF layer's hash-map iteration order)?"""
import sys, json
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle import pyoracle as po
from voxblox_b200 import scenes

def stats(a, b, voxel, md):
    ia, ib = a.block_indices(1), b.block_indices(1)
    assert ia.shape == ib.shape and (ia == ib).all()
    va = np.stack([a.block(i, 1)[0] for i in ia]); vb = np.stack([b.block(i, 1)[0] for i in ib])
    obs = va["observed"] != 0
    assert ((vb["observed"] != 0) == obs).all()
    da, db = va["distance"][obs].astype(np.float64), vb["distance"][obs].astype(np.float64)
    err = np.abs(da - db); rel = err / np.maximum(np.abs(da), 1e-3 * voxel)
    return {"observed": int(obs.sum()), "bit_exact": float((da == db).mean()), "within_1e-4_rel": float((rel <= 1e-4).mean()),
            "within_2_min_diff": float((err <= 2 * md + 1e-7).mean()) if md > 0 else None,
            "within_one_voxel": float((err <= voxel * 1.0001).mean()), "max_abs_err_m": float(err.max()),
            "rmse_m": float(np.sqrt((err ** 2).mean()))}

def run(scene, voxel, trunc, scans, md, mq, order):
    lib = po.OracleLib("reference")
    maps = []
    for which in ("hash", order):
        m = po.OracleMap(lib, po.TsdfConfig(default_truncation_distance=trunc, integrator_threads=1), voxel, 16)
        m.esdf_create(po.EsdfConfig(max_distance_m=2.0, default_distance_m=2.0, min_distance_m=trunc / 2, min_diff_m=md, multi_queue=mq))
        maps.append(m)
    a, b = maps
    rng = np.random.default_rng(1)
    for s in scans:
        a.integrate(2, s); b.integrate(2, s)
        a.esdf_update(batch=False, clear_updated_flag=True)
        idx = b.block_indices(0)
        upd = [i for i in idx if b.block(i, 0)[1] & 4]
        upd = np.array(upd, dtype=np.int32)
        if order == "reversed_sorted": upd = upd[::-1].copy()
        elif order == "random": upd = upd[rng.permutation(len(upd))].copy()
        b.esdf_update_blocks(upd, incremental=True)
        # clear the kEsdf bit like updateFromTsdfLayer(true): the harness has no direct call; emulate with an update over no blocks
        b.esdf_update(batch=False, clear_updated_flag=True)   # (all listed blocks were just processed: this re-walks them but changes nothing new?) 
    return stats(a, b, voxel, md)

if __name__ == "__main__":
    small = scenes.c3_room_sequence(n_scans=4, width=160, height=120)
    for md, mq, name in ((0.0, 1, "min_diff 0 multi_queue"), (1e-3, 0, "ROS defaults")):
        for order in ("sorted", "reversed_sorted", "random"):
            print("room_small", name, order, json.dumps(run("small", 0.1, 0.4, small, md, mq, order)), flush=True)
