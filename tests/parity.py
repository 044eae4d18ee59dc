"""Comparison helpers shared by the parity tests, smoke() and bench.py."""
from __future__ import annotations

from typing import Dict

import numpy as np


def rel_err(a: np.ndarray, b: np.ndarray, floor: float) -> np.ndarray:
    """|a - b| / max(|b|, floor): relative error with an absolute floor for values near 0."""
    return np.abs(a.astype(np.float64) - b.astype(np.float64)) / np.maximum(np.abs(b.astype(np.float64)), floor)


def compare_tsdf(layer, omap, dist_floor: float = None) -> Dict:
    """GPU Layer (voxblox_b200.Layer) against an oracle map (oracle.pyoracle.OracleMap).

    Block sets must be identical (allocation is exact); distances and weights are
    compared relatively (north_star: 1e-4), colours exactly."""
    gi = layer.getAllAllocatedBlocks()
    oi = omap.block_indices()
    rep = {"gpu_blocks": int(len(gi)), "oracle_blocks": int(len(oi)),
           "blocks_equal": gi.shape == oi.shape and bool((gi == oi).all())}
    if not rep["blocks_equal"]:
        return rep
    gv, gupd = layer.getBlocks(gi)
    ov = np.stack([omap.block(i)[0] for i in oi]) if len(oi) else gv
    oupd = np.array([omap.block(i)[1] for i in oi], dtype=np.uint8)
    floor = dist_floor if dist_floor is not None else 1e-3 * layer.voxel_size()
    de = rel_err(gv["distance"], ov["distance"], floor)
    we = rel_err(gv["weight"], ov["weight"], 1e-12)
    obs = ov["weight"] > 0
    rep.update({
        "voxels_observed": int(obs.sum()),
        "observed_equal": bool(((gv["weight"] > 0) == obs).all()),
        "max_rel_err_distance": float(de.max()) if de.size else 0.0,
        "max_rel_err_weight": float(we.max()) if we.size else 0.0,
        "n_dist_over_1e-4": int((de > 1e-4).sum()),
        "n_bit_exact": int(((gv["distance"] == ov["distance"]) & (gv["weight"] == ov["weight"])).sum()),
        "n_voxels": int(gv.size),
        "color_mismatch": int((gv["color"] != ov["color"]).any(axis=-1).sum()),
        "updated_equal": bool((gupd == oupd).all()),
    })
    rep["max_rel_err"] = max(rep["max_rel_err_distance"], rep["max_rel_err_weight"])
    return rep
