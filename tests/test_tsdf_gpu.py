"""-m gpu parity: the CUDA TSDF path through the C-ABI against the CPU oracle."""
import numpy as np
import pytest

import voxblox_b200 as vb
from oracle import pyoracle as po
from tests.parity import compare_tsdf
from voxblox_b200 import scenes

pytestmark = pytest.mark.gpu

REL_TOL = 1e-4  # north_star: TSDF distance / weight within 1e-4 relative


def _run(kind, scans, voxel_size, trunc, order, **cfg_kw):
    cfg = vb.TsdfIntegratorConfig(default_truncation_distance=trunc, integrator_threads=1, **cfg_kw)
    layer = vb.Layer(voxel_size, 16)
    integ = vb.TsdfIntegratorFactory.create(kind, cfg, layer)
    ocfg = po.TsdfConfig(default_truncation_distance=trunc, integrator_threads=1, **cfg_kw)
    omap = po.OracleMap(po.OracleLib("port"), ocfg, voxel_size, 16)
    for s in scans:
        integ.integratePointCloud((s[2], s[3]), s[0], s[1])
        omap.integrate(kind, s, order=order)
        gc, oc = integ.counters(), omap.counters()
        for k in ("rays", "clear_rays", "updates", "voxels_touched", "blocks_touched", "blocks_allocated"):
            assert gc[k] == oc[k], (k, gc, oc)
    return compare_tsdf(layer, omap)


def _assert_parity(rep):
    assert rep["blocks_equal"], rep
    assert rep["observed_equal"], rep
    assert rep["max_rel_err"] <= REL_TOL, rep
    assert rep["color_mismatch"] == 0, rep
    assert rep["updated_equal"], rep


def test_c1_simple_planar_wall():
    rep = _run(1, [scenes.c1_planar_wall()], 0.2, 0.8, po.ORDER_REFERENCE)
    print(rep)
    _assert_parity(rep)


def test_c1_merged_planar_wall():
    rep = _run(2, [scenes.c1_planar_wall()], 0.2, 0.8, po.ORDER_CANONICAL)
    print(rep)
    _assert_parity(rep)


@pytest.mark.parametrize("kind,order", [(1, po.ORDER_REFERENCE), (2, po.ORDER_CANONICAL)])
def test_room_sequence_small(kind, order):
    scans = scenes.c3_room_sequence(n_scans=4, width=160, height=120)
    rep = _run(kind, scans, 0.1, 0.4, order)
    print(rep)
    _assert_parity(rep)
