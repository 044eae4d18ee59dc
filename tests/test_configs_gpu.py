"""-m gpu parity at BASELINE.json's configurations C2-C5 at FULL size: the
oracle finishes each of these scans in well under a second, so they are compared voxel for voxel.
(C1 is tests/test_tsdf_gpu.py::test_c1_*; the bench workload is the C3/C4 cloud stream.)"""
import numpy as np
import pytest

import voxblox_b200 as vb
from oracle import pyoracle as po
from tests.parity import compare_esdf, compare_tsdf
from voxblox_b200 import scenes

pytestmark = pytest.mark.gpu


def _pair(kind, voxel_size, opts=None, **kw):
    cfg = vb.TsdfIntegratorConfig(integrator_threads=1, **kw)
    layer = vb.Layer(voxel_size, 16, engine_options=opts)
    integ = vb.TsdfIntegratorFactory.create(kind, cfg, layer)
    omap = po.OracleMap(po.OracleLib("port"), po.TsdfConfig(integrator_threads=1, **kw), voxel_size, 16)
    return layer, integ, omap


def _exact(rep):
    assert rep["blocks_equal"] and rep["observed_equal"] and rep["updated_equal"], rep
    assert rep["max_rel_err"] <= 1e-4 and rep["color_mismatch"] == 0, rep   # north_star tolerance
    assert rep["n_bit_exact"] == rep["n_voxels"], rep                        # and in fact bit for bit


@pytest.mark.parametrize("trunc", [0.4, 4.0])
def test_c2_merged_sphere_room_full_size(trunc):
    """C2: MergedTsdfIntegrator, 640x480 inside a 3 m sphere, 0.10 m voxels; truncation 4 voxels (the
    reference's convention) and the literal "4 m truncation" of BASELINE.json as the stress variant."""
    layer, integ, omap = _pair("merged", 0.1, default_truncation_distance=trunc)
    for i in (0, 1, 2):
        s = scenes.c2_sphere_scan(i)
        assert s[0].shape[0] == 640 * 480
        integ.integratePointCloud((s[2], s[3]), s[0], s[1])
        omap.integrate(2, s, order=po.ORDER_REFERENCE)
        gc, oc = integ.counters(), omap.counters()
        for k in ("rays", "clear_rays", "updates", "voxels_touched", "blocks_touched", "blocks_allocated"):
            assert gc[k] == oc[k], (k, gc, oc)
    rep = compare_tsdf(layer, omap)
    print(trunc, rep, gc)
    _exact(rep)


def test_c3_fast_full_size_statistics():
    """C3: FastTsdfIntegrator on 640x480 room scans at 0.05 m.  The Fast integrator is approximate by
    design (lossy hash sets), so the comparison is the reference's own kind: block sets, coverage
    and error statistics (test_sdf_integrators.cc:155-178)."""
    layer, integ, omap = _pair("fast", 0.05, default_truncation_distance=0.2)
    for i in range(3):
        s = scenes.c3_room_scan(i)
        integ.integratePointCloud((s[2], s[3]), s[0], s[1])
        omap.integrate(3, s)
    gi, oi = layer.getAllAllocatedBlocks(), omap.block_indices()
    gs, os_ = {tuple(b) for b in gi.tolist()}, {tuple(b) for b in oi.tolist()}
    assert len(gs & os_) >= 0.9 * len(os_), (len(gs), len(os_))
    common = np.array(sorted(gs & os_), np.int32)
    gv, _ = layer.getBlocks(common)
    ov = np.stack([omap.block(i)[0] for i in common])
    both = (gv["weight"] > 0) & (ov["weight"] > 0)
    cover = both.sum() / max(1, (ov["weight"] > 0).sum())
    rmse = float(np.sqrt(np.mean((gv["distance"][both] - ov["distance"][both]) ** 2)))
    print("fast: blocks", len(gs), len(os_), "coverage", cover, "rmse", rmse)
    assert cover > 0.9 and rmse < 0.05  # within one voxel, as in tests/test_tsdf_gpu.py


def test_c4_merged_plus_esdf_every_scan_full_size():
    """C4: Merged + updateFromTsdfLayer(true) after every scan, 640x480 at 0.05 m, min_diff_m = 0."""
    ekw = dict(max_distance_m=2.0, default_distance_m=2.0, min_distance_m=0.1, min_diff_m=0.0)
    layer, integ, omap = _pair("merged", 0.05, default_truncation_distance=0.2)
    esdf = vb.Layer(0.05, 16, voxel_type="esdf")
    eint = vb.EsdfIntegrator(vb.EsdfIntegratorConfig(**ekw), layer, esdf)
    omap.esdf_create(po.EsdfConfig(**ekw))
    for i in range(2):
        s = scenes.c3_room_scan(i)
        integ.integratePointCloud((s[2], s[3]), s[0], s[1])
        omap.integrate(2, s, order=po.ORDER_REFERENCE)
        eint.updateFromTsdfLayer(True)
        omap.esdf_update(batch=False, clear_updated_flag=True)
    _exact(compare_tsdf(layer, omap))
    rep = compare_esdf(esdf, omap, 2.0)
    print(rep, eint.counters())
    assert rep["blocks_equal"] and rep["observed_equal"] and rep["fixed_equal"], rep
    assert rep["rmse"] < 0.05 and rep["n_over_1e-4"] < 0.08 * rep["voxels_observed"], rep


def test_c5_merged_lidar_full_size():
    """C5: MergedTsdfIntegrator, 2048x128 spinning LiDAR in a 9x9x4 m hall, 0.05 m voxels, constant
    weights, max_ray_length 10 m: ~13 M voxel updates on ~1.7 M voxels in ~730 blocks per scan."""
    opts = vb.EngineOptions(max_blocks=8192, max_points_per_scan=1 << 19)
    layer, integ, omap = _pair("merged", 0.05, opts, default_truncation_distance=0.2, max_ray_length_m=10.0,
                               use_const_weight=1)
    for i in range(2):
        s = scenes.c5_lidar_scan(i)
        assert s[0].shape[0] == 2048 * 128
        integ.integratePointCloud((s[2], s[3]), s[0], s[1])
        omap.integrate(2, s, order=po.ORDER_REFERENCE)
        gc, oc = integ.counters(), omap.counters()
        for k in ("rays", "clear_rays", "updates", "voxels_touched", "blocks_touched", "blocks_allocated"):
            assert gc[k] == oc[k], (k, gc, oc)
    rep = compare_tsdf(layer, omap)
    print(rep, gc, integ.lastDeviceMs())
    assert gc["updates"] > 10_000_000
    _exact(rep)
