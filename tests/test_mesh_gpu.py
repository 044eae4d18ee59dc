"""-m gpu parity: the device mesher (vbx_mesh.cu) through the C-ABI against the CPU oracle's
MeshIntegrator restatement -- same vertices, normals and colours in the same order."""
import numpy as np
import pytest

import voxblox_b200 as vb
from oracle import pyoracle as po
from tests.parity import compare_tsdf
from voxblox_b200 import scenes

pytestmark = pytest.mark.gpu


def _compare(mesh_layer, omap):
    gi, oi = mesh_layer.getAllAllocatedMeshes(), omap.mesh_block_indices()
    assert gi.shape == oi.shape and (gi == oi).all(), (gi.shape, oi.shape)
    rep = {"blocks": int(len(oi)), "vertices": 0, "vertex_mismatch": 0, "normal_mismatch": 0, "color_mismatch": 0,
           "count_mismatch": 0, "max_vertex_err": 0.0}
    for i in oi:
        v, n, c, upd = omap.mesh_block(i)
        m = mesh_layer.getMeshPtrByIndex(i)
        assert m.updated == upd
        rep["vertices"] += len(v)
        if m.vertices.shape != v.shape:
            rep["count_mismatch"] += 1
            continue
        assert (m.indices == np.arange(len(v))).all()
        rep["vertex_mismatch"] += int((m.vertices.view(np.uint32) != v.view(np.uint32)).any(axis=-1).sum())
        rep["normal_mismatch"] += int((m.normals.view(np.uint32) != n.view(np.uint32)).any(axis=-1).sum())
        if len(v):
            rep["max_vertex_err"] = max(rep["max_vertex_err"], float(np.abs(m.vertices - v).max()))
        if c is None:
            assert m.colors.shape[0] == 0
        else:
            rep["color_mismatch"] += int((m.colors != c).any(axis=-1).sum())
    return rep


def _setup(voxel_size, trunc, vps=16):
    cfg = vb.TsdfIntegratorConfig(default_truncation_distance=trunc, integrator_threads=1)
    layer = vb.Layer(voxel_size, vps)
    integ = vb.TsdfIntegratorFactory.create("merged", cfg, layer)
    omap = po.OracleMap(po.OracleLib("port"), po.TsdfConfig(default_truncation_distance=trunc, integrator_threads=1),
                        voxel_size, vps)
    return layer, integ, omap


@pytest.mark.parametrize("use_color", [True, False])
def test_incremental_mesh_matches_oracle(use_color):
    """generateMesh(only_mesh_updated_blocks=true, clear_updated_flag=true) after every scan, as
    TsdfServer::updateMesh does (voxblox_ros/src/tsdf_server.cc:494-501)."""
    scans = scenes.c3_room_sequence(n_scans=4, width=160, height=120)
    layer, integ, omap = _setup(0.1, 0.4)
    mesh_layer = vb.MeshLayer(layer.block_size())
    mesher = vb.MeshIntegrator(vb.MeshIntegratorConfig(use_color=use_color), layer, mesh_layer)
    for s in scans:
        integ.integratePointCloud((s[2], s[3]), s[0], s[1])
        omap.integrate(2, s, order=po.ORDER_REFERENCE)
        mesher.generateMesh(True, True)
        omap.mesh_generate(use_color, 1e-4, True, True)
        assert len(layer.getAllUpdatedBlocks(1)) == 0   # kMesh bits cleared
    assert compare_tsdf(layer, omap)["max_rel_err"] == 0.0               # (incl. the updated bits)
    rep = _compare(mesh_layer, omap)
    print(rep, mesher.lastDeviceMs())
    assert rep["vertices"] > 3000
    assert rep["count_mismatch"] == 0 and rep["vertex_mismatch"] == 0 and rep["normal_mismatch"] == 0, rep
    assert rep["color_mismatch"] == 0, rep


def test_full_mesh_small_blocks_and_second_call_is_empty():
    """All blocks at once (only_mesh_updated_blocks=false, the const-layer constructor's use), at
    voxels_per_side 8; an incremental call right after a clearing one finds nothing to do."""
    scans = scenes.c3_room_sequence(n_scans=2, width=160, height=120)
    layer, integ, omap = _setup(0.1, 0.4, vps=8)
    mesh_layer = vb.MeshLayer(layer.block_size())
    mesher = vb.MeshIntegrator(vb.MeshIntegratorConfig(min_weight=0.5), layer, mesh_layer)
    for s in scans:
        integ.integratePointCloud((s[2], s[3]), s[0], s[1])
        omap.integrate(2, s, order=po.ORDER_REFERENCE)
    mesher.generateMesh(False, False)
    omap.mesh_generate(True, 0.5, False, False)
    rep = _compare(mesh_layer, omap)
    print(rep)
    assert rep["vertices"] > 500
    assert rep["count_mismatch"] == 0 and rep["vertex_mismatch"] == 0 and rep["normal_mismatch"] == 0, rep
    assert rep["color_mismatch"] == 0, rep
    assert len(layer.getAllUpdatedBlocks(1)) == len(layer.getAllAllocatedBlocks())  # not cleared
    mesher.generateMesh(True, True)
    assert mesher.last_blocks == len(layer.getAllAllocatedBlocks())
    mesher.generateMesh(True, True)
    assert mesher.last_blocks == 0 and mesher.last_vertices == 0


def test_mesh_of_empty_and_unobserved_maps():
    layer = vb.Layer(0.1, 16)
    integ = vb.TsdfIntegratorFactory.create("merged", vb.TsdfIntegratorConfig(default_truncation_distance=0.4), layer)
    mesh_layer = vb.MeshLayer(layer.block_size())
    mesher = vb.MeshIntegrator(vb.MeshIntegratorConfig(), layer, mesh_layer)
    mesher.generateMesh(False, True)
    assert mesh_layer.getNumberOfAllocatedMeshes() == 0
    # blocks uploaded with zero weight: meshes are allocated but stay empty (mesh_integrator.h:146-149)
    idx = np.array([[0, 0, 0], [1, 0, 0]], np.int32)
    layer.insertBlocks(idx, np.zeros((2, 4096), vb.TSDF_DTYPE))
    mesher.generateMesh(False, True)
    assert mesh_layer.getNumberOfAllocatedMeshes() == 2
    assert all(mesh_layer.getMeshPtrByIndex(i).size() == 0 for i in idx)
