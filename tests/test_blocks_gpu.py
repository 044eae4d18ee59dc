"""-m gpu: host-driven block management (upload / remove / clear) keeps the device map consistent."""
import numpy as np
import pytest

import voxblox_b200 as vb
from voxblox_b200 import scenes

pytestmark = pytest.mark.gpu


def _make():
    cfg = vb.TsdfIntegratorConfig(default_truncation_distance=0.4)
    layer = vb.Layer(0.1, 16, engine_options=vb.EngineOptions(max_blocks=2048, max_updates_per_pass=1 << 22))
    return layer, vb.TsdfIntegratorFactory.create("merged", cfg, layer)


def _snapshot(layer):
    idx = layer.getAllAllocatedBlocks()
    vox, upd = layer.getBlocks(idx)
    return idx, vox, upd


def test_upload_continue_equals_uninterrupted():
    scans = scenes.c3_room_sequence(n_scans=4, width=128, height=96)
    a, ia = _make()
    for s in scans[:2]:
        ia.integratePointCloud((s[2], s[3]), s[0], s[1])
    idx, vox, upd = _snapshot(a)
    b, ib = _make()
    b.insertBlocks(idx, vox, upd)  # e.g. a map loaded from disk on the host
    i2, v2, u2 = _snapshot(b)
    assert (i2 == idx).all() and v2.tobytes() == vox.tobytes() and (u2 == upd).all()
    for s in scans[2:]:
        ia.integratePointCloud((s[2], s[3]), s[0], s[1])
        ib.integratePointCloud((s[2], s[3]), s[0], s[1])
    ja, va, ua = _snapshot(a)
    jb, vb_, ub = _snapshot(b)
    assert (ja == jb).all() and va.tobytes() == vb_.tobytes() and (ua == ub).all()


def test_remove_blocks_and_clear():
    scans = scenes.c3_room_sequence(n_scans=3, width=128, height=96)
    a, ia = _make()
    for s in scans[:2]:
        ia.integratePointCloud((s[2], s[3]), s[0], s[1])
    idx, vox, upd = _snapshot(a)
    kill = idx[::3]
    a.removeBlocks(kill)                        # Layer::removeBlock, core/layer.h:163
    a.removeBlocks(np.array([[900, 900, 900]]))  # erasing a missing block is a no-op (unordered_map::erase)
    keep = np.array([i for i in range(len(idx)) if i % 3 != 0])
    i2, v2, u2 = _snapshot(a)
    assert (i2 == idx[keep]).all()
    assert v2.tobytes() == vox[keep].tobytes() and (u2 == upd[keep]).all()
    # integration keeps working: removed blocks come back from the empty state
    ia.integratePointCloud((scans[2][2], scans[2][3]), scans[2][0], scans[2][1])
    i3, v3, _ = _snapshot(a)
    assert len(i3) >= len(i2)
    fresh, ifresh = _make()
    ifresh.integratePointCloud((scans[2][2], scans[2][3]), scans[2][0], scans[2][1])
    fi, fv, _ = _snapshot(fresh)
    killed = {tuple(k) for k in kill.tolist()}
    for k, bi in enumerate(i3.tolist()):
        if tuple(bi) in killed:  # a re-created block only holds the last scan
            j = [tuple(x) for x in fi.tolist()].index(tuple(bi))
            assert v3[k].tobytes() == fv[j].tobytes()
    a.removeAllBlocks()                          # Layer::removeAllBlocks, core/layer.h:164
    assert a.getNumberOfAllocatedBlocks() == 0
    ia.integratePointCloud((scans[2][2], scans[2][3]), scans[2][0], scans[2][1])
    i4, v4, _ = _snapshot(a)
    assert (i4 == fi).all() and v4.tobytes() == fv.tobytes()


def test_mirror_updated_equals_list_plus_download():
    """vbx_mirror_updated = getAllUpdatedBlocks(bit) + block payloads + updated().reset(bit) in one
    call: same blocks, same bytes, same order; staging and direct
    (page-locked) destinations; the clear mask only touches the mirrored bit."""
    cfg = vb.TsdfIntegratorConfig(default_truncation_distance=0.4, integrator_threads=1)
    layer = vb.Layer(0.1, 16)
    integ = vb.TsdfIntegratorFactory.create("merged", cfg, layer)
    scans = scenes.c3_room_sequence(n_scans=3, width=96, height=72)
    integ.integratePointCloud((scans[0][2], scans[0][3]), scans[0][0], scans[0][1])
    MESH, ESDF = 2, 4
    want_idx = layer.getAllUpdatedBlocks(1)          # Update::kMesh
    want_vox, want_upd = layer.getBlocks(want_idx)
    idx, vox, upd = layer.mirrorUpdated(MESH, MESH)
    assert idx.tobytes() == want_idx.tobytes()
    assert vox.tobytes() == want_vox.tobytes()
    assert upd.tobytes() == np.asarray(want_upd).tobytes()
    # the mesh bit is gone, the others stay
    assert layer.getAllUpdatedBlocks(1).shape[0] == 0
    assert layer.getAllUpdatedBlocks(2).tobytes() == want_idx.tobytes()
    # next scan: only the blocks it touched come back, into a page-locked buffer, no staging
    integ.integratePointCloud((scans[1][2], scans[1][3]), scans[1][0], scans[1][1])
    want_idx = layer.getAllUpdatedBlocks(1)
    want_vox, _ = layer.getBlocks(want_idx)
    pinned = layer.hostBuffer((want_idx.shape[0] + 3, 4096), vb.TSDF_DTYPE)
    idx, vox, upd = layer.mirrorUpdated(MESH, MESH, voxels_out=pinned)
    assert idx.tobytes() == want_idx.tobytes() and vox.tobytes() == want_vox.tobytes()
    assert np.shares_memory(vox, pinned)
    # mask 0 = every block; nothing is cleared without a clear mask
    idx_all, vox_all, _ = layer.mirrorUpdated(0, 0)
    assert idx_all.tobytes() == layer.getAllAllocatedBlocks().tobytes()
    assert vox_all.tobytes() == layer.getBlocks(idx_all)[0].tobytes()
    assert layer.getAllUpdatedBlocks(2).shape[0] > 0
    # ESDF layer
    esdf = vb.Layer(0.1, 16, voxel_type="esdf")
    e = vb.EsdfIntegrator(vb.EsdfIntegratorConfig(min_distance_m=0.2), layer, esdf)
    e.updateFromTsdfLayer(True)
    eidx, evox, _ = esdf.mirrorUpdated(0, 0)
    assert eidx.tobytes() == esdf.getAllAllocatedBlocks().tobytes()
    assert evox.tobytes() == esdf.getBlocks(eidx)[0].tobytes()
