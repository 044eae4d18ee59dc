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
