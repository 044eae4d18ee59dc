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


def test_serialized_blocks_match_reference_format():
    """vbx_serialize_updated packs blocks on the device in Block::serializeToIntegers' layout
    (src/core/block.cc:159-183 TSDF, :203-234 ESDF incl. the serializeDirection sign-extension
    quirk): word for word equal to the oracle's; vbx_deserialize_blocks is its inverse
    (block.cc:65-90,110-135) -- SURVEY.md section 8f N2."""
    from oracle import pyoracle as po

    cfg = vb.TsdfIntegratorConfig(default_truncation_distance=0.4, integrator_threads=1)
    ekw = dict(max_distance_m=2.0, default_distance_m=2.0, min_distance_m=0.2, min_diff_m=0.0)
    layer = vb.Layer(0.1, 16)
    integ = vb.TsdfIntegratorFactory.create("merged", cfg, layer)
    esdf = vb.Layer(0.1, 16, voxel_type="esdf")
    eint = vb.EsdfIntegrator(vb.EsdfIntegratorConfig(**ekw), layer, esdf)
    omap = po.OracleMap(po.OracleLib("port"), po.TsdfConfig(default_truncation_distance=0.4), 0.1, 16)
    omap.esdf_create(po.EsdfConfig(**ekw))
    for s in scenes.c3_room_sequence(n_scans=2, width=96, height=72):
        integ.integratePointCloud((s[2], s[3]), s[0], s[1])
        omap.integrate(2, s, order=po.ORDER_REFERENCE)
    eint.updateFromTsdfLayer(False)
    # TSDF: the device map is bit-identical to the oracle's, so the words must be too
    idx, words, upd = layer.serializeUpdated(0, 0)
    assert idx.tobytes() == omap.block_indices(0).tobytes()
    for k, i in enumerate(idx):
        assert words[k].tobytes() == omap.serialize_block(i, 0).tobytes(), tuple(i)
    # ESDF: serialise the DEVICE's voxels with the oracle's packer (load them into a scratch map)
    scratch = po.OracleMap(po.OracleLib("port"), po.TsdfConfig(default_truncation_distance=0.4), 0.1, 16)
    scratch.esdf_create(po.EsdfConfig(**ekw))
    eidx, ewords, _ = esdf.serializeUpdated(0, 0)
    evox, _ = esdf.getBlocks(eidx)
    assert (evox["parent"] < 0).any()
    for k, i in enumerate(eidx):
        # oracle round trip of the device words = what the reference would decode ...
        scratch.deserialize_block(i, ewords[k], 1)
        dec = scratch.block(i, 1)[0]
        # ... and re-encoding that is a fixed point, equal to the device's words
        assert scratch.serialize_block(i, 1).tobytes() == ewords[k].tobytes(), tuple(i)
        assert dec["distance"].tobytes() == evox[k]["distance"].tobytes()
        for f in ("observed", "hallucinated", "in_queue", "fixed"):
            assert (dec[f] == evox[k][f]).all()
        # block.cc:8-41,203-234 restated in numpy on the raw device voxels (int64 shifts sign-extend
        # like the reference's `int8 << n`)
        par = evox[k]["parent"].astype(np.int64).clip(-128, 127)
        w2 = ((par[:, 0] << 24) | (par[:, 1] << 16) | (par[:, 2] << 8)) & 0xFFFFFFFF
        w2 |= (evox[k]["observed"] != 0) * 1 | (evox[k]["hallucinated"] != 0) * 2 | (evox[k]["in_queue"] != 0) * 4 | \
            (evox[k]["fixed"] != 0) * 8
        want = np.stack([evox[k]["distance"].view(np.uint32), w2.astype(np.uint32)], axis=1).reshape(-1)
        assert want.tobytes() == ewords[k].tobytes(), tuple(i)
    # deserialise on the device: a fresh map fed the words must hold what the reference decodes
    layer2 = vb.Layer(0.1, 16)
    integ2 = vb.TsdfIntegratorFactory.create("merged", cfg, layer2)
    esdf2 = vb.Layer(0.1, 16, voxel_type="esdf")
    vb.EsdfIntegrator(vb.EsdfIntegratorConfig(**ekw), layer2, esdf2)
    layer2.insertSerializedBlocks(idx, words, upd)
    esdf2.insertSerializedBlocks(eidx, ewords)
    assert layer2.getAllAllocatedBlocks().tobytes() == idx.tobytes()
    assert layer2.getBlocks(idx)[0].tobytes() == layer.getBlocks(idx)[0].tobytes()       # TSDF is lossless
    assert np.asarray(layer2.getBlocks(idx)[1]).tobytes() == np.asarray(upd).tobytes()
    got, _ = esdf2.getBlocks(eidx)
    for k, i in enumerate(eidx):
        want = scratch.block(i, 1)[0]
        assert got[k].tobytes() == want.tobytes(), tuple(i)
    # continuing to integrate on the reloaded map equals continuing on the original
    s = scenes.c3_room_sequence(n_scans=3, width=96, height=72)[2]
    integ.integratePointCloud((s[2], s[3]), s[0], s[1])
    integ2.integratePointCloud((s[2], s[3]), s[0], s[1])
    a = layer.getAllAllocatedBlocks()
    assert a.tobytes() == layer2.getAllAllocatedBlocks().tobytes()
    assert layer.getBlocks(a)[0].tobytes() == layer2.getBlocks(a)[0].tobytes()


def test_save_and_load_layer_file(tmp_path):
    """Layer::saveToFile / io::LoadBlocksFromFile through the device map: TSDF and ESDF layers in one
    .vxblx file (clear_file = false appends, core/layer_inl.h:96-103), reloaded into fresh layers;
    the file is also walked with a real protobuf parser."""
    from tests.test_proto_io import _messages

    cfg = vb.TsdfIntegratorConfig(default_truncation_distance=0.4, integrator_threads=1)
    layer = vb.Layer(0.1, 16)
    integ = vb.TsdfIntegratorFactory.create("merged", cfg, layer)
    esdf = vb.Layer(0.1, 16, voxel_type="esdf")
    eint = vb.EsdfIntegrator(vb.EsdfIntegratorConfig(min_distance_m=0.2), layer, esdf)
    for s in scenes.c3_room_sequence(n_scans=2, width=96, height=72):
        integ.integratePointCloud((s[2], s[3]), s[0], s[1])
    eint.updateFromTsdfLayer(True)
    path = str(tmp_path / "map.vxblx")
    assert layer.saveToFile(path, True)
    assert esdf.saveToFile(path, False)
    # walk the file with python protobuf
    BlockProto, LayerProto = _messages()
    data = open(path, "rb").read()

    def varint(pos):
        v, shift = 0, 0
        while True:
            b = data[pos]
            pos += 1
            v |= (b & 0x7F) << shift
            shift += 7
            if not b & 0x80:
                return v, pos
    pos = 0
    idx, words, _ = layer.serializeUpdated(0, 0)
    eidx, ewords, _ = esdf.serializeUpdated(0, 0)
    for typ, want_idx, want_words in (("tsdf", idx, words), ("esdf", eidx, ewords)):
        count, pos = varint(pos)
        assert count == 1 + want_idx.shape[0]
        size, pos = varint(pos)
        lm = LayerProto()
        lm.ParseFromString(data[pos:pos + size])
        pos += size
        assert lm.type == typ and lm.voxels_per_side == 16 and lm.voxel_size == float(np.float32(0.1))
        for k in range(want_idx.shape[0]):
            size, pos = varint(pos)
            bm = BlockProto()
            bm.ParseFromString(data[pos:pos + size])
            pos += size
            bs = np.float32(0.1) * np.float32(16)
            assert (bm.origin_x, bm.origin_y, bm.origin_z) == tuple(float(np.float32(v) * bs) for v in want_idx[k])
            assert bm.voxels_per_side == 16 and not bm.has_data
            assert np.asarray(bm.voxel_data, dtype=np.uint32).tobytes() == want_words[k].tobytes()
    assert pos == len(data)
    # reload
    layer2 = vb.Layer(0.1, 16)
    vb.TsdfIntegratorFactory.create("merged", cfg, layer2)
    esdf2 = vb.Layer(0.1, 16, voxel_type="esdf")
    vb.EsdfIntegrator(vb.EsdfIntegratorConfig(min_distance_m=0.2), layer2, esdf2)
    assert layer2.loadBlocksFromFile(path) == idx.shape[0]
    assert esdf2.loadBlocksFromFile(path) == eidx.shape[0]     # second layer of the file
    assert layer2.getAllAllocatedBlocks().tobytes() == idx.tobytes()
    assert layer2.getBlocks(idx)[0].tobytes() == layer.getBlocks(idx)[0].tobytes()
    assert (np.asarray(layer2.getBlocks(idx)[1]) == 7).all()   # updated().set(), layer_inl.h:227
    assert esdf2.serializeUpdated(0, 0)[1].tobytes() == ewords.tobytes()
    # an incompatible layer is refused (isCompatible, layer_inl.h:237-260)
    other = vb.Layer(0.2, 16)
    vb.TsdfIntegratorFactory.create("merged", cfg, other)
    with pytest.raises(vb.VoxbloxError):
        other.loadBlocksFromFile(path)


def test_loaded_file_with_duplicate_blocks_and_has_data(tmp_path):
    """io::LoadBlocksFromFile with kReplace (io/layer_io_inl.h:71-73): a block listed twice keeps the LAST
    payload; BlockProto.has_data (core/block_inl.h:73-109) survives a load -> save round trip, blocks the
    integrator makes keep has_data = false, and removing a block forgets the flag."""
    from tests.test_proto_io import _encode_block, _messages
    import ctypes as C

    lib = vb.load_library()
    BlockProto, LayerProto = _messages()
    vs, vps = float(np.float32(0.1)), 16
    bs = float(np.float32(0.1) * np.float32(16))
    rng = np.random.default_rng(5)
    n_words = 3 * vps ** 3

    def payload():
        w = np.zeros((vps ** 3, 3), dtype=np.uint32)
        w[:, 0] = rng.uniform(-0.4, 0.4, vps ** 3).astype(np.float32).view(np.uint32)
        w[:, 1] = rng.uniform(0.5, 9.0, vps ** 3).astype(np.float32).view(np.uint32)
        w[:, 2] = rng.integers(0, 2 ** 32, vps ** 3, dtype=np.uint32)
        return w.reshape(-1)
    first, second, third = payload(), payload(), payload()
    blocks = [((1, 0, 0), True, first), ((0, 2, -1), False, second), ((1, 0, 0), True, third)]  # (1,0,0) twice

    def varint_bytes(v):
        out = bytearray()
        while True:
            b = v & 0x7F
            v >>= 7
            out.append(b | (0x80 if v else 0))
            if not v:
                return bytes(out)
    n = C.c_uint64(0)
    hdr = np.zeros(64, dtype=np.uint8)
    assert lib.vbx_proto_encode_layer(vs, vps, b"tsdf", hdr.ctypes.data, hdr.size, C.byref(n)) == 0
    data = varint_bytes(1 + len(blocks)) + varint_bytes(n.value) + hdr[:n.value].tobytes()
    for idx, has, w in blocks:
        msg = _encode_block(lib, vps, vs, [i * bs for i in idx], has, w)
        data += varint_bytes(len(msg)) + msg
    path = str(tmp_path / "dup.vxblx")
    open(path, "wb").write(data)

    cfg = vb.TsdfIntegratorConfig(default_truncation_distance=0.4, integrator_threads=1)
    layer = vb.Layer(0.1, 16)
    integ = vb.TsdfIntegratorFactory.create("merged", cfg, layer)
    assert layer.loadBlocksFromFile(path) == 2
    idx, words, _ = layer.serializeUpdated(0, 0)
    got = {tuple(int(v) for v in idx[k]): words[k] for k in range(idx.shape[0])}
    assert set(got) == {(1, 0, 0), (0, 2, -1)}
    assert got[(1, 0, 0)].tobytes() == third.tobytes()       # the last payload, whole
    assert got[(0, 2, -1)].tobytes() == second.tobytes()
    # the integrator adds blocks of its own: has_data stays false on them
    s = scenes.c3_room_sequence(n_scans=1, width=64, height=48)[0]
    integ.integratePointCloud((s[2], s[3]), s[0], s[1])

    def saved_flags(p):
        layer.saveToFile(p, True)
        raw = open(p, "rb").read()

        def varint(pos):
            v, shift = 0, 0
            while True:
                b = raw[pos]
                pos += 1
                v |= (b & 0x7F) << shift
                shift += 7
                if not b & 0x80:
                    return v, pos
        count, pos = varint(0)
        size, pos = varint(pos)
        pos += size
        flags = {}
        for _ in range(count - 1):
            size, pos = varint(pos)
            bm = BlockProto()
            bm.ParseFromString(raw[pos:pos + size])
            pos += size
            flags[tuple(int(round(v / bs)) for v in (bm.origin_x, bm.origin_y, bm.origin_z))] = bool(bm.has_data)
        return flags
    flags = saved_flags(str(tmp_path / "resaved.vxblx"))
    assert flags[(1, 0, 0)] is True and flags[(0, 2, -1)] is False
    assert sum(flags.values()) == 1 and len(flags) > 2
    # Layer::removeBlock + a new block at the same index: a fresh Block, has_data false
    layer.removeBlock((1, 0, 0))
    layer.insertBlocks(np.array([[1, 0, 0]], dtype=np.int32), layer.getBlocks(np.array([[0, 2, -1]], dtype=np.int32))[0])
    assert saved_flags(str(tmp_path / "resaved2.vxblx"))[(1, 0, 0)] is False


def test_esdf_only_blocks_stay_out_of_the_tsdf_layer():
    """Blocks inserted into the ESDF layer at indices the TSDF layer does not hold (what
    addNewRobotPosition and loading an ESDF map do) occupy pool slots of their own: the TSDF
    layer's listings, downloads, mirror, serialisation and block count do not see them, the
    clear-updated call leaves them alone, and the first TSDF integration into such an index
    starts from an empty block like the reference's allocateBlockPtrByIndex."""
    cfg = vb.TsdfIntegratorConfig(default_truncation_distance=0.4, integrator_threads=1)
    scans = scenes.c3_room_sequence(n_scans=1, width=96, height=72)
    s = scans[0]
    # what the scan alone produces
    ref_layer = vb.Layer(0.1, 16)
    vb.TsdfIntegratorFactory.create("merged", cfg, ref_layer).integratePointCloud((s[2], s[3]), s[0], s[1])
    want_idx, want_vox, want_upd = _snapshot(ref_layer)
    # same scan into a map that first received ESDF-only blocks: two at indices the scan will
    # touch, one far away
    layer = vb.Layer(0.1, 16)
    integ = vb.TsdfIntegratorFactory.create("merged", cfg, layer)
    esdf = vb.Layer(0.1, 16, voxel_type="esdf")
    vb.EsdfIntegrator(vb.EsdfIntegratorConfig(min_distance_m=0.2), layer, esdf)
    eidx = np.concatenate([want_idx[:2], np.array([[40, -40, 7]], np.int32)])
    evox = np.zeros((3, 4096), vb.ESDF_DTYPE)
    evox["distance"] = 1.25
    evox["observed"] = 1
    esdf.insertBlocks(eidx, evox)
    assert layer.getNumberOfAllocatedBlocks() == 0 and len(layer.getAllAllocatedBlocks()) == 0
    assert layer.mirrorUpdated(0, 0)[0].shape[0] == 0
    assert esdf.getNumberOfAllocatedBlocks() == 3
    got_e, _ = esdf.getBlocks(esdf.getAllAllocatedBlocks())
    assert (got_e["distance"] == 1.25).all()
    with pytest.raises(vb.VoxbloxError):
        layer.getBlocks(eidx[:1])                  # not a TSDF block
    layer.clearUpdated(7)                          # must not disturb the internal marks
    assert layer.getNumberOfAllocatedBlocks() == 0
    integ.integratePointCloud((s[2], s[3]), s[0], s[1])
    got_idx, got_vox, got_upd = _snapshot(layer)
    assert (got_idx == want_idx).all()
    assert got_vox.tobytes() == want_vox.tobytes() and (got_upd == want_upd).all()
    assert layer.getNumberOfAllocatedBlocks() == len(want_idx)
    assert esdf.getNumberOfAllocatedBlocks() == 3   # the ESDF blocks are still there, untouched
    got_e, _ = esdf.getBlocks(esdf.getAllAllocatedBlocks())
    assert (got_e["distance"] == 1.25).all()


def test_layers_are_removed_independently():
    """Layer::removeBlock / removeAllBlocks act on ONE layer (core/layer.h:163-164): removing TSDF
    blocks keeps the ESDF blocks at those indices and vice versa; a pool slot is recycled once both
    are gone, and integration afterwards starts from empty blocks."""
    cfg = vb.TsdfIntegratorConfig(default_truncation_distance=0.4, integrator_threads=1)
    scans = scenes.c3_room_sequence(n_scans=2, width=96, height=72)
    layer = vb.Layer(0.1, 16)
    integ = vb.TsdfIntegratorFactory.create("merged", cfg, layer)
    esdf = vb.Layer(0.1, 16, voxel_type="esdf")
    eint = vb.EsdfIntegrator(vb.EsdfIntegratorConfig(min_distance_m=0.2), layer, esdf)
    s = scans[0]
    integ.integratePointCloud((s[2], s[3]), s[0], s[1])
    eint.updateFromTsdfLayer(True)
    idx, vox, _ = _snapshot(layer)
    eidx, evox, _ = _snapshot(esdf)
    assert (idx == eidx).all()
    kill = idx[::2]
    keep = np.array([i for i in range(len(idx)) if i % 2 == 1])
    layer.removeBlocks(kill)
    i2, v2, _ = _snapshot(layer)
    assert (i2 == idx[keep]).all() and v2.tobytes() == vox[keep].tobytes()
    e2, ev2, _ = _snapshot(esdf)                       # the ESDF layer is untouched
    assert (e2 == eidx).all() and ev2.tobytes() == evox.tobytes()
    esdf.removeBlocks(kill[:1])                        # now that slot is free in both layers
    e3, ev3, _ = _snapshot(esdf)
    assert (e3 == eidx[1:]).all() and ev3.tobytes() == evox[1:].tobytes()
    esdf.removeBlocks(idx[keep][:1])                   # ESDF gone, TSDF stays
    i4, v4, _ = _snapshot(layer)
    assert (i4 == idx[keep]).all() and v4.tobytes() == vox[keep].tobytes()
    assert esdf.getNumberOfAllocatedBlocks() == len(eidx) - 2
    # removeAllBlocks on the TSDF layer keeps the remaining ESDF blocks
    esdf_before = _snapshot(esdf)
    layer.removeAllBlocks()
    assert layer.getNumberOfAllocatedBlocks() == 0
    esdf_after = _snapshot(esdf)
    assert (esdf_after[0] == esdf_before[0]).all() and esdf_after[1].tobytes() == esdf_before[1].tobytes()
    # integrating again: the result equals a fresh map's (the freed / ESDF-only slots read as new blocks)
    s1 = scans[1]
    integ.integratePointCloud((s1[2], s1[3]), s1[0], s1[1])
    fresh = vb.Layer(0.1, 16)
    vb.TsdfIntegratorFactory.create("merged", cfg, fresh).integratePointCloud((s1[2], s1[3]), s1[0], s1[1])
    a, b = _snapshot(layer), _snapshot(fresh)
    assert (a[0] == b[0]).all() and a[1].tobytes() == b[1].tobytes()
    esdf.removeAllBlocks()
    assert esdf.getNumberOfAllocatedBlocks() == 0 and layer.getNumberOfAllocatedBlocks() == len(b[0])
