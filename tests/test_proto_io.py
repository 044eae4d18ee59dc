"""The .vxblx wire format (vbx_io.cu) against a real protobuf implementation (python protobuf,
messages built from the field lists of the reference's proto/voxblox/Block.proto and Layer.proto).
Host-only: no GPU needed."""
import ctypes as C

import numpy as np
import pytest

from voxblox_b200 import api

pb = pytest.importorskip("google.protobuf")
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory  # noqa: E402


def _messages():
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "voxblox_test.proto"
    fd.package = "voxblox"
    fd.syntax = "proto2"
    T = descriptor_pb2.FieldDescriptorProto
    blk = fd.message_type.add()
    blk.name = "BlockProto"   # proto/voxblox/Block.proto
    for name, num, typ, label in [("voxels_per_side", 1, T.TYPE_INT32, T.LABEL_OPTIONAL),
                                  ("voxel_size", 2, T.TYPE_DOUBLE, T.LABEL_OPTIONAL),
                                  ("origin_x", 3, T.TYPE_DOUBLE, T.LABEL_OPTIONAL),
                                  ("origin_y", 4, T.TYPE_DOUBLE, T.LABEL_OPTIONAL),
                                  ("origin_z", 5, T.TYPE_DOUBLE, T.LABEL_OPTIONAL),
                                  ("has_data", 6, T.TYPE_BOOL, T.LABEL_OPTIONAL),
                                  ("voxel_data", 7, T.TYPE_UINT32, T.LABEL_REPEATED)]:
        f = blk.field.add()
        f.name, f.number, f.type, f.label = name, num, typ, label
    lay = fd.message_type.add()
    lay.name = "LayerProto"   # proto/voxblox/Layer.proto
    for name, num, typ in [("voxel_size", 1, T.TYPE_DOUBLE), ("voxels_per_side", 2, T.TYPE_UINT32),
                           ("type", 3, T.TYPE_STRING)]:
        f = lay.field.add()
        f.name, f.number, f.type, f.label = name, num, typ, T.LABEL_OPTIONAL
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = message_factory.GetMessageClass
    return get(pool.FindMessageTypeByName("voxblox.BlockProto")), get(pool.FindMessageTypeByName("voxblox.LayerProto"))


def _encode_block(lib, vps, voxel_size, origin, has_data, words):
    o = np.asarray(origin, dtype=np.float64)
    w = np.ascontiguousarray(words, dtype=np.uint32)
    n = C.c_uint64(0)
    assert lib.vbx_proto_encode_block(vps, voxel_size, o.ctypes.data, int(has_data), w.ctypes.data, w.size, None, 0,
                                      C.byref(n)) == 0
    out = np.zeros(n.value, dtype=np.uint8)
    assert lib.vbx_proto_encode_block(vps, voxel_size, o.ctypes.data, int(has_data), w.ctypes.data, w.size,
                                      out.ctypes.data, out.size, C.byref(n)) == 0
    return out.tobytes()


def _decode_block(lib, msg, n_words):
    buf = np.frombuffer(msg, dtype=np.uint8)
    vps, has = C.c_int32(0), C.c_int(0)
    vs = C.c_double(0)
    origin = np.zeros(3, dtype=np.float64)
    words = np.zeros(n_words, dtype=np.uint32)
    n = C.c_uint64(0)
    rc = lib.vbx_proto_decode_block(buf.ctypes.data, buf.size, C.byref(vps), C.byref(vs), origin.ctypes.data,
                                    C.byref(has), words.ctypes.data, words.size, C.byref(n))
    assert rc == 0 and n.value == n_words
    return vps.value, vs.value, origin, has.value, words


def test_wire_format_equals_protobuf():
    lib = api.load_library()
    BlockProto, LayerProto = _messages()
    rng = np.random.default_rng(0)
    # words spanning every varint length, incl. float bit patterns with the top bit set
    words = np.concatenate([np.array([0, 1, 127, 128, 16383, 16384, 2 ** 21, 2 ** 28, 2 ** 32 - 1], dtype=np.uint32),
                            rng.integers(0, 2 ** 32, 4096 * 3 - 9, dtype=np.uint64).astype(np.uint32)])
    for vps, vs, origin, has in [(16, float(np.float32(0.05)), (-1.6, 0.0, 3.2), False),
                                 (8, float(np.float32(0.1)), (0.0, -0.8, 0.0), True)]:
        m = BlockProto()
        m.voxels_per_side, m.voxel_size = vps, vs
        m.origin_x, m.origin_y, m.origin_z = origin
        m.has_data = has
        m.voxel_data.extend(int(w) for w in words)
        ours = _encode_block(lib, vps, vs, origin, has, words)
        assert ours == m.SerializeToString()          # byte for byte what libprotobuf writes
        back = BlockProto()
        back.ParseFromString(ours)
        assert back == m
        d = _decode_block(lib, m.SerializeToString(), words.size)
        assert d[0] == vps and d[1] == vs and tuple(d[2]) == origin and d[3] == int(has)
        assert d[4].tobytes() == words.tobytes()
    # the parser also accepts the packed encoding of voxel_data (what a proto3 writer would emit)
    packed = b"".join(_varint(int(w)) for w in words)
    msg = b"\x08\x10" + b"\x3a" + _varint(len(packed)) + packed
    d = _decode_block(lib, msg, words.size)
    assert d[0] == 16 and d[4].tobytes() == words.tobytes()
    # layer header
    lm = LayerProto()
    lm.voxel_size, lm.voxels_per_side, lm.type = float(np.float32(0.05)), 16, "tsdf"
    n = C.c_uint64(0)
    out = np.zeros(64, dtype=np.uint8)
    assert lib.vbx_proto_encode_layer(lm.voxel_size, 16, b"tsdf", out.ctypes.data, out.size, C.byref(n)) == 0
    assert out[: n.value].tobytes() == lm.SerializeToString()


def _varint(v):
    o = bytearray()
    while v >= 0x80:
        o.append((v & 0x7F) | 0x80)
        v >>= 7
    o.append(v)
    return bytes(o)
