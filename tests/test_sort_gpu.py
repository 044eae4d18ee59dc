"""-m gpu: the engine's own radix sort (device-side element count, skipped uniform passes) and
exclusive scan against numpy."""
import ctypes as C

import numpy as np
import pytest

import voxblox_b200 as vb

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    layer = vb.Layer(0.1, 16, engine_options=vb.EngineOptions(max_blocks=1024, max_points_per_scan=1 << 20,
                                                              max_updates_per_pass=1 << 22))
    vb.TsdfIntegratorFactory.create("merged", vb.TsdfIntegratorConfig(), layer)
    c = layer._ctx
    c.lib.vbx_debug_sort.restype = C.c_int
    c.lib.vbx_debug_sort.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]
    c.lib.vbx_debug_scan.restype = C.c_int
    c.lib.vbx_debug_scan.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    return c


def _sort(ctx, keys, bits):
    out_k = np.zeros_like(keys)
    perm = np.zeros(len(keys), dtype=np.uint32)
    ctx.check(ctx.lib.vbx_debug_sort(ctx.handle, keys.ctypes.data, keys.dtype.itemsize, len(keys), bits,
                                     out_k.ctypes.data, perm.ctypes.data), "vbx_debug_sort")
    return out_k, perm


@pytest.mark.parametrize("dtype,bits", [(np.uint32, 32), (np.uint32, 19), (np.uint32, 8), (np.uint64, 64),
                                        (np.uint64, 28), (np.uint64, 41)])
@pytest.mark.parametrize("n", [0, 1, 31, 4096, 4097, 100003, 1 << 20])
def test_radix_sort_is_a_stable_sort(ctx, dtype, bits, n):
    rng = np.random.default_rng(n * 131 + bits)
    hi = (1 << bits) - 1
    keys = rng.integers(0, hi, size=n, dtype=np.uint64, endpoint=True).astype(dtype)
    if n > 1000:
        keys[: n // 3] &= dtype(0xFF)  # many duplicates: stability matters
    got_k, perm = _sort(ctx, keys, bits)
    want = np.argsort(keys, kind="stable").astype(np.uint32)
    assert (got_k == keys[want]).all()
    assert (perm == want).all()


def test_radix_sort_more_tiles_than_thread_blocks(ctx):
    """1024 tiles on a grid of at most two thread blocks per SM: every block takes several tickets per pass and
    waits for the other blocks' tiles between the passes (the whole sort is one launch)."""
    n = (1 << 22) - 5
    rng = np.random.default_rng(11)
    keys = rng.integers(0, (1 << 22) - 1, size=n, dtype=np.uint32)
    keys[: n // 4] &= np.uint32(0x3FF)
    got_k, perm = _sort(ctx, keys, 32)
    want = np.argsort(keys, kind="stable").astype(np.uint32)
    assert (got_k == keys[want]).all()
    assert (perm == want).all()


def test_radix_sort_skips_uniform_digits(ctx):
    # keys that only differ in bits 8..15: three of the four passes are identity permutations
    rng = np.random.default_rng(7)
    keys = (rng.integers(0, 256, size=50000, dtype=np.uint32) << 8) | np.uint32(0xAB0000CD)
    got_k, perm = _sort(ctx, keys, 32)
    want = np.argsort(keys, kind="stable").astype(np.uint32)
    assert (perm == want).all() and (got_k == keys[want]).all()


@pytest.mark.parametrize("n", [1, 2047, 2048, 2049, 300001, (1 << 20) + 1])
def test_exclusive_scan(ctx, n):
    rng = np.random.default_rng(n)
    v = rng.integers(0, 300, size=n, dtype=np.uint32)
    out = np.zeros(n, dtype=np.uint32)
    ctx.check(ctx.lib.vbx_debug_scan(ctx.handle, v.ctypes.data, n, out.ctypes.data), "vbx_debug_scan")
    want = np.concatenate([[0], np.cumsum(v[:-1], dtype=np.uint64)]).astype(np.uint32)
    assert (out == want).all()
