"""-m gpu: the Merged integrator's bundle order kernel against the C++ library itself.

k_bundle_order (voxblox_b200/csrc/vbx_order.cuh) must reproduce the iteration order of the
std::unordered_map that MergedTsdfIntegrator::bundleRays fills (tsdf_integrator.cc:318-322,
340-371) and integrateVoxels walks (cc:436-456).  oracle vbo_umap_order builds a real
std::unordered_map with preset hash values; the device gets the same hashes."""
import ctypes as C

import numpy as np
import pytest

import voxblox_b200 as vb
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    layer = vb.Layer(0.1, 16, engine_options=vb.EngineOptions(max_blocks=1024, max_points_per_scan=1 << 18,
                                                              max_updates_per_pass=1 << 20))
    vb.TsdfIntegratorFactory.create("merged", vb.TsdfIntegratorConfig(), layer)
    c = layer._ctx
    c.lib.vbx_debug_bundle_order.restype = C.c_int
    c.lib.vbx_debug_bundle_order.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p]
    return c


def _device_order(ctx, hashes, force_global):
    out = np.zeros(len(hashes), dtype=np.uint32)
    ctx.check(ctx.lib.vbx_debug_bundle_order(ctx.handle, hashes.ctypes.data, len(hashes), int(force_global),
                                             out.ctypes.data), "vbx_debug_bundle_order")
    return out


# sizes around every rehash threshold of libstdc++'s policy (13, 29, 59, 127, 257, 541, 1109, 2357, 5087,
# 10273, 20753, 42043, 85229) and sizes in between
SIZES = [1, 2, 12, 13, 14, 28, 29, 30, 59, 60, 127, 128, 257, 258, 541, 542, 1000, 1109, 1110, 2357, 2358, 5087, 5088,
         5300, 10273, 10274, 20753, 20754, 42043, 42044, 49152, 85229, 85230, 131072]


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("kind", ["random", "voxel_hash", "colliding"])
def test_bundle_order_equals_unordered_map(ctx, n, kind):
    if kind == "colliding" and n > 6000:
        pytest.skip("seven buckets' worth of chains: quadratic work, covered at the smaller sizes")
    rng = np.random.default_rng(n)
    if kind == "random":
        h = rng.integers(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)
    elif kind == "voxel_hash":
        # LongIndexHash of voxels on a surface patch (core/block_hash.h:52-64): strongly structured
        x = rng.integers(-60, 60, size=n).astype(np.int64)
        y = rng.integers(-60, 60, size=n).astype(np.int64)
        z = rng.integers(-3, 3, size=n).astype(np.int64)
        h = ((x + y * 17191 + z * 17191 * 17191) & 0xffffffff).astype(np.uint32)
    else:
        h = (rng.integers(0, 7, size=n, dtype=np.uint64) * 5087).astype(np.uint32)  # long bucket chains
    ref = po.umap_order(po.OracleLib("port"), h)
    for force_global in (0, 1):
        got = _device_order(ctx, h, force_global)
        assert (got == ref).all(), (n, kind, force_global, int((got != ref).sum()))
