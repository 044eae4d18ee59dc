"""TEST INFRASTRUCTURE ONLY: ctypes front end for the two CPU checkers.

  OracleLib("port")       -> oracle/libvbx_oracle.so      (restatement, oracle/vbx_oracle.cc)
  OracleLib("reference")  -> oracle/_ref/libvbx_ref.so    (the reference's own sources)

Both export oracle/vbo_api.h.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs import this module; the product package
voxblox_b200 never does (tests/test_layout.py enforces it).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libvbx_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libvbx_ref.so")
REF_O3_SO = os.path.join(HERE, "_ref", "libvbx_ref_o3.so")  # timing-only build (-O3 -march=x86-64-v3)
_PATHS = {"port": PORT_SO, "reference": REF_SO, "reference_o3": REF_O3_SO}

TSDF_DTYPE = np.dtype([("distance", "<f4"), ("weight", "<f4"), ("color", "u1", (4,))])
ESDF_DTYPE = np.dtype([("distance", "<f4"), ("observed", "u1"), ("hallucinated", "u1"),
                       ("in_queue", "u1"), ("fixed", "u1"), ("parent", "<i4", (3,))])
assert TSDF_DTYPE.itemsize == 12 and ESDF_DTYPE.itemsize == 20

SIMPLE, MERGED, FAST = 1, 2, 3
ORDER_REFERENCE = 0
LAYER_TSDF, LAYER_ESDF = 0, 1


class TsdfConfig(C.Structure):
    """vbo_tsdf_config; defaults = TsdfIntegratorBase::Config (tsdf_integrator.h:56-89)."""
    _fields_ = [("default_truncation_distance", C.c_float), ("max_weight", C.c_float),
                ("voxel_carving_enabled", C.c_int32), ("min_ray_length_m", C.c_float),
                ("max_ray_length_m", C.c_float), ("use_const_weight", C.c_int32),
                ("allow_clear", C.c_int32), ("use_weight_dropoff", C.c_int32),
                ("use_sparsity_compensation_factor", C.c_int32),
                ("sparsity_compensation_factor", C.c_float), ("integrator_threads", C.c_int32),
                ("integration_order_mode", C.c_int32), ("enable_anti_grazing", C.c_int32),
                ("start_voxel_subsampling_factor", C.c_float),
                ("max_consecutive_ray_collisions", C.c_int32),
                ("clear_checks_every_n_frames", C.c_int32), ("max_integration_time_s", C.c_float)]

    def __init__(self, **kw):
        d = dict(default_truncation_distance=0.1, max_weight=10000.0, voxel_carving_enabled=1,
                 min_ray_length_m=0.1, max_ray_length_m=5.0, use_const_weight=0, allow_clear=1,
                 use_weight_dropoff=1, use_sparsity_compensation_factor=0,
                 sparsity_compensation_factor=1.0, integrator_threads=1, integration_order_mode=0,
                 enable_anti_grazing=0, start_voxel_subsampling_factor=2.0,
                 max_consecutive_ray_collisions=2, clear_checks_every_n_frames=1,
                 max_integration_time_s=3.4028234663852886e38)
        d.update(kw)
        super().__init__(**d)


class EsdfConfig(C.Structure):
    """vbo_esdf_config; defaults = EsdfIntegrator::Config (esdf_integrator.h:29-78)."""
    _fields_ = [("full_euclidean_distance", C.c_int32), ("max_distance_m", C.c_float),
                ("min_distance_m", C.c_float), ("default_distance_m", C.c_float),
                ("min_diff_m", C.c_float), ("min_weight", C.c_float), ("num_buckets", C.c_int32),
                ("multi_queue", C.c_int32), ("add_occupied_crust", C.c_int32),
                ("clear_sphere_radius", C.c_float), ("occupied_sphere_radius", C.c_float)]

    def __init__(self, **kw):
        d = dict(full_euclidean_distance=0, max_distance_m=2.0, min_distance_m=0.2,
                 default_distance_m=2.0, min_diff_m=0.001, min_weight=1e-6, num_buckets=20,
                 multi_queue=0, add_occupied_crust=0, clear_sphere_radius=1.5,
                 occupied_sphere_radius=5.0)
        d.update(kw)
        super().__init__(**d)


class IcpConfig(C.Structure):
    """vbo_icp_config; defaults = ICP::Config (alignment/icp.h:76-108) except num_threads, which is
    hardware_concurrency() there (its threads race; 1 is the deterministic setting)."""
    _fields_ = [("refine_roll_pitch", C.c_int32), ("mini_batch_size", C.c_int32), ("min_match_ratio", C.c_float),
                ("subsample_keep_ratio", C.c_float), ("inital_translation_weighting", C.c_float),
                ("inital_rotation_weighting", C.c_float), ("num_threads", C.c_int32)]

    def __init__(self, **kw):
        d = dict(refine_roll_pitch=0, mini_batch_size=20, min_match_ratio=0.8, subsample_keep_ratio=0.5,
                 inital_translation_weighting=100.0, inital_rotation_weighting=100.0, num_threads=1)
        d.update(kw)
        super().__init__(**d)


def available(which: str) -> bool:
    return os.path.exists(_PATHS[which])


class OracleLib:
    def __init__(self, which: str = "port"):
        path = _PATHS[which]
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle` (or __graft_entry__.build())")
        self.which = which
        lib = C.CDLL(path)
        lib.vbo_impl_name.restype = C.c_char_p
        lib.vbo_create.restype = C.c_void_p
        lib.vbo_create.argtypes = [C.POINTER(TsdfConfig), C.c_float, C.c_int]
        lib.vbo_destroy.argtypes = [C.c_void_p]
        lib.vbo_integrate.restype = C.c_int
        lib.vbo_integrate.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_uint64, C.c_int, C.c_int]
        lib.vbo_last_seconds.restype = C.c_double
        lib.vbo_last_seconds.argtypes = [C.c_void_p]
        lib.vbo_last_counters.argtypes = [C.c_void_p, C.c_void_p]
        lib.vbo_num_blocks.restype = C.c_uint64
        lib.vbo_num_blocks.argtypes = [C.c_void_p, C.c_int]
        lib.vbo_block_indices.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.vbo_get_block.restype = C.c_int
        lib.vbo_get_block.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.vbo_esdf_create.restype = C.c_int
        lib.vbo_esdf_create.argtypes = [C.c_void_p, C.POINTER(EsdfConfig)]
        lib.vbo_esdf_update.restype = C.c_int
        lib.vbo_esdf_update.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.vbo_serialize_block.restype = C.c_int
        lib.vbo_serialize_block.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        lib.vbo_deserialize_block.restype = C.c_int
        lib.vbo_deserialize_block.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        lib.vbo_esdf_update_blocks.restype = C.c_int
        lib.vbo_esdf_update_blocks.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]
        lib.vbo_esdf_set_max_distance.restype = C.c_int
        lib.vbo_esdf_set_max_distance.argtypes = [C.c_void_p, C.c_float]
        lib.vbo_esdf_set_full_euclidean.restype = C.c_int
        lib.vbo_esdf_set_full_euclidean.argtypes = [C.c_void_p, C.c_int]
        lib.vbo_esdf_add_robot_position.restype = C.c_int
        lib.vbo_esdf_add_robot_position.argtypes = [C.c_void_p, C.c_void_p]
        lib.vbo_mesh_generate.restype = C.c_int
        lib.vbo_mesh_generate.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_int]
        lib.vbo_mesh_num_blocks.restype = C.c_uint64
        lib.vbo_mesh_num_blocks.argtypes = [C.c_void_p]
        lib.vbo_mesh_block_indices.argtypes = [C.c_void_p, C.c_void_p]
        lib.vbo_mesh_get.restype = C.c_uint64
        lib.vbo_mesh_get.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.vbo_umap_order.restype = None
        lib.vbo_umap_order.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        lib.vbo_icp_run.restype = C.c_int
        lib.vbo_icp_run.argtypes = [C.c_void_p, C.POINTER(IcpConfig), C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                                    C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
        lib.vbo_icp_shuffle.restype = None
        lib.vbo_icp_shuffle.argtypes = [C.c_uint64, C.c_uint32, C.c_void_p]
        lib.vbo_mc_tables.restype = C.c_int
        lib.vbo_mc_tables.argtypes = [C.c_void_p, C.c_void_p]
        self.lib = lib
        assert lib.vbo_impl_name().decode() == which.split("_")[0]


def umap_order(lib: "OracleLib", hashes: np.ndarray) -> np.ndarray:
    """Keys 0..n-1 (hash(key) = hashes[key]) in the iteration order of a std::unordered_map filled by operator[]."""
    h = np.ascontiguousarray(hashes, dtype=np.uint32)
    out = np.zeros(h.size, dtype=np.uint32)
    lib.lib.vbo_umap_order(h.ctypes.data, h.size, out.ctypes.data)
    return out


class OracleMap:
    """One TSDF (+ optional ESDF) layer with the three integrators attached."""

    def __init__(self, lib: OracleLib, cfg: TsdfConfig, voxel_size: float, voxels_per_side: int = 16):
        self.lib = lib.lib
        self.vps = voxels_per_side
        self.voxel_size = voxel_size
        self.h = self.lib.vbo_create(C.byref(cfg), voxel_size, voxels_per_side)

    def close(self):
        if self.h:
            self.lib.vbo_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def integrate(self, kind: int, scan, freespace: bool = False, order: int = ORDER_REFERENCE):
        pts, cols, q, t = scan
        pts = np.ascontiguousarray(pts, dtype=np.float32)
        cols = np.ascontiguousarray(cols, dtype=np.uint8)
        q = np.ascontiguousarray(q, dtype=np.float32)
        t = np.ascontiguousarray(t, dtype=np.float32)
        assert pts.shape[0] == cols.shape[0]
        rc = self.lib.vbo_integrate(self.h, kind, q.ctypes.data, t.ctypes.data, pts.ctypes.data,
                                    cols.ctypes.data, pts.shape[0], int(freespace), order)
        if rc != 0:
            raise RuntimeError(f"vbo_integrate rc={rc}")

    def icp(self, cfg: "IcpConfig", points, q_wxyz, t, seed: int):
        """ICP::runICP against this map's TSDF layer -> (q_wxyz, t, num_updates)."""
        pts = np.ascontiguousarray(points, dtype=np.float32)
        q = np.ascontiguousarray(q_wxyz, dtype=np.float32)
        tt = np.ascontiguousarray(t, dtype=np.float32)
        oq, ot, nu = np.zeros(4, np.float32), np.zeros(3, np.float32), C.c_uint64(0)
        rc = self.lib.vbo_icp_run(self.h, C.byref(cfg), pts.ctypes.data, pts.shape[0], q.ctypes.data, tt.ctypes.data,
                                  int(seed) & 0xffffffff, oq.ctypes.data, ot.ctypes.data, C.byref(nu))
        if rc != 0:
            raise RuntimeError(f"vbo_icp_run rc={rc}")
        return oq, ot, int(nu.value)

    def last_seconds(self) -> float:
        return float(self.lib.vbo_last_seconds(self.h))

    def counters(self) -> Dict[str, int]:
        out = np.zeros(8, dtype=np.uint64)
        self.lib.vbo_last_counters(self.h, out.ctypes.data)
        names = ["rays", "clear_rays", "updates", "voxels_touched", "blocks_touched",
                 "blocks_allocated", "valid_points", "reserved"]
        return {k: int(v) for k, v in zip(names, out)}

    def block_indices(self, layer: int = LAYER_TSDF) -> np.ndarray:
        m = int(self.lib.vbo_num_blocks(self.h, layer))
        out = np.zeros((m, 3), dtype=np.int32)
        if m:
            self.lib.vbo_block_indices(self.h, layer, out.ctypes.data)
        return out

    def block(self, idx, layer: int = LAYER_TSDF) -> Tuple[np.ndarray, int]:
        dt = TSDF_DTYPE if layer == LAYER_TSDF else ESDF_DTYPE
        vox = np.zeros(self.vps ** 3, dtype=dt)
        i3 = np.ascontiguousarray(idx, dtype=np.int32)
        upd = C.c_uint8(0)
        rc = self.lib.vbo_get_block(self.h, layer, i3.ctypes.data, vox.ctypes.data, C.byref(upd))
        if rc != 0:
            raise KeyError(tuple(int(v) for v in idx))
        return vox, int(upd.value)

    def blocks(self, layer: int = LAYER_TSDF) -> Dict[Tuple[int, int, int], np.ndarray]:
        return {tuple(int(v) for v in i): self.block(i, layer)[0] for i in self.block_indices(layer)}

    def esdf_create(self, cfg: EsdfConfig):
        rc = self.lib.vbo_esdf_create(self.h, C.byref(cfg))
        if rc != 0:
            raise RuntimeError(f"vbo_esdf_create rc={rc}")

    def esdf_update(self, batch: bool = False, clear_updated_flag: bool = True):
        rc = self.lib.vbo_esdf_update(self.h, int(batch), int(clear_updated_flag))
        if rc != 0:
            raise RuntimeError(f"vbo_esdf_update rc={rc}")

    def serialize_block(self, index, layer: int = 0) -> np.ndarray:
        """Block::serializeToIntegers (src/core/block.cc:159-183 / :203-234)"""
        idx = np.ascontiguousarray(index, dtype=np.int32)
        words = np.zeros(self.vps ** 3 * (3 if layer == 0 else 2), dtype=np.uint32)
        rc = self.lib.vbo_serialize_block(self.h, layer, idx.ctypes.data, words.ctypes.data)
        if rc != 0:
            raise KeyError(tuple(int(v) for v in idx))
        return words

    def deserialize_block(self, index, words, layer: int = 0):
        """Block::deserializeFromIntegers into a (new) block (block.cc:65-90 / :110-135)"""
        idx = np.ascontiguousarray(index, dtype=np.int32)
        w = np.ascontiguousarray(words, dtype=np.uint32)
        assert w.size == self.vps ** 3 * (3 if layer == 0 else 2)
        if self.lib.vbo_deserialize_block(self.h, layer, idx.ctypes.data, w.ctypes.data) != 0:
            raise RuntimeError("vbo_deserialize_block")

    def esdf_update_blocks(self, indices, incremental: bool = False):
        """EsdfIntegrator::updateFromTsdfBlocks (esdf_integrator.cc:124-302)"""
        idx = np.ascontiguousarray(indices, dtype=np.int32).reshape(-1, 3)
        rc = self.lib.vbo_esdf_update_blocks(self.h, idx.ctypes.data, idx.shape[0], int(incremental))
        if rc != 0:
            raise RuntimeError(f"vbo_esdf_update_blocks rc={rc}")

    def esdf_set_max_distance(self, d: float):
        if self.lib.vbo_esdf_set_max_distance(self.h, float(d)) != 0:
            raise RuntimeError("vbo_esdf_set_max_distance")

    def esdf_set_full_euclidean(self, on: bool):
        if self.lib.vbo_esdf_set_full_euclidean(self.h, int(on)) != 0:
            raise RuntimeError("vbo_esdf_set_full_euclidean")

    def esdf_add_robot_position(self, p):
        p = np.ascontiguousarray(p, dtype=np.float32)
        rc = self.lib.vbo_esdf_add_robot_position(self.h, p.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"vbo_esdf_add_robot_position rc={rc}")

    # ---- MeshIntegrator<TsdfVoxel> (mesh/mesh_integrator.h) into the map's MeshLayer
    def mesh_generate(self, use_color: bool = True, min_weight: float = 1e-4, only_mesh_updated_blocks: bool = False,
                      clear_updated_flag: bool = True):
        rc = self.lib.vbo_mesh_generate(self.h, int(use_color), float(min_weight), int(only_mesh_updated_blocks),
                                        int(clear_updated_flag))
        if rc != 0:
            raise RuntimeError(f"vbo_mesh_generate rc={rc}")

    def mesh_block_indices(self) -> np.ndarray:
        n = int(self.lib.vbo_mesh_num_blocks(self.h))
        out = np.zeros((n, 3), dtype=np.int32)
        if n:
            self.lib.vbo_mesh_block_indices(self.h, out.ctypes.data)
        return out

    def mesh_block(self, index):
        """(vertices [n,3] f32, normals [n,3] f32, colors [n,4] u8 or None, updated) of one block mesh"""
        idx = np.ascontiguousarray(index, dtype=np.int32)
        n = int(self.lib.vbo_mesh_get(self.h, idx.ctypes.data, None, None, None, None, None))
        if n >= 2 ** 63:
            raise KeyError(tuple(int(v) for v in idx))
        v = np.zeros((n, 3), dtype=np.float32)
        nr = np.zeros((n, 3), dtype=np.float32)
        col = np.zeros((n, 4), dtype=np.uint8)
        has, upd = C.c_int(0), C.c_int(0)
        self.lib.vbo_mesh_get(self.h, idx.ctypes.data, v.ctypes.data, nr.ctypes.data, col.ctypes.data,
                              C.byref(has), C.byref(upd))
        return v, nr, (col if has.value else None), bool(upd.value)
