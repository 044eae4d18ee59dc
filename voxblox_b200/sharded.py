"""Ray-range sharding of one scan over the GPUs of a box (BASELINE.json north_star, SURVEY.md
section 8e): one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch) for the one
exchange step, the C-ABI's vbx_shard_front / vbx_shard_back for the device work.

Every rank keeps a full replica of the map and is handed the full cloud.  All ranks bundle the
whole cloud (so they agree on the bundles), then rank r merges and ray-casts only the ray slots
[r*S, (r+1)*S), S = ceil(n / world).  The ranks all-gather their update records (16 B each:
global voxel key + ray slot) and their slices of the per-ray tables; every rank then applies ALL
records in ray order to its replica.  The per-voxel update order is the single-GPU order, so the
replicas and the single-GPU map are bit-identical (tests/test_sharded_gpu.py).

The reference has no multi-device path; its closest semantic model is the layer merge of
voxblox/src/utils/voxel_utils.cc:9-22, which is associative only because it averages -- the
integrator's clamp-after-every-update (tsdf_integrator.cc:205-208) is not, which is why the
exchange carries update records rather than partial sums.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np

from .api import TsdfIntegratorBase, VoxbloxError, _as_pose


class ShardLayout(C.Structure):
    """vbx_shard_layout"""
    _fields_ = [("record_capacity", C.c_uint64), ("slice", C.c_uint64), ("pack_bytes", C.c_uint64),
                ("off_ray_a", C.c_uint64), ("off_ray_c", C.c_uint64), ("off_records", C.c_uint64)]


RECORD_BYTES = 16
RECORD_GRANULE = 4096  # the exchanged prefix grows in steps of 4096 records


# ------------------------------------------------------------------ host logic (CPU-testable)
def slot_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous ray-slot range of a rank: [rank*S, (rank+1)*S) clipped to n, S = ceil(n/world)."""
    s = (n + world - 1) // world
    return min(rank * s, n), min((rank + 1) * s, n)


def exchange_bytes(off_records: int, counts: Sequence[int]) -> int:
    """Bytes of each rank's pack that must travel: the per-ray tables plus the longest record
    list, rounded up to the record granule (all ranks must send the same length)."""
    m = max(int(c) for c in counts) if len(counts) else 0
    m = (m + RECORD_GRANULE - 1) // RECORD_GRANULE * RECORD_GRANULE
    return int(off_records) + RECORD_BYTES * m


def record_prefix(counts: Sequence[int]) -> List[int]:
    """Start of every rank's records in the global (ray-ordered) record sequence."""
    out = [0]
    for c in counts:
        out.append(out[-1] + int(c))
    return out


def gather_counts(count: int, group=None):
    """All ranks' record counts (one small all-gather; gloo on CPU, NCCL on GPU)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    mine = torch.tensor([int(count)], dtype=torch.int64, device=dev)
    out = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(out, mine, group=group)
    return [int(v) for v in out.cpu().tolist()]


def gather_packs(pack, nbytes: int, gathered, group=None):
    """All-gather the first nbytes of every rank's pack into gathered[:world * nbytes]."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    dist.all_gather_into_tensor(gathered[: world * nbytes], pack[:nbytes], group=group)
    return gathered[: world * nbytes]


# ------------------------------------------------------------------------------ device path
class ShardedTsdfIntegrator:
    """integratePointCloud for one scan shared by all ranks of a torch.distributed group.

    Wraps a Simple or Merged integrator created on a Layer whose EngineOptions carry this
    process' rank and world_size."""

    def __init__(self, integrator: TsdfIntegratorBase, record_capacity: int = 1 << 21, group=None):
        import torch
        import torch.distributed as dist

        if integrator.kind not in (1, 2):
            raise VoxbloxError("ray-range sharding supports the simple and merged integrators")
        self.integ = integrator
        self.ctx = integrator._ctx
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.record_capacity = int(record_capacity)
        self.lib = self.ctx.lib
        self.lib.vbx_shard_layout_for.restype = C.c_int
        self.lib.vbx_shard_layout_for.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(ShardLayout)]
        self.lib.vbx_shard_front.restype = C.c_int
        self.lib.vbx_shard_front.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_uint64, C.c_int, C.POINTER(ShardLayout), C.c_void_p,
                                             C.POINTER(C.c_uint64)]
        self.lib.vbx_shard_back.restype = C.c_int
        self.lib.vbx_shard_back.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64,
                                            C.POINTER(ShardLayout), C.c_void_p, C.c_uint64, C.c_void_p]
        self._torch = torch
        self._pack = None
        self._gathered = None
        self.last_exchange_bytes = 0

    def _buffers(self, lay: ShardLayout):
        torch = self._torch
        if self._pack is None or self._pack.numel() < lay.pack_bytes:
            dev = torch.device("cuda", torch.cuda.current_device())
            self._pack = torch.empty(int(lay.pack_bytes), dtype=torch.uint8, device=dev)
            self._gathered = torch.empty(int(lay.pack_bytes) * self.world, dtype=torch.uint8, device=dev)
        return self._pack, self._gathered

    def integratePointCloudDevice(self, T_G_C, d_xyz: int, d_rgba: int, n: int, freespace_points: bool = False):
        q, t = _as_pose(T_G_C)
        lay = ShardLayout()
        self.ctx.check(self.lib.vbx_shard_layout_for(self.ctx.handle, n, self.record_capacity, C.byref(lay)),
                       "vbx_shard_layout_for")
        pack, gathered = self._buffers(lay)
        count = C.c_uint64(0)
        self.ctx.check(self.lib.vbx_shard_front(self.ctx.handle, self.integ.kind, q.ctypes.data, t.ctypes.data,
                                                d_xyz, d_rgba, n, int(bool(freespace_points)), C.byref(lay),
                                                pack.data_ptr(), C.byref(count)), "vbx_shard_front")
        # vbx_shard_front returns after its stream has drained, so the pack is complete
        counts = gather_counts(int(count.value), self.group)
        nbytes = exchange_bytes(int(lay.off_records), counts)
        got = gather_packs(pack, nbytes, gathered, self.group)
        self._torch.cuda.current_stream().synchronize()
        self.last_exchange_bytes = nbytes * self.world
        carr = np.asarray(counts, dtype=np.uint64)
        self.ctx.check(self.lib.vbx_shard_back(self.ctx.handle, self.integ.kind, q.ctypes.data, t.ctypes.data, n,
                                               C.byref(lay), got.data_ptr(), nbytes, carr.ctypes.data),
                       "vbx_shard_back")
        return counts
