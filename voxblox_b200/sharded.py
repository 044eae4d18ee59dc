"""One map over the GPUs of a box: block-ownership sharding (BASELINE.json north_star, SURVEY.md
section 8e; DESIGN.md "multi-GPU").  One process per GPU, torch.distributed for the plumbing.

Every rank is handed EVERY scan and runs the ordinary integrators; the engine (created with
EngineOptions(rank=r, world_size=W)) applies only the voxel updates of the blocks rank r owns and
creates only those blocks.  Bundling, the reference's bundle order, the merge and the ray walk are
the same deterministic computation on every rank, so the ranks need no exchange while integrating:
the union of the W shards is the single-GPU map bit for bit (tests/test_sharded_gpu.py), every voxel
keeping the reference's one-thread update order.

Why no reduce: the reference's closest multi-map model is the layer merge of
voxblox/src/utils/voxel_utils.cc:9-22, which is associative only because it averages; the
integrator's clamp after every update (tsdf_integrator.cc:205-208) is not, so partial TSDF sums
from different ray ranges cannot be combined into the reference's result.  Ownership sharding keeps
every voxel's update chain on one GPU instead.

This module holds the host-side logic: the ownership function (mirrors vbx_block_owner), gathering
the shards into one block dictionary, and keeping read-only replicas of the other ranks' dirty blocks
(`sync_replicas`, one all-gather of the blocks dirtied since the last call).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

from .api import EngineOptions, Layer, VoxbloxError


# ------------------------------------------------------------------ host logic (CPU-testable)
def block_owner(index, world: int) -> np.ndarray:
    """vbx_block_owner: rank owning block (bx, by, bz) = (bx + 2 by + 4 bz) mod world (non-negative)."""
    idx = np.asarray(index, dtype=np.int64).reshape(-1, 3)
    return ((idx[:, 0] + 2 * idx[:, 1] + 4 * idx[:, 2]) % int(world)).astype(np.int32)


def shard_options(rank: int, world: int, **kw) -> EngineOptions:
    """EngineOptions for rank `rank` of a `world`-way sharded map."""
    if not 0 <= rank < world:
        raise VoxbloxError("rank outside [0, world_size)")
    return EngineOptions(rank=rank, world_size=world, **kw)


def pack_blocks(indices: np.ndarray, voxels: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """(M, 3) int32 indices and (M, vps^3) voxel records -> two contiguous byte arrays."""
    idx = np.ascontiguousarray(indices, dtype=np.int32).reshape(-1, 3)
    vox = np.ascontiguousarray(voxels)
    return idx.view(np.uint8).reshape(-1), vox.view(np.uint8).reshape(-1)


def all_gather_blocks(indices: np.ndarray, voxels: np.ndarray, group=None):
    """All ranks' (indices, voxels), rank after rank.  Variable block counts per rank: the counts
    travel first, then one padded all-gather (gloo on CPU tensors, NCCL on CUDA tensors)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    nccl = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if nccl else torch.device("cpu")
    idx_b, vox_b = pack_blocks(indices, voxels)
    m = int(np.asarray(indices).reshape(-1, 3).shape[0])
    per_block = int(vox_b.size // m) if m else int(voxels.dtype.itemsize * (voxels.shape[-1] if voxels.ndim > 1 else 0))
    meta = torch.tensor([m, per_block], dtype=torch.int64, device=dev)
    metas = torch.zeros(world * 2, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(metas, meta, group=group)
    metas = metas.cpu().numpy().reshape(world, 2)
    counts = metas[:, 0]
    per_block = int(metas[:, 1].max())
    cap = int(counts.max())
    out_idx, out_vox = [], []
    if cap == 0:
        return np.zeros((0, 3), np.int32), np.zeros((0,), np.uint8), counts
    send = torch.zeros(cap * (12 + per_block), dtype=torch.uint8, device=dev)
    if m:
        send[: m * 12] = torch.from_numpy(idx_b.copy()).to(dev)
        send[cap * 12: cap * 12 + m * per_block] = torch.from_numpy(vox_b.copy()).to(dev)
    recv = torch.zeros(world * send.numel(), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.cpu().numpy().reshape(world, -1)
    for r in range(world):
        c = int(counts[r])
        out_idx.append(recv[r, : c * 12].view(np.int32).reshape(c, 3))
        out_vox.append(recv[r, cap * 12: cap * 12 + c * per_block].reshape(c, per_block))
    return np.concatenate(out_idx), np.concatenate(out_vox), counts


class ShardedLayer:
    """The block-sharded map as its callers see it.

    `layer` is this rank's shard (a Layer whose engine was created with shard_options)."""

    def __init__(self, layer: Layer, group=None):
        import torch.distributed as dist

        self.layer = layer
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def owned(self, indices) -> np.ndarray:
        return block_owner(indices, self.world) == self.rank

    def gather(self) -> Dict[Tuple[int, int, int], np.ndarray]:
        """Every block of the map (all shards), as {index: voxels} on every rank."""
        idx = self.layer.getAllAllocatedBlocks()
        vox, _ = self.layer.getBlocks(idx)
        if len(idx):
            mine = self.owned(idx)
            idx, vox = idx[mine], vox[mine]
        all_idx, all_vox, _ = all_gather_blocks(idx, vox, self.group)
        dt = vox.dtype
        return {tuple(int(v) for v in i): all_vox[k].view(dt) for k, i in enumerate(all_idx)}

    def sync_replicas(self, updated_mask: int = 1) -> int:
        """Make every rank's map a full replica for READING: all-gather the owned blocks dirtied since
        the last call (Update::kMap bit by default, cleared on the owner) and upload the other ranks'
        blocks as read-only copies (the integrators never touch blocks this rank does not own).
        Returns the number of blocks received."""
        idx, vox, _ = self.layer.mirrorUpdated(updated_mask, clear_mask=updated_mask)
        if len(idx):
            mine = self.owned(idx)
            idx, vox = idx[mine], vox[mine]
        all_idx, all_vox, counts = all_gather_blocks(idx, vox, self.group)
        if not len(all_idx):
            return 0
        theirs = ~self.owned(all_idx)
        if theirs.any():
            self.layer.insertBlocks(all_idx[theirs], all_vox[theirs].view(vox.dtype).reshape(int(theirs.sum()), -1),
                                    updated_bits=np.zeros(int(theirs.sum()), np.uint8))
        return int(theirs.sum())
