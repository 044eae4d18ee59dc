// Device-side access to the GPU-resident block hash (the device mirror of
// Layer<T>::block_map_, voxblox/include/voxblox/core/layer.h:30-32,292).
#pragma once

#include "vbx_engine.h"

namespace vbx {

// ------------------------------------------------------------------ block hash
// Find the hash position of a block, creating the entry if it is missing
// (allocateStorageAndGetVoxelPtr's find-or-emplace, cc:109-124, without the mutex:
// one CAS decides the winner).  Pool slots are assigned later by k_assign.
__device__ inline uint32_t ensure_block(const Tables& t, uint64_t key, ScanState* st) {
  uint32_t hp = hash64(key) & t.hmask;
  for (uint32_t probe = 0; probe <= t.hmask; ++probe) {
    const uint64_t k = *reinterpret_cast<volatile uint64_t*>(t.hkeys + hp);
    if (k == key) return hp;
    if (k == kEmptyKey) {
      const unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(t.hkeys + hp),
                                               (unsigned long long)kEmptyKey, (unsigned long long)key);
      if (old == kEmptyKey) {
        const uint32_t j = atomicAdd(&st->n_new, 1u);
        if (j < t.max_blocks) {
          t.new_list[j] = hp;
        } else {
          atomicOr(&st->error, kErrPoolFull);
        }
        return hp;
      }
      if (old == key) return hp;
    }
    hp = (hp + 1) & t.hmask;
  }
  atomicOr(&st->error, kErrHashFull);
  return 0xffffffffu;
}

__device__ __forceinline__ uint32_t find_block(const Tables& t, uint64_t key) {
  uint32_t hp = hash64(key) & t.hmask;
  for (uint32_t probe = 0; probe <= t.hmask; ++probe) {
    const uint64_t k = t.hkeys[hp];
    if (k == key) return hp;
    if (k == kEmptyKey) return 0xffffffffu;
    hp = (hp + 1) & t.hmask;
  }
  return 0xffffffffu;
}

// Blocks touched by the current call get dense ids 0, 1, 2, ... (update records are keyed by
// (touched id, voxel in block): a handful of bits instead of a hash position).  The per-position
// word packs (call id, touched id); the first toucher of a block in this call installs it with one
// CAS.  A thread that loses the CAS race has drawn an id nobody uses: it is marked as a hole in
// touched_list (0xffffffff) -- ids stay dense enough, n_touched counts the blocks exactly.
__device__ __forceinline__ uint32_t touch_block(const Tables& t, uint32_t hp, uint32_t epoch, ScanState* st) {
  unsigned long long* w = t.htouch + hp;
  const unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(w);
  if ((uint32_t)(cur >> 32) == epoch) return (uint32_t)cur;
  const uint32_t id = atomicAdd(&st->n_touch_ids, 1u);
  if (id >= t.touched_cap) {
    atomicOr(&st->error, kErrPoolFull);
    return 0u;
  }
  const unsigned long long want = ((unsigned long long)epoch << 32) | id;
  const unsigned long long old = atomicCAS(w, cur, want);
  if (old == cur) {
    t.touched_list[id] = hp;
    atomicAdd(&st->n_touched, 1u);
    return id;
  }
  t.touched_list[id] = 0xffffffffu;  // a hole
  return (uint32_t)old;              // (only this call's walk writes these words: the winner carries this call's id)
}

}  // namespace vbx
