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

__device__ __forceinline__ void mark_touched(const Tables& t, uint32_t hp, uint32_t epoch, ScanState* st) {
  if (*reinterpret_cast<volatile uint32_t*>(t.htouch_epoch + hp) != epoch) {
    const uint32_t old = atomicExch(t.htouch_epoch + hp, epoch);
    if (old != epoch) {
      const uint32_t j = atomicAdd(&st->n_touched, 1u);
      if (j < t.max_blocks) t.touched_list[j] = hp;
    }
  }
}

}  // namespace vbx
