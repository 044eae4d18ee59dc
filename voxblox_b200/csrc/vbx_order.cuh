// The Merged integrator's bundle order: the iteration order of the reference's
//   LongIndexHashMapType<AlignedVector<size_t>>::type voxel_map / clear_map
// (tsdf_integrator.cc:318-322), which integrateVoxels walks with one thread
// (tsdf_integrator.cc:436-456).  That map is a libstdc++ std::unordered_map filled by bundleRays
// (cc:340-371) in point order, so its iteration order is a pure function of
//   (a) the sequence of DISTINCT voxel keys in first-occurrence order,
//   (b) LongIndexHash of each key (core/block_hash.h:52-64), and
//   (c) libstdc++'s insertion / rehash rules (bits/hashtable.h _M_insert_bucket_begin,
//       _M_rehash_aux; bits/hashtable_policy.h _Prime_rehash_policy):
//       - a node whose bucket is non-empty goes to the FRONT of that bucket's group of nodes;
//       - a node whose bucket is empty goes to the FRONT of the whole list (a new group);
//       - a rehash re-inserts every node, in list order, into the new bucket array by the same
//         two rules;
//       - the bucket count grows along the policy's prime table when the element count would
//         exceed it (max_load_factor 1).
// Hence after any sequence of insertions into a table of n buckets the list is: groups in
// descending order of their creation time, nodes of a group in descending insertion time -- so
//   position(b) = #nodes in groups created later than b's group + #nodes of b's group inserted later
// which is data parallel: a chain per bucket (atomicExch), the group's creation time and size by
// walking the chain, a suffix sum over creation times.  A rehash is the same computation with
// "insertion time" = position in the list before the rehash.  The schedule of rehashes (element
// count at which it happens, new bucket count) depends on the number of insertions only and is
// computed on the host from the very policy class the C++ library ships (vbx_create), so the
// device follows the libstdc++ the engine is built with.
//
// Checked against a real std::unordered_map on the host by tests/umap_order_check.cc (same code,
// compiled for the host), and end to end by every Merged parity test (bit-exact against the
// reference's own MergedTsdfIntegrator with one thread).
#pragma once

#include <stdint.h>

namespace vbx {

struct RehashSchedule {  // event k: when the map holds m[k] elements the bucket count becomes n[k]
  int count;
  uint32_t m[30];
  uint32_t n[30];
};

struct OrderScratch {     // global tables (the clearing map's h / head_of / wp follow the normal map's)
  uint32_t* h;            // [2 * cap] LongIndexHash per element (insertion order)
  uint32_t* tau;          // [cap] current insertion time
  uint32_t* tau2;         // [cap]
  uint32_t* next;         // [cap] bucket chain
  uint32_t* bkt;          // [cap] bucket of the element in the current table
  uint32_t* A;            // [cap] group sizes by creation time -> suffix sums
  uint32_t* bhead;        // [bucket_cap]
  uint32_t* head_of;      // [2 * cap] element -> bundle id
  uint32_t* wp;           // [2 * (cap / 32 + 2)] popcount prefixes of the first-occurrence bitmaps
  uint32_t* cta_tot;      // [64] per-block totals of the grid-wide suffix sum
  uint32_t cap, bucket_cap;
};

constexpr uint32_t kOrderNil = 0xffffffffu;
constexpr int kOrderThreads = 1024;

#if defined(__CUDACC__)
// exclusive scan of one value per thread over the 1024-thread block; *total = block sum
__device__ __forceinline__ uint32_t order_block_scan(uint32_t v, uint32_t* warp_sums /* [33] shared */, uint32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = warp_sums[lane];
    uint32_t winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    warp_sums[lane] = winc - w;
    if (lane == 31) warp_sums[32] = winc;
  }
  __syncthreads();
  const uint32_t r = warp_sums[warp] + inc - v;
  *total = warp_sums[32];
  __syncthreads();  // warp_sums may be reused right away
  return r;
}

// List positions of elements [0, m) with insertion times tau[] in a table of n buckets.
// Result in tau_out[]; A, next, bkt, bhead are scratch.  `tag` (distinct per call, < 2^12) marks the
// bucket heads written by this call, so the bucket array needs no clearing between calls:
// bhead[j] = tag << 20 | element, anything with another tag reads as "empty" (elements < 2^20).
// IdxT: uint32_t, or uint16_t when m < 65535 and n <= 65535 (times, chain links, bucket numbers and the
// suffix sums all fit 16 bits then): the tables of a few thousand bundles take half the shared memory.
// h may live in shared or in global memory (it is read once per element and call).
template <typename IdxT>
__device__ inline void order_positions(const uint32_t* h, const IdxT* tau, IdxT* tau_out, IdxT* next, IdxT* bkt, IdxT* A,
                                       uint32_t* bhead, uint32_t m, uint32_t n, uint32_t tag, uint32_t* warp_sums) {
  constexpr uint32_t nil = (uint32_t)(IdxT)~(IdxT)0;
  const uint32_t tid = threadIdx.x;
  const uint32_t tg = tag << 20;
  // chains: next[b] = the element that headed b's bucket before b (nil: none)
  for (uint32_t b = tid; b < m; b += kOrderThreads) {
    const uint32_t g = h[b] % n;
    bkt[b] = (IdxT)g;
    const uint32_t old = atomicExch(&bhead[g], tg | b);
    next[b] = (old >> 20) == tag ? (IdxT)(old & 0xfffffu) : (IdxT)nil;
  }
  __syncthreads();
  // the group's creator (smallest time) publishes the group size at its creation time; every time in
  // [0, m) belongs to exactly one element, so A needs no clearing either
  for (uint32_t b = tid; b < m; b += kOrderThreads) {
    const uint32_t tb = tau[b];
    uint32_t cmin = 0xffffffffu, size = 0;
    for (uint32_t c = bhead[bkt[b]] & 0xfffffu; c != nil; c = next[c]) {
      cmin = min(cmin, (uint32_t)tau[c]);
      ++size;
    }
    A[tb] = (IdxT)(cmin == tb ? size : 0u);
  }
  __syncthreads();
  // A[t] <- number of elements in groups created after time t (exclusive suffix sum): every thread owns
  // a contiguous run of times, one block-wide scan over the run totals
  {
    const uint32_t per = (m + kOrderThreads - 1) / kOrderThreads;
    // thread 0 owns the HIGHEST times
    const uint32_t hi = m > tid * per ? m - tid * per : 0u;          // one past my highest time
    const uint32_t lo = hi > per ? hi - per : 0u;
    uint32_t sum = 0;
    for (uint32_t t = lo; t < hi; ++t) sum += A[t];
    uint32_t total;
    uint32_t run = order_block_scan(sum, warp_sums, &total);          // elements at times above my run
    for (uint32_t t = hi; t-- > lo;) {
      const uint32_t v = A[t];
      A[t] = (IdxT)run;
      run += v;
    }
  }
  __syncthreads();
  for (uint32_t b = tid; b < m; b += kOrderThreads) {
    const uint32_t tb = tau[b];
    uint32_t cmin = 0xffffffffu, later = 0;
    for (uint32_t c = bhead[bkt[b]] & 0xfffffu; c != nil; c = next[c]) {
      const uint32_t tc = tau[c];
      cmin = min(cmin, tc);
      later += tc > tb ? 1u : 0u;
    }
    tau_out[b] = (IdxT)((uint32_t)A[cmin] + later);
  }
  __syncthreads();
}

// All rehash stages + the final listing for B elements with hashes h[] (insertion order).
// On return pos[] (= one of tau / tau2, returned) holds every element's position in the iteration order.
// bhead must not hold values with tags 1.. from an earlier run: it is cleared here (n_final entries).
template <typename IdxT>
__device__ inline IdxT* order_run(const RehashSchedule& rs, uint32_t B, const uint32_t* h, IdxT* tau, IdxT* tau2, IdxT* next,
                                  IdxT* bkt, IdxT* A, uint32_t* bhead, uint32_t n_final, uint32_t* warp_sums) {
  for (uint32_t e = threadIdx.x; e < B; e += kOrderThreads) tau[e] = (IdxT)e;
  for (uint32_t j = threadIdx.x; j < n_final; j += kOrderThreads) bhead[j] = 0u;  // tag 0 = never used
  __syncthreads();
  uint32_t n_cur = 1;
  uint32_t tag = 1;
  IdxT* cur = tau;
  IdxT* oth = tau2;
  for (int k = 0; k < rs.count; ++k) {
    const uint32_t mk = rs.m[k];
    if (mk >= B) break;  // the map never reaches this size
    if (mk > 0) {
      order_positions<IdxT>(h, cur, oth, next, bkt, A, bhead, mk, n_cur, tag++, warp_sums);
      // elements inserted after the rehash keep their insertion index as time
      for (uint32_t e = mk + threadIdx.x; e < B; e += kOrderThreads) oth[e] = (IdxT)e;
      __syncthreads();
      IdxT* t = cur;
      cur = oth;
      oth = t;
    }
    n_cur = rs.n[k];
  }
  order_positions<IdxT>(h, cur, oth, next, bkt, A, bhead, B, n_cur, tag, warp_sums);
  return oth;
}

// Shared-memory words the single-block form needs for B elements in n_final buckets: five 16-bit index
// tables and the bucket heads (0xffffffff: too many elements for 16-bit tables -- such a map would not fit
// one SM's shared memory with 32-bit tables either).  Host and device use the same formula.
__host__ __device__ inline uint32_t order_smem_words_needed(uint32_t B, uint32_t n_final) {
  if (B >= 65535u || n_final > 65535u) return 0xffffffffu;
  const uint32_t pad = (B + 1u) & ~1u;  // (keeps every 16-bit table 4-byte aligned)
  return 5u * pad / 2u + n_final;
}
#endif  // __CUDACC__

}  // namespace vbx
