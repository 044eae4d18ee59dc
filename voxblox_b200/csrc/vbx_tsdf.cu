// TSDF integration on the device: the Simple / Merged / Fast integrators of
// voxblox/src/integrator/tsdf_integrator.cc re-designed as a data-parallel pipeline.
//
//   k_point_keys     transform + validate every point, key it by its end voxel
//                    (bundleRays, cc:340-371)
//   sort             stable radix sort of the point keys: bundles become runs, in the
//                    reference's point order inside a run
//   k_rays_count     one thread per ray / bundle: sequential weighted merge of the
//                    bundle (integrateVoxel cc:387-405), DDA walk (RayCaster) that
//                    creates missing blocks in the device hash and counts the voxels
//                    it will update (allocateStorageAndGetVoxelPtr cc:91-134)
//   scan + k_assign  offsets of every ray's update records; pool slots for new blocks
//   k_rays_emit      second DDA walk writing (voxel id, ray id) records
//   sort             stable radix sort by voxel id: every voxel's updates become one
//                    run, ordered by ray rank
//   k_apply          one pass over the runs applying updateTsdfVoxel (cc:150-209)
//                    sequentially per voxel -- clamp-after-every-update semantics
//                    preserved exactly, with no locks and no atomics on voxels
//
// Update order.  The reference applies a voxel's updates in whatever order its
// threads reach the voxel's mutex (cc:186); with one thread that is point order for
// Simple / Fast and unordered_map iteration order for Merged.  The device applies
// them in ray-rank order: point order (integration_order_mode) for Simple / Fast,
// and ascending (z, y, x) of the bundle voxel for Merged, normal bundles before
// clearing bundles (cc:323-335).  See DESIGN.md "update order".
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdio>
#include <cstring>

#include "vbx_engine.h"

namespace vbx {

struct ScanParams {
  Pose T;
  F3 origin;
  float voxel_size, voxel_size_inv;
  float trunc, min_ray, max_ray;
  UpdateParams up;
  int L;  // log2(voxels per side)
  int kind;
  int freespace, use_const_weight, allow_clear, carving, anti_grazing;
  int order_mode;      // 0 mixed, 1 sorted (order array)
  uint32_t n;          // points in the cloud
  uint32_t n_groups;   // n / 1024 (MixedThreadSafeIndex)
  float start_inv;     // start_voxel_subsampling_factor * voxel_size_inv
  int max_collisions;
  uint32_t set_epoch;  // generation tag of the Fast integrator's approximate sets
  uint32_t epoch;      // call id
  uint64_t max_updates;
};

// MixedThreadSafeIndex::getNextIndexImpl, integrator_utils.cc:54-63
__device__ __forceinline__ uint32_t point_order(const ScanParams& P, const uint32_t* order, uint32_t s) {
  if (P.order_mode == 1) return order[s];
  if (P.n_groups * 1024u <= s) return s;
  return (s % P.n_groups) * 1024u + s / P.n_groups;
}

__device__ __forceinline__ F3 load_point(const float* xyz, uint32_t idx) {
  return f3(__ldg(xyz + 3 * idx), __ldg(xyz + 3 * idx + 1), __ldg(xyz + 3 * idx + 2));
}
__device__ __forceinline__ uint32_t load_color(const uint8_t* rgba, uint32_t idx) {
  return __ldg(reinterpret_cast<const uint32_t*>(rgba) + idx);
}

// ------------------------------------------------------------------ block hash
// Find the hash position of a block, creating the entry if it is missing
// (allocateStorageAndGetVoxelPtr's find-or-emplace, cc:109-124, without the mutex:
// one CAS decides the winner).  Pool slots are assigned later by k_assign.
__device__ uint32_t ensure_block(const Tables& t, uint64_t key, ScanState* st) {
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

// ------------------------------------------------------------------- kernels
// Merged: key every point by its end voxel (bundleRays, cc:340-371).
__global__ void k_point_keys(ScanParams P, const float* __restrict__ xyz, const uint32_t* __restrict__ order,
                             uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, ScanState* st) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  bool valid = false;
  if (s < P.n) {
    const uint32_t idx = point_order(P, order, s);
    const F3 p = load_point(xyz, idx);
    const int cls = classify_point(p, P.min_ray, P.max_ray, P.allow_clear != 0, P.freespace != 0);
    uint64_t key = kInvalidPointKey;
    if (cls != 0) {
      const I3 v = grid_index(transform(P.T, p), P.voxel_size_inv);
      const int lim = kCoordBias - 1;
      if (v.x < -lim || v.x > lim || v.y < -lim || v.y > lim || v.z < -lim || v.z > lim) {
        atomicOr(&st->error, kErrCoordRange);
      } else {
        key = pack3(v.x, v.y, v.z) | ((uint64_t)(cls == 2) << 63);
        valid = true;
      }
    }
    keys[s] = key;
    vals[s] = idx;
  }
  const unsigned b = __ballot_sync(0xffffffffu, valid);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(&st->n_valid_points, (uint32_t)__popc(b));
}

// "sorted" integration order: key = |p|^2 (float, widened to double like
// SortedThreadSafeIndex, integrator_utils.cc:24-37); non-negative doubles order as integers.
__global__ void k_sqnorm_keys(uint32_t n, const float* __restrict__ xyz, uint64_t* __restrict__ keys,
                              uint32_t* __restrict__ vals) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const F3 p = load_point(xyz, i);
  const double d = (double)dot3(p, p);
  keys[i] = (uint64_t)__double_as_longlong(d);
  vals[i] = i;
}

// binary search over the sorted point keys: is there a NORMAL bundle ending in this voxel?
// (the voxel_map.find() of the anti-grazing test, cc:415-422)
__device__ bool bundle_exists(const uint64_t* keys, uint32_t n, uint64_t key) {
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (keys[mid] < key) {
      lo = mid + 1;
    } else {
      hi = mid;
    }
  }
  return lo < n && keys[lo] == key;
}

struct RayJob {
  F3 point_G;
  float weight;
  uint32_t color;
  bool clearing;
  uint64_t bundle_key;  // Merged: key of the bundle's own voxel (anti-grazing)
};

// Build ray i of this call, or return false when slot i casts nothing.
__device__ bool make_ray(const ScanParams& P, uint32_t i, const float* xyz, const uint8_t* rgba,
                         const uint32_t* order, const uint64_t* keys, const uint32_t* vals, RayJob* job) {
  if (P.kind == VBX_MERGED) {
    const uint64_t key = keys[i];
    if (key == kInvalidPointKey) return false;
    if (i > 0 && keys[i - 1] == key) return false;  // not the head of its bundle
    const bool clearing = (key >> 63) != 0;
    // integrateVoxel, cc:384-405: running weighted mean in the CAMERA frame, in list order
    F3 mp = f3(0.f, 0.f, 0.f);
    float mw = 0.0f;
    uint32_t mcol = 0u;
    for (uint32_t j = i; j < P.n && keys[j] == key; ++j) {
      const uint32_t idx = vals[j];
      const F3 p = load_point(xyz, idx);
      const float w = point_weight(p.z, P.use_const_weight != 0);
      if (w < VBX_EPS) continue;
      mp = div3(add3(scale3(mp, mw), scale3(p, w)), fadd(mw, w));
      mcol = blend_rgba(mcol, mw, load_color(rgba, idx), w);
      mw = fadd(mw, w);
      if (clearing) break;  // "only take first point when clearing"
    }
    job->point_G = transform(P.T, mp);
    job->weight = mw;
    job->color = mcol;
    job->clearing = clearing;
    job->bundle_key = key;
    return true;
  }
  // Simple / Fast: one ray per valid point (integrateFunction, cc:269-305 / :488-553)
  const uint32_t idx = point_order(P, order, i);
  const F3 p = load_point(xyz, idx);
  const int cls = classify_point(p, P.min_ray, P.max_ray, P.allow_clear != 0, P.freespace != 0);
  if (cls == 0) return false;
  job->point_G = transform(P.T, p);
  job->weight = point_weight(p.z, P.use_const_weight != 0);
  job->color = load_color(rgba, idx);
  job->clearing = (cls == 2);
  job->bundle_key = 0;
  return true;
}

// LongIndexHash, core/block_hash.h:52-64 (32-bit wrap of x + 17191 y + 17191^2 z)
__device__ __forceinline__ uint32_t long_index_hash(int x, int y, int z) {
  return (uint32_t)x + (uint32_t)y * 17191u + (uint32_t)z * 295530481u;
}
// ApproxHashSet::replaceHash, utils/approx_hash_array.h:125-134.  The generation tag in
// the upper word plays the role of the reference's sliding offset (h:155-168).
__device__ __forceinline__ bool replace_hash(unsigned long long* set, uint32_t h, uint32_t epoch) {
  const unsigned long long tag = ((unsigned long long)epoch << 32) | h;
  const unsigned long long old = atomicExch(set + (h & 0xfffffu), tag);
  return old != tag;
}

__global__ void k_rays_count(ScanParams P, Tables tab, const float* __restrict__ xyz,
                             const uint8_t* __restrict__ rgba, const uint32_t* __restrict__ order,
                             const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                             float4* __restrict__ ray_p, uint2* __restrict__ ray_c,
                             uint32_t* __restrict__ cnt, unsigned long long* set_start,
                             unsigned long long* set_observed, ScanState* st) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > P.n) return;
  if (i == P.n) {
    cnt[i] = 0;
    return;
  }
  RayJob job;
  if (!make_ray(P, i, xyz, rgba, order, keys, vals, &job)) {
    cnt[i] = 0;
    return;
  }
  if (P.kind == VBX_FAST) {
    // start-voxel subsampling, cc:507-519
    const I3 g = grid_index(job.point_G, P.start_inv);
    if (!replace_hash(set_start, long_index_hash(g.x, g.y, g.z), P.set_epoch)) {
      cnt[i] = 0;
      return;
    }
  }
  ray_p[i] = make_float4(job.point_G.x, job.point_G.y, job.point_G.z, job.weight);
  ray_c[i] = make_uint2(job.color, job.clearing ? 1u : 0u);
  atomicAdd(job.clearing ? &st->n_clear_rays : &st->n_rays, 1u);

  Dda d;
  dda_setup(d, P.origin, job.point_G, job.clearing, P.carving != 0, P.max_ray, P.voxel_size_inv, P.trunc,
            P.kind != VBX_FAST);
  uint32_t count = 0;
  int collisions = 0;
  int lbx = INT_MIN, lby = INT_MIN, lbz = INT_MIN;
  const int lim = (kCoordBias - 1) << P.L;
  for (unsigned int s = 0; s <= d.len; ++s, dda_advance(d)) {
    if (P.kind == VBX_MERGED && P.anti_grazing) {
      const uint64_t vkey = pack3(d.cx, d.cy, d.cz);
      if ((job.clearing || vkey != (job.bundle_key & ~(1ull << 63))) && bundle_exists(keys, P.n, vkey)) continue;
    }
    if (P.kind == VBX_FAST) {
      // cc:531-543: stop once the ray runs through voxels other rays already observed
      if (!replace_hash(set_observed, long_index_hash(d.cx, d.cy, d.cz), P.set_epoch)) {
        ++collisions;
      } else {
        collisions = 0;
      }
      if (collisions > P.max_collisions) break;
    }
    if (d.cx < -lim || d.cx > lim || d.cy < -lim || d.cy > lim || d.cz < -lim || d.cz > lim) {
      atomicOr(&st->error, kErrCoordRange);
      break;
    }
    const int bx = d.cx >> P.L, by = d.cy >> P.L, bz = d.cz >> P.L;
    if (bx != lbx || by != lby || bz != lbz) {
      const uint32_t hp = ensure_block(tab, pack3(bx, by, bz), st);
      if (hp == 0xffffffffu) break;
      mark_touched(tab, hp, P.epoch, st);
      lbx = bx;
      lby = by;
      lbz = bz;
    }
    ++count;
  }
  cnt[i] = count;
}

// Pool slots for the blocks created by this call, dense ranks for the touched ones.
__global__ void k_assign(Tables tab, const uint32_t* __restrict__ off, uint32_t n, uint32_t n_blocks_before,
                         uint64_t max_updates, ScanState* st) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n_new = min(st->n_new, tab.max_blocks);
  const uint32_t n_touched = min(st->n_touched, tab.max_blocks);
  if (j < n_new) {
    const uint32_t slot = n_blocks_before + j;
    if (slot < tab.max_blocks) {
      const uint32_t hp = tab.new_list[j];
      tab.hslot[hp] = (int32_t)slot;
      tab.slot_key[slot] = tab.hkeys[hp];
    } else {
      atomicOr(&st->error, kErrPoolFull);
    }
  }
  if (j < n_touched) tab.htouch_rank[tab.touched_list[j]] = j;
  if (j == 0) {
    st->total_updates = off[n];
    st->n_blocks = min(n_blocks_before + st->n_new, tab.max_blocks);
    if ((uint64_t)off[n] > max_updates) atomicOr(&st->error, kErrUpdatesFull);
  }
}

__global__ void k_rays_emit(ScanParams P, Tables tab, const uint64_t* __restrict__ keys,
                            const float4* __restrict__ ray_p, const uint2* __restrict__ ray_c,
                            const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ off,
                            uint32_t* __restrict__ ckeys, uint32_t* __restrict__ cvals) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n) return;
  const uint32_t c = cnt[i];
  if (c == 0) return;
  const float4 rp = ray_p[i];
  const bool clearing = (ray_c[i].y & 1u) != 0;
  const F3 point_G = f3(rp.x, rp.y, rp.z);
  Dda d;
  dda_setup(d, P.origin, point_G, clearing, P.carving != 0, P.max_ray, P.voxel_size_inv, P.trunc,
            P.kind != VBX_FAST);
  const uint64_t own = (P.kind == VBX_MERGED) ? (keys[i] & ~(1ull << 63)) : 0ull;
  uint32_t emitted = 0;
  int lbx = INT_MIN, lby = INT_MIN, lbz = INT_MIN;
  uint32_t rank = 0;
  const uint32_t base = off[i];
  const int mask = (1 << P.L) - 1;
  for (unsigned int s = 0; s <= d.len && emitted < c; ++s, dda_advance(d)) {
    if (P.kind == VBX_MERGED && P.anti_grazing) {
      const uint64_t vkey = pack3(d.cx, d.cy, d.cz);
      if ((clearing || vkey != own) && bundle_exists(keys, P.n, vkey)) continue;
    }
    const int bx = d.cx >> P.L, by = d.cy >> P.L, bz = d.cz >> P.L;
    if (bx != lbx || by != lby || bz != lbz) {
      const uint32_t hp = find_block(tab, pack3(bx, by, bz));
      rank = tab.htouch_rank[hp];
      tab.slot_updated[tab.hslot[hp]] = 7;  // (*last_block)->updated().set(), cc:128
      lbx = bx;
      lby = by;
      lbz = bz;
    }
    const uint32_t lin = (uint32_t)(d.cx & mask) | ((uint32_t)(d.cy & mask) << P.L) |
                         ((uint32_t)(d.cz & mask) << (2 * P.L));
    ckeys[base + emitted] = (rank << (3 * P.L)) | lin;
    cvals[base + emitted] = i;
    ++emitted;
  }
}

// One thread per run of equal voxel ids: the run is that voxel's updates in ray-rank
// order; apply them one after the other exactly like updateTsdfVoxel (cc:150-209).
__global__ void k_apply(ScanParams P, Tables tab, const uint32_t* __restrict__ ckeys,
                        const uint32_t* __restrict__ cvals, unsigned long long total,
                        const float4* __restrict__ ray_p, const uint2* __restrict__ ray_c, ScanState* st) {
  const unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  bool head = false;
  if (e < total) {
    const uint32_t key = ckeys[e];
    head = (e == 0) || (ckeys[e - 1] != key);
    if (head) {
      const uint32_t rank = key >> (3 * P.L);
      const uint32_t lin = key & ((1u << (3 * P.L)) - 1u);
      const uint32_t hp = tab.touched_list[rank];
      int bx, by, bz;
      unpack3(tab.hkeys[hp], &bx, &by, &bz);
      const int mask = (1 << P.L) - 1;
      const int vx = (bx << P.L) + (int)(lin & mask);
      const int vy = (by << P.L) + (int)((lin >> P.L) & mask);
      const int vz = (bz << P.L) + (int)(lin >> (2 * P.L));
      TsdfVoxel* vp = tab.tsdf + (((size_t)tab.hslot[hp]) << (3 * P.L)) + lin;
      TsdfVoxel v = *vp;
      for (unsigned long long j = e; j < total && ckeys[j] == key; ++j) {
        const uint32_t r = cvals[j];
        const float4 rp = ray_p[r];
        const float sdf = ray_sdf(P.origin, f3(rp.x, rp.y, rp.z), vx, vy, vz, P.voxel_size);
        const float w = update_weight(sdf, rp.w, P.up);
        apply_update(v, sdf, w, ray_c[r].x, P.up);
      }
      *vp = v;
    }
  }
  const unsigned b = __ballot_sync(0xffffffffu, head);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(&st->n_voxels, (uint32_t)__popc(b));
}

// --------------------------------------------------------------------- host side
static inline unsigned int grid_for(uint64_t n, int block) { return (unsigned int)((n + block - 1) / block); }

static int bits_for(uint64_t v) {
  int b = 0;
  while (v > 0) {
    ++b;
    v >>= 1;
  }
  return b;
}

static int check_state_errors(vbx_ctx* c, uint32_t err) {
  if (!err) return VBX_OK;
  std::string m = "device reported:";
  if (err & kErrPoolFull) m += " block pool full (raise vbx_engine_options.max_blocks);";
  if (err & kErrHashFull) m += " block hash full;";
  if (err & kErrCoordRange) m += " voxel coordinate outside +-2^20 blocks;";
  if (err & kErrUpdatesFull) m += " ray-voxel updates exceed max_updates_per_pass;";
  return fail(c, VBX_E_CAPACITY, m);
}

int integrate_device(vbx_ctx* c, int kind, const float q[4], const float t[3], const float* d_xyz,
                     const uint8_t* d_rgba, uint64_t n64, int freespace) {
  if (kind < VBX_SIMPLE || kind > VBX_FAST) return fail(c, VBX_E_INVALID, "Unknown TSDF integrator type");
  if (n64 > c->max_points) return fail(c, VBX_E_CAPACITY, "cloud larger than max_points_per_scan");
  const uint32_t n = (uint32_t)n64;
  cudaStream_t s = c->stream;
  const vbx_tsdf_config& cfg = c->cfg;
  std::memset(c->counters, 0, sizeof(c->counters));
  uint64_t launches = 0;

  ScanParams P;
  P.T.w = q[0];
  P.T.x = q[1];
  P.T.y = q[2];
  P.T.z = q[3];
  P.T.t = f3(t[0], t[1], t[2]);
  P.origin = P.T.t;  // T_G_C.getPosition()
  P.voxel_size = c->voxel_size;
  P.voxel_size_inv = c->voxel_size_inv;
  P.trunc = cfg.default_truncation_distance;
  P.min_ray = cfg.min_ray_length_m;
  P.max_ray = cfg.max_ray_length_m;
  P.up.trunc = cfg.default_truncation_distance;
  P.up.max_weight = cfg.max_weight;
  P.up.voxel_size = c->voxel_size;
  P.up.use_weight_dropoff = cfg.use_weight_dropoff;
  P.up.use_sparsity = cfg.use_sparsity_compensation_factor;
  P.up.sparsity_factor = cfg.sparsity_compensation_factor;
  P.L = c->L;
  P.kind = kind;
  P.freespace = freespace;
  P.use_const_weight = cfg.use_const_weight;
  P.allow_clear = cfg.allow_clear;
  P.carving = cfg.voxel_carving_enabled;
  P.anti_grazing = cfg.enable_anti_grazing;
  P.order_mode = cfg.integration_order_mode;
  P.n = n;
  P.n_groups = n / 1024u;
  P.start_inv = cfg.start_voxel_subsampling_factor * c->voxel_size_inv;
  P.max_collisions = cfg.max_consecutive_ray_collisions;
  P.max_updates = c->max_updates;
  c->epoch += 1;
  P.epoch = c->epoch;
  if (kind == VBX_FAST) {
    // resetApproxSet every clear_checks_every_n_frames calls (cc:563-568)
    if ((++c->fast_reset_counter) >= cfg.clear_checks_every_n_frames) {
      c->fast_reset_counter = 0;
      c->set_epoch += 1;
    }
  }
  P.set_epoch = c->set_epoch;

  VBX_CUDA(c, cudaEventRecord(c->ev0, s));
  VBX_CUDA(c, cudaMemsetAsync(c->d_state, 0, sizeof(ScanState), s));
  if (n == 0) {
    VBX_CUDA(c, cudaEventRecord(c->ev1, s));
    VBX_CUDA(c, cudaStreamSynchronize(s));
    c->last_ms = 0.f;
    return VBX_OK;
  }
  const int TB = 256;
  int n_marks = 0;
  int mark_stage[20];
  auto mark = [&](int stage_just_finished) {
    if (c->profiling && n_marks < 19) {
      cudaEventRecord(c->sev[n_marks + 1], s);
      mark_stage[n_marks++] = stage_just_finished;
    }
  };
  if (c->profiling) cudaEventRecord(c->sev[0], s);

  const uint32_t* order = nullptr;
  if (cfg.integration_order_mode == 1) {
    // SortedThreadSafeIndex: ascending |p|^2 (stable here; std::sort leaves ties unspecified)
    k_sqnorm_keys<<<grid_for(n, TB), TB, 0, s>>>(n, d_xyz, c->pkeys[0], c->pvals[0]);
    cub::DoubleBuffer<uint64_t> kb(c->pkeys[0], c->pkeys[1]);
    cub::DoubleBuffer<uint32_t> vb(c->pvals[0], c->pvals[1]);
    size_t tmp = c->cub_tmp_bytes;
    VBX_CUDA(c, cub::DeviceRadixSort::SortPairs(c->cub_tmp, tmp, kb, vb, (int)n, 0, 64, s));
    VBX_CUDA(c, cudaMemcpyAsync(c->order, vb.Current(), n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    order = c->order;
    launches += 10;
  }

  const uint64_t* keys = nullptr;
  const uint32_t* vals = nullptr;
  if (kind == VBX_MERGED) {
    k_point_keys<<<grid_for(n, TB), TB, 0, s>>>(P, d_xyz, order, c->pkeys[0], c->pvals[0], c->d_state);
    mark(0);
    cub::DoubleBuffer<uint64_t> kb(c->pkeys[0], c->pkeys[1]);
    cub::DoubleBuffer<uint32_t> vb(c->pvals[0], c->pvals[1]);
    size_t tmp = c->cub_tmp_bytes;
    VBX_CUDA(c, cub::DeviceRadixSort::SortPairs(c->cub_tmp, tmp, kb, vb, (int)n, 0, 64, s));
    keys = kb.Current();
    vals = vb.Current();
    launches += 10;
    mark(1);
  }

  k_rays_count<<<grid_for((uint64_t)n + 1, 128), 128, 0, s>>>(P, c->tab, d_xyz, d_rgba, order, keys, vals,
                                                                c->ray_p, c->ray_c, c->cnt, c->set_start,
                                                                c->set_observed, c->d_state);
  mark(2);
  {
    size_t tmp = c->cub_tmp_bytes;
    VBX_CUDA(c, cub::DeviceScan::ExclusiveSum(c->cub_tmp, tmp, c->cnt, c->off, (int)(n + 1), s));
  }
  mark(3);
  k_assign<<<grid_for(c->tab.max_blocks, TB), TB, 0, s>>>(c->tab, c->off, n, c->n_blocks, c->max_updates,
                                                          c->d_state);
  mark(4);
  launches += 4;
  VBX_CUDA(c, cudaMemcpyAsync(c->h_state, c->d_state, sizeof(ScanState), cudaMemcpyDeviceToHost, s));
  VBX_CUDA(c, cudaStreamSynchronize(s));
  if (int rc = check_state_errors(c, c->h_state->error)) return rc;
  const unsigned long long K = c->h_state->total_updates;
  const uint32_t n_touched = c->h_state->n_touched;
  c->n_blocks = c->h_state->n_blocks;

  if (K > 0) {
    k_rays_emit<<<grid_for(n, 128), 128, 0, s>>>(P, c->tab, keys, c->ray_p, c->ray_c, c->cnt, c->off,
                                                  c->ckeys[0], c->cvals[0]);
    mark(5);
    cub::DoubleBuffer<uint32_t> kb(c->ckeys[0], c->ckeys[1]);
    cub::DoubleBuffer<uint32_t> vb(c->cvals[0], c->cvals[1]);
    size_t tmp = c->cub_tmp_bytes;
    const int key_bits = 3 * c->L + std::max(1, bits_for(n_touched > 0 ? n_touched - 1 : 0));
    VBX_CUDA(c, cub::DeviceRadixSort::SortPairs(c->cub_tmp, tmp, kb, vb, (int)K, 0, key_bits, s));
    mark(6);
    k_apply<<<grid_for(K, TB), TB, 0, s>>>(P, c->tab, kb.Current(), vb.Current(), K, c->ray_p, c->ray_c,
                                           c->d_state);
    mark(7);
    launches += 3 + (key_bits + 7) / 8;
  }
  VBX_CUDA(c, cudaEventRecord(c->ev1, s));
  VBX_CUDA(c, cudaMemcpyAsync(c->h_state, c->d_state, sizeof(ScanState), cudaMemcpyDeviceToHost, s));
  VBX_CUDA(c, cudaStreamSynchronize(s));
  VBX_CUDA(c, cudaGetLastError());
  VBX_CUDA(c, cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
  for (int m = 0; m < n_marks; ++m) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->sev[m], c->sev[m + 1]) == cudaSuccess) {
      c->stage_ms[mark_stage[m]] += ms;
      c->stage_calls[mark_stage[m]] += 1;
    }
  }
  c->launches += launches;
  c->counters[0] = c->h_state->n_rays;
  c->counters[1] = c->h_state->n_clear_rays;
  c->counters[2] = K;
  c->counters[3] = c->h_state->n_voxels;
  c->counters[4] = n_touched;
  c->counters[5] = c->h_state->n_new;
  c->counters[6] = (kind == VBX_MERGED) ? c->h_state->n_valid_points
                                        : (uint64_t)c->h_state->n_rays + c->h_state->n_clear_rays;
  c->counters[7] = launches;
  return VBX_OK;
}

size_t cub_temp_bytes(uint32_t max_points, uint64_t max_updates) {
  size_t a = 0, b = 0, d = 0;
  cub::DoubleBuffer<uint64_t> k64(nullptr, nullptr);
  cub::DoubleBuffer<uint32_t> k32(nullptr, nullptr), v32(nullptr, nullptr);
  cub::DeviceRadixSort::SortPairs(nullptr, a, k64, v32, (int)max_points, 0, 64);
  cub::DeviceRadixSort::SortPairs(nullptr, b, k32, v32, (int)std::min<uint64_t>(max_updates, 0x7fffffffull), 0, 32);
  cub::DeviceScan::ExclusiveSum(nullptr, d, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)max_points + 1);
  return std::max(a, std::max(b, d)) + 256;
}

}  // namespace vbx
