// Host check of the formulation behind the device's bundle-order kernels
// (voxblox_b200/csrc/vbx_order.cuh): the iteration order of a libstdc++
// std::unordered_map filled by operator[] equals
//   position(b) = #elements in bucket groups created later than b's group + #elements of b's group inserted later
// applied once per rehash (with "insertion time" = list position before the rehash) along the growth
// schedule of the library's own _Prime_rehash_policy, computed with the same skipping loop as
// init_bundle_order (vbx_tsdf.cu).  Compared against a real std::unordered_map for many sizes and hash
// distributions, including the hash of the reference's voxel maps (LongIndexHash, core/block_hash.h:52-64).
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <random>
#include <unordered_map>
#include <vector>

struct Preset {
  const std::vector<uint32_t>* h;
  size_t operator()(uint32_t key) const { return static_cast<size_t>((*h)[key]); }
};

struct Schedule {
  std::vector<uint32_t> m, n;
};
// init_bundle_order's loop
static Schedule schedule(size_t limit) {
  Schedule s;
  std::__detail::_Prime_rehash_policy pol;
  size_t buckets = 1, e = 0;
  while (e < limit) {
    const auto r = pol._M_need_rehash(buckets, e, 1);
    if (r.first) {
      buckets = r.second;
      s.m.push_back((uint32_t)e);
      s.n.push_back((uint32_t)buckets);
    }
    e = std::max(e + 1, (size_t)pol._M_next_resize);
  }
  return s;
}

static void positions(const std::vector<uint32_t>& h, const std::vector<uint32_t>& tau, std::vector<uint32_t>& out,
                      uint32_t m, uint32_t n) {
  const uint32_t NIL = ~0u;
  std::vector<uint32_t> head(n, NIL), next(m), A(m, 0);
  for (uint32_t b = 0; b < m; ++b) {
    next[b] = head[h[b] % n];
    head[h[b] % n] = b;
  }
  for (uint32_t b = 0; b < m; ++b) {
    uint32_t cmin = NIL, size = 0;
    for (uint32_t c = head[h[b] % n]; c != NIL; c = next[c]) {
      cmin = std::min(cmin, tau[c]);
      ++size;
    }
    A[tau[b]] = cmin == tau[b] ? size : 0;
  }
  uint32_t run = 0;
  for (uint32_t t = m; t-- > 0;) {
    const uint32_t v = A[t];
    A[t] = run;
    run += v;
  }
  for (uint32_t b = 0; b < m; ++b) {
    uint32_t cmin = NIL, later = 0;
    for (uint32_t c = head[h[b] % n]; c != NIL; c = next[c]) {
      cmin = std::min(cmin, tau[c]);
      later += tau[c] > tau[b];
    }
    out[b] = A[cmin] + later;
  }
}

static bool check(const std::vector<uint32_t>& h) {
  const uint32_t B = (uint32_t)h.size();
  std::unordered_map<uint32_t, int, Preset> ref(0, Preset{&h});
  // (a bucket hint of 0 leaves the single static bucket and an untouched policy, like a default-constructed map)
  for (uint32_t e = 0; e < B; ++e) ref[e] = 0;
  std::vector<uint32_t> want;
  for (const auto& kv : ref) want.push_back(kv.first);
  const Schedule s = schedule(B + 1);
  std::vector<uint32_t> tau(B), oth(B);
  for (uint32_t e = 0; e < B; ++e) tau[e] = e;
  uint32_t n_cur = 1;
  for (size_t k = 0; k < s.m.size() && s.m[k] < B; ++k) {
    if (s.m[k] > 0) {
      positions(h, tau, oth, s.m[k], n_cur);
      for (uint32_t e = s.m[k]; e < B; ++e) oth[e] = e;
      tau.swap(oth);
    }
    n_cur = s.n[k];
  }
  if (n_cur != ref.bucket_count()) return false;
  positions(h, tau, oth, B, n_cur);
  std::vector<uint32_t> got(B);
  for (uint32_t e = 0; e < B; ++e) got[oth[e]] = e;
  return got == want;
}

int main() {
  std::mt19937_64 rng(7);
  int failures = 0, cases = 0;
  const uint32_t sizes[] = {1, 2, 12, 13, 14, 28, 29, 30, 59, 60, 127, 128, 257, 258, 541, 542, 1000, 1109, 1110,
                            2357, 2358, 5087, 5088, 5300, 10273, 10274, 20753, 20754, 49152};
  for (uint32_t B : sizes) {
    for (int kind = 0; kind < 3; ++kind) {
      std::vector<uint32_t> h(B);
      for (uint32_t e = 0; e < B; ++e) {
        if (kind == 0) {
          h[e] = (uint32_t)rng();
        } else if (kind == 1) {  // LongIndexHash of voxels on a surface patch
          const int64_t x = (int64_t)(rng() % 120) - 60, y = (int64_t)(rng() % 120) - 60, z = (int64_t)(rng() % 6) - 3;
          h[e] = (uint32_t)(x + y * 17191 + z * 17191 * 17191);
        } else {
          h[e] = (uint32_t)(rng() % 7) * 5087u;  // long bucket chains
        }
      }
      if (kind == 2 && B > 6000) continue;
      ++cases;
      if (!check(h)) {
        std::printf("MISMATCH B=%u kind=%d\n", B, kind);
        ++failures;
      }
    }
  }
  std::printf("%d cases, %d failures\n", cases, failures);
  return failures ? 1 : 0;
}
