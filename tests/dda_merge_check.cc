// Test harness (tests/ only): the three-way-merge form of the ray walk (vbx_math.cuh, dda_rank) against
// the sequential walk (dda_advance = nextRayIndex, integrator_utils.cc:106-125) on the host, over
// random and adversarial rays.  Prints "rays R regular G fallback F mismatches M".
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../voxblox_b200/csrc/vbx_math.cuh"

using namespace vbx;

struct V3i {
  int x, y, z;
};

static bool merge_walk(const Dda& d0, std::vector<V3i>* out, int cap) {
  if (!dda_is_regular(d0)) return false;
  const unsigned int len = d0.len;
  const unsigned int n_axis[3] = {d0.nx, d0.ny, d0.nz};
  const float t0[3] = {d0.tx, d0.ty, d0.tz}, dt[3] = {d0.dx, d0.dy, d0.dz};
  const int sg[3] = {d0.sx, d0.sy, d0.sz};
  std::vector<float> chain[3];
  int K[3];
  for (int a = 0; a < 3; ++a) {
    K[a] = (int)dda_chain_len(n_axis[a], len);
    if (K[a] > cap) return false;
    chain[a].resize(K[a] > 0 ? K[a] : 1);
    float t = t0[a];
    for (int k = 0; k < K[a]; ++k) {
      chain[a][k] = t;
      t = fadd(t, dt[a]);
    }
  }
  const float* T[3] = {chain[0].data(), chain[1].data(), chain[2].data()};
  out->assign(len + 1, V3i{0, 0, 0});
  std::vector<char> seen(len + 1, 0);
  (*out)[0] = V3i{d0.cx, d0.cy, d0.cz};
  seen[0] = 1;
  unsigned int emitted = 0;
  for (int a = 0; a < 3; ++a) {
    for (int k = 0; k < K[a]; ++k) {
      unsigned int rank;
      int c[3];
      const bool trusted = dda_rank(T, K, len, a, k, &rank, c);
      if (rank >= len) continue;
      if (!trusted) return false;
      if (seen[rank + 1]) return false;  // cannot happen for a consistent order
      seen[rank + 1] = 1;
      (*out)[rank + 1] = V3i{d0.cx + sg[0] * c[0], d0.cy + sg[1] * c[1], d0.cz + sg[2] * c[2]};
      ++emitted;
    }
  }
  return emitted == len;
}

int main(int argc, char** argv) {
  const long n_rays = argc > 1 ? std::atol(argv[1]) : 2000000;
  std::mt19937_64 rng(12345);
  std::uniform_real_distribution<float> U(-1.f, 1.f);
  long regular = 0, fallback = 0, mismatches = 0, steps = 0;
  long fb_regime[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  std::vector<V3i> par;
  for (long r = 0; r < n_rays; ++r) {
    const int regime = (int)(r % 8);
    const float vs = (regime & 1) ? 0.05f : 0.1f;
    const float inv = (float)(1.0 / vs);
    F3 origin = f3(3.f * U(rng), 3.f * U(rng), 1.5f * U(rng));
    F3 dir = f3(U(rng), U(rng), U(rng));
    float range = 0.3f + 4.6f * (0.5f + 0.5f * U(rng));
    bool clearing = false;
    if (regime == 2) {  // sensor exactly on a voxel corner, diagonal directions: ties on every step
      origin = f3(vs * (float)(int)(10 * U(rng)), vs * (float)(int)(10 * U(rng)), vs * (float)(int)(10 * U(rng)));
      dir = f3((rng() & 1) ? 1.f : -1.f, (rng() & 1) ? 1.f : -1.f, (rng() & 1) ? 1.f : -1.f);
    } else if (regime == 3) {  // directions with small-integer ratios: frequent exact ties
      dir = f3((float)((int)(rng() % 7) - 3), (float)((int)(rng() % 7) - 3), (float)((int)(rng() % 7) - 3));
      origin = f3(vs * 0.5f * (float)(int)(20 * U(rng)), vs * 0.5f * (float)(int)(20 * U(rng)), vs * 0.5f * (float)(int)(20 * U(rng)));
    } else if (regime == 4) {  // nearly axis-parallel
      dir = f3(U(rng), 1e-6f * U(rng), 1e-7f * U(rng));
    } else if (regime == 5) {  // exactly axis-parallel in one or two axes (irregular)
      dir = f3(U(rng), (rng() & 1) ? 0.f : U(rng), 0.f);
    } else if (regime == 6) {
      clearing = true;
      range = 6.f + 20.f * (0.5f + 0.5f * U(rng));
    } else if (regime == 7) {  // far from the origin of the map: large coordinates, coarse t spacing
      origin = f3(900.f + 50.f * U(rng), -700.f + 50.f * U(rng), 300.f * U(rng));
    }
    const float nrm = norm3(dir);
    if (!(nrm > 0.f)) continue;
    const F3 point = add3(origin, scale3(dir, range / nrm));
    Dda d;
    dda_setup(d, origin, point, clearing, true, 5.0f, inv, 4.f * vs, true);
    if (d.len > 4000) continue;
    Dda ds = d;
    if (merge_walk(d, &par, 256)) {
      ++regular;
      for (unsigned int s = 0; s <= ds.len; ++s, dda_advance(ds)) {
        ++steps;
        if (par[s].x != ds.cx || par[s].y != ds.cy || par[s].z != ds.cz) {
          if (mismatches < 5) {
            std::printf("MISMATCH ray %ld step %u of %u: merge (%d %d %d) walk (%d %d %d)\n", r, s, ds.len, par[s].x,
                        par[s].y, par[s].z, ds.cx, ds.cy, ds.cz);
          }
          ++mismatches;
          break;
        }
      }
    } else {
      ++fallback;
      ++fb_regime[regime];
    }
  }
  std::printf("rays %ld regular %ld fallback %ld mismatches %ld steps %ld\n", n_rays, regular, fallback, mismatches, steps);
  std::printf("fallbacks per regime:");
  for (int i = 0; i < 8; ++i) std::printf(" %ld", fb_regime[i]);
  std::printf("\n");
  return mismatches ? 1 : 0;
}
