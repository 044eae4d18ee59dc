// Test harness (tests/ only): compiles the PRODUCT's device arithmetic header
// voxblox_b200/csrc/vbx_math.cuh for the host and drives it ray by ray, so the
// float32 index / update arithmetic can be checked against the oracle without a GPU.
// It is not a fallback: nothing in voxblox_b200/ links or loads it.
#include <cstdint>
#include <cstring>
#include <map>
#include <tuple>
#include <vector>

#include "../voxblox_b200/csrc/vbx_math.cuh"

using namespace vbx;

struct Sim {
  std::map<std::tuple<int, int, int>, TsdfVoxel> vox;
};

extern "C" {
void* sim_create() { return new Sim; }
void sim_destroy(void* s) { delete static_cast<Sim*>(s); }

// Simple-integrator semantics, "mixed" point order, one thread.
void sim_integrate_simple(void* sv, const float* q, const float* t, const float* xyz, const uint8_t* rgba,
                          uint64_t n, float voxel_size, float trunc, float max_weight, float min_ray,
                          float max_ray, int from_origin) {
  Sim* S = static_cast<Sim*>(sv);
  Pose T{q[0], q[1], q[2], q[3], f3(t[0], t[1], t[2])};
  const float inv = (float)(1.0 / voxel_size);
  UpdateParams P{trunc, max_weight, voxel_size, 1, 0, 1.0f};
  const uint64_t groups = n / 1024;
  for (uint64_t s = 0; s < n; ++s) {
    const uint64_t idx = (groups * 1024 <= s) ? s : (s % groups) * 1024 + s / groups;
    const F3 p = f3(xyz[3 * idx], xyz[3 * idx + 1], xyz[3 * idx + 2]);
    const int cls = classify_point(p, min_ray, max_ray, true, false);
    if (!cls) continue;
    const F3 pg = transform(T, p);
    const float w0 = point_weight(p.z, false);
    uint32_t col;
    std::memcpy(&col, rgba + 4 * idx, 4);
    Dda d;
    dda_setup(d, T.t, pg, cls == 2, true, max_ray, inv, trunc, from_origin != 0);
    for (unsigned k = 0; k <= d.len; ++k, dda_advance(d)) {
      TsdfVoxel& v = S->vox.emplace(std::make_tuple(d.cx, d.cy, d.cz), TsdfVoxel{0.f, 0.f, 0u}).first->second;
      const float sdf = ray_sdf(T.t, pg, d.cx, d.cy, d.cz, voxel_size);
      apply_update(v, sdf, update_weight(sdf, w0, P), col, P);
    }
  }
}
uint64_t sim_count(void* sv) { return static_cast<Sim*>(sv)->vox.size(); }
void sim_dump(void* sv, int32_t* idx, float* dist, float* weight, uint32_t* color) {
  uint64_t i = 0;
  for (const auto& kv : static_cast<Sim*>(sv)->vox) {
    idx[3 * i] = std::get<0>(kv.first);
    idx[3 * i + 1] = std::get<1>(kv.first);
    idx[3 * i + 2] = std::get<2>(kv.first);
    dist[i] = kv.second.distance;
    weight[i] = kv.second.weight;
    color[i] = kv.second.color;
    ++i;
  }
}
}
