// TEST INFRASTRUCTURE ONLY (oracle/): wraps the reference's OWN integrator
// classes -- compiled verbatim from /root/reference/voxblox/src/integrator/*.cc
// against oracle/shim/ -- behind the oracle C interface (oracle/vbo_api.h).
// Built only where /root/reference exists (oracle/Makefile, target _ref); the
// resulting oracle/_ref/libvbx_ref.so is git-ignored and travels to the GPU box.
// No reference source is copied: this file only CALLS the reference's public
// API (TsdfIntegratorFactory::create, integratePointCloud, EsdfIntegrator).
#include <algorithm>
#include <chrono>
#include <cstring>
#include <memory>
#include <vector>

#include <numeric>
#include <random>

#include "voxblox/alignment/icp.h"
#include "voxblox/core/layer.h"
#include "voxblox/core/voxel.h"
#include "voxblox/integrator/esdf_integrator.h"
#include "voxblox/integrator/tsdf_integrator.h"
#include "voxblox/mesh/marching_cubes.h"
#include "voxblox/mesh/mesh_integrator.h"
#include "voxblox/mesh/mesh_layer.h"

#include "vbo_api.h"

namespace {

struct Handle {
  voxblox::TsdfIntegratorBase::Config cfg;
  std::unique_ptr<voxblox::Layer<voxblox::TsdfVoxel>> tsdf;
  std::unique_ptr<voxblox::Layer<voxblox::EsdfVoxel>> esdf;
  voxblox::TsdfIntegratorBase::Ptr integrators[4];
  std::unique_ptr<voxblox::EsdfIntegrator> esdf_integrator;
  std::unique_ptr<voxblox::MeshLayer> mesh;
  double last_seconds = 0.0;
};

static_assert(sizeof(voxblox::TsdfVoxel) == 12, "TsdfVoxel layout");
static_assert(sizeof(voxblox::EsdfVoxel) == 20, "EsdfVoxel layout");

template <typename V>
void sortedIndices(const voxblox::Layer<V>& layer, int32_t* out) {
  voxblox::BlockIndexList blocks;
  layer.getAllAllocatedBlocks(&blocks);
  std::sort(blocks.begin(), blocks.end(),
            [](const voxblox::BlockIndex& a, const voxblox::BlockIndex& b) {
              if (a.x() != b.x()) return a.x() < b.x();
              if (a.y() != b.y()) return a.y() < b.y();
              return a.z() < b.z();
            });
  for (size_t i = 0; i < blocks.size(); ++i) {
    out[3 * i + 0] = blocks[i].x();
    out[3 * i + 1] = blocks[i].y();
    out[3 * i + 2] = blocks[i].z();
  }
}

template <typename V>
int getBlock(const voxblox::Layer<V>& layer, const int32_t idx[3], void* voxels,
             uint8_t* updated_bits) {
  typename voxblox::Block<V>::ConstPtr block =
      layer.getBlockPtrByIndex(voxblox::BlockIndex(idx[0], idx[1], idx[2]));
  if (!block) return 1;
  std::memcpy(voxels, &block->getVoxelByLinearIndex(0), block->num_voxels() * sizeof(V));
  if (updated_bits) {
    *updated_bits = static_cast<uint8_t>(block->updated().to_ulong());
  }
  return 0;
}

}  // namespace

template <typename V>
static int serializeBlock(voxblox::Layer<V>& layer, const int32_t idx[3], uint32_t* words) {
  typename voxblox::Block<V>::Ptr b = layer.getBlockPtrByIndex(voxblox::BlockIndex(idx[0], idx[1], idx[2]));
  if (!b) return 1;
  std::vector<uint32_t> data;
  b->serializeToIntegers(&data);
  std::memcpy(words, data.data(), data.size() * sizeof(uint32_t));
  return 0;
}
template <typename V>
static int deserializeBlock(voxblox::Layer<V>& layer, const int32_t idx[3], const uint32_t* words, size_t per_voxel) {
  typename voxblox::Block<V>::Ptr b = layer.allocateBlockPtrByIndex(voxblox::BlockIndex(idx[0], idx[1], idx[2]));
  std::vector<uint32_t> data(words, words + b->num_voxels() * per_voxel);
  b->deserializeFromIntegers(data);
  return 0;
}

extern "C" {

const char* vbo_impl_name(void) { return "reference"; }

void* vbo_create(const vbo_tsdf_config* c, float voxel_size, int voxels_per_side) {
  Handle* h = new Handle;
  h->cfg.default_truncation_distance = c->default_truncation_distance;
  h->cfg.max_weight = c->max_weight;
  h->cfg.voxel_carving_enabled = c->voxel_carving_enabled != 0;
  h->cfg.min_ray_length_m = c->min_ray_length_m;
  h->cfg.max_ray_length_m = c->max_ray_length_m;
  h->cfg.use_const_weight = c->use_const_weight != 0;
  h->cfg.allow_clear = c->allow_clear != 0;
  h->cfg.use_weight_dropoff = c->use_weight_dropoff != 0;
  h->cfg.use_sparsity_compensation_factor = c->use_sparsity_compensation_factor != 0;
  h->cfg.sparsity_compensation_factor = c->sparsity_compensation_factor;
  h->cfg.integrator_threads = static_cast<size_t>(c->integrator_threads);
  h->cfg.integration_order_mode = c->integration_order_mode == 1 ? "sorted" : "mixed";
  h->cfg.enable_anti_grazing = c->enable_anti_grazing != 0;
  h->cfg.start_voxel_subsampling_factor = c->start_voxel_subsampling_factor;
  h->cfg.max_consecutive_ray_collisions = c->max_consecutive_ray_collisions;
  h->cfg.clear_checks_every_n_frames = c->clear_checks_every_n_frames;
  h->cfg.max_integration_time_s = c->max_integration_time_s;
  h->tsdf.reset(new voxblox::Layer<voxblox::TsdfVoxel>(voxel_size, voxels_per_side));
  h->esdf.reset(new voxblox::Layer<voxblox::EsdfVoxel>(voxel_size, voxels_per_side));
  return h;
}

void vbo_destroy(void* hv) { delete static_cast<Handle*>(hv); }

int vbo_integrate(void* hv, int kind, const float q[4], const float t[3],
                  const float* xyz, const uint8_t* rgba, uint64_t n, int freespace,
                  int bundle_order) {
  Handle* h = static_cast<Handle*>(hv);
  if (kind < 1 || kind > 3) return 2;
  if (bundle_order != VBO_ORDER_REFERENCE) return 3;  // the reference has one order
  if (!h->integrators[kind]) {
    h->integrators[kind] = voxblox::TsdfIntegratorFactory::create(
        static_cast<voxblox::TsdfIntegratorType>(kind), h->cfg, h->tsdf.get());
  }
  voxblox::Pointcloud points(n);
  voxblox::Colors colors(n);
  for (uint64_t i = 0; i < n; ++i) {
    points[i] = voxblox::Point(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    colors[i] = voxblox::Color(rgba[4 * i], rgba[4 * i + 1], rgba[4 * i + 2], rgba[4 * i + 3]);
  }
  const voxblox::Transformation T_G_C(
      voxblox::Rotation(q[0], q[1], q[2], q[3]), voxblox::Point(t[0], t[1], t[2]));
  const auto t0 = std::chrono::steady_clock::now();
  h->integrators[kind]->integratePointCloud(T_G_C, points, colors, freespace != 0);
  h->last_seconds =
      std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return 0;
}

double vbo_last_seconds(void* hv) { return static_cast<Handle*>(hv)->last_seconds; }

void vbo_last_counters(void*, uint64_t out[8]) { std::memset(out, 0, 8 * sizeof(uint64_t)); }

uint64_t vbo_num_blocks(void* hv, int layer) {
  Handle* h = static_cast<Handle*>(hv);
  return layer == VBO_LAYER_TSDF ? h->tsdf->getNumberOfAllocatedBlocks()
                                 : h->esdf->getNumberOfAllocatedBlocks();
}

void vbo_block_indices(void* hv, int layer, int32_t* out) {
  Handle* h = static_cast<Handle*>(hv);
  if (layer == VBO_LAYER_TSDF) {
    sortedIndices(*h->tsdf, out);
  } else {
    sortedIndices(*h->esdf, out);
  }
}

int vbo_get_block(void* hv, int layer, const int32_t idx[3], void* voxels,
                  uint8_t* updated_bits) {
  Handle* h = static_cast<Handle*>(hv);
  return layer == VBO_LAYER_TSDF ? getBlock(*h->tsdf, idx, voxels, updated_bits)
                                 : getBlock(*h->esdf, idx, voxels, updated_bits);
}

int vbo_serialize_block(void* hv, int layer, const int32_t idx[3], uint32_t* words) {
  Handle* h = static_cast<Handle*>(hv);
  return layer == VBO_LAYER_TSDF ? serializeBlock(*h->tsdf, idx, words) : serializeBlock(*h->esdf, idx, words);
}
int vbo_deserialize_block(void* hv, int layer, const int32_t idx[3], const uint32_t* words) {
  Handle* h = static_cast<Handle*>(hv);
  return layer == VBO_LAYER_TSDF ? deserializeBlock(*h->tsdf, idx, words, 3) : deserializeBlock(*h->esdf, idx, words, 2);
}

int vbo_esdf_create(void* hv, const vbo_esdf_config* c) {
  Handle* h = static_cast<Handle*>(hv);
  voxblox::EsdfIntegrator::Config cfg;
  cfg.full_euclidean_distance = c->full_euclidean_distance != 0;
  cfg.max_distance_m = c->max_distance_m;
  cfg.min_distance_m = c->min_distance_m;
  cfg.default_distance_m = c->default_distance_m;
  cfg.min_diff_m = c->min_diff_m;
  cfg.min_weight = c->min_weight;
  cfg.num_buckets = c->num_buckets;
  cfg.multi_queue = c->multi_queue != 0;
  cfg.add_occupied_crust = c->add_occupied_crust != 0;
  cfg.clear_sphere_radius = c->clear_sphere_radius;
  cfg.occupied_sphere_radius = c->occupied_sphere_radius;
  h->esdf_integrator.reset(new voxblox::EsdfIntegrator(cfg, h->tsdf.get(), h->esdf.get()));
  return 0;
}

int vbo_esdf_update(void* hv, int batch, int clear_updated_flag) {
  Handle* h = static_cast<Handle*>(hv);
  if (!h->esdf_integrator) return 2;
  const auto t0 = std::chrono::steady_clock::now();
  if (batch) {
    h->esdf_integrator->updateFromTsdfLayerBatch();
  } else {
    h->esdf_integrator->updateFromTsdfLayer(clear_updated_flag != 0);
  }
  h->last_seconds =
      std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return 0;
}

int vbo_esdf_update_blocks(void* hv, const int32_t* idx3, uint64_t m, int incremental) {
  Handle* h = static_cast<Handle*>(hv);
  if (!h->esdf_integrator) return 2;
  voxblox::BlockIndexList blocks;
  for (uint64_t i = 0; i < m; ++i) blocks.push_back(voxblox::BlockIndex(idx3[3 * i], idx3[3 * i + 1], idx3[3 * i + 2]));
  h->esdf_integrator->updateFromTsdfBlocks(blocks, incremental != 0);
  return 0;
}

int vbo_esdf_set_max_distance(void* hv, float max_distance) {
  Handle* h = static_cast<Handle*>(hv);
  if (!h->esdf_integrator) return 2;
  h->esdf_integrator->setEsdfMaxDistance(max_distance);
  return 0;
}

int vbo_esdf_set_full_euclidean(void* hv, int full_euclidean) {
  Handle* h = static_cast<Handle*>(hv);
  if (!h->esdf_integrator) return 2;
  h->esdf_integrator->setFullEuclidean(full_euclidean != 0);
  return 0;
}

int vbo_esdf_add_robot_position(void* hv, const float p[3]) {
  Handle* h = static_cast<Handle*>(hv);
  if (!h->esdf_integrator) return 2;
  h->esdf_integrator->addNewRobotPosition(voxblox::Point(p[0], p[1], p[2]));
  return 0;
}


int vbo_mesh_generate(void* hv, int use_color, float min_weight, int only_mesh_updated_blocks,
                      int clear_updated_flag) {
  Handle* h = static_cast<Handle*>(hv);
  if (!h->mesh) h->mesh.reset(new voxblox::MeshLayer(h->tsdf->block_size()));
  voxblox::MeshIntegratorConfig cfg;
  cfg.use_color = use_color != 0;
  cfg.min_weight = min_weight;
  cfg.integrator_threads = 1;
  voxblox::MeshIntegrator<voxblox::TsdfVoxel> integrator(cfg, h->tsdf.get(), h->mesh.get());
  const auto t0 = std::chrono::steady_clock::now();
  integrator.generateMesh(only_mesh_updated_blocks != 0, clear_updated_flag != 0);
  h->last_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return 0;
}

uint64_t vbo_mesh_num_blocks(void* hv) {
  Handle* h = static_cast<Handle*>(hv);
  if (!h->mesh) return 0;
  voxblox::BlockIndexList have;
  h->mesh->getAllAllocatedMeshes(&have);
  return have.size();
}

void vbo_mesh_block_indices(void* hv, int32_t* out) {
  Handle* h = static_cast<Handle*>(hv);
  if (!h->mesh) return;
  voxblox::BlockIndexList blocks;
  h->mesh->getAllAllocatedMeshes(&blocks);
  std::sort(blocks.begin(), blocks.end(), [](const voxblox::BlockIndex& a, const voxblox::BlockIndex& b) {
    if (a.x() != b.x()) return a.x() < b.x();
    if (a.y() != b.y()) return a.y() < b.y();
    return a.z() < b.z();
  });
  for (size_t i = 0; i < blocks.size(); ++i) {
    out[3 * i + 0] = blocks[i].x();
    out[3 * i + 1] = blocks[i].y();
    out[3 * i + 2] = blocks[i].z();
  }
}

uint64_t vbo_mesh_get(void* hv, const int32_t idx[3], float* vertices, float* normals, uint8_t* colors,
                      int* has_colors, int* updated) {
  Handle* h = static_cast<Handle*>(hv);
  const voxblox::BlockIndex bi(idx[0], idx[1], idx[2]);
  if (!h->mesh) return ~0ull;
  voxblox::BlockIndexList have;
  h->mesh->getAllAllocatedMeshes(&have);
  if (std::find(have.begin(), have.end(), bi) == have.end()) return ~0ull;
  voxblox::Mesh::ConstPtr m = static_cast<const voxblox::MeshLayer*>(h->mesh.get())->getMeshPtrByIndex(bi);
  const size_t n = m->vertices.size();
  for (size_t i = 0; i < n; ++i) {
    if (m->indices[i] != i) return ~0ull - 1;  // (never: marching_cubes.h:97-99)
    for (int k = 0; k < 3; ++k) {
      if (vertices) vertices[3 * i + k] = m->vertices[i](k);
      if (normals) normals[3 * i + k] = m->normals[i](k);
    }
    if (colors && m->colors.size() == n) {
      colors[4 * i + 0] = m->colors[i].r;
      colors[4 * i + 1] = m->colors[i].g;
      colors[4 * i + 2] = m->colors[i].b;
      colors[4 * i + 3] = m->colors[i].a;
    }
  }
  if (has_colors) *has_colors = (n > 0 && m->colors.size() == n) ? 1 : 0;
  if (updated) *updated = m->updated ? 1 : 0;
  return n;
}

int vbo_mc_tables(int32_t tri[256 * 16], int32_t edges[12 * 2]) {
  for (int i = 0; i < 256; ++i) {
    for (int j = 0; j < 16; ++j) tri[16 * i + j] = voxblox::MarchingCubes::kTriangleTable[i][j];
  }
  for (int i = 0; i < 12; ++i) {
    edges[2 * i] = voxblox::MarchingCubes::kEdgeIndexPairs[i][0];
    edges[2 * i + 1] = voxblox::MarchingCubes::kEdgeIndexPairs[i][1];
  }
  return 0;
}


}  // extern "C"
namespace {
const uint32_t* g_preset_hashes = nullptr;
struct PresetHash {
  size_t operator()(uint32_t key) const { return static_cast<size_t>(g_preset_hashes[key]); }
};
}  // namespace
extern "C" void vbo_umap_order(const uint32_t* hashes, uint64_t n, uint32_t* out) {
  g_preset_hashes = hashes;
  std::unordered_map<uint32_t, int, PresetHash> m;  // default-constructed like tsdf_integrator.cc:318-322
  for (uint64_t i = 0; i < n; ++i) m[static_cast<uint32_t>(i)] = 0;
  uint64_t p = 0;
  for (const auto& kv : m) out[p++] = kv.first;
}

extern "C" {
}  // extern "C"

// ---------------------------------------------------------------- ICP (src/alignment/icp.cc)
extern "C" int vbo_icp_run(void* hv, const vbo_icp_config* c, const float* xyz, uint64_t n, const float q[4],
                           const float t[3], uint32_t seed, float out_q[4], float out_t[3], uint64_t* num_updates) {
  Handle* h = static_cast<Handle*>(hv);
  if (c->num_threads != 1) return 3;  // the reference's threads race: only one thread is a function of its input
  voxblox::ICP::Config cfg;
  cfg.refine_roll_pitch = c->refine_roll_pitch != 0;
  cfg.mini_batch_size = c->mini_batch_size;
  cfg.min_match_ratio = c->min_match_ratio;
  cfg.subsample_keep_ratio = c->subsample_keep_ratio;
  cfg.inital_translation_weighting = c->inital_translation_weighting;
  cfg.inital_rotation_weighting = c->inital_rotation_weighting;
  cfg.num_threads = static_cast<size_t>(c->num_threads);
  voxblox::ICP icp(cfg);
  voxblox::Pointcloud points(n);
  for (uint64_t i = 0; i < n; ++i) points[i] = voxblox::Point(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
  const voxblox::Transformation T_init(voxblox::Rotation(q[0], q[1], q[2], q[3]), voxblox::Point(t[0], t[1], t[2]));
  voxblox::Transformation T_out;
  const auto t0 = std::chrono::steady_clock::now();
  const size_t updates = icp.runICP(*h->tsdf, points, T_init, &T_out, seed);
  h->last_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  out_q[0] = T_out.getRotation().w();
  out_q[1] = T_out.getRotation().x();
  out_q[2] = T_out.getRotation().y();
  out_q[3] = T_out.getRotation().z();
  for (int i = 0; i < 3; ++i) out_t[i] = T_out.getPosition()[i];
  if (num_updates) *num_updates = updates;
  return 0;
}

extern "C" void vbo_icp_shuffle(uint64_t n, uint32_t seed, uint32_t* out) {
  std::iota(out, out + n, 0u);
  std::shuffle(out, out + n, std::default_random_engine(seed));
}
