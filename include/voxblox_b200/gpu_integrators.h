// voxblox-side adapter: plugs the B200 engine in underneath voxblox's own integrator API.
//
// GpuTsdfIntegrator IS-A voxblox::TsdfIntegratorBase (voxblox/integrator/tsdf_integrator.h:51-198):
// same constructor arguments, same virtual integratePointCloud(), so TsdfServer / TsdfMap /
// user code keep working unchanged.  The voxel blocks live in HBM; the host Layer<TsdfVoxel>
// stays the reference's own container and is refreshed on demand by syncLayer() (only blocks
// whose updated() bits are set are downloaded -- SURVEY.md section 8f N1).
//
// This header is compiled against the REFERENCE's headers (it includes them); it contains no
// reference code.  See INTEGRATION.md for the three-line change to TsdfIntegratorFactory::create.
#ifndef VOXBLOX_B200_GPU_INTEGRATORS_H_
#define VOXBLOX_B200_GPU_INTEGRATORS_H_

#include <algorithm>
#include <cstring>
#include <map>
#include <vector>

#include <voxblox/alignment/icp.h>
#include <voxblox/core/layer.h>
#include <voxblox/core/voxel.h>
#include <voxblox/integrator/esdf_integrator.h>
#include <voxblox/mesh/mesh_integrator.h>
#include <voxblox/mesh/mesh_layer.h>
#include <voxblox/integrator/tsdf_integrator.h>

#include "../voxblox_b200.h"

namespace voxblox {

static_assert(sizeof(TsdfVoxel) == 12, "TsdfVoxel must be the 12-byte struct of core/voxel.h:12-16");
static_assert(sizeof(EsdfVoxel) == 20, "EsdfVoxel must be the 20-byte struct of core/voxel.h:18-37");
static_assert(sizeof(Point) == 12 && sizeof(Color) == 4, "Pointcloud / Colors must be packed xyz / rgba");

namespace gpu_detail {
inline vbx_tsdf_config toPod(const TsdfIntegratorBase::Config& c) {
  vbx_tsdf_config p;
  p.default_truncation_distance = c.default_truncation_distance;
  p.max_weight = c.max_weight;
  p.voxel_carving_enabled = c.voxel_carving_enabled;
  p.min_ray_length_m = c.min_ray_length_m;
  p.max_ray_length_m = c.max_ray_length_m;
  p.use_const_weight = c.use_const_weight;
  p.allow_clear = c.allow_clear;
  p.use_weight_dropoff = c.use_weight_dropoff;
  p.use_sparsity_compensation_factor = c.use_sparsity_compensation_factor;
  p.sparsity_compensation_factor = c.sparsity_compensation_factor;
  p.integrator_threads = static_cast<int32_t>(c.integrator_threads);
  if (c.integration_order_mode == "mixed") {
    p.integration_order_mode = 0;
  } else if (c.integration_order_mode == "sorted") {
    p.integration_order_mode = 1;
  } else {  // ThreadSafeIndexFactory::get, integrator_utils.cc:12
    LOG(FATAL) << "Unknown integration order mode: '" << c.integration_order_mode << "'!";
    p.integration_order_mode = 0;
  }
  p.enable_anti_grazing = c.enable_anti_grazing;
  p.start_voxel_subsampling_factor = c.start_voxel_subsampling_factor;
  p.max_consecutive_ray_collisions = c.max_consecutive_ray_collisions;
  p.clear_checks_every_n_frames = c.clear_checks_every_n_frames;
  p.max_integration_time_s = c.max_integration_time_s;
  return p;
}
inline vbx_esdf_config toPod(const EsdfIntegrator::Config& c) {
  vbx_esdf_config p;
  p.full_euclidean_distance = c.full_euclidean_distance;
  p.max_distance_m = c.max_distance_m;
  p.min_distance_m = c.min_distance_m;
  p.default_distance_m = c.default_distance_m;
  p.min_diff_m = c.min_diff_m;
  p.min_weight = c.min_weight;
  p.num_buckets = c.num_buckets;
  p.multi_queue = c.multi_queue;
  p.add_occupied_crust = c.add_occupied_crust;
  p.clear_sphere_radius = c.clear_sphere_radius;
  p.occupied_sphere_radius = c.occupied_sphere_radius;
  return p;
}
// the reference aborts on errors (glog CHECK / LOG(FATAL)); keep that behaviour
inline void check(vbx_ctx* ctx, int rc, const char* what) {
  if (rc != VBX_OK) LOG(FATAL) << what << " failed (" << rc << "): " << vbx_last_error(ctx);
}
// TSDF layer -> the engine context that holds it (registered by GpuTsdfIntegrator): lets
// GpuEsdfIntegrator / GpuMeshIntegrator be constructed from the reference's own constructor
// arguments (layer pointers), e.g. inside voxblox_ros's EsdfServer, without handing a
// GpuTsdfIntegrator around.  Like the reference's integrators, not thread safe.
inline std::map<const void*, vbx_ctx*>& contextOfLayer() {
  static std::map<const void*, vbx_ctx*> m;
  return m;
}
inline vbx_ctx* lookupContext(const void* tsdf_layer) {
  auto it = contextOfLayer().find(tsdf_layer);
  CHECK(it != contextOfLayer().end()) << "no GpuTsdfIntegrator is attached to this Layer<TsdfVoxel>";
  return it->second;
}
// device -> host: refresh (or create) the host blocks listed by the device.  clear_mask: updated()
// bits cleared ON THE DEVICE for the mirrored blocks -- the host block carries them from then on, so
// the next call with the same mask transfers only what changed since (core/layer.h:194-203 protocol).
template <typename VoxelType>
inline size_t downloadBlocks(vbx_ctx* ctx, int layer_id, int updated_mask, Layer<VoxelType>* layer, int clear_mask = 0) {
  // one call: dirty-block list + payloads (gathered on the device, one copy out)
  uint64_t n = 0;
  check(ctx, vbx_num_blocks(ctx, layer_id, &n), "vbx_num_blocks");
  if (n == 0) return 0;
  const size_t vpb = layer->voxels_per_side() * layer->voxels_per_side() * layer->voxels_per_side();
  std::vector<int32_t> idx;
  std::vector<VoxelType> vox;
  std::vector<uint8_t> upd;
  uint64_t cap = updated_mask ? std::min<uint64_t>(n, 64) : n;
  while (true) {
    idx.resize(3 * cap);
    vox.resize(vpb * cap);
    upd.resize(cap);
    // (a call whose buffer is too small delivers nothing and clears nothing: grow and retry)
    check(ctx, vbx_mirror_updated(ctx, layer_id, updated_mask, clear_mask, idx.data(), vox.data(), upd.data(), cap, &n),
          "vbx_mirror_updated");
    if (n <= cap) break;
    cap = n;
  }
  if (n == 0) return 0;
  for (uint64_t b = 0; b < n; ++b) {
    typename Block<VoxelType>::Ptr block =
        layer->allocateBlockPtrByIndex(BlockIndex(idx[3 * b], idx[3 * b + 1], idx[3 * b + 2]));
    std::memcpy(&block->getVoxelByLinearIndex(0), vox.data() + b * vpb, vpb * sizeof(VoxelType));
    for (int bit = 0; bit < static_cast<int>(Update::kCount); ++bit) {
      if (upd[b] & (1u << bit)) block->updated().set(bit);
    }
  }
  return n;
}
}  // namespace gpu_detail

class GpuTsdfIntegrator : public TsdfIntegratorBase {
 public:
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW

  GpuTsdfIntegrator(TsdfIntegratorType type, const Config& config, Layer<TsdfVoxel>* layer,
                    const vbx_engine_options* options = nullptr)
      : TsdfIntegratorBase(config, layer), type_(type), ctx_(nullptr), pipelined_(false), auto_sync_(false) {
    const vbx_tsdf_config pod = gpu_detail::toPod(config_);
    const int rc = vbx_create(&pod, voxel_size_, static_cast<int>(voxels_per_side_), options, &ctx_);
    if (rc != VBX_OK) LOG(FATAL) << "vbx_create failed (" << rc << "): " << vbx_last_error(ctx_);
    uploadLayer();  // a layer loaded from a file / built by another integrator becomes the device map
    gpu_detail::contextOfLayer()[layer_] = ctx_;
  }
  ~GpuTsdfIntegrator() {
    gpu_detail::contextOfLayer().erase(layer_);
    vbx_destroy(ctx_);
  }

  /// TsdfIntegratorBase::setLayer (tsdf_integrator.cc:68-80): the device map is re-targeted too -- it is
  /// emptied and `layer`'s blocks become its content.  (The base class method is not virtual: call it on
  /// the adapter type.  voxel_size / voxels_per_side of the new layer must equal the old ones, which the
  /// engine was sized for; the reference has no such restriction.)
  void setLayer(Layer<TsdfVoxel>* layer) {
    CHECK_NOTNULL(layer);
    CHECK_EQ(layer->voxel_size(), voxel_size_);
    CHECK_EQ(layer->voxels_per_side(), voxels_per_side_);
    gpu_detail::contextOfLayer().erase(layer_);
    TsdfIntegratorBase::setLayer(layer);
    gpu_detail::check(ctx_, vbx_clear(ctx_, VBX_LAYER_TSDF), "vbx_clear");
    uploadLayer();
    gpu_detail::contextOfLayer()[layer_] = ctx_;
  }

  /// Auto-sync mode for UNMODIFIED host consumers of the layer: every integratePointCloud() ends by
  /// mirroring the blocks it changed into the host Layer and setting their updated() bits there
  /// (Block::updated().set(), tsdf_integrator.cc:128), so code that reads layer_ right after the call --
  /// the reference's own MeshIntegrator, EsdfIntegrator, io::SaveLayer -- sees what it would see after
  /// the reference's call.  The transfer is incremental through the engine's own dirty mark
  /// (VBX_UPDATED_MIRROR); the three updated() bits on the device stay with the device-side consumers.
  void setAutoSync(bool on) {
    auto_sync_ = on;
    if (on) pipelined_ = false;
  }

  void integratePointCloud(const Transformation& T_G_C, const Pointcloud& points_C, const Colors& colors,
                           const bool freespace_points = false) override {
    CHECK_EQ(points_C.size(), colors.size());  // tsdf_integrator.cc:247,312,560
    const float q[4] = {T_G_C.getRotation().w(), T_G_C.getRotation().x(), T_G_C.getRotation().y(),
                        T_G_C.getRotation().z()};
    const Point p = T_G_C.getPosition();
    const float t[3] = {p.x(), p.y(), p.z()};
    const float* xyz = points_C.empty() ? nullptr : reinterpret_cast<const float*>(points_C.data());
    const uint8_t* rgba = colors.empty() ? nullptr : reinterpret_cast<const uint8_t*>(colors.data());
    if (pipelined_ && !points_C.empty()) {
      // Pageable host memory: the runtime stages the copy before the call returns, so the
      // caller may reuse points_C / colors immediately, as after the reference's call.
      gpu_detail::check(ctx_,
                        vbx_tsdf_integrate_async(ctx_, static_cast<int>(type_), q, t, xyz, rgba, points_C.size(),
                                                 freespace_points ? 1 : 0, /*inputs_on_device=*/0),
                        "vbx_tsdf_integrate_async");
    } else {
      gpu_detail::check(ctx_,
                        vbx_tsdf_integrate(ctx_, static_cast<int>(type_), q, t, xyz, rgba, points_C.size(),
                                           freespace_points ? 1 : 0),
                        "vbx_tsdf_integrate");
    }
    if (auto_sync_) gpu_detail::downloadBlocks(ctx_, VBX_LAYER_TSDF, VBX_UPDATED_MIRROR, layer_, VBX_UPDATED_MIRROR);
  }

  /// Pipelined mode: integratePointCloud() enqueues the scan and returns; the transform / bundle
  /// half of the next scan overlaps the ray-cast / update half of the current one on the device.
  /// Every read of the map (syncLayer, the ESDF update, block management) waits for the queued
  /// scans first, and the result is bit-identical to the non-pipelined calls.
  void setPipelined(bool on) {
    if (!on) gpu_detail::check(ctx_, vbx_sync(ctx_), "vbx_sync");
    pipelined_ = on;
  }

  /// Host layer -> device map (Layer::insertBlock for every allocated block).
  void uploadLayer() {
    BlockIndexList blocks;
    layer_->getAllAllocatedBlocks(&blocks);
    if (blocks.empty()) return;
    const size_t vpb = voxels_per_side_ * voxels_per_side_ * voxels_per_side_;
    std::vector<int32_t> idx(3 * blocks.size());
    std::vector<TsdfVoxel> vox(vpb * blocks.size());
    std::vector<uint8_t> upd(blocks.size());
    for (size_t b = 0; b < blocks.size(); ++b) {
      const Block<TsdfVoxel>& block = layer_->getBlockByIndex(blocks[b]);
      idx[3 * b] = blocks[b].x();
      idx[3 * b + 1] = blocks[b].y();
      idx[3 * b + 2] = blocks[b].z();
      std::memcpy(vox.data() + b * vpb, &block.getVoxelByLinearIndex(0), vpb * sizeof(TsdfVoxel));
      uint8_t bits = 0;
      for (int bit = 0; bit < static_cast<int>(Update::kCount); ++bit) {
        if (block.updated()[bit]) bits |= static_cast<uint8_t>(1u << bit);
      }
      upd[b] = bits;
    }
    gpu_detail::check(ctx_, vbx_upload_blocks(ctx_, VBX_LAYER_TSDF, idx.data(), blocks.size(), vox.data(), upd.data()),
                      "vbx_upload_blocks");
  }

  /// removeDistantBlocks (core/layer_inl.h / tsdf_server.cc:314-316) on both copies of the map.
  void removeBlocks(const BlockIndexList& blocks) {
    if (blocks.empty()) return;
    std::vector<int32_t> idx(3 * blocks.size());
    for (size_t b = 0; b < blocks.size(); ++b) {
      idx[3 * b] = blocks[b].x();
      idx[3 * b + 1] = blocks[b].y();
      idx[3 * b + 2] = blocks[b].z();
      layer_->removeBlock(blocks[b]);
    }
    gpu_detail::check(ctx_, vbx_remove_blocks(ctx_, VBX_LAYER_TSDF, idx.data(), blocks.size()), "vbx_remove_blocks");
  }

  /// Bring the host Layer<TsdfVoxel> up to date: downloads every block whose updated() bits
  /// match `updated_mask` (0 = all blocks).  Call it before host code reads the layer (meshing,
  /// saving, interpolation); like the reference, clearing the bits is the consumer's job.
  /// `clear_mask` (default: the mirrored bits themselves): cleared on the device for the blocks delivered, so
  /// that the host block owns them from then on and the next syncLayer(mask) transfers only newer changes.
  /// Pass 0 to leave the device bits alone (e.g. when the device-side ESDF / mesher consume them too).
  size_t syncLayer(int updated_mask = 0, int clear_mask = -1) {
    return gpu_detail::downloadBlocks(ctx_, VBX_LAYER_TSDF, updated_mask, layer_, clear_mask < 0 ? updated_mask : clear_mask);
  }

  vbx_ctx* context() { return ctx_; }

 private:
  TsdfIntegratorType type_;
  vbx_ctx* ctx_;
  bool pipelined_;
  bool auto_sync_;
};

/// EsdfIntegrator's update entry points (esdf_integrator.h:101-106) on the device map owned by a
/// GpuTsdfIntegrator.
class GpuEsdfIntegrator {
 public:
  typedef EsdfIntegrator::Config Config;
  GpuEsdfIntegrator(const EsdfIntegrator::Config& config, GpuTsdfIntegrator* tsdf, Layer<EsdfVoxel>* esdf_layer)
      : ctx_(CHECK_NOTNULL(tsdf)->context()), esdf_layer_(CHECK_NOTNULL(esdf_layer)), auto_sync_(false) {
    const vbx_esdf_config pod = gpu_detail::toPod(config);
    gpu_detail::check(ctx_, vbx_esdf_create(ctx_, &pod), "vbx_esdf_create");
  }
  /// EsdfIntegrator's own constructor signature (esdf_integrator.h:80-82): the TSDF layer must be the one a
  /// GpuTsdfIntegrator is attached to.  With this, `std::unique_ptr<EsdfIntegrator> esdf_integrator_`
  /// (voxblox_ros/include/voxblox_ros/esdf_server.h:107) becomes `std::unique_ptr<GpuEsdfIntegrator>` and the
  /// construction line stays as it is; auto-sync is on so that the server's host readers of the
  /// ESDF layer (publishing, planning) need no further edits.
  GpuEsdfIntegrator(const EsdfIntegrator::Config& config, Layer<TsdfVoxel>* tsdf_layer, Layer<EsdfVoxel>* esdf_layer)
      : ctx_(gpu_detail::lookupContext(CHECK_NOTNULL(tsdf_layer))), esdf_layer_(CHECK_NOTNULL(esdf_layer)), auto_sync_(true) {
    const vbx_esdf_config pod = gpu_detail::toPod(config);
    gpu_detail::check(ctx_, vbx_esdf_create(ctx_, &pod), "vbx_esdf_create");
  }
  /// mirror the ESDF blocks an update changed into the host Layer<EsdfVoxel> right after the update
  void setAutoSync(bool on) { auto_sync_ = on; }
  /// every ESDF block changed since the last call of this function (incl. the blocks the wavefront reached)
  size_t syncChanged() {
    return gpu_detail::downloadBlocks(ctx_, VBX_LAYER_ESDF, VBX_UPDATED_MIRROR, esdf_layer_, VBX_UPDATED_MIRROR);
  }
  void updateFromTsdfLayer(bool clear_updated_flag) {
    gpu_detail::check(ctx_, vbx_esdf_update(ctx_, 0, clear_updated_flag ? 1 : 0), "vbx_esdf_update");
    if (auto_sync_) syncChanged();
  }
  void updateFromTsdfLayerBatch() {
    gpu_detail::check(ctx_, vbx_esdf_update(ctx_, 1, 0), "vbx_esdf_update");
    if (auto_sync_) syncChanged();
  }
  void updateFromTsdfBlocks(const BlockIndexList& tsdf_blocks, bool incremental = false) {
    std::vector<int32_t> idx(3 * tsdf_blocks.size());
    for (size_t b = 0; b < tsdf_blocks.size(); ++b) {
      idx[3 * b] = tsdf_blocks[b].x();
      idx[3 * b + 1] = tsdf_blocks[b].y();
      idx[3 * b + 2] = tsdf_blocks[b].z();
    }
    gpu_detail::check(ctx_, vbx_esdf_update_blocks(ctx_, idx.data(), tsdf_blocks.size(), incremental ? 1 : 0),
                      "vbx_esdf_update_blocks");
    if (auto_sync_) syncChanged();
  }
  /// esdf_integrator.cc:25-92: free sphere / occupied shell around the robot, queued for the next update
  void addNewRobotPosition(const Point& position) {
    const float p[3] = {position.x(), position.y(), position.z()};
    gpu_detail::check(ctx_, vbx_esdf_add_robot_position(ctx_, p), "vbx_esdf_add_robot_position");
  }
  /// esdf_integrator.h:135-140
  void clear() { gpu_detail::check(ctx_, vbx_esdf_clear(ctx_), "vbx_esdf_clear"); }
  float getEsdfMaxDistance() const {
    vbx_esdf_config c;
    gpu_detail::check(ctx_, vbx_esdf_get_config(ctx_, &c), "vbx_esdf_get_config");
    return c.max_distance_m;
  }
  void setEsdfMaxDistance(float max_distance) {
    gpu_detail::check(ctx_, vbx_esdf_set_max_distance(ctx_, max_distance), "vbx_esdf_set_max_distance");
  }
  bool getFullEuclidean() const {
    vbx_esdf_config c;
    gpu_detail::check(ctx_, vbx_esdf_get_config(ctx_, &c), "vbx_esdf_get_config");
    return c.full_euclidean_distance != 0;
  }
  void setFullEuclidean(bool full_euclidean) {
    gpu_detail::check(ctx_, vbx_esdf_set_full_euclidean(ctx_, full_euclidean ? 1 : 0), "vbx_esdf_set_full_euclidean");
  }
  size_t syncLayer(int updated_mask = 0, int clear_mask = -1) {
    return gpu_detail::downloadBlocks(ctx_, VBX_LAYER_ESDF, updated_mask, esdf_layer_, clear_mask < 0 ? updated_mask : clear_mask);
  }

 private:
  vbx_ctx* ctx_;
  Layer<EsdfVoxel>* esdf_layer_;
  bool auto_sync_;
};

/// MeshIntegrator<TsdfVoxel>::generateMesh (mesh/mesh_integrator.h:132-160) on the device map owned by
/// a GpuTsdfIntegrator; the MeshLayer stays the caller's host object, as in the reference.
class GpuMeshIntegrator {
 public:
  GpuMeshIntegrator(const MeshIntegratorConfig& config, GpuTsdfIntegrator* tsdf, MeshLayer* mesh_layer)
      : config_(config), ctx_(CHECK_NOTNULL(tsdf)->context()), mesh_layer_(CHECK_NOTNULL(mesh_layer)) {}
  /// MeshIntegrator<TsdfVoxel>'s own constructor signature (mesh/mesh_integrator.h:75-90; tsdf_server.cc:110-112)
  GpuMeshIntegrator(const MeshIntegratorConfig& config, Layer<TsdfVoxel>* tsdf_layer, MeshLayer* mesh_layer)
      : config_(config), ctx_(gpu_detail::lookupContext(CHECK_NOTNULL(tsdf_layer))), mesh_layer_(CHECK_NOTNULL(mesh_layer)) {}

  void generateMesh(bool only_mesh_updated_blocks, bool clear_updated_flag) {
    vbx_mesh_config pod;
    pod.use_color = config_.use_color ? 1 : 0;
    pod.min_weight = config_.min_weight;
    uint64_t nb = 0, nv = 0;
    gpu_detail::check(ctx_, vbx_mesh_generate(ctx_, &pod, only_mesh_updated_blocks ? 1 : 0, clear_updated_flag ? 1 : 0, &nb, &nv),
                      "vbx_mesh_generate");
    if (nb == 0) return;
    std::vector<int32_t> idx(3 * nb);
    std::vector<uint64_t> first(nb + 1);
    std::vector<float> vertices(3 * nv), normals(3 * nv);
    std::vector<uint8_t> colors(config_.use_color ? 4 * nv : 0);
    gpu_detail::check(ctx_, vbx_mesh_download(ctx_, idx.data(), first.data(), vertices.data(), normals.data(),
                                              (config_.use_color && nv) ? colors.data() : nullptr),
                      "vbx_mesh_download");
    for (uint64_t b = 0; b < nb; ++b) {
      // allocateMeshPtrByIndex + updateMeshForBlock (mesh_integrator.h:146-149, :238-260)
      Mesh::Ptr mesh = mesh_layer_->allocateMeshPtrByIndex(BlockIndex(idx[3 * b], idx[3 * b + 1], idx[3 * b + 2]));
      mesh->clear();
      const uint64_t lo = first[b], n = first[b + 1] - first[b];
      mesh->vertices.reserve(n);
      mesh->normals.reserve(n);
      mesh->indices.reserve(n);
      if (config_.use_color) mesh->colors.reserve(n);
      for (uint64_t i = lo; i < lo + n; ++i) {
        mesh->vertices.emplace_back(vertices[3 * i], vertices[3 * i + 1], vertices[3 * i + 2]);
        mesh->normals.emplace_back(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]);
        mesh->indices.push_back(i - lo);
        if (config_.use_color) mesh->colors.emplace_back(colors[4 * i], colors[4 * i + 1], colors[4 * i + 2], colors[4 * i + 3]);
      }
      mesh->updated = true;
    }
  }

 private:
  MeshIntegratorConfig config_;
  vbx_ctx* ctx_;
  MeshLayer* mesh_layer_;
};

/// voxblox::ICP (alignment/icp.h:72-233) against the device map: same constructor and runICP signature, so
/// `icp_.reset(new ICP(getICPConfigFromRosParam(nh_private)))` in TsdfServer (voxblox_ros/src/tsdf_server.cc:114)
/// and the call at tsdf_server.cc:259-262 only change the TYPE.  The TSDF layer argument selects the engine
/// context a GpuTsdfIntegrator registered for it; scans still queued by pipelined submission are drained first.
/// Config::num_threads racing host threads become that many warps under a fixed round-robin schedule
/// (include/voxblox_b200.h), i.e. the result is deterministic where the reference's (num_threads > 1) is not.
class GpuICP {
 public:
  explicit GpuICP(const ICP::Config& config) : config_(config) {}

  size_t runICP(const Layer<TsdfVoxel>& tsdf_layer, const Pointcloud& points, const Transformation& inital_T_tsdf_sensor,
                Transformation* refined_T_tsdf_sensor,
                const unsigned seed = std::chrono::system_clock::now().time_since_epoch().count()) {
    CHECK_NOTNULL(refined_T_tsdf_sensor);
    vbx_ctx* ctx = gpu_detail::lookupContext(&tsdf_layer);
    vbx_icp_config pod;
    std::memset(&pod, 0, sizeof(pod));
    pod.refine_roll_pitch = config_.refine_roll_pitch ? 1 : 0;
    pod.mini_batch_size = config_.mini_batch_size;
    pod.min_match_ratio = config_.min_match_ratio;
    pod.subsample_keep_ratio = config_.subsample_keep_ratio;
    pod.inital_translation_weighting = config_.inital_translation_weighting;
    pod.inital_rotation_weighting = config_.inital_rotation_weighting;
    pod.num_threads = static_cast<int32_t>(std::max<size_t>(1, std::min<size_t>(config_.num_threads, 32)));
    const float q[4] = {inital_T_tsdf_sensor.getRotation().w(), inital_T_tsdf_sensor.getRotation().x(),
                        inital_T_tsdf_sensor.getRotation().y(), inital_T_tsdf_sensor.getRotation().z()};
    const Point p = inital_T_tsdf_sensor.getPosition();
    const float t[3] = {p.x(), p.y(), p.z()};
    float oq[4], ot[3];
    uint64_t num_updates = 0;
    gpu_detail::check(ctx,
                      vbx_icp_run(ctx, &pod, points.empty() ? nullptr : reinterpret_cast<const float*>(points.data()), points.size(), q, t,
                                  static_cast<uint32_t>(seed), oq, ot, &num_updates),
                      "vbx_icp_run");
    *refined_T_tsdf_sensor = Transformation(Rotation(oq[0], oq[1], oq[2], oq[3]), Point(ot[0], ot[1], ot[2]));
    return static_cast<size_t>(num_updates);
  }

  bool refiningRollPitch() { return config_.refine_roll_pitch; }

 private:
  ICP::Config config_;
};

}  // namespace voxblox

#endif  // VOXBLOX_B200_GPU_INTEGRATORS_H_
