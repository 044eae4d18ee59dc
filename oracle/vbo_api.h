/* TEST INFRASTRUCTURE ONLY -- the C interface shared by the two CPU checkers:
 *   oracle/_ref/libvbx_ref.so   the reference's own hot-path translation units
 *                               (compiled verbatim from /root/reference against
 *                               oracle/shim/), wrapped by oracle/ref_harness.cc
 *   oracle/libvbx_oracle.so     the from-scratch restatement, oracle/vbx_oracle.cc
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load either library.  The product (voxblox_b200/) never does.
 */
#ifndef VBO_API_H_
#define VBO_API_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* POD mirror of TsdfIntegratorBase::Config
 * (voxblox/include/voxblox/integrator/tsdf_integrator.h:56-89). */
typedef struct vbo_tsdf_config {
  float default_truncation_distance;
  float max_weight;
  int32_t voxel_carving_enabled;
  float min_ray_length_m;
  float max_ray_length_m;
  int32_t use_const_weight;
  int32_t allow_clear;
  int32_t use_weight_dropoff;
  int32_t use_sparsity_compensation_factor;
  float sparsity_compensation_factor;
  int32_t integrator_threads;
  int32_t integration_order_mode; /* 0 "mixed", 1 "sorted" */
  int32_t enable_anti_grazing;
  float start_voxel_subsampling_factor;
  int32_t max_consecutive_ray_collisions;
  int32_t clear_checks_every_n_frames;
  float max_integration_time_s;
} vbo_tsdf_config;

/* POD mirror of EsdfIntegrator::Config
 * (voxblox/include/voxblox/integrator/esdf_integrator.h:29-78). */
typedef struct vbo_esdf_config {
  int32_t full_euclidean_distance;
  float max_distance_m;
  float min_distance_m;
  float default_distance_m;
  float min_diff_m;
  float min_weight;
  int32_t num_buckets;
  int32_t multi_queue;
  int32_t add_occupied_crust;
  float clear_sphere_radius;
  float occupied_sphere_radius;
} vbo_esdf_config;

/* bundle_order for the Merged integrator: only the reference's own order exists --
 * libstdc++ unordered_map iteration order, as integrateVoxels walks it with one thread
 * (tsdf_integrator.cc:434-457).  (Round 1 also had a "canonical" order; it is gone: the device
 * reproduces the reference order.) */
enum { VBO_ORDER_REFERENCE = 0 };
enum { VBO_SIMPLE = 1, VBO_MERGED = 2, VBO_FAST = 3 }; /* TsdfIntegratorType */
enum { VBO_LAYER_TSDF = 0, VBO_LAYER_ESDF = 1 };

const char* vbo_impl_name(void); /* "reference" or "port" */

void* vbo_create(const vbo_tsdf_config* cfg, float voxel_size, int voxels_per_side);
void vbo_destroy(void* h);

/* TsdfIntegratorBase::integratePointCloud (tsdf_integrator.h:100-103).
 * q_wxyz + t = T_G_C; xyz = 3n floats (points_C); rgba = 4n bytes. */
int vbo_integrate(void* h, int kind, const float q_wxyz[4], const float t[3],
                  const float* xyz, const uint8_t* rgba, uint64_t n, int freespace,
                  int bundle_order);
/* seconds spent inside the last integratePointCloud / ESDF update call */
double vbo_last_seconds(void* h);
/* counters of the last integrate call (restatement only; zeros from _ref):
 * [0] normal rays/bundles cast, [1] clearing rays/bundles cast, [2] K voxel
 * updates attempted, [3] U distinct voxels touched, [4] B distinct blocks touched,
 * [5] blocks newly allocated, [6] valid points, [7] reserved */
void vbo_last_counters(void* h, uint64_t out[8]);

uint64_t vbo_num_blocks(void* h, int layer);
/* 3*M int32, sorted ascending by (x, y, z) */
void vbo_block_indices(void* h, int layer, int32_t* out);
/* voxels: vps^3 * 12 B (TsdfVoxel: f32 distance, f32 weight, u8 rgba;
 * core/voxel.h:12-16) or * 20 B (EsdfVoxel: f32 distance, u8 observed,
 * hallucinated, in_queue, fixed, i32 parent[3]; core/voxel.h:18-37), linear index
 * x + vps*(y + vps*z).  updated_bits: bit0 kMap, bit1 kMesh, bit2 kEsdf
 * (core/block.h:15-18).  Returns 0, or 1 if the block does not exist. */
int vbo_get_block(void* h, int layer, const int32_t idx[3], void* voxels,
                  uint8_t* updated_bits);

int vbo_esdf_create(void* h, const vbo_esdf_config* cfg);
/* batch=0: updateFromTsdfLayer(clear_updated_flag) (esdf_integrator.cc:104-122)
 * batch=1: updateFromTsdfLayerBatch()              (esdf_integrator.cc:94-102) */
int vbo_esdf_update(void* h, int batch, int clear_updated_flag);
/* Block::serializeToIntegers / deserializeFromIntegers (src/core/block.cc:65-90,110-135,159-183,
 * 203-234): words holds vps^3 * 3 (TSDF) or * 2 (ESDF) uint32.  The deserialising call creates
 * the block if needed. */
int vbo_serialize_block(void* h, int layer, const int32_t idx[3], uint32_t* words);
int vbo_deserialize_block(void* h, int layer, const int32_t idx[3], const uint32_t* words);
/* EsdfIntegrator::updateFromTsdfBlocks(tsdf_blocks, incremental) (esdf_integrator.cc:124-302) */
int vbo_esdf_update_blocks(void* h, const int32_t* idx3, uint64_t m, int incremental);
/* setEsdfMaxDistance / setFullEuclidean (esdf_integrator.h:139-149) */
int vbo_esdf_set_max_distance(void* h, float max_distance);
int vbo_esdf_set_full_euclidean(void* h, int full_euclidean);
/* EsdfIntegrator::addNewRobotPosition (esdf_integrator.cc:25-92) */
int vbo_esdf_add_robot_position(void* h, const float p[3]);


/* MeshIntegrator<TsdfVoxel>::generateMesh(only_mesh_updated_blocks, clear_updated_flag)
 * (mesh/mesh_integrator.h:132-160) into the map's own MeshLayer (mesh/mesh_layer.h); marching
 * cubes per mesh/marching_cubes.h:44-164, vertex colours per mesh_integrator.h:368-388. */
int vbo_mesh_generate(void* h, int use_color, float min_weight, int only_mesh_updated_blocks,
                      int clear_updated_flag);
uint64_t vbo_mesh_num_blocks(void* h);           /* MeshLayer::getAllAllocatedMeshes */
void vbo_mesh_block_indices(void* h, int32_t* out); /* 3*M int32, sorted by (x, y, z) */
/* One block mesh: returns its vertex count (UINT64_MAX if there is no mesh at idx).  vertices /
 * normals: 3 floats per vertex, colors: 4 bytes per vertex (written only when the mesh has
 * colours); any pointer may be null.  Mesh::indices is always 0..n-1 (marching_cubes.h:97-99). */
uint64_t vbo_mesh_get(void* h, const int32_t idx[3], float* vertices, float* normals, uint8_t* colors,
                      int* has_colors, int* updated);
/* MarchingCubes::kTriangleTable / kEdgeIndexPairs (src/mesh/marching_cubes.cc:33-293);
 * returns 0 from the reference library, 1 from the restatement (which holds no second copy). */
int vbo_mc_tables(int32_t tri[256 * 16], int32_t edges[12 * 2]);

/* POD mirror of ICP::Config (voxblox/include/voxblox/alignment/icp.h:76-108; field names as spelled there). */
typedef struct vbo_icp_config {
  int32_t refine_roll_pitch;
  int32_t mini_batch_size;
  float min_match_ratio;
  float subsample_keep_ratio;
  float inital_translation_weighting;
  float inital_rotation_weighting;
  int32_t num_threads;
} vbo_icp_config;
/* ICP::runICP(tsdf_layer, points, inital_T_tsdf_sensor, &refined_T_tsdf_sensor, seed) (icp.h:118-123,
 * icp.cc:219-259) against the map's TSDF layer; returns 0 and the number of successful mini batches.
 * The reference library runs the reference's own class: deterministic only with num_threads = 1 (its
 * threads race for batches and for the pose).  The restatement also accepts num_threads = T > 1 and
 * then follows ONE legal schedule of those T threads, the round-robin one: in every round thread
 * w = 0..T-1 takes the next batch (atomic_idx_ order), matches it against the pose snapshot it took at
 * its own last successful fusion, and the fusions of the round are applied in the order w = 0..T-1. */
int vbo_icp_run(void* h, const vbo_icp_config* cfg, const float* xyz, uint64_t n, const float q_wxyz[4],
                const float t[3], uint32_t seed, float out_q_wxyz[4], float out_t[3], uint64_t* num_updates);
/* std::shuffle(first, last, std::default_random_engine(seed)) of n elements (icp.cc:229-233):
 * out[i] = original position of the element that ends at position i. */
void vbo_icp_shuffle(uint64_t n, uint32_t seed, uint32_t* out);

/* Iteration order of a std::unordered_map<key, ...> with hash(key) = hashes[i] (as size_t) after
 * inserting keys 0..n-1 one by one with operator[], exactly as bundleRays fills voxel_map
 * (tsdf_integrator.cc:340-371): out[p] = the key at iteration position p.  Checks the device's
 * k_bundle_order against the C++ library itself. */
void vbo_umap_order(const uint32_t* hashes, uint64_t n, uint32_t* out);

#ifdef __cplusplus
}
#endif
#endif /* VBO_API_H_ */
