// TEST PROGRAM (oracle/, built only where /root/reference exists): drives the voxblox-side
// adapter include/voxblox_b200/gpu_integrators.h through the reference's OWN headers and
// classes, next to the reference's CPU integrators, the way test/test_sdf_integrators.cc does.
// Simple must agree voxel for voxel (same update order); Merged is compared on the block set
// and statistically (its CPU order is unordered_map iteration order).  Exit code 0 = pass.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

#include "voxblox_b200/gpu_integrators.h"

using namespace voxblox;  // NOLINT

static void makeScan(int k, Transformation* T, Pointcloud* pts, Colors* cols) {
  // a wall at z = 3 m and a floor, seen by a 96x72 pinhole camera from slightly different poses
  const float yaw = 0.05f * k, pitch = -0.03f * k;
  const float cy = std::cos(yaw / 2), sy = std::sin(yaw / 2), cp = std::cos(pitch / 2), sp = std::sin(pitch / 2);
  const Rotation rot(cy * cp, -sy * sp, cy * sp, sy * cp);  // not exactly unit: kept as is on both paths
  *T = Transformation(rot, Point(0.013f + 0.1f * k, 0.021f, 0.017f));
  pts->clear();
  cols->clear();
  for (int v = 0; v < 72; ++v) {
    for (int u = 0; u < 96; ++u) {
      const float dx = (u - 47.6f) / 80.0f, dy = (v - 35.7f) / 80.0f;
      float depth = 3.0f + 0.2f * std::sin(0.3f * u) * std::cos(0.2f * v);
      if (dy > 0.25f) depth = 1.2f / dy * 0.4f;  // "floor"
      if (depth > 4.5f || depth < 0.5f) continue;
      pts->push_back(Point(dx * depth, dy * depth, depth));
      cols->push_back(Color(static_cast<uint8_t>(u * 2), static_cast<uint8_t>(v * 3), 128, 255));
    }
  }
}

template <typename V>
static std::vector<BlockIndex> sortedBlocks(const Layer<V>& l) {
  BlockIndexList b;
  l.getAllAllocatedBlocks(&b);
  std::vector<BlockIndex> v(b.begin(), b.end());
  std::sort(v.begin(), v.end(), [](const BlockIndex& a, const BlockIndex& c) {
    if (a.x() != c.x()) return a.x() < c.x();
    if (a.y() != c.y()) return a.y() < c.y();
    return a.z() < c.z();
  });
  return v;
}

int main() {
  const float voxel_size = 0.1f;
  TsdfIntegratorBase::Config config;
  config.default_truncation_distance = 0.4f;
  config.integrator_threads = 1;
  int failures = 0;
  for (int type = 1; type <= 2; ++type) {
    Layer<TsdfVoxel> cpu_layer(voxel_size, 16), gpu_layer(voxel_size, 16);
    TsdfIntegratorBase::Ptr cpu = TsdfIntegratorFactory::create(static_cast<TsdfIntegratorType>(type), config, &cpu_layer);
    // the drop-in: same base-class pointer type, same call
    std::shared_ptr<GpuTsdfIntegrator> gpu_impl(
        new GpuTsdfIntegrator(static_cast<TsdfIntegratorType>(type), config, &gpu_layer));
    TsdfIntegratorBase::Ptr gpu = gpu_impl;
    for (int k = 0; k < 4; ++k) {
      Transformation T;
      Pointcloud pts;
      Colors cols;
      makeScan(k, &T, &pts, &cols);
      cpu->integratePointCloud(T, pts, cols);
      gpu->integratePointCloud(T, pts, cols);
    }
    gpu_impl->syncLayer(0);
    const std::vector<BlockIndex> a = sortedBlocks(cpu_layer), b = sortedBlocks(gpu_layer);
    bool same_blocks = a.size() == b.size();
    for (size_t i = 0; same_blocks && i < a.size(); ++i) same_blocks = a[i] == b[i];
    size_t n = 0, exact = 0, observed = 0;
    double se = 0;
    if (same_blocks) {
      for (const BlockIndex& bi : a) {
        const Block<TsdfVoxel>& x = cpu_layer.getBlockByIndex(bi);
        const Block<TsdfVoxel>& y = gpu_layer.getBlockByIndex(bi);
        for (size_t l = 0; l < x.num_voxels(); ++l) {
          const TsdfVoxel& p = x.getVoxelByLinearIndex(l);
          const TsdfVoxel& q = y.getVoxelByLinearIndex(l);
          ++n;
          if (p.distance == q.distance && p.weight == q.weight && p.color.r == q.color.r && p.color.g == q.color.g &&
              p.color.b == q.color.b && p.color.a == q.color.a) {
            ++exact;
          }
          if (p.weight > 0 || q.weight > 0) {
            ++observed;
            se += (p.distance - q.distance) * (p.distance - q.distance);
          }
        }
      }
    }
    const double rmse = observed ? std::sqrt(se / observed) : 0.0;
    std::printf("type %d (%s): blocks cpu %zu gpu %zu same %d, voxels %zu bit-exact %zu, rmse %.3g\n", type,
                type == 1 ? "simple" : "merged", a.size(), b.size(), same_blocks ? 1 : 0, n, exact, rmse);
    if (!same_blocks) ++failures;
    // Simple: point order.  Merged: the device reproduces the iteration order of the reference's
    // unordered_map (one integrator thread).  Both: bit-identical.
    if (exact != n) ++failures;
  }
  // Pipelined adapter, and an adapter started from a non-empty layer: both must reproduce the
  // plain adapter bit for bit.
  {
    auto digest = [](Layer<TsdfVoxel>& layer) {
      uint64_t h = 1469598103934665603ull;
      for (const BlockIndex& bi : sortedBlocks(layer)) {
        const Block<TsdfVoxel>& blk = layer.getBlockByIndex(bi);
        const unsigned char* p = reinterpret_cast<const unsigned char*>(&blk.getVoxelByLinearIndex(0));
        for (size_t i = 0; i < blk.num_voxels() * sizeof(TsdfVoxel); ++i) h = (h ^ p[i]) * 1099511628211ull;
        h = (h ^ static_cast<uint64_t>(bi.x() * 73856093 ^ bi.y() * 19349669 ^ bi.z() * 83492791)) * 1099511628211ull;
      }
      return h;
    };
    Layer<TsdfVoxel> plain_layer(voxel_size, 16), piped_layer(voxel_size, 16), resumed_layer(voxel_size, 16);
    GpuTsdfIntegrator plain(TsdfIntegratorType::kMerged, config, &plain_layer);
    GpuTsdfIntegrator piped(TsdfIntegratorType::kMerged, config, &piped_layer);
    piped.setPipelined(true);
    Transformation T;
    Pointcloud pts;
    Colors cols;
    for (int k = 0; k < 6; ++k) {
      makeScan(k, &T, &pts, &cols);
      plain.integratePointCloud(T, pts, cols);
      piped.integratePointCloud(T, pts, cols);
      pts.assign(pts.size(), Point(1e9f, 1e9f, 1e9f));  // the caller may reuse its buffers at once
      if (k == 2) {
        // hand the map over to a fresh adapter through the host layer
        plain.syncLayer(0);
        for (const BlockIndex& bi : sortedBlocks(plain_layer)) {
          Block<TsdfVoxel>::Ptr dst = resumed_layer.allocateBlockPtrByIndex(bi);
          const Block<TsdfVoxel>& src = plain_layer.getBlockByIndex(bi);
          std::memcpy(&dst->getVoxelByLinearIndex(0), &src.getVoxelByLinearIndex(0), src.num_voxels() * sizeof(TsdfVoxel));
        }
      }
    }
    GpuTsdfIntegrator resumed(TsdfIntegratorType::kMerged, config, &resumed_layer);
    for (int k = 3; k < 6; ++k) {
      makeScan(k, &T, &pts, &cols);
      resumed.integratePointCloud(T, pts, cols);
    }
    plain.syncLayer(0);
    piped.syncLayer(0);
    resumed.syncLayer(0);
    const uint64_t d0 = digest(plain_layer), d1 = digest(piped_layer), d2 = digest(resumed_layer);
    std::printf("pipelined adapter identical %d, resumed-from-host-layer identical %d (%zu blocks)\n", d0 == d1 ? 1 : 0,
                d0 == d2 ? 1 : 0, plain_layer.getNumberOfAllocatedBlocks());
    if (d0 != d1 || d0 != d2) ++failures;
  }
  // ESDF through the adapter: incremental update after a scan, then batch
  {
    Layer<TsdfVoxel> tsdf(voxel_size, 16);
    Layer<EsdfVoxel> esdf(voxel_size, 16);
    GpuTsdfIntegrator gpu(TsdfIntegratorType::kMerged, config, &tsdf);
    EsdfIntegrator::Config ec;
    ec.min_distance_m = 0.2f;
    GpuEsdfIntegrator ge(ec, &gpu, &esdf);
    Transformation T;
    Pointcloud pts;
    Colors cols;
    makeScan(0, &T, &pts, &cols);
    gpu.integratePointCloud(T, pts, cols);
    ge.updateFromTsdfLayer(true);
    const size_t nb = ge.syncLayer(0);
    std::printf("esdf blocks %zu (tsdf %zu)\n", nb, gpu.syncLayer(0));
    if (nb == 0 || esdf.getNumberOfAllocatedBlocks() != tsdf.getNumberOfAllocatedBlocks()) ++failures;
  }
  // The drop-in shapes: GpuEsdfIntegrator built from the REFERENCE's constructor arguments (layer pointers,
  // esdf_integrator.h:80-82 / esdf_server.cc) with auto-sync, the TSDF adapter in auto-sync mode feeding the
  // reference's OWN host-side MeshIntegrator without any syncLayer() call, and setLayer().
  {
    Layer<TsdfVoxel> tsdf(voxel_size, 16), tsdf_plain(voxel_size, 16);
    Layer<EsdfVoxel> esdf(voxel_size, 16);
    GpuTsdfIntegrator gpu(TsdfIntegratorType::kMerged, config, &tsdf);
    GpuTsdfIntegrator plain(TsdfIntegratorType::kMerged, config, &tsdf_plain);
    gpu.setAutoSync(true);  // every block an integratePointCloud call changes is mirrored into `tsdf`
    EsdfIntegrator::Config ec;
    ec.min_distance_m = 0.2f;
    GpuEsdfIntegrator ge(ec, &tsdf, &esdf);  // <- the reference's signature; finds the engine through the layer
    MeshLayer mesh_host(tsdf.block_size()), mesh_ref(tsdf_plain.block_size());
    MeshIntegratorConfig mc;
    mc.integrator_threads = 1;
    MeshIntegrator<TsdfVoxel> host_mesher(mc, &tsdf, &mesh_host);       // the reference's mesher on the mirrored layer
    MeshIntegrator<TsdfVoxel> ref_mesher(mc, &tsdf_plain, &mesh_ref);  // ... and on an explicitly synced layer
    Transformation T;
    Pointcloud pts;
    Colors cols;
    size_t esdf_blocks = 0;
    for (int k = 0; k < 3; ++k) {
      makeScan(k, &T, &pts, &cols);
      gpu.integratePointCloud(T, pts, cols);
      plain.integratePointCloud(T, pts, cols);
      ge.updateFromTsdfLayer(true);  // auto-sync: the ESDF blocks it touched are in `esdf` when it returns
      host_mesher.generateMesh(true, true);
      plain.syncLayer(0, 0);
      ref_mesher.generateMesh(false, true);
      esdf_blocks = esdf.getNumberOfAllocatedBlocks();
    }
    BlockIndexList ma, mb;
    mesh_host.getAllAllocatedMeshes(&ma);
    mesh_ref.getAllAllocatedMeshes(&mb);
    size_t va = 0, vb = 0;
    for (const BlockIndex& bi : ma) va += mesh_host.getMeshByIndex(bi).vertices.size();
    for (const BlockIndex& bi : mb) vb += mesh_ref.getMeshByIndex(bi).vertices.size();
    // setLayer: the device map follows the new host layer
    Layer<TsdfVoxel> other(voxel_size, 16);
    gpu.setLayer(&other);
    makeScan(0, &T, &pts, &cols);
    gpu.integratePointCloud(T, pts, cols);
    Layer<TsdfVoxel> fresh(voxel_size, 16);
    GpuTsdfIntegrator fresh_gpu(TsdfIntegratorType::kMerged, config, &fresh);
    fresh_gpu.integratePointCloud(T, pts, cols);
    fresh_gpu.syncLayer(0);
    const std::vector<BlockIndex> oa = sortedBlocks(other), ob = sortedBlocks(fresh);
    bool set_layer_ok = oa.size() == ob.size() && !oa.empty();
    for (size_t i = 0; set_layer_ok && i < oa.size(); ++i) {
      set_layer_ok = oa[i] == ob[i] &&
                     std::memcmp(&other.getBlockByIndex(oa[i]).getVoxelByLinearIndex(0),
                                 &fresh.getBlockByIndex(ob[i]).getVoxelByLinearIndex(0), 4096 * sizeof(TsdfVoxel)) == 0;
    }
    std::printf("drop-in shapes: esdf blocks %zu tsdf blocks %zu, host mesher on the auto-synced layer %zu vertices "
                "(explicit sync %zu), setLayer ok %d\n",
                esdf_blocks, tsdf.getNumberOfAllocatedBlocks(), va, vb, set_layer_ok ? 1 : 0);
    if (esdf_blocks == 0 || esdf_blocks != tsdf.getNumberOfAllocatedBlocks() || va == 0 || va != vb || !set_layer_ok) ++failures;
  }
  // addNewRobotPosition through the adapter against the reference's own EsdfIntegrator on an empty
  // map (test_clear_spheres.cc:135-136): the hallucinated free sphere / occupied shell is deterministic,
  // so both ESDF layers must agree voxel for voxel.
  {
    Layer<TsdfVoxel> tsdf_c(voxel_size, 16), tsdf_g(voxel_size, 16);
    Layer<EsdfVoxel> esdf_c(voxel_size, 16), esdf_g(voxel_size, 16);
    EsdfIntegrator::Config ec;
    ec.min_distance_m = 0.2f;
    ec.clear_sphere_radius = 1.0f;
    ec.occupied_sphere_radius = 2.5f;
    EsdfIntegrator ce(ec, &tsdf_c, &esdf_c);
    GpuTsdfIntegrator gpu(TsdfIntegratorType::kMerged, config, &tsdf_g);
    GpuEsdfIntegrator ge(ec, &gpu, &esdf_g);
    const Point pos(0.13f, -0.21f, 0.57f);
    ce.addNewRobotPosition(pos);
    ge.addNewRobotPosition(pos);
    ge.syncLayer(0);
    const std::vector<BlockIndex> a = sortedBlocks(esdf_c), b = sortedBlocks(esdf_g);
    bool same = a.size() == b.size();
    for (size_t i = 0; same && i < a.size(); ++i) same = a[i] == b[i];
    size_t diff = 0, free_v = 0, occ_v = 0;
    if (same) {
      for (const BlockIndex& bi : a) {
        const Block<EsdfVoxel>& x = esdf_c.getBlockByIndex(bi);
        const Block<EsdfVoxel>& y = esdf_g.getBlockByIndex(bi);
        for (size_t l = 0; l < x.num_voxels(); ++l) {
          const EsdfVoxel& p = x.getVoxelByLinearIndex(l);
          const EsdfVoxel& q = y.getVoxelByLinearIndex(l);
          if (p.distance != q.distance || p.observed != q.observed || p.hallucinated != q.hallucinated ||
              p.in_queue != q.in_queue || p.fixed != q.fixed || !(p.parent == q.parent)) {
            ++diff;
          }
          if (q.hallucinated) (q.distance > 0 ? free_v : occ_v) += 1;
        }
      }
    }
    std::printf("addNewRobotPosition: esdf blocks cpu %zu gpu %zu same %d, differing voxels %zu, free %zu occupied %zu, tsdf blocks %zu\n",
                a.size(), b.size(), same ? 1 : 0, diff, free_v, occ_v, gpu.syncLayer(0));
    if (!same || diff != 0 || free_v == 0 || occ_v == 0 || tsdf_g.getNumberOfAllocatedBlocks() != 0) ++failures;
  }
  // Meshing through the adapter against the reference's own MeshIntegrator on the SAME host layer
  // (the Simple integrator's device map is bit-identical to the CPU one, so every vertex must be too).
  {
    Layer<TsdfVoxel> cpu_layer(voxel_size, 16), gpu_layer(voxel_size, 16);
    TsdfIntegratorBase::Ptr cpu = TsdfIntegratorFactory::create("simple", config, &cpu_layer);
    GpuTsdfIntegrator gpu(TsdfIntegratorType::kSimple, config, &gpu_layer);
    MeshLayer mesh_c(cpu_layer.block_size()), mesh_g(gpu_layer.block_size());
    MeshIntegratorConfig mc;
    mc.integrator_threads = 1;
    MeshIntegrator<TsdfVoxel> cm(mc, &cpu_layer, &mesh_c);
    GpuMeshIntegrator gm(mc, &gpu, &mesh_g);
    size_t vertices = 0, diff = 0, blocks_differ = 0;
    for (int k = 0; k < 3; ++k) {
      Transformation T;
      Pointcloud pts;
      Colors cols;
      makeScan(k, &T, &pts, &cols);
      cpu->integratePointCloud(T, pts, cols);
      gpu.integratePointCloud(T, pts, cols);
      cm.generateMesh(true, true);
      gm.generateMesh(true, true);
    }
    BlockIndexList a, b;
    mesh_c.getAllAllocatedMeshes(&a);
    mesh_g.getAllAllocatedMeshes(&b);
    if (a.size() != b.size()) ++blocks_differ;
    for (const BlockIndex& bi : a) {
      Mesh::ConstPtr x = static_cast<const MeshLayer&>(mesh_c).getMeshPtrByIndex(bi);
      Mesh::ConstPtr y = std::find(b.begin(), b.end(), bi) != b.end() ? static_cast<const MeshLayer&>(mesh_g).getMeshPtrByIndex(bi)
                                                                       : Mesh::ConstPtr();
      if (!y || x->vertices.size() != y->vertices.size() || x->colors.size() != y->colors.size() ||
          x->indices.size() != y->indices.size()) {
        ++blocks_differ;
        continue;
      }
      vertices += x->vertices.size();
      for (size_t i = 0; i < x->vertices.size(); ++i) {
        const bool same = x->vertices[i] == y->vertices[i] && x->normals[i] == y->normals[i] &&
                          x->indices[i] == y->indices[i] && x->colors[i].r == y->colors[i].r &&
                          x->colors[i].g == y->colors[i].g && x->colors[i].b == y->colors[i].b &&
                          x->colors[i].a == y->colors[i].a;
        if (!same) ++diff;
      }
    }
    std::printf("mesh: blocks cpu %zu gpu %zu, block mismatches %zu, vertices %zu, differing %zu\n", a.size(), b.size(),
                blocks_differ, vertices, diff);
    if (blocks_differ || diff || vertices == 0) ++failures;
  }
  // ICP through the adapter: the reference's own voxblox::ICP on the host layer next to GpuICP
  // on the device map, same constructor / runICP signature, num_threads = 1 (the reference's deterministic
  // setting), same seed.  The maps are bit identical (Merged); the refined poses agree to the float tolerance
  // (a chain of SE(3) log / exp with library functions, tests/test_icp_gpu.py), the fused-batch counts exactly.
  {
    Layer<TsdfVoxel> tsdf_c(voxel_size, 16), tsdf_g(voxel_size, 16);
    TsdfIntegratorBase::Ptr cpu = TsdfIntegratorFactory::create(TsdfIntegratorType::kMerged, config, &tsdf_c);
    GpuTsdfIntegrator gpu(TsdfIntegratorType::kMerged, config, &tsdf_g);
    Transformation T;
    Pointcloud pts;
    Colors cols;
    for (int k = 0; k < 3; ++k) {
      makeScan(k, &T, &pts, &cols);
      cpu->integratePointCloud(T, pts, cols);
      gpu.integratePointCloud(T, pts, cols);
    }
    makeScan(1, &T, &pts, &cols);
    const Transformation guess(T.getRotation(), T.getPosition() + Point(0.03f, -0.02f, 0.01f));
    size_t bad = 0;
    for (int roll_pitch = 0; roll_pitch < 2; ++roll_pitch) {
      ICP::Config ic;
      ic.num_threads = 1;
      ic.refine_roll_pitch = roll_pitch != 0;
      ICP ref_icp(ic);
      GpuICP gpu_icp(ic);  // <- the reference's constructor
      Transformation Tc, Tg;
      const size_t nc = ref_icp.runICP(tsdf_c, pts, guess, &Tc, 7u);
      const size_t ng = gpu_icp.runICP(tsdf_g, pts, guess, &Tg, 7u);  // <- the reference's call
      const float dq = std::max(std::max(std::fabs(Tc.getRotation().w() - Tg.getRotation().w()),
                                         std::fabs(Tc.getRotation().x() - Tg.getRotation().x())),
                                std::max(std::fabs(Tc.getRotation().y() - Tg.getRotation().y()),
                                         std::fabs(Tc.getRotation().z() - Tg.getRotation().z())));
      const float dt = (Tc.getPosition() - Tg.getPosition()).norm();
      std::printf("icp (refine_roll_pitch %d): fused batches cpu %zu gpu %zu, |dq| %.3g, |dt| %.3g m, moved %.3g m\n", roll_pitch,
                  nc, ng, dq, dt, (Tg.getPosition() - guess.getPosition()).norm());
      if (nc != ng || nc == 0 || !(dq < 1e-4f) || !(dt < 1e-4f) || gpu_icp.refiningRollPitch() != ref_icp.refiningRollPitch()) ++bad;
    }
    if (bad) ++failures;
  }
  std::printf(failures ? "ADAPTER TEST FAILED\n" : "ADAPTER TEST OK\n");
  return failures ? 1 : 0;
}
