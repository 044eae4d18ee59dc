// TEST INFRASTRUCTURE ONLY (oracle/): in the reference this header is generated
// by protoc from voxblox/proto/voxblox/Block.proto:4-16.  protoc is absent here;
// this plain struct offers the accessors core/block_inl.h:73-109 names so that
// the (never called on this path) serialisation templates still parse.
#ifndef VBX_ORACLE_SHIM_BLOCK_PB_H_
#define VBX_ORACLE_SHIM_BLOCK_PB_H_
#include <cstdint>
#include <vector>
#include <google/protobuf/message.h>
namespace voxblox {
class BlockProto : public google::protobuf::Message {
 public:
  int voxels_per_side() const { return voxels_per_side_; }
  float voxel_size() const { return voxel_size_; }
  float origin_x() const { return ox_; }
  float origin_y() const { return oy_; }
  float origin_z() const { return oz_; }
  bool has_data() const { return has_data_; }
  const std::vector<uint32_t>& voxel_data() const { return data_; }
  int voxel_data_size() const { return static_cast<int>(data_.size()); }
  void set_voxels_per_side(int v) { voxels_per_side_ = v; }
  void set_voxel_size(float v) { voxel_size_ = v; }
  void set_origin_x(float v) { ox_ = v; }
  void set_origin_y(float v) { oy_ = v; }
  void set_origin_z(float v) { oz_ = v; }
  void set_has_data(bool v) { has_data_ = v; }
  void add_voxel_data(uint32_t w) { data_.push_back(w); }
 private:
  int voxels_per_side_ = 0;
  float voxel_size_ = 0, ox_ = 0, oy_ = 0, oz_ = 0;
  bool has_data_ = false;
  std::vector<uint32_t> data_;
};
}  // namespace voxblox
#endif
