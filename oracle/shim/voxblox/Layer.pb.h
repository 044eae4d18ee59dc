// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for the protoc output of
// voxblox/proto/voxblox/Layer.proto:4-11 (accessors used by core/layer_inl.h:25-52).
#ifndef VBX_ORACLE_SHIM_LAYER_PB_H_
#define VBX_ORACLE_SHIM_LAYER_PB_H_
#include <cstdint>
#include <string>
#include <google/protobuf/message.h>
namespace voxblox {
class LayerProto : public google::protobuf::Message {
 public:
  double voxel_size() const { return voxel_size_; }
  uint32_t voxels_per_side() const { return voxels_per_side_; }
  const std::string& type() const { return type_; }
  void set_voxel_size(double v) { voxel_size_ = v; }
  void set_voxels_per_side(uint32_t v) { voxels_per_side_ = v; }
  void set_type(const std::string& t) { type_ = t; }
 private:
  double voxel_size_ = 0;
  uint32_t voxels_per_side_ = 0;
  std::string type_;
};
}  // namespace voxblox
#endif
