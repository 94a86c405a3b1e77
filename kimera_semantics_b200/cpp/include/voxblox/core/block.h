// voxblox::Block<VoxelType> subset
#pragma once
#include "voxblox/core/common.h"
namespace voxblox {
template <typename VoxelType>
class Block {
 public:
  typedef std::shared_ptr<Block<VoxelType>> Ptr;
  typedef std::shared_ptr<const Block<VoxelType>> ConstPtr;
  Block(size_t voxels_per_side, FloatingPoint voxel_size, const Point& origin)
      : has_data_(false), voxels_per_side_(voxels_per_side), voxel_size_(voxel_size), origin_(origin), updated_(false) {
    num_voxels_ = voxels_per_side_ * voxels_per_side_ * voxels_per_side_;
    voxel_size_inv_ = 1.0 / voxel_size_;
    block_size_ = voxels_per_side_ * voxel_size_;
    block_size_inv_ = 1.0 / block_size_;
    voxels_.reset(new VoxelType[num_voxels_]);
  }
  size_t computeLinearIndexFromVoxelIndex(const VoxelIndex& i) const {
    return static_cast<size_t>(i.x() + voxels_per_side_ * (i.y() + i.z() * voxels_per_side_));
  }
  VoxelIndex computeVoxelIndexFromLinearIndex(size_t lin) const {
    int rem = (int)lin;
    VoxelIndex r;
    r[2] = rem / (int)(voxels_per_side_ * voxels_per_side_);
    rem -= r[2] * (int)(voxels_per_side_ * voxels_per_side_);
    r[1] = rem / (int)voxels_per_side_;
    r[0] = rem - r[1] * (int)voxels_per_side_;
    return r;
  }
  Point computeCoordinatesFromLinearIndex(size_t lin) const {
    const VoxelIndex v = computeVoxelIndexFromLinearIndex(lin);
    return Point(origin_.x() + (v.x() + 0.5f) * voxel_size_, origin_.y() + (v.y() + 0.5f) * voxel_size_, origin_.z() + (v.z() + 0.5f) * voxel_size_);
  }
  const VoxelType& getVoxelByLinearIndex(size_t i) const { return voxels_[i]; }
  VoxelType& getVoxelByLinearIndex(size_t i) { return voxels_[i]; }
  const VoxelType& getVoxelByVoxelIndex(const VoxelIndex& i) const { return voxels_[computeLinearIndexFromVoxelIndex(i)]; }
  VoxelType& getVoxelByVoxelIndex(const VoxelIndex& i) { return voxels_[computeLinearIndexFromVoxelIndex(i)]; }
  size_t voxels_per_side() const { return voxels_per_side_; }
  FloatingPoint voxel_size() const { return voxel_size_; }
  FloatingPoint voxel_size_inv() const { return voxel_size_inv_; }
  size_t num_voxels() const { return num_voxels_; }
  const Point& origin() const { return origin_; }
  FloatingPoint block_size() const { return block_size_; }
  BlockIndex block_index() const { return getGridIndexFromPoint(origin_, block_size_inv_); }
  bool has_data() const { return has_data_; }
  bool& has_data() { return has_data_; }
  bool updated() const { return updated_; }
  bool& updated() { return updated_; }
  VoxelType* voxel_data() { return voxels_.get(); }
 private:
  std::unique_ptr<VoxelType[]> voxels_;
  size_t num_voxels_;
  bool has_data_;
  const size_t voxels_per_side_;
  const FloatingPoint voxel_size_;
  const Point origin_;
  FloatingPoint voxel_size_inv_, block_size_, block_size_inv_;
  bool updated_;
};
}  // namespace voxblox
