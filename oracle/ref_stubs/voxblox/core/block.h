// Stand-in for voxblox/core/block.h (TEST INFRASTRUCTURE): SURVEY.md A.1.
#pragma once
#include <memory>
#include <voxblox/core/common.h>

namespace voxblox {

template <typename VoxelType>
class Block {
 public:
  typedef std::shared_ptr<Block<VoxelType>> Ptr;
  typedef std::shared_ptr<const Block<VoxelType>> ConstPtr;

  Block(size_t voxels_per_side, FloatingPoint voxel_size, const Point& origin)
      : voxels_per_side_(voxels_per_side), voxel_size_(voxel_size), origin_(origin), has_data_(false), updated_(false) {
    num_voxels_ = voxels_per_side_ * voxels_per_side_ * voxels_per_side_;
    voxel_size_inv_ = 1.0 / voxel_size_;
    block_size_ = voxels_per_side_ * voxel_size_;
    block_size_inv_ = 1.0 / block_size_;
    voxels_.reset(new VoxelType[num_voxels_]);
  }

  size_t computeLinearIndexFromVoxelIndex(const VoxelIndex& index) const {
    return static_cast<size_t>(index.x() + voxels_per_side_ * (index.y() + index.z() * voxels_per_side_));
  }
  VoxelIndex computeVoxelIndexFromLinearIndex(size_t linear_index) const {
    int rem = (int)linear_index;
    const int vps = (int)voxels_per_side_;
    const int z = rem / (vps * vps);
    rem -= z * vps * vps;
    const int y = rem / vps;
    return VoxelIndex(rem - y * vps, y, z);
  }
  Point computeCoordinatesFromVoxelIndex(const VoxelIndex& index) const {
    return origin_ + getCenterPointFromGridIndex(index, voxel_size_);
  }
  Point computeCoordinatesFromLinearIndex(size_t linear_index) const {
    return computeCoordinatesFromVoxelIndex(computeVoxelIndexFromLinearIndex(linear_index));
  }
  VoxelType& getVoxelByLinearIndex(size_t index) { return voxels_[index]; }
  const VoxelType& getVoxelByLinearIndex(size_t index) const { return voxels_[index]; }
  VoxelType& getVoxelByVoxelIndex(const VoxelIndex& index) { return voxels_[computeLinearIndexFromVoxelIndex(index)]; }
  const VoxelType& getVoxelByVoxelIndex(const VoxelIndex& index) const { return voxels_[computeLinearIndexFromVoxelIndex(index)]; }

  BlockIndex block_index() const { return getGridIndexFromPoint<BlockIndex>(origin_, block_size_inv_); }
  size_t voxels_per_side() const { return voxels_per_side_; }
  FloatingPoint voxel_size() const { return voxel_size_; }
  FloatingPoint block_size() const { return block_size_; }
  size_t num_voxels() const { return num_voxels_; }
  const Point& origin() const { return origin_; }
  bool has_data() const { return has_data_; }
  bool& has_data() { return has_data_; }
  bool updated() const { return updated_; }
  bool& updated() { return updated_; }

 private:
  std::unique_ptr<VoxelType[]> voxels_;
  size_t num_voxels_;
  const size_t voxels_per_side_;
  const FloatingPoint voxel_size_;
  Point origin_;
  FloatingPoint voxel_size_inv_, block_size_, block_size_inv_;
  bool has_data_, updated_;
};

}  // namespace voxblox
