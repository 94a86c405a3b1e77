// Stand-in for voxblox/core/layer.h (TEST INFRASTRUCTURE): SURVEY.md A.1, the subset the integrators and the
// export wrapper use.
#pragma once
#include <memory>
#include <utility>
#include <voxblox/core/block.h>
#include <voxblox/core/block_hash.h>
#include <voxblox/core/common.h>

namespace voxblox {

template <typename VoxelType>
class Layer {
 public:
  typedef std::shared_ptr<Layer> Ptr;
  typedef Block<VoxelType> BlockType;
  typedef typename AnyIndexHashMapType<typename BlockType::Ptr>::type BlockHashMap;
  typedef typename std::pair<BlockIndex, typename BlockType::Ptr> BlockMapPair;

  explicit Layer(FloatingPoint voxel_size, size_t voxels_per_side) : voxel_size_(voxel_size), voxels_per_side_(voxels_per_side) {
    CHECK_GT(voxel_size_, 0.0f);
    voxel_size_inv_ = 1.0 / voxel_size_;
    block_size_ = voxel_size_ * voxels_per_side_;
    CHECK_GT(block_size_, 0.0f);
    block_size_inv_ = 1.0 / block_size_;
    CHECK_GT(voxels_per_side_, 0u);
    voxels_per_side_inv_ = 1.0f / static_cast<FloatingPoint>(voxels_per_side_);
  }
  virtual ~Layer() {}

  typename BlockType::Ptr getBlockPtrByIndex(const BlockIndex& index) {
    typename BlockHashMap::iterator it = block_map_.find(index);
    return it != block_map_.end() ? it->second : typename BlockType::Ptr();
  }
  typename BlockType::Ptr allocateBlockPtrByIndex(const BlockIndex& index) {
    typename BlockHashMap::iterator it = block_map_.find(index);
    return it != block_map_.end() ? it->second : allocateNewBlock(index);
  }
  typename BlockType::Ptr allocateNewBlock(const BlockIndex& index) {
    auto status = block_map_.emplace(
        index, std::make_shared<BlockType>(voxels_per_side_, voxel_size_, getOriginPointFromGridIndex(index, block_size_)));
    CHECK(status.second) << "Block already exists when allocating at " << index.transpose();
    return status.first->second;
  }
  void insertBlock(const std::pair<const BlockIndex, typename BlockType::Ptr>& block_pair) {
    auto status = block_map_.insert(block_pair);
    CHECK(status.second) << "Block already exists when inserting at " << status.first->first.transpose();
  }
  void removeAllBlocks() { block_map_.clear(); }
  bool hasBlock(const BlockIndex& index) const { return block_map_.count(index) > 0; }
  void getAllAllocatedBlocks(BlockIndexList* blocks) const {
    blocks->clear();
    blocks->reserve(block_map_.size());
    for (const auto& kv : block_map_) blocks->emplace_back(kv.first);
  }
  size_t getNumberOfAllocatedBlocks() const { return block_map_.size(); }
  BlockIndex computeBlockIndexFromCoordinates(const Point& coords) const { return getGridIndexFromPoint<BlockIndex>(coords, block_size_inv_); }
  const BlockType& getBlockByIndex(const BlockIndex& index) const { return *block_map_.at(index); }

  FloatingPoint voxel_size() const { return voxel_size_; }
  FloatingPoint voxel_size_inv() const { return voxel_size_inv_; }
  FloatingPoint block_size() const { return block_size_; }
  FloatingPoint block_size_inv() const { return block_size_inv_; }
  size_t voxels_per_side() const { return voxels_per_side_; }
  FloatingPoint voxels_per_side_inv() const { return voxels_per_side_inv_; }

 private:
  FloatingPoint voxel_size_, block_size_, voxel_size_inv_, block_size_inv_, voxels_per_side_inv_;
  size_t voxels_per_side_;
  BlockHashMap block_map_;
};

}  // namespace voxblox
