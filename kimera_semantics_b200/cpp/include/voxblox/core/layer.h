// voxblox::Layer<VoxelType> subset: the block-hash surface the integrator and its callers use.
#pragma once
#include <utility>
#include "voxblox/core/block.h"
namespace voxblox {
template <typename VoxelType>
class Layer {
 public:
  typedef std::shared_ptr<Layer> Ptr;
  typedef Block<VoxelType> BlockType;
  typedef typename AnyIndexHashMapType<typename BlockType::Ptr>::type BlockHashMap;
  typedef typename std::pair<BlockIndex, typename BlockType::Ptr> BlockMapPair;

  explicit Layer(FloatingPoint voxel_size, size_t voxels_per_side) : voxel_size_(voxel_size), voxels_per_side_(voxels_per_side) {
    KSG_CHECK(voxel_size > 0.0f);
    voxel_size_inv_ = 1.0 / voxel_size_;
    block_size_ = voxel_size_ * voxels_per_side_;
    block_size_inv_ = 1.0 / block_size_;
    voxels_per_side_inv_ = 1.0f / static_cast<FloatingPoint>(voxels_per_side_);
  }
  typename BlockType::Ptr getBlockPtrByIndex(const BlockIndex& i) {
    auto it = block_map_.find(i);
    return it != block_map_.end() ? it->second : typename BlockType::Ptr();
  }
  typename BlockType::ConstPtr getBlockPtrByIndex(const BlockIndex& i) const {
    auto it = block_map_.find(i);
    return it != block_map_.end() ? it->second : typename BlockType::ConstPtr();
  }
  typename BlockType::Ptr allocateBlockPtrByIndex(const BlockIndex& i) {
    auto it = block_map_.find(i);
    return it != block_map_.end() ? it->second : allocateNewBlock(i);
  }
  typename BlockType::Ptr allocateNewBlock(const BlockIndex& i) {
    auto ins = block_map_.emplace(i, std::make_shared<BlockType>(voxels_per_side_, voxel_size_, getOriginPointFromGridIndex(i, block_size_)));
    KSG_CHECK(ins.second) << "Block already exists when allocating";
    return ins.first->second;
  }
  void insertBlock(const std::pair<const BlockIndex, typename BlockType::Ptr>& p) { block_map_.insert(p); }
  void removeBlock(const BlockIndex& i) { block_map_.erase(i); }
  void removeAllBlocks() { block_map_.clear(); }
  bool hasBlock(const BlockIndex& i) const { return block_map_.count(i) > 0; }
  void getAllAllocatedBlocks(BlockIndexList* blocks) const {
    blocks->clear();
    blocks->reserve(block_map_.size());
    for (const auto& kv : block_map_) blocks->emplace_back(kv.first);
  }
  void getAllUpdatedBlocks(BlockIndexList* blocks) const {
    blocks->clear();
    for (const auto& kv : block_map_) if (kv.second->updated()) blocks->emplace_back(kv.first);
  }
  size_t getNumberOfAllocatedBlocks() const { return block_map_.size(); }
  BlockIndex computeBlockIndexFromCoordinates(const Point& p) const { return getGridIndexFromPoint(p, block_size_inv_); }
  FloatingPoint block_size() const { return block_size_; }
  FloatingPoint block_size_inv() const { return block_size_inv_; }
  FloatingPoint voxel_size() const { return voxel_size_; }
  FloatingPoint voxel_size_inv() const { return voxel_size_inv_; }
  size_t voxels_per_side() const { return voxels_per_side_; }
  FloatingPoint voxels_per_side_inv() const { return voxels_per_side_inv_; }
  size_t getMemorySize() const { return block_map_.size() * voxels_per_side_ * voxels_per_side_ * voxels_per_side_ * sizeof(VoxelType); }
 private:
  FloatingPoint voxel_size_;
  size_t voxels_per_side_;
  FloatingPoint block_size_, voxel_size_inv_, block_size_inv_, voxels_per_side_inv_;
  BlockHashMap block_map_;
};
}  // namespace voxblox
