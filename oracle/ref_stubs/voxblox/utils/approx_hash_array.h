// Stand-in for voxblox/utils/approx_hash_array.h (TEST INFRASTRUCTURE): SURVEY.md A.4.
#pragma once
#include <atomic>
#include <limits>
#include <vector>
#include <voxblox/core/common.h>

namespace voxblox {

// 2^bits payloads addressed by the low bits of a hash; distinct indices may share one.
template <size_t unmasked_bits, typename StoredElement, typename IndexType, typename IndexTypeHasher>
class ApproxHashArray {
 public:
  ApproxHashArray() : slots_(size_t(1) << unmasked_bits) {}
  StoredElement& get(const size_t& hash) { return slots_[hash & kMask]; }
  StoredElement& get(const IndexType& index, size_t* hash) {
    *hash = hasher_(index);
    return get(*hash);
  }
  StoredElement& get(const IndexType& index) { return get(hasher_(index)); }

 private:
  static constexpr size_t kMask = (size_t(1) << unmasked_bits) - 1;
  std::vector<StoredElement> slots_;
  IndexTypeHasher hasher_;
};

// A set that only remembers the latest hash written to each of its 2^bits slots.
template <size_t unmasked_bits, size_t full_reset_threshold, typename IndexType, typename IndexTypeHasher>
class ApproxHashSet {
 public:
  ApproxHashSet() : offset_(0), slots_(size_t(1) << unmasked_bits) { wipe(); }

  bool isHashCurrentlyPresent(const size_t& hash) {
    return slots_[(hash + offset_) & kMask].load(std::memory_order_relaxed) == hash + offset_;
  }
  bool isHashCurrentlyPresent(const IndexType& index) { return isHashCurrentlyPresent(hasher_(index)); }

  // true when the value was not there (and is now); false when the slot already held it.
  bool replaceHash(const size_t& hash) {
    const size_t value = hash + offset_;
    std::atomic<size_t>& slot = slots_[value & kMask];
    if (slot.load(std::memory_order_relaxed) == value) return false;
    slot.store(value, std::memory_order_relaxed);
    return true;
  }
  bool replaceHash(const IndexType& index) { return replaceHash(hasher_(index)); }

  // Cheap invalidation: shifting every future value by one makes old entries (almost always) mismatch.
  void resetApproxSet() {
    if (++offset_ >= full_reset_threshold) {
      wipe();
      offset_ = 0;
    }
  }

 private:
  void wipe() {
    for (std::atomic<size_t>& s : slots_) s.store(0, std::memory_order_relaxed);
    slots_[0].store(std::numeric_limits<size_t>::max());  // hash 0 must not read as present
  }
  static constexpr size_t kMask = (size_t(1) << unmasked_bits) - 1;
  size_t offset_;
  std::vector<std::atomic<size_t>> slots_;
  IndexTypeHasher hasher_;
};

}  // namespace voxblox
