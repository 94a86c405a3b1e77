// Stand-in for voxblox/integrator/integrator_utils.h (TEST INFRASTRUCTURE): SURVEY.md A.3 (ThreadSafeIndex) and
// A.7 (RayCaster).
#pragma once
#include <algorithm>
#include <atomic>
#include <string>
#include <utility>
#include <vector>
#include <voxblox/core/block_hash.h>
#include <voxblox/core/common.h>

namespace voxblox {

// Hands out each point index exactly once, to any number of threads, in an order that depends on the mode.
class ThreadSafeIndex {
 public:
  explicit ThreadSafeIndex(size_t number_of_points) : next_(0), count_(number_of_points) {}
  virtual ~ThreadSafeIndex() {}
  bool getNextIndex(size_t* idx) {
    const size_t sequence = next_.fetch_add(1);
    if (sequence >= count_) return false;
    *idx = getNextIndexImpl(sequence);
    return true;
  }
  void reset() { next_.store(0); }

 protected:
  virtual size_t getNextIndexImpl(size_t sequence) = 0;
  std::atomic<size_t> next_;
  const size_t count_;
};

// Interleaves the cloud in strides of 1024 so that neighbouring pixels land far apart in time.
class MixedThreadSafeIndex : public ThreadSafeIndex {
 public:
  explicit MixedThreadSafeIndex(size_t number_of_points) : ThreadSafeIndex(number_of_points), full_groups_(number_of_points / kStride) {}

 protected:
  size_t getNextIndexImpl(size_t sequence) override {
    if (full_groups_ * kStride <= sequence) return sequence;  // ragged tail keeps its place
    return (sequence % full_groups_) * kStride + sequence / full_groups_;
  }

 private:
  static constexpr size_t kStride = 1024;
  const size_t full_groups_;
};

// Closest points first. (Upstream uses std::sort, so equal ranges come out in an unspecified order; ties are kept in
// index order here, the same choice the oracle makes.)
class SortedThreadSafeIndex : public ThreadSafeIndex {
 public:
  explicit SortedThreadSafeIndex(const Pointcloud& points_C) : ThreadSafeIndex(points_C.size()) {
    std::vector<std::pair<size_t, FloatingPoint>> keyed(points_C.size());
    for (size_t i = 0; i < points_C.size(); ++i) keyed[i] = std::make_pair(i, points_C[i].squaredNorm());
    std::stable_sort(keyed.begin(), keyed.end(),
                     [](const std::pair<size_t, FloatingPoint>& a, const std::pair<size_t, FloatingPoint>& b) { return a.second < b.second; });
    order_.resize(keyed.size());
    for (size_t i = 0; i < keyed.size(); ++i) order_[i] = keyed[i].first;
  }

 protected:
  size_t getNextIndexImpl(size_t sequence) override { return order_[sequence]; }

 private:
  std::vector<size_t> order_;
};

class ThreadSafeIndexFactory {
 public:
  static ThreadSafeIndex* get(const std::string& mode, const Pointcloud& points_C) {
    if (mode == "mixed") return new MixedThreadSafeIndex(points_C.size());
    if (mode == "sorted") return new SortedThreadSafeIndex(points_C);
    LOG(FATAL) << "Unknown integration order mode: '" << mode << "'!";
    return nullptr;
  }
};

// Voxel traversal of a segment given in world units (Amanatides-Woo stepping on the unit grid after scaling).
class RayCaster {
 public:
  RayCaster(const Point& origin, const Point& point_G, const bool is_clearing_ray, const bool voxel_carving_enabled,
            const FloatingPoint max_ray_length_m, const FloatingPoint voxel_size_inv, const FloatingPoint truncation_distance,
            const bool cast_from_origin = true) {
    const Ray unit_ray = (point_G - origin).normalized();
    Point ray_start, ray_end;
    if (is_clearing_ray) {
      FloatingPoint ray_length = (point_G - origin).norm();
      ray_length = std::min(std::max(ray_length - truncation_distance, static_cast<FloatingPoint>(0.0)), max_ray_length_m);
      ray_end = origin + unit_ray * ray_length;
      ray_start = voxel_carving_enabled ? origin : ray_end;
    } else {
      ray_end = point_G + unit_ray * truncation_distance;
      ray_start = voxel_carving_enabled ? origin : (point_G - unit_ray * truncation_distance);
    }
    const Point start_scaled = ray_start * voxel_size_inv;
    const Point end_scaled = ray_end * voxel_size_inv;
    if (cast_from_origin) {
      setupRayCaster(start_scaled, end_scaled);
    } else {
      setupRayCaster(end_scaled, start_scaled);
    }
  }
  RayCaster(const Point& start_scaled, const Point& end_scaled) { setupRayCaster(start_scaled, end_scaled); }

  // Writes the next voxel of the segment; false once the segment is exhausted.
  bool nextRayIndex(GlobalIndex* ray_index) {
    if (current_step_++ > ray_length_in_steps_) return false;
    *ray_index = curr_index_;
    int axis = 0;  // first minimum of t_to_next_boundary_
    if (t_to_next_boundary_[1] < t_to_next_boundary_[axis]) axis = 1;
    if (t_to_next_boundary_[2] < t_to_next_boundary_[axis]) axis = 2;
    curr_index_[axis] += ray_step_signs_[axis];
    t_to_next_boundary_[axis] += t_step_size_[axis];
    return true;
  }

 private:
  void setupRayCaster(const Point& start_scaled, const Point& end_scaled) {
    if (start_scaled.hasNaN() || end_scaled.hasNaN()) {
      ray_length_in_steps_ = 0;
      current_step_ = 0;
      return;
    }
    curr_index_ = getGridIndexFromPoint<GlobalIndex>(start_scaled);
    const GlobalIndex end_index = getGridIndexFromPoint<GlobalIndex>(end_scaled);
    const GlobalIndex diff_index = end_index - curr_index_;
    current_step_ = 0;
    ray_length_in_steps_ = (unsigned)(std::abs(diff_index.x()) + std::abs(diff_index.y()) + std::abs(diff_index.z()));
    const Ray ray_scaled = end_scaled - start_scaled;
    for (int k = 0; k < 3; ++k) {
      ray_step_signs_[k] = signum(ray_scaled[k]);
      const FloatingPoint corrected_step = (FloatingPoint)std::max(0, (int)ray_step_signs_[k]);
      const FloatingPoint start_scaled_shifted = start_scaled[k] - static_cast<FloatingPoint>(curr_index_[k]);
      const FloatingPoint distance_to_boundary = corrected_step - start_scaled_shifted;
      // (upstream has a guard for |ray| < 0 here, which can never fire: a zero component divides by zero)
      t_to_next_boundary_[k] = distance_to_boundary / ray_scaled[k];
      t_step_size_[k] = static_cast<FloatingPoint>(ray_step_signs_[k]) / ray_scaled[k];
    }
  }

  Ray t_to_next_boundary_;
  GlobalIndex curr_index_;
  GlobalIndex ray_step_signs_;
  Ray t_step_size_;
  unsigned ray_length_in_steps_;
  unsigned current_step_;
};

}  // namespace voxblox
