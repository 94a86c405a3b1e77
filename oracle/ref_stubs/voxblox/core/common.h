// Stand-in for voxblox/core/common.h (TEST INFRASTRUCTURE, see oracle/ref_stubs/README.md).  voxblox is not vendored by
// the reference; this restates SURVEY.md Appendix A.0, A.2 and A.8 on top of the Eigen stand-in.
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <deque>
#include <limits>
#include <list>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

#include <glog/logging.h>
#include <Eigen/Core>

namespace voxblox {

typedef float FloatingPoint;
typedef int IndexElement;
typedef int64_t LongIndexElement;

typedef Eigen::Matrix<FloatingPoint, 3, 1> Point;
typedef Eigen::Matrix<FloatingPoint, 3, 1> Ray;
typedef Eigen::Matrix<IndexElement, 3, 1> AnyIndex;
typedef AnyIndex VoxelIndex;
typedef AnyIndex BlockIndex;
typedef Eigen::Matrix<LongIndexElement, 3, 1> LongIndex;
typedef LongIndex GlobalIndex;

template <typename Type>
using AlignedVector = std::vector<Type, Eigen::aligned_allocator<Type>>;
template <typename Type>
using AlignedDeque = std::deque<Type, Eigen::aligned_allocator<Type>>;
template <typename Type>
using AlignedList = std::list<Type, Eigen::aligned_allocator<Type>>;

typedef AlignedVector<Point> Pointcloud;
typedef AlignedVector<BlockIndex> BlockIndexList;
typedef AlignedVector<GlobalIndex> GlobalIndexVector;

constexpr FloatingPoint kEpsilon = 1e-6;
constexpr FloatingPoint kFloatEpsilon = 1e-6;
constexpr FloatingPoint kCoordinateEpsilon = 1e-6;

// minkindr QuatTransformationTemplate<float>, reduced to what the integrators call (A.8).
class Transformation {
 public:
  // getRotation().toImplementation() is Eigen::Quaternionf in minkindr; callers read w() x() y() z().
  struct QuaternionImplementation {
    float qw, qx, qy, qz;
    float w() const { return qw; }
    float x() const { return qx; }
    float y() const { return qy; }
    float z() const { return qz; }
  };
  struct Rotation {
    QuaternionImplementation q;
    const QuaternionImplementation& toImplementation() const { return q; }
  };
  Transformation() : w_(1.0f), v_(0.0f, 0.0f, 0.0f), t_(0.0f, 0.0f, 0.0f) {}
  Transformation(float qw, float qx, float qy, float qz, const Point& t) : w_(qw), v_(qx, qy, qz), t_(t) {}
  const Point& getPosition() const { return t_; }
  Rotation getRotation() const { return Rotation{QuaternionImplementation{w_, v_[0], v_[1], v_[2]}}; }
  Point operator*(const Point& p) const {
    // Eigen's quaternion * vector: uv = 2 * (q.vec x p);  p + w * uv + q.vec x uv;  then + t.
    Point uv = cross(v_, p);
    uv += uv;
    return (p + uv * w_ + cross(v_, uv)) + t_;
  }

 private:
  static Point cross(const Point& a, const Point& b) {
    return Point(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
  }
  float w_;
  Point v_, t_;
};

// A.2 grid helpers.
template <typename IndexType>
inline IndexType getGridIndexFromPoint(const Point& scaled_point) {
  return IndexType((typename IndexType::Scalar)std::floor(scaled_point.x() + kCoordinateEpsilon),
                   (typename IndexType::Scalar)std::floor(scaled_point.y() + kCoordinateEpsilon),
                   (typename IndexType::Scalar)std::floor(scaled_point.z() + kCoordinateEpsilon));
}
template <typename IndexType>
inline IndexType getGridIndexFromPoint(const Point& point, const FloatingPoint grid_size_inv) {
  return IndexType((typename IndexType::Scalar)std::floor(point.x() * grid_size_inv + kCoordinateEpsilon),
                   (typename IndexType::Scalar)std::floor(point.y() * grid_size_inv + kCoordinateEpsilon),
                   (typename IndexType::Scalar)std::floor(point.z() * grid_size_inv + kCoordinateEpsilon));
}
template <typename IndexType>
inline Point getCenterPointFromGridIndex(const IndexType& idx, FloatingPoint grid_size) {
  return Point((static_cast<FloatingPoint>(idx.x()) + 0.5f) * grid_size, (static_cast<FloatingPoint>(idx.y()) + 0.5f) * grid_size,
               (static_cast<FloatingPoint>(idx.z()) + 0.5f) * grid_size);
}
template <typename IndexType>
inline Point getOriginPointFromGridIndex(const IndexType& idx, FloatingPoint grid_size) {
  return Point(static_cast<FloatingPoint>(idx.x()) * grid_size, static_cast<FloatingPoint>(idx.y()) * grid_size,
               static_cast<FloatingPoint>(idx.z()) * grid_size);
}
inline BlockIndex getBlockIndexFromGlobalVoxelIndex(const GlobalIndex& g, FloatingPoint voxels_per_side_inv) {
  return BlockIndex((IndexElement)std::floor(static_cast<FloatingPoint>(g.x()) * voxels_per_side_inv),
                    (IndexElement)std::floor(static_cast<FloatingPoint>(g.y()) * voxels_per_side_inv),
                    (IndexElement)std::floor(static_cast<FloatingPoint>(g.z()) * voxels_per_side_inv));
}
inline VoxelIndex getLocalFromGlobalVoxelIndex(const GlobalIndex& g, const int voxels_per_side) {
  CHECK((voxels_per_side & (voxels_per_side - 1)) == 0) << "voxels_per_side must be a power of two";
  constexpr LongIndexElement offset = LongIndexElement(1) << (8 * sizeof(IndexElement) - 1);
  const LongIndexElement mask = voxels_per_side - 1;
  return VoxelIndex((IndexElement)((g.x() + offset) & mask), (IndexElement)((g.y() + offset) & mask),
                    (IndexElement)((g.z() + offset) & mask));
}
inline GlobalIndex getGlobalVoxelIndexFromBlockAndVoxelIndex(const BlockIndex& b, const VoxelIndex& v, int voxels_per_side) {
  return GlobalIndex((LongIndexElement)b.x() * voxels_per_side + v.x(), (LongIndexElement)b.y() * voxels_per_side + v.y(),
                     (LongIndexElement)b.z() * voxels_per_side + v.z());
}

template <typename T>
inline int signum(T x) { return (x == 0) ? 0 : (x < 0 ? -1 : 1); }

}  // namespace voxblox

#include <voxblox/core/color.h>
