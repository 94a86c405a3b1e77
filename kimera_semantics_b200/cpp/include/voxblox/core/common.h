// Minimal, dependency-free subset of voxblox/core/common.h + block_hash.h + color.h:
// only the types that cross the kimera_semantics integrator boundary. No Eigen / minkindr / glog.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <memory>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

// glog-style aborts (the reference's error convention): no exceptions, no error codes.
namespace ksg_shim {
struct FatalStream {
  std::ostringstream os;
  const char* file; int line;
  FatalStream(const char* f, int l, const char* cond) : file(f), line(l) { os << "Check failed: " << cond << " "; }
  template <typename T> FatalStream& operator<<(const T& v) { os << v; return *this; }
  [[noreturn]] ~FatalStream() { std::fprintf(stderr, "F %s:%d] %s\n", file, line, os.str().c_str()); std::abort(); }
};
struct Voidify { void operator&(const FatalStream&) {} };
}  // namespace ksg_shim
#define KSG_CHECK(cond) (cond) ? (void)0 : ksg_shim::Voidify() & ksg_shim::FatalStream(__FILE__, __LINE__, #cond)
#define KSG_LOG_FATAL ksg_shim::FatalStream(__FILE__, __LINE__, "LOG(FATAL)")

namespace voxblox {

typedef float FloatingPoint;
typedef int IndexElement;
typedef int64_t LongIndexElement;

// 3 contiguous floats, like Eigen::Matrix<float,3,1>
struct Point {
  FloatingPoint v[3];
  Point() : v{0, 0, 0} {}
  Point(FloatingPoint x, FloatingPoint y, FloatingPoint z) : v{x, y, z} {}
  FloatingPoint& operator[](int i) { return v[i]; }
  const FloatingPoint& operator[](int i) const { return v[i]; }
  FloatingPoint x() const { return v[0]; }
  FloatingPoint y() const { return v[1]; }
  FloatingPoint z() const { return v[2]; }
  static Point Zero() { return Point(); }
};
typedef Point Ray;
template <typename T> struct Index3 {
  T v[3];
  Index3() : v{0, 0, 0} {}
  Index3(T x, T y, T z) : v{x, y, z} {}
  T& operator[](int i) { return v[i]; }
  const T& operator[](int i) const { return v[i]; }
  T x() const { return v[0]; }
  T y() const { return v[1]; }
  T z() const { return v[2]; }
  bool operator==(const Index3& o) const { return v[0] == o.v[0] && v[1] == o.v[1] && v[2] == o.v[2]; }
  bool operator!=(const Index3& o) const { return !(*this == o); }
};
typedef Index3<IndexElement> AnyIndex;
typedef AnyIndex VoxelIndex;
typedef AnyIndex BlockIndex;
typedef Index3<LongIndexElement> LongIndex;
typedef LongIndex GlobalIndex;

template <typename T> using AlignedVector = std::vector<T>;
typedef AlignedVector<Point> Pointcloud;
typedef AlignedVector<BlockIndex> BlockIndexList;

constexpr FloatingPoint kEpsilon = 1e-6f;
constexpr float kFloatEpsilon = 1e-6f;

struct Color {
  uint8_t r, g, b, a;
  Color() : r(0), g(0), b(0), a(0) {}
  Color(uint8_t r_, uint8_t g_, uint8_t b_) : Color(r_, g_, b_, 255) {}
  Color(uint8_t r_, uint8_t g_, uint8_t b_, uint8_t a_) : r(r_), g(g_), b(b_), a(a_) {}
  static const Color White() { return Color(255, 255, 255); }
  static const Color Black() { return Color(0, 0, 0); }
  static const Color Gray() { return Color(127, 127, 127); }
  static const Color Red() { return Color(255, 0, 0); }
  static const Color Green() { return Color(0, 255, 0); }
  static const Color Blue() { return Color(0, 0, 255); }
  static const Color Yellow() { return Color(255, 255, 0); }
  static const Color Orange() { return Color(255, 127, 0); }
  static const Color Purple() { return Color(127, 0, 255); }
  static const Color Teal() { return Color(0, 255, 255); }
  static const Color Pink() { return Color(255, 0, 127); }
};
typedef AlignedVector<Color> Colors;
inline Color randomColor() { return Color(std::rand() % 256, std::rand() % 256, std::rand() % 256); }   // voxblox/core/color.h

// AnyIndexHash / LongIndexHash (A.2)
struct AnyIndexHash {
  static constexpr size_t sl = 17191;
  static constexpr size_t sl2 = sl * sl;
  std::size_t operator()(const AnyIndex& i) const { return static_cast<unsigned int>(i.x() + i.y() * sl + i.z() * sl2); }
};
struct LongIndexHash {
  static constexpr size_t sl = 17191;
  static constexpr size_t sl2 = sl * sl;
  std::size_t operator()(const LongIndex& i) const { return static_cast<unsigned int>(i.x() + i.y() * sl + i.z() * sl2); }
};
template <typename V> struct AnyIndexHashMapType { typedef std::unordered_map<AnyIndex, V, AnyIndexHash> type; };
template <typename V> struct LongIndexHashMapType { typedef std::unordered_map<LongIndex, V, LongIndexHash> type; };

// minkindr QuatTransformation<float> subset: rotation quaternion (w, x, y, z) + position
class Transformation {
 public:
  Transformation() : q_{1, 0, 0, 0}, t_() {}
  Transformation(FloatingPoint qw, FloatingPoint qx, FloatingPoint qy, FloatingPoint qz, const Point& t) : q_{qw, qx, qy, qz}, t_(t) {}
  const Point& getPosition() const { return t_; }
  const FloatingPoint* getRotationWxyz() const { return q_; }
  // T * p = q (x) p (x) q^-1 + t, Eigen's quaternion-vector product (A.8)
  Point operator*(const Point& p) const {
    const FloatingPoint qv[3] = {q_[1], q_[2], q_[3]};
    FloatingPoint uv[3] = {qv[1] * p[2] - qv[2] * p[1], qv[2] * p[0] - qv[0] * p[2], qv[0] * p[1] - qv[1] * p[0]};
    for (auto& c : uv) c = c + c;
    const FloatingPoint cr[3] = {qv[1] * uv[2] - qv[2] * uv[1], qv[2] * uv[0] - qv[0] * uv[2], qv[0] * uv[1] - qv[1] * uv[0]};
    return Point(((p[0] + q_[0] * uv[0]) + cr[0]) + t_[0], ((p[1] + q_[0] * uv[1]) + cr[1]) + t_[1], ((p[2] + q_[0] * uv[2]) + cr[2]) + t_[2]);
  }
 private:
  FloatingPoint q_[4];
  Point t_;
};

inline BlockIndex getGridIndexFromPoint(const Point& p, FloatingPoint inv) {
  return BlockIndex((int)std::floor(p.x() * inv + kEpsilon), (int)std::floor(p.y() * inv + kEpsilon), (int)std::floor(p.z() * inv + kEpsilon));
}
inline Point getOriginPointFromGridIndex(const BlockIndex& i, FloatingPoint size) {
  return Point((FloatingPoint)i.x() * size, (FloatingPoint)i.y() * size, (FloatingPoint)i.z() * size);
}

}  // namespace voxblox
