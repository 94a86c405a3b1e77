// Mirror of kimera_semantics/include/kimera_semantics/common.h (reference common.h:17-38).
#pragma once
#include <array>
#include <cmath>
#include <memory>
#include "voxblox/core/common.h"
namespace kimera {
namespace vxb = voxblox;
typedef uint8_t SemanticLabel;
typedef vxb::AlignedVector<SemanticLabel> SemanticLabels;
static constexpr uint8_t kUnknownSemanticLabelId = 0u;
// The reference fixes this at compile time (common.h:26). The device path takes the class count at run time
// (ksg_config.num_labels); the host-side SemanticVoxel keeps a compile-time size, overridable at build time.
#ifndef KIMERA_TOTAL_NUMBER_OF_LABELS
#define KIMERA_TOTAL_NUMBER_OF_LABELS 21
#endif
static constexpr size_t kTotalNumberOfLabels = KIMERA_TOTAL_NUMBER_OF_LABELS;
typedef vxb::FloatingPoint SemanticProbability;
// stand-in for Eigen::Matrix<SemanticProbability, kTotalNumberOfLabels, 1>
struct SemanticProbabilities {
  std::array<SemanticProbability, kTotalNumberOfLabels> v;
  static SemanticProbabilities Constant(SemanticProbability c) { SemanticProbabilities p; p.v.fill(c); return p; }
  static SemanticProbabilities Zero() { return Constant(0.0f); }
  SemanticProbability& operator[](size_t i) { return v[i]; }
  const SemanticProbability& operator[](size_t i) const { return v[i]; }
  size_t size() const { return kTotalNumberOfLabels; }
  SemanticProbability* data() { return v.data(); }
  const SemanticProbability* data() const { return v.data(); }
  void setConstant(SemanticProbability c) { v.fill(c); }
  bool hasNaN() const { for (size_t i = 0; i < kTotalNumberOfLabels; ++i) if (v[i] != v[i]) return true; return false; }
  // Eigen norm()/normalize(): L2 norm (left-to-right accumulation here), division by it when it is positive
  SemanticProbability norm() const {
    SemanticProbability s = v[0] * v[0];
    for (size_t i = 1; i < kTotalNumberOfLabels; ++i) s += v[i] * v[i];
    return std::sqrt(s);
  }
  void normalize() { const SemanticProbability n = norm(); if (n > 0.0f) for (size_t i = 0; i < kTotalNumberOfLabels; ++i) v[i] /= n; }
  // Eigen maxCoeff(&index): first maximum wins
  SemanticProbability maxCoeff(SemanticLabel* index) const {
    size_t best = 0;
    for (size_t i = 1; i < kTotalNumberOfLabels; ++i) if (v[i] > v[best]) best = i;
    *index = static_cast<SemanticLabel>(best);
    return v[best];
  }
};
// row-major kTotalNumberOfLabels x kTotalNumberOfLabels
struct SemanticLikelihoodFunction {
  std::array<SemanticProbability, kTotalNumberOfLabels * kTotalNumberOfLabels> m;
  SemanticProbability& operator()(size_t i, size_t j) { return m[i * kTotalNumberOfLabels + j]; }
  const SemanticProbability& operator()(size_t i, size_t j) const { return m[i * kTotalNumberOfLabels + j]; }
};
template <typename T, typename... Args>
std::unique_ptr<T> make_unique(Args&&... args) { return std::unique_ptr<T>(new T(std::forward<Args>(args)...)); }
}  // namespace kimera
