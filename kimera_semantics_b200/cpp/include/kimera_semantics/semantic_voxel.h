// Mirror of kimera_semantics/include/kimera_semantics/semantic_voxel.h (reference semantic_voxel.h:14-27).
#pragma once
#include "kimera_semantics/color.h"
#include "kimera_semantics/common.h"
namespace kimera {
struct SemanticVoxel {
  SemanticLabel semantic_label = 0u;
  SemanticProbabilities semantic_priors = SemanticProbabilities::Constant(-0.60205999132);
  HashableColor color = HashableColor(vxb::Color::Gray());
};
}  // namespace kimera
