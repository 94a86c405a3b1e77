#pragma once
#include "voxblox/core/common.h"
namespace voxblox {
struct TsdfVoxel {
  float distance = 0.0f;
  float weight = 0.0f;
  Color color;
};
}  // namespace voxblox
