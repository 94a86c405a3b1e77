// Drop-in for kimera::MergedSemanticTsdfIntegrator (reference merged.h:56-120).
#pragma once
#include "kimera_semantics/gpu_integrator_core.h"
namespace kimera {
class MergedSemanticTsdfIntegrator : public vxb::MergedTsdfIntegrator, public SemanticIntegratorBase {
 public:
  MergedSemanticTsdfIntegrator(const Config& config, const SemanticConfig& semantic_config,
                               vxb::Layer<vxb::TsdfVoxel>* tsdf_layer, vxb::Layer<SemanticVoxel>* semantic_layer);
  virtual ~MergedSemanticTsdfIntegrator() = default;
  // labels encoded as colours (merged.cpp:65-95)
  virtual void integratePointCloud(const vxb::Transformation& T_G_C, const vxb::Pointcloud& points_C,
                                   const vxb::Colors& colors, const bool freespace_points = false) override;
  // label-explicit overload (merged.h:82-86, merged.cpp:97-149). As in the reference, `colors` only has to match in
  // size: the merged colour that reaches the voxels is the blend of these colours.
  void integratePointCloud(const vxb::Transformation& T_G_C, const vxb::Pointcloud& points_C, const HashableColors& colors,
                           const SemanticLabels& semantic_labels, const bool freespace_points = false);
  GpuIntegratorCore& gpu() { return core_; }
 private:
  GpuIntegratorCore core_;
};
}  // namespace kimera
