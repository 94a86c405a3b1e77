// Drop-in for kimera::SemanticTsdfIntegratorFactory (reference factory.h:49-93, factory.cpp:43-88).
#pragma once
#include <array>
#include <memory>
#include <string>
#include "kimera_semantics/semantic_integrator_base.h"
namespace kimera {
enum class SemanticTsdfIntegratorType : int { kMerged = 0, kFast = 1 };
const std::array<std::string, 2> kSemanticTsdfIntegratorTypeNames = {{/*kMerged*/ "merged", /*kFast*/ "fast"}};
class SemanticTsdfIntegratorFactory {
 public:
  static std::unique_ptr<vxb::TsdfIntegratorBase> create(const std::string& integrator_type_name,
                                                         const vxb::TsdfIntegratorBase::Config& config,
                                                         const SemanticIntegratorBase::SemanticConfig& semantic_config,
                                                         vxb::Layer<vxb::TsdfVoxel>* tsdf_layer,
                                                         vxb::Layer<SemanticVoxel>* semantic_layer);
  static std::unique_ptr<vxb::TsdfIntegratorBase> create(const SemanticTsdfIntegratorType& integrator_type,
                                                         const vxb::TsdfIntegratorBase::Config& config,
                                                         const SemanticIntegratorBase::SemanticConfig& semantic_config,
                                                         vxb::Layer<vxb::TsdfVoxel>* tsdf_layer,
                                                         vxb::Layer<SemanticVoxel>* semantic_layer);
 private:
  SemanticTsdfIntegratorFactory() = default;
  virtual ~SemanticTsdfIntegratorFactory() = default;
};
}  // namespace kimera
