// Compile-and-run check that the call patterns the reference uses at this boundary compile unchanged against the shim headers
// (SURVEY.md 7.3 item 7): SemanticTsdfServer's constructor sequence (kimera_semantics_ros/src/semantic_tsdf_server.cpp:58-79),
// the Layer/Block accessors used by kimera_semantics/src/simulation/semantic_simulation_world.cpp:62-73,99-109, SemanticConfig
// fields (ros_params.cpp:38-77) and the enum/string factory overloads (semantic_simulation_server.cpp:19-24).
// No integrator is constructed (the program is a pure API / layout check).  It uses ONLY the reference's API, and the same file is
// also compiled and run against the reference's real headers (test_shim_cpu.py::test_api_compat_source_also_builds_against_the_
// reference_headers): what compiles and passes there must compile and pass here.
#include <cstdio>
#include <fstream>
#include <type_traits>
#include <unistd.h>
#include "kimera_semantics/semantic_tsdf_integrator_factory.h"
#include "kimera_semantics/semantic_tsdf_integrator_fast.h"
#include "kimera_semantics/semantic_tsdf_integrator_merged.h"

namespace vxb = voxblox;
using namespace kimera;

static_assert(std::is_base_of<vxb::TsdfIntegratorBase, FastSemanticTsdfIntegrator>::value, "fast.h:63-66");
static_assert(std::is_base_of<SemanticIntegratorBase, FastSemanticTsdfIntegrator>::value, "fast.h:63-66");
static_assert(std::is_base_of<vxb::MergedTsdfIntegrator, MergedSemanticTsdfIntegrator>::value, "merged.h:56-59");
static_assert(std::is_base_of<SemanticIntegratorBase, MergedSemanticTsdfIntegrator>::value, "merged.h:56-59");
static_assert(static_cast<int>(SemanticTsdfIntegratorType::kMerged) == 0 && static_cast<int>(SemanticTsdfIntegratorType::kFast) == 1, "factory.h:49-52");
static_assert(static_cast<int>(ColorMode::kColor) == 0 && static_cast<int>(ColorMode::kSemantic) == 1 && static_cast<int>(ColorMode::kSemanticProbability) == 2, "base.h:54-58");
static_assert(kUnknownSemanticLabelId == 0u && std::is_same<SemanticLabel, uint8_t>::value, "common.h:17-21");
static_assert(sizeof(vxb::TsdfVoxel) == 12, "TsdfVoxel = {float distance, float weight, Color}");

// semantic_simulation_world.cpp:99-109 setSemanticVoxel-style write access
static void setSemanticVoxel(const SemanticLabel& label, const HashableColor& color, SemanticVoxel* voxel) {
  voxel->semantic_label = label;
  voxel->color = color;
}

int main() {
  // ros_params.cpp:38-77
  SemanticIntegratorBase::SemanticConfig semantic_config;
  semantic_config.semantic_measurement_probability_ = 0.8f;
  semantic_config.color_mode = ColorMode::kSemantic;
  char csv_path[] = "/tmp/api_compat_labels_XXXXXX";
  const int fd = mkstemp(csv_path);
  if (fd < 0) return 10;
  close(fd);
  {
    std::ofstream csv(csv_path);
    csv << "name,red,green,blue,alpha,id\ngrass,0,255,0,255,1\n";
  }
  semantic_config.semantic_label_to_color_ = std::make_shared<SemanticLabel2Color>(std::string(csv_path));   // color.cpp:42-67
  unlink(csv_path);
  if (semantic_config.semantic_label_to_color_->getSemanticLabelFromColor(HashableColor(0, 255, 0, 255)) != 1u) return 11;
  semantic_config.dynamic_labels_.push_back(20u);

  vxb::TsdfIntegratorBase::Config config;   // voxblox defaults (A.6)
  if (config.default_truncation_distance != 0.1f || config.max_weight != 10000.0f || config.integration_order_mode != "mixed" ||
      config.start_voxel_subsampling_factor != 2.0f || config.max_consecutive_ray_collisions != 2 || !config.voxel_carving_enabled) return 1;

  // semantic_tsdf_server.cpp:68-69 + semantic_simulation_world.cpp:62-73
  std::unique_ptr<vxb::Layer<SemanticVoxel>> semantic_layer(new vxb::Layer<SemanticVoxel>(0.1f, 16u));
  vxb::Layer<vxb::TsdfVoxel> tsdf_layer(0.1f, 16u);
  const vxb::BlockIndex block_index = semantic_layer->computeBlockIndexFromCoordinates(vxb::Point(-0.05f, 1.7f, 3.3f));
  if (!(block_index == vxb::BlockIndex(-1, 1, 2))) return 2;
  vxb::Block<SemanticVoxel>::Ptr block = semantic_layer->allocateBlockPtrByIndex(block_index);
  for (size_t i = 0; i < block->num_voxels(); ++i) {
    const vxb::Point coords = block->computeCoordinatesFromLinearIndex(i);
    (void)coords;
    setSemanticVoxel(1u, semantic_config.semantic_label_to_color_->getColorFromSemanticLabel(1u), &block->getVoxelByLinearIndex(i));
  }
  const SemanticVoxel fresh;
  if (fresh.semantic_label != 0u || fresh.semantic_priors[0] != static_cast<float>(-0.60205999132) || !(fresh.color == HashableColor(127, 127, 127, 255))) return 3;
  if (semantic_layer->getNumberOfAllocatedBlocks() != 1u || !semantic_layer->hasBlock(block_index) || semantic_layer->voxels_per_side() != 16u) return 4;
  vxb::BlockIndexList all;
  semantic_layer->getAllAllocatedBlocks(&all);
  if (all.size() != 1u || block->block_index() != block_index) return 5;

  // factory call sites (semantic_tsdf_server.cpp:71-77 string overload, semantic_simulation_server.cpp:19-24 enum overload):
  // taken as function pointers so that the signatures are checked without needing a device
  typedef std::unique_ptr<vxb::TsdfIntegratorBase> (*CreateByName)(const std::string&, const vxb::TsdfIntegratorBase::Config&,
                                                                  const SemanticIntegratorBase::SemanticConfig&, vxb::Layer<vxb::TsdfVoxel>*,
                                                                  vxb::Layer<SemanticVoxel>*);
  typedef std::unique_ptr<vxb::TsdfIntegratorBase> (*CreateByEnum)(const SemanticTsdfIntegratorType&, const vxb::TsdfIntegratorBase::Config&,
                                                                  const SemanticIntegratorBase::SemanticConfig&, vxb::Layer<vxb::TsdfVoxel>*,
                                                                  vxb::Layer<SemanticVoxel>*);
  CreateByName by_name = &SemanticTsdfIntegratorFactory::create;
  CreateByEnum by_enum = &SemanticTsdfIntegratorFactory::create;
  void (vxb::TsdfIntegratorBase::*integrate)(const vxb::Transformation&, const vxb::Pointcloud&, const vxb::Colors&, const bool) =
      &vxb::TsdfIntegratorBase::integratePointCloud;
  void (MergedSemanticTsdfIntegrator::*integrate_labels)(const vxb::Transformation&, const vxb::Pointcloud&, const HashableColors&,
                                                          const SemanticLabels&, const bool) = &MergedSemanticTsdfIntegrator::integratePointCloud;
  if (!by_name || !by_enum || !integrate || !integrate_labels) return 6;
  if (kSemanticTsdfIntegratorTypeNames[0] != "merged" || kSemanticTsdfIntegratorTypeNames[1] != "fast") return 7;

  // minkindr-style transformation: T * p and getPosition()
  const vxb::Transformation T(1.0f, 0.0f, 0.0f, 0.0f, vxb::Point(1.0f, 2.0f, 3.0f));
  const vxb::Point q = T * vxb::Point(0.5f, 0.25f, -4.0f);
  if (q.x() != 1.5f || q.y() != 2.25f || q.z() != -1.0f || T.getPosition().z() != 3.0f) return 8;
  // color.h:58-82: label -> colour table of the simulation (fixed colours for labels 0..7, random for the rest)
  const SemanticLabelToColorMap random_table = getRandomSemanticLabelToColorMap();
  if (random_table.size() != 255u || !(random_table.at(0) == HashableColor(vxb::Color::Gray())) ||
      !(random_table.at(3) == HashableColor(vxb::Color::Purple())) || !(random_table.at(7) == HashableColor(vxb::Color::Yellow()))) return 9;
  std::printf("api compat ok\n");
  return 0;
}
