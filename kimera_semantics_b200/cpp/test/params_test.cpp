// Prints what the shim's ROS-free parameter reading (kimera_semantics/params.h) makes of a "key: value" file, in the format of the
// reference-side probe used by tests/test_shim_cpu.py, followed by the voxblox-side values.   params_test <file>
#include <cstdio>
#include "kimera_semantics/params.h"
using namespace kimera;
int main(int argc, char** argv) {
  if (argc < 2) return 2;
  const ParamMap p = readParamFile(argv[1]);
  const std::string method = getSemanticTsdfIntegratorTypeFromParams(p);
  const std::string csv = getSemanticLabelToColorCsvFilepathFromParams(p);
  const SemanticIntegratorBase::SemanticConfig sc = getSemanticTsdfIntegratorConfigFromParams(p);
  std::printf("method=%s\ncsv=%s\nprobability=%.9g\ncolor_mode=%d\ndynamic=", method.c_str(), csv.c_str(),
              (double)sc.semantic_measurement_probability_, (int)sc.color_mode);
  for (size_t i = 0; i < sc.dynamic_labels_.size(); ++i) std::printf("%s%d", i ? "," : "", (int)sc.dynamic_labels_[i]);
  std::printf("\nlabels=%zu\n", sc.semantic_label_to_color_->semantic_label_to_color_map_.size());
  const vxb::TsdfIntegratorBase::Config c = getTsdfIntegratorConfigFromParams(p);
  const SemanticTsdfServer::Params s = getServerParamsFromParams(p);
  std::printf("voxel_size=%.9g vps=%zu trunc=%.9g max_ray=%.9g carving=%d const_weight=%d throttle=%.9g order=%s\n", (double)s.tsdf_voxel_size,
              s.tsdf_voxels_per_side, (double)c.default_truncation_distance, (double)c.max_ray_length_m, (int)c.voxel_carving_enabled,
              (int)c.use_const_weight, s.min_time_between_msgs_sec, c.integration_order_mode.c_str());
  return 0;
}
