// ROS-free reading of the parameters the reference takes from the ROS parameter server (SURVEY.md 8f NEXT-2):
//   kimera_semantics_ros/src/ros_params.cpp:20-77   method, semantic_label_2_color_csv_filepath, semantic_measurement_probability,
//                                                   semantic_color_mode, dynamic_semantic_labels
//   voxblox_ros getTsdfMapConfigFromRosParam / getTsdfIntegratorConfigFromRosParam (not in the reference tree; parameter names as
//   used by kimera_semantics_ros/launch/kimera_semantics.launch:98-122)
// Parameters live in a string map (what `rosparam` would hold); readParamFile() fills one from "key: value" lines, so a launch
// file's <param> block can be kept as a small text file.  Same defaults, same fatal errors as the reference.
#pragma once
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>
#include "kimera_semantics/semantic_integrator_base.h"
#include "kimera_semantics/semantic_tsdf_server.h"

namespace kimera {

typedef std::map<std::string, std::string> ParamMap;

namespace params_detail {
inline std::string trim(const std::string& s) {
  const size_t b = s.find_first_not_of(" \t\r\n\"'"), e = s.find_last_not_of(" \t\r\n\"'");
  return b == std::string::npos ? std::string() : s.substr(b, e - b + 1);
}
inline bool get(const ParamMap& p, const std::string& k, std::string* v) {
  const ParamMap::const_iterator it = p.find(k);
  if (it == p.end()) return false;
  *v = it->second;
  return true;
}
inline double getDouble(const ParamMap& p, const std::string& k, double def) { std::string v; return get(p, k, &v) ? std::atof(v.c_str()) : def; }
inline int getInt(const ParamMap& p, const std::string& k, int def) { std::string v; return get(p, k, &v) ? std::atoi(v.c_str()) : def; }
inline bool getBool(const ParamMap& p, const std::string& k, bool def) {
  std::string v;
  if (!get(p, k, &v)) return def;
  return v == "true" || v == "True" || v == "1";
}
}  // namespace params_detail

// "key: value" per line, '#' starts a comment, values may be quoted; lists are written [a, b, c].
inline ParamMap readParamFile(const std::string& path) {
  std::ifstream f(path.c_str());
  KSG_CHECK(f.good()) << "Couldn't open file: " << path;
  ParamMap out;
  std::string line;
  while (std::getline(f, line)) {
    const size_t hash = line.find('#');
    if (hash != std::string::npos) line.erase(hash);
    const size_t colon = line.find(':');
    if (colon == std::string::npos) continue;
    const std::string key = params_detail::trim(line.substr(0, colon));
    if (!key.empty()) out[key] = params_detail::trim(line.substr(colon + 1));
  }
  return out;
}

inline std::vector<int> parseIntList(const std::string& text) {
  std::vector<int> out;
  std::string s = text;
  for (size_t i = 0; i < s.size(); ++i)
    if (s[i] == '[' || s[i] == ']' || s[i] == ',') s[i] = ' ';
  std::stringstream ss(s);
  int v;
  while (ss >> v) out.push_back(v);
  return out;
}

// ros_params.cpp:20-30
inline std::string getSemanticTsdfIntegratorTypeFromParams(const ParamMap& p) {
  std::string type = "fast";
  params_detail::get(p, "method", &type);
  return type;
}

// ros_params.cpp:32-37
inline std::string getSemanticLabelToColorCsvFilepathFromParams(const ParamMap& p) {
  std::string path = "semantics2labels.csv";
  params_detail::get(p, "semantic_label_2_color_csv_filepath", &path);
  return path;
}

// ros_params.cpp:39-77
inline SemanticIntegratorBase::SemanticConfig getSemanticTsdfIntegratorConfigFromParams(const ParamMap& p) {
  SemanticIntegratorBase::SemanticConfig c;
  c.semantic_measurement_probability_ = static_cast<SemanticProbability>(
      params_detail::getDouble(p, "semantic_measurement_probability", c.semantic_measurement_probability_));
  std::string color_mode = "color";
  params_detail::get(p, "semantic_color_mode", &color_mode);
  if (color_mode == "color") {
    c.color_mode = ColorMode::kColor;
  } else if (color_mode == "semantic") {
    c.color_mode = ColorMode::kSemantic;
  } else if (color_mode == "semantic_probability") {
    c.color_mode = ColorMode::kSemanticProbability;
  } else {
    KSG_CHECK(false) << "Unknown semantic color mode: " << color_mode;
  }
  c.semantic_label_to_color_ = std::make_shared<SemanticLabel2Color>(getSemanticLabelToColorCsvFilepathFromParams(p));
  std::string labels;
  KSG_CHECK(params_detail::get(p, "dynamic_semantic_labels", &labels)) << "nh_private.getParam(\"dynamic_semantic_labels\", dynamic_labels)";
  c.dynamic_labels_.clear();
  for (const int l : parseIntList(labels)) c.dynamic_labels_.push_back(static_cast<SemanticLabel>(l));
  return c;
}

// voxblox_ros getTsdfIntegratorConfigFromRosParam: truncation defaults to 4 voxels, everything else to the Config defaults
inline vxb::TsdfIntegratorBase::Config getTsdfIntegratorConfigFromParams(const ParamMap& p) {
  using namespace params_detail;
  vxb::TsdfIntegratorBase::Config c;
  const double voxel_size = getDouble(p, "tsdf_voxel_size", 0.2);
  c.voxel_carving_enabled = getBool(p, "voxel_carving_enabled", true);
  c.default_truncation_distance = static_cast<float>(getDouble(p, "truncation_distance", static_cast<float>(voxel_size) * 4));
  c.max_ray_length_m = static_cast<float>(getDouble(p, "max_ray_length_m", c.max_ray_length_m));
  c.min_ray_length_m = static_cast<float>(getDouble(p, "min_ray_length_m", c.min_ray_length_m));
  c.max_weight = static_cast<float>(getDouble(p, "max_weight", c.max_weight));
  c.use_const_weight = getBool(p, "use_const_weight", c.use_const_weight);
  c.use_weight_dropoff = getBool(p, "use_weight_dropoff", c.use_weight_dropoff);
  c.allow_clear = getBool(p, "allow_clear", c.allow_clear);
  c.start_voxel_subsampling_factor = static_cast<float>(getDouble(p, "start_voxel_subsampling_factor", c.start_voxel_subsampling_factor));
  c.max_consecutive_ray_collisions = getInt(p, "max_consecutive_ray_collisions", c.max_consecutive_ray_collisions);
  c.clear_checks_every_n_frames = getInt(p, "clear_checks_every_n_frames", c.clear_checks_every_n_frames);
  c.max_integration_time_s = static_cast<float>(getDouble(p, "max_integration_time_s", c.max_integration_time_s));
  c.enable_anti_grazing = getBool(p, "anti_grazing", c.enable_anti_grazing);
  c.use_sparsity_compensation_factor = getBool(p, "use_sparsity_compensation_factor", c.use_sparsity_compensation_factor);
  c.sparsity_compensation_factor = static_cast<float>(getDouble(p, "sparsity_compensation_factor", c.sparsity_compensation_factor));
  std::string order = c.integration_order_mode;
  get(p, "integration_order_mode", &order);
  c.integration_order_mode = order;
  c.integrator_threads = static_cast<size_t>(getInt(p, "integrator_threads", static_cast<int>(c.integrator_threads)));
  return c;
}

// voxblox_ros getTsdfMapConfigFromRosParam (tsdf_voxel_size 0.2, tsdf_voxels_per_side 16) + TsdfServer's message throttle
inline SemanticTsdfServer::Params getServerParamsFromParams(const ParamMap& p) {
  SemanticTsdfServer::Params s;
  s.tsdf_voxel_size = static_cast<vxb::FloatingPoint>(params_detail::getDouble(p, "tsdf_voxel_size", 0.2));
  s.tsdf_voxels_per_side = static_cast<size_t>(params_detail::getInt(p, "tsdf_voxels_per_side", 16));
  s.method = getSemanticTsdfIntegratorTypeFromParams(p);
  s.min_time_between_msgs_sec = params_detail::getDouble(p, "min_time_between_msgs_sec", 0.0);
  return s;
}

}  // namespace kimera
