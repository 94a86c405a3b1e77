// oracle/_ref/libks_ref_hybrid.so -- TEST INFRASTRUCTURE ONLY (same rules as ks_oracle.cpp: only tests/ may load it).
//
// What this is: the reference's OWN kimera_semantics translation units
//     src/semantic_tsdf_integrator_fast.cpp, src/semantic_tsdf_integrator_merged.cpp,
//     src/semantic_integrator_base.cpp, src/color.cpp, src/csv_iterator.cpp      (+ their headers)
// compiled unmodified from /root/reference (never copied into this repository; see oracle/Makefile target `ref`)
// against the stand-in dependency headers in oracle/ref_stubs/ (Eigen, glog, voxblox, minkindr are not in the image
// and voxblox is not vendored by the reference).  So the Kimera half of the algorithm (frame drivers, colour->label
// lookup, dynamic-label filter, start/observed set usage, early termination, bundle merge, anti-grazing, label
// histogram, log-likelihood matrix, arg-max, colour modes, temp-block handling) is the real thing; the voxblox half
// (RayCaster, updateTsdfVoxel, ApproxHashSet, ThreadSafeIndex, bundleRays, Layer/Block, hashes) is a second restatement
// of SURVEY.md Appendix A, independent of ks_oracle.cpp.  It therefore pins the oracle's restatement of the in-tree
// reference files, not voxblox: DESIGN.md keeps saying "parity unpinned" for the voxblox half.
//
// This file is the only code of ours in the library: a C wrapper with the kso_* export layout so that
// tests/parity_utils.compare_maps can diff reference-hybrid vs oracle maps directly.
#include <algorithm>
#include <array>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>
#include <string>
#include <unistd.h>

#include "kimera_semantics/semantic_tsdf_integrator_fast.h"
#include "kimera_semantics/semantic_tsdf_integrator_merged.h"

#include "kimera_semantics_ros/ros_params.h"

#include "../include/ksg.h"

namespace {

struct Hybrid {
  ksg_config cfg;
  std::unique_ptr<voxblox::Layer<voxblox::TsdfVoxel>> tsdf;
  std::unique_ptr<voxblox::Layer<kimera::SemanticVoxel>> sem;
  std::unique_ptr<voxblox::TsdfIntegratorBase> integrator;
  kimera::SemanticIntegratorBase* semantic_base = nullptr;   // the same object seen through its second base class
  double last_integrate_seconds = 0.0;
};

std::vector<voxblox::BlockIndex> sortedBlocks(const Hybrid& h) {
  voxblox::BlockIndexList all;
  h.tsdf->getAllAllocatedBlocks(&all);
  std::vector<voxblox::BlockIndex> v(all.begin(), all.end());
  std::sort(v.begin(), v.end(), [](const voxblox::BlockIndex& a, const voxblox::BlockIndex& b) {
    if (a.z() != b.z()) return a.z() < b.z();
    if (a.y() != b.y()) return a.y() < b.y();
    return a.x() < b.x();
  });
  return v;
}

}  // namespace

extern "C" {

int kref_num_labels() { return (int)kimera::kTotalNumberOfLabels; }

void* kref_create(const ksg_config* c) {
  if (!c || c->num_labels != (int)kimera::kTotalNumberOfLabels) return nullptr;  // compile-time constant in the reference
  std::unique_ptr<Hybrid> h(new Hybrid);
  h->cfg = *c;
  h->tsdf.reset(new voxblox::Layer<voxblox::TsdfVoxel>(c->voxel_size, (size_t)c->voxels_per_side));
  h->sem.reset(new voxblox::Layer<kimera::SemanticVoxel>(c->voxel_size, (size_t)c->voxels_per_side));

  voxblox::TsdfIntegratorBase::Config tc;
  tc.default_truncation_distance = c->default_truncation_distance;
  tc.max_weight = c->max_weight;
  tc.voxel_carving_enabled = c->voxel_carving_enabled != 0;
  tc.min_ray_length_m = c->min_ray_length_m;
  tc.max_ray_length_m = c->max_ray_length_m;
  tc.use_const_weight = c->use_const_weight != 0;
  tc.allow_clear = c->allow_clear != 0;
  tc.use_weight_dropoff = c->use_weight_dropoff != 0;
  tc.use_sparsity_compensation_factor = c->use_sparsity_compensation_factor != 0;
  tc.sparsity_compensation_factor = c->sparsity_compensation_factor;
  tc.integrator_threads = (size_t)std::max(1, c->integrator_threads);
  tc.integration_order_mode = c->integration_order_mode == KSG_ORDER_SORTED ? "sorted" : "mixed";
  tc.enable_anti_grazing = c->enable_anti_grazing != 0;
  tc.start_voxel_subsampling_factor = c->start_voxel_subsampling_factor;
  tc.max_consecutive_ray_collisions = c->max_consecutive_ray_collisions;
  tc.clear_checks_every_n_frames = c->clear_checks_every_n_frames;

  // SemanticLabel2Color only reads a CSV file (color.cpp:42-67): write the config's label table as one.
  char path[] = "/tmp/kref_labels_XXXXXX";
  const int fd = mkstemp(path);
  if (fd < 0) return nullptr;
  {
    std::string csv;
    for (int l = 0; l < 256; ++l) {
      if (!c->label_color_known[l]) continue;
      char row[96];
      std::snprintf(row, sizeof(row), "label_%d,%d,%d,%d,%d,%d\n", l, c->label_color[l][0], c->label_color[l][1], c->label_color[l][2],
                    c->label_color[l][3], l);
      csv += row;
    }
    if (write(fd, csv.data(), csv.size()) != (ssize_t)csv.size()) { close(fd); unlink(path); return nullptr; }
    close(fd);
  }
  kimera::SemanticIntegratorBase::SemanticConfig sc;
  sc.semantic_measurement_probability_ = c->semantic_measurement_probability;
  sc.color_mode = static_cast<kimera::ColorMode>(c->color_mode);
  sc.semantic_label_to_color_ = std::make_shared<kimera::SemanticLabel2Color>(std::string(path));
  unlink(path);
  for (int l = 0; l < 256; ++l)
    if (c->dynamic_label[l]) sc.dynamic_labels_.push_back((kimera::SemanticLabel)l);

  if (c->integrator_type == KSG_INTEGRATOR_FAST) {
    auto* p = new kimera::FastSemanticTsdfIntegrator(tc, sc, h->tsdf.get(), h->sem.get());
    h->integrator.reset(p);
    h->semantic_base = p;
  } else {
    auto* p = new kimera::MergedSemanticTsdfIntegrator(tc, sc, h->tsdf.get(), h->sem.get());
    h->integrator.reset(p);
    h->semantic_base = p;
  }
  return h.release();
}

void kref_destroy(void* hh) { delete (Hybrid*)hh; }

// The reference boundary (fast.cpp:145-149 / merged.cpp:65-69): T_G_C, points_C, colors, freespace.  rgba must be given.
int kref_integrate_points(void* hh, const float* T, const float* xyz, const uint8_t* rgba, int64_t n, int freespace) {
  Hybrid* h = (Hybrid*)hh;
  if (!h || !rgba) return KSG_ERR_INVALID_ARGUMENT;
  const voxblox::Transformation T_G_C(T[0], T[1], T[2], T[3], voxblox::Point(T[4], T[5], T[6]));
  voxblox::Pointcloud points((size_t)n);
  voxblox::Colors colors((size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    points[i] = voxblox::Point(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    colors[i] = voxblox::Color(rgba[4 * i], rgba[4 * i + 1], rgba[4 * i + 2], rgba[4 * i + 3]);
  }
  const auto t0 = std::chrono::steady_clock::now();
  h->integrator->integratePointCloud(T_G_C, points, colors, freespace != 0);
  h->last_integrate_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return 0;
}

// Wall time of the last integratePointCloud call alone (the span of the reference's "integrate/fast" +
// "inserting_missed_blocks" / "semantic_tsdf/integrate" timers).
double kref_last_integrate_seconds(void* hh) { return ((Hybrid*)hh)->last_integrate_seconds; }

// The reference's public per-vector helpers (base.cpp:283-314, 317-350, 370-380), for checking the shim's copies of them.
void kref_update_probabilities(void* hh, const float* frequencies, float* prior_inout) {
  kimera::SemanticProbabilities f, p;
  for (size_t i = 0; i < kimera::kTotalNumberOfLabels; ++i) { f[i] = frequencies[i]; p[i] = prior_inout[i]; }
  ((Hybrid*)hh)->semantic_base->updateSemanticVoxelProbabilities(f, &p);
  for (size_t i = 0; i < kimera::kTotalNumberOfLabels; ++i) prior_inout[i] = p[i];
}
void kref_normalize_probabilities(void* hh, float* probs_inout) {   // aborts (CHECK) exactly where the reference does
  kimera::SemanticProbabilities p;
  for (size_t i = 0; i < kimera::kTotalNumberOfLabels; ++i) p[i] = probs_inout[i];
  ((Hybrid*)hh)->semantic_base->normalizeProbabilities(&p);
  for (size_t i = 0; i < kimera::kTotalNumberOfLabels; ++i) probs_inout[i] = p[i];
}
void kref_label_color(void* hh, int label, uint8_t* rgba) {
  kimera::HashableColor c;
  ((Hybrid*)hh)->semantic_base->updateSemanticVoxelColor((kimera::SemanticLabel)label, &c);
  rgba[0] = c.r; rgba[1] = c.g; rgba[2] = c.b; rgba[3] = c.a;
}
void kref_log_likelihood(void* hh, float* matrix_row_major, float* log_match, float* log_non_match) {
  const kimera::SemanticIntegratorBase* b = ((Hybrid*)hh)->semantic_base;
  const size_t C = kimera::kTotalNumberOfLabels;
  for (size_t i = 0; i < C; ++i)
    for (size_t j = 0; j < C; ++j) matrix_row_major[i * C + j] = b->semantic_log_likelihood_(i, j);
  *log_match = b->log_match_probability_;
  *log_non_match = b->log_non_match_probability_;
}

// SemanticLabel2Color's two tables for a CSV file, parsed by the reference's own reader (color.cpp:42-67, csv_iterator.cpp), in the
// text format of the shim's `color_csv_test <file> --dump`.  Returns the number of characters written (or needed).
int64_t kref_csv_dump(const char* path, char* out, int64_t capacity) {
  const kimera::SemanticLabel2Color lut{std::string(path)};
  std::string text;
  char row[64];
  for (int l = 0; l < 256; ++l) {
    const auto it = lut.semantic_label_to_color_map_.find((kimera::SemanticLabel)l);
    if (it == lut.semantic_label_to_color_map_.end()) continue;
    std::snprintf(row, sizeof(row), "L %d %d %d %d %d\n", l, it->second.r, it->second.g, it->second.b, it->second.a);
    text += row;
  }
  std::vector<std::array<int, 5>> rows;
  for (const auto& kv : lut.color_to_semantic_label_) rows.push_back({{kv.first.r, kv.first.g, kv.first.b, kv.first.a, kv.second}});
  std::sort(rows.begin(), rows.end());
  for (const auto& r : rows) {
    std::snprintf(row, sizeof(row), "C %d %d %d %d %d\n", r[0], r[1], r[2], r[3], r[4]);
    text += row;
  }
  if ((int64_t)text.size() <= capacity && out) std::memcpy(out, text.data(), text.size());
  return (int64_t)text.size();
}

// The reference's own parameter reading (kimera_semantics_ros/src/ros_params.cpp:20-77) on a "key: value" text, one pair per line.
// Writes "method=<m>\ncsv=<path>\nprobability=<%.9g>\ncolor_mode=<int>\ndynamic=<a,b,...>\n"; aborts where the reference aborts.
int64_t kref_ros_params(const char* text, char* out, int64_t capacity) {
  ros::NodeHandle nh;
  std::stringstream ss{std::string(text)};
  std::string line;
  while (std::getline(ss, line)) {
    const size_t colon = line.find(':');
    if (colon == std::string::npos) continue;
    auto trim = [](std::string s) {
      const size_t b = s.find_first_not_of(" \t"), e = s.find_last_not_of(" \t");
      return b == std::string::npos ? std::string() : s.substr(b, e - b + 1);
    };
    nh.values[trim(line.substr(0, colon))] = trim(line.substr(colon + 1));
  }
  const std::string method = kimera::getSemanticTsdfIntegratorTypeFromRosParam(nh);
  const std::string csv = kimera::getSemanticLabelToColorCsvFilepathFromRosParam(nh);
  const kimera::SemanticIntegratorBase::SemanticConfig sc = kimera::getSemanticTsdfIntegratorConfigFromRosParam(nh);
  char num[64];
  std::snprintf(num, sizeof(num), "%.9g", (double)sc.semantic_measurement_probability_);
  std::string r = "method=" + method + "\ncsv=" + csv + "\nprobability=" + num + "\ncolor_mode=" + std::to_string((int)sc.color_mode) + "\ndynamic=";
  for (size_t i = 0; i < sc.dynamic_labels_.size(); ++i) r += (i ? "," : "") + std::to_string((int)sc.dynamic_labels_[i]);
  r += "\nlabels=" + std::to_string(sc.semantic_label_to_color_->semantic_label_to_color_map_.size()) + "\n";
  if ((int64_t)r.size() <= capacity && out) std::memcpy(out, r.data(), r.size());
  return (int64_t)r.size();
}

int64_t kref_num_blocks(void* hh) { return (int64_t)((Hybrid*)hh)->tsdf->getNumberOfAllocatedBlocks(); }
int64_t kref_num_semantic_blocks(void* hh) { return (int64_t)((Hybrid*)hh)->sem->getNumberOfAllocatedBlocks(); }

int kref_export_blocks(void* hh, int64_t capacity, int32_t* block_index, float* tsdf_distance, float* tsdf_weight, uint8_t* tsdf_rgba,
                       uint8_t* sem_label, float* sem_priors, uint8_t* sem_rgba) {
  Hybrid* h = (Hybrid*)hh;
  const std::vector<voxblox::BlockIndex> idx = sortedBlocks(*h);
  if ((int64_t)idx.size() > capacity) return KSG_ERR_INVALID_ARGUMENT;
  const size_t vps = (size_t)h->cfg.voxels_per_side, V = vps * vps * vps;
  const size_t C = kimera::kTotalNumberOfLabels;
  for (size_t b = 0; b < idx.size(); ++b) {
    const voxblox::Block<voxblox::TsdfVoxel>::Ptr tb = h->tsdf->getBlockPtrByIndex(idx[b]);
    const voxblox::Block<kimera::SemanticVoxel>::Ptr sb = h->sem->getBlockPtrByIndex(idx[b]);
    if (!sb) return KSG_ERR_INVALID_ARGUMENT;  // both layers are allocated in lock step (fast.cpp:125-132)
    if (block_index) { block_index[3 * b] = idx[b].x(); block_index[3 * b + 1] = idx[b].y(); block_index[3 * b + 2] = idx[b].z(); }
    for (size_t v = 0; v < V; ++v) {
      const voxblox::TsdfVoxel& t = tb->getVoxelByLinearIndex(v);
      const kimera::SemanticVoxel& s = sb->getVoxelByLinearIndex(v);
      if (tsdf_distance) tsdf_distance[b * V + v] = t.distance;
      if (tsdf_weight) tsdf_weight[b * V + v] = t.weight;
      if (tsdf_rgba) { uint8_t* o = &tsdf_rgba[(b * V + v) * 4]; o[0] = t.color.r; o[1] = t.color.g; o[2] = t.color.b; o[3] = t.color.a; }
      if (sem_label) sem_label[b * V + v] = s.semantic_label;
      if (sem_rgba) { uint8_t* o = &sem_rgba[(b * V + v) * 4]; o[0] = s.color.r; o[1] = s.color.g; o[2] = s.color.b; o[3] = s.color.a; }
      if (sem_priors) std::memcpy(&sem_priors[(b * V + v) * C], s.semantic_priors.data(), C * sizeof(float));
    }
  }
  return 0;
}

}  // extern "C"
