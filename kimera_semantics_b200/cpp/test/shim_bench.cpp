// End-to-end throughput of the drop-in C++ classes, the way a kimera_semantics caller uses them:
//   SemanticTsdfIntegratorFactory::create(method, config, semantic_config, tsdf_layer, semantic_layer)   (factory.h:71-93)
//   integrator->integratePointCloud(T_G_C, points_C, colors)                                             (fast.h:82-86)
// with host std::vector clouds (back-projected outside the timed span, as the ROS front end does).  eager = the reference's contract:
// when the call returns the HOST layers hold the frame's result; lazy = the host layers are refreshed once at the end.
//   shim_bench <fast|merged> <frames.bin> <warmup> <eager|lazy>          (frames.bin: the format of shim_demo)
// prints one JSON line.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <vector>
#include "kimera_semantics/semantic_tsdf_integrator_factory.h"
#include "kimera_semantics/semantic_tsdf_integrator_fast.h"
#include "kimera_semantics/semantic_tsdf_integrator_merged.h"
#include "kimera_semantics/gpu_integrator_core.h"

using namespace kimera;
template <typename T> static T rd(std::ifstream& f) { T v; f.read(reinterpret_cast<char*>(&v), sizeof(T)); return v; }
struct Frame { float T[7]; vxb::Pointcloud pts; vxb::Colors cols; };

int main(int argc, char** argv) {
  if (argc < 5) { std::fprintf(stderr, "usage: shim_bench <fast|merged> frames.bin warmup <eager|lazy>\n"); return 2; }
  std::ifstream f(argv[2], std::ios::binary);
  KSG_CHECK(f.good()) << "cannot open " << argv[2];
  const int n_frames = rd<int32_t>(f);
  const float voxel_size = rd<float>(f);
  const int vps = rd<int32_t>(f);
  SemanticLabelToColorMap pal;
  const int n_pal = rd<int32_t>(f);
  for (int i = 0; i < n_pal; ++i) { uint8_t e[5]; f.read(reinterpret_cast<char*>(e), 5); pal[e[4]] = HashableColor(e[0], e[1], e[2], e[3]); }
  SemanticIntegratorBase::SemanticConfig sc;
  sc.semantic_label_to_color_ = std::make_shared<SemanticLabel2Color>(pal);
  const int n_dyn = rd<int32_t>(f);
  for (int i = 0; i < n_dyn; ++i) sc.dynamic_labels_.push_back(rd<uint8_t>(f));
  vxb::TsdfIntegratorBase::Config config;
  config.default_truncation_distance = 4.0f * voxel_size;  // voxblox_ros
  const int warmup = std::atoi(argv[3]);
  const bool lazy = std::strcmp(argv[4], "lazy") == 0;
  std::vector<Frame> frames((size_t)n_frames);
  for (int fr = 0; fr < n_frames; ++fr) {
    const int n = rd<int32_t>(f);
    f.read(reinterpret_cast<char*>(frames[fr].T), sizeof(frames[fr].T));
    frames[fr].pts.resize(n); frames[fr].cols.resize(n);
    f.read(reinterpret_cast<char*>(frames[fr].pts.data()), sizeof(float) * 3 * n);
    f.read(reinterpret_cast<char*>(frames[fr].cols.data()), 4 * (size_t)n);
  }
  vxb::Layer<vxb::TsdfVoxel> tsdf_layer(voxel_size, vps);
  vxb::Layer<SemanticVoxel> semantic_layer(voxel_size, vps);
  std::unique_ptr<vxb::TsdfIntegratorBase> integrator =
      SemanticTsdfIntegratorFactory::create(argv[1], config, sc, &tsdf_layer, &semantic_layer);
  GpuIntegratorCore* core = nullptr;
  if (auto* fi = dynamic_cast<FastSemanticTsdfIntegrator*>(integrator.get())) core = &fi->gpu();
  else if (auto* mi = dynamic_cast<MergedSemanticTsdfIntegrator*>(integrator.get())) core = &mi->gpu();
  KSG_CHECK(core != nullptr);
  core->setLayerSyncMode(lazy ? LayerSyncMode::kLazy : LayerSyncMode::kEager);
  long long updates = 0;
  std::chrono::steady_clock::time_point t0;
  for (int fr = 0; fr < n_frames; ++fr) {
    if (fr == warmup) t0 = std::chrono::steady_clock::now();
    const Frame& F = frames[fr];
    integrator->integratePointCloud(vxb::Transformation(F.T[0], F.T[1], F.T[2], F.T[3], vxb::Point(F.T[4], F.T[5], F.T[6])), F.pts, F.cols);
    if (fr >= warmup) updates += core->lastVoxelUpdates();
  }
  if (lazy) core->syncLayers();      // the one refresh of the host layers belongs to the measured span
  const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  const int timed = n_frames - warmup;
  std::printf("{\"method\": \"%s\", \"layer_sync\": \"%s\", \"frames\": %d, \"seconds\": %.6f, \"fps\": %.3f, \"voxel_updates\": %lld, \"host_blocks\": %zu}\n",
              argv[1], lazy ? "lazy" : "eager", timed, sec, timed / sec, updates, tsdf_layer.getNumberOfAllocatedBlocks());
  return 0;
}
