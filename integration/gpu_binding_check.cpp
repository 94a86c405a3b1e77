// Driver for integration/kimera_semantics/semantic_tsdf_integrator_gpu.h: the reference's REAL headers and color.cpp /
// csv_iterator.cpp / semantic_integrator_base.cpp + the binding + libksg.so.  Reads the frame file format of the shim demo
// (kimera_semantics_b200/cpp/test/shim_demo.cpp), integrates every frame through GpuSemanticTsdfIntegrator and writes both host
// layers in the shim demo's output format, so tests can compare it with the oracle.   usage: gpu_binding_check fast|merged in out
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <unistd.h>

#include "kimera_semantics/semantic_tsdf_integrator_gpu.h"

namespace {
template <typename T>
T rd(std::ifstream& f) { T v; f.read(reinterpret_cast<char*>(&v), sizeof(T)); return v; }
}  // namespace

int main(int argc, char** argv) {
  if (argc < 4) { std::fprintf(stderr, "usage: %s fast|merged frames.bin out.bin\n", argv[0]); return 2; }
  const std::string method = argv[1];
  std::ifstream f(argv[2], std::ios::binary);
  CHECK(f.good()) << "cannot open " << argv[2];
  const int32_t n_frames = rd<int32_t>(f);
  const float voxel_size = rd<float>(f);
  const int32_t vps = rd<int32_t>(f);
  const int32_t n_labels = rd<int32_t>(f);
  // label table -> the CSV file SemanticLabel2Color reads (color.cpp:42-67)
  char path[] = "/tmp/gpu_binding_labels_XXXXXX";
  const int fd = mkstemp(path);
  CHECK_GE(fd, 0);
  std::string csv;
  for (int i = 0; i < n_labels; ++i) {
    uint8_t e[5];
    f.read(reinterpret_cast<char*>(e), 5);
    char row[96];
    std::snprintf(row, sizeof(row), "label_%d,%d,%d,%d,%d,%d\n", e[4], e[0], e[1], e[2], e[3], e[4]);
    csv += row;
  }
  CHECK_EQ(write(fd, csv.data(), csv.size()), static_cast<ssize_t>(csv.size()));
  close(fd);
  kimera::SemanticIntegratorBase::SemanticConfig sc;
  sc.semantic_label_to_color_ = std::make_shared<kimera::SemanticLabel2Color>(std::string(path));
  unlink(path);
  const int32_t n_dyn = rd<int32_t>(f);
  for (int i = 0; i < n_dyn; ++i) sc.dynamic_labels_.push_back(rd<uint8_t>(f));

  voxblox::Layer<voxblox::TsdfVoxel> tsdf(voxel_size, static_cast<size_t>(vps));
  voxblox::Layer<kimera::SemanticVoxel> sem(voxel_size, static_cast<size_t>(vps));
  voxblox::TsdfIntegratorBase::Config cfg;
  cfg.default_truncation_distance = 4.0f * voxel_size;   // ROS default of the reference's launch files
  cfg.integrator_threads = 1;
  CHECK(method == "fast" || method == "merged") << "Unknown TSDF integrator type: " << method;
  kimera::GpuSemanticTsdfIntegrator integrator(method == "fast" ? KSG_INTEGRATOR_FAST : KSG_INTEGRATOR_MERGED, cfg, sc, &tsdf, &sem);

  for (int fr = 0; fr < n_frames; ++fr) {
    const int32_t n = rd<int32_t>(f);
    float T[7];
    f.read(reinterpret_cast<char*>(T), sizeof(T));
    voxblox::Pointcloud pts(static_cast<size_t>(n));
    voxblox::Colors cols(static_cast<size_t>(n));
    for (int i = 0; i < n; ++i) { float p[3]; f.read(reinterpret_cast<char*>(p), 12); pts[i] = voxblox::Point(p[0], p[1], p[2]); }
    for (int i = 0; i < n; ++i) { uint8_t c[4]; f.read(reinterpret_cast<char*>(c), 4); cols[i] = voxblox::Color(c[0], c[1], c[2], c[3]); }
    integrator.integratePointCloud(voxblox::Transformation(T[0], T[1], T[2], T[3], voxblox::Point(T[4], T[5], T[6])), pts, cols);
  }

  // output: int32 nb, then per block (sorted z,y,x): int32 index[3], V records {f32 d, f32 w, u8 rgba[4], u8 label, f32 priors[C], u8 srgba[4]}
  voxblox::BlockIndexList all;
  tsdf.getAllAllocatedBlocks(&all);
  std::vector<voxblox::BlockIndex> blocks(all.begin(), all.end());
  std::sort(blocks.begin(), blocks.end(), [](const voxblox::BlockIndex& a, const voxblox::BlockIndex& b) {
    if (a.z() != b.z()) return a.z() < b.z();
    if (a.y() != b.y()) return a.y() < b.y();
    return a.x() < b.x();
  });
  std::ofstream o(argv[3], std::ios::binary);
  const int32_t nb = static_cast<int32_t>(blocks.size());
  o.write(reinterpret_cast<const char*>(&nb), 4);
  const size_t V = static_cast<size_t>(vps) * vps * vps;
  for (const voxblox::BlockIndex& bi : blocks) {
    const int32_t idx[3] = {bi.x(), bi.y(), bi.z()};
    o.write(reinterpret_cast<const char*>(idx), 12);
    const auto tb = tsdf.getBlockPtrByIndex(bi);
    const auto sb = sem.getBlockPtrByIndex(bi);
    CHECK(sb) << "semantic block missing";
    for (size_t v = 0; v < V; ++v) {
      const voxblox::TsdfVoxel& t = tb->getVoxelByLinearIndex(v);
      const kimera::SemanticVoxel& s = sb->getVoxelByLinearIndex(v);
      o.write(reinterpret_cast<const char*>(&t.distance), 4);
      o.write(reinterpret_cast<const char*>(&t.weight), 4);
      const uint8_t c[4] = {t.color.r, t.color.g, t.color.b, t.color.a};
      o.write(reinterpret_cast<const char*>(c), 4);
      o.write(reinterpret_cast<const char*>(&s.semantic_label), 1);
      o.write(reinterpret_cast<const char*>(s.semantic_priors.data()), 4 * kimera::kTotalNumberOfLabels);
      const uint8_t sc4[4] = {s.color.r, s.color.g, s.color.b, s.color.a};
      o.write(reinterpret_cast<const char*>(sc4), 4);
    }
  }
  std::printf("gpu_binding_check: %d frames, %d blocks\n", n_frames, nb);
  return 0;
}
