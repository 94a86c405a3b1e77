// Host-only round trip of kimera::map_io (checkpoint file of both layers): save -> load into fresh layers -> every voxel equal;
// refusal of foreign / mismatching / truncated files.   map_io_test <tmp dir>
#include <cstdio>
#include <string>
#include "kimera_semantics/map_io.h"
#include "kimera_semantics/vxblx_io.h"

using namespace kimera;

static uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

static void fill(vxb::Layer<vxb::TsdfVoxel>* tsdf, vxb::Layer<SemanticVoxel>* sem, uint32_t seed) {
  const int idx[4][3] = {{0, 0, 0}, {-1, 2, 3}, {5, -7, 1}, {-100000, 99999, -3}};
  for (const auto& i : idx) {
    const vxb::BlockIndex bi(i[0], i[1], i[2]);
    auto tb = tsdf->allocateBlockPtrByIndex(bi);
    auto sb = sem->allocateBlockPtrByIndex(bi);
    for (size_t v = 0; v < tb->num_voxels(); ++v) {
      vxb::TsdfVoxel& tv = tb->getVoxelByLinearIndex(v);
      tv.distance = (float)(lcg(seed) % 2001) * 1e-3f - 1.0f;
      tv.weight = (float)(lcg(seed) % 100000) * 0.1f;
      tv.color = vxb::Color(lcg(seed) & 255, lcg(seed) & 255, lcg(seed) & 255, lcg(seed) & 255);
      SemanticVoxel& sv = sb->getVoxelByLinearIndex(v);
      sv.semantic_label = (SemanticLabel)(lcg(seed) % kTotalNumberOfLabels);
      for (size_t c = 0; c < kTotalNumberOfLabels; ++c) sv.semantic_priors[c] = -(float)(lcg(seed) % 100000) * 1e-3f;
      sv.color = HashableColor(lcg(seed) & 255, lcg(seed) & 255, lcg(seed) & 255, 255);
    }
  }
}

static bool equal(const vxb::Layer<vxb::TsdfVoxel>& a, const vxb::Layer<SemanticVoxel>& as, const vxb::Layer<vxb::TsdfVoxel>& b,
                  const vxb::Layer<SemanticVoxel>& bs) {
  if (a.getNumberOfAllocatedBlocks() != b.getNumberOfAllocatedBlocks() || as.getNumberOfAllocatedBlocks() != bs.getNumberOfAllocatedBlocks()) return false;
  for (const vxb::BlockIndex& bi : map_io::sortedBlocks(a)) {
    auto ta = a.getBlockPtrByIndex(bi), tb = b.getBlockPtrByIndex(bi);
    auto sa = as.getBlockPtrByIndex(bi), sb = bs.getBlockPtrByIndex(bi);
    if (!ta || !tb || !sa || !sb) return false;
    for (size_t v = 0; v < ta->num_voxels(); ++v) {
      const vxb::TsdfVoxel &x = ta->getVoxelByLinearIndex(v), &y = tb->getVoxelByLinearIndex(v);
      if (std::memcmp(&x.distance, &y.distance, 4) || std::memcmp(&x.weight, &y.weight, 4) || x.color.r != y.color.r || x.color.g != y.color.g ||
          x.color.b != y.color.b || x.color.a != y.color.a) return false;
      const SemanticVoxel &p = sa->getVoxelByLinearIndex(v), &q = sb->getVoxelByLinearIndex(v);
      if (p.semantic_label != q.semantic_label || std::memcmp(p.semantic_priors.data(), q.semantic_priors.data(), 4 * kTotalNumberOfLabels) ||
          !(p.color == q.color)) return false;
    }
  }
  return true;
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  const std::string dir = argv[1], file = dir + "/map.ksgm";
  vxb::Layer<vxb::TsdfVoxel> tsdf(0.05f, 16u), tsdf2(0.05f, 16u), tsdf_other(0.10f, 16u), tsdf_vps8(0.05f, 8u);
  vxb::Layer<SemanticVoxel> sem(0.05f, 16u), sem2(0.05f, 16u), sem_other(0.10f, 16u), sem_vps8(0.05f, 8u);
  fill(&tsdf, &sem, 12345u);
  KSG_CHECK(map_io::saveLayers(file, tsdf, sem));
  fill(&tsdf2, &sem2, 999u);                                   // stale contents must be replaced, not merged
  tsdf2.allocateBlockPtrByIndex(vxb::BlockIndex(42, 42, 42));
  sem2.allocateBlockPtrByIndex(vxb::BlockIndex(42, 42, 42));
  KSG_CHECK(map_io::loadLayers(file, &tsdf2, &sem2));
  KSG_CHECK(equal(tsdf, sem, tsdf2, sem2)) << "round trip changed the map";
  KSG_CHECK(!map_io::loadLayers(file, &tsdf_other, &sem_other)) << "voxel size mismatch must be refused";
  KSG_CHECK(!map_io::loadLayers(file, &tsdf_vps8, &sem_vps8)) << "voxels_per_side mismatch must be refused";
  KSG_CHECK(tsdf_other.getNumberOfAllocatedBlocks() == 0u);
  KSG_CHECK(!map_io::loadLayers(dir + "/does_not_exist", &tsdf2, &sem2));
  {  // not a KSGM file
    std::ofstream o((dir + "/junk").c_str(), std::ios::binary);
    o << "VXBLX this is something else entirely";
  }
  KSG_CHECK(!map_io::loadLayers(dir + "/junk", &tsdf2, &sem2));
  KSG_CHECK(equal(tsdf, sem, tsdf2, sem2)) << "a refused file must leave the layers untouched";
  {  // truncated file
    std::ifstream in(file.c_str(), std::ios::binary);
    std::string all((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    std::ofstream o((dir + "/cut").c_str(), std::ios::binary);
    o.write(all.data(), (std::streamsize)(all.size() / 2));
  }
  KSG_CHECK(!map_io::loadLayers(dir + "/cut", &tsdf2, &sem2));
  {  // voxblox .vxblx file of the TSDF layer: round trip, refusals
    const std::string vx = dir + "/tsdf.vxblx";
    tsdf.getBlockPtrByIndex(vxb::BlockIndex(0, 0, 0))->has_data() = true;
    KSG_CHECK(vxblx_io::saveTsdfLayer(vx, tsdf));
    vxb::Layer<vxb::TsdfVoxel> back(0.05f, 16u);
    back.allocateBlockPtrByIndex(vxb::BlockIndex(7, 7, 7));              // replaced, not merged
    KSG_CHECK(vxblx_io::loadTsdfLayer(vx, &back));
    KSG_CHECK(back.getNumberOfAllocatedBlocks() == tsdf.getNumberOfAllocatedBlocks());
    for (const vxb::BlockIndex& bi : map_io::sortedBlocks(tsdf)) {
      auto a = tsdf.getBlockPtrByIndex(bi);
      auto b = back.getBlockPtrByIndex(bi);
      KSG_CHECK(b) << "block lost";
      KSG_CHECK(a->has_data() == b->has_data());
      for (size_t v = 0; v < a->num_voxels(); ++v) {
        const vxb::TsdfVoxel &x = a->getVoxelByLinearIndex(v), &y = b->getVoxelByLinearIndex(v);
        KSG_CHECK(!std::memcmp(&x.distance, &y.distance, 4) && !std::memcmp(&x.weight, &y.weight, 4) && x.color.r == y.color.r &&
                  x.color.g == y.color.g && x.color.b == y.color.b && x.color.a == y.color.a);
      }
    }
    KSG_CHECK(!vxblx_io::loadTsdfLayer(vx, &tsdf_other)) << "voxel size mismatch must be refused";
    KSG_CHECK(!vxblx_io::loadTsdfLayer(vx, &tsdf_vps8));
    KSG_CHECK(!vxblx_io::loadTsdfLayer(dir + "/junk", &back));
    KSG_CHECK(!vxblx_io::loadTsdfLayer(dir + "/cut", &back));
    KSG_CHECK(back.getNumberOfAllocatedBlocks() == tsdf.getNumberOfAllocatedBlocks()) << "a refused file must leave the layer untouched";
  }
  std::printf("map io ok\n");
  return 0;
}
