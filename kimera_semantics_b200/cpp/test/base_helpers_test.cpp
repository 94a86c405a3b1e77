// Exercises the host-side per-vector helpers of the shim's SemanticIntegratorBase (the reference's public THREAD SAFE
// utilities, base.cpp:283-380) on vectors read from a file, so that tests can compare them with the reference's own
// implementations.   base_helpers_test in.bin out.bin
//   in : int32 n, int32 n_palette, palette n*(r,g,b,a,id), float p, then n * { float prior[C], float freq[C] }
//   out: float log_match, float log_non_match, float L[C*C] (row major), then per case
//        { float updated[C], u8 label, u8 rgba[4], float normalized[C] }
#include <cstdio>
#include <fstream>
#include "kimera_semantics/semantic_integrator_base.h"

using namespace kimera;
template <typename T> static T rd(std::ifstream& f) { T v; f.read(reinterpret_cast<char*>(&v), sizeof(T)); return v; }

int main(int argc, char** argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: base_helpers_test in.bin out.bin\n"); return 2; }
  std::ifstream f(argv[1], std::ios::binary);
  KSG_CHECK(f.good());
  const int n = rd<int32_t>(f);
  SemanticLabelToColorMap pal;
  const int n_pal = rd<int32_t>(f);
  for (int i = 0; i < n_pal; ++i) { uint8_t e[5]; f.read(reinterpret_cast<char*>(e), 5); pal[e[4]] = HashableColor(e[0], e[1], e[2], e[3]); }
  SemanticIntegratorBase::SemanticConfig sc;
  sc.semantic_measurement_probability_ = rd<float>(f);
  sc.semantic_label_to_color_ = std::make_shared<SemanticLabel2Color>(pal);
  vxb::Layer<SemanticVoxel> layer(0.1f, 16u);
  const SemanticIntegratorBase base(sc, &layer);   // host-only object: no device is involved in these helpers
  std::ofstream o(argv[2], std::ios::binary);
  o.write(reinterpret_cast<const char*>(&base.log_match_probability_), 4);
  o.write(reinterpret_cast<const char*>(&base.log_non_match_probability_), 4);
  for (size_t i = 0; i < kTotalNumberOfLabels; ++i)
    for (size_t j = 0; j < kTotalNumberOfLabels; ++j) o.write(reinterpret_cast<const char*>(&base.semantic_log_likelihood_(i, j)), 4);
  for (int k = 0; k < n; ++k) {
    SemanticProbabilities prior, freq;
    f.read(reinterpret_cast<char*>(prior.data()), 4 * kTotalNumberOfLabels);
    f.read(reinterpret_cast<char*>(freq.data()), 4 * kTotalNumberOfLabels);
    base.updateSemanticVoxelProbabilities(freq, &prior);
    o.write(reinterpret_cast<const char*>(prior.data()), 4 * kTotalNumberOfLabels);
    SemanticLabel label;
    base.calculateMaximumLikelihoodLabel(prior, &label);
    HashableColor color;
    base.updateSemanticVoxelColor(label, &color);
    const uint8_t lc[5] = {label, color.r, color.g, color.b, color.a};
    o.write(reinterpret_cast<const char*>(lc), 5);
    base.normalizeProbabilities(&prior);
    o.write(reinterpret_cast<const char*>(prior.data()), 4 * kTotalNumberOfLabels);
  }
  std::printf("base helpers ok: %d cases\n", n);
  return 0;
}
