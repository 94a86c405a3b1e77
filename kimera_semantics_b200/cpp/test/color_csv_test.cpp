// CPU-only check of SemanticLabel2Color's CSV loading against the reference semantics (color.cpp:42-94).
//   color_csv_test <file.csv>   prints "label r g b a" for labels 0..5 and "rgba -> label" for a few colours
#include <algorithm>
#include <array>
#include <cstdio>
#include <string>
#include <vector>
#include "kimera_semantics/color.h"
using namespace kimera;
int main(int argc, char** argv) {
  if (argc < 2) return 2;
  SemanticLabel2Color lut(argv[1]);
  if (argc > 2 && std::string(argv[2]) == "--dump") {   // both tables in full, in a canonical order
    for (int l = 0; l < 256; ++l) {
      const auto it = lut.semantic_label_to_color_map_.find((SemanticLabel)l);
      if (it != lut.semantic_label_to_color_map_.end()) std::printf("L %d %d %d %d %d\n", l, it->second.r, it->second.g, it->second.b, it->second.a);
    }
    std::vector<std::array<int, 5>> rows;
    for (const auto& kv : lut.color_to_semantic_label_) rows.push_back({{kv.first.r, kv.first.g, kv.first.b, kv.first.a, kv.second}});
    std::sort(rows.begin(), rows.end());
    for (const auto& r : rows) std::printf("C %d %d %d %d %d\n", r[0], r[1], r[2], r[3], r[4]);
    return 0;
  }
  for (int l = 0; l < 6; ++l) {
    const HashableColor c = lut.getColorFromSemanticLabel((SemanticLabel)l);
    std::printf("label %d -> %d %d %d %d\n", l, c.r, c.g, c.b, c.a);
  }
  const HashableColor q[] = {HashableColor(255, 0, 127, 255), HashableColor(255, 0, 0, 255), HashableColor(0, 255, 0, 255),
                             HashableColor(255, 20, 127, 255), HashableColor(255, 255, 255, 255), HashableColor(0, 0, 0, 0),
                             HashableColor(1, 2, 3, 255)};
  for (const HashableColor& c : q) std::printf("color %d %d %d %d -> %d\n", c.r, c.g, c.b, c.a, lut.getSemanticLabelFromColor(c));
  std::printf("entries %zu %zu\n", lut.semantic_label_to_color_map_.size(), lut.color_to_semantic_label_.size());
  return 0;
}
