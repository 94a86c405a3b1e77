// CPU-only check of SemanticLabel2Color's CSV loading against the reference semantics (color.cpp:42-94).
//   color_csv_test <file.csv>   prints "label r g b a" for labels 0..5 and "rgba -> label" for a few colours
#include <cstdio>
#include "kimera_semantics/color.h"
using namespace kimera;
int main(int argc, char** argv) {
  if (argc < 2) return 2;
  SemanticLabel2Color lut(argv[1]);
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
