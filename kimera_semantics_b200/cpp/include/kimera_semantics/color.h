// Mirror of kimera_semantics/include/kimera_semantics/color.h + src/color.cpp (reference color.cpp:18-94).
#pragma once
#include <string>
#include <unordered_map>
#include "kimera_semantics/common.h"
namespace kimera {
struct HashableColor : public vxb::Color {
  HashableColor() : Color() {}
  HashableColor(const vxb::Color& c) : Color(c) {}
  HashableColor(uint8_t r, uint8_t g, uint8_t b) : HashableColor(r, g, b, 255) {}
  HashableColor(uint8_t r, uint8_t g, uint8_t b, uint8_t a) : Color(r, g, b, a) {}
  bool operator==(const HashableColor& o) const { return r == o.r && g == o.g && b == o.b && a == o.a; }
  bool equal(const HashableColor& o) const { return *this == o; }
};
typedef vxb::AlignedVector<HashableColor> HashableColors;
struct ColorHasher {  // color.cpp:33-40 (alpha not hashed)
  size_t operator()(const HashableColor& k) const {
    return ((std::hash<uint8_t>()(k.r) ^ (std::hash<uint8_t>()(k.g) << 1)) >> 1) ^ (std::hash<uint8_t>()(k.b) << 1);
  }
};
typedef std::unordered_map<HashableColor, SemanticLabel, ColorHasher> ColorToSemanticLabelMap;
typedef std::unordered_map<SemanticLabel, HashableColor> SemanticLabelToColorMap;

class SemanticLabel2Color {
 public:
  // CSV with header name,red,green,blue,alpha,id (color.cpp:42-67)
  explicit SemanticLabel2Color(const std::string& filename);
  // programmatic construction (tests, simulation): label -> colour table
  explicit SemanticLabel2Color(const SemanticLabelToColorMap& label_to_color);
  SemanticLabel getSemanticLabelFromColor(const HashableColor& color) const;
  HashableColor getColorFromSemanticLabel(const SemanticLabel& semantic_label) const;
  ColorToSemanticLabelMap color_to_semantic_label_;
  SemanticLabelToColorMap semantic_label_to_color_map_;
};

// color.h:58-82 of the reference: a colour for every label 0..254, random except the eight fixed simulation colours.
inline SemanticLabelToColorMap getRandomSemanticLabelToColorMap() {
  SemanticLabelToColorMap table;
  for (int label = 0; label < 255; ++label) table[static_cast<SemanticLabel>(label)] = HashableColor(vxb::randomColor());
  const vxb::Color fixed[8] = {vxb::Color::Gray(),  vxb::Color::Green(), vxb::Color::Blue(),   vxb::Color::Purple(),
                               vxb::Color::Pink(),  vxb::Color::Teal(),  vxb::Color::Orange(), vxb::Color::Yellow()};
  for (int label = 0; label < 8; ++label) table.at(static_cast<SemanticLabel>(label)) = HashableColor(fixed[label]);
  return table;
}
}  // namespace kimera
