// Stand-in for voxblox/core/color.h (TEST INFRASTRUCTURE): SURVEY.md A.0 / A.6.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <voxblox/core/common.h>

namespace voxblox {

struct Color {
  uint8_t r, g, b, a;
  Color() : r(0), g(0), b(0), a(0) {}
  Color(uint8_t _r, uint8_t _g, uint8_t _b) : Color(_r, _g, _b, 255) {}
  Color(uint8_t _r, uint8_t _g, uint8_t _b, uint8_t _a) : r(_r), g(_g), b(_b), a(_a) {}

  static Color blendTwoColors(const Color& first_color, FloatingPoint first_weight, const Color& second_color,
                              FloatingPoint second_weight) {
    const FloatingPoint total_weight = first_weight + second_weight;
    first_weight /= total_weight;
    second_weight /= total_weight;
    Color out;
    out.r = static_cast<uint8_t>(std::round(first_color.r * first_weight + second_color.r * second_weight));
    out.g = static_cast<uint8_t>(std::round(first_color.g * first_weight + second_color.g * second_weight));
    out.b = static_cast<uint8_t>(std::round(first_color.b * first_weight + second_color.b * second_weight));
    out.a = static_cast<uint8_t>(std::round(first_color.a * first_weight + second_color.a * second_weight));
    return out;
  }

  static const Color White() { return Color(255, 255, 255); }
  static const Color Black() { return Color(0, 0, 0); }
  static const Color Gray() { return Color(127, 127, 127); }
  static const Color Red() { return Color(255, 0, 0); }
  static const Color Green() { return Color(0, 255, 0); }
  static const Color Blue() { return Color(0, 0, 255); }
  static const Color Yellow() { return Color(255, 255, 0); }
  static const Color Orange() { return Color(255, 127, 0); }
  static const Color Purple() { return Color(127, 0, 255); }
  static const Color Teal() { return Color(0, 255, 255); }
  static const Color Pink() { return Color(255, 0, 127); }
};
typedef AlignedVector<Color> Colors;

// HSV wheel with s = v = 1; h is taken modulo 1 and mapped over six 60-degree sectors.
inline Color rainbowColorMap(double h) {
  Color color;
  color.a = 255;
  h -= std::floor(h);
  h *= 6;
  const int sector = (int)std::floor(h);
  double f = h - sector;
  if (!(sector & 1)) f = 1 - f;  // even sectors ramp the other way
  const double m = 0.0;          // v * (1 - s)
  const double n = 1 - f;        // v * (1 - s * f)
  const double v = 1.0;
  double r, g, b;
  switch (sector) {
    case 6:
    case 0: r = v, g = n, b = m; break;
    case 1: r = n, g = v, b = m; break;
    case 2: r = m, g = v, b = n; break;
    case 3: r = m, g = n, b = v; break;
    case 4: r = n, g = m, b = v; break;
    case 5: r = v, g = m, b = n; break;
    default: r = 1, g = 0.5, b = 0.5; break;
  }
  color.r = (uint8_t)(255 * r);
  color.g = (uint8_t)(255 * g);
  color.b = (uint8_t)(255 * b);
  return color;
}

inline Color randomColor() { return Color((uint8_t)(std::rand() % 256), (uint8_t)(std::rand() % 256), (uint8_t)(std::rand() % 256)); }

}  // namespace voxblox
