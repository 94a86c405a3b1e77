// Stand-in for voxblox/utils/timing.h (TEST INFRASTRUCTURE): timers are no-ops.
#pragma once
#include <string>

namespace voxblox {
namespace timing {
class Timer {
 public:
  explicit Timer(const std::string&, bool = false) {}
  void Start() {}
  void Stop() {}
};
}  // namespace timing
}  // namespace voxblox
