// Stand-in for <ros/ros.h> (TEST INFRASTRUCTURE, see oracle/ref_stubs/README.md): just enough of ros::NodeHandle for the
// reference's kimera_semantics_ros/src/ros_params.cpp to compile - param() with a default and getParam() for an int list, both
// served from a string map the test fills in.
#pragma once
#include <cstdlib>
#include <map>
#include <sstream>
#include <string>
#include <vector>

namespace ros {

class NodeHandle {
 public:
  std::map<std::string, std::string> values;

  bool param(const std::string& name, std::string& out, const std::string& fallback) const {
    const auto it = values.find(name);
    out = it == values.end() ? fallback : it->second;
    return it != values.end();
  }
  bool param(const std::string& name, double& out, const double& fallback) const {
    const auto it = values.find(name);
    out = it == values.end() ? fallback : std::atof(it->second.c_str());
    return it != values.end();
  }
  bool getParam(const std::string& name, std::vector<int>& out) const {
    const auto it = values.find(name);
    if (it == values.end()) return false;
    std::string s = it->second;
    for (char& c : s)
      if (c == '[' || c == ']' || c == ',') c = ' ';
    std::stringstream ss(s);
    out.clear();
    int v;
    while (ss >> v) out.push_back(v);
    return true;
  }
};

}  // namespace ros
