// ORACLE — TEST INFRASTRUCTURE ONLY.  Stand-in for the one ROS call information_matrix_calculator.cpp makes: NodeHandle::param<T>(name, default).
#pragma once
#include <map>
#include <string>
namespace ros {
class NodeHandle {
 public:
  std::map<std::string, double> values;      // parameters set by the test; everything else takes the caller's default
  template <typename T>
  T param(const std::string& name, const T& def) const {
    auto it = values.find(name);
    return it == values.end() ? def : static_cast<T>(it->second);
  }
};
}  // namespace ros
