// tests/ros_stubs/ros/ros.h -- the slice of the roscpp API that ros1_shim/*.cpp use, with roscpp's signatures, so
// that the shim can be compiled and linked in an image without ROS (tests/test_ros1_shim.py).  Declarations only
// matter; the inline bodies do nothing.  Message types are the ros_lite structs, whose members carry the names and
// types of the genuine sensor_msgs / std_msgs definitions.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>

#include "d2pc_b200/ros_lite.hpp"

namespace ros {
using Time = ros_lite::Time;
class Subscriber {};
class Publisher {
 public:
  template <class M>
  void publish(const M &) const {}
  uint32_t getNumSubscribers() const { return 0; }
};
class NodeHandle {
 public:
  explicit NodeHandle(const std::string & = std::string()) {}
  template <class M, class T>
  Subscriber subscribe(const std::string &, uint32_t, void (T::*)(const std::shared_ptr<M const> &), T *) {
    return Subscriber();
  }
  template <class M>
  Publisher advertise(const std::string &, uint32_t, bool = false) {
    return Publisher();
  }
  template <class T>
  bool param(const std::string &, T &v, const T &def) const {
    v = def;
    return false;
  }
  bool getParam(const std::string &, int &) const { return false; }
  bool getParam(const std::string &, double &) const { return false; }
};
inline void init(int &, char **, const std::string &) {}
inline void spin() {}
inline void shutdown() {}
}  // namespace ros

#define ROS_WARN(...) std::fprintf(stderr, __VA_ARGS__)
#define ROS_ERROR(...) std::fprintf(stderr, __VA_ARGS__)
#define ROS_FATAL(...) std::fprintf(stderr, __VA_ARGS__)
#define ROS_ERROR_THROTTLE(period, ...) std::fprintf(stderr, __VA_ARGS__)
