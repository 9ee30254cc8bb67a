// stand-in for <sensor_msgs/image_encodings.h>: the three constants the shim names
#pragma once
#include <string>
namespace sensor_msgs {
namespace image_encodings {
const std::string MONO8 = "mono8";
const std::string TYPE_8UC1 = "8UC1";
const std::string RGB8 = "rgb8";
}  // namespace image_encodings
}  // namespace sensor_msgs
