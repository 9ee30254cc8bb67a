// stand-in for <sensor_msgs/Image.h>: see ../ros/ros.h
#pragma once
#include "d2pc_b200/ros_lite.hpp"
namespace sensor_msgs {
using Image = ros_lite::sensor_msgs::Image;
using ImagePtr = ros_lite::sensor_msgs::ImagePtr;
using ImageConstPtr = ros_lite::sensor_msgs::ImageConstPtr;
}  // namespace sensor_msgs
