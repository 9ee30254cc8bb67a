// stand-in for <sensor_msgs/PointCloud2.h>: see ../ros/ros.h
#pragma once
#include "d2pc_b200/ros_lite.hpp"
namespace sensor_msgs {
using PointField = ros_lite::sensor_msgs::PointField;
using PointCloud2 = ros_lite::sensor_msgs::PointCloud2;
}  // namespace sensor_msgs
