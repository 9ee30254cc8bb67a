// stand-in for <pcl_ros/point_cloud.h>: see ../ref_stubs.hpp (test infrastructure)
#pragma once
#include "../ref_stubs.hpp"
