// stand-in for <sensor_msgs/PointCloud2.h>: see ../ref_stubs.hpp (test infrastructure)
#pragma once
#include "../ref_stubs.hpp"
