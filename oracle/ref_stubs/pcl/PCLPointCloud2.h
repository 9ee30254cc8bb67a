// stand-in for <pcl/PCLPointCloud2.h>: see ../ref_stubs.hpp (test infrastructure)
#pragma once
#include "../ref_stubs.hpp"
