// stand-in for <pcl_conversions/pcl_conversions.h>: see ../ref_stubs.hpp (test infrastructure)
#pragma once
#include "../ref_stubs.hpp"
