// stand-in for <opencv2/calib3d/calib3d.hpp>: see ../ref_stubs.hpp (test infrastructure)
#pragma once
#include "../../ref_stubs.hpp"
