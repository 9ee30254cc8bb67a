// stand-in for <opencv2/highgui/highgui.hpp>: see ../ref_stubs.hpp (test infrastructure)
#pragma once
#include "../../ref_stubs.hpp"
