// stand-in for <cv_bridge/cv_bridge.h>: see ../ref_stubs.hpp (test infrastructure)
#pragma once
#include "../ref_stubs.hpp"
