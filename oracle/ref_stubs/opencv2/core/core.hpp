// stand-in for <opencv2/core/core.hpp>: see ../ref_stubs.hpp (test infrastructure)
#pragma once
#include "../../ref_stubs.hpp"
