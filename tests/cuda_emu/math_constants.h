// stub of <math_constants.h> for tests/cuda_emu
#pragma once
#include <limits>
#define CUDART_INF (std::numeric_limits<double>::infinity())
