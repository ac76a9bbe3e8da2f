// stub of <cuda_runtime.h> for tests/cuda_emu (see emu_core.h)
#pragma once
#include "emu_core.h"
#include "emu_runtime.h"
