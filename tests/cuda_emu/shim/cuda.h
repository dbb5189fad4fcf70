// TEST INFRASTRUCTURE: stands in for <cuda.h> (driver API types used for tensor maps) under the CPU emulation.
#pragma once
#include "cuda_runtime.h"
#include "../cuda_emu_sm100.h"
