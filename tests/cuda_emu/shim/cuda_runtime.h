// TEST INFRASTRUCTURE: stands in for <cuda_runtime.h> when csrc/common.cuh is compiled by g++ for the CPU emulation.
#pragma once
#include <memory>
#include "../cuda_emu.h"
