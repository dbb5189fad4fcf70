// genesis_b200 -- library-level entry points.
#include "common.cuh"

extern "C" int g2_abi_version(void) { return 1; }
