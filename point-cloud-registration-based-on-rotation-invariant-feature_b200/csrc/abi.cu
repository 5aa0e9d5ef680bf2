// abi.cu — ABI bookkeeping for libri_b200.so (include/ri_b200.h).
#include "ri_common.cuh"

extern "C" int ri_abi_version(void) { return 1; }
