#include "common.cuh"
GED_API int ged_version(void) { return 100; }          // 0.1.0
GED_API const char* ged_arch(void) { return "sm_100a"; }
