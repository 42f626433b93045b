// k_fast / k_slow specialised for nv = 0, me = 0 (run-time sizes) (see fast_tu.cuh)
#define IWBC_TU_NV 0
#define IWBC_TU_ME 0
#define IWBC_TU_ENTRY iwbc_launch_fast_0_0
#include "fast_tu.cuh"
