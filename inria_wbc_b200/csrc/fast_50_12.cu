// k_fast / k_slow specialised for nv = 50, me = 12 (see fast_tu.cuh)
#define IWBC_TU_NV 50
#define IWBC_TU_ME 12
#define IWBC_TU_ENTRY iwbc_launch_fast_50_12
#include "fast_tu.cuh"
