// k_fast / k_slow specialised for nv = 38, me = 12 (see fast_tu.cuh)
#define IWBC_TU_NV 38
#define IWBC_TU_ME 12
#define IWBC_TU_ENTRY iwbc_launch_fast_38_12
#include "fast_tu.cuh"
