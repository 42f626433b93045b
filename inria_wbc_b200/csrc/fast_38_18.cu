// k_fast / k_slow specialised for nv = 38, me = 18 (see fast_tu.cuh)
#define IWBC_TU_NV 38
#define IWBC_TU_ME 18
#define IWBC_TU_ENTRY iwbc_launch_fast_38_18
#include "fast_tu.cuh"
