// generic cycle kernels of one group width / thread-per-instance size (see cycle_tu.cuh)
#define IWBC_CYC_G 64
#include "cycle_tu.cuh"
