// k_coop.cu -- lane-cooperative kernels (added below)
#include "device_common.cuh"
