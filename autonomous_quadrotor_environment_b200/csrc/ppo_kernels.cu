// ppo_kernels.cu — translation unit of ppo_kernels.cuh (see there).
#include "quadsim_internal.cuh"
#include "ppo_kernels.cuh"
