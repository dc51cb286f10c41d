// ppo_update.cu — translation unit of ppo_update.cuh (see there).
#include "quadsim_internal.cuh"
#include "ppo_update.cuh"
