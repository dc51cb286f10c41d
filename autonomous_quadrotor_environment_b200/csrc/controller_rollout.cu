// controller_rollout.cu — translation unit of controller_rollout.cuh (see there).
#include "quadsim_internal.cuh"
#include "controller_rollout.cuh"
