// actor_rollout.cu — translation unit of actor_rollout.cuh (see there).
#include "quadsim_internal.cuh"
#include "actor_rollout.cuh"
