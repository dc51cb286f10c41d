// rollout_fast.cu — K fused env steps per launch, FP32 / RK4 production mode, two envs per thread on the packed FP32 pipe
// (rollout_pair.cuh).  Own translation unit; quadsim.cu calls launch_rollout_fast.
#include "quadsim_internal.cuh"
#include "step_warp.cuh"
#include "step_pair.cuh"
#include "rollout_pair.cuh"

template <bool DIRECT, bool SENSOR>
static bool launch_fast(qs_sim* s, const qs_rollout_args* a, cudaStream_t st) {
    // no strict in-lane resets; even shard; 8-byte aligned caller buffers; qs_set_step_loader(h, 0..2) selects the one-env-per-thread kernel
    const uintptr_t al = (uintptr_t)a->actions | (uintptr_t)a->obs_out | (uintptr_t)a->action_out | (uintptr_t)a->reward_out |
                         (uintptr_t)a->sensed_obs_out;
    if (s->step_loader != 3 || (s->cfg.flags & QS_FLAG_AUTO_RESET) || (s->N & 1) != 0 || (al & 7) != 0 || ((uintptr_t)a->done_out & 1) != 0)
        return false;
    RolloutIO<float> io{a->horizon, a->action_source, (const float*)a->actions, (float*)a->obs_out, (float*)a->action_out,
                        (float*)a->reward_out, a->done_out, (float*)a->sensed_obs_out};
    constexpr int threads = RolloutPairCfg<SENSOR>::kThreads;
    int64_t g = (s->N / 2 + threads - 1) / threads;
    if (g > (int64_t)s->sm_count * 8) g = (int64_t)s->sm_count * 8;
    constexpr size_t smem = SENSOR ? (size_t)(threads / 32) * qs::kSensorStateDim * 256 : 0;
    if (SENSOR) QS_SET_SMEM_ONCE(s, (rollout_pair_kernel<DIRECT, SENSOR>), smem);
    rollout_pair_kernel<DIRECT, SENSOR><<<(int)g, threads, smem, st>>>(s->pf, make_view<float>(s), io);
    return true;
}

bool launch_rollout_fast(qs_sim* s, const qs_rollout_args* a, bool direct, cudaStream_t st) {
    if (s->cfg.flags & QS_FLAG_SENSOR_NOISE) return direct ? launch_fast<true, true>(s, a, st) : launch_fast<false, true>(s, a, st);
    return direct ? launch_fast<true, false>(s, a, st) : launch_fast<false, false>(s, a, st);
}
