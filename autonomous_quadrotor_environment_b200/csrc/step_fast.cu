// step_fast.cu — the FP32 / RK4 production implementations of qs_step (loaders 2 and 3): per-warp cp.async pipelines with one env
// (step_warp.cuh) or two envs (step_pair.cuh, packed FP32 pipe) per lane.  Own translation unit; quadsim.cu calls launch_step_fast.
#include "quadsim_internal.cuh"
#include "step_warp.cuh"
#include "step_pair.cuh"

template <bool DIRECT, bool SENSOR>
static bool launch_fast(qs_sim* s, const StepIO<float>& io, cudaStream_t st) {
    if (s->cfg.flags & (QS_FLAG_AUX | QS_FLAG_AUTO_RESET)) return false;   // strict lock-step resets / AUX rows keep the CTA-wide kernel
    if (s->step_loader == 3) {
        constexpr int threads = SENSOR ? pr::kThreadsSensor : pr::kThreadsPlain;
        constexpr int warps = (SENSOR ? pr::kConsumerThreadsSensor : pr::kThreadsPlain) / 32;     // the warps that own chunks
        constexpr size_t smem = SENSOR ? pr::kSmemSensor : pr::kSmemPlain;
        QS_SET_SMEM_ONCE(s, (step_kernel_pair<DIRECT, SENSOR>), smem);
        const int64_t chunks = (s->slice_count + 63) / 64;
        int64_t g = (int64_t)s->sm_count;
        const int64_t need = (chunks + warps - 1) / warps;
        if (g > need) g = need;
        step_kernel_pair<DIRECT, SENSOR><<<(int)(g < 1 ? 1 : g), threads, smem, st>>>(s->pf, make_view<float>(s), io);
        return true;
    }
    if (s->step_loader >= 2) {
        constexpr size_t smem = (size_t)(SENSOR ? wp::kRowsSensor : wp::kRowsPlain) * 128 * wp::kStagesW * (kBlock / 32);
        QS_SET_SMEM_ONCE(s, (step_kernel_warp<DIRECT, SENSOR>), smem);
        const int64_t chunks = (s->slice_count + 31) / 32;
        int64_t g = (int64_t)s->sm_count * wp::kMinCtas;
        const int64_t need = (chunks + kBlock / 32 - 1) / (kBlock / 32);
        if (g > need) g = need;
        step_kernel_warp<DIRECT, SENSOR><<<(int)(g < 1 ? 1 : g), kBlock, smem, st>>>(s->pf, make_view<float>(s), io);
        return true;
    }
    return false;
}

bool launch_step_fast(qs_sim* s, const StepIO<float>& io, bool direct, bool sensor, cudaStream_t st) {
    if (direct) return sensor ? launch_fast<true, true>(s, io, st) : launch_fast<true, false>(s, io, st);
    return sensor ? launch_fast<false, true>(s, io, st) : launch_fast<false, false>(s, io, st);
}
