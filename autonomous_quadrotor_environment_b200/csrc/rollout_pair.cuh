// rollout_pair.cuh — K fused env steps per launch with TWO envs per thread (FP32 / RK4 production mode).
//
// Same contract as rollout_kernel (quadsim.cu): the env state stays in registers for the whole horizon, actions come from
// a [K][4][N] buffer or are drawn in-kernel (Philox U(-1,1), BASELINE.json configs[2]), outputs are optional.  Thread t owns
// the adjacent pair (2t, 2t+1) as aligned register pairs: the RK4 stages of drone_eq run on FFMA2/FMUL2/FADD2
// (packed_device.cuh), the scalar phases are instantiated per half (two independent chains per thread), loads and stores
// are 8 bytes per thread.  With no state traffic this is the kernel whose FP32 pipe utilisation the metric's
// "% of FP32 FMA roofline" is about.
#pragma once
#include "packed_device.cuh"

#ifndef QS_ROLLOUT_PAIR_THREADS
#define QS_ROLLOUT_PAIR_THREADS 256
#endif
#ifndef QS_ROLLOUT_SENSOR_STREAM
#define QS_ROLLOUT_SENSOR_STREAM 1              // sensor phase of the rollout kernel: 1 = streaming (state rows read / written in place in
#endif                                          // shared memory), 0 = whole state and sensed observation in registers (252 registers)
#ifndef QS_ROLLOUT_PAIR_THREADS_SENSOR
#define QS_ROLLOUT_PAIR_THREADS_SENSOR 256      // 384 (168 registers, 12 warps per SM) measured slower with recording: 96.5 vs 87.8 us
#endif
template <bool SENSOR> struct RolloutPairCfg { static constexpr int kThreads = SENSOR ? QS_ROLLOUT_PAIR_THREADS_SENSOR : QS_ROLLOUT_PAIR_THREADS; };

// Actions read from the caller's [K][4][N] tensor (QS_ACT_BUFFER) are fetched ONE STEP AHEAD: every lane copies the four float2 of
// its pair for step t+1 into its own slots of a two-stage shared-memory buffer with cp.async (8 bytes: the alignment the launcher
// already requires) before it starts on step t, so that the HBM latency of the load sits under a whole step of arithmetic
// instead of in front of it (8 warps per SM cannot hide it): 84.1 -> 78.2 us per step of 1,048,576 envs with the sensor model and
// the recorded stream, 43.0 -> 36.9 without the sensor model (-DQS_ROLLOUT_ACT_PREFETCH=0 is the A/B build;
// profiles/r02_rollout_prefetch_ab.txt).  A lane only ever reads what it copied itself: no warp synchronisation.
#ifndef QS_ROLLOUT_ACT_PREFETCH
#define QS_ROLLOUT_ACT_PREFETCH 1               // 0 = tensor actions loaded where they are consumed (A/B)
#endif
#ifndef QS_ROLLOUT_LOCKSTEP
#define QS_ROLLOUT_LOCKSTEP 0
#endif
#ifndef QS_ROLLOUT_COOP_RESET
#define QS_ROLLOUT_COOP_RESET 1                 // 0 = every finishing env is re-sampled in the lane that owns it (async_resample)
#endif
namespace rp {
struct ResetSlots { float4 out[8][4]; uint2 desc[8]; };        // per warp: the finishing envs of one pass and their sampled states
__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory"); }
}  // namespace rp

// SENSOR (QS_FLAG_SENSOR_NOISE): the packed sensor model of the step kernel (sensor_pair.cuh via pr::sensor_phase) runs after
// every step; the pair's 20 sensor-state rows live in shared memory for the whole horizon (one 256-byte row segment per warp
// and row: lane l holds its pair as one float2, conflict-free LDS.64 / STS.64), so a K-step launch moves the sensor state
// through HBM once instead of K times and the per-step traffic is the recorded stream only (sensed observation, reward, done).
template <bool DIRECT, bool SENSOR>
__global__ void __launch_bounds__(RolloutPairCfg<SENSOR>::kThreads, 1)
rollout_pair_kernel(const __grid_constant__ DevParams<float> p, const __grid_constant__ SimView<float> v,
                    const __grid_constant__ RolloutIO<float> io) {
    using pr::half_of;
    extern __shared__ __align__(16) float s_sensor[];           // SENSOR: [warps][20] rows of 64 floats (dynamic shared memory)
    __shared__ __align__(8) float2 s_act[2][4][RolloutPairCfg<SENSOR>::kThreads];      // QS_ACT_BUFFER: [stage][row][thread]
    const bool act_buf = io.action_source != QS_ACT_PHILOX_UNIFORM;
#if QS_ROLLOUT_COOP_RESET
    __shared__ __align__(16) rp::ResetSlots s_rs[RolloutPairCfg<SENSOR>::kThreads / 32];
#endif
    pr::Row* srows = reinterpret_cast<pr::Row*>(s_sensor) + (threadIdx.x >> 5) * qs::kSensorStateDim;
    const int lane = threadIdx.x & 31;
    LocalStats ls;
    ls.clear();
    bool any_end = false;
    const bool async_reset = (p.flags & F_ASYNC_RESET) != 0;
    const int64_t N2 = v.N >> 1, ld2 = v.ld >> 1;                   // v.N is even (checked by the launcher)
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // Whole warps only: the sensor phase uses warp-wide votes, so every lane of a warp that has work runs the iteration.  Lanes past
    // the last pair (m >= N2) compute on the padding columns of the handle's rows (ld is a multiple of 256 envs) and never touch the
    // caller's unpadded [K][C][N] buffers or the statistics.
    const int64_t N2r = (N2 + 31) & ~(int64_t)31;
#if QS_ROLLOUT_LOCKSTEP
    // tuning build: the warps of a CTA take every step together (see QS_PAIR_LOCKSTEP in step_pair.cuh); measured and rejected:
    // 75.9 -> 79.3 us per step with the sensor model, 36.6 -> 42.0 without.  Warps past the end of the shard only keep the barrier count
    for (int64_t mb = (int64_t)blockIdx.x * blockDim.x; mb < N2r; mb += stride) {
        const int64_t m = mb + threadIdx.x;
        if (m >= N2r) {
            for (int t = 0; t < io.horizon; ++t) __syncthreads();
            continue;
        }
#else
    for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < N2r; m += stride) {
#endif
        const bool live = m < N2;          // (`act` is the clipped-action array below)
        const int64_t nA = 2 * m;
        const float2* g2 = reinterpret_cast<const float2*>(v.obs17) + m;
        P2 y[13];
#pragma unroll
        for (int k = 0; k < 10; ++k) y[k].v = g2[(int64_t)k * ld2];
#pragma unroll
        for (int k = 0; k < 3; ++k) y[10 + k].v = g2[(int64_t)(14 + k) * ld2];
        Env<float> e[2];
        {
            float2 t;
#pragma unroll
            for (int k = 0; k < 3; ++k) { t = g2[(int64_t)(wp::kMAng + k) * ld2]; e[0].prev_ang[k] = t.x; e[1].prev_ang[k] = t.y; }
            t = g2[(int64_t)wp::kMShaping * ld2]; e[0].prev_shaping = t.x; e[1].prev_shaping = t.y;
            t = g2[(int64_t)wp::kMAbsSum * ld2]; e[0].abs_sum = t.x; e[1].abs_sum = t.y;
            t = g2[(int64_t)wp::kMEpRet * ld2]; e[0].ep_return = t.x; e[1].ep_return = t.y;
            t = g2[(int64_t)wp::kMStepI * ld2]; e[0].i = __float_as_int(t.x); e[1].i = __float_as_int(t.y);
            t = g2[(int64_t)wp::kMEpisode * ld2]; e[0].episode = __float_as_uint(t.x); e[1].episode = __float_as_uint(t.y);
            const uint32_t fl = reinterpret_cast<const uint16_t*>(v.flags)[m];
            e[0].flags = fl & 0xffu; e[1].flags = fl >> 8;
        }
        if (SENSOR) {
            const float2* gs2 = reinterpret_cast<const float2*>(v.sensor_state) + m;
#pragma unroll
            for (int k = 0; k < qs::kSensorStateDim; ++k) { const float2 t = gs2[(int64_t)k * ld2]; pr::sts2(srows, k, lane, t.x, t.y); }
        }
        StepOut<float> o[2];
        bool warm[2] = {false, false};
#if QS_ROLLOUT_ACT_PREFETCH
        if (act_buf && live) {
            const float2* at = reinterpret_cast<const float2*>(io.actions) + m;
#pragma unroll
            for (int k = 0; k < 4; ++k) rp::cp_async8(&s_act[0][k][threadIdx.x], at + (int64_t)k * N2);
        }
        rp::cp_async_commit();
#endif
        for (int t = 0; t < io.horizon; ++t) {
#if QS_ROLLOUT_LOCKSTEP
            __syncthreads();
#endif
            float a[2][4], act[2][4];
            Ctrl<float> ctl[2];
            bool was_done[2];
            if (!act_buf) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint4 u = philox_block(v.seed, v.env_id_offset + (uint32_t)(nA + h), e[h].episode, (uint32_t)e[h].i, RNG_ACTION);
                    a[h][0] = 2.f * u32_to_unit<float>(u.x) - 1.f; a[h][1] = 2.f * u32_to_unit<float>(u.y) - 1.f;
                    a[h][2] = 2.f * u32_to_unit<float>(u.z) - 1.f; a[h][3] = 2.f * u32_to_unit<float>(u.w) - 1.f;
                }
            } else {
#if !QS_ROLLOUT_ACT_PREFETCH
                const float2* at0 = reinterpret_cast<const float2*>(io.actions + (int64_t)t * 4 * v.N) + m;
#pragma unroll
                for (int k = 0; k < 4; ++k) { const float2 q = live ? at0[(int64_t)k * N2] : make_float2(0.f, 0.f); a[0][k] = q.x; a[1][k] = q.y; }
#else
                if (live && t + 1 < io.horizon) {                  // step t+1's actions start their way in now
                    const float2* at = reinterpret_cast<const float2*>(io.actions + (int64_t)(t + 1) * 4 * v.N) + m;
#pragma unroll
                    for (int k = 0; k < 4; ++k) rp::cp_async8(&s_act[(t + 1) & 1][k][threadIdx.x], at + (int64_t)k * N2);
                }
                rp::cp_async_commit();
                rp::cp_async_wait<1>();                            // everything but the group just committed has landed: step t's actions
#pragma unroll
                for (int k = 0; k < 4; ++k) { const float2 q = live ? s_act[t & 1][k][threadIdx.x] : make_float2(0.f, 0.f); a[0][k] = q.x; a[1][k] = q.y; }
#endif
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                warm[h] = async_reset ? async_warmup_prologue(p, e[h], a[h]) : false;
                was_done[h] = (e[h].flags & EF_DONE) != 0;
                ctl[h] = step_pre<float, DIRECT>(p, e[h], a[h], o[h], act[h]);
            }
            const Ctrl2 c2 = pack_ctrl(ctl[0], ctl[1]);
#if QS_SENSOR_FUSED_NORMALS
            P2 zpre[SENSOR ? 24 : 1];
            if constexpr (SENSOR) {                 // the sensor model's normals are drawn inside the unrolled RK4 stages (packed_device.cuh)
                SensorRng2 rng;
                rng.seed = v.seed; rng.rk = v.rk;
                rng.id[0] = v.env_id_offset + (uint32_t)nA; rng.id[1] = rng.id[0] + 1u;
                rng.ep[0] = e[0].episode; rng.ep[1] = e[1].episode;
                rng.step[0] = (uint32_t)e[0].i; rng.step[1] = (uint32_t)e[1].i;
                integrate_rk4_2_fused(p, c2, y, [&](int st) {
                    if (st < 3) sensor_normals_block2(rng, st, &zpre[8 * st]);
                });
            } else {
                integrate_rk4_2(p, c2, y);
            }
#else
            integrate_rk4_2(p, c2, y);
#endif
#if QS_PAIR_PACKED_POST
            pr::step_post2(p, y, e, act, o);        // phase 3 on the packed pipe; sets e[.].y; bit-identical to the scalar phase
#endif
#define QS_RP_POST(H)                                                                                         \
            {                                                                                                 \
                if (!QS_PAIR_PACKED_POST) {                                                                   \
                    _Pragma("unroll") for (int k = 0; k < 13; ++k) e[H].y[k] = half_of<H>(y[k]);              \
                    step_post(p, e[H], act[H], o[H]);                                                         \
                }                                                                                             \
                o[H].reward = warm[H] ? 0.f : o[H].reward;                                                    \
                e[H].ep_return += o[H].reward;                                                                \
            }
            QS_RP_POST(0)
            QS_RP_POST(1)
#undef QS_RP_POST
            if (SENSOR) {                // the sensor model sees the state the step left (before a finished env is re-sampled)
                pr::PairMeta pm;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    pm.episode[h] = e[h].episode; pm.step_i[h] = (uint32_t)e[h].i; pm.flags[h] = e[h].flags; pm.warm[h] = warm[h];
                }
                const bool last = (t == io.horizon - 1);
                float2* rec2 = (io.sensed_out && live) ? reinterpret_cast<float2*>(io.sensed_out + (int64_t)t * 14 * v.N) + m : nullptr;
                float2* go2 = reinterpret_cast<float2*>(v.sensed_obs) + m;
#if QS_ROLLOUT_SENSOR_STREAM
                const bool any_warm = __any_sync(0xffffffffu, warm[0] | warm[1]);
                const bool w0 = warm[0], w1 = warm[1];
                pr::sensor_phase_stream(p, v, srows, lane, nA, y, c2.f_m, pm, any_warm, [&](int k, P2 val) {
                    if (any_warm) {          // warm-up steps pass the true observation through
                        const P2 tr = k < 10 ? y[k] : pk(o[0].vq[k - 10], o[1].vq[k - 10]);
                        val = psel(w0, w1, tr, val);
                    }
                    if (rec2) rec2[(int64_t)k * N2] = val.v;
                    if (last) go2[(int64_t)k * ld2] = val.v;
                }
#if QS_SENSOR_FUSED_NORMALS
                , zpre
#endif
                );
#else
                P2 sn[qs::kSensorStateDim], so[14], vq[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) vq[k] = pk(o[0].vq[k], o[1].vq[k]);
#if QS_SENSOR_FUSED_NORMALS
                pr::sensor_phase(p, v, srows, lane, nA, y, vq, c2.f_m, pm, sn, so, zpre);
#else
                pr::sensor_phase(p, v, srows, lane, nA, y, vq, c2.f_m, pm, sn, so);
#endif
#pragma unroll
                for (int k = 0; k < qs::kSensorStateDim; ++k) pr::sts2(srows, k, lane, sn[k].v.x, sn[k].v.y);
#pragma unroll
                for (int k = 0; k < 14; ++k) {
                    if (rec2) rec2[(int64_t)k * N2] = so[k].v;
                    if (last) go2[(int64_t)k * ld2] = so[k].v;
                }
#endif
            }
#if QS_ROLLOUT_COOP_RESET
#pragma unroll
            for (int h = 0; h < 2; ++h)
                if (live && o[h].done && !was_done[h]) { count_episode(ls, p, e[h], o[h]); any_end = true; }
            {
                // Re-sampling by the whole warp.  With ~50-step episodes about one env of a warp's 64 finishes per step: run in the
                // lane that owns it, the re-sampler (four Philox4x32-10 blocks, six Box-Muller pairs, euler_quat) costs ~480
                // instructions for one or two active lanes in 72 % of all warp-steps — a quarter of the kernel's instructions.
                // Here the finishing envs of the warp (both halves) are listed in shared memory, FOUR LANES TAKE ONE ENV — lane b of
                // a group draws Philox block b of quad.reset's stream and finishes its share of the state (block 0: Euler angles ->
                // quaternion; blocks 1-3: the clipped normals) — and the owners read their 13 values back: eight envs per pass,
                // one pass in all but mass time-outs.  Same draws, same arithmetic as sample_reset_state / async_resample (the
                // bit-for-bit rollout == repeated-steps tests cover it).  Measured at 1,048,576 envs, K = 32: 82.3 -> 75.6 us per step
                // with the sensor model (94.7 -> 78.1 with actions from a tensor and the recorded stream), 37.5 -> 36.5 without; most
                // of it is the instruction cache (stall_no_instructions 16 % -> 9 %: two inlined copies of the re-sampler left the loop).
                const bool need0 = async_reset && o[0].done, need1 = async_reset && o[1].done;
                const uint32_t m0 = __ballot_sync(0xffffffffu, need0), m1 = __ballot_sync(0xffffffffu, need1);
                if (m0 | m1) {
                    rp::ResetSlots& rs = s_rs[threadIdx.x >> 5];
                    const uint32_t lt = (1u << lane) - 1u;
                    const int rank[2] = {__popc(m0 & lt), __popc(m0) + __popc(m1 & lt)};
                    const bool need[2] = {need0, need1};
                    const int total = __popc(m0) + __popc(m1);
                    for (int base = 0; base < total; base += 8) {
#pragma unroll
                        for (int h = 0; h < 2; ++h)
                            if (need[h] && (unsigned)(rank[h] - base) < 8u)
                                rs.desc[rank[h] - base] = make_uint2(v.env_id_offset + (uint32_t)(nA + h), e[h].episode + 1u);
                        __syncwarp();
                        const int grp = lane >> 2, blk = lane & 3;
                        if (base + grp < total) {
                            const uint2 d = rs.desc[grp];
                            const uint4 u = philox_block(v.seed, d.x, d.y, (uint32_t)blk, RNG_RESET);
                            float4 r;
                            if (blk == 0) {
                                const float ang[3] = {u32_to_unit<float>(u.x) - 0.5f, u32_to_unit<float>(u.y) - 0.5f, u32_to_unit<float>(u.z) - 0.5f};
                                float q[4];
                                euler_quat(ang, q);
                                r = make_float4(q[0], q[1], q[2], q[3]);
                            } else {
                                float n[4];
                                box_muller(u32_to_unit<float>(u.x), u32_to_unit<float>(u.y), &n[0], &n[1]);
                                box_muller(u32_to_unit<float>(u.z), u32_to_unit<float>(u.w), &n[2], &n[3]);
                                // block 1: x y z vx | block 2: vy vz wx wy | block 3: wz
                                const bool b1 = blk == 1, b2 = blk == 2;
                                const float lo0 = b1 ? -p.pos_clip : b2 ? -p.vel_clip : p.w_clip_lo, hi0 = b1 ? p.pos_clip : b2 ? p.vel_clip : p.w_clip_hi;
                                const float lo2 = b1 ? -p.pos_clip : p.w_clip_lo, hi2 = b1 ? p.pos_clip : p.w_clip_hi;
                                const float lo3 = b1 ? -p.vel_clip : p.w_clip_lo, hi3 = b1 ? p.vel_clip : p.w_clip_hi;
                                r = make_float4(clampr(n[0] * 2.f, lo0, hi0), clampr(n[1] * 2.f, lo0, hi0), clampr(n[2] * 2.f, lo2, hi2),
                                                clampr(n[3] * 2.f, lo3, hi3));
                            }
                            rs.out[grp][blk] = r;
                        }
                        __syncwarp();
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            if (need[h] && (unsigned)(rank[h] - base) < 8u) {
                                const float4* ro = rs.out[rank[h] - base];
                                const float4 q = ro[0], r1 = ro[1], r2 = ro[2], r3 = ro[3];
                                Env<float>& en = e[h];
                                en.y[0] = r1.x; en.y[2] = r1.y; en.y[4] = r1.z; en.y[1] = r1.w;
                                en.y[3] = r2.x; en.y[5] = r2.y; en.y[10] = r2.z; en.y[11] = r2.w; en.y[12] = r3.x;
                                en.y[6] = q.x; en.y[7] = q.y; en.y[8] = q.z; en.y[9] = q.w;
                                en.episode += 1u;
                                en.flags = (uint32_t)p.T << EF_WARM_SHIFT;          // as async_resample
                                en.i = 0; en.abs_sum = 0.f; en.ep_return = 0.f;
                                deriv_quat(&en.y[10], &en.y[6], o[h].vq);
#pragma unroll
                                for (int k = 0; k < 13; ++k) { if (h == 0) y[k].v.x = e[0].y[k]; else y[k].v.y = e[1].y[k]; }
                                if (SENSOR) {        // the sensed observation returned with done is the new episode's initial observation
                                    float* rec = (io.sensed_out && live) ? io.sensed_out + (int64_t)t * 14 * v.N + nA + h : nullptr;
                                    float* go = v.sensed_obs + nA + h;
#pragma unroll
                                    for (int k = 0; k < 14; ++k) {
                                        const float x = k < 10 ? e[h].y[k] : o[h].vq[k - 10];
                                        if (rec) rec[(int64_t)k * v.N] = x;
                                        if (t == io.horizon - 1) go[(int64_t)k * v.ld] = x;
                                    }
                                }
                            }
                        }
                        __syncwarp();
                    }
                }
            }
#else
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (live && o[h].done && !was_done[h]) { count_episode(ls, p, e[h], o[h]); any_end = true; }
                if (async_reset && o[h].done) {
                    async_resample(p, v.seed, v.env_id_offset + (uint32_t)(nA + h), e[h], o[h].vq);
#pragma unroll
                    for (int k = 0; k < 13; ++k) { if (h == 0) y[k].v.x = e[0].y[k]; else y[k].v.y = e[1].y[k]; }
                    if (SENSOR) {        // the sensed observation returned with done is the new episode's initial observation
                        float* rec = (io.sensed_out && live) ? io.sensed_out + (int64_t)t * 14 * v.N + nA + h : nullptr;
                        float* go = v.sensed_obs + nA + h;
#pragma unroll
                        for (int k = 0; k < 14; ++k) {
                            const float x = k < 10 ? e[h].y[k] : o[h].vq[k - 10];
                            if (rec) rec[(int64_t)k * v.N] = x;
                            if (t == io.horizon - 1) go[(int64_t)k * v.ld] = x;
                        }
                    }
                }
            }
#endif
            if (io.obs_out && live) {
                float2* ot = reinterpret_cast<float2*>(io.obs_out + (int64_t)t * 14 * v.N) + m;
#pragma unroll
                for (int k = 0; k < 10; ++k) ot[(int64_t)k * N2] = y[k].v;
#pragma unroll
                for (int k = 0; k < 4; ++k) ot[(int64_t)(10 + k) * N2] = make_float2(o[0].vq[k], o[1].vq[k]);
            }
            if (io.action_out && live) {
                float2* at = reinterpret_cast<float2*>(io.action_out + (int64_t)t * 4 * v.N) + m;
#pragma unroll
                for (int k = 0; k < 4; ++k) at[(int64_t)k * N2] = make_float2(a[0][k], a[1][k]);
            }
            if (io.reward_out && live) reinterpret_cast<float2*>(io.reward_out + (int64_t)t * v.N)[m] = make_float2(o[0].reward, o[1].reward);
            if (io.done_out && live)
                reinterpret_cast<uint16_t*>(io.done_out + (int64_t)t * v.N)[m] =
                    (uint16_t)(((o[0].done ? 1u : 0u) | (warm[0] ? 2u : 0u)) | (((o[1].done ? 1u : 0u) | (warm[1] ? 2u : 0u)) << 8));
        }
        // ---- state back to the handle
        float2* s2 = reinterpret_cast<float2*>(v.obs17) + m;
#pragma unroll
        for (int k = 0; k < 10; ++k) s2[(int64_t)k * ld2] = y[k].v;
#pragma unroll
        for (int k = 0; k < 4; ++k) s2[(int64_t)(10 + k) * ld2] = make_float2(o[0].vq[k], o[1].vq[k]);
#pragma unroll
        for (int k = 0; k < 3; ++k) s2[(int64_t)(14 + k) * ld2] = y[10 + k].v;
#pragma unroll
        for (int k = 0; k < 3; ++k) s2[(int64_t)(wp::kMAng + k) * ld2] = make_float2(e[0].prev_ang[k], e[1].prev_ang[k]);
        s2[(int64_t)wp::kMShaping * ld2] = make_float2(e[0].prev_shaping, e[1].prev_shaping);
        s2[(int64_t)wp::kMAbsSum * ld2] = make_float2(e[0].abs_sum, e[1].abs_sum);
        s2[(int64_t)wp::kMEpRet * ld2] = make_float2(e[0].ep_return, e[1].ep_return);
        s2[(int64_t)wp::kMStepI * ld2] = make_float2(__int_as_float(e[0].i), __int_as_float(e[1].i));
        s2[(int64_t)wp::kMEpisode * ld2] = make_float2(__uint_as_float(e[0].episode), __uint_as_float(e[1].episode));
        s2[(int64_t)wp::kMReward * ld2] = make_float2(o[0].reward, o[1].reward);
        reinterpret_cast<uint16_t*>(v.flags)[m] = (uint16_t)((e[0].flags & 0xffu) | ((e[1].flags & 0xffu) << 8));
        reinterpret_cast<uint16_t*>(v.done)[m] =
            (uint16_t)(((o[0].done ? 1u : 0u) | (warm[0] ? 2u : 0u)) | (((o[1].done ? 1u : 0u) | (warm[1] ? 2u : 0u)) << 8));
        reinterpret_cast<uint16_t*>(v.solved)[m] = (uint16_t)((o[0].solved ? 1u : 0u) | ((o[1].solved ? 1u : 0u) << 8));
        if (SENSOR) {
            float2* gs2 = reinterpret_cast<float2*>(v.sensor_state) + m;
#pragma unroll
            for (int k = 0; k < qs::kSensorStateDim; ++k) gs2[(int64_t)k * ld2] = pr::lds2(srows, k, lane);
        }
    }
    flush_stats(ls, any_end, v.stats);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&v.stats[7], (double)v.N * io.horizon);
}
