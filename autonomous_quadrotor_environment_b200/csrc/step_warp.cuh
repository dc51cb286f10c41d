// step_warp.cuh — quad.step for every env of the shard, FP32 production kernel (loader 2, the default).
//
// Every WARP runs its own software pipeline over 32-env chunks; warps never synchronise with each other
// inside the loop (the earlier CTA-wide TMA ring lost 16 % of all warp time to fast warps spinning on the
// `full` barrier until the issuing warp caught up, and 14 % at the end-of-kernel barriers — profiles/).
//
//   prefetch(chunk j+1)      cp.async (LDGSTS) into the warp's other smem stage: the 4-byte SoA rows of the handle
//                            form ONE [rows][ld] matrix, so a 16-byte cp.async per lane moves FOUR row segments
//                            (8 lanes x 16 B = one 128-byte row segment of 32 envs) — 6 instructions fetch the 21
//                            input rows of a chunk, one more the 4 action rows, one the flag bytes
//   wait(chunk j)            cp.async.wait_group + __syncwarp
//   registers <- smem        conflict-free LDS, immediate offsets (lane = env)
//   quad.step                step_core<> (shared with every other kernel)
//   smem <- results          STS in place, __syncwarp
//   HBM <- smem              LDS.128 + STG.128: 7 vector stores write the 26 output rows; the flag/done/solved
//                            bytes of the 32 envs go out as 16-byte pieces as well
// so a warp issues ~16 global memory instructions per chunk instead of ~60 scalar ones plus their 64-bit
// address arithmetic, and loads are always one whole chunk time (~5 us) ahead of their use.
//
// Resets (QS_FLAG_ASYNC_RESET).  A lane whose env finished pushes the env index on a per-CTA shared-memory
// queue AFTER its results are stored.  Whenever 32 entries are available, the next warp that finishes a chunk
// claims them and re-samples 32 envs with all lanes busy (Philox reset sampling is ~450 instructions; done
// in the finishing lane it would cost every second warp that much for one or two lanes).  Leftovers are drained
// after the CTA's last chunk.  Re-sampling needs nothing of the old episode except its counter.
#pragma once

#ifndef QS_WARP_STAGES            // shared-memory stages per warp: 2 = prefetch one chunk ahead; 1 = in-place, more resident warps
#define QS_WARP_STAGES 2
#endif
#ifndef QS_WARP_MIN_CTAS
#define QS_WARP_MIN_CTAS QS_MIN_CTAS
#endif

namespace wp {
constexpr int kStagesW = QS_WARP_STAGES;
constexpr int kMinCtas = QS_WARP_MIN_CTAS;

constexpr int kRowAct = 26;          // stage rows 26..29: the 4 action rows (in)
constexpr int kRowBytes = 30;        // stage row 30: bytes [0,32) flags (in/out), [32,64) done, [64,96) solved (out)
constexpr int kRowSensor = 31;       // stage rows 31..50: sensor_state (in/out, SENSOR only)
constexpr int kRowsPlain = 31;
constexpr int kRowsSensor = 51;
constexpr int kQueueCapPlain = 2048;   // entries of the per-CTA reset queue (claimed 32 at a time while the kernel runs;
constexpr int kQueueCapSensor = 1024;  // pushes beyond the capacity are re-sampled in their own lane)

// matrix rows (4-byte SoA rows starting at obs17, stride ld) — fixed by the row table in quadsim.cu (checked at create)
constexpr int kMAng = 17, kMShaping = 20, kMAbsSum = 21, kMEpRet = 22, kMStepI = 23, kMEpisode = 24, kMReward = 25;

__device__ __forceinline__ void cp16(uint32_t s, const void* g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(g) : "memory");
}
__device__ __forceinline__ void cp8(uint32_t s, const void* g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(g) : "memory");
}
__device__ __forceinline__ void cp4(uint32_t s, const void* g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

typedef float Row[32];

struct Ext {                 // which caller-provided arrays can take 16-byte vector accesses (uniform per launch)
    bool act_vec, obs_vec, rew_vec, done_vec, solved_vec;
};

// issue the loads of one chunk (envs n0 .. n0+31) into stage st; every lane executes the same number of commits
template <bool SENSOR>
__device__ __forceinline__ void prefetch(const SimView<float>& v, const float* __restrict__ action, const Ext& x,
                                         int64_t n0, Row* st, int lane) {
    const int r4 = lane >> 3, c4 = (lane & 7) << 2;        // 16-byte pieces: 4 rows x 8 lanes
    const int r2 = lane >> 4, c2 = (lane & 15) << 1;       //  8-byte pieces: 2 rows x 16 lanes
    const int64_t ld = v.ld;
    const float* g4 = v.obs17 + (int64_t)r4 * ld + n0 + c4;
    const uint32_t s4 = smem_u32(&st[r4][c4]);
    cp16(s4, g4);                                           // rows 0..3    x vx y vy
    cp16(s4 + 4 * 128, g4 + 4 * ld);                        // rows 4..7    z vz q0 q1
    cp8(smem_u32(&st[8 + r2][c2]), v.obs17 + (int64_t)(8 + r2) * ld + n0 + c2);   // rows 8,9  q2 q3   (10..13 = V_q: output only)
    cp16(s4 + 14 * 128, g4 + 14 * ld);                      // rows 14..17  wx wy wz ang0
    cp16(s4 + 18 * 128, g4 + 18 * ld);                      // rows 18..21  ang1 ang2 prev_shaping abs_sum
    cp16(s4 + 22 * 128, g4 + 22 * ld);                      // rows 22..25  ep_return step_i episode (reward: unused)
    if (lane < 8) cp4(smem_u32(reinterpret_cast<unsigned char*>(st[kRowBytes]) + 4 * lane), v.flags + n0 + 4 * lane);
    const bool full = n0 + 32 <= v.N;
    if (x.act_vec && full) {
        cp16(s4 + kRowAct * 128, action + (int64_t)r4 * v.N + n0 + c4);
    } else if (n0 + lane < v.N) {
#pragma unroll
        for (int k = 0; k < 4; ++k) cp4(smem_u32(&st[kRowAct + k][lane]), action + (int64_t)k * v.N + n0 + lane);
    }
    if (SENSOR) {
        const float* gs = v.sensor_state + (int64_t)r4 * ld + n0 + c4;
#pragma unroll
        for (int k = 0; k < 5; ++k) cp16(s4 + (kRowSensor + 4 * k) * 128, gs + (int64_t)(4 * k) * ld);
    }
    cp_commit();
}

// rows [R0, R0+4*G) of the stage -> 4-byte matrix at g (row stride ld elements), 16 bytes per lane
template <int G>
__device__ __forceinline__ void store_rows4(const Row* st, int R0, float* g, int64_t ld, int lane) {
    const int r4 = lane >> 3, c4 = (lane & 7) << 2;
    float* gl = g + (int64_t)r4 * ld + c4;
#pragma unroll
    for (int k = 0; k < G; ++k) {
        const float4 x = *reinterpret_cast<const float4*>(&st[R0 + 4 * k + r4][c4]);
        *reinterpret_cast<float4*>(gl + (int64_t)(4 * k) * ld) = x;
    }
}
// rows R0, R0+1 of the stage -> matrix at g, 8 bytes per lane
__device__ __forceinline__ void store_rows2(const Row* st, int R0, float* g, int64_t ld, int lane) {
    const int r2 = lane >> 4, c2 = (lane & 15) << 1;
    const float2 x = *reinterpret_cast<const float2*>(&st[R0 + r2][c2]);
    *reinterpret_cast<float2*>(g + (int64_t)r2 * ld + c2) = x;
}

// re-sample env n (QS_FLAG_ASYNC_RESET): what the finishing step of quad.step + the head of quad.reset leave behind
template <bool SENSOR>
__device__ __forceinline__ void resample_env(const DevParams<float>& p, const SimView<float>& v, const StepIO<float>& io, int64_t n) {
    Env<float> e;
    e.episode = __ldcg(v.episode + n);
    float vq[4];
    async_resample(p, v.seed, v.env_id_offset + (uint32_t)n, e, vq);
    const int64_t ld = v.ld;
    float* m = v.obs17 + n;
#pragma unroll
    for (int k = 0; k < 10; ++k) m[k * ld] = e.y[k];
#pragma unroll
    for (int k = 0; k < 4; ++k) m[(10 + k) * ld] = vq[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) m[(14 + k) * ld] = e.y[10 + k];
    v.abs_sum[n] = 0.f;
    v.ep_return[n] = 0.f;
    v.step_i[n] = 0;
    v.episode[n] = e.episode;
    v.flags[n] = (uint8_t)e.flags;
    if (SENSOR) {
        float* so = v.sensed_obs + n;
#pragma unroll
        for (int k = 0; k < 10; ++k) so[k * ld] = e.y[k];
#pragma unroll
        for (int k = 0; k < 4; ++k) so[(10 + k) * ld] = vq[k];
    }
    if (io.obs) {
        float* oo = io.obs + n;
#pragma unroll
        for (int k = 0; k < 10; ++k) oo[k * v.N] = e.y[k];
#pragma unroll
        for (int k = 0; k < 4; ++k) oo[(10 + k) * v.N] = vq[k];
    }
}

}  // namespace wp


// bytes [b*32, b*32+32) of the stage's byte row -> dst[n0 .. n0+32) as two 16-byte pieces (lanes l0, l0+1)
__device__ __forceinline__ void wp_store_bytes(const unsigned char* sb, int b, uint8_t* dst, int64_t n0, int lane, int l0) {
    if (lane == l0 || lane == l0 + 1) {
        const int half = lane - l0;
        const uint4 q = *reinterpret_cast<const uint4*>(sb + b * 32 + half * 16);
        *reinterpret_cast<uint4*>(dst + n0 + half * 16) = q;
    }
}

template <bool DIRECT, bool SENSOR>
__global__ void __launch_bounds__(kBlock, QS_WARP_MIN_CTAS)
step_kernel_warp(const __grid_constant__ DevParams<float> p, const __grid_constant__ SimView<float> v,
                 const __grid_constant__ StepIO<float> io) {
    using namespace wp;
    constexpr int kRows = SENSOR ? kRowsSensor : kRowsPlain;
    constexpr int kQueueCap = SENSOR ? kQueueCapSensor : kQueueCapPlain;
    constexpr int kWarps = kBlock / 32;
    constexpr unsigned kFull = 0xffffffffu;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint32_t s_queue[kQueueCap];
    __shared__ int s_qn, s_qhead;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    Row* ring = reinterpret_cast<Row*>(smem_raw) + (size_t)w * kStagesW * kRows;
    for (int i = tid; i < kQueueCap; i += kBlock) s_queue[i] = 0xFFFFFFFFu;
    if (tid == 0) { s_qn = 0; s_qhead = 0; }
    __syncthreads();

    Ext x;
    const bool n4 = (v.N & 3) == 0, n16 = (v.N & 15) == 0;
    x.act_vec = n4 && ((reinterpret_cast<uintptr_t>(io.action) & 15) == 0);
    x.obs_vec = n4 && ((reinterpret_cast<uintptr_t>(io.obs) & 15) == 0);
    x.rew_vec = n4 && ((reinterpret_cast<uintptr_t>(io.reward) & 15) == 0);
    x.done_vec = n16 && ((reinterpret_cast<uintptr_t>(io.done) & 15) == 0);
    x.solved_vec = n16 && ((reinterpret_cast<uintptr_t>(io.solved) & 15) == 0);
    const bool async_reset = (p.flags & F_ASYNC_RESET) != 0;
    const int64_t ld = v.ld;

    LocalStats ls;
    ls.clear();
    bool any_end = false;
    const int64_t n_chunks = (v.N + 31) >> 5;
    const int64_t stride = (int64_t)gridDim.x * kWarps;
    int64_t c = (int64_t)blockIdx.x * kWarps + w;     // CTA-major (warp-major, which helps the pair kernel's tail, costs 2 % here)
    if (c < n_chunks) prefetch<SENSOR>(v, io.action, x, c << 5, ring, lane);
    for (int j = 0; c < n_chunks; c += stride, ++j) {
        Row* st = ring + (size_t)(kStagesW == 2 ? (j & 1) : 0) * kRows;
        const int64_t cn = c + stride;
        if (kStagesW == 2 && cn < n_chunks) {  // the other stage was drained by the previous iteration's store phase
            prefetch<SENSOR>(v, io.action, x, cn << 5, ring + (size_t)((j + 1) & 1) * kRows, lane);
            cp_wait<1>();
        } else {
            cp_wait<0>();
        }
        __syncwarp();
        const int64_t n0 = c << 5;
        const int64_t n = n0 + lane;
        const bool active = n < v.N;
        const bool full = n0 + 32 <= v.N;
        unsigned char* sb = reinterpret_cast<unsigned char*>(st[kRowBytes]);
        bool push = false;
        float sobs[14];
        if (active) {
            Env<float> e;
#pragma unroll
            for (int k = 0; k < 10; ++k) e.y[k] = st[k][lane];
#pragma unroll
            for (int k = 0; k < 3; ++k) e.y[10 + k] = st[14 + k][lane];
#pragma unroll
            for (int k = 0; k < 3; ++k) e.prev_ang[k] = st[kMAng + k][lane];
            e.prev_shaping = st[kMShaping][lane];
            e.abs_sum = st[kMAbsSum][lane];
            e.ep_return = st[kMEpRet][lane];
            e.i = __float_as_int(st[kMStepI][lane]);
            e.episode = __float_as_uint(st[kMEpisode][lane]);
            e.flags = sb[lane];
            float a[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) a[k] = st[kRowAct + k][lane];
            bool warm = false;
            if (async_reset) warm = async_warmup_prologue(p, e, a);
            const bool was_done = (e.flags & EF_DONE) != 0;
            StepOut<float> o;
            Ctrl<float> ctl;
            step_core<float, 0, DIRECT>(p, e, a, o, &ctl);
            if (warm) o.reward = 0.f; else e.ep_return += o.reward;
            if (o.done && !was_done) { count_episode(ls, p, e, o); any_end = true; }
            push = async_reset && o.done;
            // ---- results -> stage, in place
#pragma unroll
            for (int k = 0; k < 10; ++k) st[k][lane] = e.y[k];
#pragma unroll
            for (int k = 0; k < 4; ++k) st[10 + k][lane] = o.vq[k];
#pragma unroll
            for (int k = 0; k < 3; ++k) st[14 + k][lane] = e.y[10 + k];
#pragma unroll
            for (int k = 0; k < 3; ++k) st[kMAng + k][lane] = e.prev_ang[k];
            st[kMShaping][lane] = e.prev_shaping;
            st[kMAbsSum][lane] = e.abs_sum;
            st[kMEpRet][lane] = e.ep_return;
            st[kMStepI][lane] = __int_as_float(e.i);
            st[kMEpisode][lane] = __uint_as_float(e.episode);
            st[kMReward][lane] = o.reward;
            sb[lane] = (unsigned char)e.flags;
            sb[32 + lane] = (unsigned char)((o.done ? 1 : 0) | (warm ? 2 : 0));
            sb[64 + lane] = (unsigned char)(o.solved ? 1 : 0);
            if (SENSOR) {
                // warm-up steps bypass the sensor; the last one re-initialises it (sensor.reset)
                const int mode = warm ? ((e.flags >> EF_WARM_SHIFT) ? 2 : 1) : 0;
                if (mode == 0) {
                    float s[kSensorStateDim], z[32], rot[9], acc_read[3];
#pragma unroll
                    for (int k = 0; k < 17; ++k) s[k] = st[kRowSensor + k][lane];
                    accel_read(p, ctl, e.y, rot, acc_read);     // trailing drone_eq call: rotation matrix, accelerometer reading
                    sensor_normals(v.seed, v.env_id_offset + (uint32_t)n, e.episode, (uint32_t)e.i, p.s_gps_blend > 0.f, z);
                    sensor_step(p, z, e.y, acc_read, rot, ctl.f_m, s, sobs);
#pragma unroll
                    for (int k = 0; k < kSensorStateDim; ++k) st[kRowSensor + k][lane] = s[k];
                } else {
                    if (mode == 1) {
                        float s[kSensorStateDim];
                        sensor_reset(p, v.seed, v.env_id_offset + (uint32_t)n, e.episode, e.y, s);
#pragma unroll
                        for (int k = 0; k < kSensorStateDim; ++k) st[kRowSensor + k][lane] = s[k];
                    }
#pragma unroll
                    for (int k = 0; k < 10; ++k) sobs[k] = e.y[k];
#pragma unroll
                    for (int k = 0; k < 4; ++k) sobs[10 + k] = o.vq[k];
                }
            }
        }
        __syncwarp();
        // ---- stage -> HBM: the 26 matrix rows, the three byte rows, the sensor state; then the caller's arrays
        store_rows4<6>(st, 0, v.obs17 + n0, ld, lane);                     // rows 0..23
        store_rows2(st, 24, v.obs17 + 24 * ld + n0, ld, lane);             // episode, reward
        if (lane < 6) {
            const int b = lane >> 1, half = lane & 1;
            const uint4 q = *reinterpret_cast<const uint4*>(sb + b * 32 + half * 16);
            uint8_t* dst = (b == 0 ? v.flags : (b == 1 ? v.done : v.solved)) + n0 + half * 16;
            *reinterpret_cast<uint4*>(dst) = q;
        }
        if (SENSOR) store_rows4<5>(st, kRowSensor, v.sensor_state + n0, ld, lane);
        if (io.obs) {
            if (x.obs_vec && full) {
                store_rows4<3>(st, 0, io.obs + n0, v.N, lane);
                store_rows2(st, 12, io.obs + 12 * v.N + n0, v.N, lane);
            } else if (active) {
#pragma unroll
                for (int k = 0; k < 14; ++k) io.obs[k * v.N + n] = st[k][lane];
            }
        }
        if (io.reward) {
            if (x.rew_vec && full) {
                if (lane < 8) *reinterpret_cast<float4*>(io.reward + n0 + 4 * lane) = *reinterpret_cast<const float4*>(&st[kMReward][4 * lane]);
            } else if (active) {
                io.reward[n] = st[kMReward][lane];
            }
        }
        if (io.done) {
            if (x.done_vec && full) wp_store_bytes(sb, 1, io.done, n0, lane, 8);
            else if (active) io.done[n] = sb[32 + lane];
        }
        if (io.solved) {
            if (x.solved_vec && full) wp_store_bytes(sb, 2, io.solved, n0, lane, 10);
            else if (active) io.solved[n] = sb[64 + lane];
        }
        __syncwarp();
        if (SENSOR) {                          // second pass: the sensed observation goes out through rows 0..13
            if (active) {
#pragma unroll
                for (int k = 0; k < 14; ++k) st[k][lane] = sobs[k];
            }
            __syncwarp();
            store_rows4<3>(st, 0, v.sensed_obs + n0, ld, lane);
            store_rows2(st, 12, v.sensed_obs + 12 * ld + n0, ld, lane);
            __syncwarp();
        }
        // single stage: the stage was drained by the stores above; the next chunk's loads overlap the reset work
        // of this warp and the arithmetic of the CTA's other warps
        if (kStagesW == 1 && cn < n_chunks) prefetch<SENSOR>(v, io.action, x, cn << 5, st, lane);
        // ---- resets: push finished envs, claim 32 queued ones if available
        if (async_reset) {
            if (__any_sync(kFull, push)) {
                __threadfence_block();         // this warp's stores of the finished envs precede the re-sampler's
                if (push) {
                    const int slot = atomicAdd(&s_qn, 1);
                    if (slot < kQueueCap) *reinterpret_cast<volatile uint32_t*>(&s_queue[slot]) = (uint32_t)n;
                    else resample_env<SENSOR>(p, v, io, n);            // queue exhausted (e.g. a whole shard timing out at once)
                }
            }
            int take = -1;
            if (lane == 0) {
                const int head = *reinterpret_cast<volatile int*>(&s_qhead);
                int qn = *reinterpret_cast<volatile int*>(&s_qn);
                qn = qn < kQueueCap ? qn : kQueueCap;
                if (qn - head >= 32 && atomicCAS(&s_qhead, head, head + 32) == head) take = head;
            }
            take = __shfl_sync(kFull, take, 0);
            if (take >= 0) {
                uint32_t ent;
                do { ent = *reinterpret_cast<volatile uint32_t*>(&s_queue[take + lane]); } while (ent == 0xFFFFFFFFu);
                __threadfence_block();
                resample_env<SENSOR>(p, v, io, (int64_t)ent);
            }
        }
    }
    __syncthreads();                           // every push of this CTA has been made
    if (async_reset) {
        const int head = s_qhead;
        const int qn = s_qn < kQueueCap ? s_qn : kQueueCap;
        __threadfence_block();
        for (int q = head + tid; q < qn; q += kBlock) resample_env<SENSOR>(p, v, io, (int64_t)s_queue[q]);
    }
    flush_stats(ls, any_end, v.stats);
    if (blockIdx.x == 0 && tid == 0) atomicAdd(&v.stats[7], (double)v.N);
}
