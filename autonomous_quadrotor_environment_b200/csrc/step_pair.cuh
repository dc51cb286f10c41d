// step_pair.cuh — quad.step for every env of the shard, FP32 production kernel with TWO environments per lane (loader 3).
//
// A warp owns 64-env chunks and lane l the ADJACENT pair (n0+2l, n0+2l+1):
//   * every quantity of the pair is one aligned 64-bit register pair {A, B}: one LDS.64 reads it from the warp's
//     shared-memory stage, one STG.64 writes it back to its SoA row (a warp-wide STG.64 = 256 contiguous bytes);
//   * the RK4 stages of drone_eq run on FFMA2/FMUL2/FADD2 (packed_device.cuh): half the issue slots per env;
//   * the scalar phases (action map, Euler angles, done/reward) are instantiated once per half: two independent
//     dependency chains per thread where the one-env kernel stalls on its single chain (ncu: 3-5x more warp-time per
//     instruction outside the RK loop than inside), and the per-chunk bookkeeping (addresses, loop, queue) is paid once
//     per 64 envs;
//   * the sensor model runs packed as well (sensor_pair.cuh).
// Pipeline per warp (no synchronisation between warps inside the loop):
//   wait(chunk j) -> LDS the whole chunk into registers -> __syncwarp -> cp.async (LDGSTS, 16 B per lane: two 256-byte row
//   segments per instruction) of chunk j+1 into the SAME stage, which has the whole arithmetic of chunk j to land ->
//   quad.step x2 -> STG.64 of the results straight from registers.
//
// With the sensor model (256 threads, ~240 registers: dynamics + packed sensor model of an env pair in one thread) the
// sensor_state rows travel as their own cp.async group: they are read after the dynamics and re-filled after the sensor
// phase, so they never sit in registers during the RK4 stages.  Measured alternatives (1,048,576 envs, T=5, async resets,
// us per step): this kernel 108-110; one env per lane (step_warp.cuh) 111; warp-specialised variant (6 dynamics + 6 sensor
// warps per CTA at 168 registers, mbarrier hand-off through shared memory) 132 — stall_no_instructions 21 %: twelve warps
// spread over ~96 KB of straight-line code thrash the instruction cache; Philox rounds rolled 115; IMAD.WIDE Philox 114.
// Resets: same per-CTA queue and opportunistic full-warp drains as step_warp.cuh.
#pragma once
#ifndef QS_PAIR_PACKED_POST
#define QS_PAIR_PACKED_POST 1
#endif
#ifndef QS_PAIR_LOCKSTEP
#define QS_PAIR_LOCKSTEP 0
#endif
#include "packed_device.cuh"
#include "sensor_pair.cuh"

namespace pr {

using wp::cp16; using wp::cp4; using wp::cp_commit; using wp::cp_wait; using wp::Ext;

// stage rows (64 floats each): 0..9 = matrix rows 0..9 (x vx y vy z vz q0..q3), 10..20 = matrix rows 14..24 (w, ang,
// prev_shaping, abs_sum, ep_return, step_i, episode), 21..24 = the 4 action rows, 25 = flag bytes [0,64),
// 26..45 = sensor_state (fused sensor mode only)
constexpr int kSW = 10;              // stage row of matrix row 14
constexpr int kRowAct = 21;
constexpr int kRowBytes = 25;
constexpr int kRowSensor = 26;
constexpr int kRowsPlain = 26;
constexpr int kRowsSensor = 46;
#ifndef QS_PAIR_THREADS
#define QS_PAIR_THREADS 384
#endif
#ifndef QS_SENSOR_STREAM
#define QS_SENSOR_STREAM 0          // 1 = streaming sensor phase in the STEP kernel (state rows stay in shared memory, 168 registers,
#endif                              // 12 warps per SM): measured slower, 127.9 vs 101.8 us per step of 1M envs; tuning builds only
#ifndef QS_SENSOR_FUSED_NORMALS
#define QS_SENSOR_FUSED_NORMALS 0   // 1 = draw the sensor model's normals (Philox + Box-Muller) inside the UNROLLED RK4 stages of the same env
#endif                              // pair (one basic block: six Philox chains + MUFU chains next to the FP32 stage arithmetic).  Measured:
                                    // step kernel 106.1 vs 101.6 us (228 registers, 3 more copies of the RHS in the instruction stream),
                                    // rollout kernel 82.6 vs 82.8 us — tuning builds only
#ifndef QS_PAIR_THREADS_SENSOR
#define QS_PAIR_THREADS_SENSOR (QS_SENSOR_STREAM ? 384 : 256)
#endif
#ifndef QS_PAIR_PRODUCER
#define QS_PAIR_PRODUCER 0          // 1 = sensor kernel: an extra warpgroup (4 warps, setmaxnreg 56) draws the sensor model's normals (Philox +
#endif                              // Box-Muller, blocks 0..2 of every env-step) one chunk ahead into shared memory for the chunk warps, whose
                                    // code then needs 161 registers instead of 210.  Measured (1,048,576 envs, us per step, with / without
                                    // resets): inline normals 100.8 / 86.8; 8 chunk warps + 4 producers 105.3 / 84.3; 12 chunk warps at 152
                                    // registers (74 B of spills, single normals buffer) + 4 producers 119.7 / 91.9 — tuning builds only
#ifndef QS_PAIR_SELF_NORMALS
#define QS_PAIR_SELF_NORMALS 0      // 1 = sensor kernel: every chunk warp draws the 24 normals of its chunk right after loading it, while few
#endif                              // registers are live, and parks them in its own rows of shared memory (each lane reads back what it wrote:
                                    // no synchronisation); the sensor phase then runs without the Philox state next to the sensor state:
                                    // 168 registers, 12 warps per SM (-DQS_PAIR_THREADS_SENSOR=384) WITHOUT spills.  Measured (us per step of
                                    // 1,048,576 envs, with / without resets): 8 warps 103.6 / 84.4, 10 warps 108.5 / 91.6, 12 warps 102.5 / 90.8
                                    // against 100.7 / 86.7 of the default — tuning builds only
constexpr int kThreadsPlain = QS_PAIR_THREADS;
constexpr int kConsumerThreadsSensor = QS_PAIR_THREADS_SENSOR;                 // the warps that own chunks
constexpr bool kSelfNormals = QS_PAIR_SELF_NORMALS != 0 && !QS_PAIR_PRODUCER && !QS_SENSOR_STREAM && !QS_SENSOR_FUSED_NORMALS;
constexpr bool kProducer = QS_PAIR_PRODUCER != 0 && !QS_SENSOR_STREAM && !QS_SENSOR_FUSED_NORMALS &&
                           (QS_PAIR_THREADS_SENSOR == 256 || QS_PAIR_THREADS_SENSOR == 384);
constexpr int kProducerWarps = kProducer ? 4 : 0;                              // one warpgroup (setmaxnreg works per warpgroup)
constexpr int kServed = (kConsumerThreadsSensor / 32) / 4;                     // chunk warps per producer warp: j serves kServed*j ...
constexpr int kThreadsSensor = kConsumerThreadsSensor + 32 * kProducerWarps;
constexpr int kZRows = 24;                                                     // normals per env-step handed over (blocks 0..2)
constexpr int kZBufs = QS_PAIR_THREADS_SENSOR == 256 ? 2 : 1;                  // normals buffers per chunk warp (shared memory: 227 KB)
constexpr int kConsumerRegs = QS_PAIR_THREADS_SENSOR == 256 ? 216 : 152;       // 8 x 32 x 216 (12 x 32 x 152) + 4 x 32 x 56 <= 65,536
constexpr int kQueueCapPlain = 2048;
constexpr int kQueueCapSensor = 1024;
constexpr size_t kSmemPlain = (size_t)26 * 256 * (kThreadsPlain / 32);
// sensor kernel: per chunk warp 46 stage rows (+ 2 x 24 rows of pre-drawn normals, double-buffered, in producer mode)
constexpr size_t kSmemSensor = (size_t)(46 + (kProducer ? kZBufs * kZRows : (kSelfNormals ? kZRows : 0))) * 256 * (kConsumerThreadsSensor / 32);

typedef float Row[64];

template <int H> __device__ __forceinline__ float half_of(const qs::P2& x) { return H == 0 ? x.v.x : x.v.y; }
template <int H> __device__ __forceinline__ float half_of(const float2& x) { return H == 0 ? x.x : x.y; }

__device__ __forceinline__ float2 lds2(const Row* st, int row, int lane) { return reinterpret_cast<const float2*>(st[row])[lane]; }
__device__ __forceinline__ void sts2(Row* st, int row, int lane, float a, float b) { reinterpret_cast<float2*>(st[row])[lane] = make_float2(a, b); }

// issue the loads of one chunk (envs n0 .. n0+63) into the warp's stage; every lane executes the same number of commits
__device__ __forceinline__ void prefetch(const SimView<float>& v, const float* __restrict__ action, const Ext& x,
                                         int64_t n0, Row* st, int lane) {
    const int r2 = lane >> 4, c4 = (lane & 15) << 2;       // 16-byte pieces: 2 rows x 16 lanes
    const int64_t ld = v.ld;
    const float* g = v.obs17 + (int64_t)r2 * ld + n0 + c4;
    const uint32_t s = smem_u32(&st[r2][c4]);
#pragma unroll
    for (int k = 0; k < 5; ++k) cp16(s + 2 * k * 256, g + (int64_t)(2 * k) * ld);                      // matrix rows 0..9
#pragma unroll
    for (int k = 0; k < 5; ++k) cp16(s + (kSW + 2 * k) * 256, g + (int64_t)(14 + 2 * k) * ld);         // matrix rows 14..23
    if (r2 == 0) cp16(s + (kSW + 10) * 256, g + (int64_t)24 * ld);                                     // matrix row 24 (episode)
    if (lane < 4) cp16(smem_u32(reinterpret_cast<unsigned char*>(st[kRowBytes]) + 16 * lane), v.flags + n0 + 16 * lane);
    const bool full = n0 + 64 <= v.N;
    if (x.act_vec && full) {
        const float* ga = action + (int64_t)r2 * v.N + n0 + c4;
        cp16(s + kRowAct * 256, ga);
        cp16(s + (kRowAct + 2) * 256, ga + 2 * v.N);
    } else {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            const int col = lane + 32 * hf;
            if (n0 + col < v.N) {
#pragma unroll
                for (int k = 0; k < 4; ++k) cp4(smem_u32(&st[kRowAct + k][col]), action + (int64_t)k * v.N + n0 + col);
            }
        }
    }
    cp_commit();
}

// the sensor_state rows of one chunk as their own cp.async group: they are read, and therefore re-filled, later in the
// iteration than the state rows
__device__ __forceinline__ void prefetch_sensor(const SimView<float>& v, int64_t n0, Row* srows, int lane) {
    const int r2 = lane >> 4, c4 = (lane & 15) << 2;
    const float* gs = v.sensor_state + (int64_t)r2 * v.ld + n0 + c4;
    const uint32_t s = smem_u32(&srows[r2][c4]);
#pragma unroll
    for (int k = 0; k < 10; ++k) cp16(s + 2 * k * 256, gs + (int64_t)(2 * k) * v.ld);
    cp_commit();
}

// row k of a [rows][ld] 4-byte matrix, columns n0+2*lane, n0+2*lane+1 <- {a, b}; g2 = (float2*)(matrix + n0) + lane, ld2 = ld / 2
__device__ __forceinline__ void stg2(float2* g2, int64_t ld2, int k, float a, float b) { g2[(int64_t)k * ld2] = make_float2(a, b); }

// what the sensor phase needs to know about a stepped env pair
struct PairMeta {
    uint32_t episode[2], step_i[2], flags[2];     // after the step
    bool warm[2];                                 // this step was a warm-up step of quad.reset
};

// Sensor phase of a pair, arithmetic only: trailing drone_eq evaluation at the new state (rotation matrix, accelerometer
// reading), the packed sensor model, then the warm-up rules (warm-up steps bypass the model: state kept, true observation
// passed through; the last one re-initialises it = sensor.reset).  srows = the pair's sensor_state rows in shared memory.
template <int ZS = 0>
__device__ __forceinline__ void sensor_phase(const DevParams<float>& p, const SimView<float>& v, const Row* srows, int lane, int64_t nA,
                                             const qs::P2 y[13], const qs::P2 vq[4], qs::P2 f_m, const PairMeta& m,
                                             qs::P2 sn[qs::kSensorStateDim], qs::P2 so[14], const qs::P2* zpre = nullptr) {
    using namespace qs;
    P2 rot[9], acc[3];
    accel_read2(p, f_m, y, rot, acc);
#pragma unroll
    for (int k = 0; k < kSensorStateDim; ++k) sn[k].v = lds2(srows, k, lane);
    SensorRng2 rng;
    rng.seed = v.seed; rng.rk = v.rk;
    rng.id[0] = v.env_id_offset + (uint32_t)nA; rng.id[1] = rng.id[0] + 1u;
    rng.ep[0] = m.episode[0]; rng.ep[1] = m.episode[1];
    rng.step[0] = m.step_i[0]; rng.step[1] = m.step_i[1];
    sensor_step2<ZS>(p, rng, y, acc, rot, f_m, sn, so, zpre);
    const bool w0 = m.warm[0], w1 = m.warm[1];
    if (__any_sync(0xffffffffu, w0 | w1)) {
#pragma unroll
        for (int k = 0; k < kSensorStateDim; ++k) { P2 old; old.v = lds2(srows, k, lane); sn[k] = psel(w0, w1, old, sn[k]); }
#pragma unroll
        for (int k = 0; k < 10; ++k) so[k] = psel(w0, w1, y[k], so[k]);
#pragma unroll
        for (int k = 0; k < 4; ++k) so[10 + k] = psel(w0, w1, vq[k], so[10 + k]);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            if (m.warm[hf] && (m.flags[hf] >> EF_WARM_SHIFT) == 0) {
                float yh[13], sh_[kSensorStateDim];
#pragma unroll
                for (int k = 0; k < 13; ++k) yh[k] = hf ? y[k].v.y : y[k].v.x;
                sensor_reset(p, v.seed, rng.id[hf], rng.ep[hf], yh, sh_);
#pragma unroll
                for (int k = 0; k < kSensorStateDim; ++k) { if (hf) sn[k].v.y = sh_[k]; else sn[k].v.x = sh_[k]; }
            }
        }
    }
}

__device__ __forceinline__ void sensor_store(const SimView<float>& v, int64_t n0, int lane, const qs::P2 sn[qs::kSensorStateDim],
                                             const qs::P2 so[14]) {
    const int64_t ld2 = v.ld >> 1;
    float2* gs2 = reinterpret_cast<float2*>(v.sensor_state + n0) + lane;
#pragma unroll
    for (int k = 0; k < qs::kSensorStateDim; ++k) gs2[(int64_t)k * ld2] = sn[k].v;
    float2* go2 = reinterpret_cast<float2*>(v.sensed_obs + n0) + lane;
#pragma unroll
    for (int k = 0; k < 14; ++k) go2[(int64_t)k * ld2] = so[k].v;
}

// Streaming sensor phase (sensor_step2_stream): state rows read / written in place in the warp's shared-memory stage, outputs
// handed to `out(k, value)` as they are produced (value = what the sensor model computed: the caller substitutes the true
// observation for envs in a warm-up step).  The last warm-up step of an episode re-initialises the sensor (sensor.reset).
template <typename Out>
__device__ __forceinline__ void sensor_phase_stream(const DevParams<float>& p, const SimView<float>& v, Row* srows, int lane, int64_t nA,
                                                    const qs::P2 y[13], qs::P2 f_m, const PairMeta& m, bool any_warm, Out&& out,
                                                    const qs::P2* zpre = nullptr) {
    using namespace qs;
    SensorMem sm;
    sm.st = reinterpret_cast<float2*>(srows[0]) + lane;
    sm.any_warm = any_warm; sm.w0 = m.warm[0]; sm.w1 = m.warm[1];
    SensorRng2 rng;
    rng.seed = v.seed; rng.rk = v.rk;
    rng.id[0] = v.env_id_offset + (uint32_t)nA; rng.id[1] = rng.id[0] + 1u;
    rng.ep[0] = m.episode[0]; rng.ep[1] = m.episode[1];
    rng.step[0] = m.step_i[0]; rng.step[1] = m.step_i[1];
    sensor_step2_stream(p, rng, y, f_m, sm, out, zpre);
    if (any_warm) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            if (m.warm[hf] && (m.flags[hf] >> EF_WARM_SHIFT) == 0) {
                float yh[13], sh_[kSensorStateDim];
#pragma unroll
                for (int k = 0; k < 13; ++k) yh[k] = hf ? y[k].v.y : y[k].v.x;
                sensor_reset(p, v.seed, rng.id[hf], rng.ep[hf], yh, sh_);
#pragma unroll
                for (int k = 0; k < kSensorStateDim; ++k) srows[k][2 * lane + hf] = sh_[k];
            }
        }
    }
}

// the pair's sensor-state rows: shared memory -> HBM
__device__ __forceinline__ void sensor_rows_store(const SimView<float>& v, int64_t n0, int lane, const Row* srows) {
    const int64_t ld2 = v.ld >> 1;
    float2* gs2 = reinterpret_cast<float2*>(v.sensor_state + n0) + lane;
#pragma unroll
    for (int k = 0; k < qs::kSensorStateDim; ++k) gs2[(int64_t)k * ld2] = lds2(srows, k, lane);
}

// fast_atan2f / fast_asinf (quad_device.cuh) for an env pair: the same operation sequence, the polynomial parts on the packed
// pipe, min / max / compare / select / MUFU per env (no packed form in the ISA).  Bit-identical to the scalar routines.
__device__ __forceinline__ qs::P2 patan2(qs::P2 y, qs::P2 x) {
    using namespace qs;
    const float ax0 = fabsf(x.v.x), ay0 = fabsf(y.v.x), ax1 = fabsf(x.v.y), ay1 = fabsf(y.v.y);
    const float mx0 = fmaxf(ax0, ay0), mn0 = fminf(ax0, ay0), mx1 = fmaxf(ax1, ay1), mn1 = fminf(ax1, ay1);
    P2 a = pmul(pk(mn0, mn1), pk(fast_rcpf(mx0), fast_rcpf(mx1)));
    a = pk(mx0 == 0.f ? 0.f : a.v.x, mx1 == 0.f ? 0.f : a.v.y);
    const P2 s2 = pmul(a, a);
    P2 r = bc(0.0027856871f);
    r = pfma(r, s2, bc(-0.0158660002f));
    r = pfma(r, s2, bc(0.042472221f));
    r = pfma(r, s2, bc(-0.0749753043f));
    r = pfma(r, s2, bc(0.106448799f));
    r = pfma(r, s2, bc(-0.142070308f));
    r = pfma(r, s2, bc(0.199934542f));
    r = pfma(r, s2, bc(-0.333331466f));
    r = pmul(r, s2);
    r = pfma(r, a, a);
    float r0 = r.v.x, r1 = r.v.y;
    r0 = (ay0 > ax0) ? (1.57079637f - r0) : r0;  r1 = (ay1 > ax1) ? (1.57079637f - r1) : r1;
    r0 = (x.v.x < 0.f) ? (3.14159274f - r0) : r0;  r1 = (x.v.y < 0.f) ? (3.14159274f - r1) : r1;
    return pk(copysignf(r0, y.v.x), copysignf(r1, y.v.y));
}
__device__ __forceinline__ qs::P2 pasin(qs::P2 x) {
    using namespace qs;
    const float ax0 = fabsf(x.v.x), ax1 = fabsf(x.v.y);
    const bool big0 = ax0 > 0.5f, big1 = ax1 > 0.5f;
    const P2 z = pk(big0 ? fmaf(-0.5f, ax0, 0.5f) : ax0 * ax0, big1 ? fmaf(-0.5f, ax1, 0.5f) : ax1 * ax1);
    const P2 sq = pk(big0 ? fast_sqrtf(z.v.x) : ax0, big1 ? fast_sqrtf(z.v.y) : ax1);
    P2 pz = bc(4.2163199048e-2f);
    pz = pfma(pz, z, bc(2.4181311049e-2f));
    pz = pfma(pz, z, bc(4.5470025998e-2f));
    pz = pfma(pz, z, bc(7.4953002686e-2f));
    pz = pfma(pz, z, bc(1.6666752422e-1f));
    const P2 r = pfma(pmul(sq, z), pz, sq);
    float r0 = r.v.x, r1 = r.v.y;
    r0 = big0 ? fmaf(-2.f, r0, 1.57079637f) : r0;  r1 = big1 ? fmaf(-2.f, r1, 1.57079637f) : r1;
    r0 = (ax0 > 1.f) ? __int_as_float(0x7fc00000) : r0;  r1 = (ax1 > 1.f) ? __int_as_float(0x7fc00000) : r1;
    return pk(copysignf(r0, x.v.x), copysignf(r1, x.v.y));
}

// Phase 3 of quad.step (step_post, quad_device.cuh: :486-498 observation tail, Euler angles, done_condition, reward_function,
// control_effort) for an env pair.  Everything both envs compute alike — quaternion normalisation, V_q, the Euler-angle
// arguments and polynomials, the sums of squares of the reward, the action penalty, the effort norm — runs on the packed
// pipe; thresholds, the reward cascade and the flag logic stay per env (branch-free, the two chains interleave).
__device__ __forceinline__ void step_post2(const DevParams<float>& p, const qs::P2 y[13], Env<float> e[2], const float act[2][4],
                                           StepOut<float> o[2]) {
    using namespace qs;
    const P2 inv = prsqrt(pfma(y[6], y[6], pfma(y[7], y[7], pfma(y[8], y[8], pmul(y[9], y[9])))));        // :488-489
    const P2 q[4] = {pmul(y[6], inv), pmul(y[7], inv), pmul(y[8], inv), pmul(y[9], inv)};
    P2 vq[4];
    deriv_quat2(&y[10], q, vq);                                                                         // :392
    // quat_euler utility:39-48
    const P2 two = bc(2.f), m2 = bc(-2.f), one = bc(1.f);
    const P2 sx = pmul(two, pfma(q[0], q[1], pmul(q[2], q[3])));
    const P2 cx = pfma(m2, pfma(q[1], q[1], pmul(q[2], q[2])), one);
    const P2 sy = pmul(two, pfma(q[0], q[2], pmul(pmul(q[3], bc(-1.f)), q[1])));
    const P2 sz = pmul(two, pfma(q[0], q[3], pmul(q[1], q[2])));
    const P2 cz = pfma(m2, pfma(q[2], q[2], pmul(q[3], q[3])), one);
    const P2 phi = patan2(sx, cx);
    const P2 theta = pasin(pk(asin_arg<float>(sy.v.x), asin_arg<float>(sy.v.y)));
    const P2 psi = patan2(sz, cz);
    // reward_function :511-573, the sums
    const P2 v2 = pfma(y[1], y[1], pfma(y[3], y[3], pmul(y[5], y[5])));
    const P2 e2 = pfma(phi, phi, pmul(theta, theta));
    const P2 psi2 = pmul(psi, psi);
    const P2 w2 = pfma(y[10], y[10], pfma(y[11], y[11], pmul(y[12], y[12])));
    const P2 cur = padd(v2, padd(padd(e2, psi2), w2));                                                   // :558
    const P2 nr2 = padd(v2, psi2);
    P2 pen = bc(0.f);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const P2 d = padd(pk(act[0][k], act[1][k]), bc(-p.zero_control[k]));
        pen = pfma(d, d, pen);
    }
    const P2 ef2 = pfma(pk(o[0].effort[0], o[1].effort[0]), pk(o[0].effort[0], o[1].effort[0]),
                   pfma(pk(o[0].effort[1], o[1].effort[1]), pk(o[0].effort[1], o[1].effort[1]),
                   pfma(pk(o[0].effort[2], o[1].effort[2]), pk(o[0].effort[2], o[1].effort[2]),
                   pmul(pk(o[0].effort[3], o[1].effort[3]), pk(o[0].effort[3], o[1].effort[3])))));
    const P2 nv = pk(fast_sqrtf(v2.v.x), fast_sqrtf(v2.v.y)), ne = pk(fast_sqrtf(e2.v.x), fast_sqrtf(e2.v.y));
    const P2 shaping0 = pmul(bc(-1.f), pfma(bc(p.sh_v), nv, pfma(bc(p.sh_psi), pk(fabsf(psi.v.x), fabsf(psi.v.y)), pmul(bc(p.sh_ang), ne))));   // :529-531
    const P2 pen_c = pmul(pen, bc(-p.p_c));
#pragma unroll
    for (int h = 0; h < 2; ++h) {                 // the per-env tail is the scalar kernels' own (post_tail, quad_device.cuh)
        const bool hi = h != 0;
#pragma unroll
        for (int k = 0; k < 13; ++k) e[h].y[k] = hi ? y[k].v.y : y[k].v.x;
        const float ang[3] = {hi ? phi.v.y : phi.v.x, hi ? theta.v.y : theta.v.x, hi ? psi.v.y : psi.v.x};
        const float vqh[4] = {hi ? vq[0].v.y : vq[0].v.x, hi ? vq[1].v.y : vq[1].v.x, hi ? vq[2].v.y : vq[2].v.x, hi ? vq[3].v.y : vq[3].v.x};
        PostSums r;
        r.v2 = hi ? v2.v.y : v2.v.x; r.e2 = hi ? e2.v.y : e2.v.x; r.ne = hi ? ne.v.y : ne.v.x; r.nr2 = hi ? nr2.v.y : nr2.v.x;
        r.cur = hi ? cur.v.y : cur.v.x; r.shaping0 = hi ? shaping0.v.y : shaping0.v.x; r.pen_c = hi ? pen_c.v.y : pen_c.v.x;
        r.ef2 = hi ? ef2.v.y : ef2.v.x;
        post_tail(p, e[h], o[h], ang, vqh, r);
    }
}

// QS_FLAG_ASYNC_RESET, once per chunk: push the finished envs of this lane's pair on the CTA's queue (AFTER all of their
// rows have been stored), then claim 32 queued envs if available and re-sample them with all lanes busy
template <bool SENSOR, int kQueueCap>
__device__ __forceinline__ void reset_queue_step(const DevParams<float>& p, const SimView<float>& v, const StepIO<float>& io,
                                                 uint32_t* s_queue, int* s_qn, int* s_qhead, int lane, int64_t nA, bool push0, bool push1) {
    constexpr unsigned kFull = 0xffffffffu;
    if (__any_sync(kFull, push0 || push1)) {
        __threadfence_block();                 // this warp's stores of the finished envs precede the re-sampler's
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            if (hf ? push1 : push0) {
                const int slot = atomicAdd(s_qn, 1);
                if (slot < kQueueCap) *reinterpret_cast<volatile uint32_t*>(&s_queue[slot]) = (uint32_t)(nA + hf);
                else wp::resample_env<SENSOR>(p, v, io, nA + hf);   // queue exhausted (e.g. a whole shard timing out at once)
            }
        }
    }
    int take = -1;
    if (lane == 0) {
        const int head = *reinterpret_cast<volatile int*>(s_qhead);
        int qn = *reinterpret_cast<volatile int*>(s_qn);
        qn = qn < kQueueCap ? qn : kQueueCap;
        if (qn - head >= 32 && atomicCAS(s_qhead, head, head + 32) == head) take = head;
    }
    take = __shfl_sync(kFull, take, 0);
    if (take >= 0) {
        uint32_t ent;
        do { ent = *reinterpret_cast<volatile uint32_t*>(&s_queue[take + lane]); } while (ent == 0xFFFFFFFFu);
        __threadfence_block();
        wp::resample_env<SENSOR>(p, v, io, (int64_t)ent);
    }
}

// Producer warp of the sensor kernel: for the chunks of its two chunk warps, in their order, draw blocks 0..2 of the sensor stream of
// every env-step (sensor_normals_block2: Philox4x32-10 + Box-Muller, the same arithmetic the chunk warp would run inline) into the
// chunk warp's normals buffer.  Counters of the step being taken: (global env id, episode, step_i + 1) read from the handle's rows —
// the chunk warp stores the incremented step counter only after it has waited for this buffer (the episode counter changes only
// when a finished env is re-sampled, after its chunk's sensor phase).  full / empty: one
// mbarrier per (chunk warp, buffer), count 32 (every lane arrives for itself).
__device__ __forceinline__ void normals_producer(const SimView<float>& v, Row* zbase, uint64_t* s_full, uint64_t* s_empty, int pw, int lane,
                                                 int64_t n_chunks, int64_t stride) {
    using namespace qs;
    for (int it = 0;; ++it) {
        bool any = false;
#pragma unroll 1
        for (int j = 0; j < kServed; ++j) {
            const int cw = kServed * pw + j;
            const int64_t c = (int64_t)cw * gridDim.x + blockIdx.x + (int64_t)it * stride;
            if (c >= n_chunks) continue;
            any = true;
            const int slot = cw * kZBufs + (it % kZBufs);
            mbar_wait(&s_empty[slot], ((uint32_t)(it / kZBufs) & 1u) ^ 1u);      // free from the start: the first wait per buffer falls through
            const int64_t n0 = c << 6;
            const float2 si = reinterpret_cast<const float2*>(v.obs17 + (int64_t)wp::kMStepI * v.ld + n0)[lane];
            const float2 ep = reinterpret_cast<const float2*>(v.obs17 + (int64_t)wp::kMEpisode * v.ld + n0)[lane];
            SensorRng2 rng;
            rng.seed = v.seed; rng.rk = v.rk;
            rng.id[0] = v.env_id_offset + (uint32_t)(n0 + 2 * lane); rng.id[1] = rng.id[0] + 1u;
            rng.ep[0] = __float_as_uint(ep.x); rng.ep[1] = __float_as_uint(ep.y);
            rng.step[0] = (uint32_t)(__float_as_int(si.x) + 1); rng.step[1] = (uint32_t)(__float_as_int(si.y) + 1);
            Row* zr = zbase + (size_t)slot * kZRows;
#pragma unroll 1
            for (int b = 0; b < 3; ++b) {
                P2 z[8];
                sensor_normals_block2(rng, b, z);
#pragma unroll
                for (int k = 0; k < 8; ++k) reinterpret_cast<float2*>(zr[8 * b + k])[lane] = z[k].v;
            }
            mbar_arrive(&s_full[slot]);                          // every lane releases its own stores (count 32)
        }
        if (!any) break;
    }
}

}  // namespace pr

template <bool DIRECT, bool SENSOR>
__global__ void __launch_bounds__(SENSOR ? pr::kThreadsSensor : pr::kThreadsPlain, 1)
step_kernel_pair(const __grid_constant__ DevParams<float> p, const __grid_constant__ SimView<float> v,
                 const __grid_constant__ StepIO<float> io) {
    using namespace pr;
    constexpr int kRows = SENSOR ? kRowsSensor : kRowsPlain;
    constexpr int kQueueCap = SENSOR ? kQueueCapSensor : kQueueCapPlain;
    constexpr int kThreads = SENSOR ? kThreadsSensor : kThreadsPlain;                 // all threads of the CTA
    constexpr bool kProd = SENSOR && kProducer;
    constexpr int kWarps = (SENSOR ? kConsumerThreadsSensor : kThreadsPlain) / 32;    // chunk warps
    constexpr int kChunkThreads = kWarps * 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint32_t s_queue[kQueueCap];
    __shared__ int s_qn, s_qhead;
    __shared__ uint64_t s_full[kProd ? kZBufs * kWarps : 1], s_empty[kProd ? kZBufs * kWarps : 1];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    Row* st = reinterpret_cast<Row*>(smem_raw) + (size_t)(w < kWarps ? w : 0) * kRows;
    Row* srows = st + kRowSensor;                            // SENSOR: the pair's sensor_state rows
    Row* zbase = reinterpret_cast<Row*>(smem_raw) + (size_t)kWarps * kRows;          // [chunk warp][kZBufs][24] rows of normals
    for (int i = tid; i < kQueueCap; i += kThreads) s_queue[i] = 0xFFFFFFFFu;
    if (tid == 0) {
        s_qn = 0; s_qhead = 0;
        if (kProd) {
            for (int i = 0; i < kZBufs * kWarps; ++i) { mbar_init(&s_full[i], 32); mbar_init(&s_empty[i], 32); }
            mbar_fence_init();
        }
    }
    __syncthreads();
    if (kProd) {
        if (w >= kWarps) {
            asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
            normals_producer(v, zbase, s_full, s_empty, w - kWarps, lane, (v.N + 63) >> 6, (int64_t)gridDim.x * kWarps);
            __syncthreads();                   // the chunk warps' "every push has been made"
            LocalStats none;
            none.clear();
            flush_stats(none, false, v.stats);
            return;
        }
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kConsumerRegs));
    }

    Ext x;
    const bool n2 = (v.N & 1) == 0, n4 = (v.N & 3) == 0;
    x.act_vec = n4 && ((reinterpret_cast<uintptr_t>(io.action) & 15) == 0);
    x.obs_vec = n2 && ((reinterpret_cast<uintptr_t>(io.obs) & 7) == 0);         // 8-byte pair stores into the caller's arrays
    x.rew_vec = n2 && ((reinterpret_cast<uintptr_t>(io.reward) & 7) == 0);
    x.done_vec = n2 && ((reinterpret_cast<uintptr_t>(io.done) & 1) == 0);       // 2-byte pair stores
    x.solved_vec = n2 && ((reinterpret_cast<uintptr_t>(io.solved) & 1) == 0);
    const bool async_reset = (p.flags & F_ASYNC_RESET) != 0;
    const int64_t ld2 = v.ld >> 1, N2 = v.N >> 1;

    LocalStats ls;
    ls.clear();
    bool any_end = false;
    const int64_t n_chunks = (v.N + 63) >> 6;
    const int64_t stride = (int64_t)gridDim.x * kWarps;
    // warp-major chunk order: the LAST, partial round of the grid-stride loop (16,384 chunks over 1,776 warps at 1M envs = 9.2
    // rounds) is spread over all CTAs as a few warps each instead of keeping the first CTAs full and leaving the others idle
    // (49.3 -> 47.6 us per step of 1,048,576 envs)
    int64_t c = (int64_t)w * gridDim.x + blockIdx.x;

    if (c < n_chunks) {
        prefetch(v, io.action, x, c << 6, st, lane);
        if (SENSOR) prefetch_sensor(v, c << 6, srows, lane);
    }
    uint32_t it = 0;                                         // chunks done by this warp (normals buffer = it & 1)
#if QS_PAIR_LOCKSTEP
    // -DQS_PAIR_LOCKSTEP=1 (tuning build): the chunk warps of a CTA start every chunk together.  Motive: the hot path of the loop
    // body (2,041 instructions = 32.6 KB with the sensor model) sits right at the SM's 32 KB instruction cache — ncu counts 13.5 %
    // misses there (sm__icc_request_hit_rate 86.5 %) and the GPC-level cache behind it at 80 % of its request rate
    // (gcc__cache_requests_type_instruction) — and in step one warp's miss would fetch the line for all.  Measured and rejected
    // (gpurun_out/r2w_lockstep_ab.txt): 101.6 -> 111.4 us per step of 1,048,576 envs with the sensor model, 47.1 -> 59.6 without:
    // warps in step also wait for memory and for the MUFU / FMA pipes in step.
    const uint32_t iters_max = blockIdx.x < n_chunks ? (uint32_t)((n_chunks - 1 - blockIdx.x) / stride + 1) : 0u;   // warp 0's count
    for (; it < iters_max; c += stride, ++it) {
        asm volatile("bar.sync 1, %0;" ::"n"(kChunkThreads) : "memory");
        if (c >= n_chunks) continue;
#else
    for (; c < n_chunks; c += stride, ++it) {
#endif
        // cp.async groups complete in order.  SENSOR: pending here = {state(c), sensor(c)}; the state rows are needed now
        if (SENSOR) cp_wait<1>(); else cp_wait<0>();
        __syncwarp();
        const int64_t n0 = c << 6;
        const int64_t nA = n0 + 2 * lane;                       // env A; env B = nA + 1
        const bool act_[2] = {nA < v.N, nA + 1 < v.N};
        // ---- stage -> registers: the state stays packed {A, B} from the LDS.64 to the end of the integration
        P2 y[13];
#pragma unroll
        for (int k = 0; k < 10; ++k) y[k].v = lds2(st, k, lane);
#pragma unroll
        for (int k = 0; k < 3; ++k) y[10 + k].v = lds2(st, kSW + k, lane);
        float2 pa[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) pa[k] = lds2(st, kSW + 3 + k, lane);
        const float2 sh = lds2(st, kSW + 6, lane), as = lds2(st, kSW + 7, lane), er = lds2(st, kSW + 8, lane);
        const float2 si = lds2(st, kSW + 9, lane), ep = lds2(st, kSW + 10, lane);
        const uint32_t fl = reinterpret_cast<const uint16_t*>(st[kRowBytes])[lane];
        float2 a2[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) a2[k] = lds2(st, kRowAct + k, lane);
        __syncwarp();                          // every lane has read its pair: the stage is free for the next chunk
        const bool has_next = c + stride < n_chunks;
        if (has_next) prefetch(v, io.action, x, (c + stride) << 6, st, lane);
        else if (SENSOR) cp_commit();           // keep the group count uniform: {sensor(c), state(next) or empty}

        if constexpr (SENSOR && kSelfNormals) {
            // counters of the step being taken: (global env id, episode, step_i + 1); 24 normals -> this warp's rows, lane's own columns
            SensorRng2 rng;
            rng.seed = v.seed; rng.rk = v.rk;
            rng.id[0] = v.env_id_offset + (uint32_t)nA; rng.id[1] = rng.id[0] + 1u;
            rng.ep[0] = __float_as_uint(ep.x); rng.ep[1] = __float_as_uint(ep.y);
            rng.step[0] = (uint32_t)(__float_as_int(si.x) + 1); rng.step[1] = (uint32_t)(__float_as_int(si.y) + 1);
            Row* zr = zbase + (size_t)w * kZRows;
#pragma unroll 1
            for (int b = 0; b < 3; ++b) {
                P2 z[8];
                sensor_normals_block2(rng, b, z);
#pragma unroll
                for (int k = 0; k < 8; ++k) reinterpret_cast<float2*>(zr[8 * b + k])[lane] = z[k].v;
            }
        }
        Env<float> e[2];
        StepOut<float> o[2];
        Ctrl<float> ctl[2];
        float act[2][4];
        bool warm[2], was_done[2], push[2];
        // ---- phase 1 per env: warm-up override, action clip, rotor map
#define QS_PAIR_PRE(H)                                                                                        \
        {                                                                                                     \
            _Pragma("unroll") for (int k = 0; k < 3; ++k) e[H].prev_ang[k] = half_of<H>(pa[k]);              \
            e[H].prev_shaping = half_of<H>(sh); e[H].abs_sum = half_of<H>(as); e[H].ep_return = half_of<H>(er); \
            e[H].i = __float_as_int(half_of<H>(si)); e[H].episode = __float_as_uint(half_of<H>(ep));          \
            e[H].flags = (fl >> (8 * H)) & 0xffu;                                                             \
            float a[4];                                                                                       \
            _Pragma("unroll") for (int k = 0; k < 4; ++k) a[k] = half_of<H>(a2[k]);                           \
            warm[H] = async_reset ? async_warmup_prologue(p, e[H], a) : false;                                \
            was_done[H] = (e[H].flags & EF_DONE) != 0;                                                        \
            ctl[H] = step_pre<float, DIRECT>(p, e[H], a, o[H], act[H]);                                       \
        }
        QS_PAIR_PRE(0)
        QS_PAIR_PRE(1)
#undef QS_PAIR_PRE
        // ---- phase 2: both envs through the RK4 stages on the packed pipe
        const Ctrl2 c2 = pack_ctrl(ctl[0], ctl[1]);
#if QS_SENSOR_FUSED_NORMALS
        P2 zpre[SENSOR ? 24 : 1];
        if constexpr (SENSOR) {
            SensorRng2 rng;                                    // counters of the step being taken: (env, episode, i after the increment)
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
        // ---- phase 3 per env: observation tail, Euler angles, done, reward
        // phase 3 on the packed pipe (step_post2): bit-identical to the scalar FP32 phase of the one-env-per-thread kernels (both are
        // the same sequence of explicitly rounded operations, quad_device.cuh); -DQS_PAIR_PACKED_POST=0 instantiates the scalar phase
        // per half instead (A/B: 102.0 vs 100.6 us per step of 1,048,576 envs with the sensor model)
#if QS_PAIR_PACKED_POST
#define QS_PAIR_POST_CALL(H)
        step_post2(p, y, e, act, o);
#else
#define QS_PAIR_POST_CALL(H) step_post(p, e[H], act[H], o[H]);
#endif
#define QS_PAIR_POST(H)                                                                                       \
        {                                                                                                     \
            _Pragma("unroll") for (int k = 0; k < 13; ++k) e[H].y[k] = half_of<H>(y[k]);                      \
            QS_PAIR_POST_CALL(H)                                                                              \
            o[H].reward = warm[H] ? 0.f : o[H].reward;                                                        \
            e[H].ep_return += o[H].reward;                                                                    \
            push[H] = async_reset & act_[H] & o[H].done;                                                      \
        }
        QS_PAIR_POST(0)
        QS_PAIR_POST(1)
        if (act_[0] && o[0].done && !was_done[0]) { count_episode(ls, p, e[0], o[0]); any_end = true; }
        if (act_[1] && o[1].done && !was_done[1]) { count_episode(ls, p, e[1], o[1]); any_end = true; }
#undef QS_PAIR_POST
#undef QS_PAIR_POST_CALL
        // ---- results -> HBM straight from the register pairs (rows of the handle are padded to whole chunks)
        {
            float2* g2 = reinterpret_cast<float2*>(v.obs17 + n0) + lane;
#pragma unroll
            for (int k = 0; k < 10; ++k) g2[(int64_t)k * ld2] = y[k].v;
#pragma unroll
            for (int k = 0; k < 4; ++k) stg2(g2, ld2, 10 + k, o[0].vq[k], o[1].vq[k]);
#pragma unroll
            for (int k = 0; k < 3; ++k) g2[(int64_t)(14 + k) * ld2] = y[10 + k].v;
#pragma unroll
            for (int k = 0; k < 3; ++k) stg2(g2, ld2, wp::kMAng + k, e[0].prev_ang[k], e[1].prev_ang[k]);
            stg2(g2, ld2, wp::kMShaping, e[0].prev_shaping, e[1].prev_shaping);
            stg2(g2, ld2, wp::kMAbsSum, e[0].abs_sum, e[1].abs_sum);
            stg2(g2, ld2, wp::kMEpRet, e[0].ep_return, e[1].ep_return);
            // producer mode: the producer warp reads the step counters of this chunk from these rows — they are stored only after
            // its normals have been waited for (sensor phase below)
            if (!kProd) stg2(g2, ld2, wp::kMStepI, __int_as_float(e[0].i), __int_as_float(e[1].i));
            stg2(g2, ld2, wp::kMEpisode, __uint_as_float(e[0].episode), __uint_as_float(e[1].episode));
            stg2(g2, ld2, wp::kMReward, o[0].reward, o[1].reward);
            const uint16_t bf = (uint16_t)((e[0].flags & 0xffu) | ((e[1].flags & 0xffu) << 8));
            const uint16_t bd = (uint16_t)(((o[0].done ? 1u : 0u) | (warm[0] ? 2u : 0u)) | (((o[1].done ? 1u : 0u) | (warm[1] ? 2u : 0u)) << 8));
            const uint16_t bs = (uint16_t)((o[0].solved ? 1u : 0u) | ((o[1].solved ? 1u : 0u) << 8));
            reinterpret_cast<uint16_t*>(v.flags + n0)[lane] = bf;
            reinterpret_cast<uint16_t*>(v.done + n0)[lane] = bd;
            reinterpret_cast<uint16_t*>(v.solved + n0)[lane] = bs;
            if (io.obs) {
                if (x.obs_vec && act_[1]) {
                    float2* q2 = reinterpret_cast<float2*>(io.obs + n0) + lane;
#pragma unroll
                    for (int k = 0; k < 10; ++k) q2[(int64_t)k * N2] = y[k].v;
#pragma unroll
                    for (int k = 0; k < 4; ++k) stg2(q2, N2, 10 + k, o[0].vq[k], o[1].vq[k]);
                } else {
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        if (act_[hf]) {
#pragma unroll
                            for (int k = 0; k < 10; ++k) io.obs[k * v.N + nA + hf] = hf ? y[k].v.y : y[k].v.x;
#pragma unroll
                            for (int k = 0; k < 4; ++k) io.obs[(10 + k) * v.N + nA + hf] = o[hf].vq[k];
                        }
                    }
                }
            }
            if (io.reward) {
                if (x.rew_vec && act_[1]) reinterpret_cast<float2*>(io.reward + n0)[lane] = make_float2(o[0].reward, o[1].reward);
                else {
                    if (act_[0]) io.reward[nA] = o[0].reward;
                    if (act_[1]) io.reward[nA + 1] = o[1].reward;
                }
            }
            if (io.done) {
                if (x.done_vec && act_[1]) reinterpret_cast<uint16_t*>(io.done + n0)[lane] = bd;
                else {
                    if (act_[0]) io.done[nA] = (uint8_t)(bd & 0xff);
                    if (act_[1]) io.done[nA + 1] = (uint8_t)(bd >> 8);
                }
            }
            if (io.solved) {
                if (x.solved_vec && act_[1]) reinterpret_cast<uint16_t*>(io.solved + n0)[lane] = bs;
                else {
                    if (act_[0]) io.solved[nA] = (uint8_t)(bs & 0xff);
                    if (act_[1]) io.solved[nA + 1] = (uint8_t)(bs >> 8);
                }
            }
        }
        if (SENSOR) {
            PairMeta m;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                m.episode[hf] = e[hf].episode; m.step_i[hf] = (uint32_t)e[hf].i; m.flags[hf] = e[hf].flags; m.warm[hf] = warm[hf];
            }
            cp_wait<1>();                      // pending = {sensor(c), state(next) | empty}: the sensor rows have landed
            __syncwarp();
#if QS_SENSOR_STREAM
            {
                const bool any_warm = __any_sync(0xffffffffu, warm[0] | warm[1]);
                float2* go2 = reinterpret_cast<float2*>(v.sensed_obs + n0) + lane;
                const float2* gt2 = reinterpret_cast<const float2*>(v.obs17 + n0) + lane;     // rows 0..13 = the true observation, stored above
                const bool w0 = warm[0], w1 = warm[1];
                sensor_phase_stream(p, v, srows, lane, nA, y, c2.f_m, m, any_warm, [&](int k, P2 val) {
                    if (any_warm) { P2 t; t.v = gt2[(int64_t)k * ld2]; val = psel(w0, w1, t, val); }   // warm-up steps pass the true observation through
                    go2[(int64_t)k * ld2] = val.v;
                });
                sensor_rows_store(v, n0, lane, srows);
                __syncwarp();                  // every lane is done with the sensor rows of the stage
                if (has_next) prefetch_sensor(v, (c + stride) << 6, srows, lane); else cp_commit();
            }
#else
            {
                P2 sn[kSensorStateDim], so[14], vq[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) vq[k] = pk(o[0].vq[k], o[1].vq[k]);
#if QS_SENSOR_FUSED_NORMALS
                sensor_phase(p, v, srows, lane, nA, y, vq, c2.f_m, m, sn, so, zpre);
#else
                if constexpr (kProd) {
                    const int slot = w * kZBufs + (int)(it % kZBufs);
                    mbar_wait(&s_full[slot], (it / kZBufs) & 1u);      // the producer warp has drawn this chunk's normals
                    stg2(reinterpret_cast<float2*>(v.obs17 + n0) + lane, ld2, wp::kMStepI, __int_as_float(e[0].i), __int_as_float(e[1].i));
                    const P2* zp = reinterpret_cast<const P2*>(zbase[(size_t)slot * kZRows]) + lane;
                    sensor_phase<32>(p, v, srows, lane, nA, y, vq, c2.f_m, m, sn, so, zp);
                    mbar_arrive(&s_empty[slot]);                   // every lane has read its normals (count 32)
                } else if constexpr (kSelfNormals) {
                    const P2* zp = reinterpret_cast<const P2*>(zbase[(size_t)w * kZRows]) + lane;   // written by this lane above
                    sensor_phase<32>(p, v, srows, lane, nA, y, vq, c2.f_m, m, sn, so, zp);
                } else {
                    sensor_phase(p, v, srows, lane, nA, y, vq, c2.f_m, m, sn, so);
                }
#endif
                __syncwarp();                  // every lane is done with the sensor rows of the stage
                if (has_next) prefetch_sensor(v, (c + stride) << 6, srows, lane); else cp_commit();
                sensor_store(v, n0, lane, sn, so);
            }
#endif
        }
        if (async_reset)
            reset_queue_step<SENSOR, kQueueCap>(p, v, io, s_queue, &s_qn, &s_qhead, lane, nA, push[0], push[1]);
    }
    __syncthreads();                           // every push of this CTA has been made
    if (async_reset) {
        const int head = s_qhead;
        const int qn = s_qn < kQueueCap ? s_qn : kQueueCap;
        __threadfence_block();
        for (int q = head + tid; q < qn; q += kChunkThreads) wp::resample_env<SENSOR>(p, v, io, (int64_t)s_queue[q]);
    }
    flush_stats(ls, any_end, v.stats);
    if (blockIdx.x == 0 && tid == 0) atomicAdd(&v.stats[7], (double)v.N);
}
