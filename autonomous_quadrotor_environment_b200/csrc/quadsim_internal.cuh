// quadsim_internal.cuh — what the translation units of libquadsim.so share: the handle, the SoA view the kernels take, the
// per-env load / store helpers, episode statistics and the launch-side helpers.  (One .cu per kernel family, compiled in
// parallel by __graft_entry__.build(); no relocatable device code: nothing on the device crosses a translation unit.)
#pragma once
#include "../../include/quadsim.h"
#include "quad_device.cuh"
#include "sensor_device.cuh"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <atomic>

using namespace qs;

// ------------------------------------------------------------------------------------------------
// error plumbing (quadsim.cu owns the thread-local message)
// ------------------------------------------------------------------------------------------------
__attribute__((visibility("hidden"))) int qs_fail_(int code, const char* fmt, const char* a = "", const char* b = "");
#define fail qs_fail_
#define QS_CUDA(call)                                                                   \
    do {                                                                                \
        cudaError_t e_ = (call);                                                        \
        if (e_ != cudaSuccess) return fail(QS_ECUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)


// ------------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------------
template <typename R> struct SimView {
    int64_t N, ld;
    R* obs17;
    R* prev_ang;
    R* prev_shaping;
    R* abs_sum;
    R* ep_return;
    R* reward;
    int32_t* step_i;
    uint32_t* episode;
    uint8_t* flags;
    uint8_t* done;
    uint8_t* solved;
    R* ang_vel;       // AUX (nullable)
    R* step_effort;
    R* w;
    R* accel;
    R* acc_read;
    R* mat_rot;
    R* clipped_action;
    R* fm;
    R* sensed_obs;    // SENSOR (nullable)
    R* sensor_state;
    int32_t* gust_count;   // ROBUST (nullable): per-env gust counter of robust_control.wind
    double* stats;
    uint64_t seed;
    uint32_t env_id_offset;
    uint32_t rk[20];       // Philox round keys of seed (philox_round_keys)
};

struct Slot { void* ptr; int32_t channels; int32_t elem; };

struct qs_sim {
    qs_config cfg;
    int64_t N, ld;
    int rs;                 // sizeof(real)
    char* ws;
    size_t ws_bytes;
    bool owns_ws;
    Slot slot[QS_FIELD_COUNT_];
    void* obs17;
    void* action_stage;     // [4][ld] staging for qs_step_host
    double* stats;
    DevParams<float> pf;
    DevParams<double> pd;
    int sm_count;
    uint64_t seed;
    int64_t slice_begin, slice_count;   // env range the step launchers address (whole shard except inside qs_step_host's pipeline)
    cudaStream_t host_streams[4];       // qs_step_host: slices of the shard flow H2D -> step -> D2H on these, overlapping both PCIe directions
    cudaEvent_t host_ev[5];
    bool host_pipe_ready;
    int step_loader;        // 0 direct LDG, 1 CTA-wide TMA ring, 2 per-warp cp.async pipeline, 3 per-warp pipeline with env pairs (packed FP32)
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

template <typename R> static inline SimView<R> make_view(const qs_sim* s) {
    SimView<R> v;
    memset(&v, 0, sizeof(v));
    v.N = s->N; v.ld = s->ld;
    v.obs17 = (R*)s->obs17;
    v.prev_ang = (R*)s->slot[QS_FIELD_ANG].ptr;
    v.prev_shaping = (R*)s->slot[QS_FIELD_PREV_SHAPING].ptr;
    v.abs_sum = (R*)s->slot[QS_FIELD_ABS_SUM].ptr;
    v.ep_return = (R*)s->slot[QS_FIELD_EP_RETURN].ptr;
    v.reward = (R*)s->slot[QS_FIELD_REWARD].ptr;
    v.step_i = (int32_t*)s->slot[QS_FIELD_I].ptr;
    v.episode = (uint32_t*)s->slot[QS_FIELD_EPISODE].ptr;
    v.flags = (uint8_t*)s->slot[QS_FIELD_FLAGS].ptr;
    v.done = (uint8_t*)s->slot[QS_FIELD_DONE].ptr;
    v.solved = (uint8_t*)s->slot[QS_FIELD_SOLVED].ptr;
    v.ang_vel = (R*)s->slot[QS_FIELD_ANG_VEL].ptr;
    v.step_effort = (R*)s->slot[QS_FIELD_STEP_EFFORT].ptr;
    v.w = (R*)s->slot[QS_FIELD_W].ptr;
    v.accel = (R*)s->slot[QS_FIELD_ACCEL].ptr;
    v.acc_read = (R*)s->slot[QS_FIELD_ACC_READ].ptr;
    v.mat_rot = (R*)s->slot[QS_FIELD_MAT_ROT].ptr;
    v.clipped_action = (R*)s->slot[QS_FIELD_CLIPPED_ACTION].ptr;
    v.fm = (R*)s->slot[QS_FIELD_FM].ptr;
    v.sensed_obs = (R*)s->slot[QS_FIELD_SENSED_OBS].ptr;
    v.sensor_state = (R*)s->slot[QS_FIELD_SENSOR_STATE].ptr;
    v.gust_count = (int32_t*)s->slot[QS_FIELD_GUST_COUNT].ptr;
    v.stats = s->stats;
    v.seed = s->seed;
    v.env_id_offset = (uint32_t)s->cfg.env_id_offset;
    qs::philox_round_keys(s->seed, v.rk);
    if (s->slice_begin != 0 || s->slice_count != s->N) {       // a 256-aligned sub-range of the shard: same rows, shifted columns
        const int64_t b = s->slice_begin;
        R** real_rows[] = {&v.obs17, &v.prev_ang, &v.prev_shaping, &v.abs_sum, &v.ep_return, &v.reward, &v.ang_vel, &v.step_effort,
                           &v.w, &v.accel, &v.acc_read, &v.mat_rot, &v.clipped_action, &v.fm, &v.sensed_obs, &v.sensor_state};
        for (R** r : real_rows) if (*r) *r += b;
        v.step_i += b; v.episode += b; v.flags += b; v.done += b; v.solved += b;
        if (v.gust_count) v.gust_count += b;
        v.N = s->slice_count;
        v.env_id_offset += (uint32_t)b;
    }
    return v;
}

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
#ifndef QS_BLOCK
#define QS_BLOCK 256          // threads (= envs) per CTA tile
#endif
#ifndef QS_MIN_CTAS
#define QS_MIN_CTAS 2         // resident CTAs per SM the step / rollout kernels are compiled for
#endif
#ifndef QS_STAGES
#define QS_STAGES 3           // depth of the TMA staging ring
#endif
constexpr int kBlock = QS_BLOCK;
// The FP64 replica of SciPy's RK45 keeps seven stage vectors of 13 doubles: its kernels are compiled for ONE resident CTA per SM (255
// registers per thread) instead of QS_MIN_CTAS.
template <typename R, int INTEG> constexpr int qs_min_ctas() { return (sizeof(R) == 8 && INTEG == 1) ? 1 : QS_MIN_CTAS; }

template <typename R>
__device__ __forceinline__ void load_env(const SimView<R>& v, int64_t n, Env<R>& e) {
#pragma unroll
    for (int k = 0; k < 10; ++k) e.y[k] = v.obs17[k * v.ld + n];
#pragma unroll
    for (int k = 0; k < 3; ++k) e.y[10 + k] = v.obs17[(14 + k) * v.ld + n];
#pragma unroll
    for (int k = 0; k < 3; ++k) e.prev_ang[k] = v.prev_ang[k * v.ld + n];
    e.prev_shaping = v.prev_shaping[n];
    e.abs_sum = v.abs_sum[n];
    e.ep_return = v.ep_return[n];
    e.i = v.step_i[n];
    e.flags = v.flags[n];
    e.episode = v.episode[n];
}

template <typename R>
__device__ __forceinline__ void store_env(const SimView<R>& v, int64_t n, const Env<R>& e, const R vq[4]) {
#pragma unroll
    for (int k = 0; k < 10; ++k) v.obs17[k * v.ld + n] = e.y[k];
#pragma unroll
    for (int k = 0; k < 4; ++k) v.obs17[(10 + k) * v.ld + n] = vq[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) v.obs17[(14 + k) * v.ld + n] = e.y[10 + k];
#pragma unroll
    for (int k = 0; k < 3; ++k) v.prev_ang[k * v.ld + n] = e.prev_ang[k];
    v.prev_shaping[n] = e.prev_shaping;
    v.abs_sum[n] = e.abs_sum;
    v.ep_return[n] = e.ep_return;
    v.step_i[n] = e.i;
    v.flags[n] = (uint8_t)e.flags;
    v.episode[n] = e.episode;
}

// AUX attributes the single-env compatibility class exposes (quad.ang_vel, step_effort, w, accel,
// accelerometer_read, mat_rot); evaluated at the new state like the reference's trailing drone_eq call.
template <typename R, bool ROBUST = false>
__device__ __noinline__ void store_aux(const DevParams<R>& p, const SimView<R>& v, int64_t n, const Env<R> e,
                                       const StepOut<R> o, const Ctrl<R> c) {   // by VALUE: callers' structs stay in registers
#pragma unroll
    for (int k = 0; k < 3; ++k) v.ang_vel[k * v.ld + n] = o.ang_vel[k];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        v.step_effort[k * v.ld + n] = o.effort[k]; v.w[k * v.ld + n] = o.w[k];
        v.clipped_action[k * v.ld + n] = o.clipped[k]; v.fm[k * v.ld + n] = o.fm[k];
    }
    R dy[13], qn[4], r[9];
    drone_rhs<R, ROBUST>(p, c, e.y, dy);
    quat_normalize(&e.y[6], qn);
    quat_rot_mat(qn, r);
    R a[3] = {dy[1], dy[3], dy[5]};
#pragma unroll
    for (int k = 0; k < 3; ++k) v.accel[k * v.ld + n] = a[k];
    R g[3] = {a[0], a[1], a[2] - p.g};                                   // :371  R^T (accel + [0,0,-G])
    v.acc_read[0 * v.ld + n] = r[0] * g[0] + r[3] * g[1] + r[6] * g[2];
    v.acc_read[1 * v.ld + n] = r[1] * g[0] + r[4] * g[1] + r[7] * g[2];
    v.acc_read[2 * v.ld + n] = r[2] * g[0] + r[5] * g[1] + r[8] * g[2];
#pragma unroll
    for (int k = 0; k < 9; ++k) v.mat_rot[k * v.ld + n] = r[k];
}

// thread-local episode statistics, reduced warp -> block -> device accumulators
struct LocalStats {
    float v[7];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int k = 0; k < 7; ++k) v[k] = 0.f;
    }
};

__device__ __forceinline__ void flush_stats(const LocalStats& ls, bool any_local, double* stats) {
    __shared__ float s_acc[7];
    __shared__ int s_any;
    if (threadIdx.x == 0) s_any = 0;
    if (threadIdx.x < 7) s_acc[threadIdx.x] = 0.f;
    __syncthreads();
    const unsigned full = 0xffffffffu;
    if (__any_sync(full, any_local)) {
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            float x = ls.v[k];
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(full, x, d);
            if ((threadIdx.x & 31) == 0 && x != 0.f) atomicAdd(&s_acc[k], x);
        }
        if ((threadIdx.x & 31) == 0) s_any = 1;
    }
    __syncthreads();
    if (s_any && threadIdx.x < 7 && s_acc[threadIdx.x] != 0.f) atomicAdd(&stats[threadIdx.x], (double)s_acc[threadIdx.x]);
}

template <typename R>
__device__ __forceinline__ void count_episode(LocalStats& ls, const DevParams<R>& p, const Env<R>& e, const StepOut<R>& o) {
    ls.v[0] += (float)e.ep_return;
    ls.v[1] += (float)(e.i - p.T);
    ls.v[2] += 1.f;
    ls.v[3] += o.solved ? 1.f : 0.f;
    ls.v[4] += o.broken ? 1.f : 0.f;
    ls.v[5] += o.timeout ? 1.f : 0.f;
    ls.v[6] += (float)e.abs_sum;
}

// Sensor sub-pass for one env (QS_FLAG_SENSOR_NOISE).  mode 0: one step of the sensor model; mode 1: sensor.reset
// from the true state (end of an episode's warm-up) and pass the true observation through; mode 2: pass-through only.
template <typename R>
__device__ __forceinline__ void sensor_update_inl(const DevParams<R>& p, const SimView<R>& v, int64_t n, const Env<R>& e,
                                                  const Ctrl<R>& c, const R vq0, const R vq1, const R vq2, const R vq3, int mode) {
    R obs[14];
    if (mode == 0) {
        R s[kSensorStateDim], z[32], dy[13], qn[4], rot[9];
#pragma unroll
        for (int k = 0; k < 17; ++k) s[k] = v.sensor_state[k * v.ld + n];
        drone_rhs(p, c, e.y, dy);                                  // trailing drone_eq call: accel at the new state
        quat_normalize(&e.y[6], qn);
        quat_rot_mat(qn, rot);
        const R g[3] = {dy[1], dy[3], dy[5] - p.g};               // :371
        const R acc_read[3] = {rot[0] * g[0] + rot[3] * g[1] + rot[6] * g[2], rot[1] * g[0] + rot[4] * g[1] + rot[7] * g[2],
                               rot[2] * g[0] + rot[5] * g[1] + rot[8] * g[2]};
        sensor_normals(v.seed, v.env_id_offset + (uint32_t)n, e.episode, (uint32_t)e.i, p.s_gps_blend > R(0), z);
        sensor_step(p, z, e.y, acc_read, rot, c.f_m, s, obs);
#pragma unroll
        for (int k = 0; k < kSensorStateDim; ++k) v.sensor_state[k * v.ld + n] = s[k];
    } else {
        if (mode == 1) {
            R s[kSensorStateDim];
            sensor_reset(p, v.seed, v.env_id_offset + (uint32_t)n, e.episode, e.y, s);
#pragma unroll
            for (int k = 0; k < kSensorStateDim; ++k) v.sensor_state[k * v.ld + n] = s[k];
        }
#pragma unroll
        for (int k = 0; k < 10; ++k) obs[k] = e.y[k];
        obs[10] = vq0; obs[11] = vq1; obs[12] = vq2; obs[13] = vq3;
    }
#pragma unroll
    for (int k = 0; k < 14; ++k) v.sensed_obs[k * v.ld + n] = obs[k];
}

// out-of-line variant for cold paths (explicit / strict resets): arguments by value, callers' structs stay in registers
template <typename R>
__device__ __noinline__ void sensor_update(const DevParams<R>& p, const SimView<R>& v, int64_t n, const Env<R> e,
                                           const Ctrl<R> c, const R vq0, const R vq1, const R vq2, const R vq3, int mode) {
    sensor_update_inl(p, v, n, e, c, vq0, vq1, vq2, vq3, mode);
}

template <typename R> struct StepIO {
    const R* action;     // [4][N]
    R* obs;              // [14][N] or NULL
    R* reward;           // [N] or NULL
    uint8_t* done;       // [N] or NULL
    uint8_t* solved;     // [N] or NULL
};

constexpr int kResetQueueCap = 4096;

// K fused env steps per launch: state stays in registers, only actions/outputs stream through HBM.
template <typename R> struct RolloutIO {
    int32_t horizon;
    int32_t action_source;
    const R* actions;
    R* obs_out;
    R* action_out;
    R* reward_out;
    uint8_t* done_out;
    R* sensed_out;       // [K][14][N] or NULL (QS_FLAG_SENSOR_NOISE handles)
};


// ------------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------------
static inline int grid_for(const qs_sim* s, int64_t n) {
    int64_t blocks = (n + kBlock - 1) / kBlock;
    int64_t cap = (int64_t)s->sm_count * 8;           // grid-stride beyond 8 CTAs per SM ...
    int64_t need = (n + kResetQueueCap - 1) / kResetQueueCap;   // ... but a block never owns more envs than its reset queue holds
    if (cap < need) cap = need;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

#define QS_DISPATCH(s, FN, ...)                                                                          \
    do {                                                                                                 \
        const bool direct_ = ((s)->cfg.flags & QS_FLAG_DIRECT_CONTROL) != 0;                             \
        const bool rk45_ = (s)->cfg.integrator == QS_RK45;                                               \
        if ((s)->cfg.precision == QS_F32) {                                                              \
            if (rk45_) { if (direct_) FN<float, 1, true>(__VA_ARGS__); else FN<float, 1, false>(__VA_ARGS__); } \
            else       { if (direct_) FN<float, 0, true>(__VA_ARGS__); else FN<float, 0, false>(__VA_ARGS__); } \
        } else {                                                                                         \
            if (rk45_) { if (direct_) FN<double, 1, true>(__VA_ARGS__); else FN<double, 1, false>(__VA_ARGS__); } \
            else       { if (direct_) FN<double, 0, true>(__VA_ARGS__); else FN<double, 0, false>(__VA_ARGS__); } \
        }                                                                                                \
    } while (0)

template <typename R> inline const DevParams<R>& params_of(const qs_sim* s);
template <> inline const DevParams<float>& params_of<float>(const qs_sim* s) { return s->pf; }
template <> inline const DevParams<double>& params_of<double>(const qs_sim* s) { return s->pd; }

// the per-warp pipeline exists for the production configuration only: FP32, fixed-step RK4, no AUX rows, resets
// either asynchronous or none (strict lock-step resets run T serial hover steps per env and keep the CTA-wide kernel)
template <typename R, int INTEG> struct WarpKernelOk { static constexpr bool value = false; };
template <> struct WarpKernelOk<float, 0> { static constexpr bool value = true; };

// Launches go to the handle's device even when the calling thread's current device is another one (one process driving
// several GPUs); the caller's current device is restored when the entry point returns.
struct QsDeviceGuard {
    int prev = -1;
    bool switched = false;
    cudaError_t err = cudaSuccess;
    explicit QsDeviceGuard(int dev) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) { err = cudaSetDevice(dev); switched = (err == cudaSuccess); }
    }
    ~QsDeviceGuard() { if (switched) cudaSetDevice(prev); }
    QsDeviceGuard(const QsDeviceGuard&) = delete;
    QsDeviceGuard& operator=(const QsDeviceGuard&) = delete;
};
#define QS_USE_DEVICE(h)                                                               \
    QsDeviceGuard dev_guard_((h)->cfg.device);                                         \
    if (dev_guard_.err != cudaSuccess) return fail(QS_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(dev_guard_.err))

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (call site, device): function attributes are per context.  The bit
// set is atomic so that two host threads driving two handles / GPUs may race here (setting the attribute twice is harmless).
#define QS_SET_SMEM_ONCE(s, KERNEL, BYTES)                                                              \
    do {                                                                                                \
        static std::atomic<uint64_t> done_{0};                                                          \
        const uint64_t bit_ = 1ull << ((s)->cfg.device & 63);                                           \
        if (!(done_.load(std::memory_order_acquire) & bit_)) {                                          \
            cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(BYTES));    \
            done_.fetch_or(bit_, std::memory_order_release);                                            \
        }                                                                                               \
    } while (0)

// the FP32 / RK4 production kernels live in their own translation units; both return false when the handle's configuration
// is not theirs (the caller then launches the generic kernel)
bool launch_step_fast(qs_sim* s, const StepIO<float>& io, bool direct, bool sensor, cudaStream_t st);       // step_fast.cu
bool launch_rollout_fast(qs_sim* s, const qs_rollout_args* a, bool direct, cudaStream_t st);                // rollout_fast.cu
