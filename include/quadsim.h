/*
 * quadsim.h — C ABI of the B200-native batched quadrotor simulator (libquadsim.so).
 *
 * Drop-in boundary for ONE hot path of rafaelcostafrf/autonomous_quadrotor_environment:
 * `quad.reset` / `quad.step` of environment/quadrotor_env.py (rotor map -> drone_eq -> per-step
 * ODE integration -> quaternion/Euler conversion -> done/reward), re-designed for N independent
 * environments advanced in lock-step, one CUDA thread per environment.
 *
 * The reference has no FFI of its own (its boundary is the Python class API);
 * every entry point below names the reference interface it replaces (paths relative to the
 * reference root).  Python/PyTorch host code binds these with ctypes.
 *
 * Conventions
 *  - return 0 (QS_OK) on success, a negative QS_E* code otherwise; qs_last_error() gives the
 *    thread-local message of the last failure.  No exceptions cross the boundary.
 *  - every data pointer is a DEVICE pointer owned by the caller unless the name ends in `_host`;
 *    element type is float for QS_F32 handles and double for QS_F64 handles.
 *  - layout is structure-of-arrays with the env index fastest: a buffer documented as [C][N] holds
 *    channel c of env n at  ptr[c*N + n].
 *  - calls are asynchronous and ordered on the `stream` argument (a cudaStream_t passed as void*,
 *    NULL = legacy default stream).  There is no host synchronisation inside qs_step: resets are
 *    handled by a predicated sub-pass inside the kernel.
 *  - a handle is not thread-safe (one host thread per handle / GPU).
 *  - there is NO CPU fallback: every entry point that computes returns QS_ECUDA without a device.
 */
#ifndef QUADSIM_H
#define QUADSIM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QS_VERSION 100

/* error codes */
#define QS_OK 0
#define QS_EINVAL (-1)   /* bad argument */
#define QS_ECUDA (-2)    /* CUDA runtime error (no device, launch failure, ...) */
#define QS_ENOMEM (-3)   /* device allocation failed */
#define QS_ESTATE (-4)   /* call not valid for this handle's configuration */

/* precision (arithmetic AND boundary element type) */
#define QS_F32 0
#define QS_F64 1

/* integrator */
#define QS_RK4 0    /* fixed-step classical RK4, cfg.substeps sub-intervals (FP32 production mode)        */
#define QS_RK45 1   /* per-thread replica of scipy.integrate.solve_ivp's default RK45 as the reference     */
                    /* calls it at environment/quadrotor_env.py:483 (rtol 1e-3, atol 1e-6) — parity mode   */

/* qs_config.flags */
#define QS_FLAG_DIRECT_CONTROL 0x01u /* quad(direct_control=1): actions are 4 normalised rotor thrusts       */
#define QS_FLAG_CLIPPED        0x02u /* quad(clipped=True): indirect-mode mixer clips rotor speeds^2          */
#define QS_FLAG_TRAINING       0x04u /* quad(training=True): "solved" ends the episode                        */
#define QS_FLAG_AUTO_RESET     0x08u /* re-sample + T warm-up steps inside the step kernel when done          */
#define QS_FLAG_SENSOR_NOISE   0x10u /* run the `sensor` model (quadrotor_env.py:579-724) each step           */
#define QS_FLAG_AUX            0x20u /* also store ang_vel, step_effort, w, accel, mat_rot, accelerometer_read */
#define QS_FLAG_ASYNC_RESET    0x40u /* auto-reset with ASYNCHRONOUS warm-up (production mode): an env that returns   */
                                     /* done is re-sampled at the end of that step (the observation returned with done */
                                     /* is the new episode's initial observation), and the T hover steps of quad.reset */
                                     /* (:447-453) run as its next T ordinary lock-step steps (caller's action         */
                                     /* ignored, reward 0, bit1 of the done byte set).  Per-env sequences are exactly  */
                                     /* those of reset()+step(); no lane ever runs T serial steps.                     */
#define QS_FLAG_ROBUST         0x80u /* quad.robust_control = True (quadrotor_env.py:84-109,:183): per-episode perturbation   */
                                     /* of rotor thrust (K_F), mass, rotor inertia and J plus linearly ramping wind gusts,    */
                                     /* every draw a pure function of (seed, global env id, episode / gust counter).  Runs on */
                                     /* the generic step / reset kernels (any precision / integrator); not available with     */
                                     /* QS_FLAG_SENSOR_NOISE, qs_rollout or qs_policy_rollout (qs_control_rollout supports it). */

/* Physical / reward constants.  Defaults (qs_default_config) = environment/quadrotor_env.py:30-80. */
typedef struct qs_params {
    double mass, gravity;              /* M, G                        :39          */
    double rho, c_d;                   /* RHO, C_D                    :42,45       */
    double k_f, k_m, i_r, t2wr;        /* K_F, K_M, I_R, T2WR         :48-51       */
    double j[3];                       /* diag(J)                     :55-57       */
    double arm;                        /* D                           :60          */
    double beam_thickness;             /* BEAM_THICKNESS              :63          */
    double bb_vel, bb_ang, bb_pos;     /* BB_VEL, BB_ANG, BB_POS      :31-34       */
    double solved_reward, broken_reward, shaping_weight;   /*         :70-72       */
    double shaping_internal_weights[3];                    /*         :73          */
    double p_c;                        /* P_C                         :76          */
    double tr[3], tr_p[3];             /* TR, TR_P                    :80-81       */
    /* sensor model defaults: sensor.__init__ :587-591 */
    double accel_std, accel_bias_drift, gyro_std, gyro_bias_drift;
    double magnet_std, magnet_bias_drift, gps_std_p, gps_std_v;
    double gps_blend;                  /* GPS_P of visual_landing/math_trajectory.py:71-77: per cent of the GPS reading blended into the
                                          dead-reckoned position/velocity each step (and written back); 0 = off (the script's GPS = False) */
    /* robust_control.__init__ :85-93 (QS_FLAG_ROBUST): D_KF, D_KM (unused by the reference too), D_M, D_IR, D_J, gust_std, gust_period */
    double robust_d_kf, robust_d_km, robust_d_m, robust_d_ir;
    double robust_d_j[3], robust_gust_std[3];
    int32_t robust_gust_period, reserved_;
} qs_params;

/* Mirrors quad.__init__(t_step, n, training, euler, direct_control, T, clipped)  quadrotor_env.py:112 */
typedef struct qs_config {
    int64_t  n_envs;          /* environments owned by this handle (this GPU's shard)                      */
    int64_t  env_id_offset;   /* global id of local env 0; Philox streams are keyed by GLOBAL env id        */
    double   t_step;          /* quad(t_step)                                                               */
    int32_t  n_max;           /* quad(n): step limit (the reference adds T internally, :157)                */
    int32_t  T;               /* quad(T): warm-up hover steps performed by reset (:447-453)                 */
    int32_t  substeps;        /* RK4 sub-intervals per env step (>=1); ignored by QS_RK45                   */
    int32_t  precision;       /* QS_F32 | QS_F64                                                            */
    int32_t  integrator;      /* QS_RK4 | QS_RK45                                                           */
    uint32_t flags;           /* QS_FLAG_*                                                                  */
    uint64_t seed;            /* Philox key (quad.seed, :189-193)                                           */
    int32_t  device;          /* CUDA device ordinal                                                        */
    int32_t  reserved;
    void*    workspace;       /* optional caller-owned device memory (>= qs_workspace_bytes) so that the    */
                              /* host framework's allocator owns the bytes; NULL -> cudaMalloc              */
    qs_params params;
} qs_config;

typedef struct qs_sim* qs_handle;

/* Fields addressable through qs_get / qs_set / qs_field_info.  Names follow the attributes of the
 * reference `quad` object that callers read. */
typedef enum qs_field {
    QS_FIELD_OBS = 0,         /* [14][N] real  quat_state = state[0:10] ++ V_q          :486             */
    QS_FIELD_STATE = 1,       /* [13][N] real  [x,vx,y,vy,z,vz,q0..q3,wx,wy,wz]         :399-405,:485    */
    QS_FIELD_ANG = 2,         /* [3][N]  real  Euler angles of the normalised quaternion :488-491        */
    QS_FIELD_ANG_VEL = 3,     /* [3][N]  real  finite-difference Euler rates (AUX)       :492             */
    QS_FIELD_STEP_EFFORT = 4, /* [4][N]  real  (AUX)                                     :474,:476,:243  */
    QS_FIELD_W = 5,           /* [4][N]  real  rotor speeds rad/s (AUX)                  :288,:476       */
    QS_FIELD_REWARD = 6,      /* [N]     real  reward of the last step                   :511-573        */
    QS_FIELD_DONE = 7,        /* [N]     u8    bit0 = done returned by the last step (:498); bit1 = the step was an  */
                              /*               asynchronous warm-up step (QS_FLAG_ASYNC_RESET only)                  */
    QS_FIELD_SOLVED = 8,      /* [N]     u8    quad.solved                               :564            */
    QS_FIELD_I = 9,           /* [N]     i32   quad.i step counter                       :467            */
    QS_FIELD_ABS_SUM = 10,    /* [N]     real  accumulated control effort                :575-577        */
    QS_FIELD_PREV_SHAPING = 11,/*[N]     real  reward shaping memory                     :545-547        */
    QS_FIELD_EP_RETURN = 12,  /* [N]     real  sum of rewards since the last reset                        */
    QS_FIELD_EPISODE = 13,    /* [N]     u32   episode counter (Philox counter word)                      */
    QS_FIELD_FLAGS = 14,      /* [N]     u8    bit0 sticky done (:509), bit1 prev_shaping valid, bit2 solved,        */
                              /*               bits 3..7 warm-up steps still owed (QS_FLAG_ASYNC_RESET)               */
    QS_FIELD_ACCEL = 15,      /* [3][N]  real  inertial acceleration (AUX)               :368            */
    QS_FIELD_ACC_READ = 16,   /* [3][N]  real  accelerometer_read (AUX)                  :371            */
    QS_FIELD_MAT_ROT = 17,    /* [9][N]  real  body->inertial rotation, row-major (AUX)  :315            */
    QS_FIELD_SENSED_OBS = 18, /* [14][N] real  sensor-based observation (SENSOR_NOISE)   visual_landing/rl_worker.py:171-174 */
    QS_FIELD_SENSOR_STATE = 19,/*[QS_SENSOR_STATE_DIM][N] real (SENSOR_NOISE)                             */
    QS_FIELD_CLIPPED_ACTION = 20,/*[4][N] real quad.clipped_action (AUX)                  :472,:477          */
    QS_FIELD_FM = 21,         /* [4][N]  real  body thrust + moments applied [F,Mx,My,Mz] (AUX) :287-291        */
    QS_FIELD_GUST_COUNT = 22, /* [N]     i32   gusts drawn so far by robust_control.wind (ROBUST)  :104-109          */
    QS_FIELD_COUNT_
} qs_field;

#define QS_SENSOR_STATE_DIM 20

typedef struct qs_field_desc {
    int32_t channels;     /* C of [C][N]                                             */
    int32_t elem_bytes;   /* 1, 4 or 8                                               */
    int64_t ld;           /* element stride between channels inside the handle       */
    void*   ptr;          /* handle-owned device pointer (zero-copy view), or NULL   */
    int64_t ws_offset;    /* byte offset of ptr inside the workspace                 */
} qs_field_desc;

/* Episode statistics accumulated on the device by warp-shuffle/block reductions inside the step
 * kernels (replaces the host-side aggregation of worker results, environment/controller/ppo.py:371-382). */
#define QS_STATS_DIM 8
typedef struct qs_stats {
    double sum_return;    /* sum over finished episodes of the episode return               */
    double sum_length;    /* sum of episode lengths (steps after the T warm-up steps)       */
    double n_episodes;
    double n_solved;
    double n_broken;      /* ended by a bounding-box breach                                 */
    double n_timeout;     /* ended by i >= n                                                */
    double sum_effort;    /* sum of quad.abs_sum at episode end                             */
    double n_steps;       /* env steps executed                                             */
} qs_stats;

/* Optional fused rollout (K env steps per launch, state kept in registers). */
#define QS_ACT_BUFFER 0          /* actions read from `actions` [K][4][N]                          */
#define QS_ACT_PHILOX_UNIFORM 1  /* a ~ U(-1,1)^4 drawn in-kernel (BASELINE.json configs[2])        */
typedef struct qs_rollout_args {
    int32_t horizon;             /* K                                                              */
    int32_t action_source;       /* QS_ACT_*                                                       */
    const void* actions;         /* [K][4][N] or NULL                                              */
    void* obs_out;               /* [K][14][N] or NULL                                             */
    void* action_out;            /* [K][4][N] or NULL                                              */
    void* reward_out;            /* [K][N] or NULL                                                 */
    uint8_t* done_out;           /* [K][N] or NULL                                                 */
    void* sensed_obs_out;        /* [K][14][N] or NULL: the sensor-based observation of every step  */
                                 /* (QS_FLAG_SENSOR_NOISE handles; visual_landing/rl_worker.py:171) */
} qs_rollout_args;

/* Actor of the reference's PPO controller (environment/controller/model.py:27-34): Linear(in_dim,H)-Tanh-Linear(H,H)-
 * Tanh-Linear(H,4)-Tanh with a fixed-sigma diagonal Normal (model.py:60-66).  FP32 device pointers, PyTorch Linear
 * layout [out][in] (the keys actor.{0,2,4}.{weight,bias} of the shipped solved/*.pth files). */
typedef struct qs_actor {
    const float* w1; const float* b1;   /* [hidden][in_dim], [hidden] */
    const float* w2; const float* b2;   /* [hidden][hidden], [hidden] */
    const float* w3; const float* b3;   /* [4][hidden], [4]           */
    int32_t hidden;                     /* 128 */
    int32_t in_dim;                     /* 75 = 15 floats x history T=5 (environment/controller/dl_auxiliary.py:15-23) */
    float   action_std;                 /* sigma; <= 0 -> deterministic (evaluation, ppo_quad_eval.py:53) */
    int32_t reserved;
    /* optional critic (model.py:36-43: Linear(in_dim,H)-Tanh-Linear(H,H)-Tanh-Linear(H,1); keys critic.{0,2,4}.{weight,bias}): when
     * cw1 is not NULL qs_policy_rollout also evaluates it on every network input (same history tile, same tensor-core path) and
     * can record the state values PPO's memory.values holds (model.py:68). */
    const float* cw1; const float* cb1; /* [hidden][in_dim], [hidden] */
    const float* cw2; const float* cb2; /* [hidden][hidden], [hidden] */
    const float* cw3; const float* cb3; /* [1][hidden], [1]           */
} qs_actor;

/* Fused PPO rollout (BASELINE.json configs[4]): per step  history -> actor MLP (tcgen05 tensor cores, BF16 operands,
 * FP32 accumulate) -> a ~ N(mean, sigma) (Philox) -> quad.step -> history push, K steps per launch with the env state in
 * registers.  Replaces the per-step loop of environment/controller/ppo.py:238-257.  Output buffers are [K][C][N].
 * The output layer's weights are staged in constant memory by a stream-ordered copy at every call: rollouts with DIFFERENT
 * actors must not run concurrently on different streams of one device. */
typedef struct qs_policy_rollout_args {
    int32_t horizon;
    int32_t reserved;
    void* obs_out;        /* [K][14][N] float or NULL */
    void* action_out;     /* [K][4][N]  float or NULL : sampled (unclipped) actions                     */
    void* logprob_out;    /* [K][4][N]  float or NULL : per-dimension log-probabilities (model.py:66)   */
    void* reward_out;     /* [K][N]     float or NULL */
    uint8_t* done_out;    /* [K][N]     or NULL : bit0 done, bit1 warm-up step                          */
    void* hist;           /* [75][N] float in/out: dl_in_gen.deep_learning_input per env, oldest first; NULL = zeros */
    void* value_out;      /* [K+1][N]   float or NULL : critic value of the network input of step t (needs qs_actor.cw1..cb3); row K = */
                          /*                            value of the input after the last step (the GAE bootstrap, ppo.py:125-141)      */
    void* sensed_obs_out; /* [K][14][N] float or NULL : QS_FLAG_SENSOR_NOISE handles — there the HISTORY takes the sensed observation  */
                          /*                            (the policy flies on sensor_sp's states_sens, visual_landing/rl_worker.py:164) */
} qs_policy_rollout_args;

/* Classical comparison controllers of the reference as in-kernel control laws: LQR
 * (environment/controller/lqr_quad.py:129-157) and the cascaded velocity/attitude PID (environment/controller/
 * pid_vel_control.py:29-127).  Both command [F, Mx, My, Mz]: indirect-control (direct_control=0) handles only. */
#define QS_CTRL_LQR 0
#define QS_CTRL_PID 1
typedef struct qs_controller {
    int32_t kind;              /* QS_CTRL_*                                                                     */
    int32_t reserved;
    double k_t[3][6];          /* LQR translational gain K_t   lqr_quad.py:107-111 (solution of an ARE: computed by the host) */
    double k_att[4][6];        /* LQR attitude gain K_att      lqr_quad.py:82-86                                 */
    double pid_xy[3], pid_z[3], pid_att[3], pid_psi[3];   /* P, I, D of the velocity / attitude loops  pid_vel_control.py:17-27 */
    double target_vel[3];      /* velocity set-point xd        pid_vel_control.py:150                            */
    double target_psi;         /* yaw set-point psd                                                              */
    double pid_ts;             /* time step of class pid (its default argument, 0.01, :114); <= 0 -> 0.01        */
} qs_controller;

/* Controller memory per env, [QS_CTRL_STATE_DIM][N]: rows 0..2 quad.ang_vel as of the last step (the LQR's rate feedback),
 * 3..8 x_old and 9..14 ix of the six scalar PIDs [vx, vy, vz, phi, theta, psi], 15..17 ang_d_ant, 18..21 the PID action
 * computed after the previous step and applied at the next (pid_vel_control.py:144-153; [M*G,0,0,0] for a fresh controller). */
#define QS_CTRL_STATE_DIM 22
typedef struct qs_control_rollout_args {
    int32_t horizon;
    int32_t reserved;
    void* ctrl_state;          /* [QS_CTRL_STATE_DIM][N] in/out, or NULL = fresh controller, ang_vel = 0, not written back */
    void* obs_out;             /* [K][14][N] or NULL */
    void* action_out;          /* [K][4][N]  or NULL : the [F,Mx,My,Mz] command applied at step t               */
    void* reward_out;          /* [K][N]     or NULL */
    uint8_t* done_out;         /* [K][N]     or NULL */
    void* aux_out;             /* [K][10][N] or NULL : ang(3), ang_vel(3), step_effort(4) — with the velocities of obs_out the */
                               /*                      13 columns of the reference's classical_controller_results logs         */
    const void* target_traj;   /* [K][3] or NULL (PID only): velocity set-point in force at step t, e.g. the `velocity` array of */
                               /* a mission (mission_control/mission_control.py); NULL = the constant qs_controller.target_vel  */
} qs_control_rollout_args;

/* ---- lifecycle ----------------------------------------------------------------------------- */
/* Fill *cfg with the reference defaults (constants quadrotor_env.py:30-80; quad() keyword defaults :112). */
int qs_default_config(qs_config* cfg);
/* Bytes of device memory a handle built from *cfg needs (for caller-provided workspaces). */
int64_t qs_workspace_bytes(const qs_config* cfg);
/* quad.__init__  quadrotor_env.py:112-187 */
int qs_create(qs_handle* out, const qs_config* cfg);
int qs_destroy(qs_handle h);
/* quad.seed  :189-193 — re-keys the Philox streams */
int qs_seed(qs_handle h, uint64_t seed);

/* ---- the hot path --------------------------------------------------------------------------- */
/* quad.reset(det_state)  :408-454.  det_state [13][N] or NULL (random branch, Philox in place of the
 * NumPy global RNG); mask u8[N] or NULL (= all envs); obs_hist [T][14][N] / act_hist [T][4][N] or NULL. */
int qs_reset(qs_handle h, const void* det_state, const uint8_t* mask, void* obs_hist, void* act_hist, void* stream);
/* quad.step(action)  :458-498 (+ f2F :247-272 / f2w :197-245, drone_eq :274-406, solve_ivp :483,
 * quat_euler utility:39-48, done_condition :500-509, reward_function :511-573, control_effort :575-577).
 * action [4][N]; obs [14][N], reward [N], done u8[N], solved u8[N] may each be NULL (the results are
 * always available zero-copy through qs_field_info(QS_FIELD_OBS/REWARD/DONE/SOLVED)). */
int qs_step(qs_handle h, const void* action, void* obs, void* reward, uint8_t* done, uint8_t* solved, void* stream);
/* K fused env steps (see qs_rollout_args). */
int qs_rollout(qs_handle h, const qs_rollout_args* args, void* stream);
/* K fused policy steps (see qs_policy_rollout_args); FP32 / RK4 / direct-control handles only; with QS_FLAG_SENSOR_NOISE the sensor
 * model runs after every step and the actor's observation history is built from the SENSED observation. */
int qs_policy_rollout(qs_handle h, const qs_actor* actor, const qs_policy_rollout_args* args, void* stream);
/* PID gains of pid_vel_control.py:17-27 (clipped / not clipped); the LQR gains are left zero (host computes the AREs). */
int qs_default_controller(qs_controller* c, int kind, int clipped);
/* K fused steps driven by an in-kernel control law (see qs_controller); resets: none, or QS_FLAG_ASYNC_RESET (the
 * controller memory is cleared after every reset like `controller = pid_control(drone)`, pid_vel_control.py:143).
 * On QS_FLAG_AUX handles quad.ang_vel is read from / written back to the handle's ANG_VEL row (rows 0..2 of ctrl_state are
 * then ignored on input), so reset() -> control_rollout() behaves like the reference scripts. */
int qs_control_rollout(qs_handle h, const qs_controller* c, const qs_control_rollout_args* args, void* stream);
/* PPO.get_advantages (environment/controller/ppo.py:125-141) on time-major rollout buffers: backward GAE scan per env.
 * reward [K][N], value [K+1][N] (row K = bootstrap value of the state after the last step; the reference appends 0, :384),
 * done u8 [K][N] (bit0 = done -> mask 0; bit1 = asynchronous warm-up step -> not a transition).  Writes returns [K][N] and the
 * UNNORMALISED advantages [K][N], and ACCUMULATES {count, sum, sum of squares} of the valid advantages into moments[3]
 * (device doubles, zeroed by the caller; all-reduce them across ranks before normalising).  FP32. */
int qs_gae(int64_t n_envs, int32_t horizon, float gamma, float lambda, const float* reward, const float* value,
           const uint8_t* done, float* returns_out, float* adv_out, double* moments, void* stream);
/* ppo.py:141  adv <- (adv - mean) / (std + 1e-10) in place (population std), 0 for warm-up steps; weight (nullable) <- 1/0. */
int qs_adv_normalize(int64_t total, const uint8_t* done, const double* moments, float* adv, float* weight, void* stream);
/* The network update of the reference's PPO trainer (environment/controller/ppo.py:143-209, model.py:19-88) on the time-major rollout
 * buffers, hand-written for sm_100a: forward AND backward of one 75-128-128-{4,1} network on tcgen05 tensor cores (BF16 operands, FP32
 * accumulation in tensor memory), one full-batch gradient per call.  All pointers device, FP32. */
typedef struct qs_ppo_batch {
    int64_t n_envs;               /* N                                                                                         */
    int32_t horizon;              /* K recorded steps                                                                          */
    int32_t flags;                /* QS_PPO_RECORD_LOGP or 0                                                                    */
    const float* hist0;           /* [75][N]     dl_in_gen buffer at rollout start, oldest entry first (dl_auxiliary.py:15-23)  */
    const float* entries;         /* [K][15][N]  the entry pushed after step t: action(4), v(3), q(4), dq(4) (dl_auxiliary.py:27-30);
                                                    NULL: taken from `actions` and rows 1,3,5,6..13 of `obs` (no materialised copy)     */
    const float* obs;             /* [K][14][N]  observations recorded by qs_policy_rollout (read when entries == NULL)                */
    const float* actions;         /* [K][4][N]   memory.actions (actor)                                                        */
    float* logp_old;              /* [K][4][N]   memory.logprobs (actor); WRITTEN by the call under QS_PPO_RECORD_LOGP               */
    const float* adv;             /* [K][N]      normalised advantages (actor; qs_gae + qs_adv_normalize)                      */
    const float* ret;             /* [K][N]      returns (critic)                                                              */
    const float* weight;          /* [K][N]      1 = transition, 0 = warm-up step of an asynchronous reset                     */
} qs_ppo_batch;
typedef struct qs_ppo_net {       /* one network, PyTorch Linear layout [out][in]: (128,75) (128) (128,128) (128) (OUT,128) (OUT) */
    const float* w1; const float* b1; const float* w2; const float* b2; const float* w3; const float* b3;
} qs_ppo_net;
/* flags: the call's own forward pass supplies memory.logprobs — the reference's policy_old is an exact copy of policy when the update
 * starts (ppo.py:206), so the first epoch's ratio is exactly 1; the per-dimension log-probs are stored to logp_old for the later epochs */
#define QS_PPO_RECORD_LOGP 1
#define QS_PPO_ACTOR 0            /* OUT = 4, tanh output; loss = -min(r A, clip(r, 1 -+ eps) A), r = exp(sum logp - sum logp_old)  ppo.py:187-195 */
#define QS_PPO_CRITIC 1           /* OUT = 1; loss = 0.5 (V - R)^2                                                           ppo.py:194     */
/* ACCUMULATES d(sum of weight * loss / count)/d(parameters) into *grad (same six shapes as *net; zeroed by the caller — gradients of
 * several ranks or batches add up) and the summed weighted loss / count into *loss_sum (device double, nullable).  count = the GLOBAL
 * number of valid transitions (loss.mean() of ppo.py:203).  The history at any step is a window of `entries`, so the launch splits
 * every 128-env tile's steps into chunks for load balance; summation order (FP32 atomics) is not fixed from run to run. */
int qs_ppo_grad(const qs_ppo_batch* batch, const qs_ppo_net* net, const qs_ppo_net* grad, int which, float sigma, float eps_clip,
                double count, double* loss_sum, void* stream);
/* torch.optim.Adam.step (the optimizer of ppo.py:105; eps 1e-8 there) on one flat FP32 parameter vector; step = 1, 2, ... */
int qs_adam_step(int64_t n, float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int32_t step, float lr, float beta1,
                 float beta2, float eps, void* stream);
/* Same contract as qs_step but with HOST buffers (pinned or pageable): H2D of the actions, the step
 * kernel and D2H of obs/reward/done are enqueued on `stream` and the call returns after they finish. */
int qs_step_host(qs_handle h, const void* action_host, void* obs_host, void* reward_host, uint8_t* done_host, void* stream);

/* Select the implementation of qs_step for this handle (all produce the same results to FP32 rounding): 0 = plain coalesced
 * loads, 1 = CTA-wide TMA-staged ring, 2 = per-warp cp.async pipeline with 16-byte vector loads/stores (one env per lane),
 * 3 = per-warp pipeline with two envs per lane, the RK4 stages on the packed FP32 instructions (FFMA2).  Default for
 * FP32/RK4 handles: 3 (with or without QS_FLAG_SENSOR_NOISE); other configurations fall back to 1.  Tuning / A-B testing
 * only.  qs_get_step_loader returns the loader a handle will use. */
int qs_get_step_loader(qs_handle h);
int qs_set_step_loader(qs_handle h, int loader);

/* ---- state access --------------------------------------------------------------------------- */
int qs_field_info(qs_handle h, qs_field f, qs_field_desc* out);
int qs_get(qs_handle h, qs_field f, void* dst /* device, [C][N] contiguous */, void* stream);
int qs_set(qs_handle h, qs_field f, const void* src /* device, [C][N] contiguous */, void* stream);

/* ---- statistics ----------------------------------------------------------------------------- */
/* Device pointer to the QS_STATS_DIM doubles (for an in-place NCCL all-reduce by the host framework). */
int qs_stats_device(qs_handle h, double** dptr);
/* Copy the accumulators to host (synchronises `stream`), optionally zeroing them afterwards. */
int qs_stats_read(qs_handle h, qs_stats* out_host, int reset_after, void* stream);

/* ---- stateless batched device functions (unit-testable pieces of the path) ------------------ */
/* precision: QS_F32/QS_F64; all pointers device, SoA [C][n]. */
int qs_euler_quat(int precision, int64_t n, const void* ang /*[3][n]*/, void* q /*[4][n]*/, void* stream);      /* utility:17-36 */
int qs_quat_euler(int precision, int64_t n, const void* q /*[4][n]*/, void* ang /*[3][n]*/, void* stream);      /* utility:39-48 */
int qs_deriv_quat(int precision, int64_t n, const void* w /*[3][n]*/, const void* q, void* dq, void* stream);   /* utility:58-69 */
int qs_quat_rot_mat(int precision, int64_t n, const void* q, void* R /*[9][n] row-major*/, void* stream);       /* utility:71-80 */
/* drone_eq :274-406.  direct!=0: action = normalised rotor thrusts (f2F); else action = [F,Mx,My,Mz] and
 * w_rotor [4][n] are the rotor speeds f2w left in self.w. */
int qs_drone_eq(int precision, const qs_params* p, int64_t n, int direct, const void* x /*[13][n]*/,
                const void* action /*[4][n]*/, const void* w_rotor /*[4][n] or NULL*/, void* dx /*[13][n]*/, void* stream);
/* f2w :197-245 -> step_effort [4][n], w [4][n], FM_new [4][n] */
int qs_f2w(int precision, const qs_params* p, int64_t n, int clipped, const void* fm /*[4][n]*/,
           void* step_effort, void* w, void* fm_new, void* stream);
/* Philox4x32-10 raw blocks: out[4][n] u32 for counter (env_id0+i, episode, block, stream_id). */
int qs_philox_raw(uint64_t seed, int64_t env_id0, int64_t n, uint32_t episode, uint32_t block, uint32_t stream_id,
                  uint32_t* out /*[4][n]*/, void* stream);

/* The reference's `sensor` class (environment/quadrotor_env.py:579-724), one METHOD per call, for n independent sensors whose random
 * draws are supplied by the caller: z [k][n] holds the STANDARD-normal draws the method's np.random.normal calls consume, in call
 * order (QS_SENSOR_RESET: the three U(0,1) draws of bias_reset :600-608).  This is the entry point behind the single-env drop-in
 * `sensor` (which draws from the global NumPy stream in the reference's order, so a seeded script sees the reference's readings) and
 * what pins the in-kernel sensor model (QS_FLAG_SENSOR_NOISE: same device functions, Philox draws) to the reference class.
 *   sensor_state [QS_SENSOR_STATE_DIM][n] in/out: 0 a_b_accel, 1 g_b, 2 a_b_d, 3 g_b_d, 4..6 velocity_t0, 7..9 position_t0,
 *                10..13 quaternion_t0, 14..16 third column of self.R (the part :658 reads), 17..19 acceleration_t0
 *   quad_state [13][n] = quad.state, acc_read [3][n] = quad.accelerometer_read (:371), mat_rot [9][n] = quad.mat_rot row-major (:315),
 *   f_m [n] = quad.f_in[2] / M (:658); each may be NULL when the method does not read it.
 *   method            z rows  out rows
 *   QS_SENSOR_RESET     3       0    sensor.reset :630-640 (+ bias_reset); reads quad_state
 *   QS_SENSOR_ACCEL     3       3    sensor.accel :611-620 -> accelerometer reading
 *   QS_SENSOR_GYRO      3       3    sensor.gyro :622-628 -> gyro reading
 *   QS_SENSOR_GPS       6       6    sensor.gps :642-647 -> position(3), velocity(3)
 *   QS_SENSOR_TRIAD     6      13    sensor.triad :649-697 -> q scalar-first (4), R row-major (9); z = accel(3), magnetometer(3)
 *   QS_SENSOR_ACCEL_INT 9      18    sensor.accel_int :700-715 -> acceleration(3), velocity(3), position(3), self.R row-major (9);
 *                                    z = accel(3), triad(6)
 *   QS_SENSOR_GYRO_INT  3       4    sensor.gyro_int :717-724 -> q before normalisation
 *   QS_SENSOR_STEP     27      14    accel_int, gyro_int, gyro, gps, triad in the order of every caller (visual_landing/rl_worker.py:
 *                                    164-175, math_trajectory.py:61-83 incl. the GPS blend p->gps_blend) -> sensed observation */
#define QS_SENSOR_RESET 0
#define QS_SENSOR_ACCEL 1
#define QS_SENSOR_GYRO 2
#define QS_SENSOR_GPS 3
#define QS_SENSOR_TRIAD 4
#define QS_SENSOR_ACCEL_INT 5
#define QS_SENSOR_GYRO_INT 6
#define QS_SENSOR_STEP 7
int qs_sensor_call(int precision, const qs_params* p, double t_step, int64_t n, int method, void* sensor_state,
                   const void* quad_state, const void* acc_read, const void* mat_rot, const void* f_m, const void* z, void* out,
                   void* stream);

/* tcgen05 self-test: D[128][N] = A[128][K] * B[N][K]^T with BF16 operands / FP32 accumulation through the same
 * shared-memory operand layout, descriptors and TMEM read-back the fused actor rollout uses (row-major fp32 in/out). */
int qs_umma_selftest(int N, int K, const float* A, const float* B, float* D, void* stream);
/* The same product with the A operand staged in TENSOR memory (tcgen05.st by the thread that owns the row, TS form of
 * tcgen05.mma): 16 <= N <= 128 (multiple of 16), 32 <= K <= 128 (multiple of 32). */
int qs_umma_selftest_ts(int N, int K, const float* A, const float* B, float* D, void* stream);
/* The MN-major operand path of qs_ppo_grad (operands stored with K = the row index, so one tile serves a forward product and a
 * weight-gradient product without a transpose), K = 128: mode 0: D[128][N] = At^T Bt with At [128 k][128 m], Bt [128 k][N];
 * mode 1: D[128][N] = A Bt with A [128 m][128 k] K-major, Bt [128 k][N] MN-major. */
int qs_umma_selftest_mn(int mode, int N, const float* A, const float* B, float* D, void* stream);

/* ---- misc ----------------------------------------------------------------------------------- */
const char* qs_last_error(void);
int qs_version(void);
/* peak-FP32 micro-benchmark: every thread runs 8 independent chains of `iters` dependent FFMAs
 * (FLOPs = 2 * 8 * iters * blocks * threads); *ms_out = elapsed ms of one launch, so that bench.py can
 * state the FP32 roof it compares with. */
int qs_fp32_peak_probe(int blocks, int threads, int iters, float* ms_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QUADSIM_H */
