"""ctypes binding of libquadsim.so (the C ABI declared in include/quadsim.h).

There is deliberately no fallback: if the CUDA library is missing the import of any compute entry point
raises, and every compute call returns QS_ECUDA without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QUADSIM_LIB") or os.path.join(_HERE, "_C", "libquadsim.so")   # override: tuning builds only

QS_OK, QS_EINVAL, QS_ECUDA, QS_ENOMEM, QS_ESTATE = 0, -1, -2, -3, -4
QS_F32, QS_F64 = 0, 1
QS_RK4, QS_RK45 = 0, 1
QS_FLAG_DIRECT_CONTROL = 0x01
QS_FLAG_CLIPPED = 0x02
QS_FLAG_TRAINING = 0x04
QS_FLAG_AUTO_RESET = 0x08
QS_FLAG_SENSOR_NOISE = 0x10
QS_FLAG_AUX = 0x20
QS_FLAG_ASYNC_RESET = 0x40
QS_FLAG_ROBUST = 0x80
QS_ACT_BUFFER, QS_ACT_PHILOX_UNIFORM = 0, 1
QS_STATS_DIM = 8
QS_SENSOR_STATE_DIM = 20

(QS_FIELD_OBS, QS_FIELD_STATE, QS_FIELD_ANG, QS_FIELD_ANG_VEL, QS_FIELD_STEP_EFFORT, QS_FIELD_W, QS_FIELD_REWARD,
 QS_FIELD_DONE, QS_FIELD_SOLVED, QS_FIELD_I, QS_FIELD_ABS_SUM, QS_FIELD_PREV_SHAPING, QS_FIELD_EP_RETURN,
 QS_FIELD_EPISODE, QS_FIELD_FLAGS, QS_FIELD_ACCEL, QS_FIELD_ACC_READ, QS_FIELD_MAT_ROT, QS_FIELD_SENSED_OBS,
 QS_FIELD_SENSOR_STATE, QS_FIELD_CLIPPED_ACTION, QS_FIELD_FM, QS_FIELD_GUST_COUNT, QS_FIELD_COUNT) = range(24)

# every symbol include/quadsim.h declares (tests check that the library exports all of them)
EXPORTED_SYMBOLS = [
    "qs_default_config", "qs_workspace_bytes", "qs_create", "qs_destroy", "qs_seed", "qs_reset", "qs_step",
    "qs_rollout", "qs_policy_rollout", "qs_control_rollout", "qs_default_controller", "qs_gae", "qs_adv_normalize", "qs_step_host", "qs_set_step_loader", "qs_get_step_loader", "qs_field_info", "qs_get", "qs_set", "qs_stats_device", "qs_stats_read",
    "qs_euler_quat", "qs_quat_euler", "qs_deriv_quat", "qs_quat_rot_mat", "qs_drone_eq", "qs_f2w", "qs_philox_raw",
    "qs_sensor_call", "qs_ppo_grad", "qs_adam_step", "qs_last_error", "qs_version", "qs_fp32_peak_probe", "qs_umma_selftest", "qs_umma_selftest_ts", "qs_umma_selftest_mn",
]


class qs_params(C.Structure):
    _fields_ = [
        ("mass", C.c_double), ("gravity", C.c_double), ("rho", C.c_double), ("c_d", C.c_double),
        ("k_f", C.c_double), ("k_m", C.c_double), ("i_r", C.c_double), ("t2wr", C.c_double),
        ("j", C.c_double * 3), ("arm", C.c_double), ("beam_thickness", C.c_double),
        ("bb_vel", C.c_double), ("bb_ang", C.c_double), ("bb_pos", C.c_double),
        ("solved_reward", C.c_double), ("broken_reward", C.c_double), ("shaping_weight", C.c_double),
        ("shaping_internal_weights", C.c_double * 3), ("p_c", C.c_double),
        ("tr", C.c_double * 3), ("tr_p", C.c_double * 3),
        ("accel_std", C.c_double), ("accel_bias_drift", C.c_double), ("gyro_std", C.c_double),
        ("gyro_bias_drift", C.c_double), ("magnet_std", C.c_double), ("magnet_bias_drift", C.c_double),
        ("gps_std_p", C.c_double), ("gps_std_v", C.c_double), ("gps_blend", C.c_double),
        ("robust_d_kf", C.c_double), ("robust_d_km", C.c_double), ("robust_d_m", C.c_double), ("robust_d_ir", C.c_double),
        ("robust_d_j", C.c_double * 3), ("robust_gust_std", C.c_double * 3),
        ("robust_gust_period", C.c_int32), ("reserved_", C.c_int32),
    ]


class qs_config(C.Structure):
    _fields_ = [
        ("n_envs", C.c_int64), ("env_id_offset", C.c_int64), ("t_step", C.c_double),
        ("n_max", C.c_int32), ("T", C.c_int32), ("substeps", C.c_int32), ("precision", C.c_int32),
        ("integrator", C.c_int32), ("flags", C.c_uint32), ("seed", C.c_uint64),
        ("device", C.c_int32), ("reserved", C.c_int32), ("workspace", C.c_void_p),
        ("params", qs_params),
    ]


class qs_field_desc(C.Structure):
    _fields_ = [("channels", C.c_int32), ("elem_bytes", C.c_int32), ("ld", C.c_int64), ("ptr", C.c_void_p),
                ("ws_offset", C.c_int64)]


class qs_stats(C.Structure):
    _fields_ = [("sum_return", C.c_double), ("sum_length", C.c_double), ("n_episodes", C.c_double),
                ("n_solved", C.c_double), ("n_broken", C.c_double), ("n_timeout", C.c_double),
                ("sum_effort", C.c_double), ("n_steps", C.c_double)]


class qs_rollout_args(C.Structure):
    _fields_ = [("horizon", C.c_int32), ("action_source", C.c_int32), ("actions", C.c_void_p),
                ("obs_out", C.c_void_p), ("action_out", C.c_void_p), ("reward_out", C.c_void_p),
                ("done_out", C.c_void_p), ("sensed_obs_out", C.c_void_p)]


class qs_actor(C.Structure):
    _fields_ = [("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p), ("w3", C.c_void_p),
                ("b3", C.c_void_p), ("hidden", C.c_int32), ("in_dim", C.c_int32), ("action_std", C.c_float),
                ("reserved", C.c_int32), ("cw1", C.c_void_p), ("cb1", C.c_void_p), ("cw2", C.c_void_p), ("cb2", C.c_void_p),
                ("cw3", C.c_void_p), ("cb3", C.c_void_p)]


class qs_policy_rollout_args(C.Structure):
    _fields_ = [("horizon", C.c_int32), ("reserved", C.c_int32), ("obs_out", C.c_void_p), ("action_out", C.c_void_p),
                ("logprob_out", C.c_void_p), ("reward_out", C.c_void_p), ("done_out", C.c_void_p), ("hist", C.c_void_p),
                ("value_out", C.c_void_p), ("sensed_obs_out", C.c_void_p)]


class qs_controller(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("k_t", (C.c_double * 6) * 3), ("k_att", (C.c_double * 6) * 4),
                ("pid_xy", C.c_double * 3), ("pid_z", C.c_double * 3), ("pid_att", C.c_double * 3), ("pid_psi", C.c_double * 3),
                ("target_vel", C.c_double * 3), ("target_psi", C.c_double), ("pid_ts", C.c_double)]


class qs_control_rollout_args(C.Structure):
    _fields_ = [("horizon", C.c_int32), ("reserved", C.c_int32), ("ctrl_state", C.c_void_p), ("obs_out", C.c_void_p),
                ("action_out", C.c_void_p), ("reward_out", C.c_void_p), ("done_out", C.c_void_p), ("aux_out", C.c_void_p),
                ("target_traj", C.c_void_p)]


class qs_ppo_batch(C.Structure):
    _fields_ = [("n_envs", C.c_int64), ("horizon", C.c_int32), ("flags", C.c_int32), ("hist0", C.c_void_p), ("entries", C.c_void_p), ("obs", C.c_void_p),
                ("actions", C.c_void_p), ("logp_old", C.c_void_p), ("adv", C.c_void_p), ("ret", C.c_void_p), ("weight", C.c_void_p)]


class qs_ppo_net(C.Structure):
    _fields_ = [("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p), ("w3", C.c_void_p), ("b3", C.c_void_p)]


QS_PPO_ACTOR, QS_PPO_CRITIC = 0, 1
QS_PPO_RECORD_LOGP = 1
QS_CTRL_LQR, QS_CTRL_PID = 0, 1
(QS_SENSOR_RESET, QS_SENSOR_ACCEL, QS_SENSOR_GYRO, QS_SENSOR_GPS, QS_SENSOR_TRIAD, QS_SENSOR_ACCEL_INT, QS_SENSOR_GYRO_INT,
 QS_SENSOR_STEP) = range(8)
SENSOR_Z_ROWS = (3, 3, 3, 6, 6, 9, 3, 27)        # rows of z / out per method (include/quadsim.h, qs_sensor_call)
SENSOR_OUT_ROWS = (0, 3, 3, 6, 13, 18, 4, 14)
QS_CTRL_STATE_DIM = 22


class QuadSimError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libquadsim error %d: %s" % (code, msg))
        self.code = code


_lib = None


def load_library():
    """Load libquadsim.so (built in-tree by __graft_entry__.build()).  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            "CUDA library %s not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(this package has no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i64, i32, u32, u64 = C.c_void_p, C.c_int64, C.c_int32, C.c_uint32, C.c_uint64
    P = C.POINTER
    sig = {
        "qs_default_config": (C.c_int, [P(qs_config)]),
        "qs_workspace_bytes": (i64, [P(qs_config)]),
        "qs_create": (C.c_int, [P(vp), P(qs_config)]),
        "qs_destroy": (C.c_int, [vp]),
        "qs_seed": (C.c_int, [vp, u64]),
        "qs_reset": (C.c_int, [vp, vp, vp, vp, vp, vp]),
        "qs_step": (C.c_int, [vp, vp, vp, vp, vp, vp, vp]),
        "qs_rollout": (C.c_int, [vp, P(qs_rollout_args), vp]),
        "qs_policy_rollout": (C.c_int, [vp, P(qs_actor), P(qs_policy_rollout_args), vp]),
        "qs_step_host": (C.c_int, [vp, vp, vp, vp, vp, vp]),
        "qs_default_controller": (C.c_int, [P(qs_controller), C.c_int, C.c_int]),
        "qs_control_rollout": (C.c_int, [vp, P(qs_controller), P(qs_control_rollout_args), vp]),
        "qs_gae": (C.c_int, [i64, C.c_int32, C.c_float, C.c_float, vp, vp, vp, vp, vp, vp, vp]),
        "qs_adv_normalize": (C.c_int, [i64, vp, vp, vp, vp, vp]),
        "qs_set_step_loader": (C.c_int, [vp, C.c_int]),
        "qs_get_step_loader": (C.c_int, [vp]),
        "qs_field_info": (C.c_int, [vp, C.c_int, P(qs_field_desc)]),
        "qs_get": (C.c_int, [vp, C.c_int, vp, vp]),
        "qs_set": (C.c_int, [vp, C.c_int, vp, vp]),
        "qs_stats_device": (C.c_int, [vp, P(vp)]),
        "qs_stats_read": (C.c_int, [vp, P(qs_stats), C.c_int, vp]),
        "qs_euler_quat": (C.c_int, [C.c_int, i64, vp, vp, vp]),
        "qs_quat_euler": (C.c_int, [C.c_int, i64, vp, vp, vp]),
        "qs_deriv_quat": (C.c_int, [C.c_int, i64, vp, vp, vp, vp]),
        "qs_quat_rot_mat": (C.c_int, [C.c_int, i64, vp, vp, vp]),
        "qs_drone_eq": (C.c_int, [C.c_int, P(qs_params), i64, C.c_int, vp, vp, vp, vp, vp]),
        "qs_f2w": (C.c_int, [C.c_int, P(qs_params), i64, C.c_int, vp, vp, vp, vp, vp]),
        "qs_philox_raw": (C.c_int, [u64, i64, i64, u32, u32, u32, vp, vp]),
        "qs_sensor_call": (C.c_int, [C.c_int, P(qs_params), C.c_double, i64, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp]),
        "qs_ppo_grad": (C.c_int, [P(qs_ppo_batch), P(qs_ppo_net), P(qs_ppo_net), C.c_int, C.c_float, C.c_float, C.c_double, vp, vp]),
        "qs_adam_step": (C.c_int, [i64, vp, vp, vp, vp, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float, vp]),
        "qs_umma_selftest_mn": (C.c_int, [C.c_int, C.c_int, vp, vp, vp, vp]),
        "qs_last_error": (C.c_char_p, []),
        "qs_version": (C.c_int, []),
        "qs_fp32_peak_probe": (C.c_int, [C.c_int, C.c_int, C.c_int, P(C.c_float), vp]),
        "qs_umma_selftest": (C.c_int, [C.c_int, C.c_int, vp, vp, vp, vp]),
        "qs_umma_selftest_ts": (C.c_int, [C.c_int, C.c_int, vp, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != QS_OK:
        raise QuadSimError(rc, load_library().qs_last_error().decode("utf-8", "replace"))


def default_config() -> qs_config:
    cfg = qs_config()
    check(load_library().qs_default_config(C.byref(cfg)))
    return cfg
