"""Drop-in for environment/quadrotor_env.py: the reference's single-environment `quad` class, backed by the
CUDA library (a BatchedQuad with N=1, FP64 + the SciPy-RK45 replica so that trajectories match the
reference's own NumPy/SciPy step to ~1e-12).

The controller scripts of the reference (environment/controller/{lqr_quad,pid_vel_control,ppo_quad_eval,
ppo}.py) drive this class unchanged through the overlay module compat/environment/quadrotor_env.py.

What is preserved:
  * constructor signature `quad(t_step, n, training=True, euler=0, direct_control=1, T=1, clipped=True)` (:112)
  * `seed`, `reset(det_state=None) -> ((T,14),(T,4))`, `step(action) -> ((1,14) float64, float, bool)`
  * attributes callers read: state, ang, ang_vel, step_effort, w, done, solved, reward, i, n, t_step, T,
    state_size, action_size, abs_sum (read/write), mass, gravity, J_mat, clipped_action, accel, mat_rot,
    f_in, accelerometer_read, target_state, current_state, quat_state, zero_control, previous_state,
    action_hist
  * quirks the shipped logs depend on: prev_ang survives reset, sticky done, first reward has no shaping
    term, asymmetric angular-rate clip, reset = T real hover steps, 12 RNG draws consumed by
    `robust_control.reset` at HEAD (disable with ``robust_rng_draws=False`` to reproduce the 2021 logs).
  * random initial states come from the global NumPy RNG in the reference's draw order (host-side RNG
    plumbing, so `env.seed(1)` reproduces the reference's episodes); the batched API uses Philox instead.
"""
from __future__ import annotations

import numpy as np

from . import _lib as L

## SIMULATION BOUNDING BOXES ## (reference :30-80; values only)
BB_POS = 5
BB_VEL = 10
BB_CONTROL = 9
BB_ANG = np.pi / 2
M, G = 1.03, 9.82
RHO = 1.2041
C_D = 1.1
K_F = 1.435e-5
K_M = 2.4086e-7
I_R = 5e-5
T2WR = 2
J = np.array([[16.83e-3, 0, 0], [0, 16.83e-3, 0], [0, 0, 28.34e-3]])
D = 0.26
BEAM_THICKNESS = 0.05
A_X = BEAM_THICKNESS * 2 * D
A_Y = BEAM_THICKNESS * 2 * D
A_Z = BEAM_THICKNESS * 2 * D * 2
A = np.array([[A_X, A_Y, A_Z]]).T
SOLVED_REWARD = 20
BROKEN_REWARD = -20
SHAPING_WEIGHT = 5
SHAPING_INTERNAL_WEIGHTS = [15, 4, 1]
P_C = 0.003
P_C_D = 0
TR = [0.005, 0.01, 0.1]
TR_P = [3, 2, 1]


class robust_control():
    """Per-episode parameter perturbation of the reference (:84-109).  Dead code at HEAD
    (`quad.robust_control = False`, :183) — only its RNG draw pattern is reproduced."""

    def __init__(self):
        self.D_KF, self.D_KM, self.D_M, self.D_IR = 0.1, 0.1, 0.3, 0.1
        self.D_J = np.ones(3) * 0.1
        self.reset()

    def reset(self):
        self.episode_kf = np.random.random(4) * self.D_KF
        self.episode_m = np.random.normal(0, self.D_M, 1)
        self.episode_ir = np.random.random(4) * self.D_IR
        self.episode_J = np.eye(3) * np.random.normal(np.zeros(3), self.D_J, [3])


class quad():
    def __init__(self, t_step, n, training=True, euler=0, direct_control=1, T=1, clipped=True, *,
                 precision="f64", integrator=None, substeps=1, robust_rng_draws=True, device=None, verbose=True):
        self.clipped = clipped
        self.ppo_training = bool(training)
        self.mass = M
        self.gravity = G
        self.i = 0
        self.T = T
        self.bb_cond = np.array([BB_VEL, BB_VEL, BB_VEL, BB_ANG, BB_ANG, 3 / 4 * np.pi,
                                 BB_VEL * 2, BB_VEL * 2, BB_VEL * 2])
        self.state_size = 13
        self.action_size = 4
        self.done = True
        self.n = n + self.T
        self._n_arg = n
        self.t_step = t_step
        if direct_control:
            self.zero_control = np.ones(4) * (2 / T2WR - 1)
        else:
            self.zero_control = np.array([M * G, 0, 0, 0])
        self.direct_control_flag = direct_control
        self.ang_vel = np.zeros(3)
        self.prev_ang = np.zeros(3)
        self.ang = np.zeros(3)
        self.J_mat = J
        self.abs_sum = 0
        self.solved = 0
        self.reward = 0
        self.robust_control = False
        self._robust_rng_draws = robust_rng_draws
        if robust_rng_draws:
            self.robust_parameters = robust_control()          # consumes RNG draws like the reference (:182)
        self._precision = precision
        self._integrator = integrator
        self._substeps = substeps
        self._device = device
        self._sim = None                                       # created lazily (keeps the object picklable)
        self._fields = None
        self.action_hist = []
        self.state = np.zeros(13)
        self.previous_state = np.zeros(13)
        self.target_state = 9 * (TR[0] ** 2)
        if verbose:
            ev_cd = 'Training' if self.ppo_training else 'Eval'
            print('Environment Condition: ' + ev_cd)

    # -- pickling (environment/controller/ppo.py ships workers to a multiprocessing pool) --------------
    def __getstate__(self):
        d = dict(self.__dict__)
        if self._sim is not None:
            d["_ws_host"] = self._sim._ws.cpu()
        d["_sim"] = None
        d["_fields"] = None
        return d

    def __setstate__(self, d):
        ws = d.pop("_ws_host", None)
        self.__dict__.update(d)
        if ws is not None:
            self._ensure_sim()
            self._sim._ws.copy_(ws.to(self._sim._ws.device))

    # -- device plumbing -----------------------------------------------------------------------------------
    def _ensure_sim(self):
        if self._sim is None:
            from .batched import BatchedQuad
            self._sim = BatchedQuad(1, self.t_step, self._n_arg, training=self.ppo_training,
                                    direct_control=self.direct_control_flag, T=self.T, clipped=self.clipped,
                                    precision=self._precision, integrator=self._integrator, substeps=self._substeps,
                                    aux=True, device=self._device)
            # host map of the workspace: field -> (offset, channels, ld, dtype)
            import ctypes as C
            fm = {}
            real = np.float64 if self._precision == "f64" else np.float32
            for name, f in [("obs", L.QS_FIELD_OBS), ("ang", L.QS_FIELD_ANG), ("ang_vel", L.QS_FIELD_ANG_VEL),
                            ("step_effort", L.QS_FIELD_STEP_EFFORT), ("w", L.QS_FIELD_W), ("reward", L.QS_FIELD_REWARD),
                            ("done", L.QS_FIELD_DONE), ("solved", L.QS_FIELD_SOLVED), ("i", L.QS_FIELD_I),
                            ("abs_sum", L.QS_FIELD_ABS_SUM), ("accel", L.QS_FIELD_ACCEL),
                            ("acc_read", L.QS_FIELD_ACC_READ), ("mat_rot", L.QS_FIELD_MAT_ROT),
                            ("clipped_action", L.QS_FIELD_CLIPPED_ACTION), ("fm", L.QS_FIELD_FM)]:
                d = L.qs_field_desc()
                L.check(self._sim.lib.qs_field_info(self._sim._h, f, C.byref(d)))
                dt = {1: np.uint8, 4: np.int32 if name == "i" else np.float32, 8: np.float64}[d.elem_bytes]
                if d.elem_bytes == 4 and name != "i":
                    dt = real
                fm[name] = (d.ws_offset, d.channels, d.ld, dt)
            # obs17: 17 rows starting at the OBS offset
            o = fm["obs"]
            fm["obs17"] = (o[0], 17, o[2], o[3])
            self._fields = fm
        return self._sim

    def _pull(self):
        """One D2H copy of the (few-KB) workspace, then slice the reference's attributes out of it."""
        ws = self._sim._ws.cpu().numpy()

        def get(name):
            off, c, ld, dt = self._fields[name]
            item = np.dtype(dt).itemsize
            return np.frombuffer(ws, dtype=dt, count=c * ld, offset=off).reshape(c, ld)[:, 0].astype(
                np.float64 if dt in (np.float32, np.float64) else dt)

        o17 = get("obs17")
        self.state = np.concatenate((o17[0:10], o17[14:17]))
        self.V_q = o17[10:14].copy()
        self.quat_state = np.array([np.concatenate((self.state[0:10], self.V_q))])
        self.previous_state = self.state
        self.ang = get("ang")
        self.prev_ang = self.ang
        self.ang_vel = get("ang_vel")
        self.step_effort = get("step_effort")
        self.w = get("w").reshape(4, 1)
        self.reward = float(get("reward")[0])
        self.done = bool(get("done")[0])
        self.solved = int(get("solved")[0])
        self.i = int(get("i")[0])
        self._abs_sum = float(get("abs_sum")[0])
        self.accel = get("accel").reshape(3, 1)
        self.accelerometer_read = get("acc_read")
        self.mat_rot = get("mat_rot").reshape(3, 3)
        self.clipped_action = get("clipped_action")
        fm = get("fm")
        self.f_in = np.array([[0, 0, fm[0]]]).T
        self.current_state = float(np.sum(np.square(np.concatenate((self.state[1:6:2], self.ang, self.state[-3:])))))

    @property
    def abs_sum(self):
        return self._abs_sum

    @abs_sum.setter
    def abs_sum(self, v):
        self._abs_sum = v
        if getattr(self, "_sim", None) is not None:
            import torch
            self._sim._field(L.QS_FIELD_ABS_SUM).fill_(float(v))

    # -- reference API ---------------------------------------------------------------------------------------
    def seed(self, seed):
        """quad.seed (:189-193): seeds the global NumPy RNG (and re-keys the device Philox streams)."""
        np.random.seed(seed)
        if self._sim is not None:
            self._sim.seed(seed)

    def reset(self, det_state=None):
        """quad.reset (:408-454)."""
        sim = self._ensure_sim()
        self.action_hist = []
        if self._robust_rng_draws:
            self.robust_parameters.reset()                                        # :426 (12 draws, unused)
        if det_state is not None:
            init = np.asarray(det_state, dtype=np.float64).reshape(13)
        else:
            from .quaternion_euler_utility import euler_quat
            init = np.zeros(13)
            ang = np.random.rand(3) - 0.5                                         # :440
            Q_in = euler_quat(ang)                                                # :441 (device)
            init[0:5:2] = np.clip((np.random.normal([0, 0, 0], 2)), -BB_POS / 2, BB_POS / 2)
            init[1:6:2] = np.clip((np.random.normal([0, 0, 0], 2)), -BB_VEL / 2, BB_VEL / 2)
            init[6:10] = Q_in.T
            init[10:13] = np.clip((np.random.normal([0, 0, 0], 2)), -BB_VEL * 1.5, BB_POS * 1.5)
        obs_h, act_h = sim.reset(init.reshape(1, 13))
        state = obs_h[:, 0, :].cpu().numpy().astype(np.float64)
        action = act_h[:, 0, :].cpu().numpy().astype(np.float64)
        for _ in range(self.T):                                                   # :447-450
            self.action = self.zero_control
            self.action_hist.append(self.action)
            self.action_hist.append(self.zero_control)
        self._pull()
        return state, action

    def step(self, action):
        """quad.step (:458-498)."""
        sim = self._ensure_sim()
        a = np.asarray(action, dtype=np.float64).reshape(-1)[:4]
        self.action = np.clip(a, -1, 1) if self.direct_control_flag else a
        sim.step(a.reshape(1, 4))
        self._pull()
        self.action_hist.append(self.clipped_action)
        return self.quat_state, self.reward, self.done


class sensor():
    """Single-env `sensor` (reference :579-724) is not part of the accelerated path; the batched in-kernel
    sensor model is `BatchedQuad(sensor_noise=True)` (sensed_obs / sensor_state fields)."""

    def __init__(self, env, *a, **k):
        raise NotImplementedError(
            "use BatchedQuad(sensor_noise=True): the single-env host sensor object is outside the accelerated path")


class plotter():
    """No-op stand-in for the reference's matplotlib/pgf plotter (:727-836) so controller scripts run."""

    def __init__(self, env, velocity_plot=False, depth_plot=False):
        self.env = env
        self.states, self.times = [], []
        self.axs = _Noop()

    def add(self, target=None):
        pass

    def clear(self):
        pass

    def plot(self, *a, **k):
        pass


class _Noop:
    def __getitem__(self, i):
        return self

    def __getattr__(self, name):
        return lambda *a, **k: None
