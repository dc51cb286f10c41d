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
    _instances_created = 0          # diagnostic: lets a harness check that a script really resolved to this class

    def __init__(self, t_step, n, training=True, euler=0, direct_control=1, T=1, clipped=True, *,
                 precision="f64", integrator=None, substeps=1, robust_rng_draws=True, device=None, verbose=True):
        quad._instances_created += 1
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
        for k in ("_ws_pin", "_a_host", "_a_dev"):              # staging buffers are rebuilt on demand
            d.pop(k, None)
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
        sim = self._sim
        import torch
        if getattr(self, "_ws_pin", None) is None:               # pinned mirror: one asynchronous copy + one stream synchronisation
            self._ws_pin = torch.empty(sim._ws.shape, dtype=sim._ws.dtype, pin_memory=True)
        self._ws_pin.copy_(sim._ws, non_blocking=True)
        torch.cuda.current_stream(sim.device).synchronize()
        ws = self._ws_pin.numpy()                               # every attribute below is a copy (astype), never a view of the mirror

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
        a = np.array(action, dtype=np.float64).reshape(-1)[:4]
        self.action = np.clip(a, -1, 1) if self.direct_control_flag else a
        import torch
        if getattr(self, "_a_host", None) is None:               # pinned staging of the 4-float action and its device twin
            self._a_host = torch.empty(4, 1, dtype=sim.dtype, pin_memory=True)
            self._a_dev = torch.empty(4, 1, dtype=sim.dtype, device=sim.device)
        self._a_host[:, 0] = torch.from_numpy(a)
        self._a_dev.copy_(self._a_host, non_blocking=True)
        sim.step_soa(self._a_dev)
        self._pull()
        self.action_hist.append(self.clipped_action)
        return self.quat_state, self.reward, self.done


class sensor():
    """Drop-in for the reference's single-env `sensor` (:579-724): same constructor, methods (`bias_reset`, `reset`, `accel`, `gyro`,
    `gps`, `triad`, `accel_int`, `gyro_int`), attributes and — given the same NumPy seed — the same readings.

    Every random draw comes from the global NumPy stream in the reference's order (`np.random.normal(loc, scale, n)` is
    `loc + scale * z` over NumPy's own gaussians, so drawing the standard z here consumes the stream identically); the
    arithmetic of each method runs on the device through the C-ABI entry point `qs_sensor_call` — the device functions the
    in-kernel model of `BatchedQuad(sensor_noise=True)` is made of.  Not reproduced: the aliasing of `sensor.reset`
    (:636-638: it keeps VIEWS of `quad.state`, and the first `gyro_int` of an episode writes through them into the true
    quaternion); this class copies."""

    def __init__(self, env,
                 accel_std=0.1, accel_bias_drift=0.0005,
                 gyro_std=0.035, gyro_bias_drift=0.00015,
                 magnet_std=15, magnet_bias_drift=0.075,
                 gps_std_p=1.71, gps_std_v=0.5, *, precision="f64", device=None):
        self.std = [accel_std, gyro_std, magnet_std, gps_std_p, gps_std_v]
        self.b_d = [accel_bias_drift, gyro_bias_drift, magnet_bias_drift]
        self.quad = env
        self.error = True
        self.bias_reset()
        self.R = np.eye(3)
        self.a_b_grav = self.a_b_accel = self.m_b = self.g_b = 0
        self.acceleration_t0 = np.zeros(3)
        self.position_t0 = np.zeros(3)
        self.velocity_t0 = np.zeros(3)
        self.quaternion_t0 = np.array([1.0, 0.0, 0.0, 0.0])
        self._precision = precision
        self._device = device
        self._buf = None

    def bias_reset(self):                                          # :600-608 (three uniform draws)
        self.a_std = self.std[0] * self.error
        self.a_b_d = (np.random.random() - 0.5) * 2 * self.b_d[0] * self.error
        self.g_std = self.std[1] * self.error
        self.g_b_d = (np.random.random() - 0.5) * 2 * self.b_d[1] * self.error
        self.m_std = self.std[2] * self.error
        self.m_b_d = (np.random.random() - 0.5) * 2 * self.b_d[2] * self.error
        self.gps_std_p = self.std[3] * self.error
        self.gps_std_v = self.std[4] * self.error

    def reset(self):                                               # :630-640
        self.a_b_grav = 0
        self.a_b_accel = 0
        self.m_b = 0
        self.g_b = 0
        self.acceleration_t0 = np.zeros(3)
        st = np.asarray(self.quad.state, dtype=np.float64).reshape(-1)
        self.position_t0 = st[0:5:2].copy()
        self.velocity_t0 = st[1:6:2].copy()
        self.quaternion_t0 = st[6:10].copy()
        self.bias_reset()

    # -- device plumbing: one H2D of the packed inputs, one launch of the method, one D2H of state + outputs ------------------
    def _call(self, method):
        import ctypes as C
        import torch
        lib = L.load_library()
        nz, no = L.SENSOR_Z_ROWS[method], L.SENSOR_OUT_ROWS[method]
        z = np.random.normal(0.0, 1.0, nz)
        q = self.quad
        host = np.zeros(20 + 13 + 3 + 9 + 1 + 27 + 18, dtype=np.float64)
        host[0:4] = [self.a_b_accel, self.g_b, self.a_b_d, self.g_b_d]
        host[4:7] = np.asarray(self.velocity_t0, dtype=np.float64).reshape(-1)
        host[7:10] = np.asarray(self.position_t0, dtype=np.float64).reshape(-1)
        host[10:14] = np.asarray(self.quaternion_t0, dtype=np.float64).reshape(-1)
        host[14:17] = np.asarray(self.R, dtype=np.float64)[:, 2]
        host[17:20] = np.asarray(self.acceleration_t0, dtype=np.float64).reshape(-1)
        host[20:33] = np.asarray(q.state, dtype=np.float64).reshape(-1)
        host[33:36] = np.asarray(q.accelerometer_read, dtype=np.float64).reshape(-1)
        host[36:45] = np.asarray(q.mat_rot, dtype=np.float64).reshape(-1)
        host[45] = float(np.asarray(q.f_in, dtype=np.float64).reshape(-1)[2]) / M
        host[46:46 + nz] = z
        f64 = self._precision == "f64"
        dev = self._device or (q._sim.device if getattr(q, "_sim", None) is not None else "cuda")
        buf = torch.as_tensor(host if f64 else host.astype(np.float32)).to(dev)
        rs = buf.element_size()
        base = buf.data_ptr()
        par = L.default_config().params
        par.accel_std, par.gyro_std, par.magnet_std = float(self.a_std), float(self.g_std), float(self.m_std)
        par.gps_std_p, par.gps_std_v, par.gps_blend = float(self.gps_std_p), float(self.gps_std_v), 0.0
        with torch.cuda.device(buf.device):
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            L.check(lib.qs_sensor_call(L.QS_F64 if f64 else L.QS_F32, C.byref(par), float(q.t_step), 1, method,
                                       C.c_void_p(base), C.c_void_p(base + 20 * rs), C.c_void_p(base + 33 * rs),
                                       C.c_void_p(base + 36 * rs), C.c_void_p(base + 45 * rs), C.c_void_p(base + 46 * rs),
                                       C.c_void_p(base + 73 * rs), st))
        res = buf.cpu().numpy().astype(np.float64)
        self.a_b_accel, self.g_b = float(res[0]), float(res[1])
        self.velocity_t0, self.position_t0, self.quaternion_t0 = res[4:7].copy(), res[7:10].copy(), res[10:14].copy()
        self.acceleration_t0 = res[17:20].copy()
        return res[73:73 + no], res[14:17]

    # -- the reference's methods --------------------------------------------------------------------------------------------
    def accel(self):                                               # :611-620
        return self._call(L.QS_SENSOR_ACCEL)[0].copy()

    def gyro(self):                                                # :622-628
        return self._call(L.QS_SENSOR_GYRO)[0].copy()

    def gps(self):                                                 # :642-647
        o = self._call(L.QS_SENSOR_GPS)[0]
        return o[0:3].copy(), o[3:6].copy()

    def triad(self):                                               # :649-697
        o = self._call(L.QS_SENSOR_TRIAD)[0]
        self.R = o[4:13].reshape(3, 3).copy()
        return o[0:4].copy(), self.R

    def accel_int(self):                                           # :700-715
        o = self._call(L.QS_SENSOR_ACCEL_INT)[0]
        self.R = o[9:18].reshape(3, 3).copy()
        return o[0:3].copy(), o[3:6].copy(), o[6:9].copy()

    def gyro_int(self):                                            # :717-724
        return self._call(L.QS_SENSOR_GYRO_INT)[0].copy()


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
