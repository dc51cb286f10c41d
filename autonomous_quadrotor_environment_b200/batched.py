"""BatchedQuad — N independent reference `quad` environments advanced in lock-step on one B200.

Host-side mirror of the reference's environment API (`quad.__init__/seed/reset/step`,
environment/quadrotor_env.py:111-498) over the C ABI of libquadsim.so.  PyTorch is used for device
memory (the handle's workspace is a torch tensor, so every field below is a zero-copy torch view),
streams and torch.distributed; all arithmetic happens in the hand-written CUDA kernels.

Shapes follow the reference with a leading env axis: state (N,13), observation (N,14), action (N,4).
The underlying storage is structure-of-arrays ([C][N], env index fastest), so these tensors are
*transposed views* — pass actions as ``a.t()`` of a contiguous (4,N) tensor (or use ``step_soa``) to
avoid a transpose copy.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L

_TORCH_DTYPE = {4: torch.float32, 8: torch.float64}


class BatchedQuad:
    def __init__(self, n_envs: int, t_step: float = 0.01, n: int = 1000, training: bool = True, euler: int = 0,
                 direct_control: int = 1, T: int = 1, clipped: bool = True, *, precision: str = "f32",
                 integrator: Optional[str] = None, substeps: int = 1, auto_reset: bool = False,
                 async_reset: bool = False, sensor_noise: bool = False, aux: bool = False, seed: int = 0, device=None,
                 env_id_offset: int = 0, params: Optional[dict] = None, robust_control: bool = False):
        if not torch.cuda.is_available():
            raise RuntimeError("BatchedQuad needs a CUDA device: the simulator has no CPU fallback")
        self.lib = L.load_library()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        if precision not in ("f32", "f64"):
            raise ValueError("precision must be 'f32' or 'f64'")
        if integrator is None:
            integrator = "rk4" if precision == "f32" else "rk45"
        if integrator not in ("rk4", "rk45"):
            raise ValueError("integrator must be 'rk4' or 'rk45'")
        cfg = L.default_config()
        cfg.n_envs = int(n_envs)
        cfg.env_id_offset = int(env_id_offset)
        cfg.t_step = float(t_step)
        cfg.n_max = int(n)
        cfg.T = int(T)
        cfg.substeps = int(substeps)
        cfg.precision = L.QS_F64 if precision == "f64" else L.QS_F32
        cfg.integrator = L.QS_RK45 if integrator == "rk45" else L.QS_RK4
        flags = 0
        flags |= L.QS_FLAG_DIRECT_CONTROL if direct_control else 0
        flags |= L.QS_FLAG_CLIPPED if clipped else 0
        flags |= L.QS_FLAG_TRAINING if training else 0
        flags |= L.QS_FLAG_AUTO_RESET if auto_reset else 0
        flags |= L.QS_FLAG_ASYNC_RESET if async_reset else 0
        flags |= L.QS_FLAG_SENSOR_NOISE if sensor_noise else 0
        flags |= L.QS_FLAG_AUX if aux else 0
        flags |= L.QS_FLAG_ROBUST if robust_control else 0     # quad.robust_control = True (quadrotor_env.py:183)
        cfg.flags = flags
        cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        cfg.device = self.device.index
        if params:
            for k, v in params.items():
                cur = getattr(cfg.params, k)
                if hasattr(cur, "__len__"):
                    for i, x in enumerate(v):
                        cur[i] = float(x)
                else:
                    setattr(cfg.params, k, type(cur)(v))
        nbytes = self.lib.qs_workspace_bytes(C.byref(cfg))
        if nbytes < 0:
            L.check(int(nbytes))
        # the workspace is owned by torch's allocator; the handle carves its SoA rows out of it
        self._ws = torch.zeros(int(nbytes), dtype=torch.uint8, device=self.device)
        cfg.workspace = self._ws.data_ptr()
        self._cfg = cfg
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            L.check(self.lib.qs_create(C.byref(h), C.byref(cfg)))
        self._h = h
        self.N = int(n_envs)
        self.T = int(T)
        self.n = int(n) + int(T)                       # quadrotor_env.py:157
        self.t_step = float(t_step)
        self.state_size, self.action_size = 13, 4
        self.precision = precision
        self.integrator = integrator
        self.dtype = torch.float64 if precision == "f64" else torch.float32
        self.direct_control_flag = int(bool(direct_control))
        self.flags = flags
        self._views = {}
        self._act = torch.zeros(4, self.N, dtype=self.dtype, device=self.device)
        self._obs_hist = None
        self._act_hist = None
        zc = [2.0 / cfg.params.t2wr - 1.0] * 4 if direct_control else [cfg.params.mass * cfg.params.gravity, 0, 0, 0]
        self.zero_control = torch.tensor(zc, dtype=self.dtype, device=self.device)

    # ------------------------------------------------------------------ plumbing
    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self.lib.qs_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _field(self, f: int) -> torch.Tensor:
        """Zero-copy (C,N) torch view of a handle-owned field."""
        v = self._views.get(f)
        if v is None:
            d = L.qs_field_desc()
            L.check(self.lib.qs_field_info(self._h, f, C.byref(d)))
            if d.ws_offset < 0:
                raise L.QuadSimError(L.QS_ESTATE, "field %d has no zero-copy view" % f)
            nbytes = d.channels * d.ld * d.elem_bytes
            raw = self._ws[d.ws_offset:d.ws_offset + nbytes]
            if d.elem_bytes == 1:
                t = raw
            elif f in (L.QS_FIELD_I, L.QS_FIELD_GUST_COUNT):
                t = raw.view(torch.int32)
            elif f in (L.QS_FIELD_EPISODE,):
                t = raw.view(torch.int32)          # torch has limited uint32 support; values < 2^31 in practice
            else:
                t = raw.view(_TORCH_DTYPE[d.elem_bytes])
            v = t.view(d.channels, d.ld)[:, :self.N]
            self._views[f] = v
        return v

    def _as_soa(self, x, channels: int, layout=None) -> torch.Tensor:
        """Return a contiguous (C,N) tensor holding x, given as (N,C) — the reference's layout with a leading env axis — or (C,N).
        layout: "nc" / "cn" / None = by shape; with N == C the shape does not tell and the layout must be named."""
        x = torch.as_tensor(x, dtype=self.dtype, device=self.device)
        if x.dim() == 1 and self.N == 1:
            x = x.view(1, channels)
        if layout not in (None, "nc", "cn"):
            raise ValueError("layout must be 'nc', 'cn' or None")
        if layout is None and self.N == channels and self.N > 1 and x.shape == (self.N, channels):
            raise ValueError("a (%d,%d) tensor is ambiguous for %d envs x %d channels: pass layout='nc' (env-major, the reference's "
                             "layout) or layout='cn' (channel-major)" % (self.N, channels, self.N, channels))
        if layout == "cn" and x.shape == (channels, self.N):
            return x.contiguous()
        if layout == "cn":
            raise ValueError("expected shape (%d,%d), got %s" % (channels, self.N, tuple(x.shape)))
        if layout == "nc" and x.shape != (self.N, channels):
            raise ValueError("expected shape (%d,%d), got %s" % (self.N, channels, tuple(x.shape)))
        if x.shape == (self.N, channels):
            xt = x.t()
            return xt if xt.is_contiguous() else xt.contiguous()
        if x.shape == (channels, self.N):
            return x.contiguous()
        raise ValueError("expected shape (%d,%d) or (%d,%d), got %s" % (self.N, channels, channels, self.N, tuple(x.shape)))

    def set_step_loader(self, loader: int):
        """Pick the qs_step implementation (0 plain loads, 1 CTA-wide TMA ring, 2 per-warp cp.async pipeline, 3 per-warp
        pipeline with two envs per lane on the packed FP32 pipe); A/B testing."""
        L.check(self.lib.qs_set_step_loader(self._h, int(loader)))
        return self

    @property
    def step_loader(self) -> int:
        """The qs_step implementation this handle launches (see set_step_loader)."""
        r = int(self.lib.qs_get_step_loader(self._h))
        if r < 0:
            L.check(r)
        return r

    # ------------------------------------------------------------------ reference API
    def seed(self, seed: int):
        """quad.seed (quadrotor_env.py:189-193): re-keys the Philox streams."""
        L.check(self.lib.qs_seed(self._h, int(seed) & 0xFFFFFFFFFFFFFFFF))
        self._cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF       # get_checkpoint() saves the key in force

    def reset(self, det_state=None, mask=None, layout=None):
        """quad.reset (quadrotor_env.py:408-454) for the masked envs (all if mask is None).

        det_state: (N,13) initial states, or None for the random branch (sampled on the device).
        Returns (obs_hist (T,N,14), act_hist (T,N,4)) — rows of envs outside the mask are stale.
        """
        if self._obs_hist is None:
            self._obs_hist = torch.zeros(self.T, 14, self.N, dtype=self.dtype, device=self.device)
            self._act_hist = torch.zeros(self.T, 4, self.N, dtype=self.dtype, device=self.device)
        det = None if det_state is None else self._as_soa(det_state, 13, layout)
        m = None
        if mask is not None:
            m = torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
            if m.shape != (self.N,):
                raise ValueError("mask must have shape (N,)")
        with torch.cuda.device(self.device):
            L.check(self.lib.qs_reset(self._h, None if det is None else C.c_void_p(det.data_ptr()),
                                      None if m is None else C.c_void_p(m.data_ptr()),
                                      C.c_void_p(self._obs_hist.data_ptr()), C.c_void_p(self._act_hist.data_ptr()),
                                      self._stream()))
        return self._obs_hist.transpose(1, 2), self._act_hist.transpose(1, 2)

    def step_soa(self, action_soa: torch.Tensor):
        """quad.step with a contiguous (4,N) action tensor; returns zero-copy views (obs (N,14), reward (N,), done (N,))."""
        if action_soa.shape != (4, self.N) or action_soa.dtype != self.dtype or not action_soa.is_contiguous():
            raise ValueError("step_soa needs a contiguous (4,N) %s tensor" % self.dtype)
        L.check(self.lib.qs_step(self._h, C.c_void_p(action_soa.data_ptr()), None, None, None, None, self._stream()))
        return self.obs, self.reward, self.done

    def step(self, action, layout=None):
        """quad.step (quadrotor_env.py:458-498): action (N,4) -> (obs (N,14), reward (N,), done (N,) uint8)."""
        return self.step_soa(self._as_soa(action, 4, layout))

    def rollout(self, horizon: int, actions=None, record_obs=False, record_actions=False, record_reward=False,
                record_done=False, record_sensed=False):
        """K fused env steps in ONE launch (state stays in registers).  actions: (K,4,N) tensor, or None to draw
        a ~ U(-1,1)^4 in-kernel with Philox.  Returns a dict of the recorded (K,C,N) buffers."""
        a = L.qs_rollout_args()
        a.horizon = int(horizon)
        out = {}
        if actions is None:
            a.action_source = L.QS_ACT_PHILOX_UNIFORM
        else:
            if actions.shape != (horizon, 4, self.N) or not actions.is_contiguous() or actions.dtype != self.dtype:
                raise ValueError("actions must be a contiguous (K,4,N) %s tensor" % self.dtype)
            a.action_source = L.QS_ACT_BUFFER
            a.actions = actions.data_ptr()
        if record_obs:
            out["obs"] = torch.empty(horizon, 14, self.N, dtype=self.dtype, device=self.device)
            a.obs_out = out["obs"].data_ptr()
        if record_actions:
            out["actions"] = torch.empty(horizon, 4, self.N, dtype=self.dtype, device=self.device)
            a.action_out = out["actions"].data_ptr()
        if record_reward:
            out["reward"] = torch.empty(horizon, self.N, dtype=self.dtype, device=self.device)
            a.reward_out = out["reward"].data_ptr()
        if record_done:
            out["done"] = torch.empty(horizon, self.N, dtype=torch.uint8, device=self.device)
            a.done_out = out["done"].data_ptr()
        if record_sensed:                      # sensor_noise=True handles: the sensor-based observation of every step
            out["sensed_obs"] = torch.empty(horizon, 14, self.N, dtype=self.dtype, device=self.device)
            a.sensed_obs_out = out["sensed_obs"].data_ptr()
        L.check(self.lib.qs_rollout(self._h, C.byref(a), self._stream()))
        return out

    # ------------------------------------------------------------------ fused actor rollout (config 5)
    def load_actor(self, source, action_std: float = 0.1, critic=None):
        """Prepare the reference's actor (environment/controller/model.py:27-34) for policy_rollout.
        source: path to a solved/*.pth state dict, a state dict, or a dict/npz with keys actor_{0,2,4}_{weight,bias}.
        critic: None, True (take critic.{0,2,4}.* from `source`) or a dict with those keys: the critic head (model.py:36-43) is then
        evaluated by the same kernel on every network input and policy_rollout(record_values=True) returns the state values."""
        if isinstance(source, (str, bytes)):
            source = torch.load(source, map_location="cpu")
        if critic is True:
            critic = source

        def get(i, kind):
            for key in ("actor.%d.%s" % (i, kind), "actor_%d_%s" % (i, kind), "%d.%s" % (i, kind)):
                if key in source:
                    return torch.as_tensor(source[key]).to(device=self.device, dtype=torch.float32).contiguous()
            raise KeyError("actor layer %d %s not found" % (i, kind))

        w = {"w1": get(0, "weight"), "b1": get(0, "bias"), "w2": get(2, "weight"), "b2": get(2, "bias"),
             "w3": get(4, "weight"), "b3": get(4, "bias")}
        if w["w1"].shape != (128, 75) or w["w2"].shape != (128, 128) or w["w3"].shape != (4, 128):
            raise ValueError("only the 75-128-128-4 actor is supported by the fused kernel")
        if critic is not None:                                  # model.py:36-43: Linear(75,128)-Tanh-Linear(128,128)-Tanh-Linear(128,1)
            def getc(i, kind):
                for key in ("critic.%d.%s" % (i, kind), "critic_%d_%s" % (i, kind), "%d.%s" % (i, kind)):
                    if key in critic:
                        return torch.as_tensor(critic[key]).to(device=self.device, dtype=torch.float32).contiguous()
                raise KeyError("critic layer %d %s not found" % (i, kind))
            w.update({"cw1": getc(0, "weight"), "cb1": getc(0, "bias"), "cw2": getc(2, "weight"), "cb2": getc(2, "bias"),
                      "cw3": getc(4, "weight"), "cb3": getc(4, "bias")})
            if w["cw1"].shape != (128, 75) or w["cw2"].shape != (128, 128) or w["cw3"].numel() != 128:
                raise ValueError("only the 75-128-128-1 critic is supported by the fused kernel")
        a = L.qs_actor()
        for k, t in w.items():
            setattr(a, k, t.data_ptr())
        a.hidden, a.in_dim, a.action_std = 128, 75, float(action_std)
        self._actor = (a, w)                                    # keep the tensors alive
        return self

    @property
    def history(self) -> torch.Tensor:
        """(N,75) view of dl_in_gen.deep_learning_input per env (oldest entry first); persists across policy_rollout calls."""
        if getattr(self, "_hist", None) is None:
            self._hist = torch.zeros(75, self.N, dtype=torch.float32, device=self.device)
        return self._hist.t()

    def policy_rollout(self, horizon: int, record_obs=False, record_actions=True, record_logprob=True,
                       record_reward=True, record_done=True, record_values=False, record_sensed=False):
        """K fused steps of  history -> actor MLP (tcgen05) -> Normal sample -> quad.step -> history push  in ONE launch
        (the loop of environment/controller/ppo.py:238-257).  On a sensor_noise handle the sensor model runs after every step and the
        history takes the SENSED observation (record_sensed -> out["sensed_obs"], (K,14,N)).  Returns the recorded (K,C,N) buffers."""
        if getattr(self, "_actor", None) is None:
            raise RuntimeError("call load_actor() first")
        _ = self.history
        a = L.qs_policy_rollout_args()
        a.horizon = int(horizon)
        a.hist = self._hist.data_ptr()
        out = {}
        f32 = torch.float32
        if record_obs:
            out["obs"] = torch.empty(horizon, 14, self.N, dtype=f32, device=self.device); a.obs_out = out["obs"].data_ptr()
        if record_actions:
            out["actions"] = torch.empty(horizon, 4, self.N, dtype=f32, device=self.device); a.action_out = out["actions"].data_ptr()
        if record_logprob:
            out["logprob"] = torch.empty(horizon, 4, self.N, dtype=f32, device=self.device); a.logprob_out = out["logprob"].data_ptr()
        if record_reward:
            out["reward"] = torch.empty(horizon, self.N, dtype=f32, device=self.device); a.reward_out = out["reward"].data_ptr()
        if record_done:
            out["done"] = torch.empty(horizon, self.N, dtype=torch.uint8, device=self.device); a.done_out = out["done"].data_ptr()
        if record_values:                      # (K+1,N): V of the network input of every step + the bootstrap row (load_actor(critic=...))
            out["value"] = torch.empty(horizon + 1, self.N, dtype=f32, device=self.device); a.value_out = out["value"].data_ptr()
        if record_sensed:
            out["sensed_obs"] = torch.empty(horizon, 14, self.N, dtype=f32, device=self.device); a.sensed_obs_out = out["sensed_obs"].data_ptr()
        L.check(self.lib.qs_policy_rollout(self._h, C.byref(self._actor[0]), C.byref(a), self._stream()))
        return out

    def controller_state(self) -> torch.Tensor:
        """Fresh controller memory (QS_CTRL_STATE_DIM, N): zeros, pending PID action = hover [M*G,0,0,0]
        (pid_vel_control.py:143-144).  Pass it to control_rollout(ctrl_state=...) to continue a controller across launches."""
        cs = torch.zeros(L.QS_CTRL_STATE_DIM, self.N, dtype=self.dtype, device=self.device)
        cs[18] = float(self._cfg.params.mass * self._cfg.params.gravity)
        return cs

    def control_rollout(self, controller, horizon: int, ctrl_state=None, record_obs=False, record_actions=False,
                        record_reward=False, record_done=False, record_aux=False, target_traj=None):
        """K fused env steps driven by an in-kernel classical control law (controllers.lqr_controller / pid_controller:
        environment/controller/lqr_quad.py:129-157, pid_vel_control.py:29-127) in ONE launch.  Needs direct_control=0.
        record_aux -> (K,10,N): ang(3), ang_vel(3), step_effort(4), i.e. with obs[:, (1,3,5)] the 13 columns of the reference's
        classical_controller_results logs.  target_traj: (K,3) velocity set-points of the PID law, one per step (e.g.
        mission(...).velocity_setpoints(K): mission_control/mission_control.py); None = controller.target_vel.
        Returns the recorded buffers."""
        a = L.qs_control_rollout_args()
        a.horizon = int(horizon)
        if target_traj is not None:
            target_traj = torch.as_tensor(target_traj, dtype=self.dtype, device=self.device).contiguous()
            if target_traj.shape != (horizon, 3):
                raise ValueError("target_traj must have shape (horizon, 3)")
            a.target_traj = target_traj.data_ptr()
        if ctrl_state is not None:
            if ctrl_state.shape != (L.QS_CTRL_STATE_DIM, self.N) or ctrl_state.dtype != self.dtype or not ctrl_state.is_contiguous():
                raise ValueError("ctrl_state must be a contiguous (%d,N) %s tensor" % (L.QS_CTRL_STATE_DIM, self.dtype))
            a.ctrl_state = ctrl_state.data_ptr()
        out = {}
        mk = lambda c: torch.empty(horizon, c, self.N, dtype=self.dtype, device=self.device)
        if record_obs:
            out["obs"] = mk(14); a.obs_out = out["obs"].data_ptr()
        if record_actions:
            out["actions"] = mk(4); a.action_out = out["actions"].data_ptr()
        if record_aux:
            out["aux"] = mk(10); a.aux_out = out["aux"].data_ptr()
        if record_reward:
            out["reward"] = torch.empty(horizon, self.N, dtype=self.dtype, device=self.device); a.reward_out = out["reward"].data_ptr()
        if record_done:
            out["done"] = torch.empty(horizon, self.N, dtype=torch.uint8, device=self.device); a.done_out = out["done"].data_ptr()
        L.check(self.lib.qs_control_rollout(self._h, C.byref(controller), C.byref(a), self._stream()))
        return out

    # ------------------------------------------------------------------ attributes of the reference `quad`
    @property
    def obs(self):            # quat_state :486
        return self._field(L.QS_FIELD_OBS).t()

    @property
    def quat_state(self):
        return self.obs

    @property
    def reward(self):
        return self._field(L.QS_FIELD_REWARD)[0]

    @property
    def done_flags(self):
        """Raw done byte: bit0 = done, bit1 = asynchronous warm-up step (async_reset=True only)."""
        return self._field(L.QS_FIELD_DONE)[0]

    @property
    def done(self):
        d = self._field(L.QS_FIELD_DONE)[0]
        return (d & 1) if (self.flags & L.QS_FLAG_ASYNC_RESET) else d

    @property
    def warmup(self):
        """1 where the last step was one of the T hover steps of an asynchronous reset (transition to ignore)."""
        return (self._field(L.QS_FIELD_DONE)[0] >> 1) & 1

    @property
    def solved(self):
        return self._field(L.QS_FIELD_SOLVED)[0]

    @property
    def state(self):          # (N,13) copy, gathered from the obs17 rows
        out = torch.empty(13, self.N, dtype=self.dtype, device=self.device)
        L.check(self.lib.qs_get(self._h, L.QS_FIELD_STATE, C.c_void_p(out.data_ptr()), self._stream()))
        return out.t()

    def set_state(self, state, layout=None):
        s = self._as_soa(state, 13, layout)
        L.check(self.lib.qs_set(self._h, L.QS_FIELD_STATE, C.c_void_p(s.data_ptr()), self._stream()))

    @property
    def ang(self):
        return self._field(L.QS_FIELD_ANG).t()

    @property
    def prev_ang(self):
        return self._field(L.QS_FIELD_ANG).t()

    @property
    def ang_vel(self):
        return self._field(L.QS_FIELD_ANG_VEL).t()

    @property
    def step_effort(self):
        return self._field(L.QS_FIELD_STEP_EFFORT).t()

    @property
    def w(self):
        return self._field(L.QS_FIELD_W).t()

    @property
    def accel(self):
        return self._field(L.QS_FIELD_ACCEL).t()

    @property
    def accelerometer_read(self):
        return self._field(L.QS_FIELD_ACC_READ).t()

    @property
    def mat_rot(self):
        return self._field(L.QS_FIELD_MAT_ROT).t().reshape(self.N, 3, 3)

    @property
    def i(self):
        return self._field(L.QS_FIELD_I)[0]

    @property
    def abs_sum(self):
        return self._field(L.QS_FIELD_ABS_SUM)[0]

    @property
    def ep_return(self):
        return self._field(L.QS_FIELD_EP_RETURN)[0]

    @property
    def episode(self):
        return self._field(L.QS_FIELD_EPISODE)[0]

    @property
    def env_flags(self):
        return self._field(L.QS_FIELD_FLAGS)[0]

    @property
    def sensed_obs(self):
        return self._field(L.QS_FIELD_SENSED_OBS).t()

    @property
    def sensor_state(self):
        return self._field(L.QS_FIELD_SENSOR_STATE).t()

    @property
    def gust_count(self):
        """robust_control=True: wind gusts drawn so far by each env (robust_control.wind, quadrotor_env.py:104-109)."""
        return self._field(L.QS_FIELD_GUST_COUNT)[0]

    # ------------------------------------------------------------------ checkpoint / resume
    def get_checkpoint(self) -> dict:
        """Everything needed to resume: the raw workspace (all SoA rows + statistics) and the RNG key."""
        return {"workspace": self._ws.clone(), "seed": int(self._cfg.seed), "n_envs": self.N,
                "precision": self.precision}

    def set_checkpoint(self, ck: dict):
        if ck["n_envs"] != self.N or ck["precision"] != self.precision or ck["workspace"].numel() != self._ws.numel():
            raise ValueError("checkpoint does not match this handle's configuration")
        self._ws.copy_(ck["workspace"])
        self.seed(ck["seed"])

    # ------------------------------------------------------------------ statistics
    def stats_tensor(self) -> torch.Tensor:
        """Zero-copy (8,) float64 view of the device accumulators (see qs_stats in include/quadsim.h)."""
        v = self._views.get("stats")
        if v is None:
            p = C.c_void_p()
            L.check(self.lib.qs_stats_device(self._h, C.byref(p)))
            off = p.value - self._ws.data_ptr()
            v = self._ws[off:off + 8 * L.QS_STATS_DIM].view(torch.float64)
            self._views["stats"] = v
        return v

    def stats(self, reset: bool = False, all_reduce: bool = False) -> dict:
        """Episode statistics; with all_reduce=True they are summed over all ranks with one NCCL all-reduce
        (the only collective on the path — environment/controller/ppo.py:371-382 does this on the host)."""
        t = self.stats_tensor()
        r = t.clone()
        if reset:
            t.zero_()
        if all_reduce:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                dist.all_reduce(r, op=dist.ReduceOp.SUM)
        names = ["sum_return", "sum_length", "n_episodes", "n_solved", "n_broken", "n_timeout", "sum_effort", "n_steps"]
        vals = r.cpu().tolist()
        out = dict(zip(names, vals))
        ne = max(out["n_episodes"], 1.0)
        out["mean_return"] = out["sum_return"] / ne
        out["mean_length"] = out["sum_length"] / ne
        out["solved_frac"] = out["n_solved"] / ne
        return out
