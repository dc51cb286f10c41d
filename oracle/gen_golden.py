"""TEST INFRASTRUCTURE — generate tests/golden/*.npz by running the UNMODIFIED reference in the build
container (it cannot travel to the GPU box).  Re-run with:  python oracle/gen_golden.py

Fixtures (all float64 unless noted):
  utility_vectors.npz   euler_quat / quat_euler / deriv_quat / quat_rot_mat on random inputs
  drone_eq_vectors.npz  quad.drone_eq in direct and indirect mode (+ quad.f2w clipped / unclipped)
  step_direct.npz       32 envs x 120 steps of quad.step, direct control, random actions, T=1, training
  step_indirect.npz     16 envs x 80 steps, indirect control ([F,M] actions), clipped mixer, T=5
  step_eval.npz         8 envs x 200 steps, training=False + small actions (solved does not end the episode)
  lqr_log.npz           slice of the reference's SHIPPED log classical_controller_results/lqr_log_same_start.npy
                        (written by the author's 2021 run) + the LQR gains of lqr_quad.py:25-111 + initial states
  pid_log.npz           same for pid_log_same_start.npy (first episodes)
  sensor_stats.npz      noise statistics of the reference `sensor` class at hover (canonical call order)
  actor_128.npz         float32 weights of solved/nn_old_solved_128_32000_*.pth (actor only) and a reference
                        closed-loop episode driven by it (ppo_quad_eval.py protocol)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402
from oracle import quad_oracle as qo  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
REF = ref_import.REFERENCE_ROOT


def quiet_quad(ref, *a, **k):
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return ref.quad(*a, **k)


def gen_utility(ref_u, rng):
    n = 128
    ang = rng.uniform(-np.pi, np.pi, (n, 3)) * np.array([1, 0.49, 1])
    q = rng.normal(0, 1, (n, 4))
    qn = q / np.linalg.norm(q, axis=1, keepdims=True)
    w = rng.normal(0, 3, (n, 3))
    out = dict(ang=ang, q=q, qn=qn, w=w)
    out["euler_quat"] = np.array([ref_u.euler_quat(a).flatten() for a in ang])
    out["quat_euler"] = np.array([ref_u.quat_euler(x.reshape(4, 1)) for x in qn])
    out["deriv_quat"] = np.array([ref_u.deriv_quat(ww, x.reshape(4, 1)) for ww, x in zip(w, qn)])
    out["quat_rot_mat"] = np.array([ref_u.quat_rot_mat(x) for x in qn])
    np.savez_compressed(os.path.join(OUT, "utility_vectors.npz"), **out)


def gen_drone_eq(ref, rng):
    n = 256
    x = rng.normal(0, 2, (n, 13))
    x[:, 6:10] = rng.normal(0, 1, (n, 4))          # un-normalised on purpose (drone_eq normalises on read)
    x[:, 10:13] = rng.normal(0, 4, (n, 3))
    a = rng.uniform(-1, 1, (n, 4))
    env = quiet_quad(ref, 0.01, 1000, training=True, direct_control=1, T=1)
    dx_direct = np.array([env.drone_eq(0, xi, ai) for xi, ai in zip(x, a)])
    envi = quiet_quad(ref, 0.01, 1000, training=True, direct_control=0, T=1, clipped=True)
    fm = np.stack([rng.uniform(0, 25, n), rng.normal(0, 0.4, n), rng.normal(0, 0.4, n), rng.normal(0, 0.05, n)], axis=1)
    eff_c, w_c, fmn_c, dx_ind = [], [], [], []
    for xi, f in zip(x, fm):
        se, w, F_new, M_new = envi.f2w(f[0], f[1:4].reshape(3, 1))
        envi.w = w
        u = np.append([F_new], M_new)
        eff_c.append(se); w_c.append(w.flatten()); fmn_c.append(u)
        dx_ind.append(envi.drone_eq(0, xi, u))
    envu = quiet_quad(ref, 0.01, 1000, training=True, direct_control=0, T=1, clipped=False)
    eff_u, w_u, fmn_u = [], [], []
    for f in fm:
        se, w, F_new, M_new = envu.f2w(f[0], f[1:4].reshape(3, 1))
        eff_u.append(se); w_u.append(w.flatten()); fmn_u.append(np.append([F_new], M_new))
    np.savez_compressed(os.path.join(OUT, "drone_eq_vectors.npz"), x=x, a=a, dx_direct=dx_direct, fm=fm,
                        effort_clipped=np.array(eff_c), w_clipped=np.array(w_c), fm_new_clipped=np.array(fmn_c),
                        dx_indirect=np.array(dx_ind), effort_unclipped=np.array(eff_u), w_unclipped=np.array(w_u),
                        fm_new_unclipped=np.array(fmn_u))


def run_reference_batch(ref, n_env, steps, init_states, actions, **kw):
    """Drive n_env reference quads with given (steps,n_env,4) actions; returns dict of per-step records."""
    import scipy.integrate as integ
    envs = [quiet_quad(ref, 0.01, kw.pop("n", 1000) if False else kw.get("n", 1000),
                       training=kw.get("training", True), direct_control=kw.get("direct_control", 1),
                       T=kw.get("T", 1), clipped=kw.get("clipped", True)) for _ in range(n_env)]
    T = kw.get("T", 1)
    rec = dict(obs=np.zeros((steps, n_env, 14)), reward=np.zeros((steps, n_env)), done=np.zeros((steps, n_env), bool),
               state=np.zeros((steps, n_env, 13)), ang=np.zeros((steps, n_env, 3)), ang_vel=np.zeros((steps, n_env, 3)),
               solved=np.zeros((steps, n_env), np.int64), step_effort=np.zeros((steps, n_env, 4)),
               w=np.zeros((steps, n_env, 4)), accel=np.zeros((steps, n_env, 3)), abs_sum=np.zeros((steps, n_env)),
               nfev=np.zeros((steps, n_env), np.int64), clipped_action=np.zeros((steps, n_env, 4)),
               acc_read=np.zeros((steps, n_env, 3)), mat_rot=np.zeros((steps, n_env, 3, 3)))
    reset_obs = np.zeros((T, n_env, 14))
    reset_state = np.zeros((n_env, 13))
    orig = integ.solve_ivp
    last = {}

    def spy(*a, **k):
        r = orig(*a, **k)
        last["nfev"] = r.nfev
        return r

    ref.integrate.solve_ivp = spy
    try:
        for j, e in enumerate(envs):
            s, _ = e.reset(init_states[j].copy())
            reset_obs[:, j] = s
            reset_state[j] = e.state
        for t in range(steps):
            for j, e in enumerate(envs):
                o, r, d = e.step(actions[t, j])
                rec["obs"][t, j] = o[0]; rec["reward"][t, j] = r; rec["done"][t, j] = d
                rec["state"][t, j] = e.state; rec["ang"][t, j] = e.ang; rec["ang_vel"][t, j] = e.ang_vel
                rec["solved"][t, j] = e.solved; rec["step_effort"][t, j] = e.step_effort
                rec["w"][t, j] = e.w.flatten(); rec["accel"][t, j] = e.accel.flatten()
                rec["abs_sum"][t, j] = e.abs_sum; rec["nfev"][t, j] = last["nfev"]
                rec["clipped_action"][t, j] = e.clipped_action
                rec["acc_read"][t, j] = np.asarray(e.accelerometer_read).flatten()
                rec["mat_rot"][t, j] = e.mat_rot
    finally:
        ref.integrate.solve_ivp = orig
    rec["reset_obs"] = reset_obs
    rec["reset_state"] = reset_state
    return rec


def gen_steps(ref, rng):
    # direct control, random actions (episodes end by bounding box within ~40-90 steps; stepping continues, done sticky)
    n_env, steps = 32, 120
    init, _ = qo.sample_reset_state(2024, np.arange(n_env), 0)
    actions = rng.uniform(-1.2, 1.2, (steps, n_env, 4))            # some outside [-1,1] to exercise the clip
    rec = run_reference_batch(ref, n_env, steps, init, actions, n=100, training=True, direct_control=1, T=1)
    np.savez_compressed(os.path.join(OUT, "step_direct.npz"), init=init, actions=actions, n=100, T=1, **rec)

    n_env, steps = 16, 80
    init, _ = qo.sample_reset_state(2025, np.arange(n_env), 0)
    actions = np.stack([rng.uniform(5, 16, (steps, n_env)), rng.normal(0, 0.25, (steps, n_env)),
                        rng.normal(0, 0.25, (steps, n_env)), rng.normal(0, 0.04, (steps, n_env))], axis=2)
    rec = run_reference_batch(ref, n_env, steps, init, actions, n=1000, training=True, direct_control=0, T=5, clipped=True)
    np.savez_compressed(os.path.join(OUT, "step_indirect.npz"), init=init, actions=actions, n=1000, T=5, **rec)

    n_env, steps = 8, 200
    init = np.zeros((n_env, 13)); init[:, 6] = 1.0
    init[:, 1:6:2] = rng.normal(0, 0.02, (n_env, 3))
    init[:, 10:13] = rng.normal(0, 0.02, (n_env, 3))
    init[0, 1:6:2] = 0; init[0, 10:13] = 0                         # env 0 starts exactly solved
    actions = rng.uniform(-0.01, 0.01, (steps, n_env, 4))
    rec = run_reference_batch(ref, n_env, steps, init, actions, n=150, training=False, direct_control=1, T=1)
    np.savez_compressed(os.path.join(OUT, "step_eval.npz"), init=init, actions=actions, n=150, T=1, **rec)


def lqr_gains(clipped=True):
    """Gains exactly as lqr_quad.py:25-111 computes them (clipped branch)."""
    from scipy.linalg import solve_continuous_are as solve_lqr
    I_xx, I_yy, I_zz, M = 16.83e-3, 16.83e-3, 28.34e-3, 1.03
    Q_att = np.array([[5, 0, 0, 0, 0, 0], [0, 1, 0, 0, 0, 0], [0, 0, 5, 0, 0, 0], [0, 0, 0, 1, 0, 0],
                      [0, 0, 0, 0, 0.05, 0], [0, 0, 0, 0, 0, 0.01]]) * 50
    R_att = np.diag(np.ones(4)) * 40
    Q_t = np.array([[1e-08, 0, 0, 0, 0, 0], [0, 1, 0, 0, 0, 0], [0, 0, 1e-08, 0, 0, 0], [0, 0, 0, 1, 0, 0],
                    [0, 0, 0, 0, 1e-08, 0], [0, 0, 0, 0, 0, 0.8]]) * 10
    R_t = np.diag(np.ones(3)) * 10
    A_att = np.array([[0, 1, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0], [0, 0, 0, 1, 0, 0], [0, 0, 0, 0, 0, 0],
                      [0, 0, 0, 0, 0, 1], [0, 0, 0, 0, 0, 0]])
    B_att = np.array([[0, 0, 0, 0], [0, 1 / I_xx, 0, 0], [0, 0, 0, 0], [0, 0, 1 / I_yy, 0], [0, 0, 0, 0],
                      [0, 0, 0, 1 / I_zz]])
    K_att = -np.dot(np.linalg.inv(R_att), np.dot(B_att.T, solve_lqr(A_att, B_att, Q_att, R_att)))
    A_t = A_att
    B_t = np.array([[0, 0, 0], [1, 0, 0], [0, 0, 0], [0, 1, 0], [0, 0, 0], [0, 0, 1]]) / M
    K_t = -np.dot(np.linalg.inv(R_t), np.dot(B_t.T, solve_lqr(A_t, B_t, Q_t, R_t)))
    return K_t, K_att


def gen_logs(ref):
    res = os.path.join(REF, "environment", "controller", "classical_controller_results")
    lqr = np.load(os.path.join(res, "lqr_log_same_start.npy"))
    pid = np.load(os.path.join(res, "pid_log_same_start.npy"))
    K_t, K_att = lqr_gains()
    # initial states of the 20 episodes for seed 1 (robust RNG draws post-date the logs: suppressed)
    np.random.seed(1)
    inits = []
    u = ref_import.load_reference_utility()
    for _ in range(20):
        ang = np.random.rand(3) - 0.5
        s = np.zeros(13)
        Q_in = u.euler_quat(ang)
        s[0:5:2] = np.clip((np.random.normal([0, 0, 0], 2)), -2.5, 2.5)
        s[1:6:2] = np.clip((np.random.normal([0, 0, 0], 2)), -5, 5)
        s[6:10] = Q_in.T
        s[10:13] = np.clip((np.random.normal([0, 0, 0], 2)), -15, 7.5)
        inits.append(s)
    np.savez_compressed(os.path.join(OUT, "lqr_log.npz"), log=lqr[:6], K_t=K_t, K_att=K_att, inits=np.array(inits),
                        source="environment/controller/classical_controller_results/lqr_log_same_start.npy[:6]")
    np.savez_compressed(os.path.join(OUT, "pid_log.npz"), log=pid[:4], inits=np.array(inits),
                        source="environment/controller/classical_controller_results/pid_log_same_start.npy[:4]")


def gen_actor(ref):
    import glob
    import torch
    f = glob.glob(os.path.join(REF, "environment", "controller", "solved", "nn_old_solved_128_32000_*.pth"))[0]
    sd = torch.load(f, map_location="cpu")
    W = {k.replace(".", "_"): v.to(torch.float32).numpy() for k, v in sd.items() if k.startswith("actor.")}
    # reference closed-loop episode, ppo_quad_eval.py:32-66 protocol (training=False, T=5, fp32 forward)
    sys.path.insert(0, REF)
    from environment.controller.dl_auxiliary import dl_in_gen
    env = quiet_quad(ref, 0.01, 500, training=False, euler=0, direct_control=1, T=5)
    aux = dl_in_gen(5, 13, 4)
    init, _ = qo.sample_reset_state(99, np.arange(1), 0)
    state, action = env.reset(init[0].copy())
    aux.reset()
    in_nn = aux.dl_input(state, action)

    def actor(x):
        h = np.tanh(W["actor_0_weight"] @ x + W["actor_0_bias"])
        h = np.tanh(W["actor_2_weight"] @ h + W["actor_2_bias"])
        return np.tanh(W["actor_4_weight"] @ h + W["actor_4_bias"])

    model = torch.nn.Sequential(torch.nn.Linear(75, 128), torch.nn.Tanh(), torch.nn.Linear(128, 128), torch.nn.Tanh(),
                                torch.nn.Linear(128, 4), torch.nn.Tanh())
    model.load_state_dict({k.replace("actor.", ""): v.float() for k, v in sd.items() if k.startswith("actor.")})
    steps = 300
    obs = np.zeros((steps, 14)); acts = np.zeros((steps, 4)); nn_in = np.zeros((steps, 75), np.float32)
    states = np.zeros((steps, 13))
    for t in range(steps):
        nn_in[t] = in_nn
        a = model(torch.FloatTensor(in_nn)).detach().numpy()
        s, _, _ = env.step(a)
        in_nn = aux.dl_input(s, np.array([a]))
        obs[t] = s[0]; acts[t] = a; states[t] = env.state
    np.savez_compressed(os.path.join(OUT, "actor_128.npz"), init=init[0], reset_obs=state, obs=obs, actions=acts,
                        nn_in=nn_in, states=states, **W)


def gen_sensor_stats(ref):
    """Statistics of the reference `sensor` class at hover in the canonical call order (rl_worker.py:164-175):
    the Philox-driven model cannot match NumPy's MT19937 draws, so distributions are compared instead."""
    sys.path.insert(0, REF)
    from environment.quaternion_euler_utility import deriv_quat
    qv_std, dv_std, pos_end = [], [], []
    for seed in range(4):
        env = quiet_quad(ref, 0.01, 10 ** 6, training=False, direct_control=1, T=1)
        np.random.seed(seed)
        init = np.zeros(13); init[6] = 1
        env.reset(init.copy())
        sen = ref.sensor(env)
        env.state = env.state.copy()
        sen.reset()
        # neutralise the aliasing of sensor.reset (:636-638) so that the TRUE state is not perturbed
        sen.quaternion_t0 = sen.quaternion_t0.copy(); sen.position_t0 = sen.position_t0.copy(); sen.velocity_t0 = sen.velocity_t0.copy()
        obs = []
        for t in range(1500):
            env.step(np.zeros(4))
            _, v, p = sen.accel_int(); qg = sen.gyro_int().copy(); w = sen.gyro(); qv = deriv_quat(w, qg); sen.gps(); sen.triad()
            obs.append(np.concatenate([[p[0], v[0], p[1], v[1], p[2], v[2]], qg, qv]))
        obs = np.array(obs)
        qv_std.append(obs[:, 11:14].std(0)); dv_std.append(np.diff(obs[:, [1, 3, 5]], axis=0).std(0)); pos_end.append(obs[-1, [0, 2, 4]])
    np.savez_compressed(os.path.join(OUT, "sensor_stats.npz"), qv_std=np.array(qv_std), dv_std=np.array(dv_std),
                        pos_end=np.array(pos_end), steps=1500)


def _lift_function(path, name):
    """Source of a top-level function of a reference script that cannot be imported (module-level side effects), via ast."""
    import ast
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            return compile(ast.Module(body=[node], type_ignores=[]), os.path.basename(path) + ":" + name, "exec")
    raise KeyError(name)


SENSOR_STATE = lambda sen: np.concatenate([[sen.a_b_accel, sen.g_b, sen.a_b_d, sen.g_b_d], np.asarray(sen.velocity_t0).flatten(),
                                           np.asarray(sen.position_t0).flatten(), np.asarray(sen.quaternion_t0).flatten(),
                                           sen.R[:, 2], np.asarray(sen.acceleration_t0).flatten()])


def gen_sensor_vectors(ref):
    """The reference's `sensor` class (quadrotor_env.py:579-724) driven DETERMINISTICALLY: np.random.normal / np.random.random
    are replaced by a replay of recorded standard draws (oracle/replay_rng.py), so every output of every method — incl. TRIAD,
    GPS and the bias drift — is a pure function of the stored inputs.  Three scenarios:
      A  4 sensors x 40 steps, canonical call order accel_int, gyro_int, gyro, gps, triad (rl_worker.py:164-175)
      B  2 sensors x 40 steps through the reference's OWN sensor_sp (visual_landing/math_trajectory.py:61-83, lifted with ast)
         with its GPS blend switched on (GPS = True, GPS_P = 30) — and the same function with GPS = False on scenario A's
         inputs must reproduce scenario A's observation (asserted here)
      C  2 sensors x 20 steps, methods called in another order (gyro, triad, gps, gyro_int, accel, accel_int): every method on its own
    Harness-side fix: sensor.reset keeps VIEWS of quad.state (:636-638) and gyro_int then writes through them into the TRUE
    quaternion once per episode; the views are replaced by copies (no implementation reproduces that aliasing)."""
    from oracle.replay_rng import ReplayRNG
    ref_u = ref_import.load_reference_utility()
    rng = np.random.default_rng(20261019)
    sp_code = _lift_function(os.path.join(REF, "visual_landing", "math_trajectory.py"), "sensor_sp")

    def sensor_sp_with(gps, gps_p):
        ns = {"np": np, "deriv_quat": ref_u.deriv_quat, "GPS": gps, "GPS_P": gps_p}
        exec(sp_code, ns)
        return ns["sensor_sp"]

    def make(n_env, steps, nz, seed_off):
        init = np.zeros((n_env, 13)); init[:, 6] = 1
        init[:, 0:5:2] = rng.normal(0, 1.0, (n_env, 3))
        init[:, 1:6:2] = rng.normal(0, 0.3, (n_env, 3)); init[:, 10:13] = rng.normal(0, 0.3, (n_env, 3))
        ang = rng.uniform(-0.3, 0.3, (n_env, 3))
        init[:, 6:10] = np.array([ref_u.euler_quat(a).flatten() for a in ang])
        return dict(init=init, actions=rng.uniform(-0.25, 0.25, (steps, n_env, 4)), z=rng.standard_normal((steps, n_env, nz)),
                    u=rng.random((n_env, 6)))

    def start(d, j):
        env = quiet_quad(ref, 0.01, 10 ** 6, training=False, direct_control=1, T=2)
        env.reset(d["init"][j].copy())
        with ReplayRNG([], d["u"][j]):
            sen = ref.sensor(env)                  # __init__: bias_reset (3 uniforms)
            sen.reset()                            # reset: bias_reset again (3 uniforms)
        sen.quaternion_t0 = sen.quaternion_t0.copy(); sen.position_t0 = sen.position_t0.copy(); sen.velocity_t0 = sen.velocity_t0.copy()
        d.setdefault("reset_state", np.zeros((d["init"].shape[0], 13)))[j] = env.state       # quad.state sensor.reset reads
        return env, sen

    def truth(env):
        return dict(state=env.state.copy(), acc_read=np.asarray(env.accelerometer_read).flatten().copy(), mat_rot=env.mat_rot.copy(),
                    f_m=float(env.f_in.flatten()[2]) / ref.M)

    out = {}
    # ---- A: canonical order
    nA, KA = 4, 40
    A = make(nA, KA, 27, 0)
    recA = {k: [] for k in ("state", "acc_read", "mat_rot", "f_m", "accel_int", "gyro_int", "gyro", "gps", "triad_q", "triad_R", "sens_state", "obs")}
    sp_off = sensor_sp_with(False, 30.0)
    for j in range(nA):
        env, sen = start(A, j)
        env2, sen2 = start(A, j)                   # the same sensor through sensor_sp(GPS = False)
        rows = {k: [] for k in recA}
        for t in range(KA):
            env.step(A["actions"][t, j]); env2.step(A["actions"][t, j])
            for k, v in truth(env).items():
                rows[k].append(v)
            with ReplayRNG(A["z"][t, j]):
                acc, vel, pos = sen.accel_int(); qg = np.array(sen.gyro_int()).copy(); w = sen.gyro(); pg, vg = sen.gps(); qt, Rt = sen.triad()
            with ReplayRNG(A["z"][t, j]):
                obs_sp = sp_off(sen2)[0]
            qv = ref_u.deriv_quat(w, qg).flatten()
            obs = np.concatenate([[pos[0], vel[0], pos[1], vel[1], pos[2], vel[2]], qg, qv])
            assert np.array_equal(obs, obs_sp), "sensor_sp(GPS=False) != canonical composition"
            rows["accel_int"].append(np.concatenate([acc, vel, pos])); rows["gyro_int"].append(qg); rows["gyro"].append(w)
            rows["gps"].append(np.concatenate([pg, vg])); rows["triad_q"].append(qt); rows["triad_R"].append(Rt.copy())
            rows["sens_state"].append(SENSOR_STATE(sen)); rows["obs"].append(obs)
        for k in recA:
            recA[k].append(np.array(rows[k]))
    for k, v in recA.items():
        out["A_" + k] = np.swapaxes(np.array(v), 0, 1)            # (steps, env, ...)
    for k, v in A.items():
        out["A_" + k] = v
    # ---- B: the reference's sensor_sp with the GPS blend on
    nB, KB, gps_p = 2, 40, 30.0
    B = make(nB, KB, 27, 1)
    sp_on = sensor_sp_with(True, gps_p)
    recB = {k: [] for k in ("state", "acc_read", "mat_rot", "f_m", "obs", "sens_state")}
    for j in range(nB):
        env, sen = start(B, j)
        rows = {k: [] for k in recB}
        for t in range(KB):
            env.step(B["actions"][t, j])
            for k, v in truth(env).items():
                rows[k].append(v)
            with ReplayRNG(B["z"][t, j]):
                rows["obs"].append(sp_on(sen)[0])
            rows["sens_state"].append(SENSOR_STATE(sen))
        for k in recB:
            recB[k].append(np.array(rows[k]))
    for k, v in recB.items():
        out["B_" + k] = np.swapaxes(np.array(v), 0, 1)
    for k, v in B.items():
        out["B_" + k] = v
    out["B_gps_blend"] = gps_p
    # ---- C: every method on its own, other order
    nC, KC = 2, 20
    Cc = make(nC, KC, 30, 2)
    recC = {k: [] for k in ("state", "acc_read", "mat_rot", "f_m", "gyro", "triad_q", "triad_R", "gps", "gyro_int", "accel", "accel_int", "sens_state")}
    for j in range(nC):
        env, sen = start(Cc, j)
        rows = {k: [] for k in recC}
        for t in range(KC):
            env.step(Cc["actions"][t, j])
            for k, v in truth(env).items():
                rows[k].append(v)
            with ReplayRNG(Cc["z"][t, j]):
                w = sen.gyro(); qt, Rt = sen.triad(); pg, vg = sen.gps(); qg = np.array(sen.gyro_int()).copy(); ac = sen.accel()
                acc, vel, pos = sen.accel_int()
            rows["gyro"].append(w); rows["triad_q"].append(qt); rows["triad_R"].append(Rt.copy()); rows["gps"].append(np.concatenate([pg, vg]))
            rows["gyro_int"].append(qg); rows["accel"].append(ac); rows["accel_int"].append(np.concatenate([acc, vel, pos]))
            rows["sens_state"].append(SENSOR_STATE(sen))
        for k in recC:
            recC[k].append(np.array(rows[k]))
    for k, v in recC.items():
        out["C_" + k] = np.swapaxes(np.array(v), 0, 1)
    for k, v in Cc.items():
        out["C_" + k] = v
    np.savez_compressed(os.path.join(OUT, "sensor_vectors.npz"), **out)


def gen_script_logs():
    """What the reference's UNMODIFIED controller scripts produce HERE (bytecode build oracle/_ref, oracle/ref_runtime.py)
    against its five shipped logs: per-episode max |diff|.  The scripts are chaotic in a few episodes (the LQR diverges), so
    the author's 2021 logs reproduce to 1e-13 in most episodes and not at all in some; the test of the CUDA-backed drop-in
    (tests/test_reference_scripts.py) holds it to 1e-8 exactly where the reference reproduces itself."""
    from oracle import build_ref, ref_runtime as rr
    build_ref.build()
    out = {}
    for key, script, sw, log in rr_cases():
        got = list(rr.run_script(script, overlay=False, switches=sw).values())[0]
        ref_log = rr.shipped_log(log)
        out[key + "_self_err"] = np.nanmax(np.abs(got - ref_log), axis=(1, 2))
        out[key + "_self_err_first50"] = np.nanmax(np.abs(got[:, :50] - ref_log[:, :50]), axis=(1, 2))
        if key != "rl":                 # episodes the 2021 log does not pin: keep what the reference produces here (first 100 steps)
            for ep in np.nonzero(out[key + "_self_err"] > 1e-9)[0]:
                out["%s_here_ep%d" % (key, ep)] = got[ep, :100]
    np.savez_compressed(os.path.join(OUT, "script_logs_selfcheck.npz"), **out)


def rr_cases():
    return [("lqr", "environment/controller/lqr_quad", {}, "lqr_log_same_start.npy"),
            ("lqr_nc", "environment/controller/lqr_quad", {"clipped": False}, "lqr_log_same_start_not_clipped.npy"),
            ("pid", "environment/controller/pid_vel_control", {}, "pid_log_same_start.npy"),
            ("pid_nc", "environment/controller/pid_vel_control", {"clipped": False}, "pid_log_same_start_not_clipped.npy"),
            ("rl", "environment/controller/ppo_quad_eval", {}, "rl_log_same_start.npy")]


def gen_ppo_vectors():
    """PPO.get_advantages and the loss of PPO.update (environment/controller/ppo.py:125-141, :183-201) evaluated by the
    reference's own code: ppo.py is a script (argparse + training loop at import), so the two method bodies are lifted
    out of its source with `ast` and executed unmodified against random rollouts."""
    import ast
    import torch
    src = open(os.path.join(REF, "environment", "controller", "ppo.py")).read()
    tree = ast.parse(src)
    fn = None
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef) and node.name == "PPO":
            for f in node.body:
                if isinstance(f, ast.FunctionDef) and f.name == "get_advantages":
                    fn = f
    assert fn is not None
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = {"np": np}
    exec(compile(mod, "ppo.py:get_advantages", "exec"), ns)
    rng = np.random.default_rng(77)
    out = {}
    for i, L_ in enumerate((37, 400)):
        rewards = rng.normal(0, 1, L_)
        values = np.concatenate([rng.normal(0, 2, L_), [0.0]])             # memory.values.append(0)  ppo.py:384
        term = rng.random(L_) < 0.05
        term[-1] = True                                                     # a worker only returns after `done` (:283)
        ret, adv = ns["get_advantages"](None, torch.tensor(values), np.logical_not(term), rewards)
        out.update({"rewards%d" % i: rewards, "values%d" % i: values, "terminals%d" % i: term, "returns%d" % i: ret, "adv%d" % i: adv})
    # model.py:62-66,74-88 + ppo.py:183-201: loss of one update step on a random batch, by the reference's ActorCritic
    sys.path.insert(0, REF)
    from environment.controller.model import ActorCritic
    torch.manual_seed(5)
    pol = ActorCritic(32, 75, 4, 0.1, True).double()
    B = 64
    st = torch.randn(B, 1, 75, dtype=torch.double); ac = torch.randn(B, 1, 4, dtype=torch.double) * 0.3
    old_lp = torch.randn(B, 1, 4, dtype=torch.double) * 0.1 + 1.0
    adv_t = torch.randn(B, dtype=torch.double); ret_t = torch.randn(B, dtype=torch.double)
    logprobs, state_values, dist_entropy = pol.evaluate(st, ac)
    ratios = torch.exp(logprobs.sum(axis=2).flatten() - old_lp.sum(axis=2).detach().flatten())
    surr1 = ratios * adv_t
    surr2 = torch.clamp(ratios, 1 - 0.2, 1 + 0.2) * adv_t
    critic_loss = 0.5 * torch.nn.MSELoss()(state_values, ret_t)
    loss = (-torch.min(surr1, surr2) + critic_loss - 0.006 * dist_entropy.sum(axis=2).flatten()).mean()
    loss.backward()
    sd = {k.replace(".", "_"): v.detach().numpy() for k, v in pol.state_dict().items()}
    grads = {"grad_" + k.replace(".", "_"): p_.grad.numpy() for k, p_ in pol.named_parameters() if p_.grad is not None}
    out.update(dict(loss_states=st.numpy()[:, 0], loss_actions=ac.numpy()[:, 0], loss_old_logprobs=old_lp.numpy()[:, 0],
                    loss_adv=adv_t.numpy(), loss_returns=ret_t.numpy(), loss_value=float(loss.detach()), **{"w_" + k: v for k, v in sd.items()}, **grads))
    np.savez_compressed(os.path.join(OUT, "ppo_vectors.npz"), **out)


def gen_robust_vectors(ref, rng):
    """robust_control (quadrotor_env.py:84-109 and the guarded branches :235,:265,:318,:341,:360,:381) is dead code in the
    reference (quad.robust_control is hard-wired False, :183) and one branch does not run as written on current NumPy
    (episode_m is a length-1 array -> ragged return value of drone_eq): the perturbations are injected directly, episode_m as
    the scalar the arithmetic intends, and wind() is replaced by a fixed vector (its ramp is pinned separately below)."""
    n = 128
    x = rng.normal(0, 2, (n, 13))
    x[:, 6:10] = rng.normal(0, 1, (n, 4))
    x[:, 10:13] = rng.normal(0, 4, (n, 3))
    a = rng.uniform(-1, 1, (n, 4))
    kf = rng.random((n, 4)) * 0.1
    ir = rng.random((n, 4)) * 0.1
    m = rng.normal(0, 0.3, n)
    Jp = rng.normal(0, 0.1, (n, 3))
    wind = rng.normal(0, [5, 5, 2], (n, 3))
    env = quiet_quad(ref, 0.01, 1000, training=True, direct_control=1, T=1)
    env.robust_control = True
    dx_direct = []
    for k in range(n):
        rp = env.robust_parameters
        rp.episode_kf = kf[k].copy(); rp.episode_m = float(m[k]); rp.episode_ir = ir[k].copy(); rp.episode_J = np.eye(3) * Jp[k]
        rp.wind = lambda i, w=wind[k]: w.reshape(3, 1).copy()
        dx_direct.append(np.array(env.drone_eq(0, x[k], a[k]), dtype=np.float64))
    envi = quiet_quad(ref, 0.01, 1000, training=True, direct_control=0, T=1, clipped=True)
    envi.robust_control = True
    fm = np.stack([rng.uniform(0, 25, n), rng.normal(0, 0.4, n), rng.normal(0, 0.4, n), rng.normal(0, 0.05, n)], axis=1)
    eff, wr, fmn, dx_ind = [], [], [], []
    for k in range(n):
        rp = envi.robust_parameters
        rp.episode_kf = kf[k].copy(); rp.episode_m = float(m[k]); rp.episode_ir = ir[k].copy(); rp.episode_J = np.eye(3) * Jp[k]
        rp.wind = lambda i, w=wind[k]: w.reshape(3, 1).copy()
        se, w, F_new, M_new = envi.f2w(fm[k, 0], fm[k, 1:4].reshape(3, 1))
        envi.w = w
        u = np.append([F_new], M_new)
        eff.append(np.asarray(se).flatten()); wr.append(w.flatten()); fmn.append(u)
        dx_ind.append(np.array(envi.drone_eq(0, x[k], u), dtype=np.float64))
    # wind(): the stateful ramp, called once per step over two "episodes" (i restarts at 1), NumPy's stream seeded
    rc = ref.robust_control()
    np.random.seed(11)
    rc.reset()
    i_seq = np.concatenate([np.arange(1, 701), np.arange(1, 1201)])
    winds, gusts = [], []
    for i in i_seq:
        winds.append(np.asarray(rc.wind(int(i))).flatten().copy())
        gusts.append(np.asarray(rc.gust).flatten().copy())
    np.savez_compressed(os.path.join(OUT, "robust_vectors.npz"), x=x, a=a, kf=kf, ir=ir, m=m, J=Jp, wind=wind,
                        dx_direct=np.array(dx_direct), fm=fm, effort=np.array(eff), w_rotor=np.array(wr), fm_new=np.array(fmn),
                        dx_indirect=np.array(dx_ind), wind_i=i_seq, wind_out=np.array(winds), wind_gust=np.array(gusts),
                        gust_period=rc.gust_period)


sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import MISSION_CASES  # noqa: E402  (the same argument sets the test replays)


def gen_mission_vectors():
    """mission_control/mission_control.py executed as it is: trajectories, velocities and the get_error stream (three calls
    past the end, so the extrapolation branch :69-70 is recorded) of the cases in MISSION_CASES."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_mission", os.path.join(REF, "mission_control", "mission_control.py"))
    rm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rm)
    out = {}
    for name, f in MISSION_CASES.items():
        m = rm.mission(0.01)
        f(m)
        out[name + "_trajectory"] = m.trajectory.copy()
        out[name + "_velocity"] = m.velocity.copy()
        out[name + "_errors"] = np.array([m.get_error(0) for _ in range(m.trajectory_total_steps + 3)])
    np.savez_compressed(os.path.join(OUT, "mission_vectors.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    only = [a for a in sys.argv[1:] if not a.startswith("-")]
    ref = ref_import.load_reference()
    ref_u = ref_import.load_reference_utility()
    rng = np.random.default_rng(20261017)
    gens = [("utility", lambda: gen_utility(ref_u, rng)), ("drone_eq", lambda: gen_drone_eq(ref, rng)), ("steps", lambda: gen_steps(ref, rng)),
            ("logs", lambda: gen_logs(ref)), ("actor", lambda: gen_actor(ref)), ("sensor_stats", lambda: gen_sensor_stats(ref)),
            ("ppo", gen_ppo_vectors), ("robust", lambda: gen_robust_vectors(ref, np.random.default_rng(20261018))),
            ("mission", gen_mission_vectors), ("sensor", lambda: gen_sensor_vectors(ref)), ("script_logs", gen_script_logs)]
    # utility / drone_eq / steps share one generator stream: regenerate them together (no argument) or not at all
    for name, fn in gens:
        if not only or name in only:
            fn()
    for f in sorted(os.listdir(OUT)):
        print("%-28s %8d bytes" % (f, os.path.getsize(os.path.join(OUT, f))))


if __name__ == "__main__":
    main()
