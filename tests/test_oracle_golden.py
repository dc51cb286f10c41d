"""CPU tests: the oracle (oracle/quad_oracle.py) against fixtures generated from the unmodified reference
(oracle/gen_golden.py) and against the reference's own shipped trajectory log."""
import numpy as np
import pytest

from conftest import load_golden, rel_err, lqr_action
from oracle import quad_oracle as qo


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32_10
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, exp in kat:
        assert [int(v) for v in qo.philox4x32_10([ctr], [key])[0]] == exp


def test_utility_vectors():
    g = load_golden("utility_vectors.npz")
    assert rel_err(qo.euler_quat(g["ang"]), g["euler_quat"]) < 1e-14
    assert rel_err(qo.quat_euler(g["qn"]), g["quat_euler"]) < 1e-13
    assert rel_err(qo.deriv_quat(g["w"], g["qn"]), g["deriv_quat"]) < 1e-14
    assert rel_err(qo.quat_rot_mat(g["qn"]), g["quat_rot_mat"]) < 1e-14
    # round trip (SURVEY.md §8(c) known-answer identity)
    ang = g["ang"] * np.array([0.9, 0.9, 0.9])
    assert np.max(np.abs(qo.quat_euler(qo.euler_quat(ang)) - ang)) < 1e-12


def test_drone_eq_vectors():
    g = load_golden("drone_eq_vectors.npz")
    w, F, M = qo.f2F(g["a"])
    assert rel_err(qo.drone_eq(g["x"], F, M, w), g["dx_direct"]) < 1e-10
    eff, w, Fn, Mn = qo.f2w(g["fm"][:, 0], g["fm"][:, 1:4], clipped=True)
    assert rel_err(eff, g["effort_clipped"]) < 1e-12
    assert rel_err(w, g["w_clipped"]) < 1e-12
    assert rel_err(np.concatenate([Fn[:, None], Mn], axis=1), g["fm_new_clipped"]) < 1e-12
    assert rel_err(qo.drone_eq(g["x"], Fn, Mn, w), g["dx_indirect"]) < 1e-10
    eff, w, Fn, Mn = qo.f2w(g["fm"][:, 0], g["fm"][:, 1:4], clipped=False)
    assert rel_err(eff, g["effort_unclipped"]) < 1e-12
    assert rel_err(w, g["w_unclipped"]) < 1e-12


def test_hover_is_equilibrium():
    # hover: a=0 (direct), level, at rest -> all derivatives zero (az = 4*c8/M - G = 0)
    x = np.zeros((1, 13)); x[0, 6] = 1
    w, F, M = qo.f2F(np.zeros((1, 4)))
    assert np.max(np.abs(qo.drone_eq(x, F, M, w))) < 1e-14
    eff, w, Fn, Mn = qo.f2w(np.array([qo.M * qo.G]), np.zeros((1, 3)))
    assert abs(w[0, 0] - 419.777) < 1e-3


@pytest.mark.parametrize("name,direct,training", [("step_direct.npz", 1, True), ("step_indirect.npz", 0, True),
                                                   ("step_eval.npz", 1, False)])
def test_step_trajectories(name, direct, training):
    """Closed trajectories: oracle driven with the reference's initial states and actions reproduces every
    per-step quantity the reference produced, with identical done flags and RHS-evaluation counts."""
    g = load_golden(name)
    n_env = g["init"].shape[0]
    env = qo.BatchQuadOracle(n_env, 0.01, int(g["n"]), training=training, direct_control=direct, T=int(g["T"]),
                             integrator="rk45")
    oh, _ = env.reset(g["init"])
    assert rel_err(oh, g["reset_obs"]) < 1e-11
    worst = 0.0
    for t in range(g["actions"].shape[0]):
        obs, rew, done = env.step(g["actions"][t])
        ok = ~np.isnan(g["obs"][t]).any(axis=1)
        worst = max(worst, rel_err(obs[ok], g["obs"][t][ok]), rel_err(rew[ok], g["reward"][t][ok]),
                    rel_err(env.ang[ok], g["ang"][t][ok]), rel_err(env.ang_vel[ok], g["ang_vel"][t][ok], floor=1.0),
                    rel_err(env.step_effort[ok], g["step_effort"][t][ok]), rel_err(env.w[ok], g["w"][t][ok]),
                    rel_err(env.accel[ok], g["accel"][t][ok]), rel_err(env.abs_sum[ok], g["abs_sum"][t][ok]),
                    rel_err(env.accelerometer_read[ok], g["acc_read"][t][ok]),
                    rel_err(env.mat_rot[ok], g["mat_rot"][t][ok]),
                    rel_err(env.clipped_action[ok], g["clipped_action"][t][ok]))
        assert np.array_equal(done, g["done"][t]), "done differs at step %d" % t
        assert np.array_equal(env.solved, g["solved"][t]), "solved differs at step %d" % t
        assert np.array_equal(env.nfev[ok], g["nfev"][t][ok]), "nfev differs at step %d" % t
    assert worst < 1e-9, worst


def test_shipped_lqr_log():
    """The reference's own golden log (written by the author's 2021 NumPy/SciPy) is reproduced by the oracle
    driven by a restatement of lqr_quad.py's control law: pins RK45 + drone_eq + f2w + quat_euler + ang_vel."""
    g = load_golden("lqr_log.npz")
    env = qo.BatchQuadOracle(1, 0.01, 500, training=True, direct_control=0, T=1, clipped=True, integrator="rk45")
    for ep in (1, 2):
        # episodes run back-to-back in the script: prev_ang carries over (quirk), so replay episode ep-1's tail cheaply
        # by seeding prev_ang from the log (last Euler angles of the previous episode)
        env.prev_ang = g["log"][ep - 1, -1, 3:6][None, :].copy()
        env.reset(g["inits"][ep][None, :])
        euler_t_ant = env.ang[0]
        worst = 0.0
        for i in range(200):
            action, euler_t = lqr_action(g["K_t"], g["K_att"], env.state[0], env.ang[0], env.ang_vel[0], euler_t_ant)
            euler_t_ant = euler_t
            env.step(action[None, :])
            row = np.concatenate((env.state[0, 1:6:2], env.ang[0], env.ang_vel[0], env.step_effort[0]))
            worst = max(worst, float(np.max(np.abs(row - g["log"][ep, i]))))
        assert worst < 1e-9, (ep, worst)


def test_rk4_converges_to_rk45():
    """The fixed-step RK4 (production integrator) agrees with the RK45 replica within the FP32-mode bound."""
    g = load_golden("step_direct.npz")
    n_env = g["init"].shape[0]
    a = qo.BatchQuadOracle(n_env, 0.01, 100, integrator="rk45")
    b = qo.BatchQuadOracle(n_env, 0.01, 100, integrator="rk4", substeps=1)
    a.reset(g["init"]); b.reset(g["init"])
    for t in range(10):
        oa, _, _ = a.step(g["actions"][t])
        b.previous_state = a.previous_state.copy() if t else b.previous_state   # teacher-forced after step 0
        ob, _, _ = b.step(g["actions"][t]) if t == 0 else b.step(g["actions"][t])
    # single teacher-forced step comparison
    b.previous_state = a.previous_state.copy()
    oa, _, _ = a.step(g["actions"][10]); ob, _, _ = b.step(g["actions"][10])
    ok = ~np.isnan(oa).any(axis=1)
    assert np.max(np.abs(oa[ok] - ob[ok]) / (1e-5 + 1e-4 * np.abs(oa[ok]))) < 0.5


def test_reset_sampler_distribution():
    st, ang = qo.sample_reset_state(3, np.arange(20000), 0)
    assert np.all(np.abs(ang) <= 0.5)
    assert np.all(np.abs(st[:, 0:5:2]) <= 2.5) and np.all(np.abs(st[:, 1:6:2]) <= 5)
    assert st[:, 10:13].max() <= 7.5 and st[:, 10:13].min() >= -15          # asymmetric clip (reference :445)
    assert abs(np.linalg.norm(st[:, 6:10], axis=1) - 1).max() < 1e-12
    v = st[:, 10:13]
    inner = v[(v > -7) & (v < 7)]
    assert abs(inner.mean()) < 0.05 and abs(inner.std() - 2.0) < 0.08       # N(0,2) truncated at 3.5 sigma
    # different episodes / envs give different streams
    st2, _ = qo.sample_reset_state(3, np.arange(20000), 1)
    assert np.abs(st - st2).max() > 1


def test_sensor_oracle_statistics_match_reference_sensor():
    """The Philox-driven sensor restatement has the noise statistics of the reference `sensor` class at hover."""
    g = load_golden("sensor_stats.npz")
    init = np.zeros((4, 13)); init[:, 6] = 1
    o = qo.BatchQuadOracle(4, 0.01, 10 ** 6, training=False, T=1, integrator="rk4")
    o.reset(init)
    s = qo.SensorOracle(4, 0.01)
    ids, ep = np.arange(4), np.zeros(4, dtype=np.int64)
    s.reset(0, ids, ep, o.state)
    obs = []
    for t in range(int(g["steps"])):
        o.step(np.zeros((4, 4)))
        z = qo.sensor_normals(0, ids, ep, o.i)
        obs.append(s.step(z, o.state, o.accelerometer_read, o.mat_rot, o.f_in / qo.M))
    obs = np.array(obs)                                   # (steps, 4, 14)
    qv_std = obs[:, :, 11:14].std(axis=0).mean()
    dv_std = np.diff(obs[:, :, [1, 3, 5]], axis=0).std(axis=0).mean()
    assert abs(qv_std / g["qv_std"].mean() - 1) < 0.06, (qv_std, g["qv_std"].mean())
    assert abs(dv_std / g["dv_std"].mean() - 1) < 0.06, (dv_std, g["dv_std"].mean())
    # dead-reckoning drift after 15 s is a random walk of the same scale (well inside 4x of the reference's spread)
    ref_scale = np.abs(g["pos_end"]).mean()
    assert 0.2 * ref_scale < np.abs(obs[-1][:, [0, 2, 4]]).mean() < 5 * ref_scale


def test_shipped_pid_log_and_lqr_gains():
    """Controller restatements are pinned by the reference's own shipped logs: the cascaded PID of
    pid_vel_control.py:29-127 driving the oracle reproduces classical_controller_results/pid_log_same_start.npy
    (T=5, 500 steps, episodes back to back -> prev_ang carried over), and lqr_gains() equals the gains the script computes."""
    g = load_golden("lqr_log.npz")
    K_t, K_att = qo.lqr_gains()
    assert np.abs(K_t - g["K_t"]).max() < 1e-12 and np.abs(K_att - g["K_att"]).max() < 1e-12
    gp = load_golden("pid_log.npz")
    E = 3
    env = qo.BatchQuadOracle(E, 0.01, 500, training=True, direct_control=0, T=5, clipped=True, integrator="rk45")
    env.prev_ang[1:] = gp["log"][:E - 1, -1, 3:6]
    env.reset(gp["inits"][:E])
    ctl = qo.PidControllerOracle(E)
    action = np.tile(np.array([9.82 * 1.03, 0, 0, 0]), (E, 1))                # pid_vel_control.py:144
    worst = 0.0
    for j in range(300):
        env.step(action)
        action = ctl.control(env.state, env.ang, np.zeros(3), 0.0)
        row = np.concatenate([env.state[:, 1:6:2], env.ang, env.ang_vel, env.step_effort], axis=1)
        worst = max(worst, float(np.abs(row - gp["log"][:E, j]).max()))
    assert worst < 1e-9, worst


def test_lqr_law_equals_script_restatement():
    """The batched lqr_law equals the per-env restatement used to pin the oracle against the shipped LQR log."""
    g = load_golden("lqr_log.npz")
    rng = np.random.default_rng(0)
    st = rng.normal(size=(64, 13)); ang = rng.uniform(-1, 1, (64, 3)); av = rng.normal(size=(64, 3))
    a = qo.lqr_law(g["K_t"], g["K_att"], st, ang, av)
    for n in range(64):
        ref, _ = lqr_action(g["K_t"], g["K_att"], st[n], ang[n], av[n], None)
        assert np.allclose(a[n], ref, rtol=1e-13, atol=1e-13)


def test_robust_control_restatement_vs_reference_vectors():
    """robust_control (quadrotor_env.py:84-109; dead code upstream, quad.robust_control hard-wired False): the oracle's perturbed
    f2F / f2w / drone_eq against the reference's own guarded branches executed with injected perturbations, and the linear
    wind ramp against the reference's stateful wind() over two episodes (fixture: oracle/gen_golden.py:gen_robust_vectors)."""
    g = load_golden("robust_vectors.npz")
    rb = dict(wind=g["wind"], ir=g["ir"], m=g["m"], J=g["J"])
    w, F, Ma = qo.f2F(g["a"], g["kf"])
    assert rel_err(qo.drone_eq(g["x"], F, Ma, w, rb=rb), g["dx_direct"]) < 1e-12
    se, w, F, Ma = qo.f2w(g["fm"][:, 0], g["fm"][:, 1:4], True, g["kf"])
    assert rel_err(se, g["effort"]) < 1e-12 and rel_err(w, g["w_rotor"]) < 1e-12
    assert rel_err(np.c_[F, Ma], g["fm_new"]) < 1e-12
    assert rel_err(qo.drone_eq(g["x"], F, Ma, w, rb=rb), g["dx_indirect"]) < 1e-12
    # the perturbations matter: the unperturbed RHS differs
    w0, F0, Ma0 = qo.f2F(g["a"])
    assert rel_err(qo.drone_eq(g["x"], F0, Ma0, w0), g["dx_direct"]) > 1e-2
    # wind(): between two gusts np.linspace(last, gust, P)[(i % P) - 1]; a new gust whenever i % P == 1
    i, gu, P = g["wind_i"], g["wind_gust"], int(g["gust_period"])
    cur, prev, lasts, n_new = np.zeros(3), np.zeros(3), [], 0
    for k in range(len(i)):
        if not np.array_equal(gu[k], cur):
            prev, cur = cur, gu[k]
            n_new += 1
            assert i[k] % P == 1
        lasts.append(prev.copy())
    assert n_new == 5
    assert np.abs(qo.wind_ramp(np.array(lasts), gu, i, P) - g["wind_out"]).max() < 1e-12


def test_robust_oracle_streams_and_gust_counter():
    """Philox-derived perturbations: ranges / moments of robust_control.reset (:98-102), gust counter advancing once per episode
    start and once per period, perturbations constant inside an episode."""
    N = 4096
    ids = np.arange(N) + 7
    e = qo.robust_episode(5, ids, np.zeros(N, dtype=np.int64))
    assert 0 <= e["kf"].min() and e["kf"].max() <= 0.1 and abs(e["kf"].mean() - 0.05) < 2e-3
    assert 0 <= e["ir"].min() and e["ir"].max() <= 0.1
    assert abs(e["m"].std() - 0.3) < 0.02 and abs(e["J"].std() - 0.1) < 0.01
    gst = qo.robust_gust(5, ids, np.full(N, 3))
    assert np.all(np.abs(gst.std(axis=0) - np.array([5, 5, 2])) < [0.3, 0.3, 0.15])
    assert np.all(qo.robust_gust(5, ids, np.zeros(N, dtype=np.int64)) == 0)
    n = 16
    o = qo.BatchQuadOracle(n, 0.01, 10 ** 6, training=False, T=1, integrator="rk4", robust=dict(seed=5, env_id=np.arange(n)))
    st = np.zeros((n, 13)); st[:, 6] = 1
    o.reset(st)
    assert np.all(o.gust_count == 1)
    kf0 = o.rb["kf"].copy()
    for t in range(505):
        o.step(np.zeros((n, 4)))
    assert np.all(o.gust_count == 2) and np.array_equal(o.rb["kf"], kf0)


def test_mission_generator_is_bit_identical_to_the_reference():
    """mission.py (whole-array NumPy) against mission_control/mission_control.py (Python loops) run as it is
    (oracle/gen_golden.py:gen_mission_vectors): trajectories, velocities and the get_error stream incl. the extrapolation past
    the end, identical to the last bit."""
    from conftest import MISSION_CASES
    from autonomous_quadrotor_environment_b200.mission import mission
    g = load_golden("mission_vectors.npz")
    for name, f in MISSION_CASES.items():
        m = mission(0.01)
        f(m)
        assert np.array_equal(m.trajectory, g[name + "_trajectory"]), name
        assert np.array_equal(m.velocity, g[name + "_velocity"]), name
        err = np.array([m.get_error(0) for _ in range(m.trajectory_total_steps + 3)])
        assert np.array_equal(err, g[name + "_errors"]), name
    with pytest.raises(ValueError):
        mission(0.01).gen_trajectory(300, 100, np.zeros(3), velocity=np.ones(3))


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the arm the driver runs beside ours): one JSON line on stdout with the contract's keys, produced by
    the reference's own quad.step on the host cores (bytecode build oracle/_ref, built here from /root/reference) or — where that
    build is absent — by the C port, and saying which."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1",
                          "--ref-seconds", "4"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("env-steps/sec") and d["unit"] == "env-steps/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "workload" in d["config"]
    if os.path.isdir("/root/reference"):
        assert d["cpu_baseline"]["kind"] == "reference"
