"""GPU parity tests (run on the B200 box): the CUDA path, called through the C ABI, against
  (1) fixtures generated from the unmodified reference (tests/golden, oracle/gen_golden.py),
  (2) the reference's own shipped trajectory log,
  (3) the NumPy oracle on the same seeded inputs (BASELINE.json configs[1]: 4,096 envs, FP64),
  (4) size-independent properties at BASELINE.json's full size (1,048,576 envs).
Tolerances (BASELINE.json north_star): FP64 mode 1e-9 relative per step; FP32 production mode
1e-4 relative + 1e-5 absolute; done flags / reset indices bit-exact."""
import ctypes as C

import numpy as np
import pytest

from conftest import load_golden, rel_err, bound_err, lqr_action

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from autonomous_quadrotor_environment_b200 import BatchedQuad, _lib as L
    from autonomous_quadrotor_environment_b200 import quaternion_euler_utility as U
from oracle import quad_oracle as qo

DEV = "cuda:0"


def T64(x):
    return torch.as_tensor(np.asarray(x, dtype=np.float64), device=DEV)


def T32(x):
    return torch.as_tensor(np.asarray(x, dtype=np.float32), device=DEV)


def npy(t):
    return t.detach().double().cpu().numpy()


# ----------------------------------------------------------------------------------------------------
# device functions vs reference vectors (A3, A4, A5, A8)
# ----------------------------------------------------------------------------------------------------
def test_utility_functions_vs_reference_vectors():
    g = load_golden("utility_vectors.npz")
    assert rel_err(npy(U.euler_quat_batch(T64(g["ang"]))), g["euler_quat"]) < 1e-13
    assert rel_err(npy(U.quat_euler_batch(T64(g["qn"]))), g["quat_euler"]) < 1e-12
    assert rel_err(npy(U.deriv_quat_batch(T64(g["w"]), T64(g["qn"]))), g["deriv_quat"]) < 1e-13
    assert rel_err(npy(U.quat_rot_mat_batch(T64(g["qn"]))), g["quat_rot_mat"]) < 1e-13
    # FP32 production precision
    assert bound_err(npy(U.euler_quat_batch(T32(g["ang"]))), g["euler_quat"]) < 1
    assert bound_err(npy(U.quat_euler_batch(T32(g["qn"]))), g["quat_euler"]) < 1
    assert bound_err(npy(U.quat_rot_mat_batch(T32(g["qn"]))), g["quat_rot_mat"]) < 1
    # the reference-shaped single-quaternion functions
    q = U.euler_quat(g["ang"][0])
    assert q.shape == (4, 1) and rel_err(q.flatten(), g["euler_quat"][0]) < 1e-13
    assert rel_err(U.quat_euler(q), qo.quat_euler(q.T)[0]) < 1e-12
    assert U.quat_rot_mat(q).shape == (3, 3) and U.deriv_quat(g["w"][0], q).shape == (4,)


def _drone_eq(prec, x, action, direct, w=None):
    lib = L.load_library()
    mk = T64 if prec == L.QS_F64 else T32
    xs, a = mk(x.T).contiguous(), mk(action.T).contiguous()
    ws = None if w is None else mk(w.T).contiguous()
    out = torch.empty_like(xs)
    L.check(lib.qs_drone_eq(prec, None, x.shape[0], direct, xs.data_ptr(), a.data_ptr(),
                            None if ws is None else ws.data_ptr(), out.data_ptr(), None))
    return npy(out).T


def test_drone_eq_and_mixer_vs_reference_vectors():
    g = load_golden("drone_eq_vectors.npz")
    assert rel_err(_drone_eq(L.QS_F64, g["x"], g["a"], 1), g["dx_direct"]) < 1e-10
    assert rel_err(_drone_eq(L.QS_F64, g["x"], g["fm_new_clipped"], 0, g["w_clipped"]), g["dx_indirect"]) < 1e-10
    d32 = _drone_eq(L.QS_F32, g["x"], g["a"], 1)
    assert np.max(np.abs(d32 - g["dx_direct"]) / (1e-3 + 1e-4 * np.abs(g["dx_direct"]))) < 1
    lib = L.load_library()
    n = g["fm"].shape[0]
    for clipped, sfx in ((1, "clipped"), (0, "unclipped")):
        fm = T64(g["fm"].T).contiguous()
        eff, w, fmn = torch.empty_like(fm), torch.empty_like(fm), torch.empty_like(fm)
        L.check(lib.qs_f2w(L.QS_F64, None, n, clipped, fm.data_ptr(), eff.data_ptr(), w.data_ptr(), fmn.data_ptr(), None))
        assert rel_err(npy(eff).T, g["effort_" + sfx]) < 1e-11
        assert rel_err(npy(w).T, g["w_" + sfx]) < 1e-11
        assert rel_err(npy(fmn).T, g["fm_new_" + sfx]) < 1e-11


def test_philox_bit_exact():
    lib = L.load_library()
    n = 1000
    out = torch.empty(4, n, dtype=torch.int32, device=DEV)
    for seed, env0, ep, blk, sid in [(0, 0, 0, 0, 0), (0xDEADBEEFCAFEF00D, 123456, 7, 3, 1), (1, 2 ** 24 - 500, 99, 2, 2)]:
        L.check(lib.qs_philox_raw(seed, env0, n, ep, blk, sid, out.data_ptr(), None))
        got = out.cpu().numpy().view(np.uint32).T
        exp = qo.philox_block(seed, np.arange(n) + env0, ep, blk, sid)
        assert np.array_equal(got, exp)


# ----------------------------------------------------------------------------------------------------
# quad.step / quad.reset: FP64 parity mode vs trajectories recorded from the reference
# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,direct,training", [("step_direct.npz", 1, True), ("step_indirect.npz", 0, True),
                                                   ("step_eval.npz", 1, False)])
def test_f64_rk45_reproduces_reference_trajectories(name, direct, training):
    g = load_golden(name)
    n_env = g["init"].shape[0]
    env = BatchedQuad(n_env, 0.01, int(g["n"]), training=training, direct_control=direct, T=int(g["T"]),
                      precision="f64", integrator="rk45", aux=True, device=DEV)
    oh, ah = env.reset(T64(g["init"]))
    assert rel_err(npy(oh), g["reset_obs"]) < 1e-9
    worst = 0.0
    for t in range(g["actions"].shape[0]):
        obs, rew, done = env.step(T64(g["actions"][t]))
        ok = ~np.isnan(g["obs"][t]).any(axis=1)
        worst = max(worst, rel_err(npy(obs)[ok], g["obs"][t][ok]), rel_err(npy(rew)[ok], g["reward"][t][ok]),
                    rel_err(npy(env.state)[ok], g["state"][t][ok]), rel_err(npy(env.ang)[ok], g["ang"][t][ok]),
                    rel_err(npy(env.ang_vel)[ok], g["ang_vel"][t][ok], floor=1.0),
                    rel_err(npy(env.step_effort)[ok], g["step_effort"][t][ok]), rel_err(npy(env.w)[ok], g["w"][t][ok]),
                    rel_err(npy(env.accel)[ok], g["accel"][t][ok]), rel_err(npy(env.abs_sum)[ok], g["abs_sum"][t][ok]),
                    rel_err(npy(env.accelerometer_read)[ok], g["acc_read"][t][ok]),
                    rel_err(npy(env.mat_rot)[ok], g["mat_rot"][t][ok]),
                    rel_err(npy(env._field(L.QS_FIELD_CLIPPED_ACTION).t())[ok], g["clipped_action"][t][ok]))
        assert np.array_equal(done.cpu().numpy().astype(bool), g["done"][t]), "done differs at step %d" % t
        assert np.array_equal(env.solved.cpu().numpy().astype(np.int64), g["solved"][t]), "solved differs at %d" % t
    assert worst < 1e-9, worst


def test_config2_4096_envs_f64_vs_oracle():
    """BASELINE.json configs[1]: 4,096 envs, FP64 mode, random actions, trajectory equivalence."""
    N, steps = 4096, 120
    init, _ = qo.sample_reset_state(1234, np.arange(N), 0)
    rng = np.random.default_rng(5678)
    env = BatchedQuad(N, 0.01, 1000, training=True, direct_control=1, T=1, precision="f64", integrator="rk45", device=DEV)
    ora = qo.BatchQuadOracle(N, 0.01, 1000, training=True, direct_control=1, T=1, integrator="rk45")
    env.reset(T64(init)); ora.reset(init)
    worst, n_done = 0.0, 0
    for t in range(steps):
        a = rng.uniform(-1, 1, (N, 4))
        obs, rew, done = env.step(T64(a))
        o_ref, r_ref, d_ref = ora.step(a)
        ok = ~np.isnan(o_ref).any(axis=1) & (np.abs(o_ref).max(axis=1) < 1e6)
        worst = max(worst, rel_err(npy(obs)[ok], o_ref[ok]), rel_err(npy(rew)[ok], r_ref[ok]))
        assert np.array_equal(done.cpu().numpy().astype(bool), d_ref), "done flags differ at step %d" % t
        n_done = int(d_ref.sum())
    assert n_done > N // 2          # random actions end most episodes well inside the horizon
    assert worst < 1e-9, worst


def test_f64_auto_reset_matches_oracle():
    """In-kernel predicated reset sub-pass: reset indices bit-exact, post-reset trajectories within 1e-9."""
    N, steps, seed, off = 512, 150, 42, 1000
    env = BatchedQuad(N, 0.01, 60, training=True, direct_control=1, T=5, precision="f64", integrator="rk45",
                      auto_reset=True, seed=seed, env_id_offset=off, device=DEV)
    ora = qo.BatchQuadOracle(N, 0.01, 60, training=True, direct_control=1, T=5, integrator="rk45")
    init, _ = qo.sample_reset_state(seed, np.arange(N) + off, 0)
    env.reset(T64(init)); ora.reset(init)
    rng = np.random.default_rng(9)
    worst, resets = 0.0, 0
    for t in range(steps):
        a = rng.uniform(-1, 1, (N, 4)) * 0.6
        obs, rew, done = env.step(T64(a))
        o_ref, r_ref, d_ref = ora.step_autoreset(a, seed, off)
        assert np.array_equal(done.cpu().numpy().astype(bool), d_ref), "reset indices differ at step %d" % t
        worst = max(worst, rel_err(npy(obs), o_ref), rel_err(npy(rew), r_ref))
        resets += int(d_ref.sum())
    assert resets > N
    assert worst < 1e-9, worst
    assert np.array_equal(env.episode.cpu().numpy(), ora.episode)
    s = env.stats()
    for k in ("n_episodes", "n_solved", "n_broken", "n_timeout"):
        assert s[k] == ora.stats[k], k
    assert abs(s["sum_length"] - ora.stats["sum_length"]) < 0.5
    assert abs(s["sum_return"] - ora.stats["sum_return"]) < 1e-3 * max(1, abs(ora.stats["sum_return"]))
    assert s["n_steps"] == N * steps


def test_f64_async_reset_matches_oracle():
    """QS_FLAG_ASYNC_RESET: warm-up steps run as ordinary lock-step steps; done / warm-up flags bit-exact."""
    N, steps, seed, off = 512, 160, 43, 77
    env = BatchedQuad(N, 0.01, 60, training=True, direct_control=1, T=5, precision="f64", integrator="rk45",
                      async_reset=True, seed=seed, env_id_offset=off, device=DEV)
    ora = qo.BatchQuadOracle(N, 0.01, 60, training=True, direct_control=1, T=5, integrator="rk45")
    init, _ = qo.sample_reset_state(seed, np.arange(N) + off, 0)
    env.reset(T64(init)); ora.reset(init)
    rng = np.random.default_rng(10)
    worst, resets, warms = 0.0, 0, 0
    for t in range(steps):
        a = rng.uniform(-1, 1, (N, 4)) * 0.6
        obs, rew, done = env.step(T64(a))
        o_ref, r_ref, d_ref, w_ref = ora.step_async(a, seed, off)
        assert np.array_equal(done.cpu().numpy().astype(bool), d_ref), "done differs at step %d" % t
        assert np.array_equal(env.warmup.cpu().numpy().astype(bool), w_ref), "warm-up mask differs at step %d" % t
        worst = max(worst, rel_err(npy(obs), o_ref), rel_err(npy(rew), r_ref))
        resets += int(d_ref.sum()); warms += int(w_ref.sum())
    assert resets > N and warms >= 5 * (resets - N)
    assert worst < 1e-9, worst
    assert np.array_equal(env.episode.cpu().numpy(), ora.episode)
    s = env.stats()
    for k in ("n_episodes", "n_solved", "n_broken", "n_timeout"):
        assert s[k] == ora.stats[k], k
    assert abs(s["sum_return"] - ora.stats["sum_return"]) < 1e-3 * max(1, abs(ora.stats["sum_return"]))


def test_async_rollout_equals_async_steps():
    N, K, seed = 3000, 40, 6
    mk = lambda: BatchedQuad(N, 0.01, 30, T=3, precision="f32", async_reset=True, seed=seed, device=DEV)
    a, b = mk(), mk()
    a.reset(); b.reset()
    acts = (torch.rand(K, 4, N, device=DEV) * 2 - 1)
    rec = a.rollout(K, actions=acts, record_obs=True, record_reward=True, record_done=True)
    for t in range(K):
        obs, rew, done = b.step_soa(acts[t].contiguous())
        assert torch.equal(b.done_flags, rec["done"][t])
        assert torch.allclose(obs.t(), rec["obs"][t], rtol=1e-6, atol=1e-6)
        assert torch.allclose(rew, rec["reward"][t], rtol=1e-5, atol=1e-5)
    assert torch.equal(a.episode, b.episode) and int(a.episode.max()) > 1


@pytest.mark.parametrize("sensor", [False, True])
def test_rollout_mass_timeout_is_resampled_by_the_whole_warp(sensor):
    """Fused rollout with a 6-step time limit: every env that is still flying times out in the same step, so a warp of the pair
    kernel has up to 64 finishing envs at once — the warp-wide re-sampler (four lanes per env, eight envs per pass) needs several
    passes — while in between only stragglers finish.  Episode counters, done bytes and the re-sampled states must be those of
    repeated qs_step launches (queue + compacted re-sampling), through both action sources' shared path."""
    N, K, seed = 4160 + 62, 40, 23                      # ragged: the last warp is partly padding
    mk = lambda: BatchedQuad(N, 0.01, 6, T=2, precision="f32", async_reset=True, sensor_noise=sensor, seed=seed, device=DEV)
    a, b = mk(), mk()
    a.reset(); b.reset()
    ep0 = int(a.episode.max())
    assert int(a.episode.min()) == ep0
    g = torch.Generator(device=DEV); g.manual_seed(2)
    acts = (torch.rand(K, 4, N, device=DEV, generator=g) * 0.2 - 0.1).contiguous()     # gentle: nearly every env reaches the limit
    rec = a.rollout(K, actions=acts, record_obs=True, record_reward=True, record_done=True)
    most = 0
    for t in range(K):
        obs, rew, done = b.step_soa(acts[t].contiguous())
        assert torch.equal(b.done_flags, rec["done"][t]), t
        assert torch.allclose(obs.t(), rec["obs"][t], rtol=1e-6, atol=1e-6), t
        assert torch.allclose(rew, rec["reward"][t], rtol=1e-5, atol=1e-5), t
        most = max(most, int((rec["done"][t] & 1).sum()))
    assert most > 0.9 * N                                                             # the mass time-out happened
    assert torch.equal(a.episode, b.episode) and int(a.episode.min()) >= ep0 + 4
    assert torch.allclose(a.state, b.state, rtol=1e-6, atol=1e-6)
    st, _ = qo.sample_reset_state(seed, np.arange(N), ep0 + 1)                              # the first mass re-sampling against the oracle's sampler:
    first = min(t for t in range(K) if int((rec["done"][t] & 1).sum()) > 0.9 * N)
    fin = ((rec["done"][first] & 1) != 0).cpu().numpy()
    earlier = (rec["done"][:first] & 1).sum(dim=0).cpu().numpy()
    once = fin & (earlier == 0)                              # envs whose first episode ends here: the observation returned with done
    got = npy(rec["obs"][first].t())[once][:, :10]           # is the new episode's initial observation
    assert once.sum() > 0.8 * N and np.max(np.abs(got - st[once][:, :10]) / (1 + np.abs(st[once][:, :10]))) < 2e-5


def test_c_abi_error_codes_of_the_rollout_entry_point():
    """include/quadsim.h: every entry point returns 0 or a negative QS_E* code and leaves a message for qs_last_error(); nothing is
    launched on a refused call (the handle's state is untouched)."""
    N = 256
    env = BatchedQuad(N, 0.01, 100, T=2, precision="f32", async_reset=True, seed=1, device=DEV)
    env.reset()
    before = env._ws.clone()
    lib, st = env.lib, C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def rc_of(**kw):
        a = L.qs_rollout_args()
        a.horizon = 4
        a.action_source = L.QS_ACT_PHILOX_UNIFORM
        for k, v in kw.items():
            setattr(a, k, v)
        rc = lib.qs_rollout(env._h, C.byref(a), st)
        return rc, lib.qs_last_error().decode()

    for kw, code in (({"horizon": 0}, L.QS_EINVAL), ({"action_source": L.QS_ACT_BUFFER}, L.QS_EINVAL), ({"action_source": 77}, L.QS_EINVAL),
                     ({"sensed_obs_out": before.data_ptr()}, L.QS_ESTATE)):
        rc, msg = rc_of(**kw)
        assert rc == code and "qs_rollout" in msg, (kw, rc, msg)
    assert lib.qs_rollout(None, None, st) == L.QS_EINVAL
    torch.cuda.synchronize()
    assert torch.equal(env._ws, before)                              # refused calls launched nothing
    aux = BatchedQuad(N, 0.01, 100, T=1, precision="f64", integrator="rk45", aux=True, seed=1, device=DEV)
    aux.reset()
    with pytest.raises(L.QuadSimError) as ei:
        aux.rollout(4)                                               # AUX rows are a single-step feature
    assert ei.value.code == L.QS_ESTATE
    rc, _ = rc_of()                                                  # and a well-formed call goes through
    assert rc == L.QS_OK
    torch.cuda.synchronize()
    assert not torch.equal(env._ws, before)


def test_reset_queue_overflow_falls_back_in_lane():
    """Every env of a never-reset handle is done (quad.__init__ :154), so the first step finishes 2M episodes at
    once: far more than the per-block reset queue holds.  All of them must still be re-sampled correctly."""
    N, seed = 1 << 21, 17
    env = BatchedQuad(N, 0.01, 1000, T=3, precision="f32", async_reset=True, seed=seed, device=DEV)
    obs, rew, done = env.step_soa(torch.zeros(4, N, device=DEV))
    assert bool(done.all())
    assert torch.equal(env.episode, torch.ones_like(env.episode)) and int(env.i.abs().max()) == 0
    assert torch.equal(env.env_flags, torch.full_like(env.env_flags, 3 << 3))
    idx = np.concatenate([np.arange(0, 4096), np.arange(N - 4096, N), np.arange(1_000_000, 1_004_096)])
    st, _ = qo.sample_reset_state(seed, idx, 1)
    got = npy(env.state)[idx]
    assert np.max(np.abs(got - st) / (1 + np.abs(st))) < 2e-5
    obs2, _, done2 = env.step_soa(torch.zeros(4, N, device=DEV))
    assert int(done2.sum()) == 0 and bool(env.warmup.all())


def test_square_action_tensor_needs_a_named_layout():
    """4 envs x 4 action channels: the shape cannot tell (N,C) from (C,N) — step() refuses to guess; both named layouts agree."""
    mk = lambda: BatchedQuad(4, 0.01, 100, training=True, direct_control=1, T=1, precision="f64", seed=3, device=DEV)
    a, b = mk(), mk()
    st0, _ = qo.sample_reset_state(5, np.arange(4), 0)
    a.reset(torch.as_tensor(st0, device=DEV)); b.reset(torch.as_tensor(st0, device=DEV))
    act = torch.rand(4, 4, dtype=torch.float64, device=DEV) * 2 - 1            # (N,4)
    with pytest.raises(ValueError):
        a.step(act)
    oa, _, _ = a.step(act, layout="nc")
    ob, _, _ = b.step(act.t().contiguous(), layout="cn")
    assert torch.equal(oa, ob)
    ora = qo.BatchQuadOracle(4, 0.01, 100, training=True, direct_control=1, T=1, integrator="rk45")
    ora.reset(st0)
    o_ref, _, _ = ora.step(act.cpu().numpy())
    assert np.abs(oa.cpu().numpy() - o_ref).max() < 1e-9


def test_random_reset_on_device_matches_oracle_sampler():
    N, seed = 2048, 77
    for prec, tol in (("f64", 1e-12), ("f32", 2e-5)):
        env = BatchedQuad(N, 0.01, 1000, T=1, precision=prec, integrator="rk4", seed=seed, env_id_offset=5, device=DEV)
        env.reset()                                   # random branch, episode counter 0 -> 1
        st, _ = qo.sample_reset_state(seed, np.arange(N) + 5, 1)
        ora = qo.BatchQuadOracle(N, 0.01, 1000, T=1, integrator="rk4")
        ora.reset(st)
        assert np.max(np.abs(npy(env.state) - ora.state) / (1 + np.abs(ora.state))) < tol


# ----------------------------------------------------------------------------------------------------
# FP32 production mode (fixed-step RK4)
# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("substeps", [1, 2])
def test_f32_rk4_teacher_forced_vs_reference(substeps):
    """Teacher-forced protocol: restart every step from the reference state, same action;
    bound 1e-5 + 1e-4*|x_ref| on the 14-float observation, reward and Euler angles."""
    g = load_golden("step_direct.npz")
    n_env = g["init"].shape[0]
    env = BatchedQuad(n_env, 0.01, int(g["n"]), training=True, direct_control=1, T=1, precision="f32",
                      integrator="rk4", substeps=substeps, device=DEV)
    env.reset(T32(g["init"]))
    worst, flips = 0.0, 0
    steps = g["actions"].shape[0]
    for t in range(steps):
        prev = g["reset_state"] if t == 0 else g["state"][t - 1]
        ok = ~np.isnan(g["obs"][t]).any(axis=1) & ~np.isnan(prev).any(axis=1) & (np.abs(prev).max(axis=1) < 1e3)
        env.set_state(T32(np.nan_to_num(prev)))
        obs, rew, done = env.step(T32(g["actions"][t]))
        worst = max(worst, bound_err(npy(obs)[ok], g["obs"][t][ok]), bound_err(npy(env.ang)[ok], g["ang"][t][ok]))
        first = ok & ~(g["done"][t - 1] if t else np.zeros(n_env, bool))
        flips += int((done.cpu().numpy().astype(bool)[first] != g["done"][t][first]).sum())
    assert worst < 1.0, worst
    assert flips == 0


def _actor_forward_np(W, x):
    h = np.tanh(x @ W["actor_0_weight"].T + W["actor_0_bias"])
    h = np.tanh(h @ W["actor_2_weight"].T + W["actor_2_bias"])
    return np.tanh(h @ W["actor_4_weight"].T + W["actor_4_bias"])


def test_f32_closed_loop_1000_steps_with_trained_actor():
    """Closed-loop protocol: the reference's trained N=128 actor drives both the FP64 oracle and the FP32 CUDA
    path for 1000 steps; all non-position states stay within the FP32 bound (positions are not fed back by a
    velocity controller and may drift: checked at 20x the bound)."""
    W = load_golden("actor_128.npz")
    N, steps, T = 16, 1000, 5
    init, _ = qo.sample_reset_state(31, np.arange(N), 0)
    env = BatchedQuad(N, 0.01, 5000, training=False, direct_control=1, T=T, precision="f32", integrator="rk4", device=DEV)
    ora = qo.BatchQuadOracle(N, 0.01, 5000, training=False, direct_control=1, T=T, integrator="rk45")
    oh_g, ah_g = env.reset(T32(init))
    oh_o, ah_o = ora.reset(init)

    def push(hist, obs, act):        # dl_in_gen.dl_input (environment/controller/dl_auxiliary.py:25-32)
        s = np.concatenate([act, obs[:, 1:6:2], obs[:, 6:14]], axis=1)
        return np.concatenate([hist[:, 15:], s], axis=1)

    hg, ho = np.zeros((N, 75)), np.zeros((N, 75))
    oh_g, ah_g = npy(oh_g), npy(ah_g)
    for k in range(T):
        hg = push(hg, oh_g[k], ah_g[k]); ho = push(ho, oh_o[k], ah_o[k])
    worst_np, worst_p = 0.0, 0.0
    nonpos = [1, 3, 5, 6, 7, 8, 9, 10, 11, 12, 13]
    for t in range(steps):
        ag, ao = _actor_forward_np(W, hg), _actor_forward_np(W, ho)
        obs, _, _ = env.step(T32(ag))
        o_ref, _, _ = ora.step(ao)
        og = npy(obs)
        hg = push(hg, og, ag); ho = push(ho, o_ref, ao)
        worst_np = max(worst_np, bound_err(og[:, nonpos], o_ref[:, nonpos]))
        worst_p = max(worst_p, bound_err(og[:, [0, 2, 4]], o_ref[:, [0, 2, 4]]))
    assert worst_np < 1.0, worst_np
    assert worst_p < 20.0, worst_p


# ----------------------------------------------------------------------------------------------------
# sensor model (A13): same Philox -> normal mapping on both sides, so the comparison is deterministic
# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("gps_blend", [0.0, 30.0])
def test_sensor_model_matches_oracle_f64(gps_blend):
    """gps_blend > 0: the complementary GPS blend of the landing stack (visual_landing/math_trajectory.py:71-77)."""
    N, steps, seed, off = 256, 60, 5, 9
    env = BatchedQuad(N, 0.01, 1000, training=False, direct_control=1, T=2, precision="f64", integrator="rk45",
                      sensor_noise=True, seed=seed, env_id_offset=off, device=DEV, params={"gps_blend": gps_blend})
    ora = qo.BatchQuadOracle(N, 0.01, 1000, training=False, direct_control=1, T=2, integrator="rk45")
    sen = qo.SensorOracle(N, 0.01, gps_blend=gps_blend)
    init = np.zeros((N, 13)); init[:, 6] = 1
    rng = np.random.default_rng(3)
    init[:, 1:6:2] = rng.normal(0, 0.3, (N, 3)); init[:, 10:13] = rng.normal(0, 0.3, (N, 3))
    env.reset(T64(init)); ora.reset(init)
    ids = np.arange(N) + off
    ep = np.zeros(N, dtype=np.int64)
    sen.reset(seed, ids, ep, ora.state)                       # sensor.reset after quad.reset's warm-up steps
    assert rel_err(npy(env.sensed_obs), np.concatenate([ora.state[:, 0:10], ora.V_q], axis=1)) < 1e-9
    worst = 0.0
    for t in range(steps):
        a = rng.uniform(-0.2, 0.2, (N, 4))
        env.step(T64(a))
        ora.step(a)
        z = qo.sensor_normals(seed, ids, ep, ora.i)
        ref = sen.step(z, ora.state, ora.accelerometer_read, ora.mat_rot, ora.f_in / qo.M)
        worst = max(worst, rel_err(npy(env.sensed_obs), ref))
    assert worst < 1e-8, worst
    # the noise is really there: sensed attitude rate differs from the true one at the gyro-noise scale
    d = npy(env.sensed_obs)[:, 10:14] - npy(env.obs)[:, 10:14]
    assert 0.005 < d.std() < 0.05


def test_sensor_model_f32_statistics_and_async_reset():
    N, seed = 1 << 16, 8
    env = BatchedQuad(N, 0.01, 200, T=3, precision="f32", async_reset=True, sensor_noise=True, seed=seed, device=DEV)
    env.reset()
    acts = torch.zeros(4, N, device=DEV)
    gyro_dev = []
    for t in range(120):
        env.step_soa(acts)
        if t % 10 == 9:
            alive = (env.i > env.T + 1)
            so, to = env.sensed_obs[alive], env.obs[alive]
            assert torch.isfinite(so).all()
            gyro_dev.append((so[:, 10:14] - to[:, 10:14]).std().item())
    # 1/2*Omega(w_noise)*q with sigma_gyro = 0.035 rad/s -> ~0.5*0.035*sqrt(3)/2 per component, plus INS attitude drift
    assert 0.008 < np.mean(gyro_dev) < 0.04, gyro_dev
    # warm-up steps pass the true observation through
    w = env.warmup.bool()
    if w.any():
        assert torch.equal(env.sensed_obs[w], env.obs[w])


# ----------------------------------------------------------------------------------------------------
# structure: rollout fusion, sharding, checkpoint, host-buffer entry point
# ----------------------------------------------------------------------------------------------------
# ----------------------------------------------------------------------------------------------------
# the three implementations of qs_step (plain loads, CTA-wide TMA ring, per-warp cp.async pipeline) are the same function
# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("loader", [2, 3])
@pytest.mark.parametrize("N,sensor,direct,ext", [(5000, False, 1, False), (4099, True, 1, True), (1 << 16, True, 1, False),
                                                 (2500, False, 0, True), (31, False, 1, True), (33, True, 1, False),
                                                 (65, False, 1, True), (4160, True, 0, True)])
def test_step_loaders_agree(N, sensor, direct, ext, loader):
    """Same inputs through loader 2 (per-warp pipeline, 16-byte vector loads/stores, opportunistic reset drains) and loader 1
    (CTA-wide ring, scalar stores), teacher-forced (handle b restarts every step from a's workspace): every field must agree
    after every step — floats to FP32 rounding (the two kernels are compiled separately, so FMA contraction may differ by an
    ulp), integer/byte fields exactly except where a threshold comparison sits within that ulp — including ragged sizes
    (N % 32 != 0, N % 4 != 0), the sensor rows, indirect control and caller-provided output arrays."""
    seed, K = 11, 60
    mk = lambda ld: BatchedQuad(N, 0.01, 25, T=3, precision="f32", direct_control=direct, async_reset=True,
                                sensor_noise=sensor, seed=seed, device=DEV).set_step_loader(ld)
    a, b = mk(loader), mk(1)
    a.reset(); b.reset()
    g = torch.Generator(device=DEV); g.manual_seed(5)
    ffields = [L.QS_FIELD_OBS, L.QS_FIELD_ANG, L.QS_FIELD_REWARD, L.QS_FIELD_ABS_SUM, L.QS_FIELD_PREV_SHAPING, L.QS_FIELD_EP_RETURN]
    ifields = [L.QS_FIELD_DONE, L.QS_FIELD_SOLVED, L.QS_FIELD_I, L.QS_FIELD_EPISODE, L.QS_FIELD_FLAGS]
    if sensor:
        ffields += [L.QS_FIELD_SENSED_OBS, L.QS_FIELD_SENSOR_STATE]
    outs = []
    for env in (a, b):
        outs.append((torch.full((14, N), -7.0, device=DEV), torch.full((N,), -7.0, device=DEV),
                     torch.full((N,), 9, dtype=torch.uint8, device=DEV), torch.full((N,), 9, dtype=torch.uint8, device=DEV)))
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    flips = 0
    for t in range(K):
        b._ws.copy_(a._ws)
        if direct:
            act = (torch.rand(4, N, device=DEV, generator=g) * 2 - 1).contiguous()
        else:
            act = torch.stack([torch.rand(N, device=DEV, generator=g) * 20, *(torch.rand(3, N, device=DEV, generator=g) - 0.5)]).contiguous()
        for env, (o, r, d, sv) in zip((a, b), outs):
            if ext:
                L.check(env.lib.qs_step(env._h, C.c_void_p(act.data_ptr()), C.c_void_p(o.data_ptr()), C.c_void_p(r.data_ptr()),
                                        C.c_void_p(d.data_ptr()), C.c_void_p(sv.data_ptr()), st))
            else:
                env.step_soa(act)
        same = torch.ones(N, dtype=torch.bool, device=DEV)
        for f in ifields:
            same &= (a._field(f) == b._field(f)).all(dim=0)
        flips += int((~same).sum())
        for f in ffields:
            fa, fb = a._field(f)[:, same], b._field(f)[:, same]
            assert torch.allclose(fa, fb, rtol=2e-5, atol=2e-5, equal_nan=True), (t, f, float((fa - fb).abs().max()))
        if ext:     # the caller's arrays hold exactly what the handle's own rows hold
            o, r, d, sv = outs[0]
            assert torch.equal(o, a._field(L.QS_FIELD_OBS)) and torch.equal(r, a._field(L.QS_FIELD_REWARD)[0])
            assert torch.equal(d, a._field(L.QS_FIELD_DONE)[0]) and torch.equal(sv, a._field(L.QS_FIELD_SOLVED)[0])
            o, r, d, sv = outs[1]
            assert torch.equal(o, b._field(L.QS_FIELD_OBS)) and torch.equal(d, b._field(L.QS_FIELD_DONE)[0])
    assert flips <= 2 + N * K // 100000, flips
    sa = a.stats()
    assert sa["n_episodes"] > 0 and int(a.episode.max()) >= 2


@pytest.mark.parametrize("gps_blend", [0.0, 30.0])
def test_packed_sensor_model_equals_scalar_sensor_model(gps_blend):
    """The sensor model on the packed FP32 pipe (sensor_pair.cuh: loader 3, two envs per lane, Philox blocks drawn on demand,
    accelerometer reading from the closed form f_b/M - 2G R^T z) against the scalar routine (sensor_device.cuh, loader 1),
    teacher-forced, with and without the complementary GPS blend: sensed observation and the 20 sensor-state rows agree to
    FP32 rounding at every step, through warm-up steps and sensor resets.  Half way both handles are re-seeded with a seed whose
    high word is set: the packed kernel takes its Philox round keys from the view (SimView::rk, precomputed on the host per
    launch), the scalar one derives them from the seed in the kernel — the streams must stay identical, and must change."""
    N, K, seed = 4099, 80, 17
    mk = lambda ld: BatchedQuad(N, 0.01, 30, T=3, precision="f32", async_reset=True, sensor_noise=True, seed=seed, device=DEV,
                                params={"gps_blend": gps_blend}).set_step_loader(ld)
    a, b = mk(3), mk(1)
    a.reset(); b.reset()
    g = torch.Generator(device=DEV); g.manual_seed(6)
    worst = 0.0
    for t in range(K):
        b._ws.copy_(a._ws)
        act = (torch.rand(4, N, device=DEV, generator=g) * 0.6 - 0.3).contiguous()
        if t == K // 2:                                       # same state and action under the old and the new seed
            ck = a.get_checkpoint()
            a.step_soa(act)
            old_noise = a.sensed_obs.clone()
            a.set_checkpoint(ck)
            for env in (a, b):
                env.seed(0x9E3779B97F4A7C15)
        a.step_soa(act); b.step_soa(act)
        if t == K // 2:
            assert float((a.sensed_obs - old_noise).abs().max()) > 1e-4     # the new seed drew different noise
        same = (a._field(L.QS_FIELD_DONE) == b._field(L.QS_FIELD_DONE)).all(dim=0) & (a._field(L.QS_FIELD_FLAGS) == b._field(L.QS_FIELD_FLAGS)).all(dim=0)
        assert int((~same).sum()) <= 1
        for f in (L.QS_FIELD_SENSED_OBS, L.QS_FIELD_SENSOR_STATE):
            fa, fb = a._field(f)[:, same], b._field(f)[:, same]
            assert torch.allclose(fa, fb, rtol=2e-5, atol=2e-5), (t, f, float((fa - fb).abs().max()))
            worst = max(worst, float((fa - fb).abs().max()))
    assert int(a.episode.max()) >= 2                      # resets (and therefore sensor resets) happened
    d = a.sensed_obs[:, 0:6] - a.obs[:, 0:6]
    assert float(d.abs().max()) > 1e-4                    # the INS estimate really differs from the truth


@pytest.mark.parametrize("N,src", [(4096, "philox"), (1002, "buffer"), (4096, "buffer")])
def test_pair_rollout_equals_scalar_rollout(N, src):
    """rollout_pair_kernel (two envs per thread, RK4 on FFMA2; the default of FP32 handles) vs rollout_kernel (one env per
    thread) on the same envs, teacher-forced every K=4 steps (handle b restarts from a's workspace): all fields agree to FP32
    rounding, integer fields exactly except threshold cases within that rounding; statistics and Philox streams identical."""
    K, rounds, seed = 4, 12, 21
    mk = lambda ld: BatchedQuad(N, 0.01, 25, T=3, precision="f32", async_reset=True, seed=seed, device=DEV).set_step_loader(ld)
    a, b = mk(3), mk(2)
    a.reset(); b.reset()
    g = torch.Generator(device=DEV); g.manual_seed(5)
    flips = 0
    for r in range(rounds):
        b._ws.copy_(a._ws)
        acts = (torch.rand(K, 4, N, device=DEV, generator=g) * 2 - 1) if src == "buffer" else None
        ra = a.rollout(K, actions=acts, record_obs=True, record_reward=True, record_done=True, record_actions=True)
        rb = b.rollout(K, actions=acts, record_obs=True, record_reward=True, record_done=True, record_actions=True)
        same = (ra["done"] == rb["done"]).all(dim=0) & (a.episode == b.episode) & (a.i == b.i)
        flips += int((~same).sum())
        assert torch.equal(ra["actions"][:, :, same], rb["actions"][:, :, same])             # same Philox draws / same buffer
        assert torch.allclose(ra["obs"][:, :, same], rb["obs"][:, :, same], rtol=1e-4, atol=1e-4)
        assert torch.allclose(ra["reward"][:, same], rb["reward"][:, same], rtol=1e-3, atol=2e-3)
        assert torch.allclose(a.state[same], b.state[same], rtol=1e-4, atol=1e-4)
    assert flips <= 2 + N * K * rounds // 50000, flips
    assert a.stats()["n_episodes"] > 0


@pytest.mark.parametrize("prec,integ", [("f32", "rk4"), ("f64", "rk45")])
def test_rollout_kernel_equals_repeated_steps(prec, integ):
    N, K, seed = 3000, 24, 5
    dt = torch.float32 if prec == "f32" else torch.float64
    mk = lambda: BatchedQuad(N, 0.01, 40, T=2, precision=prec, integrator=integ, auto_reset=True, seed=seed, device=DEV)
    a, b = mk(), mk()
    a.reset(); b.reset()
    acts = (torch.rand(K, 4, N, device=DEV, dtype=dt) * 2 - 1)
    rec = a.rollout(K, actions=acts, record_obs=True, record_reward=True, record_done=True)
    for t in range(K):
        obs, rew, done = b.step_soa(acts[t].contiguous())
        assert torch.equal(done, rec["done"][t])
        assert torch.allclose(obs.t(), rec["obs"][t], rtol=1e-6 if prec == "f32" else 1e-13, atol=1e-6 if prec == "f32" else 1e-13)
        assert torch.allclose(rew, rec["reward"][t], rtol=1e-5 if prec == "f32" else 1e-12, atol=1e-5 if prec == "f32" else 1e-12)
    assert torch.equal(a.episode, b.episode)
    sa, sb = a.stats(), b.stats()
    assert sa["n_episodes"] == sb["n_episodes"] > 0 and sa["n_steps"] == sb["n_steps"] == N * K


def test_sharding_is_invisible():
    """Philox is keyed by GLOBAL env id, so k shards reproduce the 1-GPU result bit-for-bit."""
    N, K, seed = 4096, 40, 11
    whole = BatchedQuad(N, 0.01, 30, T=1, precision="f32", auto_reset=True, seed=seed, device=DEV)
    parts = [BatchedQuad(N // 2, 0.01, 30, T=1, precision="f32", auto_reset=True, seed=seed, env_id_offset=o, device=DEV)
             for o in (0, N // 2)]
    whole.reset()
    for p in parts:
        p.reset()
    whole.rollout(K)
    for p in parts:
        p.rollout(K)
    assert torch.equal(whole.state, torch.cat([p.state for p in parts], dim=0))
    tot = sum(p.stats()["n_episodes"] for p in parts)
    assert tot == whole.stats()["n_episodes"] > 0


def test_checkpoint_resume():
    N = 1024
    a = BatchedQuad(N, 0.01, 50, T=1, precision="f32", auto_reset=True, seed=3, device=DEV)
    a.reset(); a.rollout(10)
    ck = a.get_checkpoint()
    a.rollout(15)
    b = BatchedQuad(N, 0.01, 50, T=1, precision="f32", auto_reset=True, seed=999, device=DEV)
    b.set_checkpoint(ck); b.rollout(15)
    assert torch.equal(a.state, b.state) and torch.equal(a.episode, b.episode)
    a.seed(12345)                                         # a checkpoint taken after quad.seed carries the key in force
    ck = a.get_checkpoint()
    assert ck["seed"] == 12345
    a.rollout(40)
    b.set_checkpoint(ck); b.rollout(40)
    assert int(a.episode.max()) >= 1                      # resets drew from the re-keyed streams
    assert torch.equal(a.state, b.state) and torch.equal(a.episode, b.episode)


@pytest.mark.parametrize("N,sensor", [(2048, False), (3 * 65536 + 777, False), (1 << 18, True)])
def test_step_host_entry_point(N, sensor):
    """qs_step_host (host buffers in, host buffers out) == qs_step on device tensors, bit for bit — also when the shard is
    large enough for the sliced H2D -> step -> D2H pipeline (>= 2 x 65,536 envs; ragged last slice), over several steps with
    auto-reset so that the slices' Philox streams (global env ids) are exercised too."""
    mk = lambda: BatchedQuad(N, 0.01, 20, T=2, precision="f32", async_reset=True, sensor_noise=sensor, seed=9, device=DEV)
    a, b = mk(), mk()
    a.reset(); b.reset()
    obs_h = torch.empty(14, N).pin_memory(); rew_h = torch.empty(N).pin_memory()
    done_h = torch.empty(N, dtype=torch.uint8).pin_memory()
    g = torch.Generator(); g.manual_seed(3)
    for t in range(30):
        act = torch.rand(4, N, generator=g) * 2 - 1
        act_p = act.pin_memory()
        torch.cuda.synchronize()
        L.check(a.lib.qs_step_host(a._h, act_p.data_ptr(), obs_h.data_ptr(), rew_h.data_ptr(), done_h.data_ptr(), None))
        obs, rew, done = b.step_soa(act.to(DEV))
        assert torch.equal(obs.t().cpu(), obs_h), t
        assert torch.equal(rew.cpu(), rew_h) and torch.equal(b.done_flags.cpu(), done_h), t      # raw byte: bit1 = warm-up step
    assert torch.equal(a.state, b.state) and torch.equal(a.episode, b.episode)
    sa, sb = a.stats(), b.stats()
    assert sa["n_episodes"] == sb["n_episodes"] > 0 and sa["n_steps"] == sb["n_steps"]


# ----------------------------------------------------------------------------------------------------
# A14: actor MLP on tcgen05 tensor cores fused into the rollout (BASELINE.json configs[4])
# ----------------------------------------------------------------------------------------------------
def test_umma_selftest_gemm():
    lib = L.load_library()
    torch.manual_seed(0)
    for N, K in [(16, 16), (128, 80), (128, 128), (16, 128)]:
        A = torch.randn(128, K, device=DEV); B = torch.randn(N, K, device=DEV); D = torch.zeros(128, N, device=DEV)
        L.check(lib.qs_umma_selftest(N, K, A.data_ptr(), B.data_ptr(), D.data_ptr(), None))
        ref = A.bfloat16().float() @ B.bfloat16().float().t()
        assert (D - ref).abs().max().item() < 1e-3


def test_umma_selftest_gemm_a_operand_in_tensor_memory():
    """TS form of tcgen05.mma: the thread that owns row r writes its packed BF16 row into TMEM with tcgen05.st and the MMA reads
    the A operand from there (the data path of hidden activations that never touch shared memory)."""
    lib = L.load_library()
    torch.manual_seed(1)
    for N, K in [(16, 32), (64, 128), (128, 128), (128, 64)]:
        A = torch.randn(128, K, device=DEV); B = torch.randn(N, K, device=DEV); D = torch.zeros(128, N, device=DEV)
        L.check(lib.qs_umma_selftest_ts(N, K, A.data_ptr(), B.data_ptr(), D.data_ptr(), None))
        ref = A.bfloat16().float() @ B.bfloat16().float().t()
        assert (D - ref).abs().max().item() < 1e-3, (N, K, (D - ref).abs().max().item())


def test_umma_selftest_gemm_mn_major_operands():
    """MN-major operands (K = row index of the stored tile): the weight-gradient products of qs_ppo_grad read the forward pass's own
    activation tiles, no transposed copies (mode 0: both operands MN-major; mode 1: A K-major, B MN-major = dZ W)."""
    lib = L.load_library()
    torch.manual_seed(2)
    for N in (16, 80, 128):
        A = torch.randn(128, 128, device=DEV); B = torch.randn(128, N, device=DEV); D = torch.zeros(128, N, device=DEV)
        L.check(lib.qs_umma_selftest_mn(0, N, A.data_ptr(), B.data_ptr(), D.data_ptr(), None))
        ref = A.bfloat16().float().t() @ B.bfloat16().float()
        assert (D - ref).abs().max().item() < 1e-3, (0, N, (D - ref).abs().max().item())
        L.check(lib.qs_umma_selftest_mn(1, N, A.data_ptr(), B.data_ptr(), D.data_ptr(), None))
        ref = A.bfloat16().float() @ B.bfloat16().float()
        assert (D - ref).abs().max().item() < 1e-3, (1, N, (D - ref).abs().max().item())


def _torch_actor(W, x):
    h = torch.tanh(x @ W["actor_0_weight"].t() + W["actor_0_bias"])
    h = torch.tanh(h @ W["actor_2_weight"].t() + W["actor_2_bias"])
    return torch.tanh(h @ W["actor_4_weight"].t() + W["actor_4_bias"])


def test_fused_actor_rollout_deterministic():
    """sigma = 0: (1) every recorded action equals the FP32 torch actor evaluated on the history rebuilt from the
    kernel's own recorded (action, obs) stream, up to BF16-operand error; (2) replaying the recorded actions through
    the plain rollout kernel reproduces the recorded observations bit-for-bit (same dynamics code path)."""
    g = load_golden("actor_128.npz")
    W = {k: torch.as_tensor(v, device=DEV) for k, v in g.items() if k.startswith("actor_")}
    N, K, seed = 1000, 48, 12                                   # N not a multiple of 128: ragged last tile
    mk = lambda: BatchedQuad(N, 0.01, 300, training=False, direct_control=1, T=5, precision="f32", async_reset=True,
                             seed=seed, device=DEV)
    env, ref = mk(), mk()
    oh, ah = env.reset(); ref.reset()
    hist = torch.zeros(N, 75, device=DEV)
    for k in range(5):                                          # dl_in_gen.dl_input over the T warm-up pairs
        s = torch.cat([ah[k], oh[k][:, 1:6:2], oh[k][:, 6:14]], dim=1)
        hist = torch.cat([hist[:, 15:], s], dim=1)
    env.history.copy_(hist)
    env.load_actor(g, action_std=0.0)
    rec = env.policy_rollout(K, record_obs=True)
    worst, n_warm = 0.0, 0
    for t in range(K):
        mean = _torch_actor(W, hist.bfloat16().float())         # the kernel's A operand is the BF16-rounded history
        a_t, o_t = rec["actions"][t].t(), rec["obs"][t].t()
        warm = ((rec["done"][t] >> 1) & 1).bool()
        n_warm += int(warm.sum())
        assert bool((a_t[warm] == 0).all())                     # warm-up steps apply (and record) zero_control
        worst = max(worst, (a_t - mean)[~warm].abs().max().item())
        # what the kernel pushed: the action it applied and the observation it returned
        hist = torch.cat([hist[:, 15:], torch.cat([a_t, o_t[:, 1:6:2], o_t[:, 6:14]], dim=1)], dim=1)
    assert worst < 0.03, worst                                  # BF16 weights/activations, MUFU tanh
    assert n_warm > 0                                           # some envs broke and went through an async reset
    assert torch.allclose(env.history, hist.bfloat16().float(), atol=0, rtol=0)
    ref.set_step_loader(2)                                      # one env per thread: the scalar step_core the policy kernel runs
    replay = ref.rollout(K, actions=rec["actions"].contiguous(), record_obs=True, record_done=True)
    assert torch.equal(replay["done"], rec["done"])
    assert torch.equal(replay["obs"], rec["obs"])
    assert torch.equal(ref.state, env.state)


def test_fused_actor_rollout_on_sensed_observations():
    """QS_FLAG_SENSOR_NOISE handle (the loop of visual_landing/rl_worker.py:164-175): the policy flies on the SENSED
    observation.  sigma = 0: (1) every action equals the FP32 torch actor on the history rebuilt from the recorded (action, sensed obs)
    stream; (2) replaying the recorded actions through the plain sensor rollout reproduces the true AND the sensed observations
    bit-for-bit (same dynamics, same sensor model, same Philox counters); (3) sensed != true on ordinary steps."""
    g = load_golden("actor_128.npz")
    W = {k: torch.as_tensor(v, device=DEV) for k, v in g.items() if k.startswith("actor_")}
    N, K, seed = 700, 40, 21
    mk = lambda: BatchedQuad(N, 0.01, 300, training=False, direct_control=1, T=5, precision="f32", async_reset=True, sensor_noise=True,
                             seed=seed, device=DEV)
    env, ref = mk(), mk()
    oh, ah = env.reset(); ref.reset()
    hist = torch.zeros(N, 75, device=DEV)
    for k in range(5):
        hist = torch.cat([hist[:, 15:], torch.cat([ah[k], oh[k][:, 1:6:2], oh[k][:, 6:14]], dim=1)], dim=1)
    env.history.copy_(hist)
    env.load_actor(g, action_std=0.0)
    rec = env.policy_rollout(K, record_obs=True, record_sensed=True)
    worst = 0.0
    for t in range(K):
        mean = _torch_actor(W, hist.bfloat16().float())
        a_t, s_t = rec["actions"][t].t(), rec["sensed_obs"][t].t()
        warm = ((rec["done"][t] >> 1) & 1).bool()
        worst = max(worst, (a_t - mean)[~warm].abs().max().item())
        hist = torch.cat([hist[:, 15:], torch.cat([a_t, s_t[:, 1:6:2], s_t[:, 6:14]], dim=1)], dim=1)
    assert worst < 0.03, worst
    assert torch.allclose(env.history, hist.bfloat16().float(), atol=0, rtol=0)
    live = ((rec["done"] & 3) == 0).unsqueeze(1).expand_as(rec["obs"])
    diff = (rec["sensed_obs"] - rec["obs"]).abs()
    assert float(diff[live].max()) > 1e-4                       # the sensor model is in the loop ...
    assert float(diff[:, :6][live[:, :6]].max()) < 1.0          # ... and dead-reckons position / velocity close to the truth
    ref.set_step_loader(2)                                      # the scalar step_core + sensor sub-pass the policy kernel runs
    replay = ref.rollout(K, actions=rec["actions"].contiguous(), record_obs=True, record_done=True, record_sensed=True)
    assert torch.equal(replay["done"], rec["done"])
    assert torch.equal(replay["obs"], rec["obs"])
    assert torch.equal(replay["sensed_obs"], rec["sensed_obs"])
    assert torch.equal(ref.state, env.state)
    # a plain handle refuses the sensed record; the critic head composes with the sensor variant
    plain = BatchedQuad(256, 0.01, 300, training=True, direct_control=1, T=5, precision="f32", async_reset=True, seed=1, device=DEV)
    plain.reset(); plain.load_actor(g, action_std=0.1)
    with pytest.raises(RuntimeError):
        plain.policy_rollout(4, record_sensed=True)


def test_policy_rollout_sharding_and_launch_splitting_are_invisible():
    """configs[4] across ranks: the action noise of the fused policy (Philox stream RNG_POLICY) and the resets are keyed by GLOBAL env id,
    episode and step, so two half-size handles with env-id offsets reproduce the whole bit for bit (actions, log-probs, values,
    rewards, state, history); and one 48-step launch == three 16-step launches (the history carries over)."""
    g = load_golden("actor_128.npz")
    crit = {"critic_0_weight": g["actor_0_weight"], "critic_0_bias": g["actor_0_bias"], "critic_2_weight": g["actor_2_weight"],
            "critic_2_bias": g["actor_2_bias"], "critic_4_weight": g["actor_4_weight"][:1], "critic_4_bias": g["actor_4_bias"][:1]}
    N, K, seed = 2048, 48, 9
    mk = lambda n, off: BatchedQuad(n, 0.01, 200, training=True, direct_control=1, T=5, precision="f32", async_reset=True, seed=seed,
                                    env_id_offset=off, device=DEV)
    whole, lo, hi, split = mk(N, 0), mk(N // 2, 0), mk(N // 2, N // 2), mk(N, 0)
    recs = []
    for e in (whole, lo, hi, split):
        e.reset()
        e.load_actor(g, action_std=0.1, critic=crit)
    kw = dict(record_obs=True, record_values=True)
    rw, rl, rh = whole.policy_rollout(K, **kw), lo.policy_rollout(K, **kw), hi.policy_rollout(K, **kw)
    for key in ("actions", "logprob", "reward", "done", "obs"):
        assert torch.equal(rw[key], torch.cat([rl[key], rh[key]], dim=-1)), key
    assert torch.equal(rw["value"], torch.cat([rl["value"], rh["value"]], dim=-1))
    assert torch.equal(whole.state, torch.cat([lo.state, hi.state], dim=0))
    assert torch.equal(whole.history, torch.cat([lo.history, hi.history], dim=0))
    parts = [split.policy_rollout(16, **kw) for _ in range(3)]
    for key in ("actions", "logprob", "reward", "done", "obs"):
        assert torch.equal(rw[key], torch.cat([p[key] for p in parts], dim=0)), key
    assert torch.equal(rw["value"][:K], torch.cat([p["value"][:16] for p in parts], dim=0))
    assert torch.equal(whole.state, split.state) and torch.equal(whole.history, split.history)
    assert int(whole.episode.max()) >= 1                        # some envs were re-sampled on the way


def test_fused_actor_rollout_closed_loop_solves_and_samples():
    g = load_golden("actor_128.npz")
    N = 8192
    env = BatchedQuad(N, 0.01, 1000, training=True, direct_control=1, T=5, precision="f32", async_reset=True, seed=3, device=DEV)
    env.reset()
    env.load_actor(g, action_std=0.1)
    rec = None
    for _ in range(6):
        rec = env.policy_rollout(128)
    s = env.stats()
    # the shipped controller solves ~95 % of the training episodes (training_log/log_128_*.csv); BF16 MLP + noise sigma=0.1
    assert s["n_episodes"] > N // 2 and s["solved_frac"] > 0.85, s
    # log-prob of a fixed-sigma Normal: -z^2/2 - log(sigma) - log(sqrt(2 pi)), z ~ N(0,1)
    lp = rec["logprob"][:, :, :].flatten()
    z2 = -2 * (lp + np.log(0.1) + 0.5 * np.log(2 * np.pi))
    assert abs(z2.mean().item() - 1.0) < 0.02 and z2.min().item() > -1e-4


# ----------------------------------------------------------------------------------------------------
# the reference's own shipped log through the drop-in `quad` class
# ----------------------------------------------------------------------------------------------------
def test_dropin_quad_reproduces_shipped_lqr_log():
    from autonomous_quadrotor_environment_b200.quadrotor_env import quad
    g = load_golden("lqr_log.npz")
    # lqr_quad.py:117-118 — the 2021 logs pre-date the robust-RNG draws
    env = quad(0.01, 500, training=True, euler=0, direct_control=0, T=1, clipped=True, robust_rng_draws=False, verbose=False)
    env.seed(1)
    for ep in range(3):
        state, action = env.reset()
        assert state.shape == (1, 14) and action.shape == (1, 4)
        euler_t_ant = env.ang
        n_steps = 60 if ep == 0 else 500       # episode 0 diverges chaotically (reference re-run: 5e-7); check its head
        worst = 0.0
        for i in range(500):
            action, euler_t = lqr_action(g["K_t"], g["K_att"], env.state, env.ang, env.ang_vel, euler_t_ant)
            euler_t_ant = euler_t
            obs, rew, done = env.step(action)
            assert obs.shape == (1, 14) and isinstance(rew, float) and isinstance(done, bool)
            if i < n_steps:
                row = np.concatenate((env.state[1:6:2], env.ang, env.ang_vel, env.step_effort))
                worst = max(worst, float(np.max(np.abs(row - g["log"][ep, i]))))
        assert worst < 1e-8, (ep, worst)


def test_dropin_quad_survives_pickling_mid_episode():
    """environment/controller/ppo.py ships its environments to a multiprocessing pool (the compatibility class must
    be picklable).  A drop-in `quad` pickled in the middle of an episode and loaded again continues exactly like the original:
    same observations, rewards, done flags and attribute surface, bit for bit (FP64 + RK45 replica), through the episode's end and
    the next reset (the NumPy global stream both draw from is re-seeded alike)."""
    import pickle
    from autonomous_quadrotor_environment_b200.quadrotor_env import quad
    a = quad(0.01, 60, training=True, euler=0, direct_control=1, T=3, clipped=True, verbose=False)
    a.seed(7)
    s0, a0 = a.reset()
    assert s0.shape == (3, 14) and a0.shape == (3, 4)
    rng = np.random.default_rng(3)
    acts = rng.uniform(-0.3, 0.3, (80, 4))
    for k in range(25):
        a.step(acts[k])
    blob = pickle.dumps(a)
    b = pickle.loads(blob)
    assert b is not a and b._sim is not a._sim
    for name in ("state", "ang", "ang_vel", "step_effort", "i", "abs_sum", "done", "solved"):
        assert np.array_equal(np.asarray(getattr(a, name)), np.asarray(getattr(b, name))), name
    fresh_blob = pickle.dumps(quad(0.01, 60, training=True, direct_control=1, T=3, verbose=False))     # never stepped: no device state yet
    assert pickle.loads(fresh_blob)._sim is None
    for k in range(25, 80):
        oa, ra, da = a.step(acts[k])
        ob, rb, db = b.step(acts[k])
        assert np.array_equal(oa, ob) and ra == rb and da == db, k
        if da:
            break
    np.random.seed(11); sa, _ = a.reset()
    np.random.seed(11); sb, _ = b.reset()
    assert np.array_equal(sa, sb) and np.array_equal(a.state, b.state)


# ----------------------------------------------------------------------------------------------------
# the reference's classical comparison controllers as in-kernel control laws
# ----------------------------------------------------------------------------------------------------
def _log_rows(rec):
    """(K,N,13) = [vel(3), ang(3), ang_vel(3), step_effort(4)]: the columns of classical_controller_results/*.npy"""
    obs, aux = rec["obs"].cpu().numpy(), rec["aux"].cpu().numpy()
    return np.concatenate([obs[:, (1, 3, 5)], aux], axis=1).transpose(0, 2, 1)


@pytest.mark.parametrize("kind", ["lqr", "pid"])
def test_control_rollout_reproduces_shipped_controller_logs(kind):
    """FP64 + RK45 replica, one env per logged episode, ONE launch of 500 fused steps with the control law in-kernel vs the
    logs the reference ships (written by the author's 2021 run).  Episodes ran back to back in the scripts, so prev_ang of
    episode k is the last Euler angle of episode k-1 (never cleared by reset, quadrotor_env.py:171-172)."""
    from autonomous_quadrotor_environment_b200 import controllers as ctl
    g = load_golden("%s_log.npz" % kind)
    E = g["log"].shape[0]
    T = 1 if kind == "lqr" else 5                        # lqr_quad.py:114, pid_vel_control.py:132
    env = BatchedQuad(E, 0.01, 500, training=True, direct_control=0, T=T, clipped=True, precision="f64", aux=True, device=DEV)
    env.prev_ang[1:] = T64(g["log"][:E - 1, -1, 3:6])
    env.reset(T64(g["inits"][:E]))
    c = ctl.lqr_controller(K_t=g["K_t"], K_att=g["K_att"]) if kind == "lqr" else ctl.pid_controller()
    rec = env.control_rollout(c, 500, record_obs=True, record_aux=True)
    rows = _log_rows(rec)                                # (500, E, 13)
    for ep in range(E):
        n_steps = 60 if (kind == "lqr" and ep == 0) else 300     # LQR episode 0 diverges chaotically (reference re-run: 5e-7)
        err = np.abs(rows[:n_steps, ep] - g["log"][ep, :n_steps]).max()
        assert err < 1e-8, (kind, ep, err)
    # the AUX attributes of the handle are those of the last step, like the reference object's
    assert np.allclose(env.ang_vel.cpu().numpy(), rows[-1][:, 6:9], rtol=0, atol=1e-12)
    assert np.allclose(env.step_effort.cpu().numpy(), rows[-1][:, 9:13], rtol=0, atol=1e-12)


@pytest.mark.parametrize("kind", ["lqr", "pid"])
def test_control_rollout_f64_vs_oracle_random_states(kind):
    """256 envs from the reset distribution, 120 closed-loop steps: in-kernel law + RK45 replica vs the oracle driven by the
    restated law, every step within 1e-9 (relative to max(|x|, 1e-3)); controller memory round-trips through ctrl_state."""
    from autonomous_quadrotor_environment_b200 import controllers as ctl
    N, K = 256, 120
    init, _ = qo.sample_reset_state(21, np.arange(N), 0)
    env = BatchedQuad(N, 0.01, 500, training=False, direct_control=0, T=2, clipped=True, precision="f64", aux=True, device=DEV)
    ora = qo.BatchQuadOracle(N, 0.01, 500, training=False, direct_control=0, T=2, clipped=True, integrator="rk45")
    env.reset(T64(init)); ora.reset(init)
    K_t, K_att = qo.lqr_gains()
    c = ctl.lqr_controller() if kind == "lqr" else ctl.pid_controller(target_vel=(0.3, -0.2, 0.1), target_psi=0.05)
    pid = qo.PidControllerOracle(N)
    cs = env.controller_state()
    recs = [env.control_rollout(c, K // 2, ctrl_state=cs, record_obs=True, record_aux=True, record_actions=True) for _ in range(2)]
    rows = np.concatenate([_log_rows(r) for r in recs])
    acts = np.concatenate([r["actions"].cpu().numpy() for r in recs]).transpose(0, 2, 1)
    action = np.tile(np.array([9.82 * 1.03, 0, 0, 0]), (N, 1))
    worst = 0.0
    for t in range(K):
        if kind == "lqr":
            action = qo.lqr_law(K_t, K_att, ora.state, ora.ang, ora.ang_vel)
        worst = max(worst, float(rel_err(acts[t], action)))
        ora.step(action)
        if kind == "pid":
            action = pid.control(ora.state, ora.ang, np.array([0.3, -0.2, 0.1]), 0.05)
        ref = np.concatenate([ora.state[:, 1:6:2], ora.ang, ora.ang_vel, ora.step_effort], axis=1)
        worst = max(worst, float(rel_err(rows[t], ref)))
    assert worst < 1e-9, worst


def test_pid_tracks_a_mission_velocity_profile_f64_vs_oracle():
    """SURVEY 8(f)3, set-point generation: the velocity profile of a `mission` (mission_control/mission_control.py; mission.py
    is bit-identical to it) drives the in-kernel velocity PID through target_traj, one set-point per step, in two launches;
    the oracle's PID restatement fed the same rows step by step agrees within 1e-9 (spiral mission), and with a straight-line
    mission the quads end up flying at the commanded velocity."""
    from autonomous_quadrotor_environment_b200 import controllers as ctl
    from autonomous_quadrotor_environment_b200.mission import mission
    N, K = 128, 240
    init, _ = qo.sample_reset_state(33, np.arange(N), 0)
    init[:, 10:13] *= 0.2
    init[:, 1:6:2] *= 0.1
    env = BatchedQuad(N, 0.01, 10 ** 5, training=False, direct_control=0, T=1, clipped=True, precision="f64", aux=True, device=DEV)
    ora = qo.BatchQuadOracle(N, 0.01, 10 ** 5, training=False, direct_control=0, T=1, clipped=True, integrator="rk45")
    env.reset(T64(init)); ora.reset(init)
    ms = mission(0.01)
    ms.spiral_trajectory(150, K, 0.5, 1.5, 1.0, np.zeros(3))
    vel = ms.velocity.copy()
    c = ctl.pid_controller(target_psi=0.0)
    cs = env.controller_state()
    recs = [env.control_rollout(c, K // 2, ctrl_state=cs, record_obs=True, record_aux=True, record_actions=True,
                                target_traj=ms.velocity_setpoints(K // 2, device=DEV, dtype=torch.float64)) for _ in range(2)]
    assert ms.trajectory_step == K
    rows = np.concatenate([_log_rows(r) for r in recs])
    acts = np.concatenate([r["actions"].cpu().numpy() for r in recs]).transpose(0, 2, 1)
    pid = qo.PidControllerOracle(N)
    action = np.tile(np.array([9.82 * 1.03, 0, 0, 0]), (N, 1))
    worst = 0.0
    for t in range(K):
        worst = max(worst, float(rel_err(acts[t], action)))
        ora.step(action)
        action = pid.control(ora.state, ora.ang, vel[t], 0.0)
        ref = np.concatenate([ora.state[:, 1:6:2], ora.ang, ora.ang_vel, ora.step_effort], axis=1)
        worst = max(worst, float(rel_err(rows[t], ref)))
    assert worst < 1e-9, worst
    line = mission(0.01)
    line.gen_trajectory(600, 600, np.array([6.0, -3.0, 1.5]))          # constant velocity (1, -0.5, 0.25) m/s from step 1 on
    rec = env.control_rollout(c, 600, record_obs=True, target_traj=line.velocity_setpoints(600, device=DEV, dtype=torch.float64))
    v_end = rec["obs"][-1][(1, 3, 5), :].t().cpu().numpy()
    v_sp = line.velocity[-1]
    assert np.median(np.linalg.norm(v_end - v_sp[None, :], axis=1)) < 0.25 * np.linalg.norm(v_sp)
    with pytest.raises(L.QuadSimError):
        env.control_rollout(ctl.lqr_controller(), 4, target_traj=np.zeros((4, 3)))


# ----------------------------------------------------------------------------------------------------
# robust_control: per-episode parameter perturbations + wind gusts
# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("direct,integ", [(1, "rk45"), (0, "rk45"), (1, "rk4")])
def test_robust_control_f64_matches_oracle(direct, integ):
    """quad.robust_control = True on the device (QS_FLAG_ROBUST: Philox-derived episode_kf / episode_m / episode_ir / episode_J,
    gust counter, linear wind ramp) against the oracle's restatement, which is itself pinned to the reference's guarded
    branches (tests/golden/robust_vectors.npz): every step within 1e-9, gust counters identical, including a short gust
    period so that several ramps and the index -1 element are crossed, a second deterministic reset and a Philox reset."""
    N, steps, seed, off = 192, 70, 13, 5
    par = dict(qo.ROBUST_DEFAULTS, gust_period=16)
    env = BatchedQuad(N, 0.01, 10 ** 6, training=False, direct_control=direct, T=2, precision="f64", integrator=integ, aux=True,
                      robust_control=True, seed=seed, env_id_offset=off, device=DEV, params={"robust_gust_period": 16})
    ids = np.arange(N) + off
    ora = qo.BatchQuadOracle(N, 0.01, 10 ** 6, training=False, direct_control=direct, T=2, integrator=integ,
                             robust=dict(seed=seed, env_id=ids, par=par))
    rng = np.random.default_rng(4)
    init = np.zeros((N, 13)); init[:, 6] = 1
    init[:, 1:6:2] = rng.normal(0, 0.5, (N, 3)); init[:, 10:13] = rng.normal(0, 0.5, (N, 3))
    worst = 0.0
    for phase in range(2):
        oh, _ = env.reset(T64(init)); ro, _ = ora.reset(init)
        worst = max(worst, float(rel_err(npy(oh), ro)))
        for t in range(steps):
            if direct:
                a = rng.uniform(-0.4, 0.4, (N, 4))
            else:
                a = np.stack([rng.uniform(8, 12, N), *(rng.normal(0, 0.02, (3, N)))], axis=1)
            obs, rew, done = env.step(T64(a))
            o_ref, r_ref, d_ref = ora.step(a)
            worst = max(worst, float(rel_err(npy(obs), o_ref)), float(rel_err(npy(rew), r_ref)))
            worst = max(worst, float(rel_err(npy(env.accel), ora.accel)))
            assert np.array_equal(npy(done).astype(bool), d_ref)
        assert np.array_equal(npy(env.gust_count), ora.gust_count)
    assert worst < 1e-9, worst
    assert int(ora.gust_count.min()) >= 2 * (1 + steps // 16)
    # the perturbed plant really differs from the nominal one
    nom = BatchedQuad(N, 0.01, 10 ** 6, training=False, direct_control=direct, T=2, precision="f64", integrator=integ, device=DEV)
    nom.reset(T64(init))
    env.reset(T64(init))
    a = np.zeros((N, 4)) if direct else np.tile([10.0, 0, 0, 0], (N, 1))
    for t in range(20):
        o_n, _, _ = nom.step(T64(a)); o_r, _, _ = env.step(T64(a))
    assert float((o_n - o_r).abs().max()) > 1e-2
    # Philox reset: the episode counter advances, so the perturbations are re-drawn
    env.reset()
    ora.episode += 1
    e1 = qo.robust_episode(seed, ids, ora.episode, par)
    assert not np.allclose(e1["kf"], ora.rb["kf"])


def test_robust_control_f32_async_reset_and_controller_rollout():
    """FP32 production arithmetic with robust_control: (1) lock-step qs_step with asynchronous resets stays finite, counts
    gusts per env (one at every episode start, one per period) and ends episodes; (2) the batched LQR law against the
    perturbed plant (qs_control_rollout) still stabilises most envs but is measurably worse than on the nominal plant;
    (3) the fused rollouts that do not model it refuse the handle."""
    from autonomous_quadrotor_environment_b200 import controllers as ctl
    N = 1 << 14
    env = BatchedQuad(N, 0.01, 300, T=3, precision="f32", async_reset=True, robust_control=True, seed=2, device=DEV)
    env.reset()
    assert env.step_loader in (0, 1)
    g = torch.Generator(device=DEV); g.manual_seed(1)
    for t in range(150):
        env.step_soa((torch.rand(4, N, device=DEV, generator=g) * 2 - 1).contiguous())
    assert torch.isfinite(env.obs).all()
    s = env.stats()
    assert s["n_episodes"] > N
    # T + 150 < gust_period: exactly one gust per episode whose first step has run (an env re-sampled at the end of the last
    # step has i == 0 and draws its gust on its next step)
    assert torch.equal(env.gust_count, env.episode - (env.i == 0).int())
    with pytest.raises(L.QuadSimError):
        env.rollout(4)
    res = {}
    for robust in (False, True):
        e = BatchedQuad(N, 0.01, 1000, training=False, direct_control=0, T=1, precision="f32", robust_control=robust, seed=6, device=DEV)
        e.reset()
        e.control_rollout(ctl.lqr_controller(), 400)
        v = e.state[:, 1:6:2].norm(dim=1)
        res[robust] = (float(v.nanmedian()), float(torch.isfinite(v).float().mean()), float((v < 5.0).float().mean()))
    # N(0, 0.3) mass perturbations leave a handful of envs with a near-zero or negative mass (as in the reference's model)
    assert res[False][1] == 1.0 and res[True][1] > 0.99, res
    assert res[False][2] > 0.9 and res[True][2] > 0.5, res
    assert res[True][0] > 2 * res[False][0], res                  # gusts of N(0, 5) m/s and a wrong mass leave a larger residual speed


def test_control_rollout_f32_million_env_comparison():
    """The README's controller comparison at scale: FP32 RK4, 262,144 envs from the reset distribution, 400 fused steps of
    each law.  Property checks (size-independent): both laws stabilise (the median speed at least halves within 4 s),
    nothing becomes non-finite, and FP32 agrees with the FP64 kernel on a sub-sample within the closed-loop FP32 bound."""
    from autonomous_quadrotor_environment_b200 import controllers as ctl
    N, K = 1 << 18, 400
    for c in (ctl.lqr_controller(), ctl.pid_controller()):
        env = BatchedQuad(N, 0.01, 10 ** 6, training=False, direct_control=0, T=1, clipped=True, precision="f32", seed=4, device=DEV)
        env.reset()
        st0 = env.state.clone()
        env.control_rollout(c, K)
        st = env.state
        assert bool(torch.isfinite(st).all())
        v0, v1 = st0[:, (1, 3, 5)].norm(dim=1).median(), st[:, (1, 3, 5)].norm(dim=1).median()
        assert float(v1) < 0.5 * float(v0), (float(v0), float(v1))
        sub = 512
        e64 = BatchedQuad(sub, 0.01, 10 ** 6, training=False, direct_control=0, T=1, clipped=True, precision="f64", integrator="rk4",
                          seed=4, device=DEV)
        e64.reset(st0[:sub].double())
        e32 = BatchedQuad(sub, 0.01, 10 ** 6, training=False, direct_control=0, T=1, clipped=True, precision="f32", seed=4, device=DEV)
        e32.reset(st0[:sub])
        e64.control_rollout(c, 100); e32.control_rollout(c, 100)
        a, b = e32.state.double()[:, 1:].cpu().numpy(), e64.state[:, 1:].cpu().numpy()     # positions integrate the error: skip x
        ok = np.abs(a - b) <= 2e-3 + 2e-3 * np.abs(b)
        assert ok.mean() > 0.999, ok.mean()


# ----------------------------------------------------------------------------------------------------
# BASELINE.json full size (configs[2]): size-independent properties
# ----------------------------------------------------------------------------------------------------
def test_full_size_properties_1M_envs():
    N, K = 1 << 20, 64
    a = BatchedQuad(N, 0.01, 1000, T=5, precision="f32", auto_reset=True, seed=0, device=DEV)
    b = BatchedQuad(N, 0.01, 1000, T=5, precision="f32", auto_reset=True, seed=0, device=DEV)
    a.reset(); b.reset()
    a.rollout(K)
    for _ in range(K // 16):
        b.rollout(16)
    # determinism + fusion-invariance: one 64-step launch == four 16-step launches
    assert torch.equal(a.state, b.state)
    st = a.state
    assert torch.isfinite(st).all()
    qn = st[:, 6:10].norm(dim=1)
    assert (qn - 1).abs().max() < 1e-3                      # integrated quaternion stays near unit norm
    # bounding boxes hold for every env that is alive (done envs were reset in-kernel)
    assert (st[:, 1:6:2].abs() < 10).all() and (st[:, 10:13].abs() < 20).all()
    s = a.stats()
    assert s["n_steps"] == N * K and s["n_episodes"] > 0
    assert s["n_solved"] + s["n_broken"] + s["n_timeout"] == s["n_episodes"]
    # episode counters: every finished episode bumped exactly one counter (plus the initial reset)
    assert int(a.episode.sum().item()) == N + int(s["n_episodes"])


def test_full_size_headline_config_properties():
    """BASELINE.json configs[2] as bench.py times it — 1,048,576 envs, FP32 RK4, sensor noise, asynchronous auto-reset, T=5, the
    two-envs-per-lane step kernel with the packed sensor model — through properties that do not need the oracle: (1) sharding is
    invisible (two half-size handles with env-id offsets == the whole, bit for bit: state, sensed observation, episode counters);
    (2) K fused steps in one launch (sensor state on chip) == K single-step launches on the same action stream; (3) accounting."""
    N, K = 1 << 20, 24
    mk = lambda n, off: BatchedQuad(n, 0.01, 1000, T=5, precision="f32", async_reset=True, sensor_noise=True, seed=0,
                                    env_id_offset=off, device=DEV)
    whole, lo, hi, fused = mk(N, 0), mk(N // 2, 0), mk(N // 2, N // 2), mk(N, 0)
    for e in (whole, lo, hi, fused):
        e.reset()
    assert whole.step_loader == 3
    g = torch.Generator(device=DEV); g.manual_seed(7)
    acts = (torch.rand(K, 4, N, device=DEV, generator=g) * 2 - 1).contiguous()
    for t in range(K):
        whole.step_soa(acts[t])
        lo.step_soa(acts[t, :, :N // 2].contiguous()); hi.step_soa(acts[t, :, N // 2:].contiguous())
    fused.rollout(K, actions=acts)
    for name in ("state", "sensed_obs", "episode", "reward", "done_flags"):
        w = getattr(whole, name)
        parts = torch.cat([getattr(lo, name), getattr(hi, name)], dim=0)
        assert torch.equal(w, parts), name
        assert torch.equal(w, getattr(fused, name)), "fused rollout: " + name
    st, so = whole.state, whole.sensed_obs
    assert torch.isfinite(st).all() and torch.isfinite(so).all()
    assert (st[:, 6:10].norm(dim=1) - 1).abs().max() < 1e-3
    s = whole.stats()
    assert s["n_steps"] == N * K and s["n_episodes"] > N // 100
    assert s["n_solved"] + s["n_broken"] + s["n_timeout"] == s["n_episodes"]
    assert int(whole.episode.sum().item()) == N + int(s["n_episodes"])
    sl, sh = lo.stats(), hi.stats()
    assert sl["n_episodes"] + sh["n_episodes"] == s["n_episodes"] and abs(sl["sum_return"] + sh["sum_return"] - s["sum_return"]) < 1e-6 * abs(s["sum_return"])
    # the sensor is in the loop: on ordinary steps the sensed velocity differs from the true one, by centimetres per second
    live = (whole.done_flags & 3) == 0
    dv = (so[:, 1:6:2] - st[:, 1:6:2]).abs()[live]
    assert 1e-5 < float(dv.mean()) < 0.5


@pytest.mark.gpu
def test_handle_on_second_device_while_first_is_current():
    """One process driving two GPUs: a handle created on cuda:1 keeps launching there (kernels, shared-memory attributes,
    host-buffer pipeline) while the thread's current device is cuda:0, and produces the same trajectory as on cuda:0."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    N, seed = 4096, 4
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        torch.cuda.set_device(0)
        env = BatchedQuad(N, 0.01, 50, T=3, precision="f32", async_reset=True, sensor_noise=True, seed=seed, device=dev)
        env.reset()
        g = torch.Generator(device="cpu"); g.manual_seed(2)
        for t in range(40):
            torch.cuda.set_device(0)
            a = (torch.rand(4, N, generator=g) * 2 - 1).to(dev).contiguous()
            torch.cuda.synchronize(dev)
            env.step_soa(a)
        torch.cuda.synchronize(dev)
        outs.append((env.obs.cpu(), env.sensed_obs.cpu(), env.episode.cpu()))
    assert torch.equal(outs[0][2], outs[1][2])
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


# ----------------------------------------------------------------------------------------------------
# sensor model against the REFERENCE class driven by a replayed draw stream (tests/golden/sensor_vectors.npz)
# ----------------------------------------------------------------------------------------------------
def _sv(g, pre, t, mk):
    soa = lambda x: mk(np.ascontiguousarray(np.asarray(x).reshape(x.shape[0], -1).T))
    return dict(quad_state=soa(g[pre + "state"][t]), acc_read=soa(g[pre + "acc_read"][t]), mat_rot=soa(g[pre + "mat_rot"][t]),
                f_m=mk(g[pre + "f_m"][t]))


def _sensor_state0(g, pre, mk):
    from autonomous_quadrotor_environment_b200 import sensors as S
    n = g[pre + "u"].shape[0]
    s = S.new_state(n, dtype=mk(np.zeros(1)).dtype, device=DEV)
    S.sensor_call("reset", s, mk(np.ascontiguousarray(g[pre + "u"][:, 3:6].T)), quad_state=mk(np.ascontiguousarray(g[pre + "reset_state"].T)))
    return s


def test_sensor_methods_f64_vs_reference_class():
    """qs_sensor_call, one reference method per launch, fed the very draws the reference's `sensor` consumed: every output of
    accel_int / gyro_int / gyro / gps / triad (incl. SciPy's matrix->quaternion) and the sensor's whole internal state
    within 1e-9, free-running over the episode; scenario C calls the methods in another order."""
    from autonomous_quadrotor_environment_b200 import sensors as S
    g = load_golden("sensor_vectors.npz")
    cat = lambda *a: np.concatenate(a, axis=1)
    for pre, order in (("A_", [("accel_int", 9, lambda g_, t: g_["A_accel_int"][t]), ("gyro_int", 3, lambda g_, t: g_["A_gyro_int"][t]),
                               ("gyro", 3, lambda g_, t: g_["A_gyro"][t]), ("gps", 6, lambda g_, t: g_["A_gps"][t]),
                               ("triad", 6, lambda g_, t: cat(g_["A_triad_q"][t], g_["A_triad_R"][t].reshape(-1, 9)))]),
                       ("C_", [("gyro", 3, lambda g_, t: g_["C_gyro"][t]),
                               ("triad", 6, lambda g_, t: cat(g_["C_triad_q"][t], g_["C_triad_R"][t].reshape(-1, 9))),
                               ("gps", 6, lambda g_, t: g_["C_gps"][t]), ("gyro_int", 3, lambda g_, t: g_["C_gyro_int"][t]),
                               ("accel", 3, lambda g_, t: g_["C_accel"][t]), ("accel_int", 9, lambda g_, t: g_["C_accel_int"][t])])):
        s = _sensor_state0(g, pre, T64)
        K = g[pre + "z"].shape[0]
        for t in range(K):
            kw = _sv(g, pre, t, T64)
            zc = 0
            for name, nz, want in order:
                z = T64(np.ascontiguousarray(g[pre + "z"][t][:, zc:zc + nz].T)); zc += nz
                out = npy(S.sensor_call(name, s, z, **kw)).T
                ref = want(g, t)
                assert rel_err(out[:, :ref.shape[1]], ref) < 1e-9, (pre, t, name)
            assert rel_err(npy(s).T, g[pre + "sens_state"][t], floor=1e-6) < 1e-9, (pre, t)


@pytest.mark.parametrize("pre", ["A_", "B_"])
def test_sensor_step_vs_reference_sensor_sp(pre):
    """The fused canonical step (the function the step kernels run per env, here with the reference's draws) against the
    reference's own sensor_sp (visual_landing/math_trajectory.py:61-83): B_ = its GPS blend switched on (GPS_P = 30).
    FP64 free-running at 1e-9; FP32 teacher-forced (state restarted from the reference every step) within the production
    bound 1e-5 + 1e-4 |x|."""
    from autonomous_quadrotor_environment_b200 import sensors as S
    g = load_golden("sensor_vectors.npz")
    blend = float(g["B_gps_blend"]) if pre == "B_" else 0.0
    s64 = _sensor_state0(g, pre, T64)
    s32 = _sensor_state0(g, pre, T32)
    worst32 = 0.0
    for t in range(g[pre + "z"].shape[0]):
        z = np.ascontiguousarray(g[pre + "z"][t].T)
        o64 = npy(S.sensor_call("step", s64, T64(z), params={"gps_blend": blend}, **_sv(g, pre, t, T64))).T
        assert rel_err(o64, g[pre + "obs"][t]) < 1e-9, t
        assert rel_err(npy(s64).T, g[pre + "sens_state"][t], floor=1e-6) < 1e-9, t
        o32 = npy(S.sensor_call("step", s32, T32(z), params={"gps_blend": blend}, **_sv(g, pre, t, T32))).T
        worst32 = max(worst32, bound_err(o32, g[pre + "obs"][t]), bound_err(npy(s32).T[:, 4:17], g[pre + "sens_state"][t][:, 4:17]))
        s32.copy_(T32(np.ascontiguousarray(g[pre + "sens_state"][t].T)))                  # teacher forcing
    assert worst32 < 1.0, worst32


def test_dropin_sensor_class_reproduces_reference_readings():
    """The single-env drop-in `sensor(quad)` of the compat overlay, driven like the reference was when the fixture was
    recorded (same NumPy-level draws through oracle/replay_rng.py): same readings, method by method, to 1e-9."""
    from autonomous_quadrotor_environment_b200.quadrotor_env import quad, sensor
    from oracle.replay_rng import ReplayRNG
    g = load_golden("sensor_vectors.npz")
    for j in range(2):
        env = quad(0.01, 10 ** 6, training=False, direct_control=1, T=2, verbose=False, robust_rng_draws=False)
        env.reset(g["C_init"][j].copy())
        with ReplayRNG([], g["C_u"][j]):
            sen = sensor(env)
            sen.reset()
        assert rel_err(env.state, g["C_reset_state"][j]) < 1e-9
        for t in range(g["C_z"].shape[0]):
            env.step(g["C_actions"][t, j])
            with ReplayRNG(g["C_z"][t, j]):
                w = sen.gyro(); qt, Rt = sen.triad(); pg, vg = sen.gps(); qg = sen.gyro_int(); ac = sen.accel(); acc, vel, pos = sen.accel_int()
            for got, key in ((w, "C_gyro"), (qt, "C_triad_q"), (Rt, "C_triad_R"), (np.concatenate([pg, vg]), "C_gps"), (qg, "C_gyro_int"),
                             (ac, "C_accel"), (np.concatenate([acc, vel, pos]), "C_accel_int")):
                assert rel_err(got, g[key][t, j]) < 1e-9, (j, t, key)
            assert rel_err(sen.R[:, 2], g["C_sens_state"][t, j][14:17]) < 1e-9 and abs(sen.a_b_accel - g["C_sens_state"][t, j][0]) < 1e-15


# ----------------------------------------------------------------------------------------------------
# sensor model inside the fused K-step rollout (BASELINE.json configs[2], "with sensor noise and auto-reset")
# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("prec,integ,N,src", [("f32", "rk4", 4096, "philox"), ("f32", "rk4", 4098, "buffer"), ("f32", "rk4", 1001, "buffer"),
                                              ("f64", "rk45", 512, "buffer")])
def test_sensor_rollout_equals_repeated_steps(prec, integ, N, src):
    """qs_rollout on a QS_FLAG_SENSOR_NOISE handle (FP32: rollout_pair_kernel<.,SENSOR>, sensor state on chip for the whole horizon;
    odd N / FP64: generic kernel) against K calls of qs_step on a twin handle: recorded sensed observation, true observation,
    reward, done of every step, then the sensor state and every other row of the workspace — through asynchronous resets,
    warm-up steps and sensor resets.  FP32 compares two differently-fused kernels: teacher-forced every 8 steps, tolerance of
    FP32 rounding; the Philox streams (sensor noise, re-sampling, in-kernel actions) must be identical."""
    K, chunk, seed = 48, 8, 21
    dt = torch.float32 if prec == "f32" else torch.float64
    mk = lambda: BatchedQuad(N, 0.01, 30, T=3, precision=prec, integrator=integ, async_reset=True, sensor_noise=True, seed=seed,
                             device=DEV, params={"gps_blend": 20.0})
    a, b = mk(), mk()
    a.reset(); b.reset()
    tol = dict(rtol=3e-5, atol=3e-5) if prec == "f32" else dict(rtol=1e-11, atol=1e-11)
    g = torch.Generator(device=DEV); g.manual_seed(9)
    n_flip = 0
    for c0 in range(0, K, chunk):
        b._ws.copy_(a._ws)
        if src == "buffer":
            acts = (torch.rand(chunk, 4, N, device=DEV, dtype=dt, generator=g) * 2 - 1)
            rec = a.rollout(chunk, actions=acts, record_obs=True, record_reward=True, record_done=True, record_sensed=True, record_actions=True)
        else:
            rec = a.rollout(chunk, record_obs=True, record_reward=True, record_done=True, record_sensed=True, record_actions=True)
            acts = rec["actions"]
        same = torch.ones(N, dtype=torch.bool, device=DEV)
        for t in range(chunk):
            obs, rew, done = b.step_soa(acts[t].contiguous())
            same &= (b.done_flags == rec["done"][t])
            n_flip += int((~same).sum()) if t == chunk - 1 else 0
            assert torch.allclose(b.sensed_obs[same], rec["sensed_obs"][t].t()[same], **tol), (c0, t)
            assert torch.allclose(obs[same], rec["obs"][t].t()[same], **tol), (c0, t)
            assert torch.allclose(rew[same], rec["reward"][t][same], **tol), (c0, t)
        for f in (L.QS_FIELD_SENSOR_STATE, L.QS_FIELD_SENSED_OBS, L.QS_FIELD_OBS, L.QS_FIELD_ANG, L.QS_FIELD_ABS_SUM):
            assert torch.allclose(a._field(f)[:, same], b._field(f)[:, same], **tol), (c0, f)
        for f in (L.QS_FIELD_I, L.QS_FIELD_EPISODE, L.QS_FIELD_FLAGS):
            assert torch.equal(a._field(f)[:, same], b._field(f)[:, same]), (c0, f)
    assert n_flip <= (2 if prec == "f32" else 0), n_flip
    assert int(a.episode.max()) >= 2                                   # resets, warm-up steps and sensor resets happened
    d = a.sensed_obs[:, 10:14] - a.obs[:, 10:14]
    assert 0.003 < float(d.std()) < 0.06                               # the noise is there (gyro sigma 0.035 rad/s)


def test_fused_critic_head_values_match_torch_critic():
    """Critic head of the fused policy kernel (model.py:36-43 on the same history tile as the actor): (1) the rollout itself is
    untouched — actions, observations, dones bit-identical to the kernel without the critic; (2) the recorded state values,
    rows 0..K-1 = V(input of step t) and row K = V(input after the last step), equal the FP32 torch critic evaluated on the
    reconstructed (BF16-rounded) network inputs up to BF16-operand error."""
    from autonomous_quadrotor_environment_b200 import ppo as P
    g = load_golden("actor_128.npz")
    N, K, seed = 1000, 24, 4
    torch.manual_seed(3)
    ac = P.ActorCritic(128, 75, 4, 0.1).to(DEV)
    with torch.no_grad():                                            # a critic with O(1) outputs and both signs in every layer
        for m in ac.critic:
            if hasattr(m, "weight"):
                m.weight.mul_(2.0); m.bias.normal_(0, 0.3)
    crit = {"critic_%d_%s" % (i, k): getattr(ac.critic[i], k).detach() for i in (0, 2, 4) for k in ("weight", "bias")}
    mk = lambda: BatchedQuad(N, 0.01, 300, training=True, direct_control=1, T=5, precision="f32", async_reset=True, seed=seed, device=DEV)
    a, b = mk(), mk()
    a.reset(); b.reset()
    a.load_actor(g, action_std=0.1, critic=crit)
    b.load_actor(g, action_std=0.1)
    hist0 = a.history.t().contiguous().clone()
    ra = a.policy_rollout(K, record_obs=True, record_values=True)
    rb = b.policy_rollout(K, record_obs=True)
    for k in ("actions", "obs", "done", "reward", "logprob"):
        assert torch.equal(ra[k], rb[k]), k
    entries = P.BatchedPPO.history_entries(ra)
    x = P.BatchedPPO.network_inputs(None, hist0, entries, 0, N)      # (K+1, N, 75)
    with torch.no_grad():
        v_ref = ac.critic(x).squeeze(-1)
    err = (ra["value"] - v_ref).abs()
    assert float(v_ref.abs().mean()) > 0.2                           # the comparison is not about zeros
    assert float(err.max()) < 0.06 and float(err.mean()) < 0.01, (float(err.max()), float(err.mean()))
    with pytest.raises(L.QuadSimError):
        b.policy_rollout(4, record_values=True)                      # no critic loaded
