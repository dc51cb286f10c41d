"""Developer tool: distribution of the FP32-vs-FP64 closed-loop error over 4,096 envs x 1000 steps (see tests/test_gpu_fp32_parity.py)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import test_gpu_fp32_parity as T
from conftest import bound_err

N, steps = 4096, 1000
W, env, ora, hg, ho, mk = T._closed_loop("f32", "rk4", N, steps, 31)
nonpos = [1, 3, 5, 6, 7, 8, 9, 10, 11, 12, 13]
per_env = np.zeros(N); per_env_rew = np.zeros(N); when = np.zeros(N, int)
alive = np.ones(N, bool)
max_ang = np.zeros(N); max_w = np.zeros(N); max_v = np.zeros(N)
for t in range(steps):
    ag, ao = T._np_actor(W, hg), T._np_actor(W, ho)
    obs, rew, done = env.step(mk(ag))
    o_ref, r_ref, d_ref = ora.step(ao)
    og = T.npy(obs)
    hg = T._push(hg, og, ag); ho = T._push(ho, o_ref, ao)
    alive &= ~d_ref & np.isfinite(o_ref).all(axis=1)
    ang = T.qo.quat_euler(ora.state[:, 6:10] / np.linalg.norm(ora.state[:, 6:10], axis=1, keepdims=True))
    e = np.max(np.abs(og[:, nonpos] - o_ref[:, nonpos]) / (1e-5 + 1e-4 * np.abs(o_ref[:, nonpos])), axis=1)
    er = np.abs(T.npy(rew) - r_ref) / (1e-5 + 1e-4 * np.abs(r_ref))
    upd = alive & (e > per_env)
    when[upd] = t
    per_env = np.where(alive, np.maximum(per_env, e), per_env)
    ok = alive & T.threshold_margin_ok(ora.state, ang)
    per_env_rew = np.where(ok, np.maximum(per_env_rew, er), per_env_rew)
    max_ang = np.maximum(max_ang, np.where(alive, np.abs(ang).max(1), 0)); max_w = np.maximum(max_w, np.where(alive, np.abs(ora.state[:, 10:13]).max(1), 0))
    max_v = np.maximum(max_v, np.where(alive, np.abs(ora.state[:, 1:6:2]).max(1), 0))
print("alive", alive.mean())
for q in (0.5, 0.9, 0.99, 0.999, 1.0):
    print("quantile %.3f: obs err/bound %.3g   reward err/bound %.3g" % (q, np.quantile(per_env[alive], q), np.quantile(per_env_rew[alive], q)))
print("envs over bound:", int((per_env[alive] > 1).sum()), "of", int(alive.sum()))
worst = np.argsort(-per_env * alive)[:12]
for j in worst:
    print("env %4d err/bound %8.3g at step %4d  max|ang| %.2f max|w| %.2f max|v| %.2f rew %.3g" % (j, per_env[j], when[j], max_ang[j], max_w[j], max_v[j], per_env_rew[j]))
# how the error depends on how violent the transient was
for lo, hi in ((0, 0.6), (0.6, 1.0), (1.0, 1.3), (1.3, 1.6)):
    m = alive & (max_ang >= lo) & (max_ang < hi)
    if m.any():
        print("max|ang| in [%.1f,%.1f): %4d envs, worst err/bound %.3g, over bound %d" % (lo, hi, m.sum(), per_env[m].max(), (per_env[m] > 1).sum()))
