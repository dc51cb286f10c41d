"""BASELINE.json configs[0]: one env through the drop-in `quad` (N = 1 handle, FP64 + RK45 replica): wall time per quad.step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from autonomous_quadrotor_environment_b200.quadrotor_env import quad
for direct in (1, 0):
    env = quad(0.01, 100000, training=False, euler=0, direct_control=direct, T=1, clipped=True, verbose=False)
    env.seed(1)
    env.reset()
    a = np.zeros(4) if direct else np.array([env.mass * env.gravity if hasattr(env, "gravity") else 10.1, 0, 0, 0])
    for _ in range(50):
        env.step(a)
    t0 = time.perf_counter()
    K = 1000
    for k in range(K):
        obs, r, d = env.step(a + (0.01 * np.sin(k) if direct else 0.0))
    dt = (time.perf_counter() - t0) / K
    print("drop-in quad.step direct_control=%d: %.1f us per step (%.0f steps/s)" % (direct, dt * 1e6, 1 / dt), flush=True)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for k in range(300): env.step(a)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
