import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


@pytest.fixture(scope="session")
def golden():
    return load_golden


def rel_err(x, ref, floor=1e-3):
    """max |x-ref| / max(|ref|, floor) — the per-step metric of BASELINE.json configs[1]."""
    x, ref = np.asarray(x, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return float(np.max(np.abs(x - ref) / np.maximum(np.abs(ref), floor)))


def bound_err(x, ref, rtol=1e-4, atol=1e-5):
    """max |x-ref| / (atol + rtol*|ref|) — FP32 production-mode bound of BASELINE.json (<= 1 passes)."""
    x, ref = np.asarray(x, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return float(np.max(np.abs(x - ref) / (atol + rtol * np.abs(ref))))


def lqr_action(K_t, K_att, env_state, env_ang, env_ang_vel, euler_t_ant, dt=0.01, M=1.03, G=9.82):
    """Test-side restatement of the control law of environment/controller/lqr_quad.py:129-157.
    Returns (action[F,Mx,My,Mz], euler_t)."""
    state_t = np.array([0, env_state[1], 0, env_state[3], 0, env_state[5]])
    F = np.dot(K_t, state_t)
    theta_t = np.arctan2(F[0], (F[2] + G))
    phi_t = np.arctan2(-F[1] * np.cos(theta_t), (F[2] + G))
    euler_t = np.array([phi_t, theta_t, 0])
    U_1 = M * (F[2] + G) / (np.cos(theta_t) * np.cos(phi_t))
    euler = env_ang - euler_t
    d_euler = env_ang_vel
    state_att = np.array([euler[0], d_euler[0], euler[1], d_euler[1], euler[2], d_euler[2]])
    action = np.dot(K_att, state_att)
    action[0] = U_1
    return action, euler_t


# argument sets of the mission-generator fixture (oracle/gen_golden.py:gen_mission_vectors) and of the test that replays them
MISSION_CASES = {
    "gen": lambda m: m.gen_trajectory(500, 200, np.array([1., -2., 3.])),
    "gen_add": lambda m: m.gen_trajectory(300, 100, np.array([1., -2., 3.]), additive=np.arange(14) * 0.1),
    "gen_vel": lambda m: m.gen_trajectory(250, 250, np.zeros(3), velocity=np.array([1., .5, -.2])),
    "sin": lambda m: m.sin_trajectory(400, 2.0, 0.3, np.array([1., 2., 0.]), np.array([1., 0.5, 0.])),
    "spiral": lambda m: m.spiral_trajectory(150, 400, 0.5, 1.5, 2.0, np.array([0., 1., 2.])),
}

