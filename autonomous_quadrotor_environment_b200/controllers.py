"""The reference's classical comparison controllers for the batched simulator.

`lqr_controller()` / `pid_controller()` build the `qs_controller` description the in-kernel control laws of
`qs_control_rollout` consume (csrc/controller_rollout.cuh):
  * LQR — environment/controller/lqr_quad.py: the two gain matrices are the solutions of two continuous-time algebraic
    Riccati equations (:82-111), solved ONCE on the host with SciPy exactly as the script does; the per-step law
    (:129-157) runs on the device.
  * PID — environment/controller/pid_vel_control.py: cascaded velocity -> attitude loops (:29-127); gains :17-27.
Both command [F, Mx, My, Mz]: use a `BatchedQuad(..., direct_control=0)`.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L

# environment/quadrotor_env.py:39,55-57 — the scripts hard-code the same numbers (lqr_quad.py:16-20)
_M, _G = 1.03, 9.82
_J = (16.83e-3, 16.83e-3, 28.34e-3)


def lqr_gains(clipped: bool = True):
    """(K_t (3,6), K_att (4,6)) exactly as lqr_quad.py:25-111 computes them."""
    from scipy.linalg import solve_continuous_are
    if clipped:                                                               # :25-43
        Q_att = np.diag([5, 1, 5, 1, 0.05, 0.01]) * 50.0
        Q_t = np.diag([1e-08, 1, 1e-08, 1, 1e-08, 0.8]) * 10.0
        R_t = np.eye(3) * 10.0
    else:                                                                     # :44-62
        Q_att = np.diag([5, 0.3, 5, 0.3, 2, 0.3]) * 160.0
        Q_t = np.diag([1e-08, 1, 1e-08, 1, 1e-08, 0.5]) * 60.0
        R_t = np.eye(3) * 5.0
    R_att = np.eye(4) * 40.0
    A = np.zeros((6, 6)); A[0, 1] = A[2, 3] = A[4, 5] = 1.0                   # :67-72, :88-93
    B_att = np.zeros((6, 4)); B_att[1, 1] = 1 / _J[0]; B_att[3, 2] = 1 / _J[1]; B_att[5, 3] = 1 / _J[2]
    B_t = np.zeros((6, 3)); B_t[1, 0] = B_t[3, 1] = B_t[5, 2] = 1 / _M
    K_att = -np.linalg.inv(R_att) @ (B_att.T @ solve_continuous_are(A, B_att, Q_att, R_att))   # :82-86
    K_t = -np.linalg.inv(R_t) @ (B_t.T @ solve_continuous_are(A, B_t, Q_t, R_t))               # :107-111
    return K_t, K_att


def _new(kind: int, clipped: bool) -> L.qs_controller:
    c = L.qs_controller()
    L.check(L.load_library().qs_default_controller(C.byref(c), kind, int(bool(clipped))))
    return c


def lqr_controller(clipped: bool = True, K_t=None, K_att=None) -> L.qs_controller:
    """LQR of lqr_quad.py (gains computed here unless given)."""
    c = _new(L.QS_CTRL_LQR, clipped)
    if K_t is None or K_att is None:
        K_t, K_att = lqr_gains(clipped)
    K_t, K_att = np.asarray(K_t, dtype=np.float64), np.asarray(K_att, dtype=np.float64)
    if K_t.shape != (3, 6) or K_att.shape != (4, 6):
        raise ValueError("K_t must be (3,6) and K_att (4,6)")
    for r in range(3):
        for k in range(6):
            c.k_t[r][k] = K_t[r, k]
    for r in range(4):
        for k in range(6):
            c.k_att[r][k] = K_att[r, k]
    return c


def pid_controller(clipped: bool = True, target_vel=(0.0, 0.0, 0.0), target_psi: float = 0.0) -> L.qs_controller:
    """Cascaded PID of pid_vel_control.py with its gains (:17-27) and the script's set-points (:150-153) by default."""
    c = _new(L.QS_CTRL_PID, clipped)
    for k in range(3):
        c.target_vel[k] = float(target_vel[k])
    c.target_psi = float(target_psi)
    return c
