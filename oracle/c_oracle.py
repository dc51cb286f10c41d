"""TEST INFRASTRUCTURE ONLY — ctypes wrapper of oracle/quad_oracle.c (plain-C restatement of the reference path).
Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libquad_oracle.so")


class qo_envs(C.Structure):
    _fields_ = [("n", C.c_int64), ("integrator", C.c_int), ("substeps", C.c_int), ("direct", C.c_int),
                ("clipped", C.c_int), ("training", C.c_int), ("n_limit", C.c_int), ("dt", C.c_double),
                ("state", C.c_void_p), ("prev_ang", C.c_void_p), ("prev_shaping", C.c_void_p), ("flags", C.c_void_p),
                ("step_i", C.c_void_p), ("abs_sum", C.c_void_p)]


def load():
    if not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "quad_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    lib = C.CDLL(_SO)
    lib.qo_step_batch.restype = None
    lib.qo_step_batch.argtypes = [C.POINTER(qo_envs), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    lib.qo_max_threads.restype = C.c_int
    return lib


class COracle:
    """N reference `quad` envs in plain C.  integrator: 'rk45' (what the reference runs) or 'rk4'."""

    def __init__(self, n_envs, t_step, n, training=True, direct_control=1, T=1, clipped=True, integrator="rk45",
                 substeps=1, threads=0):
        self.lib = load()
        self.N, self.T, self.threads = n_envs, T, threads
        self.state = np.zeros((n_envs, 13))
        self.prev_ang = np.zeros((n_envs, 3))
        self.prev_shaping = np.zeros(n_envs)
        self.flags = np.ones(n_envs, dtype=np.uint8)
        self.step_i = np.zeros(n_envs, dtype=np.int32)
        self.abs_sum = np.zeros(n_envs)
        self.obs = np.zeros((n_envs, 14))
        self.reward = np.zeros(n_envs)
        self.done = np.zeros(n_envs, dtype=np.uint8)
        self.nfev = np.zeros(n_envs, dtype=np.int32)
        self.direct = int(bool(direct_control))
        self.zero_control = np.zeros(4) if self.direct else np.array([1.03 * 9.82, 0, 0, 0])
        self.e = qo_envs(n_envs, 1 if integrator == "rk45" else 0, substeps, self.direct, int(bool(clipped)),
                         int(bool(training)), n + T, t_step, self.state.ctypes.data, self.prev_ang.ctypes.data,
                         self.prev_shaping.ctypes.data, self.flags.ctypes.data, self.step_i.ctypes.data,
                         self.abs_sum.ctypes.data)

    def reset(self, det_state):
        self.state[:] = det_state
        self.flags[:] = 0
        self.step_i[:] = 0
        self.abs_sum[:] = 0
        out = []
        for _ in range(self.T):
            o, _, _ = self.step(np.broadcast_to(self.zero_control, (self.N, 4)))
            out.append(o.copy())
        return np.array(out)

    def step(self, action):
        a = np.ascontiguousarray(action, dtype=np.float64)
        self.lib.qo_step_batch(C.byref(self.e), a.ctypes.data, self.obs.ctypes.data, self.reward.ctypes.data,
                               self.done.ctypes.data, self.nfev.ctypes.data, self.threads)
        return self.obs, self.reward, self.done.astype(bool)
