"""Batched `sensor` methods with caller-provided draws — host-side mirror of `qs_sensor_call` (include/quadsim.h).

The reference's `sensor` class (environment/quadrotor_env.py:579-724) one method per call for n independent sensors held as
structure-of-arrays torch tensors on the device.  `BatchedQuad(sensor_noise=True)` runs the same device functions inside the
step kernels with Philox draws; this entry point takes the draws as an argument, which is what the single-env drop-in
`quadrotor_env.sensor` (NumPy's global stream) and the parity tests against the reference class use."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L

METHODS = {"reset": L.QS_SENSOR_RESET, "accel": L.QS_SENSOR_ACCEL, "gyro": L.QS_SENSOR_GYRO, "gps": L.QS_SENSOR_GPS,
           "triad": L.QS_SENSOR_TRIAD, "accel_int": L.QS_SENSOR_ACCEL_INT, "gyro_int": L.QS_SENSOR_GYRO_INT, "step": L.QS_SENSOR_STEP}


def new_state(n: int, dtype=torch.float64, device="cuda") -> torch.Tensor:
    """Fresh (QS_SENSOR_STATE_DIM, n) sensor state: biases 0, self.R = I (sensor.__init__ :597)."""
    s = torch.zeros(L.QS_SENSOR_STATE_DIM, n, dtype=dtype, device=device)
    s[16] = 1.0
    return s


def sensor_call(method, sensor_state: torch.Tensor, z: torch.Tensor, quad_state=None, acc_read=None, mat_rot=None, f_m=None,
                t_step: float = 0.01, params: dict | None = None) -> torch.Tensor | None:
    """Run one method for all n sensors.  sensor_state (20,n) is updated in place; z (k,n) standard normals (uniforms for
    "reset"); quad_state (13,n), acc_read (3,n), mat_rot (9,n) row-major, f_m (n,) as the method needs them.  Returns (m,n)."""
    m = METHODS[method] if isinstance(method, str) else int(method)
    dt, dev, n = sensor_state.dtype, sensor_state.device, sensor_state.shape[1]
    if dt not in (torch.float32, torch.float64) or sensor_state.shape[0] != L.QS_SENSOR_STATE_DIM or not sensor_state.is_contiguous():
        raise ValueError("sensor_state must be a contiguous (%d,n) float32/float64 tensor" % L.QS_SENSOR_STATE_DIM)

    def chk(t, rows, name):
        if t is None:
            return None
        if t.dtype != dt or t.device != dev or not t.is_contiguous() or t.numel() != rows * n:
            raise ValueError("%s must be a contiguous (%d,n) tensor of the state's dtype and device" % (name, rows))
        return C.c_void_p(t.data_ptr())

    nz, no = L.SENSOR_Z_ROWS[m], L.SENSOR_OUT_ROWS[m]
    if z.shape[0] < nz:
        raise ValueError("method needs %d rows of draws" % nz)
    par = L.default_config().params
    for k, v in (params or {}).items():
        setattr(par, k, float(v))
    out = torch.empty(max(no, 1), n, dtype=dt, device=dev)
    with torch.cuda.device(dev):
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        L.check(L.load_library().qs_sensor_call(L.QS_F64 if dt == torch.float64 else L.QS_F32, C.byref(par), float(t_step), n, m,
                                                C.c_void_p(sensor_state.data_ptr()), chk(quad_state, 13, "quad_state"),
                                                chk(acc_read, 3, "acc_read"), chk(mat_rot, 9, "mat_rot"), chk(f_m, 1, "f_m"),
                                                chk(z[:nz].contiguous() if z.shape[0] != nz else z, nz, "z"), C.c_void_p(out.data_ptr()), st))
    return out[:no] if no else None
