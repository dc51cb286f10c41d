"""Drop-in for environment/quaternion_euler_utility.py, computed by the CUDA library.

Same names, argument shapes and return shapes as the reference (`euler_quat` :17, `quat_euler` :39,
`quat_euler_2` :50, `deriv_quat` :58, `quat_rot_mat` :71) for single quaternions, plus ``*_batch``
variants over torch tensors with a leading env axis.  No CPU fallback: needs libquadsim.so and a GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L


def _prec(t: torch.Tensor) -> int:
    if t.dtype == torch.float64:
        return L.QS_F64
    if t.dtype == torch.float32:
        return L.QS_F32
    raise TypeError("float32 or float64 tensors only")


def _soa(x: torch.Tensor, c: int) -> torch.Tensor:
    if x.dim() != 2 or x.shape[1] != c:
        raise ValueError("expected shape (N,%d)" % c)
    return x.t().contiguous()


def _stream(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def euler_quat_batch(ang: torch.Tensor) -> torch.Tensor:
    """(N,3) -> (N,4)"""
    lib = L.load_library()
    a = _soa(ang, 3)
    out = torch.empty(4, a.shape[1], dtype=a.dtype, device=a.device)
    L.check(lib.qs_euler_quat(_prec(a), a.shape[1], a.data_ptr(), out.data_ptr(), _stream(a)))
    return out.t()


def quat_euler_batch(q: torch.Tensor) -> torch.Tensor:
    """(N,4) -> (N,3)"""
    lib = L.load_library()
    a = _soa(q, 4)
    out = torch.empty(3, a.shape[1], dtype=a.dtype, device=a.device)
    L.check(lib.qs_quat_euler(_prec(a), a.shape[1], a.data_ptr(), out.data_ptr(), _stream(a)))
    return out.t()


def deriv_quat_batch(w: torch.Tensor, q: torch.Tensor) -> torch.Tensor:
    """(N,3),(N,4) -> (N,4)"""
    lib = L.load_library()
    ww, qq = _soa(w, 3), _soa(q, 4)
    out = torch.empty_like(qq)
    L.check(lib.qs_deriv_quat(_prec(qq), qq.shape[1], ww.data_ptr(), qq.data_ptr(), out.data_ptr(), _stream(qq)))
    return out.t()


def quat_rot_mat_batch(q: torch.Tensor) -> torch.Tensor:
    """(N,4) -> (N,3,3)"""
    lib = L.load_library()
    qq = _soa(q, 4)
    out = torch.empty(9, qq.shape[1], dtype=qq.dtype, device=qq.device)
    L.check(lib.qs_quat_rot_mat(_prec(qq), qq.shape[1], qq.data_ptr(), out.data_ptr(), _stream(qq)))
    return out.t().reshape(-1, 3, 3)


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError("quaternion_euler_utility needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _one(x, c):
    return torch.as_tensor(np.asarray(x, dtype=np.float64).reshape(1, c), device=_dev())


def euler_quat(ang):
    """reference :17-36 — returns a (4,1) float64 array."""
    return euler_quat_batch(_one(ang, 3)).cpu().numpy().reshape(4, 1)


def quat_euler(q):
    """reference :39-48 — q of shape (4,1); returns array([phi, theta, psi])."""
    out = quat_euler_batch(_one(q, 4)).cpu().numpy().reshape(3)
    if np.any(np.isnan(out)):
        print('Divergencia na conversao Quaternion - Euler')
    return out


def quat_euler_2(q):
    """reference :50-56 — same math for a flat (4,) quaternion."""
    return quat_euler(q)


def deriv_quat(w, q):
    """reference :58-69 — returns a flat (4,) array."""
    return deriv_quat_batch(_one(w, 3), _one(q, 4)).cpu().numpy().reshape(4)


def quat_rot_mat(q):
    """reference :71-80 — returns a (3,3) array."""
    return quat_rot_mat_batch(_one(q, 4)).cpu().numpy().reshape(3, 3)
