"""Overlay for `environment.quaternion_euler_utility` (see compat/environment/quadrotor_env.py)."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.append(_root)
from autonomous_quadrotor_environment_b200.quaternion_euler_utility import (  # noqa: F401,E402
    euler_quat, quat_euler, quat_euler_2, deriv_quat, quat_rot_mat)
