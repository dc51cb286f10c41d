"""Overlay module: put `<repo>/compat` on sys.path BEFORE the reference root and the reference's
`from environment.quadrotor_env import quad, sensor, plotter` resolves here (the reference has no
__init__.py files, so `environment` is a PEP 420 namespace package and every other
`environment.*` module still resolves from the reference tree).  See INTEGRATION.md."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.append(_root)
from autonomous_quadrotor_environment_b200.quadrotor_env import *  # noqa: F401,F403,E402
from autonomous_quadrotor_environment_b200.quadrotor_env import quad, sensor, plotter, robust_control  # noqa: F401,E402
