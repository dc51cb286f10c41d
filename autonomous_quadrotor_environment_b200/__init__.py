"""B200-native batched quadrotor simulator — drop-in for the `quad` hot path of
rafaelcostafrf/autonomous_quadrotor_environment (environment/quadrotor_env.py).

    from autonomous_quadrotor_environment_b200 import BatchedQuad      # N envs in lock-step (torch tensors)
    from autonomous_quadrotor_environment_b200.quadrotor_env import quad  # the reference's single-env class

All compute runs in hand-written sm_100a CUDA kernels behind the C ABI of include/quadsim.h
(libquadsim.so, built in-tree by __graft_entry__.build()); there is no CPU fallback.
"""
from ._lib import LIB_PATH, QuadSimError, load_library  # noqa: F401


def __getattr__(name):
    if name == "BatchedQuad":
        from .batched import BatchedQuad
        return BatchedQuad
    if name == "quad":
        from .quadrotor_env import quad
        return quad
    raise AttributeError(name)
