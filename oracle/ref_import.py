"""TEST INFRASTRUCTURE ONLY — importer for the *unmodified* reference (Python).

Only usable in the build container (where /root/reference exists).  Used by
``oracle/gen_golden.py`` to generate the committed fixtures in ``tests/golden``
and by the ``not gpu`` tests that validate the restatement in
``oracle/quad_oracle.py`` against the real reference when it is present.

Nothing under the product package imports this file.

Harness shims (environment fixes, not code changes):
  * ``matplotlib`` is not installed and ``environment/quadrotor_env.py:5-7,20-28``
    imports/configures it at module scope -> stub modules are injected.
  * ``robust_control.reset`` (``environment/quadrotor_env.py:97-101``) consumes 12
    NumPy RNG draws per ``reset()`` that post-date the shipped golden logs; pass
    ``suppress_robust_rng=True`` to neutralise it when reproducing those logs.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("QUADSIM_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "environment", "quadrotor_env.py"))


def _install_matplotlib_stub():
    if "matplotlib" in sys.modules and not getattr(sys.modules["matplotlib"], "_quadsim_stub", False):
        return

    class _Anything:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return _Anything()

        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            return _Anything()

        def __iter__(self):
            return iter([_Anything(), _Anything()])

        def __getitem__(self, i):
            return _Anything()

    mpl = types.ModuleType("matplotlib")
    mpl._quadsim_stub = True
    mpl.use = lambda *a, **k: None
    mpl.rcParams = {}
    plt = types.ModuleType("matplotlib.pyplot")

    def _plt_getattr(name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    plt.__getattr__ = _plt_getattr
    mpl.pyplot = plt
    tk = types.ModuleType("mpl_toolkits")
    tk3 = types.ModuleType("mpl_toolkits.mplot3d")
    tk3.Axes3D = _Anything
    tk.mplot3d = tk3
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)
    sys.modules.setdefault("mpl_toolkits", tk)
    sys.modules.setdefault("mpl_toolkits.mplot3d", tk3)


def load_reference(suppress_robust_rng: bool = False):
    """Return the reference module ``environment.quadrotor_env`` (unmodified source)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_matplotlib_stub()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # make sure the *reference* module is the one resolved, not the compat overlay
    for name in ("environment.quadrotor_env", "environment.quaternion_euler_utility"):
        mod = sys.modules.get(name)
        if mod is not None and not getattr(mod, "__file__", "").startswith(REFERENCE_ROOT):
            del sys.modules[name]
    import importlib

    env_mod = importlib.import_module("environment.quadrotor_env")
    if not env_mod.__file__.startswith(REFERENCE_ROOT):
        raise RuntimeError("resolved %s instead of the reference" % env_mod.__file__)
    if suppress_robust_rng:
        env_mod.robust_control.reset = lambda self: None
    return env_mod


def load_reference_utility():
    load_reference()
    import importlib

    return importlib.import_module("environment.quaternion_euler_utility")
