"""TEST INFRASTRUCTURE ONLY — run the reference's UNMODIFIED modules and controller scripts from the bytecode
build in oracle/_ref (oracle/build_ref.py), here and on the GPU box (where /root/reference does not exist).

Two uses:
  * `reference_tree(overlay=True)`  + `run_script(...)`: the reference's controller scripts
    (environment/controller/{lqr_quad,pid_vel_control,ppo_quad_eval}.py) executed as they are against the
    compat/ overlay, i.e. with `environment.quadrotor_env` resolving to the CUDA-backed drop-in and every other
    module (`environment.controller.model`, `dl_auxiliary`, `mission_control` ...) to the reference's own code;
  * `reference_tree(overlay=False)`: the same with the reference's own `environment.quadrotor_env`
    (bench.py --impl reference times it on the box's host cores).

Environment fixes applied around the scripts — harness side, no source change:
  matplotlib stub (not installed), `os.chdir` neutralised and the author's home-directory prefix mapped to
  oracle/_ref/data for `torch.load`, `np.save` captured instead of written, stdout swallowed, and the RNG
  draws of `robust_control.reset` — added to the reference after its shipped logs were recorded —
  suppressed when `suppress_robust_rng` is set.
"""
import contextlib
import io
import marshal
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BUILD = os.path.join(ROOT, "oracle", "_ref")
ARCHIVE = os.path.join(REF_BUILD, "ref_build.bin")
COMPAT = os.path.join(ROOT, "compat")
AUTHOR_HOME = "/home/rafaelcostaf/mestrado/quadrotor_environment/"

_REF_PACKAGES = ("environment", "mission_control")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_BUILD, "MANIFEST.json")) and os.path.isfile(ARCHIVE)


_unpacked = None


def unpacked() -> str:
    """The archive of oracle/build_ref.py unpacked into a per-process temporary directory (pyc/ and data/ below it)."""
    global _unpacked
    if _unpacked is None:
        import atexit
        import shutil
        import tempfile
        import zipfile
        if not available():
            raise RuntimeError("oracle/_ref not built: run `python oracle/build_ref.py` where /root/reference exists")
        d = tempfile.mkdtemp(prefix="quadsim_ref_")
        with zipfile.ZipFile(ARCHIVE) as zf:
            zf.extractall(d)
        atexit.register(shutil.rmtree, d, ignore_errors=True)
        _unpacked = d
    return _unpacked


def pyc_root() -> str:
    return os.path.join(unpacked(), "pyc")


def data_root() -> str:
    return os.path.join(unpacked(), "data")


def _purge_modules():
    for name in list(sys.modules):
        if name.split(".")[0] in _REF_PACKAGES:
            del sys.modules[name]


@contextlib.contextmanager
def reference_tree(overlay: bool):
    """sys.path = [compat (if overlay), oracle/_ref/pyc] + rest; reference packages purged from sys.modules on both sides."""
    if not available():
        raise RuntimeError("oracle/_ref not built: run `python oracle/build_ref.py` where /root/reference exists")
    from oracle.ref_import import _install_matplotlib_stub
    _install_matplotlib_stub()
    saved_path = list(sys.path)
    _purge_modules()
    sys.path[:0] = ([COMPAT] if overlay else []) + [pyc_root()]
    try:
        yield
    finally:
        sys.path[:] = saved_path
        _purge_modules()


def load_code(rel_module: str):
    """Code object of a module of the bytecode build, e.g. 'environment/controller/lqr_quad'."""
    with open(os.path.join(pyc_root(), rel_module + ".pyc"), "rb") as f:
        data = f.read()
    return marshal.loads(data[16:])


class _Switches(dict):
    """Module namespace in which some top-level names are pinned: the scripts choose their variant with an in-file
    switch (`clipped = True`); pinning the name is the harness-side equivalent of the author flipping it."""

    def __init__(self, pinned):
        super().__init__()
        self._pinned = dict(pinned)

    def __setitem__(self, k, v):
        super().__setitem__(k, self._pinned.get(k, v))


def run_script(rel_module: str, overlay: bool = True, switches=None, suppress_robust_rng: bool = True, quiet: bool = True):
    """Execute a reference controller script (module-level code) and return {basename: array} of what it np.save()d."""
    import numpy as np
    saved = {}
    with reference_tree(overlay):
        import importlib
        env_mod = importlib.import_module("environment.quadrotor_env")
        rc_cls, rc_reset = env_mod.robust_control, env_mod.robust_control.reset     # (the overlay re-exports the package's class)
        if suppress_robust_rng:
            rc_cls.reset = lambda self: None
        orig_save, orig_chdir = np.save, os.chdir
        torch = sys.modules.get("torch")
        if torch is None:
            import torch
        orig_load = torch.load

        def fake_save(path, arr, *a, **k):
            saved[os.path.basename(str(path))] = np.array(arr)

        def fake_load(f, *a, **k):
            if isinstance(f, str) and f.startswith(AUTHOR_HOME):
                f = os.path.join(data_root(), "solved", os.path.basename(f))
            return orig_load(f, *a, **k)

        np.save, os.chdir, torch.load = fake_save, (lambda p: None), fake_load
        ns = _Switches(switches or {})
        ns["__name__"] = "__main__"
        ns["__builtins__"] = __builtins__
        try:
            out = io.StringIO()
            with (contextlib.redirect_stdout(out) if quiet else contextlib.nullcontext()):
                exec(load_code(rel_module), ns, ns)
        finally:
            np.save, os.chdir, torch.load = orig_save, orig_chdir, orig_load
            rc_cls.reset = rc_reset
    return saved


def shipped_log(name: str):
    import numpy as np
    return np.load(os.path.join(data_root(), "classical_controller_results", name))


def import_reference_env():
    """The reference's own environment.quadrotor_env (bytecode build), for a process that does nothing else with these
    package names (bench.py's CPU legs): leaves the unpacked bytecode tree on sys.path."""
    if not available():
        raise RuntimeError("oracle/_ref not built")
    from oracle.ref_import import _install_matplotlib_stub
    _install_matplotlib_stub()
    _purge_modules()
    if pyc_root() not in sys.path:
        sys.path.insert(0, pyc_root())
    import importlib
    return importlib.import_module("environment.quadrotor_env")
