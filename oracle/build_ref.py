"""TEST INFRASTRUCTURE ONLY — "build" of the reference for the GPU box (oracle/_ref/).

The reference is plain Python with no package metadata (nothing to pip-install) and
/root/reference does not exist on the GPU box.  Like a C reference compiled from its own sources into
oracle/_ref/*.so, the modules of the hot path and the controller scripts that drive it are compiled
FROM THE SOURCES WHERE THEY LIE (/root/reference, unmodified) into sourceless CPython bytecode
(`py_compile`, same interpreter on both sides) and packed, together with the binary data files the scripts
read, into ONE archive (the box's snapshot drops *.pyc files, so they cannot travel loose):

    oracle/_ref/ref_build.bin   (zip)
        pyc/environment/quadrotor_env.pyc                  the reference `quad` / `sensor`
        pyc/environment/quaternion_euler_utility.pyc
        pyc/environment/controller/{lqr_quad,pid_vel_control,ppo_quad_eval,dl_auxiliary,model,
                                    response_analyzer,target_parser}.pyc
        pyc/mission_control/mission_control.pyc
        data/solved/nn_old_solved_128_32000_*.pth          the trained actor ppo_quad_eval.py loads
        data/classical_controller_results/*_same_start*.npy    the author's five shipped logs
    oracle/_ref/MANIFEST.json   source path + sha256 of every input

oracle/_ref/ is git-ignored (no reference source or derivative enters the history) but travels to the
GPU box with the working tree; oracle/ref_runtime.py unpacks the archive into a temporary directory at first use.
Users:
  * tests/test_reference_scripts.py — executes the UNMODIFIED controller scripts against the compat/ overlay;
  * bench.py --impl reference / cpu_baseline — times the reference's own quad.step on the box's host cores.
Nothing under the product package reads this directory.

    python oracle/build_ref.py            # (re)build; a no-op when /root/reference is absent
"""
import glob
import hashlib
import json
import os
import py_compile
import shutil
import sys
import tempfile
import warnings
import zipfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("QUADSIM_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(ROOT, "oracle", "_ref")

MODULES = [
    "environment/quadrotor_env.py",
    "environment/quaternion_euler_utility.py",
    "environment/controller/lqr_quad.py",
    "environment/controller/pid_vel_control.py",
    "environment/controller/ppo_quad_eval.py",
    "environment/controller/dl_auxiliary.py",
    "environment/controller/model.py",
    "environment/controller/response_analyzer.py",
    "environment/controller/target_parser.py",
    "mission_control/mission_control.py",
]
DATA = [
    ("environment/controller/solved/nn_old_solved_128_32000_*.pth", "solved"),
    ("environment/controller/classical_controller_results/*_same_start*.npy", "classical_controller_results"),
]


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


ARCHIVE = os.path.join(OUT, "ref_build.bin")


def available() -> bool:
    return os.path.isfile(os.path.join(OUT, "MANIFEST.json")) and os.path.isfile(ARCHIVE)


def build(force: bool = False) -> bool:
    """Returns True when oracle/_ref is present (built now or before)."""
    if not os.path.isfile(os.path.join(REF, "environment", "quadrotor_env.py")):
        return available()
    manifest = {"python": sys.version.split()[0], "reference_root": REF, "files": {}}
    for rel in MODULES:
        manifest["files"][rel] = _sha(os.path.join(REF, rel))
    data_files = []
    for pattern, sub in DATA:
        for src in sorted(glob.glob(os.path.join(REF, pattern))):
            data_files.append((src, sub))
            manifest["files"][os.path.relpath(src, REF)] = _sha(src)
    mpath = os.path.join(OUT, "MANIFEST.json")
    if not force and available():
        try:
            if json.load(open(mpath)) == manifest:
                return True
        except Exception:
            pass
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    with tempfile.TemporaryDirectory() as tmp, zipfile.ZipFile(ARCHIVE, "w", zipfile.ZIP_DEFLATED) as zf, warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)            # the reference's matplotlib labels use "\p" escapes
        for rel in MODULES:
            dst = os.path.join(tmp, os.path.basename(rel) + "c")
            py_compile.compile(os.path.join(REF, rel), cfile=dst, dfile="<reference>/" + rel, doraise=True,
                               invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
            zf.write(dst, "pyc/" + rel[:-3] + ".pyc")
        for src, sub in data_files:
            zf.write(src, "data/%s/%s" % (sub, os.path.basename(src)))
    with open(mpath, "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "ready" if ok else "reference tree not present, nothing built")
