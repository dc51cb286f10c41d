"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/quadsim.h declares,
the ctypes mirror of its structs has the C layout, and compute entry points fail loudly without a GPU."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

from conftest import ROOT
from autonomous_quadrotor_environment_b200 import _lib as L

HEADER = os.path.join(ROOT, "include", "quadsim.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qs_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = L.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "libquadsim.so does not export %s" % name
    assert sorted(L.EXPORTED_SYMBOLS) == declared
    assert lib.qs_version() == 100


def test_struct_layouts_match_header(tmp_path):
    prog = tmp_path / "sz.c"
    prog.write_text('#include "quadsim.h"\n#include <stdio.h>\n#include <stddef.h>\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %d %zu %zu\\n",'
                    'sizeof(qs_params),sizeof(qs_config),sizeof(qs_field_desc),sizeof(qs_stats),sizeof(qs_rollout_args),'
                    'sizeof(qs_controller),sizeof(qs_control_rollout_args),sizeof(qs_policy_rollout_args),sizeof(qs_actor),'
                    'offsetof(qs_config,params),offsetof(qs_config,workspace),(int)QS_FIELD_COUNT_,sizeof(qs_ppo_batch),sizeof(qs_ppo_net));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split()
    sizes = [int(x) for x in out]
    assert sizes[0] == C.sizeof(L.qs_params)
    assert sizes[1] == C.sizeof(L.qs_config)
    assert sizes[2] == C.sizeof(L.qs_field_desc)
    assert sizes[3] == C.sizeof(L.qs_stats)
    assert sizes[4] == C.sizeof(L.qs_rollout_args)
    assert sizes[5] == C.sizeof(L.qs_controller)
    assert sizes[6] == C.sizeof(L.qs_control_rollout_args)
    assert sizes[7] == C.sizeof(L.qs_policy_rollout_args)
    assert sizes[8] == C.sizeof(L.qs_actor)
    assert sizes[9] == L.qs_config.params.offset
    assert sizes[10] == L.qs_config.workspace.offset
    assert sizes[11] == L.QS_FIELD_COUNT
    assert sizes[12] == C.sizeof(L.qs_ppo_batch) and sizes[13] == C.sizeof(L.qs_ppo_net)


def test_default_config_is_the_reference_constants():
    c = L.default_config()
    p = c.params
    # environment/quadrotor_env.py:30-80
    assert (p.mass, p.gravity, p.rho, p.c_d) == (1.03, 9.82, 1.2041, 1.1)
    assert (p.k_f, p.k_m, p.i_r, p.t2wr) == (1.435e-5, 2.4086e-7, 5e-5, 2)
    assert list(p.j) == [16.83e-3, 16.83e-3, 28.34e-3]
    assert (p.arm, p.beam_thickness, p.bb_vel, p.bb_pos) == (0.26, 0.05, 10, 5)
    assert list(p.tr) == [0.005, 0.01, 0.1] and list(p.tr_p) == [3, 2, 1]
    assert (p.solved_reward, p.broken_reward, p.shaping_weight, p.p_c) == (20, -20, 5, 0.003)
    # quad() keyword defaults :112
    assert c.T == 1 and c.flags == (L.QS_FLAG_DIRECT_CONTROL | L.QS_FLAG_CLIPPED | L.QS_FLAG_TRAINING)


def test_argument_validation():
    lib = L.load_library()
    c = L.default_config()
    c.n_envs = 0
    assert lib.qs_workspace_bytes(C.byref(c)) == L.QS_EINVAL
    c = L.default_config(); c.T = 0
    h = C.c_void_p()
    assert lib.qs_create(C.byref(h), C.byref(c)) == L.QS_EINVAL
    assert b"T must be" in lib.qs_last_error()
    c = L.default_config(); c.n_envs = 1 << 20
    nbytes = lib.qs_workspace_bytes(C.byref(c))
    # 17+3+4 real rows + i/episode + 3 byte rows + action staging: ~ (24+4)*4 + 8 + 3 bytes per env
    assert 100 * (1 << 20) < nbytes < 140 * (1 << 20)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = L.load_library()
    c = L.default_config()
    h = C.c_void_p()
    rc = lib.qs_create(C.byref(h), C.byref(c))
    assert rc == L.QS_ECUDA and b"no CPU fallback" in lib.qs_last_error()
    from autonomous_quadrotor_environment_b200 import BatchedQuad
    with pytest.raises(RuntimeError):
        BatchedQuad(4)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "autonomous_quadrotor_environment_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"(from|import)\s+oracle|oracle[/.]|quad_oracle|c_oracle", src), \
                    "%s references the oracle" % f
