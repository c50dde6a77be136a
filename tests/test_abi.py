"""CPU-side checks of the C-ABI boundary: the library loads and exports every symbol include/fgnn.h
declares; the host wrapper refuses to run without a GPU (no CPU fallback)."""
import os
import re

import pytest

from conftest import ROOT
from multiagent_gnn_policies_b200 import engine


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fgnn.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fgnn_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(engine.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(engine.LIB_PATH):
        from multiagent_gnn_policies_b200 import build
        build.build()
    lib = engine.load_library()
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.fgnn_version() >= 100


def test_config_struct_layout_matches_header():
    import ctypes
    assert ctypes.sizeof(engine.FgnnConfig) == 18 * 4 + 3 * 8
    assert ctypes.sizeof(engine.FgnnStats) == 8 + 8 + 4 + 4 + 8 + 8 + 8       # ... + n_ghosts


def test_engine_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(engine.FgnnError):
        engine.FlockEngine(n_agents=10)


def test_header_compiles_as_plain_c(tmp_path):
    """The drop-in boundary is a C ABI: include/fgnn.h must be valid C99 (no C++ in the signatures), and every declared
    entry point must be addressable with the declared prototype."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    names = declared_symbols()
    body = "\n".join(f"    p[{i}] = (void*)&{n};" for i, n in enumerate(names))
    src = tmp_path / "abi_check.c"
    src.write_text('#include "fgnn.h"\n#include <stddef.h>\n'
                   f"void* fgnn_abi_table[{len(names)}];\n"
                   "void fgnn_abi_fill(void) {\n    void** p = fgnn_abi_table;\n" + body + "\n}\n"
                   "size_t fgnn_cfg_size(void) { return sizeof(fgnn_config) + sizeof(fgnn_stats); }\n")
    out = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-Wno-pedantic", "-I", os.path.join(ROOT, "include"), "-c",
                          str(src), "-o", str(tmp_path / "abi_check.o")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr


def test_argument_errors_are_reported_without_a_gpu():
    """Error convention of the boundary: non-zero status + fgnn_last_error() text; argument checks come before any CUDA call."""
    import ctypes
    lib = engine.load_library()
    h = ctypes.c_void_p()
    assert lib.fgnn_create(None, ctypes.byref(h)) != 0
    assert b"null" in lib.fgnn_last_error()
    bad = engine.FgnnConfig(100, 1, 9, 6, 2, 32, 2, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1.0, 0.01, 10.0)       # k = 9
    assert lib.fgnn_create(ctypes.byref(bad), ctypes.byref(h)) != 0
    assert b"k must be in 1..4" in lib.fgnn_last_error()
    bad = engine.FgnnConfig(100, 1, 3, 5, 2, 32, 2, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1.0, 0.01, 10.0)       # n_states = 5
    assert lib.fgnn_create(ctypes.byref(bad), ctypes.byref(h)) != 0
    assert b"n_states" in lib.fgnn_last_error()
    tr = ctypes.c_void_p()
    assert lib.fgnn_trainer_create(7, 32, 2, 0, ctypes.byref(tr)) != 0
    assert b"k must be in 1..4" in lib.fgnn_last_error()
    assert lib.fgnn_trainer_step(None, 1, 1, None, None, None, None, None, 1, 1e-3, 0.9, 0.999, 1e-8, 1, None, None, None) != 0
    assert lib.fgnn_step(None, None, None, None) != 0 and lib.fgnn_policy(None, None, None) != 0
    assert lib.fgnn_destroy(None) == 0 and lib.fgnn_trainer_destroy(None) == 0       # destroying nothing is fine


def test_general_actor_forward_argument_checks():
    """fgnn_actor_forward_general / fgnn_actor_general_workspace (learner/actor.py:45-86 for any ind_agg): shapes are
    checked before any CUDA call, the workspace size is 2 ping-pong buffers of the widest activation."""
    import ctypes
    lib = engine.load_library()
    widths = (ctypes.c_int32 * 4)(6, 16, 24, 2)
    assert lib.fgnn_actor_general_workspace(2, 40, 3, 3, widths) == 2 * (2 * 24 * 3 * 40) * 4
    assert lib.fgnn_actor_general_workspace(2, 40, 3, 3, None) == -1
    assert lib.fgnn_actor_general_workspace(0, 40, 3, 3, widths) == -1
    bad = (ctypes.c_int32 * 4)(6, 0, 24, 2)
    assert lib.fgnn_actor_general_workspace(2, 40, 3, 3, bad) == -1
    assert b"widths" in lib.fgnn_last_error()
    assert lib.fgnn_actor_forward_general(0, 2, 40, 3, 3, widths, 1, None, None, None, None, None, None, None) != 0
    assert b"null" in lib.fgnn_last_error()
    # ind_agg beyond the layers with K > 1 leaves K rows: the reference's final view fails, so does this call
    one = ctypes.c_float(0.0)
    ptrs = (ctypes.c_void_p * 3)(*[ctypes.addressof(one)] * 3)
    p = ctypes.addressof(one)
    assert lib.fgnn_actor_forward_general(0, 2, 40, 3, 3, widths, 3, ptrs, ptrs, p, p, p, p, None) != 0
    assert b"ind_agg" in lib.fgnn_last_error()
