"""CPU tests of the C-ABI boundary: the library builds, loads and exports what include/vfk.h declares."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "vfk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vfk_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_bound_entry_points():
    from vfclik_b200 import _lib
    assert header_functions() == sorted(_lib.EXPORTS)


def test_library_loads_and_exports_every_symbol(built_lib):
    from vfclik_b200 import _lib
    lib = _lib.load()
    for name in header_functions():
        assert hasattr(lib, name), name
    assert lib.vfk_version() == 100


def test_struct_layouts_match_the_header(built_lib, tmp_path):
    """ctypes mirrors of vfk_chain_desc / vfk_params / vfk_buffers have the sizes the header implies,
    and vfk_default_params fills the documented defaults."""
    import subprocess
    from vfclik_b200 import _lib
    # the header is plain C: compile it with gcc and ask the compiler for the layouts
    prog = tmp_path / "sizes.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "vfk.h"\n'
                    'int main(void){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(vfk_chain_desc), sizeof(vfk_params),'
                    ' sizeof(vfk_buffers), offsetof(vfk_params, ns_mode), offsetof(vfk_buffers, qdot),'
                    ' offsetof(vfk_chain_desc, tip)); return 0;}\n')
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert sizes == [C.sizeof(_lib.ChainDescC), C.sizeof(_lib.ParamsC), C.sizeof(_lib.BuffersC),
                     _lib.ParamsC.ns_mode.offset, _lib.BuffersC.qdot.offset, _lib.ChainDescC.tip.offset]
    lib = _lib.load()
    p = _lib.ParamsC()
    lib.vfk_default_params(C.byref(p), 7)
    assert (p.ik_lambda, p.ns_gain, p.ns_lookahead, p.jp_delta) == (0.1, 0.5, 0.3, 0.087)
    assert list(p.mixer_w) == [1.0, 1.0, 0.0, 0.0, 0.0, 0.0]          # scripts/bridge:596
    assert list(p.tool) == [1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0] and p.direct_control == -1


def test_python_defaults_match_the_library(built_lib):
    from vfclik_b200 import _lib
    from vfclik_b200.engine import Params
    lib = _lib.load()
    want = _lib.ParamsC()
    lib.vfk_default_params(C.byref(want), 7)
    got = Params().to_c(7)
    assert bytes(got) == bytes(want)


def test_argument_errors_are_codes_not_exceptions(built_lib, lwr):
    """Every call returns 0 or a negative code with a message (no exceptions cross the ABI)."""
    from vfclik_b200 import _lib
    lib = _lib.load()
    chain, _ = lwr
    h = C.c_void_p()
    cd = _lib.chain_to_c(chain)
    assert lib.vfk_create(C.byref(h), C.byref(cd), 16, 0) == _lib.VFK_ERR_INVALID
    assert b"precision" in lib.vfk_last_error(None)
    cd.n_joints = 0
    assert lib.vfk_create(C.byref(h), C.byref(cd), 32, 0) == _lib.VFK_ERR_INVALID
    cd.n_joints = 99
    assert lib.vfk_create(C.byref(h), C.byref(cd), 32, 0) == _lib.VFK_ERR_INVALID
    assert lib.vfk_step(None, None, 1, 32, 0, 4, 1, None) == _lib.VFK_ERR_INVALID


def test_no_cpu_fallback(built_lib, lwr):
    """Without a B200 the engine refuses to exist: there is no CPU compute path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from vfclik_b200 import _lib
    from vfclik_b200.engine import Engine
    with pytest.raises(_lib.VfkError) as ei:
        Engine(lwr[0], precision=32)
    assert ei.value.code == _lib.VFK_ERR_NO_DEVICE


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from vfclik_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libvfk.so"))
    with pytest.raises(_lib.VfkLibraryError):
        _lib.load()


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under vfclik_b200/ may import or execute it."""
    pkg = os.path.join(ROOT, "vfclik_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "oracle/" not in text or f.endswith((".cu", ".cuh")) or "oracle/batch.py" in text, f


def test_integration_md_ctypes_stub_matches_the_header():
    """The reference-side stub printed in INTEGRATION.md must describe the same structs as include/vfk.h (via _lib.py)."""
    import ctypes as C
    from vfclik_b200 import _lib
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = text[text.index("class ChainDesc(C.Structure):"):text.index("lib.vfk_last_error.restype")]
    block = "\n".join(l for l in block.splitlines() if not l.startswith("#"))
    ns = {"C": C}
    exec(block, ns)
    assert C.sizeof(ns["ChainDesc"]) == C.sizeof(_lib.ChainDescC)
    assert C.sizeof(ns["Params"]) == C.sizeof(_lib.ParamsC)
    assert [f[0] for f in ns["Params"]._fields_] == [f[0] for f in _lib.ParamsC._fields_]
