"""CPU-side checks of the boundary: the shared library builds, loads without a GPU, exports
every symbol include/zaremba_b200.h declares, and refuses to compute without a device."""
import ctypes as C
import os
import re

import pytest
import torch

from zaremba_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "zaremba_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(zrb_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_loads_without_gpu():
    path = build.build()
    assert os.path.exists(path)
    lib = _lib.load()
    assert b"sm_100a" in lib.zrb_version()


def test_every_declared_symbol_is_exported_and_bound():
    lib = C.CDLL(build.LIB)
    names = _header_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert names == _lib.exported_symbols(), "python binding table out of step with the header"


def test_struct_layouts_match_header():
    assert C.sizeof(_lib.ZrbConfig) == 32
    assert C.sizeof(_lib.ZrbParams) == 8 * (3 + 4 * _lib.MAX_LAYERS)
    assert C.sizeof(_lib.ZrbStates) == 8 * 2 * _lib.MAX_LAYERS


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    lib = _lib.load()
    cfg = _lib.ZrbConfig(10, 8, 1, 2, 2, 0, 0.0, 0)
    h = C.c_void_p()
    rc = lib.zrb_ctx_create(C.byref(cfg), C.byref(h))
    assert rc == -2 and b"no CPU path" in lib.zrb_last_error()
    import zaremba_b200
    m = zaremba_b200.Model(10, 8, 1, 0.0, 0.1)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(2, 2, dtype=torch.long), m.state_init(2))


def test_product_never_imports_the_oracle():
    """SPEC: only tests/, __graft_entry__.smoke() and bench.py's cpu legs may touch oracle/."""
    pkg = os.path.join(ROOT, "zaremba_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"(import\s+oracle|from\s+oracle|oracle[/.]\w)", txt), f
