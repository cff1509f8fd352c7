"""The C-ABI library loads without a GPU and exports exactly what include/ltt_b200.h declares (no compute calls)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ltt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ltt_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    from layoutllm_t2i_b200 import _lib
    assert sorted(_lib.EXPORTS) == _declared()


def test_library_loads_and_exports_every_symbol():
    from layoutllm_t2i_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    lib = _lib.load()          # getattr on every declared symbol: AttributeError if one is missing
    assert lib.ltt_version().decode().startswith("ltt_b200")
    assert lib.ltt_launch_count(None) == 0
    assert lib.ltt_last_error() is not None


def test_missing_library_fails_loudly(monkeypatch):
    from layoutllm_t2i_b200 import _lib
    monkeypatch.setattr(_lib, "LIB_PATH", os.path.join(ROOT, "does_not_exist.so"))
    with pytest.raises(_lib.LttError, match="no CPU fallback"):
        _lib.load()


def test_engine_needs_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from layoutllm_t2i_b200 import _lib
    from layoutllm_t2i_b200.engine import Engine
    with pytest.raises(_lib.LttError, match="no CPU fallback"):
        Engine({}, 0)


def test_no_oracle_import_in_product_package():
    pkg = os.path.join(ROOT, "layoutllm_t2i_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, os.path.join(d, f)
