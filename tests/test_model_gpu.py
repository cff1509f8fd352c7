"""Model-level parity (pytest -m gpu): whole UNet forward and PLMS loop through the C-ABI against the oracle and the
reference-generated goldens.  The case list lives in tests/model_checks.py."""
import pytest

pytestmark = pytest.mark.gpu


def pytest_generate_tests(metafunc):
    if "model_case" in metafunc.fixturenames:
        import os
        import sys
        sys.path.insert(0, os.path.dirname(__file__))
        try:
            import model_checks as mc
            cases = mc.ALL
        except Exception:
            cases = [("libltt_b200.so missing", None, {}, 0.0)]
        metafunc.parametrize("model_case", cases, ids=[c[0] for c in cases])


def test_model(model_case):
    import torch
    name, fn, kw, tol = model_case
    assert fn is not None, "layoutllm_t2i_b200/libltt_b200.so is not built"
    err = fn(**kw)
    torch.cuda.synchronize()
    assert err < tol, f"{name}: rel-L2 {err:.3e} >= {tol:g}"
