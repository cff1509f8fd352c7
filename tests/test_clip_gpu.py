"""Conditioning-prep (CLIP text tower) parity (pytest -m gpu).  The case list lives in tests/clip_checks.py."""
import pytest

pytestmark = pytest.mark.gpu


def pytest_generate_tests(metafunc):
    if "clip_case" in metafunc.fixturenames:
        import os
        import sys
        sys.path.insert(0, os.path.dirname(__file__))
        try:
            import clip_checks as cc
            cases = cc.ALL
        except Exception as ex:
            cases = [(f"unavailable: {ex!r}"[:120], None, {}, 0.0)]
        metafunc.parametrize("clip_case", cases, ids=[c[0] for c in cases])


def test_clip(clip_case):
    import torch
    name, fn, kw, tol = clip_case
    assert fn is not None, name
    err = fn(**kw)
    torch.cuda.synchronize()
    assert err < tol, f"{name}: {err:.3e} >= {tol:g}"
