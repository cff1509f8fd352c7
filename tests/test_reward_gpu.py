"""Reward-path (CLIP vision tower, reward head, Reward mirror) parity (pytest -m gpu).  The case list lives in tests/reward_checks.py."""
import pytest

pytestmark = pytest.mark.gpu


def pytest_generate_tests(metafunc):
    if "reward_case" in metafunc.fixturenames:
        import os
        import sys
        sys.path.insert(0, os.path.dirname(__file__))
        try:
            import reward_checks as cc
            cases = cc.ALL
        except Exception as ex:
            cases = [(f"unavailable: {ex!r}"[:120], None, {}, 0.0)]
        metafunc.parametrize("reward_case", cases, ids=[c[0] for c in cases])


def test_reward(reward_case):
    import torch
    name, fn, kw, tol = reward_case
    assert fn is not None, name
    err = fn(**kw)
    torch.cuda.synchronize()
    assert err < tol, f"{name}: {err:.3e} >= {tol:g}"
