"""VAE decode + image post-processing parity (pytest -m gpu) against the unmodified reference AutoencoderKL (oracle/_ref).
The case list lives in tests/vae_checks.py."""
import pytest

pytestmark = pytest.mark.gpu


def pytest_generate_tests(metafunc):
    if "vae_case" in metafunc.fixturenames:
        import os
        import sys
        sys.path.insert(0, os.path.dirname(__file__))
        try:
            import vae_checks as vc
            cases = vc.ALL
        except Exception as ex:
            cases = [(f"unavailable: {ex!r}"[:120], None, {}, 0.0)]
        metafunc.parametrize("vae_case", cases, ids=[c[0] for c in cases])


def test_vae(vae_case):
    import torch
    from oracle import ref_loader as rl
    name, fn, kw, tol = vae_case
    assert fn is not None, name
    assert rl.available(), "oracle/_ref is not staged (python oracle/make_ref.py in the build container)"
    err = fn(**kw)
    torch.cuda.synchronize()
    assert err < tol, f"{name}: {err:.3e} >= {tol:g}"
