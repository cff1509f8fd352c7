"""compute-sanitizer memcheck over a small end-to-end workload of the sm_100a library (tools/sanitize_smoke.py): tiny UNet
forward at both gate values, a fused 5-step PLMS loop with CUDA-graph replay, a small VAE decode, tiny CLIP towers, the image
preprocessing and the cluster reward head.  (pytest -m gpu)"""
import os
import shutil
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_memcheck_clean():
    tool = shutil.which("compute-sanitizer") or "/usr/local/cuda/bin/compute-sanitizer"
    if not os.path.exists(tool):
        pytest.skip("compute-sanitizer not installed on this box")
    env = dict(os.environ, LTT_NO_AUTOTUNE="1")          # the tuner only repeats launches the run makes anyway
    out = subprocess.run([tool, "--tool", "memcheck", "--error-exitcode", "77", sys.executable,
                          os.path.join(ROOT, "tools", "sanitize_smoke.py")], capture_output=True, text=True, timeout=1500, cwd=ROOT, env=env)
    tail = (out.stdout + out.stderr)[-3000:]
    assert "SANITIZE_SMOKE_DONE" in out.stdout, tail
    assert out.returncode == 0 and "ERROR SUMMARY: 0 errors" in (out.stdout + out.stderr), tail
