#!/bin/bash
# GPU session B (round 2): whole gpu suite without -x, per-layer error growth vs the fp16 noise floor, VAE timing.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/b_pytest.log | tail -30
timeout 600 python tools/gpu_tap_diff.py full 1.0 981 > gpurun_out/b_tapdiff_full_gate1.txt 2>&1; echo "tapdiff rc=$?"
tail -3 gpurun_out/b_tapdiff_full_gate1.txt
timeout 600 python tools/time_vae.py > gpurun_out/b_time_vae.txt 2>&1; echo "vae rc=$?"
cat gpurun_out/b_time_vae.txt | tail -5
