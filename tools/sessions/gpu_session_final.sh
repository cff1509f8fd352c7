#!/bin/bash
# Final measurement session of round 2: gpu suite, parity tables, headline bench + configs 3/4/5 at N=1.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/z_pytest.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/z_pytest.log | tail -8
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/z_smoke.log 2>&1; tail -1 gpurun_out/z_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/z_bench_B1.json 2> gpurun_out/z_bench_B1.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/z_bench_B1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/z_bench_reference.json 2> gpurun_out/z_bench_reference.err; cut -c1-300 gpurun_out/z_bench_reference.json
for bx in 1 6 30; do
  timeout 600 python bench.py --steps 2 --warmup 3 --batch 8 --boxes $bx --no-cpu-baseline > gpurun_out/z_bench_B8_boxes$bx.json 2> gpurun_out/z_bench_B8_boxes$bx.err
  cut -c1-200 gpurun_out/z_bench_B8_boxes$bx.json
done
timeout 900 python bench.py --steps 2 --warmup 3 --size 768 --batch 4 --no-cpu-baseline > gpurun_out/z_bench_768_B4.json 2> gpurun_out/z_bench_768_B4.err
cut -c1-200 gpurun_out/z_bench_768_B4.json
timeout 900 python bench.py --steps 1 --warmup 3 --batch 64 --no-cpu-baseline > gpurun_out/z_bench_B64.json 2> gpurun_out/z_bench_B64.err
cut -c1-200 gpurun_out/z_bench_B64.json
timeout 900 python tools/gpu_parity_steps.py > gpurun_out/z_parity.log 2>&1; echo "parity rc=$?"
timeout 600 python tools/gpu_tap_diff.py full 1.0 981 > gpurun_out/z_tapdiff_full_gate1.txt 2>&1; tail -1 gpurun_out/z_tapdiff_full_gate1.txt
timeout 600 python tools/time_vae.py > gpurun_out/z_time_vae.txt 2>&1; cat gpurun_out/z_time_vae.txt | tail -3
