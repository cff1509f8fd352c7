#!/bin/bash
# GPU session A (round 2): full gpu test suite, parity tables vs the unmodified reference, bench lines for configs 2-5.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -15 gpurun_out/a_pytest.log
timeout 900 python tools/gpu_parity_steps.py > gpurun_out/a_parity.log 2>&1; echo "parity rc=$?"
tail -5 gpurun_out/a_parity.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/a_bench_B1.json 2> gpurun_out/a_bench_B1.err; echo "bench rc=$?"
cat gpurun_out/a_bench_B1.json | cut -c1-600
for bx in 1 6 30; do
  timeout 600 python bench.py --steps 2 --warmup 3 --batch 8 --boxes $bx --no-cpu-baseline > gpurun_out/a_bench_B8_boxes$bx.json 2> gpurun_out/a_bench_B8_boxes$bx.err
  cut -c1-300 gpurun_out/a_bench_B8_boxes$bx.json
done
timeout 900 python bench.py --steps 2 --warmup 3 --size 768 --batch 4 --no-cpu-baseline > gpurun_out/a_bench_768_B4.json 2> gpurun_out/a_bench_768_B4.err
cut -c1-300 gpurun_out/a_bench_768_B4.json
timeout 900 python bench.py --steps 1 --warmup 3 --batch 64 --no-cpu-baseline > gpurun_out/a_bench_B64.json 2> gpurun_out/a_bench_B64.err
cut -c1-300 gpurun_out/a_bench_B64.json
