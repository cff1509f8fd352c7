#!/bin/bash
# Final measurement session of round 2 (after the conditioning-prep / reward rows): gpu suite, smoke, headline bench +
# reference arm, configs 3/4/5 at N=1, CLIP / reward timing.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/y_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/y_pytest.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/y_pytest.log | tail -8
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/y_smoke.log 2>&1; tail -1 gpurun_out/y_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/y_bench_B1.json 2> gpurun_out/y_bench_B1.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/y_bench_B1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/y_bench_reference.json 2> gpurun_out/y_bench_reference.err; cut -c1-300 gpurun_out/y_bench_reference.json
timeout 600 python bench.py --steps 2 --warmup 3 --batch 8 --boxes 6 --no-cpu-baseline > gpurun_out/y_bench_B8_boxes6.json 2> gpurun_out/y_bench_B8.err
cut -c1-200 gpurun_out/y_bench_B8_boxes6.json
timeout 900 python bench.py --steps 2 --warmup 3 --size 768 --batch 4 --no-cpu-baseline > gpurun_out/y_bench_768_B4.json 2> gpurun_out/y_bench_768_B4.err
cut -c1-200 gpurun_out/y_bench_768_B4.json
timeout 900 python bench.py --steps 1 --warmup 3 --batch 64 --no-cpu-baseline > gpurun_out/y_bench_B64.json 2> gpurun_out/y_bench_B64.err
cut -c1-200 gpurun_out/y_bench_B64.json
python tools/time_reward.py 8 2>&1 | grep -v Warning > gpurun_out/y_reward_timing.txt; python tools/time_reward.py 1 2>&1 | grep -v Warning >> gpurun_out/y_reward_timing.txt; cat gpurun_out/y_reward_timing.txt
python tools/time_clip.py 6 3 2>&1 | grep -v "Warning\|detach\|print(" > gpurun_out/y_clip_timing.txt; python tools/time_clip.py 30 5 2>&1 | grep -v "Warning\|detach\|print(" >> gpurun_out/y_clip_timing.txt; cat gpurun_out/y_clip_timing.txt
python tests/reward_checks.py 2>&1 | grep -v Warning > gpurun_out/y_reward_checks.txt; python tests/clip_checks.py 2>&1 | grep -v Warning > gpurun_out/y_clip_checks.txt; grep -c "^ok" gpurun_out/y_reward_checks.txt gpurun_out/y_clip_checks.txt; grep -h "FAIL\|EXC" gpurun_out/y_reward_checks.txt gpurun_out/y_clip_checks.txt
