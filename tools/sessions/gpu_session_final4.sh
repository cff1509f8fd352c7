#!/bin/bash
# Closing session: gpu suite, smoke, headline bench on the final build; parity tables of the new rows.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/w_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/w_pytest.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/w_pytest.log | tail -8
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/w_smoke.log 2>&1; tail -1 gpurun_out/w_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/w_bench_B1.json 2> gpurun_out/w_bench_B1.err; echo "bench rc=$?"
cut -c1-260 gpurun_out/w_bench_B1.json
python tests/reward_checks.py 2>&1 | grep -v Warning > gpurun_out/w_reward_checks.txt; python tests/clip_checks.py 2>&1 | grep -v Warning > gpurun_out/w_clip_checks.txt
grep -c "^ok" gpurun_out/w_reward_checks.txt gpurun_out/w_clip_checks.txt; grep -h "FAIL\|EXC" gpurun_out/w_reward_checks.txt gpurun_out/w_clip_checks.txt
