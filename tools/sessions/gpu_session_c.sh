#!/bin/bash
# GPU session C (round 2): LayerNorm fold on/off A/B, full gpu suite, parity tables (true-fp32 references), bench with decode.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/c_pytest.log | tail -30
timeout 600 python bench.py --steps 5 --warmup 3 --no-decode --no-cpu-baseline > gpurun_out/c_bench_B1_fold.json 2> gpurun_out/c_bench_B1_fold.err; echo "bench fold rc=$?"
LTT_NO_LNFOLD=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-decode --no-cpu-baseline > gpurun_out/c_bench_B1_nofold.json 2> gpurun_out/c_bench_B1_nofold.err; echo "bench nofold rc=$?"
timeout 600 python bench.py --steps 2 --warmup 3 --batch 8 --no-decode --no-cpu-baseline > gpurun_out/c_bench_B8_fold.json 2> gpurun_out/c_bench_B8_fold.err
LTT_NO_LNFOLD=1 timeout 600 python bench.py --steps 2 --warmup 3 --batch 8 --no-decode --no-cpu-baseline > gpurun_out/c_bench_B8_nofold.json 2> gpurun_out/c_bench_B8_nofold.err
for f in c_bench_B1_fold c_bench_B1_nofold c_bench_B8_fold c_bench_B8_nofold; do python - <<PY
import json
d=json.load(open("gpurun_out/$f.json")); r=d["roofline"]
print("$f", round(d["value"],3), "img/s  ms", round(d["ms_per_step"],1), "launches/step", d["gpu_launches"]//d["steps"], {k:v["ms"] for k,v in r["classes"].items()})
PY
done
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/c_bench_B1_full.json 2> gpurun_out/c_bench_B1_full.err; echo "bench full rc=$?"
cut -c1-400 gpurun_out/c_bench_B1_full.json
timeout 900 python tools/gpu_parity_steps.py > gpurun_out/c_parity.log 2>&1; echo "parity rc=$?"
timeout 600 python tools/gpu_tap_diff.py full 1.0 981 > gpurun_out/c_tapdiff_full_gate1.txt 2>&1; echo "tapdiff rc=$?"
tail -2 gpurun_out/c_tapdiff_full_gate1.txt
