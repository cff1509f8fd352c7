#!/bin/bash
mkdir -p gpurun_out
./tools/ubench/tmem_mufu > gpurun_out/e_ubench_tmem_mufu.txt 2>&1; cat gpurun_out/e_ubench_tmem_mufu.txt
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e_pytest.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/e_pytest.log | tail -12
ab() { # name, env, args, steps
  env $2 timeout 600 python bench.py --steps $4 --warmup 3 --no-decode --no-cpu-baseline $3 > gpurun_out/e_$1.json 2> gpurun_out/e_$1.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/e_$1.json")); r=d["roofline"]
    print("$1", round(d["value"],3), "img/s  ms", round(d["ms_per_step"],1), "launches/step", d["gpu_launches"]//d["steps"], {k:v["ms"] for k,v in r["classes"].items()})
except Exception as ex:
    print("$1 failed", ex); print(open("gpurun_out/e_$1.err").read()[-1500:])
PY
}
ab B1_tune LTT_X=1 "" 5
ab B1_notune LTT_NO_AUTOTUNE=1 "" 5
ab B1_tune2 LTT_X=1 "" 5
ab B8_tune LTT_X=1 "--batch 8" 2
ab B8_notune LTT_NO_AUTOTUNE=1 "--batch 8" 2
LTT_VERBOSE=1 timeout 600 python bench.py --steps 1 --warmup 3 --no-decode --no-cpu-baseline 2>&1 | grep "tuned" > gpurun_out/e_tuned_B1.txt; wc -l gpurun_out/e_tuned_B1.txt
