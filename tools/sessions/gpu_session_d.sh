#!/bin/bash
mkdir -p gpurun_out
./tools/ubench/tmem_mufu > gpurun_out/d_ubench_tmem_mufu.txt 2>&1; cat gpurun_out/d_ubench_tmem_mufu.txt
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "LayerNorm" 2>&1 | tail -5
ab() { # name, env, args
  env $2 timeout 600 python bench.py --steps $4 --warmup 3 --no-decode --no-cpu-baseline $3 > gpurun_out/d_$1.json 2> gpurun_out/d_$1.err
  python - <<PY
import json
d=json.load(open("gpurun_out/d_$1.json")); r=d["roofline"]
print("$1", round(d["value"],3), "img/s  ms", round(d["ms_per_step"],1), "launches/step", d["gpu_launches"]//d["steps"], {k:v["ms"] for k,v in r["classes"].items()})
PY
}
ab B1_fold LTT_X=1 "" 5
ab B1_nofold LTT_NO_LNFOLD=1 "" 5
ab B1_fold2 LTT_X=1 "" 5
ab B1_nofold2 LTT_NO_LNFOLD=1 "" 5
ab B8_fold LTT_X=1 "--batch 8" 2
ab B8_nofold LTT_NO_LNFOLD=1 "--batch 8" 2
