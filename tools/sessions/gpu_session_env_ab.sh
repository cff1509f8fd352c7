#!/bin/bash
# same-box A/B of one environment switch: usage gpu_session_env_ab.sh VAR VALUE_A VALUE_B  (sampler-only bench lines, interleaved)
mkdir -p gpurun_out
for i in 1 2; do
  for v in "$2" "$3"; do
    env "$1=$v" python bench.py --steps 4 --warmup 3 --no-decode --no-cpu-baseline 2>gpurun_out/env_ab_err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['roofline']['classes']
print('$1=$v', round(d['ms_per_step'],2), {k: round(x['ms'],1) for k,x in c.items()})"
  done
done | tee gpurun_out/env_ab.txt
