#!/bin/bash
# same-box A/B of two library builds (LTT_LIB): sampler-only bench lines, interleaved
mkdir -p gpurun_out
for i in 1 2; do
  for lib in old b200; do
    LTT_LIB_PARTIAL=1 LTT_LIB=$PWD/layoutllm_t2i_b200/libltt_$lib.so python bench.py --steps 4 --warmup 3 --no-decode --no-cpu-baseline 2>gpurun_out/ab_err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['roofline']['classes']
print('$lib', round(d['ms_per_step'],2), {k: round(v['ms'],1) for k,v in c.items()})"
  done
done | tee gpurun_out/ab_lib.txt
