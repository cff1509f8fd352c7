#!/bin/bash
mkdir -p gpurun_out
{
for lib in layoutllm_t2i_b200/libltt_b200.so layoutllm_t2i_b200/libltt_lite.so; do
  echo "== LTT_LIB=$lib"
  LTT_LIB=$lib timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "linear or conv3x3 or qkv" 2>&1 | tail -2
  LTT_LIB=$lib timeout 300 python tools/bench_ops.py linear 2>&1 | grep -E "linear"
  LTT_LIB=$lib timeout 300 python tools/bench_ops.py conv 2>&1 | grep -E "conv"
  LTT_LIB=$lib timeout 600 python bench.py --steps 4 --warmup 3 --no-decode --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('bench B1', round(d['value'],3), 'img/s ms', round(d['ms_per_step'],1), {k:v['ms'] for k,v in r['classes'].items()})"
done
} > gpurun_out/lite_ab.txt 2>&1
cat gpurun_out/lite_ab.txt
