#!/bin/bash
mkdir -p gpurun_out
{
for cfg in "LTT_ATTN_PTM=1" "LTT_ATTN_PTM=2"; do
  echo "== $cfg"
  env $cfg timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "attention" 2>&1 | tail -2
  env $cfg timeout 300 python tools/bench_ops.py attn 2>&1 | grep -E "attention"
done
} > gpurun_out/j_attn.txt 2>&1
cat gpurun_out/j_attn.txt
