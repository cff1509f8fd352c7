#!/bin/bash
mkdir -p gpurun_out
for v in 0 1; do
  echo "== LTT_ATTN_PTM=$v"
  LTT_ATTN_PTM=$v timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "attention and 40" 2>&1 | tail -3
  LTT_ATTN_PTM=$v timeout 300 python tools/bench_ops.py attn 2>&1 | grep -E "d= 40 nq= 4096 nk= 41|B=16"
done > gpurun_out/i_attn_ptm.txt 2>&1
cat gpurun_out/i_attn_ptm.txt
