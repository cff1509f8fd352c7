#!/bin/bash
mkdir -p gpurun_out
for p in 0 8 4 3 2; do
  echo "== LTT_ATTN_POLY=$p"
  LTT_ATTN_POLY=$p timeout 300 python tools/bench_ops.py attn 2>&1 | grep -E "d= 40 nq= 4096 nk= 41|B=16"
  LTT_ATTN_POLY=$p timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "attention and 40" 2>&1 | tail -1
done > gpurun_out/g_attn_poly.txt 2>&1
cat gpurun_out/g_attn_poly.txt
