#!/bin/bash
mkdir -p gpurun_out
for p in 0 16 12 8 6; do
  echo "== LTT_ATTN_POLY=$p"
  LTT_ATTN_POLY=$p timeout 300 python tools/bench_ops.py attn 2>&1 | grep -E "d= 40 nq= 4096 nk= 41|B=16"
done > gpurun_out/h_attn_poly.txt 2>&1
cat gpurun_out/h_attn_poly.txt
for v in 8 9; do
  echo "== LTT_ATTN40=$v"
  LTT_ATTN40=$v timeout 300 python tools/bench_ops.py attn 2>&1 | grep -E "d= 40 nq= 4096 nk= 41|B=16"
  LTT_ATTN40=$v timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "attention and 40" 2>&1 | tail -1
done >> gpurun_out/h_attn_poly.txt 2>&1
tail -8 gpurun_out/h_attn_poly.txt
