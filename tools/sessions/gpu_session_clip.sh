#!/bin/bash
# CLIP text tower (row f3): parity checks, timing vs the reference call pattern; LayerNorm grid-stride A/B
mkdir -p gpurun_out
python tests/clip_checks.py > gpurun_out/clip_checks.txt 2>&1; grep -v Warning gpurun_out/clip_checks.txt | tail -30
python tools/time_clip.py 6 3 > gpurun_out/clip_timing.txt 2>&1; python tools/time_clip.py 30 5 >> gpurun_out/clip_timing.txt 2>&1; grep -v "Warning\|detach\|print(" gpurun_out/clip_timing.txt | tail -12
{
echo "== LTT_LN_ROWS=0 (one row per warp)"; LTT_LN_ROWS=0 python tools/bench_ops.py norms 2>&1 | grep layernorm
echo "== LTT_LN_ROWS=1 (>= 16384 rows, C <= 512: grid-stride rows, gamma/beta in registers, raw-vector prefetch ring)"; python tools/bench_ops.py norms 2>&1 | grep layernorm
} > gpurun_out/ln_ab2.txt 2>&1
python -m pytest tests/test_ops_gpu.py -q -x -k "layernorm or attention" 2>&1 | tail -3
