#!/bin/bash
# LayerNorm grid-stride variant A/B + full gpu suite + bench on the final build
mkdir -p gpurun_out
{
echo "== LTT_LN_ROWS=0 (one row per warp)"; LTT_LN_ROWS=0 python tools/bench_ops.py norms 2>&1 | grep layernorm
echo "== LTT_LN_ROWS=1 (grid-stride rows, gamma/beta in registers)"; python tools/bench_ops.py norms 2>&1 | grep layernorm
} > gpurun_out/ln_ab.txt 2>&1
python -m pytest tests -m gpu -q -x > gpurun_out/gputests.txt 2>&1; tail -3 gpurun_out/gputests.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_B1.json 2> gpurun_out/bench_B1.err
LTT_LN_ROWS=0 python bench.py --steps 3 --warmup 3 --batch 8 --no-cpu-baseline > gpurun_out/bench_B8_lnrows0.json 2>> gpurun_out/bench_B1.err
python bench.py --steps 3 --warmup 3 --batch 8 --no-cpu-baseline > gpurun_out/bench_B8.json 2>> gpurun_out/bench_B1.err
cat gpurun_out/ln_ab.txt; cat gpurun_out/bench_B1.json gpurun_out/bench_B8_lnrows0.json gpurun_out/bench_B8.json | cut -c1-400
