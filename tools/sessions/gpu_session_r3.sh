#!/bin/bash
# relation block norm3 statistics from the producer GEMM epilogue (LTT_LNFOLD=r3) vs the separate statistics pass
mkdir -p gpurun_out
python -m pytest tests/test_model_gpu.py -q -x -k "norm3 or LayerNorm" 2>&1 | tail -2
for i in 1 2; do
  for mode in off r3; do
    if [ $mode = off ]; then unset LTT_LNFOLD; else export LTT_LNFOLD=r3; fi
    python bench.py --steps 4 --warmup 3 --no-decode --no-cpu-baseline 2>gpurun_out/r3_err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['roofline']['classes']
print('$mode', round(d['ms_per_step'],2), d['gpu_launches'], {k: round(v['ms'],1) for k,v in c.items()})"
  done
done | tee gpurun_out/r3_ab.txt
unset LTT_LNFOLD
